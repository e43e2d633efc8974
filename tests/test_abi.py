"""CPU: the C-ABI library loads and exports every symbol include/*.h declares (no compute calls)."""
import ctypes
import os
import re

from imagevis3d_b200 import _lib as L

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared(header):
    src = open(os.path.join(ROOT, "include", header)).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(tvk[r]?_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = ctypes.CDLL(L.LIB_PATH)          # raises if the library was not built
    for header in sorted(os.listdir(os.path.join(ROOT, "include"))):
        names = _declared(header)
        assert names, header
        for n in names:
            assert hasattr(lib, n), "%s declared in %s but not exported" % (n, header)


def test_python_binding_lists_the_same_symbols():
    assert sorted(L.EXPORTS) == _declared("tvk_cuda.h")


def test_struct_layouts_match_header_sizes():
    # sizes the C compiler gives the PODs (guards against ctypes drift)
    import subprocess, tempfile
    code = r'''
#include <stdio.h>
#include "tvk_cuda.h"
int main(void){printf("%zu %zu %zu %zu %zu %zu\n", sizeof(tvk_pool_desc), sizeof(tvk_visibility),
  sizeof(tvk_raycast_params), sizeof(tvk_stats), sizeof(tvk_build_desc), sizeof(tvk_build_info));return 0;}'''
    with tempfile.TemporaryDirectory() as td:
        c = os.path.join(td, "s.c")
        open(c, "w").write(code)
        exe = os.path.join(td, "s")
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), c, "-o", exe])
        sizes = [int(v) for v in subprocess.check_output([exe]).split()]
    got = [ctypes.sizeof(t) for t in (L.PoolDesc, L.Visibility, L.RaycastParams, L.Stats, L.BuildDesc, L.BuildInfo)]
    assert got == sizes


def test_build_plan_is_host_only_and_matches_oracle():
    from oracle import oracle as O
    from imagevis3d_b200.context import TvkContext
    for vol in [(512, 512, 512), (2048, 2048, 2048), (8192, 8192, 4096), (33, 17, 9), (8, 8, 1)]:
        info = TvkContext.build_plan(vol, 64, 2, 1)
        p = O.plan(vol, 64, 2)
        assert info.lod_count == p.lod_count and info.total_bricks == p.total_bricks
        for l in range(p.lod_count):
            assert list(info.lod_size[l]) == list(p.lod_size[l])
            assert list(info.lod_bricks[l]) == list(p.lod_bricks[l])
            assert info.lod_offset[l] == p.lod_offset[l]
    # SURVEY 8a sizes (bricks down to the first single-brick LOD, i.e. what the pool addresses):
    # C2 890 bricks, C3 49,597, C5 1,487,165
    def pool_total(vol):
        info = TvkContext.build_plan(vol)
        for l in range(info.lod_count):
            if info.lod_bricks[l][0] * info.lod_bricks[l][1] * info.lod_bricks[l][2] == 1:
                return info.lod_offset[l] + 1
    assert pool_total((512,) * 3) == 890
    assert pool_total((2048,) * 3) == 49597
    assert pool_total((8192, 8192, 4096)) == 1487165
