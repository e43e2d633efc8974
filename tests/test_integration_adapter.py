"""CPU: the reference-side adapter (integration/CudaGridLeaperRenderer.{h,cpp}, the AbstrRenderer subclass a Tuvok
maintainer adds) compiles against the REFERENCE's own headers -- every override carries `override`, so a signature
that differs from Tuvok/Renderer/AbstrRenderer.h is a compile error.  Needs /root/reference (skipped on the GPU box)."""
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/Tuvok"
SHIM = os.path.join(ROOT, "oracle", "ref_build", "shim")

needs_ref = pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not mounted")


def _flags():
    return ["-fsyntax-only", "-std=c++14", "-fpermissive", "-w", "-DTUVOK_NO_QT=1", "-D_FILE_OFFSET_BITS=64", "-DNDEBUG",
            "-I" + SHIM, "-include", os.path.join(SHIM, "Controller", "Controller.h"), "-include", os.path.join(SHIM, "boost_ns.h"),
            "-I" + REF, "-I" + REF + "/Basics", "-I" + REF + "/IO", "-I" + REF + "/IO/3rdParty/boost", "-I" + REF + "/Basics/3rdParty",
            "-I" + REF + "/IO/UVF", "-I" + REF + "/IO/exception", "-I" + REF + "/3rdParty/LUA",
            "-I" + os.path.join(ROOT, "include"), "-I" + os.path.join(ROOT, "imagevis3d_b200", "csrc", "host")]


@needs_ref
def test_adapter_compiles_against_the_reference_headers():
    r = subprocess.run(["g++"] + _flags() + [os.path.join(ROOT, "integration", "CudaGridLeaperRenderer.cpp")],
                       capture_output=True, text=True)
    assert r.returncode == 0 and "error" not in r.stderr, r.stderr[-4000:]


@needs_ref
def test_adapter_overrides_every_pure_virtual_of_abstrrenderer():
    hdr = open(os.path.join(REF, "Renderer", "AbstrRenderer.h")).read()
    pure = re.findall(r"virtual\s+[^;{]*?\b(\w+)\s*\([^;{]*\)\s*(?:const\s*)?=\s*0\s*;", hdr)
    assert len(pure) == 11, pure                       # eleven pure virtuals
    mine = open(os.path.join(ROOT, "integration", "CudaGridLeaperRenderer.h")).read()
    for name in pure:
        assert re.search(r"\b%s\s*\([^;]*\)\s*(?:const\s*)?override" % name, mine), name


@needs_ref
def test_adapter_is_instantiable():
    """no pure virtual is left over: `new CudaGridLeaperRenderer` must compile (abstract classes cannot be created)"""
    src = '#include "%s"\ntuvok::AbstrRenderer* make() { return new tuvok::CudaGridLeaperRenderer(nullptr, false, false, false); }\n' % (
        os.path.join(ROOT, "integration", "CudaGridLeaperRenderer.h"))
    r = subprocess.run(["g++"] + _flags() + ["-x", "c++", "-"], input=src, capture_output=True, text=True)
    assert r.returncode == 0 and "error" not in r.stderr, r.stderr[-4000:]
