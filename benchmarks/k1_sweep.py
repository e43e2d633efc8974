#!/usr/bin/env python
"""Builds variants of the K1 ray-march kernel (tuning macros of csrc/tvk_raycast.cu) into
build_variants/*.so here (no GPU needed), and -- with --run, on the GPU box -- benches each one:

  python benchmarks/k1_sweep.py build  NAME:-DTVK_NB_TEX=8,-DTVK_OCC_TEX=6 ...
  python benchmarks/k1_sweep.py run    [sampler ...]          # every build_variants/*.so x sampler
"""
import glob
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
VAR = os.path.join(ROOT, "build_variants")


def build(specs):
    from imagevis3d_b200 import build as B
    B.build()
    os.makedirs(VAR, exist_ok=True)
    objdir = os.path.join(ROOT, "imagevis3d_b200", "build")
    others = [os.path.join(objdir, os.path.basename(s) + ".o") for s in B.sources() if not s.endswith("tvk_raycast.cu")]
    procs = []
    for spec in specs:
        name, _, flags = spec.partition(":")
        obj = os.path.join(VAR, name + ".o")
        cmd = [B._nvcc()] + B.NVCC_FLAGS + [f for f in flags.split(",") if f] + ["-c", os.path.join(B.CSRC, "tvk_raycast.cu"), "-o", obj]
        procs.append((name, obj, subprocess.Popen(cmd)))
    for name, obj, p in procs:
        if p.wait():
            raise SystemExit("nvcc failed for " + name)
        subprocess.check_call([B._nvcc(), "-shared", "-o", os.path.join(VAR, "libtvk_%s.so" % name), obj] + others +
                              ["-gencode", "arch=compute_100a,code=sm_100a", "-lcudart", "-lz", "-ldl"])
        os.remove(obj)
        print("built", name)


def run(samplers):
    for lib in sorted(glob.glob(os.path.join(VAR, "libtvk_*.so"))):
        for s in samplers or ["q4", "tex3d"]:
            env = dict(os.environ, TVK_CUDA_LIB=lib, TVK_SAMPLER=s)
            out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--no-cpu", "--steps", "20"], env=env,
                                 capture_output=True, text=True)
            try:
                d = json.loads(out.stdout.strip().splitlines()[-1])
                print("%-28s %-6s %8.4f ms  %7.1f Gsamples/s" % (os.path.basename(lib), s, d["ms_per_step"], d["value"] / 1e9), flush=True)
            except Exception:
                print(os.path.basename(lib), s, "FAILED", out.stderr[-400:], flush=True)


if __name__ == "__main__":
    (build if sys.argv[1] == "build" else run)(sys.argv[2:])
