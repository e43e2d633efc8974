"""Generates tests/golden/glsl_frames.npz: every scene of tests/glsl_scenes.py rendered to convergence BY THE REFERENCE'S
OWN SHADERS (oracle/glslref.py executes Tuvok/Shaders/*.glsl and the GLSL GLVolumePool / GLHashTable generate), with the
request table of every sub-frame.  Run in the build container (needs /root/reference); the GPU box, which has no
reference, checks the CUDA path against these frames (tests/test_gpu_glsl_golden.py).

    python tests/golden/make_golden_glsl.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import glslref as G          # noqa: E402
from tests import glsl_scenes as S       # noqa: E402


def main():
    assert G.available(), "/root/reference is not mounted"
    out = {}
    for name in sorted(S.SCENES):
        r = S.render_progressive(G.raycast, S.build(name))
        out[name + "/image"] = r["image"]
        out[name + "/subframes"] = np.int64(r["subframes"])
        for k, req in enumerate(r["requests"]):
            out["%s/req%d" % (name, k)] = req.astype(np.uint16)
        print("%-18s %d sub-frames, %d requests, %d lit pixels" % (name, r["subframes"], sum(len(q) for q in r["requests"]),
                                                                    int((r["image"][..., 3] > 0).sum())))
    path = os.path.join(ROOT, "tests", "golden", "glsl_frames.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
