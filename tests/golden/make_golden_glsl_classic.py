"""Generates tests/golden/classic_frames.npz: every scene of tests/classic_scenes.py rendered BY THE REFERENCE'S OWN SHADERS of
the classic ray-caster (oracle/classicref.py executes GLRaycaster-*-FS.glsl, Compose-FS.glsl, Compose-CV-FS.glsl ...).  Run in
the build container (needs /root/reference); the GPU box, which has no reference, checks the CUDA kernel against these
frames (tests/test_gpu_classic_golden.py).

    python tests/golden/make_golden_glsl_classic.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import classicref as R          # noqa: E402
from tests import classic_scenes as S       # noqa: E402


def main():
    assert R.available(), "/root/reference is not mounted"
    out = {}
    for name in S.NAMES:
        p, bricks, store, tf, cv = S.build(name)
        if cv is None:
            img = R.render(p, bricks, store, tf)
        else:
            img, hits = R.render(p, bricks, store, None, cv)
            out[name + "/hit_mask"] = (hits[0][..., 3] != 0)
            out[name + "/hit2_mask"] = (hits[2][..., 3] != 0)
        out[name + "/image"] = img.astype(np.float32)
        print("%-14s %d bricks, %d lit pixels" % (name, len(bricks), int((img[..., 3] > 0).sum())))
    path = os.path.join(ROOT, "tests", "golden", "classic_frames.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
