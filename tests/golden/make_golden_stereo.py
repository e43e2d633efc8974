"""Writes tests/golden/stereo_matrices.json: FLOATMATRIX4::BuildStereoLookAtAndProjection of the reference
(oracle/_ref/stereo_ref, built from /root/reference/Tuvok/Basics/Vectors.h by oracle/ref_build/Makefile) for a few cameras.
Run here (needs /root/reference); the GPU box reads the committed file."""
import json
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
BIN = os.path.join(HERE, "..", "..", "oracle", "_ref", "stereo_ref")

CASES = [
    # eye, at, up, fovy, aspect, znear, zfar, focal length, eye distance
    [0, 0, 1.6, 0, 0, 0, 0, 1, 0, 50, 1920 / 1080, 0.01, 1000, 1.0, 0.02],          # the reference's defaults
    [0.3, -0.2, 1.9, 0.1, 0, -0.1, 0, 1, 0, 50, 203 / 149, 0.01, 1000, 1.0, 0.02],
    [1.0, 0.5, 1.2, 0, 0, 0, 0.1, 0.9, 0.2, 35, 1.0, 0.05, 100, 2.5, 0.06],
    [0, 0, 2.5, 0, 0.1, 0, 0, 1, 0, 70, 0.75, 0.1, 50, 0.8, 0.1],
]


def main():
    out = []
    for c in CASES:
        args = [repr(float(v)) for v in c]
        txt = subprocess.check_output([BIN] + args).decode()
        rows = [[float(t) for t in line.split()] for line in txt.strip().splitlines()]
        assert len(rows) == 4 and all(len(r) == 16 for r in rows)
        out.append({"args": [float(v) for v in c], "view_left": rows[0], "view_right": rows[1], "proj_left": rows[2], "proj_right": rows[3]})
    with open(os.path.join(HERE, "stereo_matrices.json"), "w") as f:
        json.dump(out, f, indent=1)
    print("wrote %d cases" % len(out))


if __name__ == "__main__":
    main()
