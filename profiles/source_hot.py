#!/usr/bin/env python
"""Per-source-line view of an `ncu --set full --import-source on` report (needs -lineinfo at compile time):
warp instructions executed, thread instructions, average active threads and stall samples per source line of the
kernel, sorted by executed instructions.

  python profiles/source_hot.py gpurun_out/x.ncu-rep [top_n] > profiles/rNN_x_source.txt"""
import csv
import io
import subprocess
import sys


def main():
    path = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = None
    lines = []
    for r in rows:
        if r and r[0] == "Line No":
            hdr = r
            continue
        if hdr and len(r) == len(hdr) and r[0].isdigit():
            lines.append(r)
    ie, te, sm = hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed"), hdr.index("# Samples")
    lsb = hdr.index("stall_long_sb")
    tot_i = sum(int(r[ie]) for r in lines)
    tot_s = sum(int(r[sm]) for r in lines)
    print("# %s: %d warp instructions, %d stall samples over %d source lines" % (path, tot_i, tot_s, len(lines)))
    print("%5s %12s %6s %7s %8s %8s  %s" % ("line", "warp_inst", "share", "thr/ins", "samples", "long_sb", "source"))
    for r in sorted(lines, key=lambda r: -int(r[ie]))[:top]:
        i = int(r[ie])
        print("%5s %12d %5.1f%% %7.1f %8s %8s  %s" % (r[0], i, 100.0 * i / max(tot_i, 1), int(r[te]) / max(i, 1), r[sm], r[lsb], r[1].strip()[:110]))


if __name__ == "__main__":
    main()
