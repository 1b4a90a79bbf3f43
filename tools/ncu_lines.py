#!/usr/bin/env python
"""Per-source-line totals from `ncu -i X.ncu-rep --page source --csv --print-source cuda,sass --kernel-name regex:K`:
instructions executed and stall samples by file:line, top N."""
import csv
import sys


def main(path, top=30):
    rows = list(csv.reader(open(path, errors="replace")))
    cur_file = ""
    out = []
    hdr = None
    for r in rows:
        if len(r) >= 2 and r[0] == "File Path":
            cur_file = r[1].rsplit("/", 1)[-1]
            continue
        if r and r[0] == "Line No":
            hdr = r
            continue
        if hdr is None or not r or not r[0].strip().isdigit():
            continue
        try:
            i_inst = hdr.index("Instructions Executed")
            i_samp = hdr.index("# Samples")
            i_thr = hdr.index("Thread Instructions Executed")
            out.append((cur_file, int(r[0]), r[1].strip()[:90], int(r[i_inst] or 0), int(r[i_samp] or 0), int(r[i_thr] or 0)))
        except (ValueError, IndexError):
            pass
    tot_i = sum(o[3] for o in out) or 1
    tot_s = sum(o[4] for o in out) or 1
    print(f"total warp-instructions {tot_i}, samples {tot_s}")
    for o in sorted(out, key=lambda o: -o[4])[:top]:
        print(f"{o[0]}:{o[1]:4d} inst {100 * o[3] / tot_i:5.1f}%  samples {100 * o[4] / tot_s:5.1f}%  thr/inst {o[5] / max(1, o[3]):4.1f} | {o[2]}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 30)
