"""Summaries of ncu outputs: per-kernel shares of a launch list (csv) and key metrics of a full capture (.ncu-rep)."""
import collections
import csv
import subprocess
import sys


def launches(path):
    rows = [l for l in open(path) if l.startswith('"')]
    r = list(csv.reader(rows))
    hdr = r[0]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = collections.defaultdict(lambda: [0, 0.0, 0.0])
    for x in r[1:]:
        try:
            v = float(x[vi].replace(",", ""))
        except ValueError:
            continue
        a = agg[x[ki][:64]]
        a[0] += 1
        a[1] += v
        a[2] = max(a[2], v)
    tot = sum(v[1] for v in agg.values())
    out = []
    for k, (c, v, mx) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.append(f"{k:66s} n={c:4d} total={v / 1e6:9.3f} ms  max={mx / 1e6:8.3f} ms  share={100 * v / tot:5.1f}%")
    return "\n".join(out)


WANT = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "lts__t_bytes.sum",
        "l1tex__t_bytes.sum", "launch__registers_per_thread", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "launch__grid_size", "launch__block_size",
        "sm__inst_executed_pipe_tensor.sum", "lts__t_sectors_srcunit_tex_op_read.sum"]


def full(path):
    txt = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    r = list(csv.reader(txt.splitlines()))
    hdr, units = r[0], r[1]
    out = []
    for row in r[2:]:
        d = {}
        for w in WANT:
            if w in hdr:
                i = hdr.index(w)
                d[w] = f"{row[i]} {units[i]}".strip()
        out.append(d)
    return out


if __name__ == "__main__":
    p = sys.argv[1]
    if p.endswith(".csv"):
        print(launches(p))
    else:
        for d in full(p):
            print("----")
            for k, v in d.items():
                print(f"  {k}: {v}")
