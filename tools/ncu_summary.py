"""Reads ncu reports/launch lists brought back in gpurun_out/ (needs only the ncu CLI, no GPU) and prints
the handful of metrics profiles/README.md quotes."""
import collections
import csv
import subprocess
import sys

WANT = ["Kernel Name", "launch__grid_size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__registers_per_thread", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum"]


def raw(path):
    out = subprocess.check_output(["ncu", "-i", path, "--page", "raw", "--csv"], stderr=subprocess.DEVNULL).decode()
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    res = []
    for vals in rows[2:]:
        res.append({h: (v, u) for h, v, u in zip(hdr, vals, units)})
    return res


def launches(path):
    rows = list(csv.reader(open(path)))
    for i, r in enumerate(rows):
        if r and r[0] == "ID":
            hdr, start = r, i + 1
            break
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    tot, cnt = collections.defaultdict(float), collections.Counter()
    for r in rows[start:]:
        if len(r) <= vi:
            continue
        name = r[ki].split("(")[0].replace("<unnamed>::", "").replace("void ", "")
        tot[name] += float(r[vi].replace(",", ""))
        cnt[name] += 1
    s = sum(tot.values())
    for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
        print("%-36s n=%4d  share %.3f  avg %.1f us" % (k[:36], cnt[k], v / s, v / cnt[k] / 1e3))


if __name__ == "__main__":
    for p in sys.argv[1:]:
        print("==", p)
        if p.endswith(".csv"):
            launches(p)
        else:
            for k in raw(p):
                for w in WANT:
                    if w in k:
                        print("  %-62s %s %s" % (w, k[w][0][:60], k[w][1]))
