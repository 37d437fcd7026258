"""Turn ncu output into the committed summaries under profiles/.

    python tools/summarize_ncu.py launches <launches.csv> <out.md> "<command line that was profiled>"
    python tools/summarize_ncu.py raw <report.ncu-rep> <out.json>      (per-kernel key metrics of a --set full capture)
"""
import collections
import csv
import json
import re
import subprocess
import sys


def launches(path, out, cmd):
    rows = list(csv.reader(open(path, errors="ignore")))
    hdr, agg = None, collections.OrderedDict()
    for r in rows:
        if "Kernel Name" in r:
            hdr = r
            continue
        if hdr is None or len(r) < len(hdr):
            continue
        d = dict(zip(hdr, r))
        if d.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = re.sub(r"^void ", "", d["Kernel Name"])
        name = re.sub(r"\(.*", "", name)
        name = re.sub(r"<unnamed>::", "", name)[-70:]
        v = float(d["Metric Value"].replace(",", ""))
        u = d["Metric Unit"]
        ms = v / 1e6 if u.startswith("n") else (v / 1e3 if u.startswith("u") else (v if u.startswith("m") else v * 1e3))
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += ms
    tot = sum(t for _, t in agg.values())
    with open(out, "w") as f:
        f.write(f"# ncu launch list\n\n`{cmd}`\n\n(per-launch times are cold-cache and serialised; shares of the total are what to compare)\n\n")
        f.write("| kernel | launches | mean ms | total ms | share |\n|---|---:|---:|---:|---:|\n")
        for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| `{k}` | {n} | {t / n:.3f} | {t:.2f} | {100 * t / tot:.1f}% |\n")
    print(open(out).read())


KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_op_umma_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_uniform.sum",
        "lts__t_sector_hit_rate.pct", "lts__t_bytes.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "l1tex__t_bytes.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed"]


def raw(rep, out):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, units = rows[0], rows[1]
    res = []
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        k = {"kernel": re.sub(r"\(.*", "", d.get("Kernel Name", ""))[-60:]}
        for key in hdr:
            if any(key == x or key.startswith(x) for x in KEYS) or "umma" in key or "tensor" in key:
                try:
                    k[key + " [" + units[hdr.index(key)] + "]"] = float(d[key].replace(",", ""))
                except Exception:
                    pass
        res.append(k)
    json.dump(res, open(out, "w"), indent=1)
    for k in res:
        print(json.dumps(k)[:1500])


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3], sys.argv[4] if len(sys.argv) > 4 else "")
    else:
        raw(sys.argv[2], sys.argv[3])
