"""Turn ncu outputs into the small, committed summaries under profiles/.

  python tools/summarize_ncu.py launches <launches.csv> <out.md> [step_ms]
  python tools/summarize_ncu.py full <report.ncu-rep> <out.md>
"""
import collections
import csv
import re
import subprocess
import sys


def short(name):
    name = re.sub(r"void |pmr::|\(anonymous namespace\)::|<unnamed>::|unnamed>::", "", name)
    return name.split("(")[0][:90]


def launches(path, out, note=""):
    lines = [l for l in open(path) if not l.startswith("==")]
    rows = list(csv.DictReader(lines))
    agg = collections.OrderedDict()
    for r in rows:
        if "Metric Name" in r and not r["Metric Name"].startswith("gpu__time_duration"):
            continue
        k = short(r["Kernel Name"])
        t = float(r["Metric Value"].replace(",", ""))
        if r["Metric Unit"] in ("us", "usecond"):
            t *= 1e3
        elif r["Metric Unit"] in ("ms", "msecond"):
            t *= 1e6
        elif r["Metric Unit"] in ("s", "second"):
            t *= 1e9
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += t
    tot = sum(v[1] for v in agg.values())
    with open(out, "w") as f:
        nl = sum(v[0] for v in agg.values())
        f.write(f"# ncu launch list summary ({nl} launches, {tot/1e6:.1f} ms of kernel time)\n\n")
        f.write("`ncu --metrics gpu__time_duration.sum --clock-control none` (cold-cache, serialised: "
                "compare shares, not absolutes).\n\n")
        if note:
            f.write(note + "\n\n")
        f.write("| kernel | launches | total ms | share | avg us |\n|---|---:|---:|---:|---:|\n")
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| `{k}` | {v[0]} | {v[1]/1e6:.2f} | {100*v[1]/tot:.1f}% | {v[1]/1e3/v[0]:.1f} |\n")


KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__ops_path_tensor_src_fp64.sum.pct_of_peak_sustained_elapsed",
    "sm__ops_path_tensor_src_fp64.sum.per_second",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second",
    "dram__bytes_write.sum.per_second", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "sm__cycles_elapsed.avg.per_second",
]


def full(rep, out):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    with open(out, "w") as f:
        f.write(f"# ncu --set full summary of `{rep.split('/')[-1]}`\n\n")
        for r in rows[2:]:
            d = dict(zip(hdr, r))
            f.write(f"## `{short(d.get('Kernel Name', '?'))}`  grid {d.get('Grid Size')} block {d.get('Block Size')}\n\n")
            f.write("| metric | value | unit |\n|---|---:|---|\n")
            for k in KEYS:
                if k in d:
                    f.write(f"| {k} | {d[k]} | {units[hdr.index(k)]} |\n")
            f.write("\n")


def traffic(path, out, label, pattern, note=""):
    """Per-launch DRAM bytes (read + write) of the launches whose kernel name matches ``pattern``
    from a multi-metric launch list (gpu__time_duration.sum, dram__bytes_read.sum,
    dram__bytes_write.sum) -> {label: {...}} merged into the JSON file ``out``."""
    import json
    import os

    lines = [l for l in open(path) if not l.startswith("==")]
    rows = list(csv.DictReader(lines))
    per = collections.OrderedDict()
    for r in rows:
        if not re.search(pattern, r["Kernel Name"]):
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"].lower()
        if r["Metric Name"].startswith("dram__bytes"):
            mult = {"byte": 1.0, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(unit, 1.0)
            per.setdefault(r["ID"], [0.0, 0.0])[0] += v * mult
        elif r["Metric Name"].startswith("gpu__time_duration"):
            mult = {"ns": 1e-6, "nsecond": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1.0, "msecond": 1.0}.get(unit, 1e-6)
            per.setdefault(r["ID"], [0.0, 0.0])[1] += v * mult
    n = len(per)
    data = {}
    if os.path.exists(out):
        data = json.load(open(out))
    data[label] = {"dram_bytes_per_launch": sum(v[0] for v in per.values()) / max(1, n), "launches": n,
                   "total_dram_bytes": sum(v[0] for v in per.values()),
                   "total_ms_under_ncu": sum(v[1] for v in per.values()),
                   "source": f"ncu dram__bytes_read.sum + dram__bytes_write.sum over the {n} launches matching "
                             f"/{pattern}/ of one bench.py run ({os.path.basename(path)}); {note}"}
    json.dump(data, open(out, "w"), indent=1)


if __name__ == "__main__":
    if sys.argv[1] == "traffic":
        traffic(sys.argv[2], sys.argv[3], sys.argv[4], sys.argv[5], sys.argv[6] if len(sys.argv) > 6 else "")
    elif sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3], sys.argv[4] if len(sys.argv) > 4 else "")
    else:
        full(sys.argv[2], sys.argv[3])
