"""Trim the raw `ncu --page raw --csv` exports of scripts/ncu_r02.sh (gpurun_out/ncu/*.csv, hundreds of columns) into the
small tracked summaries profiles/<tag>_ncu_<name>.csv, and the launch list into <tag>_launches_decode_step.csv.
Usage: python scripts/summarize_ncu_csv.py r02"""
import collections
import csv
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
SRC, DST = ROOT / "gpurun_out" / "ncu", ROOT / "profiles"
tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
KEYS = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static",
        "launch__cluster_dim_x", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_sector_hit_rate.pct",
        "lts__t_bytes.sum", "sm__inst_executed_pipe_tensor.sum", "smsp__cycles_active.avg",
        "smsp__average_warp_latency_issue_stalled_barrier.ratio", "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio"]
for f in sorted(SRC.glob("*.csv")):
    if f.name == "launches.csv":
        continue
    rows = [r for r in csv.reader(f.open()) if r]
    hdr_i = next((i for i, r in enumerate(rows) if "Kernel Name" in r), None)
    if hdr_i is None:
        print("skip", f.name)
        continue
    hdr, units = rows[hdr_i], rows[hdr_i + 1]
    idx = [i for i, h in enumerate(hdr) if h in KEYS]
    out = DST / f"{tag}_ncu_{f.stem}.csv"
    with out.open("w", newline="") as g:
        w = csv.writer(g)
        w.writerow([hdr[i] + (f" [{units[i]}]" if units[i] else "") for i in idx])
        for r in rows[hdr_i + 2:]:
            if len(r) == len(hdr):
                w.writerow([r[i] for i in idx])
    print("wrote", out.name, len(rows) - hdr_i - 2, "launches")
ll = SRC / "launches.csv"
if ll.exists():
    rows = [r for r in csv.reader(ll.open()) if r]
    hdr_i = next((i for i, r in enumerate(rows) if "Kernel Name" in r), None)
    if hdr_i is not None:
        hdr = rows[hdr_i]
        kn, mv = hdr.index("Kernel Name"), hdr.index("Metric Value")
        launches = [(r[kn].split("(")[0], float(r[mv].replace(",", ""))) for r in rows[hdr_i + 1:] if len(r) == len(hdr)]
        # the last decode step = everything after the second-to-last sampler launch
        samp = [i for i, (k, _) in enumerate(launches) if "sampler" in k]
        step = launches[samp[-2] + 1: samp[-1] + 1] if len(samp) >= 2 else launches
        agg = collections.OrderedDict()
        for k, v in step:
            a = agg.setdefault(k, [0, 0.0])
            a[0] += 1
            a[1] += v
        tot = sum(v for _, v in step)
        with (DST / f"{tag}_launches_decode_step.csv").open("w", newline="") as g:
            w = csv.writer(g)
            w.writerow(["kernel", "launches", "total_ns (ncu, serialised, cold cache)", "share"])
            for k, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
                w.writerow([k, c, round(v), round(v / tot, 4)])
        print("wrote launch list:", len(step), "launches in the last decode step")
