"""Turns gpurun_out/ artefacts into the small tracked summaries under profiles/.
usage: python tools/summarize_profiles.py <round-tag> [launches.csv] [name=report.ncu-rep ...]"""
import collections
import csv
import re
import subprocess
import sys

tag = sys.argv[1]
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__cycles_active.avg",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "sm__inst_executed_pipe_tmem.avg.pct_of_peak_sustained_active",
        "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__cycles_active.avg"]

for arg in sys.argv[2:]:
    if arg.endswith(".csv"):
        lines = [l for l in open(arg) if not l.startswith("==")]
        agg = collections.defaultdict(lambda: [0, 0.0])
        for row in csv.DictReader(lines):
            try:
                t = float(row["Metric Value"].replace(",", ""))
            except Exception:
                continue
            u = row["Metric Unit"]
            t = t / 1e3 if u == "ns" else (t * 1e3 if u == "ms" else t)
            name = re.sub(r"\(.*", "", row["Kernel Name"])
            name = re.sub(r"<unnamed>::|void |gb::", "", name)[:90]
            agg[name][0] += 1
            agg[name][1] += t
        tot = sum(v[1] for v in agg.values())
        with open(f"profiles/{tag}_launches_summary.txt", "w") as f:
            f.write(f"# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised): "
                    f"{sum(v[0] for v in agg.values())} launches, {tot:.0f} us total; compare SHARES\n")
            f.write(f"{'share%':>7} {'total_us':>10} {'n':>6} {'avg_us':>9}  kernel\n")
            for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
                f.write(f"{100 * v[1] / tot:7.2f} {v[1]:10.1f} {v[0]:6d} {v[1] / v[0]:9.1f}  {k}\n")
    elif "=" in arg:
        name, rep = arg.split("=", 1)
        out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(out.splitlines()))
        hdr, units = rows[0], rows[1]
        with open(f"profiles/{tag}_ncu_{name}.txt", "w") as f:
            f.write(f"# ncu --set full --clock-control none --import-source on  ({rep})\n")
            for vals in rows[2:]:
                d = dict(zip(hdr, vals))
                f.write(f"kernel: {d.get('Kernel Name')}  grid {d.get('Grid Size')} block {d.get('Block Size')}\n")
                for i, h in enumerate(hdr):
                    if h in KEYS:
                        f.write(f"  {h} = {vals[i]} {units[i]}\n")
