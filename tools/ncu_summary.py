"""Summarise `ncu --set full` captures into small tracked files under profiles/.

  python tools/ncu_summary.py gpurun_out/prof_knn.ncu-rep gpurun_out/prof_edge.ncu-rep --tag r01

writes profiles/<tag>_ncu_<report>.md (one table per launch: duration, DRAM bytes, throughputs, pipes,
occupancy, registers) and merges per-kernel DRAM traffic (bytes per launch, mean over the captured launches)
into profiles/ncu_traffic.json, which bench.py reads for `roofline.traffic`.
"""
import argparse
import csv
import io
import json
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

METRICS = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "dram read"),
    ("dram__bytes_write.sum", "dram write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram % of peak"),
    ("lts__t_sector_hit_rate.pct", "L2 hit %"),
    ("l1tex__t_sector_hit_rate.pct", "L1 hit %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue active %"),
    ("smsp__thread_inst_executed_per_inst_executed.ratio", "active lanes / inst"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "pipe alu %"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "pipe fma %"),
    ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "pipe fp64 %"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "pipe lsu %"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe %"),
    ("sm__inst_executed_pipe_tensor.avg.pct_of_peak_sustained_active", "tensor inst %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
    ("launch__registers_per_thread", "registers / thread"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__shared_mem_per_block_dynamic", "dyn smem / block"),
    ("smsp__inst_executed.sum", "warp instructions"),
]


def to_bytes(val, unit):
    v = float(val)
    u = unit.lower()
    return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("reports", nargs="+")
    ap.add_argument("--tag", default="r01")
    a = ap.parse_args()
    traffic_path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    try:
        traffic = json.load(open(traffic_path))
    except Exception:
        traffic = {}
    for rep in a.reports:
        if rep.endswith(".csv"):     # `ncu -i x.ncu-rep --page raw --csv > x.csv` exported on the GPU box (reports > 64 MiB do not travel)
            raw = open(rep).read()
        else:
            raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(raw)))
        head, units, body = rows[0], rows[1], rows[2:]
        col = {h: i for i, h in enumerate(head)}
        name = os.path.splitext(os.path.basename(rep))[0]
        out = [f"# ncu --set full summary: {os.path.basename(rep)}", "",
               "Per-launch metrics (cold caches, serialised, ~40 replays: compare shares, not absolutes).", ""]
        per_kernel = {}
        for r in body:
            kname = r[col["Kernel Name"]]
            short = re.sub(r"^void ", "", kname).split("(")[0]
            out += [f"## {short}", "", "| metric | value | unit |", "|---|---|---|"]
            for m, label in METRICS:
                if m in col:
                    out.append(f"| {label} (`{m}`) | {r[col[m]]} | {units[col[m]]} |")
            out.append("")
            if "dram__bytes_read.sum" in col:
                tb = to_bytes(r[col["dram__bytes_read.sum"]], units[col["dram__bytes_read.sum"]]) + \
                    to_bytes(r[col["dram__bytes_write.sum"]], units[col["dram__bytes_write.sum"]])
                per_kernel.setdefault(short, []).append(tb)
        best = {}
        for kshort, vals in per_kernel.items():
            base = kshort.split("<")[0]
            entry = {"bytes_per_launch": sum(vals) / len(vals), "launches": len(vals), "instance": kshort,
                     "report": f"profiles/{a.tag}_ncu_{name}.md"}
            # one entry per base name: the largest instance of this capture (the one that dominates the step)
            if base not in best or best[base]["bytes_per_launch"] < entry["bytes_per_launch"]:
                best[base] = entry
        traffic.update(best)
        with open(os.path.join(ROOT, "profiles", f"{a.tag}_ncu_{name}.md"), "w") as f:
            f.write("\n".join(out))
    with open(traffic_path, "w") as f:
        json.dump(traffic, f, indent=1, sort_keys=True)
    print(json.dumps(traffic, indent=1))


if __name__ == "__main__":
    main()
