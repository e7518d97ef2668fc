"""Turn the ncu exports of a round into the tracked summaries under profiles/.
usage: ncu_summaries.py raw_page.csv launches.csv algorithmic_bytes"""
import collections
import csv
import json
import sys

raw, launches, alg = sys.argv[1], sys.argv[2], int(sys.argv[3])
rows = list(csv.reader(open(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
d = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
names = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
         "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
         "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
         "smsp__thread_inst_executed_per_inst_executed.ratio", "launch__registers_per_thread", "launch__grid_size",
         "launch__block_size", "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static",
         "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
         "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_sector_hit_rate.pct",
         "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
         "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"]
names += [f"smsp__average_warps_issue_stalled_{k}_per_issue_active.ratio" for k in
          ("long_scoreboard", "no_instruction", "wait", "short_scoreboard", "branch_resolving", "barrier",
           "math_pipe_throttle", "not_selected")]
with open("profiles/r01_scan_ws_kernel_ncu_full_4Mreads.csv", "w") as fh:
    fh.write("Kernel Name,," + d.get("Kernel Name", ("scan_ws_kernel", ""))[0].replace(",", ";") + "\n")
    for n in names:
        if n in d:
            fh.write(f"{n},{d[n][1]},{d[n][0]}\n")


def val(n):
    v, u = d[n]
    return float(v) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}.get(u, 1)


rd, wr = val("dram__bytes_read.sum"), val("dram__bytes_write.sum")
json.dump({"kernel": "scan_ws_kernel",
           "source": "profiles/r01_scan_ws_kernel_ncu_full_4Mreads.csv (ncu --set full, 4,000,000 reads)",
           "algorithmic_bytes": alg, "dram_bytes_read": int(rd), "dram_bytes_write": int(wr),
           "traffic_per_algorithmic_byte": round((rd + wr) / alg, 4)},
          open("profiles/r01_scan_traffic.json", "w"), indent=1)
rows = list(csv.reader(l for l in open(launches) if not l.startswith("==")))
h = rows[0]
ix = {k: i for i, k in enumerate(h)}
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows[1:]:
    if len(r) < len(h) or r[ix["Metric Name"]] != "gpu__time_duration.sum":
        continue
    k = r[ix["Kernel Name"]].split("(")[0]
    v = float(r[ix["Metric Value"]].replace(",", ""))
    v *= {"ns": 1e-3, "us": 1, "ms": 1e3, "nsecond": 1e-3, "usecond": 1, "msecond": 1e3}.get(r[ix["Metric Unit"]], 1)
    agg[k][0] += 1
    agg[k][1] += v
tot = sum(v for k, (n, v) in agg.items() if "synth" not in k)
with open("profiles/r01_launch_summary_bench_20Mreads.csv", "w") as fh:
    fh.write("kernel,launches,total_us,share_of_non_generator_time\n")
    for k, (n, v) in sorted(agg.items(), key=lambda x: -x[1][1]):
        share = "nan" if "synth" in k else f"{100 * v / tot:.1f}"
        fh.write(f"\"{k}\",{n},{v:.1f},{share}%\n")
print(open("profiles/r01_scan_ws_kernel_ncu_full_4Mreads.csv").read())
print(open("profiles/r01_scan_traffic.json").read())
print(open("profiles/r01_launch_summary_bench_20Mreads.csv").read()[:900])
