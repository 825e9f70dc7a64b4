"""Turn the files tools/evidence_run.sh leaves in gpurun_out/ into the tracked summaries under profiles/:
bench lines, the ncu launch list, the ncu --set full summary (one column per kernel) and the DRAM traffic of
the dominant kernel pair that bench.py reports as roofline.traffic.  Runs here (no GPU): ncu only reads the report.

    python tools/collect_profiles.py [prefix]        # prefix defaults to r1b
"""
import csv, json, os, shutil, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
pre = sys.argv[1] if len(sys.argv) > 1 else "r2"
for f in (f"{pre}_bench_c2_1gpu.json", f"{pre}_bench_c5_1gpu.json", f"{pre}_bench_reference_arm.json",
          f"{pre}_bench_c2_8gpu.json", f"{pre}_bench_c5_8gpu.json", f"{pre}_row_table.md", f"{pre}_row_measurements.json"):
    if os.path.exists(os.path.join(G, f)):
        shutil.copy(os.path.join(G, f), os.path.join(P, f))
if os.path.exists(os.path.join(G, f"{pre}_launches.csv")):
    shutil.copy(os.path.join(G, f"{pre}_launches.csv"), os.path.join(P, f"{pre}_launch_list_bench_c2.csv"))
KEYS = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__block_size",
        "launch__grid_size", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "dram__bytes_read.sum",
        "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__t_sector_hit_rate.pct", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__thread_inst_executed_per_inst_executed.ratio"]


def summarise(report, out_csv):
    """One column per kernel (its last captured launch) of an ncu --set full report -> profiles/<out_csv>."""
    raw = subprocess.run(["ncu", "-i", report, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ki = hdr.index("Kernel Name")
    names = [r[ki].split("(")[0].replace("mop::", "").replace("void ", "") for r in data]
    last = {nm: i for i, nm in enumerate(names)}
    idx = sorted(last.values()); names = [names[i] for i in idx]; data = [data[i] for i in idx]
    with open(os.path.join(P, out_csv), "w", newline="") as fh:
        w = csv.writer(fh)
        w.writerow(["metric", "unit"] + names)
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                w.writerow([k, units[i]] + [r[i] for r in data])
    print(out_csv, ":", names)


for extra in ("c5", "producers"):
    r2 = os.path.join(G, f"{pre}_{extra}.ncu-rep")
    if os.path.exists(r2):
        summarise(r2, f"{pre}_ncu_{extra}_summary.csv")
rep = os.path.join(G, f"{pre}_full.ncu-rep")
if os.path.exists(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ki = hdr.index("Kernel Name")
    names = [r[ki].split("(")[0].replace("mop::", "").replace("void ", "") for r in data]
    last = {nm: i for i, nm in enumerate(names)}           # the last launch of every kernel
    idx = sorted(last.values()); names = [names[i] for i in idx]; data = [data[i] for i in idx]
    keys = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__block_size",
            "launch__grid_size", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "dram__bytes_read.sum",
            "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
            "l1tex__t_sector_hit_rate.pct", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
            "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
            "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__thread_inst_executed_per_inst_executed.ratio"]
    with open(os.path.join(P, f"{pre}_ncu_summary.csv"), "w", newline="") as fh:
        w = csv.writer(fh)
        w.writerow(["metric", "unit"] + names)
        for k in keys:
            if k in hdr:
                i = hdr.index(k)
                w.writerow([k, units[i]] + [r[i] for r in data])
    scale = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1}

    def nbytes(name, key):
        i = hdr.index(key)
        return int(float(data[names.index(name)][i]) * scale[units[i]])
    tr = {"source": f"ncu --set full --clock-control none, tools/prof_step.py with DIAG_B=1024 (profiles/{pre}_ncu_summary.csv)",
          "batch": 1024, "n": 150}
    for nm in names:
        tr[nm] = {"dram_bytes_read": nbytes(nm, "dram__bytes_read.sum"), "dram_bytes_write": nbytes(nm, "dram__bytes_write.sum")}
    pair = [k for k in names if k.startswith("k_tridiag_blk") or k.startswith("k_spectrum_step")]
    tr["dominant_pair_bytes_per_launch"] = sum(tr[k]["dram_bytes_read"] + tr[k]["dram_bytes_write"] for k in pair)
    json.dump(tr, open(os.path.join(P, f"{pre}_traffic.json"), "w"), indent=1)
    print("kernels:", names, "dominant pair bytes:", tr["dominant_pair_bytes_per_launch"])
b = os.path.join(P, f"{pre}_bench_c2_1gpu.json")
if os.path.exists(b):
    d = json.loads(open(b).read().strip().split("\n")[-1])
    print("bench c2: value %.4g  ms/step %.3f  e2e %.4g  resident %.4g  roofline frac %.3f" %
          (d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e_hessian_resident"]["value"], d["roofline"]["frac"]))
