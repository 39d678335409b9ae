"""Print the metrics that matter from an ncu report: python tools/ncu_summary.py X.ncu-rep [kernel-regex]"""
import csv, subprocess, sys, io
rep = sys.argv[1]
cmd = ["ncu", "-i", rep, "--page", "raw", "--csv"]
if len(sys.argv) > 2: cmd += ["--kernel-name", "regex:" + sys.argv[2]]
out = subprocess.run(cmd, capture_output=True, text=True).stdout
r = list(csv.reader(io.StringIO(out)))
h = r[0]
keys = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__cycles_active.avg", "sm__cycles_elapsed.avg", "lts__t_bytes.sum", "l1tex__t_bytes.sum",
        "smsp__inst_executed_op_local_ld.sum", "smsp__inst_executed_op_local_st.sum", "sm__sass_inst_executed_op_local.sum"]
stall = [k for k in h if "smsp__average_warp" in k and "issue_stalled" in k and k.endswith("_per_warp_active.pct")] or \
        [k for k in h if "smsp__average_warps_issue_stalled" in k]
for row in r[2:]:
    d = dict(zip(h, row))
    print("=" * 100)
    for k in keys:
        if k in d: print(f"{k:75s} {d[k]}")
    st = sorted(((float(d[k].replace(',', '')) if d[k] not in ('', 'n/a') else 0.0, k) for k in stall), reverse=True)
    for v, k in st[:10]:
        print(f"   stall {k.replace('smsp__average_warps_issue_stalled_', '').replace('smsp__average_warp_latency_issue_stalled_', ''):60s} {v:.2f}")
