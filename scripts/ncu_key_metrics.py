"""Key metrics of one kernel launch from an ncu report -> CSV lines (name,unit,value), plus per-basis figures.
usage: ncu_key_metrics.py report.ncu-rep n_bases > profiles/<name>_key_metrics.csv"""
import csv, io, subprocess, sys

rep, n_bases = sys.argv[1], int(sys.argv[2])
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
col = {h: (u, v) for h, u, v in zip(hdr, units, vals)}
WANT = """gpu__time_duration.sum dram__bytes_read.sum dram__bytes_write.sum launch__registers_per_thread launch__grid_size
launch__block_size smsp__inst_executed.sum sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active
sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active
sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum
l1tex__data_pipe_lsu_wavefronts_mem_shared.sum smsp__sass_inst_executed_op_local_ld.sum smsp__sass_inst_executed_op_local_st.sum
sm__warps_active.avg.pct_of_peak_sustained_active smsp__issue_active.avg.pct_of_peak_sustained_active sm__cycles_elapsed.avg
smsp__sass_thread_inst_executed_op_dfma_pred_on.sum.per_cycle_elapsed smsp__sass_thread_inst_executed_op_dmul_pred_on.sum.per_cycle_elapsed
smsp__sass_thread_inst_executed_op_dadd_pred_on.sum.per_cycle_elapsed smsp__thread_inst_executed_per_inst_executed.ratio""".split()
def num(name):
    return float(col[name][1].replace(",", ""))
for k in WANT:
    if k in col:
        print(f"{k},{col[k][0]},{col[k][1]}")
for k in sorted(col):
    if "issue_stalled" in k and k.endswith("per_issue_active.ratio") and "not_issued" not in k:
        print(f"{k},{col[k][0]},{col[k][1]}")
try:
    cyc = num("sm__cycles_elapsed.avg")
    dfma, dmul, dadd = (num(f"smsp__sass_thread_inst_executed_op_{o}_pred_on.sum.per_cycle_elapsed") * cyc for o in ("dfma", "dmul", "dadd"))
    print(f"derived_dfma_thread_inst_per_basis,,{dfma / n_bases:.2f}")
    print(f"derived_dmul_thread_inst_per_basis,,{dmul / n_bases:.2f}")
    print(f"derived_dadd_thread_inst_per_basis,,{dadd / n_bases:.2f}")
    print(f"derived_executed_flops_per_basis,,{(2 * dfma + dmul + dadd) / n_bases:.2f}")
except KeyError:
    pass
print(f"derived_warp_inst_per_basis,,{num('smsp__inst_executed.sum') / n_bases:.2f}")
