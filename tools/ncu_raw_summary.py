"""Key metrics of every kernel in an ncu report (raw page) as a compact text table.
Usage: python tools/ncu_raw_summary.py REPORT.ncu-rep"""
import csv
import io
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__grid_size", "grid"), ("launch__block_size", "block"), ("launch__registers_per_thread", "regs/thread"),
    ("launch__occupancy_limit_registers", "occupancy limit (regs), blocks/SM"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("smsp__thread_inst_executed_per_inst_executed.ratio", "active threads / warp inst"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("smsp__inst_executed.avg.per_cycle_active", "IPC per scheduler"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "FMA pipe %"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "ALU pipe %"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "XU (MUFU) pipe %"),
    ("smsp__sass_thread_inst_executed_op_ffma_pred_on.sum", "FFMA thread insts"),
    ("smsp__sass_thread_inst_executed_op_fmul_pred_on.sum", "FMUL thread insts"),
    ("smsp__sass_thread_inst_executed_op_fadd_pred_on.sum", "FADD thread insts"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("dram__bytes_read.sum", "dram read"), ("dram__bytes_write.sum", "dram write"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram throughput %"),
    ("lts__t_bytes.sum", "L2 bytes"), ("l1tex__t_bytes.sum", "L1 bytes"),
    ("smsp__cycles_active.avg", "active cycles / scheduler"),
    ("sm__cycles_elapsed.max", "elapsed cycles"),
]


def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print(f"== {r[hdr.index('Kernel Name')]}  (launch id {r[hdr.index('ID')]})")
        for k, label in KEYS:
            if k in hdr:
                j = hdr.index(k)
                print(f"   {label:42s} {r[j]:>18s} {units[j]:12s} [{k}]")
        print()


if __name__ == "__main__":
    main()
