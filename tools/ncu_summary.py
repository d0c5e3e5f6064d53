"""Text summary of an .ncu-rep (`ncu --set full`) for profiles/: one block per captured launch with the metrics the
roofline discussion uses.   python tools/ncu_summary.py report.ncu-rep > profiles/ncu_<what>_summary.txt"""
import csv
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__grid_size", "grid"), ("launch__block_size", "block"), ("launch__registers_per_thread", "regs/thread"),
    ("launch__shared_mem_per_block_dynamic", "dyn smem/block"),
    ("sm__cycles_active.avg", "SM active cycles"),
    ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM write"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "DRAM % of peak"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 % of peak"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "tensor pipe active %"),
    ("sm__ops_path_tensor_op_utchmma_src_fp16_dst_fp32_sparsity_off.avg.pct_of_peak_sustained_elapsed", "tcgen05 fp16 ops % of peak"),
    ("sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active", "mma.sync (HMMA) issue %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("sm__inst_executed_pipe_xu_realtime.avg.pct_of_peak_sustained_elapsed", "XU (MUFU) pipe %"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "shared-memory wavefronts % of peak"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "shared bank conflicts"),
    ("smsp__mem_tensor_reads_op_ldt.sum.pct_of_peak_sustained_elapsed", "TMEM loads % of peak"),
]
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
h, units = rows[0], rows[1]
for row in rows[2:]:
    d = dict(zip(h, row))
    u = dict(zip(h, units))
    print("=" * 100)
    print(d.get("Kernel Name", "?")[:160])
    for k, label in KEYS:
        if k in d and d[k] != "":
            print(f"  {label:38s} {d[k]:>18s} {u.get(k, '')}")
    stalls = {k.replace("smsp__pcsamp_warps_issue_stalled_", ""): float(v) for k, v in d.items()
              if k.startswith("smsp__pcsamp_warps_issue_stalled_") and "not_issued" not in k and v not in ("", "0")}
    tot = sum(stalls.values()) or 1.0
    top = sorted(stalls.items(), key=lambda kv: -kv[1])[:6]
    print("  warp stall samples: " + ", ".join(f"{k} {100 * v / tot:.0f}%" for k, v in top))
