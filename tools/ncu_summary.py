"""Key per-kernel metrics of an ncu report (--set full) as text.
usage: python tools/ncu_summary.py rep.ncu-rep | raw.csv   (raw.csv = `ncu -i rep --page raw --csv`)

FP64 pipe counters (calibrated on tools/fp64_peak.cu, profiles/ncu_peak_pipes_r3.txt): a pure DMMA loop reads 99.98 % on
sm__pipe_tensor_subpipe_dmma_cycles_active (the *_realtime / pct_of_peak_sustained_elapsed tensor counter reads only 65 % on
the same loop), a pure DFMA loop 91.8 % on sm__pipe_fp64_cycles_active, and BOTH read ~100 % on
sm__pipe_shared_cycles_active: DMMA and DFMA share one FP64 pipe, which is the roof of every kernel here."""
import csv
import subprocess
import sys

rep = sys.argv[1]
out = open(rep).read() if rep.endswith(".csv") else subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
WANT = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__grid_size", "grid"), ("launch__block_size", "block"),
    ("launch__registers_per_thread", "regs/thread"),
    ("launch__shared_mem_per_block_dynamic", "dyn smem/block"),
    ("launch__occupancy_limit_shared_mem", "CTAs/SM (smem limit)"),
    ("launch__occupancy_limit_registers", "CTAs/SM (reg limit)"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("dram__bytes_read.sum", "dram read"), ("dram__bytes_write.sum", "dram write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram % of peak"),
    ("lts__t_sector_hit_rate.pct", "L2 hit %"),
    ("sm__pipe_shared_cycles_active.avg.pct_of_peak_sustained_active", "FP64 pipe (DMMA + DFMA share it) active %"),
    ("sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active", "  of which DMMA (tensor sub-pipe) %"),
    ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "  of which DFMA / DMUL / DADD %"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "shared-mem wavefronts % of peak"),
    ("derived__memory_l1_wavefronts_shared_excessive", "shared wavefronts excessive (bank conflicts)"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "shared wavefronts total"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("sm__inst_executed.avg.per_cycle_active", "IPC"),
]
STALL = "smsp__average_warps_issue_stalled_"
for r in rows[2:]:
    d = dict(zip(hdr, r))
    u = dict(zip(hdr, units))
    print("==", d["Kernel Name"])
    for key, label in WANT:
        c = [h for h in hdr if h.endswith(key) or h == key]
        if c:
            print(f"   {label:48s} {d[c[0]]} {u[c[0]]}")
    st = sorted(((float(d[h].replace(',', '') or 0), h) for h in hdr
                 if STALL in h and h.endswith("_per_issue_active.ratio") and "not_issued" not in h), reverse=True)[:5]
    print("   top stalls (warps per issue): " + ", ".join(f"{h.split(STALL)[1].split('_per_issue')[0]} {v:.2f}" for v, h in st))
