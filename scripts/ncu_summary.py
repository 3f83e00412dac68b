"""Condense an .ncu-rep (read with `ncu -i`) into the few numbers DESIGN.md / profiles/ quote.
usage: python scripts/ncu_summary.py report.ncu-rep [--source]"""
import csv
import io
import subprocess
import sys
from collections import defaultdict

KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "dram__bytes_read.sum",
        "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__warps_eligible.avg.per_cycle_active", "smsp__inst_executed.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "sm__cycles_elapsed.avg", "sm__cycles_elapsed.avg.per_second", "lts__t_bytes.sum"]


def page(rep, name):
    out = subprocess.run(["ncu", "-i", rep, "--page", name, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    rep = sys.argv[1]
    rows = page(rep, "raw")
    hdr, units = rows[0], rows[1]
    for vals in rows[2:]:
        d = dict(zip(hdr, vals))
        u = dict(zip(hdr, units))
        print("##", d.get("Kernel Name", "?"))
        for k in KEYS:
            if k in d:
                print(f"{k:88s} {d[k]:>16s} {u[k]}")
        stalls = {k.replace("smsp__pcsamp_warps_issue_stalled_", ""): float(v) for k, v in d.items()
                  if k.startswith("smsp__pcsamp_warps_issue_stalled_") and "not_issued" not in k and v}
        tot = sum(stalls.values()) or 1.0
        top = sorted(stalls.items(), key=lambda kv: -kv[1])[:7]
        print("top stall reasons: " + ", ".join(f"{k} {100 * v / tot:.1f}%" for k, v in top))
    if "--source" in sys.argv:
        rows = page(rep, "source")
        hdr = rows[1]                                   # row 0 is the kernel name
        si, samp, inst = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
        agg = defaultdict(lambda: [0.0, 0.0])
        seq = []
        for r in rows[2:]:
            try:
                op = r[si].split()[0] if not r[si].strip().startswith("@") else r[si].split()[1]
                op = op.split(".")[0]
                agg[op][0] += float(r[samp] or 0)
                agg[op][1] += float(r[inst] or 0)
                seq.append((op, float(r[samp] or 0)))
            except (ValueError, IndexError):
                pass
        tot = sum(v[0] for v in agg.values()) or 1.0
        print("warp samples by SASS opcode (share of samples, warp instructions executed):")
        for op, (sm, n) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:14]:
            print(f"  {100 * sm / tot:5.1f}%  {n:12.0f}  {op}")
        # phases: split the instruction stream at BAR.SYNC
        phase, acc = 0, defaultdict(float)
        for op, sm in seq:
            acc[phase] += sm
            if op == "BAR":
                phase += 1
        print("warp samples by phase (instruction stream split at BAR.SYNC): " +
              ", ".join(f"phase {k}: {100 * v / tot:.1f}%" for k, v in sorted(acc.items())))


if __name__ == "__main__":
    main()
