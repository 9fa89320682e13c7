#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, no GPU needed): per kernel the numbers DESIGN.md / bench.py quote.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep [--sass KERNEL_REGEX] > profiles/<name>.txt
"""
import csv
import io
import subprocess
import sys

METRICS = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "dram read"),
    ("dram__bytes_write.sum", "dram write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram % of peak"),
    ("lts__t_bytes.sum", "L2 bytes"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM % of peak"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("launch__registers_per_thread", "registers/thread"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__shared_mem_per_block_dynamic", "dyn smem/block"),
    ("launch__occupancy_limit_registers", "occ limit regs"),
    ("launch__occupancy_limit_shared_mem", "occ limit smem"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem bank conflicts"),
    ("sm__inst_executed_pipe_fp64.sum", "fp64 pipe instr"),
    ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "fp64 pipe busy %"),
]
STALLS = ["barrier", "long_scoreboard", "short_scoreboard", "wait", "math_pipe_throttle", "mio_throttle", "lg_throttle",
          "membar", "no_instruction", "branch_resolving", "dispatch_stall", "not_selected", "selected", "sleeping", "drain"]


def raw(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    return hdr, units, rows[2:]


def main():
    path = sys.argv[1]
    hdr, units, rows = raw(path)
    ki = hdr.index("Kernel Name")
    print(f"# {path}")
    for r in rows:
        print(f"\n== {r[ki]}")
        for key, label in METRICS:
            if key in hdr:
                i = hdr.index(key)
                print(f"  {label:24s} {r[i]:>16s} {units[i]}")
        parts = []
        for s in STALLS:
            key = f"smsp__average_warps_issue_stalled_{s}_per_issue_active.ratio"
            if key in hdr:
                parts.append((float(r[hdr.index(key)] or 0), s))
        parts.sort(reverse=True)
        print("  stalls (warps per issue): " + ", ".join(f"{n} {v:.2f}" for v, n in parts[:7]))
    if "--sass" in sys.argv:
        pat = sys.argv[sys.argv.index("--sass") + 1]
        out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--kernel-name", f"regex:{pat}"],
                             capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(out)))
        h = next(i for i, r in enumerate(rows) if "Source" in r and "Instructions Executed" in r)
        hdr = rows[h]
        si, ii, sm = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
        body = []
        for r in rows[h + 1:]:
            try:
                body.append((int(r[ii]), int(r[sm]), r[si].strip()))
            except (ValueError, IndexError):
                pass
        tot_i = sum(b[0] for b in body) or 1
        tot_s = sum(b[1] for b in body) or 1
        print(f"\n== SASS of {pat}: {len(body)} instructions, {tot_i} executed, {tot_s} samples; by opcode")
        agg = {}
        for n, s, src in body:
            op = src.split()[0] if not src.startswith("@") else src.split()[1]
            op = op.split(".")[0]
            a = agg.setdefault(op, [0, 0])
            a[0] += n
            a[1] += s
        for op, (n, s) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:24]:
            print(f"  {op:12s} {100 * n / tot_i:6.2f}% of instructions  {100 * s / tot_s:6.2f}% of samples")
        print("  hottest instructions by samples:")
        for n, s, src in sorted(body, key=lambda b: -b[1])[:24]:
            print(f"    {100 * s / tot_s:5.2f}%  x{n:<10d} {src[:100]}")


if __name__ == "__main__":
    main()
