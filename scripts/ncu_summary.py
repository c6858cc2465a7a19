"""Summarise an .ncu-rep (read here, no GPU needed): headline metrics, stall reasons, opcode mix and the
most-sampled SASS instructions of the first captured launch.
Usage: ncu_summary.py report.ncu-rep [out.txt] [--samples N]   (N = samples the captured launch rendered; bench.py turns the
summary's DRAM bytes into bytes per sample with it)"""
import collections
import csv
import io
import re
import subprocess
import sys

argv = list(sys.argv[1:])
samples = None
if "--samples" in argv:
    i = argv.index("--samples")
    samples = int(argv[i + 1])
    del argv[i:i + 2]
rep = argv[0]
out = open(argv[1], "w") if len(argv) > 1 else sys.stdout


def ncu(*args):
    return subprocess.run(["ncu", "-i", rep, *args], capture_output=True, text=True).stdout


def P(*a):
    print(*a, file=out)


rows = list(csv.reader(io.StringIO(ncu("--page", "raw", "--csv"))))
hdr, units = rows[0], rows[1]
want = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_adu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "smsp__warps_eligible.avg.per_cycle_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "l1tex__t_sector_hit_rate.pct", "sm__cycles_active.avg", "smsp__cycles_active.avg",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "sm__icc_request_hit_rate.pct", "idc__request_cycles_active.avg.pct_of_peak_sustained_elapsed"]
for r in rows[2:]:
    P("=== kernel:", r[hdr.index("Kernel Name")][:110])
    if samples:
        P(f"  {'samples_in_launch':72s} {samples:>18d}")
    for k in want:
        if k in hdr:
            P(f"  {k:72s} {r[hdr.index(k)]:>18s} {units[hdr.index(k)]}")
    if samples and "smsp__inst_executed.sum" in hdr:
        P(f"  {'warp_instructions_per_sample':72s} {float(r[hdr.index('smsp__inst_executed.sum')].replace(',', '')) / samples:>18.1f}")
    st = []
    for i, k in enumerate(hdr):
        if k.startswith("smsp__average_warps_issue_stalled") and k.endswith("_per_issue_active.ratio"):
            try:
                st.append((float(r[i].replace(",", "")), k[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]))
            except ValueError:
                pass
    P("  stalls (warps per issue):", ", ".join(f"{n}={v:.2f}" for v, n in sorted(st, reverse=True)[:8]))

rows = list(csv.reader(io.StringIO(ncu("--page", "source", "--csv", "--print-source", "sass"))))
hdr = rows[1]
isrc, ismp, iex, ith = (hdr.index(x) for x in ("Source", "# Samples", "Instructions Executed", "Thread Instructions Executed"))
data = []
for r in rows[2:]:
    if len(r) < len(hdr) or r[0] in ("Kernel Name", "Address"):
        if data:
            break
        continue
    data.append(r)
tot_ex = sum(int(r[iex]) for r in data) or 1
tot_th = sum(int(r[ith]) for r in data)
tot_smp = sum(int(r[ismp]) for r in data) or 1
P(f"=== SASS of first launch: {len(data)} instructions, {tot_ex} warp-instructions executed, "
  f"thread efficiency {tot_th / tot_ex / 32:.3f}, {tot_smp} stall samples")
ops, smp = collections.Counter(), collections.Counter()
for r in data:
    m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[isrc])
    op = m.group(2).split(".")[0] if m else "?"
    ops[op] += int(r[iex])
    smp[op] += int(r[ismp])
P("  opcode      executed    share   stall-sample share")
for op, c in ops.most_common(24):
    P(f"  {op:10s} {c:11d}  {100 * c / tot_ex:5.1f}%   {100 * smp[op] / tot_smp:5.1f}%")
P("  most-sampled instructions (# samples, executed, SASS):")
for r in sorted(data, key=lambda r: -int(r[ismp]))[:16]:
    P(f"  {r[ismp]:>6s} {r[iex]:>10s}  {r[isrc].strip()[:100]}")
