#!/usr/bin/env python
"""Attributes an `ncu --set full --import-source on` capture of a simplex wave kernel to source lines /
functions of gomilp_b200/csrc/simplex_cta.cuh, by joining the report's per-SASS-instruction samples with the
line table of the shipped cubin (nvdisasm -g).  Usage:
  python profiles/attribute.py gpurun_out/prof.ncu-rep simplex_wave_reg > profiles/<name>.txt
"""
import collections
import csv
import glob
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep, kname = sys.argv[1], sys.argv[2]
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(ROOT, "gomilp_b200/_build/libgomilp_b200.so")], cwd=tmp,
               check=True, stdout=subprocess.DEVNULL)
dis = []
for cubin in glob.glob(os.path.join(tmp, "*.cubin")):
    out = subprocess.run(["nvdisasm", "-g", "-c", cubin], stdout=subprocess.PIPE, text=True).stdout
    if re.search(kname + r"E", out):   # the mangled name ends the identifier with E: not a prefix of another kernel
        dis = out.split("\n")
        break
src_csv = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], stdout=subprocess.PIPE, text=True).stdout
raw_csv = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout

start = [i for i, l in enumerate(dis) if l.startswith(".text.") and re.search(kname + r"E", l)][0]
pat = re.compile(r'//## File "([^"]+)", line (\d+)')
chain_pat = re.compile(r'inlined at "([^"]+)", line (\d+)')
cur, insts, i = None, [], start + 1
while i < len(dis):
    l = dis[i]
    if l.startswith("//-----") and i > start + 3:
        break
    m = pat.search(l)
    if m:
        cur = [(m.group(1).split("/")[-1], int(m.group(2)))] + [(a.split("/")[-1], int(b)) for a, b in chain_pat.findall(l)]
    elif re.match(r"\s+/\*[0-9a-f]{4,}\*/", l):
        insts.append(cur)
    i += 1
rows = list(csv.reader(src_csv.split("\n")))
rows = [r for r in rows if r]
hdr, data = rows[1], rows[2:]
ci, ie = hdr.index("# Samples"), hdr.index("Instructions Executed")
src = open(os.path.join(ROOT, "gomilp_b200/csrc/simplex_cta.cuh")).read().split("\n")
funcs = []
for n, l in enumerate(src, 1):
    m = re.match(r"\s+(?:template <class F>\s*)?GM_DEV\s+[\w:<>\s\*&]+?\s+(\w+)\(", l)
    if m:
        funcs.append((n, m.group(1)))


def func_of(ln):
    name = "?"
    for n, f in funcs:
        if n <= ln:
            name = f
        else:
            break
    return name


by_line = collections.defaultdict(lambda: [0, 0])
by_func = collections.defaultdict(lambda: [0, 0])
for k, r in enumerate(data):
    if k >= len(insts):
        break
    ch = insts[k] or [("none", 0)]
    by_line[ch[0]][0] += int(r[ci]); by_line[ch[0]][1] += int(r[ie])
    fr = [c for c in ch if c[0] == "simplex_cta.cuh"]
    key = func_of(fr[-1][1]) if fr else ch[-1][0]
    by_func[key][0] += int(r[ci]); by_func[key][1] += int(r[ie])
tot = sum(v[0] for v in by_line.values()); toti = sum(v[1] for v in by_line.values())
rr = list(csv.reader(raw_csv.split("\n")))
rr = [r for r in rr if r]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
        "smsp__inst_executed.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "launch__grid_size", "launch__block_size",
        "smsp__pcsamp_warps_issue_stalled_barrier", "smsp__pcsamp_warps_issue_stalled_short_scoreboard",
        "smsp__pcsamp_warps_issue_stalled_long_scoreboard", "smsp__pcsamp_warps_issue_stalled_math_pipe_throttle",
        "smsp__pcsamp_warps_issue_stalled_wait", "smsp__pcsamp_warps_issue_stalled_mio_throttle",
        "smsp__pcsamp_warps_issue_stalled_not_selected", "smsp__pcsamp_warps_issue_stalled_selected",
        "smsp__pcsamp_warps_issue_stalled_branch_resolving", "smsp__pcsamp_warps_issue_stalled_dispatch_stall",
        "smsp__pcsamp_warps_issue_stalled_no_instructions", "smsp__pcsamp_warps_issue_stalled_lg_throttle"]
print(f"# {os.path.basename(rep)} kernel {kname}: {len(insts)} SASS instructions, {tot} samples, {toti} warp-instructions")
for j, h in enumerate(rr[0]):
    if h in want:
        print(f"{h} [{rr[1][j]}] = {rr[2][j]}")
print("\n## by function (outermost frame in simplex_cta.cuh): share of stall samples / share of executed warp-instructions")
for k, v in sorted(by_func.items(), key=lambda kv: -kv[1][0])[:25]:
    print(f"{k:28s} samples={v[0] / tot:.3f} inst={v[1] / toti:.3f}")
print("\n## top source lines")
for (f, ln), v in sorted(by_line.items(), key=lambda kv: -kv[1][0])[:40]:
    text = src[ln - 1].strip()[:100] if f == "simplex_cta.cuh" and 0 < ln <= len(src) else ""
    print(f"{f}:{ln:4d} samples={v[0] / tot:.3f} inst={v[1] / toti:.3f}  {text}")
