#!/usr/bin/env python
"""Summarise an .ncu-rep (raw + source pages) into the few numbers the roofline discussion needs.
usage: python profiles/ncu_summary.py gpurun_out/x.ncu-rep [kernel-index]"""
import csv, subprocess, sys, io, json

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2 + (int(sys.argv[2]) if len(sys.argv) > 2 else 0)]
m = {h: (vals[i], units[i]) for i, h in enumerate(hdr)}
keys = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "launch__grid_size", "launch__block_size", "smsp__inst_executed.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum"]
out = {}
for k in keys:
    if k in m:
        out[k] = " ".join(m[k]).strip()
for h in hdr:
    if "issue_stalled" in h and h.endswith("per_issue_active.ratio"):
        v = float(m[h][0])
        if v >= 0.3:
            out[h.replace("smsp__average_warps_issue_stalled_", "stall_").replace("_per_issue_active.ratio", "")] = round(v, 2)
print(json.dumps(out, indent=1))

src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h = rows[1]; data = rows[2:]
isrc, isamp, iex = h.index("Source"), h.index("# Samples"), h.index("Instructions Executed")
tot = sum(int(r[isamp]) for r in data); totex = sum(int(r[iex]) for r in data)
bars = [i for i, r in enumerate(data) if "BAR.SYNC" in r[isrc]]
seg = [0] + bars + [len(data)]
print(f"SASS instructions {len(data)}, stall samples {tot}, warp-instructions executed {totex}")
for a, b in zip(seg[:-1], seg[1:]):
    s_ = sum(int(r[isamp]) for r in data[a:b]); e_ = sum(int(r[iex]) for r in data[a:b])
    print(f"  between barriers [{a},{b}): samples {100*s_/max(tot,1):5.1f}%  executed {100*e_/max(totex,1):5.1f}%")
stalls = [x for x in h if x.startswith("stall_") and "Not Issued" not in x]
# finer regions: split at barriers, calls, unconditional exits / returns and the mbarrier wait (phase boundaries)
import re
marks = [i for i, r in enumerate(data) if re.search(r"^\s*(BAR\.SYNC|CALL|EXIT|RET|SYNCS\.PHASECHK)", r[isrc].strip())]
seg2 = sorted(set([0] + [i + 1 for i in marks] + [len(data)]))
print("regions (split at BAR.SYNC / CALL / EXIT / RET / mbarrier wait):")
for a, b in zip(seg2[:-1], seg2[1:]):
    s_ = sum(int(r[isamp]) for r in data[a:b]); e_ = sum(int(r[iex]) for r in data[a:b])
    if s_ * 200 < tot and e_ * 200 < totex:
        continue
    st = sorted(((x, sum(int(r[h.index(x)]) for r in data[a:b])) for x in stalls), key=lambda x: -x[1])[:4]
    st = [(n, round(100 * v / max(tot, 1), 1)) for n, v in st]
    print(f"  [{a:5d},{b:5d}) ends at '{data[b-1][isrc].strip()[:28]}': samples {100*s_/max(tot,1):5.1f}%  executed {100*e_/max(totex,1):5.1f}%  {st}")
print("top stall sites:")
for r in sorted(data, key=lambda r: -int(r[isamp]))[:14]:
    st = sorted(((x, int(r[h.index(x)])) for x in stalls), key=lambda x: -x[1])[:2]
    print(f"  {data.index(r):5d} {100*int(r[isamp])/max(tot,1):5.1f}% exec {r[iex]:>10} {r[isrc].strip()[:60]:60s} {st}")
