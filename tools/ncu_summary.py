"""Condenses an Nsight Compute report (.ncu-rep, read here without a GPU) into the text summary kept under profiles/:
key raw metrics per kernel launch, warp-stall totals, and the hottest SASS instructions by stall samples.

    python tools/ncu_summary.py gpurun_out/prof_tile.ncu-rep [launch index ...] > profiles/r02_step_kernels_ncu.txt
"""
import csv
import re
import subprocess
import sys

rep = sys.argv[1]
SASS_OF = [int(a) for a in sys.argv[2:]] or [0]  # which launches get the SASS view (default: the first)
WANT = re.compile(
    r"^(gpu__time_duration\.sum|dram__bytes_read\.sum|dram__bytes_write\.sum|dram__throughput\.avg\.pct_of_peak_sustained_elapsed|"
    r"lts__t_bytes\.sum|lts__throughput\.avg\.pct_of_peak_sustained_elapsed|l1tex__throughput\.avg\.pct_of_peak_sustained_elapsed|"
    r"sm__throughput\.avg\.pct_of_peak_sustained_elapsed|sm__warps_active\.avg\.pct_of_peak_sustained_active|"
    r"smsp__issue_active\.avg\.pct_of_peak_sustained_active|smsp__inst_executed\.sum|smsp__cycles_active\.avg|"
    r"launch__grid_size|launch__block_size|launch__registers_per_thread|launch__shared_mem_per_block_dynamic|"
    r"launch__occupancy_limit_(registers|shared_mem|warps)|launch__waves_per_multiprocessor|"
    r"l1tex__data_pipe_lsu_wavefronts_mem_shared\.sum|l1tex__data_bank_conflicts_pipe_lsu_mem_shared\.sum|"
    r"lts__t_requests_srcunit_tex_op_write\.sum|lts__t_sectors_srcunit_tex_op_write\.sum|lts__t_sector_hit_rate\.pct|"
    r"sm__inst_executed_pipe_lsu\.avg\.pct_of_peak_sustained_active|"
    r"smsp__average_warps_issue_stalled_.*_per_issue_active\.ratio)$")

raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
name_col = hdr.index("Kernel Name")
for r in rows[2:]:
    print("=" * 100)
    print("kernel:", r[name_col])
    for h, u, v in zip(hdr, units, r):
        if WANT.match(h):
            print(f"  {h:90s} {v} {u}")

src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
kernels, cur, hdr = [], None, None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "data": []}
        kernels.append(cur)
    elif r and r[0] == "Address":
        hdr = r
    elif cur is not None and hdr is not None and len(r) == len(hdr):
        cur["data"].append(r)
for kinfo in [kernels[i] for i in SASS_OF if i < len(kernels)]:
    data = kinfo["data"]
    si, ii, sc = hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("Source")
    tot = sum(int(r[si]) for r in data) or 1
    toti = sum(int(r[ii]) for r in data) or 1
    print("=" * 100)
    print("SASS view of:", kinfo["name"])
    print(f"warp-stall samples {tot}, warp instructions executed {toti}, SASS instructions {len(data)}")
    stall_cols = [j for j, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    tots = {hdr[j]: sum(int(r[j] or 0) for r in data) for j in stall_cols}
    print("stall totals:", {k: v for k, v in sorted(tots.items(), key=lambda x: -x[1]) if v})
    ops = {}
    for r in data:
        toks = [t for t in r[sc].split() if not t.startswith("@")]
        if toks:
            op = toks[0].split(".")[0]
            ops[op] = ops.get(op, 0) + int(r[ii])
    print("executed by opcode:", {k: v for k, v in sorted(ops.items(), key=lambda x: -x[1])[:14]})
    print("segments between barriers / barrier waits (samples, executed):")
    seg = acc = acci = start = 0
    for i, r in enumerate(data):
        acc += int(r[si]); acci += int(r[ii])
        if re.search(r"BAR\.SYNC|SYNCS\.PHASECHK", r[sc]):
            print(f"  seg {seg:2d} inst {start:4d}-{i:4d}: samples {100 * acc / tot:5.1f}%  executed {100 * acci / toti:5.1f}%   ends: {r[sc].strip()[:70]}")
            seg += 1; acc = acci = 0; start = i + 1
    print(f"  seg {seg:2d} inst {start:4d}-{len(data):4d}: samples {100 * acc / tot:5.1f}%  executed {100 * acci / toti:5.1f}%")
    print("hottest instructions by stall samples:")
    for i, r in sorted(enumerate(data), key=lambda x: -int(x[1][si]))[:20]:
        print(f"  #{i:4d} {100 * int(r[si]) / tot:5.1f}%  exec {r[ii]:>8s}  {r[sc].strip()[:90]}")
