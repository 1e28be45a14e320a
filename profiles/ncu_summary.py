#!/usr/bin/env python
"""Summarises an .ncu-rep: key raw metrics per kernel, opcode mix, stall reasons, hottest SASS.
usage: python profiles/ncu_summary.py report.ncu-rep [kernel_index]"""
import collections, csv, io, re, subprocess, sys

rep = sys.argv[1]
kid = int(sys.argv[2]) if len(sys.argv) > 2 else None
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
keys = ['Kernel Name', 'launch__grid_size', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_bytes.sum',
        'launch__registers_per_thread', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'sm__cycles_elapsed.avg']
for n, r in enumerate(rows[2:]):
    if kid is not None and n != kid:
        continue
    print("== kernel", n)
    for k in keys:
        if k in hdr:
            i = hdr.index(k)
            print("  %-68s %s %s" % (k, r[i], units[i]))
for n in range(len(rows) - 2):
    if kid is not None and n != kid:
        continue
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-id", ":::%d" % (n + 1)],
                         capture_output=True, text=True).stdout
    srows = list(csv.reader(io.StringIO(src)))
    h = srows[1]
    iS, iE, iSm = h.index('Source'), h.index('Instructions Executed'), h.index('# Samples')
    st = [i for i, x in enumerate(h) if x.startswith('stall_') and 'Not Issued' not in x]
    ops, tot, stalls, top, seen = collections.Counter(), 0, collections.Counter(), [], set()
    for r in srows[2:]:
        if len(r) < len(h) or r[0] in seen:
            continue
        seen.add(r[0])
        m = re.match(r'\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)', r[iS])
        try:
            e, s = int(r[iE]), int(r[iSm])
        except ValueError:
            continue
        if m:
            ops[m.group(2).split('.')[0]] += e
        tot += e
        for i in st:
            try:
                stalls[h[i]] += int(r[i])
            except ValueError:
                pass
        top.append((s, r[iS].strip()[:70], e))
    print("== kernel", n, "executed warp-instructions", tot)
    print("  opcode mix:", ", ".join("%s %.1f%%" % (o, 100.0 * e / tot) for o, e in ops.most_common(14)))
    ts = sum(stalls.values()) or 1
    print("  stalls:", ", ".join("%s %.1f%%" % (k[6:], 100.0 * v / ts) for k, v in stalls.most_common(9)))
    top.sort(reverse=True)
    print("  hottest:", "; ".join("%d:%s" % (s, t) for s, t, e in top[:8]))
