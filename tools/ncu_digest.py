"""Digest of an ncu report exported with --page raw --csv and --page source --csv (developer tool)."""
import csv, sys
raw, src = sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else None
rows = list(csv.reader(open(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ['gpu__time_duration.sum', 'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct', 'sm__warps_active.avg.pct_of_peak',
        'sm__inst_executed_pipe_fma.avg.pct', 'sm__inst_executed_pipe_alu.avg.pct', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg', 'lts__throughput.avg',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'smsp__average_warps_issue_stalled', 'lts__t_sector_hit_rate',
        'launch__grid_size', 'launch__registers_per_thread', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'launch__occupancy_limit']
for i, h in enumerate(hdr):
    if any(h.startswith(w) for w in want) and not any(x in h for x in ('.max', '.min', 'per_second', '.sum.pct', '.sum.peak')):
        try:
            if float(vals[i].replace(',', '')) == 0: continue
        except ValueError: pass
        print("%-90s %-12s %s" % (h, units[i], vals[i]))
if src:
    rows = list(csv.reader(open(src)))
    hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}; data = rows[2:]
    def f(r, k):
        try: return float(r[ix[k]])
        except Exception: return 0.0
    keys = ['# Samples', 'Instructions Executed', 'stall_long_sb', 'stall_no_inst', 'stall_mio', 'stall_short_sb', 'stall_wait', 'stall_barrier', 'stall_not_selected', 'stall_selected', 'L1 Wavefronts Shared Excessive']
    blk = int(sys.argv[3]) if len(sys.argv) > 3 else 150
    print("total samples", sum(f(r, '# Samples') for r in data), "static instructions", len(data))
    for b in range(0, len(data), blk):
        seg = data[b:b + blk]
        s = {k: sum(f(r, k) for r in seg) for k in keys}
        ops = {}
        for r in seg:
            t = r[ix['Source']].split()
            if not t: continue
            op = t[1] if t[0].startswith('@') and len(t) > 1 else t[0]
            ops[op.split('.')[0]] = ops.get(op.split('.')[0], 0) + 1
        top = sorted(ops.items(), key=lambda x: -x[1])[:3]
        print(b, ' '.join('%s=%d' % (k.replace('stall_', '').replace('Instructions Executed', 'inst').replace('# Samples', 'smp').replace('L1 Wavefronts Shared Excessive', 'bankx'), s[k]) for k in keys), top)
