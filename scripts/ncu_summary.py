"""Summarise an .ncu-rep (raw + source pages) into text: python tests/ncu_summary.py <rep> [kernel#]"""
import collections, csv, re, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, data = rows[0], rows[1], rows[2:]
keys = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'smsp__inst_executed.sum', 'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct',
        'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'sm__cycles_elapsed.max',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__warps_eligible.avg.per_cycle_active',
        'smsp__warps_active.avg.per_cycle_active', 'lts__t_bytes.sum', 'sm__inst_executed_pipe_lsu.sum']
for k in keys:
    if k in hdr:
        i = hdr.index(k)
        print("%-70s %-10s %s" % (k, units[i], [r[i][:60] for r in data]))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
idx = [i for i, r in enumerate(rows) if r and r[0] == 'Address']
which = int(sys.argv[2]) if len(sys.argv) > 2 else 0
h = rows[idx[which]]
end = idx[which + 1] - 1 if len(idx) > which + 1 else len(rows)
d = [r for r in rows[idx[which] + 1:end] if len(r) > 8]
ci, ti, si = h.index('Instructions Executed'), h.index('Thread Instructions Executed'), h.index('# Samples')
ops, opt, smp = collections.Counter(), collections.Counter(), collections.Counter()
tot = 0
for r in d:
    m = re.match(r'\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)', r[1])
    op = m.group(2).split('.')[0] if m else '?'
    n = int(r[ci]); ops[op] += n; opt[op] += int(r[ti]); smp[op] += int(r[si]); tot += n
print("static SASS instrs", len(d), "warp insts", tot)
for op, n in ops.most_common(22):
    print("  %-10s %12d %5.1f%%  lanes %.1f  stall-samples %d" % (op, n, 100 * n / tot, opt[op] / max(n, 1), smp[op]))
stall = [c for c in h if c.startswith('stall_') and 'Not Issued' not in c]
tots = {c: sum(int(r[h.index(c)]) for r in d if len(r) > h.index(c)) for c in stall}
s = sum(tots.values())
print("stall reasons:", ", ".join("%s %.1f%%" % (c[6:], 100 * v / s) for c, v in sorted(tots.items(), key=lambda x: -x[1])[:9]))
mix = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(mix.splitlines()))
hi = [i for i, r in enumerate(rows) if r and r[0] == 'Line No']
h = rows[hi[0]]
ci, si = h.index('Instructions Executed'), h.index('# Samples')
fn = [i for i, r in enumerate(rows) if r and r[0] == 'Function Name']
end = fn[which + 1] if len(fn) > which + 1 else len(rows)
start = fn[which] - 1   # the 'File Path' row precedes 'Function Name'
cur, lines = None, []
for r in rows[start:end]:
    if r and r[0] == 'File Path':
        cur = r[1]; continue
    if len(r) > ci and r[0].isdigit() and r[ci].isdigit() and cur and cur.endswith('.cu'):
        lines.append(r)
tot = sum(int(r[ci]) for r in lines) or 1
print("hottest source lines (% of warp instructions, stall samples):")
for r in sorted(sorted(lines, key=lambda r: -int(r[ci]))[:32], key=lambda r: int(r[0])):
    print("  %4s %5.1f%% smp %6s | %s" % (r[0], 100 * int(r[ci]) / tot, r[si], r[1].strip()[:105]))
