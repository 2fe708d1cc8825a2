"""Per-source-line share of executed warp instructions from an .ncu-rep: python scripts/ncu_lines.py <rep> [kernel#] [min%]"""
import csv, subprocess, sys
rep = sys.argv[1]; which = int(sys.argv[2]) if len(sys.argv) > 2 else 0; thr = float(sys.argv[3]) if len(sys.argv) > 3 else 0.15
mix = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(mix.splitlines()))
hi = [i for i, r in enumerate(rows) if r and r[0] == 'Line No']
h = rows[hi[0]]
ci, si, ti = h.index('Instructions Executed'), h.index('# Samples'), h.index('Thread Instructions Executed')
fn = [i for i, r in enumerate(rows) if r and r[0] == 'Function Name']
end = fn[which + 1] - 1 if len(fn) > which + 1 else len(rows)
lines = [r for r in rows[fn[which]:end] if len(r) > ci and r[0].isdigit() and r[ci].isdigit()]
tot = sum(int(r[ci]) for r in lines) or 1
stot = sum(int(r[si]) for r in lines) or 1
print("total warp instr %d, samples %d" % (tot, stot))
for r in lines:
    p = 100 * int(r[ci]) / tot
    if p >= thr:
        print("%4s %5.2f%% smp %5.2f%% lanes %4.1f | %s" % (r[0], p, 100 * int(r[si]) / stot, int(r[ti]) / max(1, int(r[ci])), r[1].strip()[:110]))
