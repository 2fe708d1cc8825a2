"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel: python scripts/launch_summary.py <csv>"""
import collections, csv, re, sys
rows = [r for r in csv.reader(open(sys.argv[1], errors="replace")) if len(r) > 14 and r[0].isdigit()]
tot, cnt = collections.Counter(), collections.Counter()
for r in rows:
    name = re.sub(r"\(.*", "", r[4]).strip()
    tot[name] += int(float(r[14])); cnt[name] += 1
allns = sum(tot.values()) or 1
print("# per kernel: launches, total ns, share of all captured launches, average")
for name, ns in tot.most_common():
    print("%-78s %5d %14d ns %6.1f%%  avg %12d ns" % (name[:78], cnt[name], ns, 100.0 * ns / allns, ns // cnt[name]))
