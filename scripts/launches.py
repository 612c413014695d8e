"""Print the kernels of an `ncu --metrics gpu__time_duration.sum --csv` log whose name matches argv[2] (default: all of ours)."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
h = rows[hi]
pat = sys.argv[2] if len(sys.argv) > 2 else 'lgcn::'
for r in rows[hi + 1:]:
    d = dict(zip(h, r))
    if d.get('Metric Name') == 'gpu__time_duration.sum' and pat in d['Kernel Name']:
        print(f"{d['Kernel Name'][:70]:70s} {d['Grid Size']:>16s} {float(d['Metric Value'])/1000:10.1f} us")
