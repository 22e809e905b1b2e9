"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list: per-kernel count / mean / share."""
import csv, collections, sys
for path in sys.argv[1:]:
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr = rows[0]; ki = hdr.index('Kernel Name'); vi = hdr.index('Metric Value')
    agg = collections.OrderedDict()
    for r in rows[1:]:
        try: agg.setdefault(r[ki].split('(')[0][-48:], []).append(float(r[vi].replace(',', '')))
        except ValueError: pass
    tot = sum(sum(v) for v in agg.values())
    print(f"# {path}: {sum(len(v) for v in agg.values())} launches, {tot/1e3:.1f} us total")
    for k, v in agg.items():
        print(f"{k:50s} n={len(v):4d} mean={sum(v)/len(v)/1e3:9.2f} us  min={min(v)/1e3:9.2f}  max={max(v)/1e3:9.2f}  share={100*sum(v)/tot:5.1f}%")
