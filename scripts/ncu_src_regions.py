"""Summarise an `ncu --page source --csv` dump into runs of SASS with (nearly) equal execution counts."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
which = int(sys.argv[2]) if len(sys.argv) > 2 else 0
hdr_idx = [i for i, r in enumerate(rows) if r and r[0] == 'Address']
start = hdr_idx[which]; end = hdr_idx[which + 1] - 1 if len(hdr_idx) > which + 1 else len(rows)
print(rows[start - 1][:2])
hdr = rows[start]; body = rows[start + 1:end]
ie = hdr.index('Instructions Executed'); si = hdr.index('# Samples'); src = hdr.index('Source')
tot = sum(int(r[ie]) for r in body if r[ie].isdigit())
print('total warp instr', tot, 'n sass', len(body))
run, prev = [], None
def flush():
    if run:
        c = int(run[0][ie]); n = len(run); s = sum(int(r[si]) for r in run)
        tot_run = sum(int(r[ie]) for r in run)
        if tot_run > 0.004 * tot:
            print(f"{run[0][0][-5:]}..{run[-1][0][-5:]} n_sass={n:4d} exec~{c:8d} sum={tot_run:9d} ({100*tot_run/tot:4.1f}%) samples={s:5d}  {run[0][src].strip()[:70]}")
for r in body:
    if not r[ie].isdigit(): continue
    if prev is not None and abs(int(r[ie]) - prev) > max(50, 0.05 * prev):
        flush(); run = []
    run.append(r); prev = int(r[ie])
flush()
