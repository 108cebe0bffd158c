"""Hot SASS lines of one .ncu-rep (source page): samples, executed count and dominant stall per line.
    python tools/ncu_hot.py gpurun_out/x.ncu-rep [min_share]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
share = float(sys.argv[2]) if len(sys.argv) > 2 else 0.012
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
h, body = rows[1], rows[2:]
ia, isamp, iex = h.index('Source'), h.index('# Samples'), h.index('Instructions Executed')
cols = [c for c in h if c.startswith('stall_') and 'Not Issued' not in c]
tot = sum(int(r[isamp] or 0) for r in body)
totx = sum(int(r[iex] or 0) for r in body)
print(f'{len(body)} SASS lines, {tot} samples, {totx} warp instructions')
for s in range(0, len(body), 50):
    ex = sum(int(r[iex] or 0) for r in body[s:s + 50])
    sm = sum(int(r[isamp] or 0) for r in body[s:s + 50])
    if sm > tot * 0.02:
        print(f'  lines {s:5d}+50: {ex / 1e6:8.2f} M instr {sm:6d} samples')
for k, r in enumerate(body):
    s = int(r[isamp] or 0)
    if s > tot * share:
        st = sorted(((int(r[h.index(c)] or 0), c) for c in cols), reverse=True)[:2]
        print(k, r[ia][:72].ljust(72), s, r[iex], st)
