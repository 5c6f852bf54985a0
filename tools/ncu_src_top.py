"""Top SASS lines by warp-stall samples from `ncu -i X.ncu-rep --page source --csv --print-source sass` (so that a GPU visit
can leave a text summary instead of the report).   python tools/ncu_src_top.py file.csv [n]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
h0 = next(i for i, r in enumerate(rows) if r and r[0] == 'Address')
hdr = rows[h0]
data = rows[h0 + 1:]
c_all = hdr.index('Warp Stall Sampling (All Samples)')
c_long = hdr.index('stall_long_sb')
c_exec = hdr.index('Instructions Executed')


def num(x):
    try:
        return float(x.replace(',', ''))
    except ValueError:
        return 0.0


print(rows[0][1] if rows and len(rows[0]) > 1 else '')
print('instructions', len(data), 'samples', sum(num(r[c_all]) for r in data), 'long_sb', sum(num(r[c_long]) for r in data))
stall_cols = [i for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
tot = {hdr[i]: sum(num(r[i]) for r in data) for i in stall_cols}
print('by reason:', ', '.join(f"{k[6:]} {int(v)}" for k, v in sorted(tot.items(), key=lambda kv: -kv[1]) if v))
# the kernel in chunks of 256 instructions: where the samples are
for c0 in range(0, len(data), 256):
    chunk = data[c0:c0 + 256]
    n = sum(num(r[c_all]) for r in chunk)
    if n >= 0.01 * max(1.0, sum(num(r[c_all]) for r in data)):
        best = max(stall_cols, key=lambda i: sum(num(r[i]) for r in chunk))
        print(f"  instr {c0:6d}-{c0 + len(chunk) - 1:6d}: {int(n):6d} samples, executed {chunk[0][c_exec]:>9s}.., most: {hdr[best][6:]} {int(sum(num(r[best]) for r in chunk))}")
top = sorted(range(len(data)), key=lambda i: -num(data[i][c_all]))[:int(sys.argv[2]) if len(sys.argv) > 2 else 40]
for i in sorted(top):
    r = data[i]
    print(f"{i:5d} {r[0][-5:]} {r[1][:90]:90s} all {r[c_all]:>6s} long_sb {r[c_long]:>6s} exec {r[c_exec]}")
