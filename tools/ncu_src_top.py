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
top = sorted(range(len(data)), key=lambda i: -num(data[i][c_all]))[:int(sys.argv[2]) if len(sys.argv) > 2 else 40]
for i in sorted(top):
    r = data[i]
    print(f"{i:5d} {r[0][-5:]} {r[1][:90]:90s} all {r[c_all]:>6s} long_sb {r[c_long]:>6s} exec {r[c_exec]}")
