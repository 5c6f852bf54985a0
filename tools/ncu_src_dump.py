"""Every SASS line of a `ncu --page source --csv --print-source sass` export with its stall samples (all / long scoreboard) and
execution count — a compact text a GPU visit can bring back instead of the report.   python tools/ncu_src_dump.py file.csv"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
h0 = next(i for i, r in enumerate(rows) if r and r[0] == 'Address')
hdr = rows[h0]
ca, cl, ce = hdr.index('Warp Stall Sampling (All Samples)'), hdr.index('stall_long_sb'), hdr.index('Instructions Executed')
for i, r in enumerate(rows[h0 + 1:]):
    print(i, r[0][-5:], r[1][:100], '|', r[ca], r[cl], r[ce])
