"""Attribute warp-state samples of an ncu source page (csv) to code regions delimited by execution-count changes."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[2:] if len(r) == len(hdr)]
def f(r, k):
    try: return float(r[ix[k]])
    except Exception: return 0.0
tot = sum(f(r, "# Samples") for r in data)
reasons = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
# split into chunks of N instructions
N = int(sys.argv[2]) if len(sys.argv) > 2 else 64
print(f"total samples {int(tot)}")
for i in range(0, len(data), N):
    ch = data[i:i + N]
    s = sum(f(r, "# Samples") for r in ch)
    ex = sum(f(r, "Instructions Executed") for r in ch)
    if s < tot * 0.002: continue
    top = sorted(((k[6:], sum(f(r, k) for r in ch)) for k in reasons), key=lambda x: -x[1])[:4]
    addr = ch[0][ix["Address"]][-5:]
    ops = {}
    for r in ch:
        op = r[ix["Source"]].split()[0] if r[ix["Source"]].split() else "?"
        if op.startswith("@"): op = r[ix["Source"]].split()[1]
        ops[op] = ops.get(op, 0) + 1
    topops = ",".join(f"{k}:{v}" for k, v in sorted(ops.items(), key=lambda x: -x[1])[:4])
    print(f"{addr} {s/tot*100:5.1f}%  exec {ex/1e6:7.2f}M  " + " ".join(f"{k}:{v/max(s,1)*100:.0f}%" for k, v in top) + "  | " + topops)
