"""Static check of a kernel's SASS for prefetch loads that are waited on right after they are issued.

ptxas encodes, per instruction, the scoreboard it sets (write barrier) and the scoreboards it waits for; a scoreboard is a
counter, so a wait is a wait for EVERYTHING outstanding on it.  A global load issued as a PREFETCH for the next loop iteration
must not be followed, a few instructions later, by an instruction that waits on the same scoreboard without consuming one of
the loads — that exposes a full DRAM round trip per iteration.  Round 2: the M=32 kernel with K-side outlier records went from
120 to 165 us per launch when a change in an unrelated function made ptxas restore a register after the prefetch with such a wait
(profiles/r02_dm4_scoreboard_regression.txt).

    python tools/sass_sb_check.py libmillion_b200.so '<substring of the mangled kernel name>' [max_distance=64]

Lists every LDG inside a loop whose scoreboard is waited on within `max_distance` instructions of straight-line code by an
instruction that reads none of the registers loaded on that scoreboard; exit status 1 if there is one.  Branch structure is only
followed as far as "the path that issued a predicated load leaves at a branch with the same predicate": a conservative listing
to read, not a proof.
"""
import re
import subprocess
import sys


def decode(text):
    """[(addr, text, write_barrier, read_barrier, wait_mask)] from `cuobjdump -sass` output (two lines per instruction; the
    control bits are in the upper word: stall 41-44, yield 45, write barrier 46-48, read barrier 49-51, wait mask 52-57)."""
    lines = text.split('\n')
    ins = []
    i = 0
    while i < len(lines):
        m = re.match(r'\s+/\*([0-9a-f]{4,6})\*/\s+(.*?);\s*/\* (0x[0-9a-f]+) \*/', lines[i])
        if m and i + 1 < len(lines):
            m2 = re.match(r'\s+/\* (0x[0-9a-f]+) \*/', lines[i + 1])
            if m2:
                hi = int(m2.group(1), 16)
                ins.append({"addr": int(m.group(1), 16), "txt": re.sub(r'\s+', ' ', m.group(2).strip()),
                            "wr": (hi >> 46) & 7, "rd": (hi >> 49) & 7, "wait": (hi >> 52) & 0x3f})
                i += 2
                continue
        i += 1
    return ins


def kernels(lib, name):
    names = subprocess.run(["cuobjdump", "-elf", lib], capture_output=True, text=True).stdout
    found = set(re.findall(r'(_Z\w*' + re.escape(name) + r'\w*)', names))
    return sorted(c for c in found if '_param_' not in c)


def _load_dsts(txt):
    m = re.search(r'LDG[.\w]* (R\d+)', txt)
    if not m:
        return set()
    r = int(m.group(1)[1:])
    return {f"R{r}", f"R{r + 1}"} if '.64' in txt else ({f"R{r + i}" for i in range(4)} if '.128' in txt else {f"R{r}"})


def hazards(lib, fn, maxd=64, loads=r'LDG\.E(\.U8|\.U16|\.64)?\.CONSTANT'):
    """Loop-resident loads matching `loads` whose scoreboard is waited on within maxd instructions by a non-consumer."""
    text = subprocess.run(["cuobjdump", "-sass", "-fun", fn, lib], capture_output=True, text=True).stdout
    ins = decode(text)
    loops = []
    for k in ins:
        m = re.search(r'BRA(?:\.\w+)* (?:\w+, )?0x([0-9a-f]+)', k["txt"])
        if m and int(m.group(1), 16) < k["addr"]:
            loops.append((int(m.group(1), 16), k["addr"]))
    out = []
    for j, k in enumerate(ins):
        if not re.search(loads, k["txt"]) or k["wr"] == 7:
            continue
        if not any(lo <= k["addr"] <= hi and hi - lo > 0x800 for lo, hi in loops):
            continue
        pred = re.match(r'(@!?U?P\d+) ', k["txt"])
        pred = pred.group(1) if pred else None
        dsts = set()
        for b in range(max(0, j - 16), j):                     # loads issued just before on the same scoreboard
            if ins[b]["wr"] == k["wr"]:
                dsts |= _load_dsts(ins[b]["txt"])
        for d in range(0, maxd + 1):
            if j + d >= len(ins):
                break
            n = ins[j + d]
            if n["wr"] == k["wr"]:
                dsts |= _load_dsts(n["txt"])
            if d == 0:
                continue
            b = re.match(r'(@!?U?P\d+ )?(BRA|EXIT|RET|CALL)', n["txt"])
            if b and (b.group(1) is None or (pred and b.group(1).strip() == pred)):
                break                                          # the path that issued the load leaves the straight line here
            if n["wait"] >> k["wr"] & 1:
                srcs = n["txt"].split(',', 1)[1] if ',' in n["txt"] else ''
                if not any(re.search(r'\b' + r + r'\b', srcs) for r in dsts) and 'LDG' not in n["txt"]:
                    out.append((k["addr"], k["txt"], d, n["txt"]))
                break
    return len(ins), out


def main():
    lib, name = sys.argv[1], sys.argv[2]
    maxd = int(sys.argv[3]) if len(sys.argv) > 3 else 64
    fns = kernels(lib, name)
    if not fns:
        print("no kernel matches", name)
        return 2
    bad = 0
    for fn in fns:
        n, hz = hazards(lib, fn, maxd)
        print(f"{fn}: {n} instructions, {len(hz)} early waits")
        for addr, txt, d, ntxt in hz:
            print(f"  {addr:06x} {txt[:72]:72s} -> +{d} {ntxt[:60]}")
        bad += len(hz)
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
