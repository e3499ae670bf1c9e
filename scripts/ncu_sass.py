"""SASS-level warp-stall attribution of one kernel of an .ncu-rep: marker instructions (TMA / MMA / TMEM / barriers) and every
instruction above a sample share, with its top stall reasons.   python scripts/ncu_sass.py rep.ncu-rep <kernel-substring> [min_pct]"""
import collections
import csv
import subprocess
import sys

rep, pat = sys.argv[1], sys.argv[2]
min_pct = float(sys.argv[3]) if len(sys.argv) > 3 else 0.8
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr_i = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
seen = set()
for bi, hi in enumerate(hdr_i):
    name = rows[hi - 1][1] if rows[hi - 1][0] == "Kernel Name" else "?"
    if pat not in name or name in seen:
        continue
    seen.add(name)
    h = rows[hi]
    end = hdr_i[bi + 1] - 2 if bi + 1 < len(hdr_i) else len(rows)
    data = [r for r in rows[hi + 1:end] if len(r) == len(h)]
    i_s, i_src = h.index("# Samples"), h.index("Source")
    stall = [(i, c[6:]) for i, c in enumerate(h) if c.startswith("stall_") and "Not Issued" not in c]
    tot = sum(int(r[i_s] or 0) for r in data)
    print(f"# {name[:100]}  samples {tot}")
    agg = collections.Counter()
    for k, r in enumerate(data):
        n = int(r[i_s] or 0)
        for i, c in stall:
            agg[c] += int(r[i] or 0)
        src = r[i_src]
        mark = any(t in src for t in ("UTMALDG", "UTCHMMA", "UTCBAR", "LDTM", "STTM", "BAR.SYNC", "SYNCS.ARRIVE", "SYNCS.PHASECHK"))
        if n > tot * min_pct / 100 or mark:
            st = sorted(((int(r[i] or 0), c) for i, c in stall), reverse=True)[:2]
            print(f"{k:5d} {100 * n / max(tot, 1):5.1f}% {src.strip()[:72]:72s} {[f'{c}={v}' for v, c in st if v]}")
    print("# stall totals:", agg.most_common(8))
