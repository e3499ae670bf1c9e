"""Per-source-line warp-stall samples of one kernel of an .ncu-rep (needs -lineinfo + --import-source on).
    python scripts/ncu_lines.py report.ncu-rep [launch_index] [top]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
idx = int(sys.argv[2]) if len(sys.argv) > 2 else 0
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
# the report is a sequence of blocks: "File Path", "Function Name", header ("Line No", ...), data rows
blocks, cur = [], None
for r in rows:
    if r and r[0] == "File Path":
        cur = {"file": r[1], "rows": [], "fn": "?"}
        blocks.append(cur)
    elif r and r[0] == "Function Name" and cur is not None:
        cur["fn"] = r[1]
    elif r and r[0] == "Line No" and cur is not None and len(r) > 4:
        cur["hdr"] = r
    elif cur is not None and "hdr" in cur and r and r[0].isdigit() and len(r) == len(cur["hdr"]) and r[2] == "-":
        cur["rows"].append(r)
print(f"# {rep}: {len(blocks)} kernel launch(es) with source; showing #{idx}")
blocks = [b for b in blocks if "hdr" in b]
b = blocks[idx]
h = b["hdr"]
i_s = h.index("# Samples")
i_ins = h.index("Instructions Executed")
stall_cols = [(i, c) for i, c in enumerate(h) if c.startswith("stall_") and "Not Issued" not in c]
tot = sum(int(r[i_s] or 0) for r in b["rows"])
print(f"# {b['fn'][:90]}  total samples {tot}")
for r in sorted(b["rows"], key=lambda r: -int(r[i_s] or 0))[:top]:
    n = int(r[i_s] or 0)
    st = sorted(((int(r[i] or 0), c) for i, c in stall_cols), reverse=True)[:3]
    print(f"{100.0 * n / max(tot, 1):5.1f}%  L{r[0]:>4}  inst={r[i_ins]:>9}  {', '.join(f'{c[6:]}={v}' for v, c in st if v)}  | {r[1].strip()[:100]}")
