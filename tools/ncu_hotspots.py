"""Compact per-kernel summary of an ncu report with source counters: run ON the GPU box, writes small text files.
    python tools/ncu_hotspots.py report.ncu-rep outdir"""
import csv
import io
import subprocess
import sys

rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
open(f"{out}/others_raw.csv", "w").write(raw)
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
blocks, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "rows": []}
        blocks.append(cur)
    elif r and r[0] == "Address":
        cur["hdr"] = r
    elif cur is not None and len(r) > 5:
        cur["rows"].append(r)
seen = set()
with open(f"{out}/others_hotspots.txt", "w") as f:
    for b in blocks:
        key = b["name"].split("(")[0]
        if key in seen or "hdr" not in b:
            continue
        seen.add(key)
        h = b["hdr"]
        iS, iSrc, iEx = h.index("# Samples"), h.index("Source"), h.index("Instructions Executed")
        cols = [c for c in h if c.startswith("stall_") and "Not Issued" not in c]
        tot = sum(int(r[iS]) for r in b["rows"]) or 1
        ex = sum(int(r[iEx]) for r in b["rows"])
        f.write(f"\n=== {b['name'][:140]}\n    samples {tot}, warp instructions executed {ex}, SASS lines {len(b['rows'])}\n")
        # opcode histogram by executed instructions
        ops = {}
        for r in b["rows"]:
            op = r[iSrc].strip().split()[0] if r[iSrc].strip() else "?"
            if op.startswith("@"):
                op = r[iSrc].strip().split()[1]
            op = op.split(".")[0]
            ops[op] = ops.get(op, 0) + int(r[iEx])
        f.write("    executed by opcode: " + ", ".join(f"{k} {100 * v / max(ex, 1):.1f}%" for k, v in sorted(ops.items(), key=lambda kv: -kv[1])[:14]) + "\n")
        for r in sorted(b["rows"], key=lambda r: -int(r[iS]))[:22]:
            st = sorted(((int(r[h.index(c)]), c) for c in cols), reverse=True)[:2]
            f.write(f"    {100 * int(r[iS]) / tot:5.1f}%  ex {int(r[iEx]):>11d}  {r[iSrc].strip()[:64]:64s} {st[0][1]}:{st[0][0]} {st[1][1]}:{st[1][0]}\n")
print(open(f"{out}/others_hotspots.txt").read()[:3000])
