"""profiles/r02_sor_traffic.json from an `ncu --set full` capture of the solver launches of one bench-sized step:
    ncu -i gpurun_out/.../prof_sor.ncu-rep --page raw --csv > raw.csv ; python tools/ncu_traffic.py raw.csv f32 25
DRAM bytes (read + write) of the captured launches per frame and level voxel -- what bench.py multiplies back up for
`roofline.traffic` (only when kernel and solver state match the benched ones)."""
import csv
import json
import sys
from pathlib import Path

raw, state, B = sys.argv[1], sys.argv[2], int(sys.argv[3])
levels = [8 * 134 * 134, 10 * 168 * 168]
rows = list(csv.reader(open(raw)))
hdr, units = rows[0], rows[1]
col = {h: i for i, h in enumerate(hdr)}
out = []
for r in rows[2:]:
    name = r[col["Kernel Name"]]
    if "fr3d_sor" not in name:
        continue

    def val(k):
        v = float(r[col[k]].replace(",", ""))
        u = units[col[k]].lower()
        return v * {"gbyte": 1e9, "mbyte": 1e6, "kbyte": 1e3, "byte": 1.0, "tbyte": 1e12}.get(u, 1.0)
    out.append({"kernel": name, "ms": float(r[col["gpu__time_duration.sum"]].replace(",", "")) *
                {"ms": 1.0, "us": 1e-3, "ns": 1e-6, "s": 1e3}[units[col["gpu__time_duration.sum"]].lower()],
                "dram_read": val("dram__bytes_read.sum"), "dram_write": val("dram__bytes_write.sum")})
assert len(out) == len(levels), [o["kernel"] for o in out]
tot = sum(o["dram_read"] + o["dram_write"] for o in out)
kern = "fr3d_sor_tiles" if "tiles" in out[0]["kernel"] else "fr3d_sor_staged" if "staged" in out[0]["kernel"] else "fr3d_sor_wavefront"
res = {"kernel": kern, "state": state, "B": B, "launches": out,
       "dram_bytes_per_frame_voxel": tot / (B * sum(levels)),
       "source": f"ncu --set full of this build, B = {B}, {state} state: dram__bytes_read.sum + dram__bytes_write.sum of the "
                 f"two solver launches of one step ({raw})"}
out_file = Path(__file__).resolve().parent.parent.joinpath("profiles", "r02_sor_traffic.json")
allres = json.loads(out_file.read_text()) if out_file.exists() else {}
if "kernel" in allres:      # old single-entry layout
    allres = {}
allres[state] = res
out_file.write_text(json.dumps(allres, indent=1))
print(json.dumps(res)[:600])
