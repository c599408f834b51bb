"""Turn an `ncu --page raw --csv` export of ONE k_decode_persistent launch into profiles/r2_ncu_traffic.json:
DRAM bytes (dram__bytes_read.sum + dram__bytes_write.sum) per decode position, which bench.py reports as
roofline.traffic (x positions per launch).  Usage: python tools/ncu_traffic.py raw.csv n_positions [out.json]"""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
path, n_pos = sys.argv[1], int(sys.argv[2])
out = sys.argv[3] if len(sys.argv) > 3 else os.path.join(ROOT, "profiles", "r2_ncu_traffic.json")
rows = list(csv.reader(open(path)))
hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
names, units = rows[hdr], rows[hdr + 1]
rec = next(r for r in rows[hdr + 2:] if any("k_decode_persistent" in c for c in r))
col = {n: i for i, n in enumerate(names)}


def val(metric):
    v, u = float(rec[col[metric]].replace(",", "")), units[col[metric]].lower()
    scale = {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "tbyte": 1e12}
    return v * scale.get(u, 1)


rd, wr = val("dram__bytes_read.sum"), val("dram__bytes_write.sum")
res = {"dram_bytes_per_position": (rd + wr) / n_pos, "dram_read_bytes": rd, "dram_write_bytes": wr, "positions": n_pos,
       "duration": rec[col["gpu__time_duration.sum"]] + " " + units[col["gpu__time_duration.sum"]] if "gpu__time_duration.sum" in col else None,
       "source": f"ncu --set full --clock-control none, one k_decode_persistent launch of {n_pos} positions at context ~50 (tools/prof_decode.py 1 {n_pos})"}
json.dump(res, open(out, "w"), indent=1)
print(json.dumps(res))
