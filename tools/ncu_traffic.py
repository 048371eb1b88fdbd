"""profiles/ncu_traffic.json from the raw CSV of an `ncu --set full` capture (dram__bytes_read.sum +
dram__bytes_write.sum per kernel launch; bench.py's roofline.traffic reads the file and rescales it to the images per
launch of its own run).   python tools/ncu_traffic.py raw.csv <frames per launch> <source note> > profiles/ncu_traffic.json"""
import csv
import json
import sys

NAMES = {"composite_kernel": "composite", "emit_scatter_kernel": "emit_scatter", "bind_preprocess_kernel": "bind_preprocess",
         "rs_onesweep_kernel": "depth_sort_pass", "face_frames_kernel": "face_frames", "flame_lbs_kernel": "flame_lbs",
         "flame_blend_tc": "flame_blend_tc", "png_strip_kernel": "png_strip", "tile_scan_kernel": "tile_scan"}
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}

rows = list(csv.reader(open(sys.argv[1])))
hdr = next(r for r in rows if "Kernel Name" in r)
units = rows[rows.index(hdr) + 1]
ki = hdr.index("Kernel Name")
ri, wi = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
out = {}
for r in rows[rows.index(hdr) + 2:]:
    if len(r) <= wi:
        continue
    key = next((v for k, v in NAMES.items() if k in r[ki]), None)
    if key is None or key in out:
        continue
    total = float(r[ri].replace(",", "")) * UNIT[units[ri]] + float(r[wi].replace(",", "")) * UNIT[units[wi]]
    out[key] = {"dram_bytes_per_launch": total, "frames_per_launch": int(sys.argv[2]), "source": sys.argv[3]}
print(json.dumps(out, indent=1))
