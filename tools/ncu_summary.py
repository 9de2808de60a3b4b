"""Transposes `ncu -i <rep> --page raw --csv` into one `metric,value,unit` row per metric (what profiles/*_ncu_summary.csv
hold and bench.py reads the DRAM traffic from).   python tools/ncu_summary.py gpurun_out/x.ncu-rep profiles/x_ncu_summary.csv"""
import csv
import subprocess
import sys

rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
with open(out, "w", newline="") as f:
    w = csv.writer(f)
    for k, vals in enumerate(rows[2:]):
        if k:
            w.writerow(["--- launch %d ---" % k, "", ""])
        for h, v, u in zip(hdr, vals, units):
            w.writerow([h, v, u])
print("wrote", out, "(%d launches, %d metrics)" % (len(rows) - 2, len(hdr)))
