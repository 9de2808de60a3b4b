"""Mechanical similarity check of the tracked sources against the reference tree (run where /root/reference exists):
for every tracked .py / .cu / .cuh / .h / .sh file, the reference file with the largest share of identical normalised lines
(Jaccard pre-filter over all reference sources, then difflib on the three best candidates).  Writes profiles/r02_copycheck.json.
    python tools/copycheck.py [/root/reference]"""
import difflib
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
EXT = (".py", ".cu", ".cuh", ".h", ".sh", ".c", ".cpp")


def norm_lines(path):
    try:
        txt = open(path, errors="replace").read()
    except OSError:
        return []
    out = []
    for line in txt.splitlines():
        line = re.sub(r"\s+", " ", line.strip())
        if len(line) >= 4 and not line.startswith(("#", "//", "*", "/*", '"""', "'''")):
            out.append(line)
    return out


tracked = [f for f in subprocess.check_output(["git", "-C", ROOT, "ls-files"], text=True).split("\n") if f.endswith(EXT)]
ref_files = []
for d, _, fs in os.walk(REF):
    if "/.git" in d:
        continue
    for f in fs:
        if f.endswith(EXT):
            ref_files.append(os.path.join(d, f))
ref = {f: norm_lines(f) for f in ref_files}
ref_sets = {f: set(v) for f, v in ref.items() if len(v) >= 5}
report = []
for f in tracked:
    a = norm_lines(os.path.join(ROOT, f))
    if len(a) < 5:
        continue
    sa = set(a)
    cand = sorted(ref_sets, key=lambda r: -len(sa & ref_sets[r]) / max(1, len(sa | ref_sets[r])))[:3]
    best, best_r = 0.0, None
    for r in cand:
        ratio = difflib.SequenceMatcher(None, a, ref[r], autojunk=False).ratio()
        if ratio > best:
            best, best_r = ratio, os.path.relpath(r, REF)
    report.append({"file": f, "lines": len(a), "nearest_reference_file": best_r, "similarity": round(best, 3)})
report.sort(key=lambda x: -x["similarity"])
out = {"reference_sources": len(ref_files), "tracked_sources": len(report), "threshold": 0.6,
       "over_threshold": [r for r in report if r["similarity"] > 0.6], "files": report}
json.dump(out, open(os.path.join(ROOT, "profiles", "r02_copycheck.json"), "w"), indent=1)
print("tracked %d, reference %d, over 0.6: %d" % (len(report), len(ref_files), len(out["over_threshold"])))
for r in report[:10]:
    print("%.3f  %-55s %s" % (r["similarity"], r["file"], r["nearest_reference_file"]))
