"""diagnostic: run a brittle example on the GPU drop-in and locate the first record / broken-bond line that differs from the serial golden"""
import os, subprocess, sys, tempfile
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parents[1]
name, gold = (sys.argv[1], sys.argv[2]) if len(sys.argv) > 2 else ("bending_sq", "c3_bending_sq")
d = Path(tempfile.mkdtemp())
with open(d / "run.log", "w") as log:
    try:
        subprocess.run([str(ROOT / "oracle/_ref" / f"{name}_b200")], cwd=d, stdout=log, stderr=subprocess.STDOUT, timeout=150)
    except subprocess.TimeoutExpired:
        pass
def table(p):
    rows = []
    for ln in Path(p).read_text().splitlines():
        try: rows.append([float(x) for x in ln.split()])
        except ValueError: pass
    return rows
G = ROOT / "tests/golden"
for what in ("force", "disp"):
    a, b = table(d / f"result_{what}.txt"), table(G / f"{gold}_result_{what}.txt")
    n = min(len(a), len(b))
    rel = [abs(a[k][1] - b[k][1]) / max(abs(b[k][1]), 1e-30) for k in range(n)]
    bad = [k for k in range(n) if rel[k] > 2e-8]
    print(what, "records", len(a), len(b), "first differing record", bad[:5], "values", [(a[k], b[k]) for k in bad[:3]], "max rel", max(rel))
    big = [k for k in range(n) if rel[k] > 1e-7]
    print(what, "records above 1e-7:", [(k, a[k], b[k], f"{rel[k]:.2e}") for k in big[:40]])
ba = [l.strip() for l in (d / "result_brokenbonds.txt").read_text().split("\n") if l.strip()]
bb = [l.strip() for l in (G / f"{gold}_result_brokenbonds.txt").read_text().split("\n") if l.strip()]
k = 0
while k < min(len(ba), len(bb)) and ba[k] == bb[k]:
    k += 1
print("broken-bond logs: lines", len(ba), len(bb), "first differing line", k, ba[max(0,k-3):k+4], "|", bb[max(0,k-3):k+4])
log = (d / "run.log").read_text()
import re
gn = [int(x) for x in re.findall(r"has finished in (\d+) iterations", log)]
print("GPU newton iterations per finished step:", " ".join(map(str, gn)))
print("GPU newton passes", len(re.findall(r"has finished in", log)), "golden summary:", (G / f"{gold}_log_summary.txt").read_text()[:200])
