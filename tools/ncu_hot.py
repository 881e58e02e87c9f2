"""Top stalled SASS instructions of a kernel from an .ncu-rep source page (needs -lineinfo + --import-source on).
    python tools/ncu_hot.py rep.ncu-rep [N]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
# first line: kernel name; second: header
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hdr_i]
ix = {h: i for i, h in enumerate(hdr)}
body = [r for r in rows[hdr_i + 1:] if len(r) == len(hdr)]
tot = sum(int(r[ix["# Samples"]] or 0) for r in body)
print(f"{rep}: {len(body)} SASS lines, {tot} samples")
stall_cols = [h for h in hdr if h.startswith("stall_") and "(Not Issued)" not in h]
for k, r in sorted(enumerate(body), key=lambda kr: -int(kr[1][ix["# Samples"]] or 0))[:n]:
    s = int(r[ix["# Samples"]] or 0)
    top = sorted(((int(r[ix[c]] or 0), c) for c in stall_cols), reverse=True)[:2]
    print(f"{k:5d} {100.0 * s / tot:5.1f}%  {r[ix['Source']].strip()[:70]:70s} {' '.join(f'{c[6:]}={v}' for v, c in top)}")
