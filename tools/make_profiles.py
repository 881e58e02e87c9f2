"""Turns the ncu artefacts a `tools/profile_gpu.sh <tag>` run left in gpurun_out/ into the committed evidence under profiles/:
    profiles/<tag>_launches_dit_step.txt    per-kernel device time of ONE eager denoise step (ncu launch list, shares of the step)
    profiles/<tag>_launches_decoder.txt     same for one decoder forward
    profiles/<tag>_ncu_<kernel>.txt         --set full summary (tensor / MUFU / DRAM / stalls) + the hottest SASS lines
    profiles/ncu_traffic.json               dram bytes (read + write) per launch of the dominant kernels (bench.py roofline.traffic)
Run here (no GPU needed):  python tools/make_profiles.py <tag>
"""
import collections
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "profiles")
GP = os.path.join(ROOT, "gpurun_out")


def launches(path, title):
    rows = list(csv.reader(open(path)))
    h = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    hdr = rows[h]
    ix = {k: i for i, k in enumerate(hdr)}
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows[h + 1:]:
        if len(r) < len(hdr):
            continue
        name = r[ix["Kernel Name"]].split("(")[0].replace("void ", "")[:80]
        if r[ix["Metric Name"]] != "gpu__time_duration.sum":
            continue
        v = float(r[ix["Metric Value"]].replace(",", ""))
        u = r[ix["Metric Unit"]]
        v = v / 1e3 if u == "ns" else v * 1e3 if u == "ms" else v
        agg[name][0] += 1
        agg[name][1] += v
    tot = sum(v[1] for v in agg.values())
    lines = [f"# {title}", f"# source: {os.path.basename(path)} (ncu --metrics gpu__time_duration.sum --clock-control none; cold-cache, serialised: compare SHARES)",
             f"# {sum(v[0] for v in agg.values())} launches, {tot / 1e3:.3f} ms summed device time", f"{'kernel':82s} {'launches':>8s} {'us':>12s} {'share':>7s}"]
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        lines.append(f"{k:82s} {v[0]:8d} {v[1]:12.1f} {100 * v[1] / tot:6.1f}%")
    return "\n".join(lines) + "\n"


def ncu_text(rep):
    a = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_summary.py"), rep], capture_output=True, text=True).stdout
    b = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_hot.py"), rep, "25"], capture_output=True, text=True).stdout
    return a + "\n# hottest SASS lines (warp-state samples)\n" + b


def traffic(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    vals = []
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        u = dict(zip(hdr, units))
        tot = 0.0
        for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u[k]]
            tot += float(d[k]) * scale
        vals.append(tot)
    return sum(vals) / len(vals)


def main():
    tag = sys.argv[1]
    os.makedirs(OUT, exist_ok=True)
    for stem, title in (("launches", "one eager denoise step of bench.py (cond+uncond DiT forward at L=4096, CFG, UniPC)"),
                        ("launches_decoder", "one stitched-decoder forward (13 views x 448x448)"),
                        ("launches_voxel", "one voxelised fusion of 2 609 152 points x 83 features (32 launches; radix passes beyond the key width exit at once)")):
        p = os.path.join(GP, f"{stem}_{tag}.csv")
        if os.path.exists(p):
            name = "launches_dit_step" if stem == "launches" else stem
            open(os.path.join(OUT, f"{tag}_{name}.txt"), "w").write(launches(p, title))
    tr = {}
    for stem, key in (("fmha", "fmha_tcgen05"), ("fmha64", None), ("gemm", "gemm_tcgen05"), ("gemm_tf32conv", None), ("gauss", "gaussian_epilogue"), ("voxel", "voxel_reduce")):
        rep = os.path.join(GP, f"{stem}_{tag}.ncu-rep")
        if os.path.exists(rep):
            open(os.path.join(OUT, f"{tag}_ncu_{stem}.txt"), "w").write(ncu_text(rep))
            if key:
                tr[key] = traffic(rep)
    if tr:
        tr["_source"] = f"ncu --set full captures of round tag {tag} (tools/profile_gpu.sh); bytes per launch, dram read + write"
        json.dump(tr, open(os.path.join(OUT, "ncu_traffic.json"), "w"), indent=1)
    print(sorted(os.listdir(OUT)))


if __name__ == "__main__":
    main()
