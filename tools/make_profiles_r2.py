"""Turns the ncu artefacts of `tools/profile_gpu_r2.sh <tag>` (gpurun_out/) into the committed evidence under profiles/:
    profiles/<tag>_launches_dit_step.txt / _decoder.txt / _vae.txt   per-kernel device time (ncu launch lists, shares)
    profiles/<tag>_ncu_targets.txt      --set full summary of every hot kernel at its BASELINE shape (tensor / MUFU / issue / DRAM / stalls)
    profiles/<tag>_ncu_targets_table.md one line per kernel: duration, tensor pipe %, MUFU %, issue %, DRAM GB/s and bytes
    profiles/ncu_traffic.json           DRAM bytes per launch of the dominant kernels (bench.py roofline.traffic)
Run here (no GPU needed):  python tools/make_profiles_r2.py <tag>"""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import make_profiles as MP  # noqa: E402

OUT, GP = MP.OUT, MP.GP


def num(d, u, key):
    v = d.get(key, "")
    if v == "":
        return None
    x = float(v.replace(",", ""))
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(u.get(key, ""), 1)
    return x * scale


def main():
    tag = sys.argv[1]
    for stem, name, title in (("launches", "launches_dit_step", "one eager denoise step of bench.py (cond+uncond DiT forward at L=4096, CFG, UniPC)"),
                              ("launches_decoder", "launches_decoder", "one stitched-decoder forward (13 views x 448x448)"),
                              ("launches_vae", "launches_vae", "one Wan VAE decode (latent [1,16,4,64,64] -> 13 x 512 x 512)")):
        p = os.path.join(GP, f"{stem}_{tag}.csv")
        if os.path.exists(p):
            open(os.path.join(OUT, f"{tag}_{name}.txt"), "w").write(MP.launches(p, title))
    raw = os.path.join(GP, f"targets_{tag}_raw.csv")
    if not os.path.exists(raw):
        return
    txt = open(os.path.join(GP, f"targets_{tag}_summary.txt")).read().replace("/tmp/", "")
    open(os.path.join(OUT, f"{tag}_ncu_targets.txt"), "w").write(txt)
    names = [l.strip() for l in open(os.path.join(GP, f"ncu_targets_{tag}.log")) if l.strip() and " " not in l.strip() and not l.startswith("=")]
    rows = list(csv.reader(open(raw)))
    hdr, units = rows[0], rows[1]
    u = dict(zip(hdr, units))
    lines = [f"# ncu --set full --clock-control none, one launch per kernel at its BASELINE shape (tools/ncu_targets.py); tag {tag}",
             "# duration under ncu (cold caches, serialised); pipe numbers are % of peak over the ACTIVE cycles; DRAM = read + write",
             "| target (launch order) | kernel | us | tensor % | MUFU % | FMA % | issue % | DRAM MB | DRAM GB/s | regs | waves/SM |", "|---|---|---|---|---|---|---|---|---|---|---|"]
    traffic = {}
    ordered = [dict(zip(hdr, r)) for r in rows[2:]]
    for i, d in enumerate(ordered):
        kern = d.get("Kernel Name", "?").split("(")[0].replace("void ", "").replace("v3a::", "")[:60]
        dur = num(d, u, "gpu__time_duration.sum")
        rd, wr = num(d, u, "dram__bytes_read.sum") or 0.0, num(d, u, "dram__bytes_write.sum") or 0.0
        g = lambda k: (d.get(k) or "-")  # noqa: E731
        lines.append(f"| {i} | `{kern}` | {dur:.1f} | {g('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active')} | "
                     f"{g('sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active')} | {g('sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active')} | "
                     f"{g('sm__inst_issued.avg.pct_of_peak_sustained_active')} | {(rd + wr) / 1e6:.1f} | {(rd + wr) / dur / 1e3:.0f} | "
                     f"{g('launch__registers_per_thread')} | {g('launch__waves_per_multiprocessor')} |")
        traffic.setdefault(kern, []).append(rd + wr)
    lines.append("")
    lines.append("launch order of the targets: " + ", ".join(names))
    open(os.path.join(OUT, f"{tag}_ncu_targets_table.md"), "w").write("\n".join(lines) + "\n")
    # traffic of the dominant kernels: the FFN1 GEMM launch (third gemm launch: 8192 x 8960 x 1536 + GELU), the DiT self-attention kernel
    tr = json.load(open(os.path.join(OUT, "ncu_traffic.json"))) if os.path.exists(os.path.join(OUT, "ncu_traffic.json")) else {}
    gem = [v for k, vs in traffic.items() if k.startswith("gemm_tcgen05_kernel<256, 2, 0") for v in vs]
    if len(gem) >= 3:
        tr["gemm_tcgen05"] = gem[2]
    pair = [v for k, vs in traffic.items() if k.startswith("fmha_pair_kernel") for v in vs]
    if pair:
        tr["fmha_tcgen05"] = pair[0]
    dec = [v for k, vs in traffic.items() if k.startswith("gemm_tcgen05_kernel<256, 2, 0") for v in vs]
    if len(dec) >= 5:
        tr["gemm_tcgen05|decoder"] = dec[4]   # 13377 x 4096 x 1024 + GELU-erf
    tr["_source"] = f"ncu --set full captures: gemm / fmha from round tag {tag} (tools/profile_gpu_r2.sh, tools/ncu_targets.py), the others from r1b; bytes per launch, dram read + write"
    json.dump(tr, open(os.path.join(OUT, "ncu_traffic.json"), "w"), indent=1)
    print(sorted(f for f in os.listdir(OUT) if f.startswith(tag)))


if __name__ == "__main__":
    main()
