#!/usr/bin/env python
"""Benchmark of the VIST3A hot path on B200 (BASELINE.json metric: denoise-steps/sec & Gaussians/sec,
VIST3A-1.3B, 512x512x13 views).

    python bench.py --gpus N --steps K --warmup W            # this framework (sm_100a kernels)
    python bench.py --impl reference --gpus N --steps K ...   # the reference's CPU path (oracle port)

A "step" is one denoise step of one prompt: cond forward + uncond forward of the Wan-1.3B DiT at
L = 4096 latent tokens (13 views @ 512x512) and 512 text tokens, CFG combine, UniPC update.
Workload = BASELINE.json configs[1].  Under torchrun every rank runs its own prompt (prompts shard
over GPUs with no data-path collective in the denoise loop: weak scaling).  The second half of the metric
(Gaussians/sec, configs[2]) is measured in the same run and reported under "gaussians": the stitched
latent->3DGS decoder at 13 views x 448x448 (2 609 152 Gaussians per prompt), decoder-only and end to end
(50 denoise steps + decode + the NCCL all-gather of the Gaussian tensors when N > 1).
Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

STEP_TFLOP = 27.756  # one denoise step (cond + uncond), 1.3B, L=4096, Lt=512
FWD_BLOCK_GFLOP = 462.25
METRIC = "denoise_steps_per_sec"
UNIT = "steps/s"
WORKLOAD = "VIST3A-1.3B DiT 50-step denoise, 512x512x13 views (latent [1,16,4,64,64], L=4096, 512 text tokens), batch 1 per GPU"
DECODER_TFLOP = 50.28  # transformer 43.75 (bf16) + heads 6.53 (tf32) per prompt at 13 views
VIEWS, IMG = 13, 448
N_GAUSS = VIEWS * IMG * IMG


def step_tflop(D, F, layers, L, Lt, text_dim=4096):
    """Algorithmic TFLOP of one denoise step (cond + uncond forward), SURVEY §8(d) accounting: per block QKV + self-attention +
    out + cross (q, text k/v, attention, out) + FFN; plus patch embedding, text / time embedders and the output head."""
    block = 2 * L * D * 3 * D + 4 * L * L * D + 2 * L * D * D + (2 * L * D * D + 2 * Lt * D * 2 * D + 4 * L * Lt * D + 2 * L * D * D) + 4 * L * D * F
    extra = 2 * L * 64 * D * 2 + 2 * Lt * (text_dim * D + D * D) + 2 * (256 * D + D * D + D * 6 * D)
    return 2.0 * (layers * block + extra) / 1e12


def configure(args):
    """BASELINE configs: 1.3B / 13 views (headline, configs[1-2]) or 14B / 21 views (configs[3]); sets the module-level workload."""
    global WORKLOAD, STEP_TFLOP, DECODER_TFLOP, VIEWS, N_GAUSS
    VIEWS = args.views
    N_GAUSS = VIEWS * IMG * IMG
    T = (VIEWS - 1) // 4 + 1
    L = T * 32 * 32
    D, F, layers, params = (1536, 8960, 30, "1.419 B") if args.model == "1.3b" else (5120, 13824, 40, "14.29 B")
    STEP_TFLOP = step_tflop(D, F, layers, L, 512)
    DECODER_TFLOP = {13: 50.28, 21: 98.70}[VIEWS]  # 43.75 + 6.53 at 13 views, 88.16 + 10.54 at 21
    name = "1.3B" if args.model == "1.3b" else "14B"
    WORKLOAD = (f"VIST3A-{name} DiT 50-step denoise, 512x512x{VIEWS} views (latent [1,16,{T},64,64], L={L}, 512 text tokens), "
                f"batch {args.prompts_per_gpu} per GPU")
    return T, params


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "bf16_burst": d["bf16_tflops"], "bf16_sustained": d["bf16_tflops_sustained"], "src": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_burst": 1590.0, "bf16_sustained": 1400.0, "src": "fallback"}


def _graph_time_fmha(B, H, Lq, Lk, D, iters=10):
    """device ms per attention call (kernel + merge kernel where the key split applies) inside a CUDA graph of `iters` back-to-back calls --
    how the step itself runs them; an eager event bracket also counts the host-side gap between the two launches of a call"""
    import torch

    from vist3a_b200 import ops
    q = torch.randn(B, Lq, H, D, device="cuda").bfloat16()
    k = torch.randn(B, Lk, H, D, device="cuda").bfloat16()
    v = torch.randn(B, Lk, H, D, device="cuda").bfloat16()
    o = torch.empty_like(q)
    for _ in range(3):
        ops.fmha(q, k, v, out=o)
    torch.cuda.synchronize()
    st = torch.cuda.Stream()
    st.wait_stream(torch.cuda.current_stream())
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.stream(st):
        with torch.cuda.graph(gr, stream=st):
            for _ in range(iters):
                ops.fmha(q, k, v, out=o)
    torch.cuda.synchronize()
    for _ in range(3):   # clocks ramp after the host-side work in front of this measurement
        gr.replay()
    ts = []
    for _ in range(7):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        gr.replay()
        e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e) / iters)
    return statistics.median(ts)


def _attention_roofline(detail, pk):
    """north_star: "achieved fraction of the attention-GEMM roofline" -- the fmha calls of one step by shape (self / cross attention):
    algorithmic 4*B*H*Lq*Lk*d FLOPs over their device time inside a CUDA graph (`achieved`; `eager_ms` = the CUDA-event brackets of the
    eager step, which include the host-side gap between the attention and merge launches), against the sustained and burst measured peaks"""
    out = {}
    for k, v in detail.items():
        if not k.startswith("fmha_tcgen05|") or not v["flops"]:
            continue
        dims = [int(x) for x in k.split("|")[1].split("x")]
        name = "self" if dims[2] == dims[3] else "cross"
        ms_call = _graph_time_fmha(*dims)
        tf = v["flops"] / v["launches"] / (ms_call / 1e3) / 1e12
        out[name] = {"shape_BxHxLqxLkxD": k.split("|")[1], "launches": v["launches"], "ms": round(ms_call * v["launches"], 4), "eager_ms": round(v["ms"], 4),
                     "achieved": tf, "unit": "TFLOP/s", "timing": "CUDA graph of 10 back-to-back calls (median of 7 replays), random q/k/v of the step's shape",
                     "frac": tf / pk["bf16_sustained"], "frac_of_burst": tf / pk["bf16_burst"]}
    return out


def _ncu_traffic(kernel):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, from the committed `ncu --set full`
    capture (profiles/ncu_traffic.json, written by tools/ncu_summary.py --traffic); None when no capture is committed."""
    p = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    try:
        return json.load(open(p)).get(kernel)
    except Exception:
        return None


class ClockSampler:
    """nvidia-smi clock / throttle-reason samples during the timed region (B200_PROFILING.md recipe)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self._stop, self._th = index, [], threading.Event(), None

    def _run(self):
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            self._stop.wait(0.2)

    def __enter__(self):
        self._th = threading.Thread(target=self._run, daemon=True)
        self._th.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._th.join(timeout=6)

    def summary(self):
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows if len(r) >= 7 for n, v in zip(names, r[3:7]) if v.lower().startswith("active")})
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(self.rows)}


# --------------------------------------------------------------------------------------------------
# CPU baseline (oracle port of the DiT; diffusers is un-vendored so the reference itself cannot run)
# --------------------------------------------------------------------------------------------------
def cpu_sample(nb: int, repeats: int = 1):
    """Time `nb` full-size 1.3B blocks (+embed/head) of the oracle on the host cores; returns (seconds, est steps/s)."""
    import dataclasses

    import torch

    from oracle import wan_dit_ref as R

    cfg = dataclasses.replace(R.WAN_1_3B, num_layers=nb)
    sd = R.init_state_dict(cfg, seed=0, round_bf16=False)
    lat, txt = R.synthetic_inputs(cfg, frames=4, hw=64, text_len=512, text_valid=200, seed=0)
    t = torch.tensor([999.0])
    times = []
    for _ in range(repeats):
        t0 = time.perf_counter()
        R.wan_forward(sd, cfg, lat, t, txt)
        times.append(time.perf_counter() - t0)
    return times


def cpu_decoder_sample(full: bool = True):
    """One forward of the decoder oracle (full widths: 1024-dim tokens, 22 + 48 blocks, DPT heads) on the host cores.  full: the BASELINE
    size, 13 views x 448x448 -> 2 609 152 Gaussians ("one full-size forward", ~1-2 min); else 5 views x 224x224 (1/10.4 of
    it).  Returns (seconds, Gaussians, description)."""
    from oracle import decoder_ref as D

    sd = D.init_state_dict(D.FULL, seed=1, round_bf16=False)
    if full:
        lat, img = D.synthetic_inputs(D.FULL, views_latent=4, latent_hw=64, image_hw=448, seed=2)
        res, what = 512, "13 views x 448x448 (2 609 152 Gaussians: the whole BASELINE decoder workload, one forward)"
    else:
        lat, img = D.synthetic_inputs(D.FULL, views_latent=2, latent_hw=32, image_hw=224, seed=2)
        res, what = 256, "5 views x 224x224 (250 880 Gaussians; 1/10.4 of the 13 x 448x448 workload)"
    t0 = time.perf_counter()
    out = D.decoder_forward(sd, D.FULL, lat, img, resolution=res)
    dt = time.perf_counter() - t0
    n = out["means"].shape[1]
    return dt, n, "oracle/decoder_ref.py fp32, full-width model, " + what


def cpu_full_steps(n_steps: int, warmup: int):
    """The reference's CPU path for one denoise step, run as it is -- no extrapolation: cond forward + uncond forward of the whole 30-block
    1.3B DiT at L = 4096 / 512 text tokens (two sequential B = 1 calls, as WanPipeline issues them), CFG combine, UniPC update; oracle
    restatement (oracle/wan_dit_ref.py + oracle/unipc_ref.py), fp32, all host threads.  Returns the seconds of each timed step."""
    import torch

    from oracle import unipc_ref as U
    from oracle import wan_dit_ref as R

    cfg = R.WAN_1_3B
    sd = R.init_state_dict(cfg, seed=0, round_bf16=False)
    lat, txt_c = R.synthetic_inputs(cfg, frames=4, hw=64, text_len=512, text_valid=200, seed=0)
    _, txt_u = R.synthetic_inputs(cfg, frames=4, hw=64, text_len=512, text_valid=60, seed=1)
    txt_c, txt_u = txt_c.float(), txt_u.float()
    sch = U.UniPCFlowRef(flow_shift=5.0)
    sch.set_timesteps(50)
    x = lat.float()
    times = []
    for i in range(warmup + n_steps):
        if sch.step_index >= 50:       # a new prompt every 50 steps
            sch.set_timesteps(50)
            x = lat.float()
        t0 = time.perf_counter()
        t = sch.timesteps[sch.step_index].expand(1)
        c = R.wan_forward(sd, cfg, x, t, txt_c, cast_fp32=False)
        u = R.wan_forward(sd, cfg, x, t, txt_u, cast_fp32=False)
        x = sch.step(u + 6.0 * (c - u), x)
        times.append(time.perf_counter() - t0)
    return times[warmup:]


def _host_threads():
    """all the host threads this process may use: torchrun exports OMP_NUM_THREADS=1 to its workers, which would time a 1-thread CPU arm"""
    import torch

    try:
        avail = len(os.sched_getaffinity(0))
    except AttributeError:
        avail = os.cpu_count() or 1
    if torch.get_num_threads() < avail:
        torch.set_num_threads(avail)
    return torch.get_num_threads()


def run_reference(args):
    """`--impl reference`: the reference's own CPU implementation of the path on the box's host cores.  The DiT lives in un-vendored
    diffusers (kind "port": the oracle restatement).  Every timed step is one WHOLE denoise step (two 30-block forwards + CFG + UniPC,
    ~14 s on 16 cores): ms_per_step is what was measured, nothing is extrapolated.  `--ref-blocks < 30` (opt-in, for smoke runs only)
    shortens the forwards and says so in `sample`."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = _host_threads()
    nb = max(1, min(30, args.ref_blocks))
    if nb == 30:
        times = cpu_full_steps(args.steps, args.warmup)
        est_step = times
        sample = ("every timed step is one whole denoise step of the 1.3B DiT (cond + uncond 30-block forwards at L=4096, Lt=512, fp32, "
                  "CFG combine, UniPC update): measured, not extrapolated")
    else:
        times = cpu_sample(nb, repeats=args.warmup + args.steps)[args.warmup:]
        est_step = [2.0 * 30.0 / nb * t for t in times]
        sample = (f"SMOKE MODE (--ref-blocks {nb}): each step times {nb} of 30 blocks of ONE forward and extrapolates, steps/s = 1 / (2 * 30/{nb} * t); "
                  "not a measurement of the full step")
    sps = 1.0 / statistics.mean(est_step)
    gauss = None
    if not args.no_decoder:
        dt, n, what = cpu_decoder_sample(full=not args.cpu_decoder_reduced)
        gauss = {"decoder_gaussians_per_sec": n / dt, "unit": "Gaussians/s", "cores": cores, "kind": "port", "sample": what + f", {dt:.1f} s"}
    line = {"impl": "reference", "metric": METRIC, "value": sps, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * statistics.mean(est_step), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": {"workload": WORKLOAD},
            "cpu_baseline": {"value": sps, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": sps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gaussians": gauss,
            "note": "reference DiT lives in un-vendored diffusers==0.33.1; timed arm is the oracle restatement (oracle/wan_dit_ref.py)"}
    print(json.dumps(line), flush=True)


def run_torch(args):
    """`--impl torch`: same-box GPU comparator ("torch 2.11's cuBLAS/SDPA/cuDNN executing the same graph").  The DiT graph of
    the oracle restatement with bf16 weights under CUDA autocast -- i.e. what the reference does on a GPU (fp16 pipeline under bf16
    autocast, inference_t23d.py:73,87): F.linear -> cuBLAS, F.scaled_dot_product_attention -> flash / cuDNN, LayerNorm in fp32 -- two
    sequential B = 1 forwards per step as WanPipeline issues them, CFG + UniPC in torch.  None of this repo's kernels run here."""
    import torch

    from oracle import unipc_ref as U
    from oracle import wan_dit_ref as R

    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    torch.cuda.set_device(dev)
    configure(args)
    cfg = R.WAN_1_3B if args.model == "1.3b" else R.WAN_14B
    g = torch.Generator(device=dev).manual_seed(0)
    sd = {}
    for k, shp in R.param_shapes(cfg).items():   # the benchmark's init, generated on the device
        if k.endswith("scale_shift_table"):
            sd[k] = torch.randn(shp, device=dev, generator=g) / cfg.inner_dim ** 0.5
        elif "norm" in k and k.endswith(".weight"):
            sd[k] = torch.ones(shp, device=dev)
        elif k.endswith(".bias"):
            sd[k] = torch.zeros(shp, device=dev, dtype=torch.bfloat16)
        else:
            sd[k] = (torch.randn(shp, device=dev, generator=g) * 0.02).bfloat16()
    T = (args.views - 1) // 4 + 1
    lat, txt_c = R.synthetic_inputs(cfg, frames=T, hw=64, text_len=512, text_valid=200, seed=0)
    _, txt_u = R.synthetic_inputs(cfg, frames=T, hw=64, text_len=512, text_valid=60, seed=1)
    lat, txt_c, txt_u = lat.to(dev), txt_c.to(dev), txt_u.to(dev)
    sch = U.UniPCFlowRef(flow_shift=5.0)
    sch.set_timesteps(50)          # the scheduler's scalars stay 0-dim CPU tensors (they combine with CUDA tensors as scalars)
    state = {"x": lat.float()}

    def step(i):
        if sch.step_index >= 50:
            sch.set_timesteps(50)
            state["x"] = lat.float()
        t = sch.timesteps[sch.step_index].to(dev).expand(1)
        x = state["x"].bfloat16()
        with torch.autocast("cuda", dtype=torch.bfloat16):
            c = R.wan_forward(sd, cfg, x, t, txt_c, cast_fp32=False)
            u = R.wan_forward(sd, cfg, x, t, txt_u, cast_fp32=False)
        state["x"] = sch.step((u.float() + 6.0 * (c.float() - u.float())), state["x"])

    with torch.no_grad():
        for i in range(max(args.warmup, 3)):
            step(i)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with ClockSampler(dev.index or 0) as clk:
            e0.record()
            for i in range(args.steps):
                step(i)
            e1.record()
            torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    line = {"impl": "torch", "metric": METRIC, "value": 1e3 / ms, "unit": UNIT, "n_gpus": 1, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": WORKLOAD, "what": "torch eager: oracle DiT graph, bf16 weights under CUDA autocast (cuBLAS GEMMs, torch SDPA, fp32 LayerNorm), "
                                                      "cond and uncond as two sequential B=1 forwards, CFG + UniPC in torch; comparator only"},
            "tflops_model": STEP_TFLOP * 1e3 / ms, "clocks": clk.summary(), "torch": torch.__version__}
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------------
# our arm
# --------------------------------------------------------------------------------------------------
def build_models(device, with_decoder=True, model_name="1.3b", voxelize=False):
    """Random-init weights of the named architectures, generated on the device (no checkpoint is reachable offline)."""
    from types import SimpleNamespace

    from vist3a_b200 import stitched_decoder as SD
    from vist3a_b200 import wan_dit as WD

    cfg = WD.WAN_1_3B_CONFIG if model_name == "1.3b" else WD.WAN_14B_CONFIG
    sd = WD.random_state_dict(cfg, 0, device)
    model = WD.WanTransformer3DModelB200.from_state_dict(sd, cfg, device=device)
    del sd
    dec = None
    if with_decoder:
        dcfg = SD.DecoderConfig(voxelize=voxelize)
        sd = SD.random_state_dict(dcfg, 0, device)
        dec = SD.StitchVAE3DB200.from_state_dict(sd, dcfg, device=device)
        del sd
    return SimpleNamespace(**cfg), model, dec


def run_ours(args):
    import torch
    import torch.distributed as dist

    from vist3a_b200 import _lib, ops
    from vist3a_b200.pipeline import DenoiseEngine
    from vist3a_b200.t23d import WAN_LATENTS_MEAN, WAN_LATENTS_STD, all_gather_gaussians, all_gather_gaussians_async

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: the product path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    # stdout carries exactly ONE JSON line: library chatter written to fd 1 (NCCL prints its version banner there) goes to stderr instead
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    _lib.load()
    T_lat, n_params = configure(args)
    cfg, model, dec = build_models(device, with_decoder=not args.no_decoder, model_name=args.model, voxelize=args.voxelize)
    B, T, HW, Lt = args.prompts_per_gpu, T_lat, 64, 512
    g = torch.Generator().manual_seed(1000 + rank)  # every rank denoises its own prompt
    noise_h = torch.randn(B, 16, T, HW, HW, generator=g).pin_memory()
    tc_h = torch.randn(B, Lt, cfg.text_dim, generator=g).bfloat16()
    tc_h[:, 200:] = 0
    tu_h = torch.randn(B, Lt, cfg.text_dim, generator=g).bfloat16()
    tu_h[:, 60:] = 0
    tc_h, tu_h = tc_h.pin_memory(), tu_h.pin_memory()
    out_h = torch.empty_like(noise_h).pin_memory()

    eng = DenoiseEngine(model, noise_h.shape, Lt, num_inference_steps=50, guidance_scale=6.0, flow_shift=5.0,
                        use_graph=not args.no_graph)
    eng.set_text(tc_h, tu_h)
    eng.set_noise(noise_h)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, n):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(n):
            fn(i)
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            tms = torch.tensor([ms], device=device)
            dist.all_reduce(tms, op=dist.ReduceOp.MAX)
            ms = float(tms)
        return ms

    # ---- device-resident throughput
    nsteps = 50

    def dev_step(i):
        if eng.sampler.i >= nsteps:
            eng.sampler.reset()
        eng.step(eng.sampler.i)

    for i in range(args.warmup):
        dev_step(i)
    if args.ncu_step:
        # one denoise step between cudaProfilerStart/Stop: `ncu --profile-from-start off ... bench.py --no-graph --ncu-step`
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
        dev_step(0)
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
        return
    n0 = _lib.launch_count()
    with ClockSampler(local) as clk:
        ms = timed(dev_step, args.steps)
    launches_eager = _lib.launch_count() - n0
    ms_step = ms / args.steps
    value = world * B * args.steps / (ms / 1e3)   # one denoise step per prompt and timed iteration

    # ---- end to end through the public call with HOST buffers (H2D inputs + D2H result inside the timed region)
    h2d = noise_h.numel() * 4 + tc_h.numel() * 2 + tu_h.numel() * 2
    d2h = out_h.numel() * 4

    def e2e_step(i):
        if eng.sampler.i >= nsteps:
            eng.sampler.reset()
        eng.x.copy_(noise_h, non_blocking=True)       # this step's latents from pinned host memory
        eng.set_text(tc_h, tu_h)                      # this step's text embeddings (H2D + text projections)
        eng.step(eng.sampler.i)
        out_h.copy_(eng.x, non_blocking=True)         # the step's result back to the host
        torch.cuda.current_stream().synchronize()

    for i in range(min(args.warmup, 3)):
        e2e_step(i)
    ms_e2e = timed(e2e_step, args.steps)
    e2e_value = world * B * args.steps / (ms_e2e / 1e3)

    # ---- Gaussians/s: stitched decoder at 13 views x 448x448 on the (de-normalised) denoised latent, + the Gaussian all-gather
    gauss = None
    dec_launches = 0
    if dec is not None:
        mean = torch.tensor(WAN_LATENTS_MEAN, device=device).view(1, 16, 1, 1, 1)
        std = torch.tensor(WAN_LATENTS_STD, device=device).view(1, 16, 1, 1, 1)
        lat = eng.x.clamp(-4, 4) * std + mean   # synthetic-weight DiT output, clamped to the range of real VAE latents
        img_h = (torch.rand(B, 3, VIEWS, IMG, IMG, generator=g) * 2 - 1).pin_memory()
        img = img_h.to(device)
        outs = {}

        # the decoder forward as the prompt pipeline runs it (t23d.TextTo3DGS): replayed from a CUDA graph, outputs copied out per call
        dec_fwd = dec.forward_with_latent if args.no_decoder_graph else dec.forward_with_latent_graph

        def dec_step(i):
            outs["o"] = dec_fwd(lat, img)

        n2 = _lib.launch_count()
        dec.forward_with_latent(lat, img)              # (eager: the launch counter does not see graph replays)
        dec_launches_fwd = _lib.launch_count() - n2
        for i in range(2):
            dec_step(i)
        kd = max(2, min(args.steps, args.decoder_iters))
        ms_dec = timed(dec_step, kd) / kd
        dec_launches = dec_launches_fwd * kd
        ms_gather = 0.0
        if world > 1:
            def gather_step(i):
                outs["all"] = all_gather_gaussians(outs["o"].gaussians, fixed_count=not args.voxelize)

            gather_step(0)
            ms_gather = timed(gather_step, 3) / 3
            outs.pop("all", None)
        # Wan VAE decode between the denoiser and the stitched decoder (inference_t23d.py:114-123): random-init weights of the released
        # architecture; its 13 x 512 x 512 frames, resized to 448 x 448, are the decoder's feedforward_image in the end-to-end prompt
        vae, ms_vae = None, None
        if not args.no_vae:
            from vist3a_b200.t23d import views_from_vae
            from vist3a_b200.wan_vae import WanVAEDecoderB200, random_state_dict as vae_random_state_dict

            vae = WanVAEDecoderB200.from_state_dict(vae_random_state_dict(device), None, device)
            for i in range(2):
                views_from_vae(vae, lat)
            ms_vae = timed(lambda i: views_from_vae(vae, lat), kd) / kd

        # end to end per prompt: H2D of text + noise, 50 denoise steps, de-normalise, VAE decode + resize (or H2D of given views), stitched
        # decode, gather, D2H of one scalar (scene scale)
        # With several GPUs the all-gather of prompt p is issued asynchronously and waited for after prompt p + 1 has been queued, so it runs
        # under that prompt's denoising (t23d.generate_sharded does the same): two prompts per rank are timed and the time per prompt reported.
        n_e2e = 2 if world > 1 else 1

        def e2e_prompts(i):
            pending = None
            for _ in range(n_e2e):
                eng.set_text(tc_h, tu_h)
                eng.set_noise(noise_h)
                for k in range(nsteps):
                    eng.step(k)
                lat_i = eng.x.clamp(-4, 4) * std + mean
                if vae is not None:
                    views = views_from_vae(vae, lat_i)
                else:
                    img.copy_(img_h, non_blocking=True)
                    views = img
                o = dec_fwd(lat_i, views)
                h = all_gather_gaussians_async(o.gaussians, fixed_count=not args.voxelize) if world > 1 else None
                if pending is not None:
                    pending[1].wait()
                    pending[0].infos["scene_scale"].cpu()
                pending = (o, h)
            if pending[1] is not None:
                pending[1].wait()
            pending[0].infos["scene_scale"].cpu()

        ms_prompt = timed(e2e_prompts, 1) / n_e2e
        gather_bytes = (world - 1) * B * N_GAUSS * (11 + 3 * 25) * 4 if world > 1 else 0   # received per GPU (own segment stays local)
        gauss = {"n_per_prompt": N_GAUSS, "unit": "Gaussians/s",
                 "decoder_gaussians_per_sec": world * B * N_GAUSS / (ms_dec / 1e3), "decoder_ms": ms_dec,
                 "decoder_tflops": B * DECODER_TFLOP / (ms_dec / 1e3), "gather_ms": ms_gather,
                 "gather": None if world == 1 else {"ms_inline": ms_gather, "bytes_received_per_gpu": gather_bytes,
                                                    "GBps_per_gpu": gather_bytes / (ms_gather / 1e3) / 1e9,
                                                    "frac_of_nvlink5_900GBps": gather_bytes / (ms_gather / 1e3) / 900e9,
                                                    "what": "one all_gather_into_tensor of the field-major Gaussian buffers (no count collective: fixed N); in the "
                                                            "end-to-end prompt it is asynchronous and overlaps the next prompt's denoising"},
                 "e2e_gaussians_per_sec": world * B * N_GAUSS / (ms_prompt / 1e3), "e2e_prompt_ms": ms_prompt,
                 "vae_decode_ms": ms_vae, "vae_what": None if vae is None else "WanVAEDecoderB200: latent [1,16,T,64,64] -> frames 512x512 (29.5 TFLOP at 13 views) + resize to 448x448",
                 "e2e_what": ("one prompt per GPU" if world == 1 else "two prompts per GPU, time per prompt") + ": H2D text + noise, text projections, 50 CFG denoise steps, de-normalise, " +
                             ("Wan VAE decode + 448 resize, " if vae is not None else "H2D views, ") + "stitched decode" +
                             (", asynchronous NCCL all-gather of all ranks' Gaussians (under the next prompt's denoising)" if world > 1 else "") + ", D2H of scene_scale",
                 "decoder_launches_per_forward": dec_launches // kd, "decoder_cuda_graph": not args.no_decoder_graph,
                 "latent": "denoised latent of the random-weight DiT, clamped to [-4, 4] before de-normalisation (real VAE latents are O(1))",
                 "workload": f"VIST3A-{'1.3B' if args.model == '1.3b' else '14B'} full stitched path: DiT -> conv3d_k5x3x3 stitch -> AnySplat "
                             f"enc_blocks_2 -> 3DGS, 512x512x{VIEWS}v"}
        # roofline of the decoder's dominant kernel class: CUDA events around every call of one forward
        with ops.OpTimer() as tmd:
            dec.forward_with_latent(lat, img)
        sd_ = tmd.summary()
        totd = sum(v["ms"] for v in sd_.values())
        topd = max((k for k in sd_ if sd_[k]["flops"] > 0), key=lambda k: sd_[k]["ms"])
        pkd = _peaks()
        dd = sd_[topd]
        achd = dd["flops"] / (dd["ms"] / 1e3) / 1e12
        gauss["roofline"] = {"bound": "tensor", "kernel": topd, "achieved": achd, "peak": pkd["bf16_sustained"], "unit": "TFLOP/s",
                             "frac": achd / pkd["bf16_sustained"], "frac_of_burst": achd / pkd["bf16_burst"], "peak_src": pkd["src"] + " (sustained)",
                             "traffic": _ncu_traffic(topd + "|decoder"), "launches_per_forward": dd["launches"], "avg_launch_ms": dd["ms"] / dd["launches"],
                             "share_of_forward": dd["ms"] / ms_dec, "share_of_eager_kernel_sum": dd["ms"] / totd,
                             "forward_model_tflops": B * DECODER_TFLOP / (ms_dec / 1e3),
                             "forward_frac_of_sustained": B * DECODER_TFLOP / (ms_dec / 1e3) / pkd["bf16_sustained"],
                             "by_kernel": {k: {"ms": round(v["ms"], 4), "launches": v["launches"],
                                               "tflops": (v["flops"] / (v["ms"] / 1e3) / 1e12) if v["flops"] else None,
                                               "gbs": v["bytes"] / (v["ms"] / 1e3) / 1e9} for k, v in sd_.items()}}
        if args.render:    # consumer of the Gaussians: re-render the 13 context views from the predicted cameras (rasteriser forward)
            from vist3a_b200.renderer import DecoderSplattingB200

            rend = DecoderSplattingB200((1.0, 1.0, 1.0))
            rend.render_context_views(outs["o"], (IMG, IMG))
            ms_r = timed(lambda i: rend.render_context_views(outs["o"], (IMG, IMG)), 2) / 2
            gauss["render"] = {"views": VIEWS, "image": f"{IMG}x{IMG}", "ms_per_view": ms_r / (B * VIEWS), "views_per_sec": world * B * VIEWS / (ms_r / 1e3),
                               "what": "DecoderSplattingB200.rendering_fn over the predicted context cameras, all Gaussians of the prompt per view"}
        if args.voxelize:  # voxelised fusion (released AnySplat configs): fewer, fused Gaussians; Gaussians/s above still counts pixels decoded
            gauss["voxelize"] = {"voxel_size": dec.cfg.voxel_size, "voxels_per_prompt": int(outs["o"].gaussians.means.shape[1]),
                                 "voxelize_ratio": float(outs["o"].infos["voxelize_ratio"])}
        del outs

    # ---- per-kernel device timing of one eager step (roofline of the dominant kernel)
    roof = None
    launches_per_step = None
    if rank == 0:
        model_fwd = lambda: eng._forward()
        model_fwd()
        torch.cuda.synchronize()
        n1 = _lib.launch_count()
        with ops.OpTimer() as tm:
            eng.xin[:B].copy_(eng.x)
            eng.xin[B:].copy_(eng.x)
            model_fwd()
            ops.cfg_combine(eng.out[:B], eng.out[B:], eng.g, out=eng.eps)
            eng.sampler.reset()
            eng.sampler.step(eng.eps, eng.x)
        launches_per_step = _lib.launch_count() - n1
        summ = tm.summary()
        if args.detail:
            for k, v in sorted(tm.summary(detail=True).items(), key=lambda kv: -kv[1]["ms"]):
                tf = v["flops"] / (v["ms"] / 1e3) / 1e12 if v["flops"] else 0.0
                print(f"# {k:55s} {v['launches']:4d} launches {v['ms']:9.3f} ms  {tf:8.1f} TFLOP/s  {v['bytes'] / (v['ms'] / 1e3) / 1e9:8.1f} GB/s",
                      file=sys.stderr)
        tot = sum(d["ms"] for d in summ.values())
        top = max((k for k in summ if summ[k]["flops"] > 0), key=lambda k: summ[k]["ms"])
        pk = _peaks()
        d = summ[top]
        ach = d["flops"] / (d["ms"] / 1e3) / 1e12
        roof = {"bound": "tensor", "kernel": top, "achieved": ach, "peak": pk["bf16_sustained"], "unit": "TFLOP/s",
                "frac": ach / pk["bf16_sustained"], "frac_of_burst": ach / pk["bf16_burst"], "peak_src": pk["src"] + " (sustained: kernel timed inside a long step)",
                "traffic": _ncu_traffic(top), "launches_per_step": d["launches"], "avg_launch_ms": d["ms"] / d["launches"],
                "share_of_step": d["ms"] / ms_step, "share_of_eager_kernel_sum": d["ms"] / tot,
                "share_note": "kernel time from CUDA events around every call of one EAGER step; share_of_step divides it by the CUDA-graphed step time",
                "attention": _attention_roofline(tm.summary(detail=True), pk),
                "step_model_tflops": B * STEP_TFLOP / (ms_step / 1e3), "step_frac_of_sustained": B * STEP_TFLOP / (ms_step / 1e3) / pk["bf16_sustained"],
                "by_kernel": {k: {"ms": round(v["ms"], 4), "launches": v["launches"],
                                  "tflops": (v["flops"] / (v["ms"] / 1e3) / 1e12) if v["flops"] else None,
                                  "gbs": v["bytes"] / (v["ms"] / 1e3) / 1e9} for k, v in summ.items()}}

    # ---- CPU baseline (rank 0, N=1 only): bounded sample of the oracle port on the host cores
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline and args.model == "1.3b" and args.views == 13:
        nb = max(1, min(30, args.cpu_blocks))
        try:   # a host-side problem (memory, threads) must not cost the measured GPU line
            cpu_sample(1)  # spins the thread pool up
            t = cpu_sample(nb)[-1]
            sps = 1.0 / (2.0 * 30.0 / nb * t)
            cpu = {"value": sps, "unit": UNIT, "cores": _host_threads(), "kind": "port",
                   "sample": f"{nb} of 30 full-size fp32 blocks of one cond forward incl. embed/head (oracle/wan_dit_ref.py), {t:.2f} s; "
                             f"a step is two such forwards: steps/s = 1/(2*30/{nb}*t)"}
            if gauss is not None:
                dt, n, what = cpu_decoder_sample(full=not args.cpu_decoder_reduced)
                gauss["cpu_baseline"] = {"decoder_gaussians_per_sec": n / dt, "cores": torch.get_num_threads(), "kind": "port",
                                         "sample": what + f", {dt:.1f} s"}
        except Exception as e:  # noqa: BLE001
            cpu = {"value": None, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port", "sample": f"failed: {type(e).__name__}: {e}"}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
                "data": "synthetic",
                "config": {"workload": WORKLOAD, "weights": f"random-init Wan-{args.model.upper()} ({n_params} params)", "cfg": "cond+uncond batched B=2",
                           "cuda_graph": not args.no_graph, "l2": "working set (bf16 weights of every layer + activations) >> 126 MB L2; no flush needed",
                           "prompts_per_gpu": B},
                "roofline": roof, "cpu_baseline": cpu, "clocks": clk.summary(), "gaussians": gauss,
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "ms_per_step": ms_e2e / args.steps},
                "gpu_launches": launches_eager if args.no_graph else (launches_per_step or 0) * args.steps,
                "gpu_launches_note": "graph replays re-launch the captured kernels; count = kernels per step x steps" if not args.no_graph else "eager",
                "tflops_model": STEP_TFLOP * value}
        sys.stdout.flush()
        os.write(json_fd, (json.dumps(line) + "\n").encode())
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference", "torch"])
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--ncu-step", action="store_true", help="profile exactly one eager denoise step (for ncu --profile-from-start off)")
    ap.add_argument("--detail", action="store_true", help="print per-shape kernel timings of one eager step to stderr")
    ap.add_argument("--ref-blocks", type=int, default=30, help="--impl reference: 30 (default) = every timed step is a whole measured step; "
                    "< 30 = smoke mode (shortened forwards, extrapolated, flagged in the line)")
    ap.add_argument("--cpu-decoder-reduced", action="store_true", help="decoder CPU leg at 5 views x 224x224 instead of the full 13 x 448x448 forward")
    ap.add_argument("--cpu-blocks", type=int, default=30, help="full-size blocks of the cpu_baseline leg (30 = one whole forward, ~10 s on 16 cores)")
    ap.add_argument("--prompts-per-gpu", type=int, default=1, help="prompts batched per GPU (BASELINE configs[4] sweep: 1/2/4/8)")
    ap.add_argument("--no-decoder", action="store_true", help="skip the Gaussians/s leg (decoder + gather)")
    ap.add_argument("--no-decoder-graph", action="store_true", help="eager stitched-decoder forward instead of the CUDA-graph replay (A/B)")
    ap.add_argument("--decoder-iters", type=int, default=3)
    ap.add_argument("--no-vae", action="store_true", help="end-to-end prompt without the Wan VAE decode (views copied from the host instead)")
    ap.add_argument("--model", default="1.3b", choices=["1.3b", "14b"], help="Wan DiT size (BASELINE configs[1-2] / configs[3])")
    ap.add_argument("--views", type=int, default=13, choices=[13, 21], help="views per prompt (13: latent T=4, L=4096; 21: T=6, L=6144)")
    ap.add_argument("--voxelize", action="store_true", help="decoder with voxelised Gaussian fusion (voxel_size 0.002)")
    ap.add_argument("--render", action="store_true", help="also time the rasteriser on the decoded Gaussians (13 context views)")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    elif args.impl == "torch":
        run_torch(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
