"""Flow-matching UniPC sampler + classifier-free-guidance denoise loop on the device.

Replaces, for the VIST3A text-to-3D path, the sampler the reference builds at
/root/reference/inference_t23d.py:65-70 (diffusers `UniPCMultistepScheduler(prediction_type=
"flow_prediction", use_flow_sigmas=True, flow_shift=...)`, bh2, order 2) and the loop of
`WanPipeline.__call__` (inference_t23d.py:94-103: cond forward, uncond forward,
`uncond + g * (cond - uncond)`, `scheduler.step`).

B200-first restructuring: every UniPC update is linear in {sample, last_sample, x0 history, model
output}, so all coefficients of all steps are computed once on the host (fp64) from the sigma
schedule; a step is then two fused elementwise launches (CFG combine; n-term axpby) with no host
synchronisation, and the cond/uncond forwards run as one B=2 batch.
"""
from __future__ import annotations

import math
from typing import List, Optional

import numpy as np
import torch

from . import ops


def flow_sigmas(num_inference_steps: int, flow_shift: float, num_train_timesteps: int = 1000):
    alphas = np.linspace(1, 1 / num_train_timesteps, num_inference_steps + 1)
    sigmas = 1.0 - alphas
    sigmas = np.flip(flow_shift * sigmas / (1 + (flow_shift - 1) * sigmas))[:-1].copy()
    timesteps = (sigmas * num_train_timesteps).astype(np.int64)
    sigmas = np.concatenate([sigmas, [0.0]]).astype(np.float32).astype(np.float64)
    return sigmas, timesteps


def _lam(s):
    return math.log(1 - s) - (math.log(s) if s > 0 else -math.inf)


def _phi(hh):
    """expm1(hh), (expm1(hh)/hh - 1) with the hh -> -inf limit of the final step (sigma = 0)."""
    if math.isinf(hh):
        return -1.0, -1.0
    e = math.expm1(hh)
    return e, e / hh - 1.0


class UniPCFlowSchedule:
    """Per-step linear-combination coefficients of UniPC-bh2 (order <= 2, predict-x0, flow sigmas).

    step i consumes eps_i (the guided model output at x_i) and the state {x_i, last_i = sample the
    previous predictor started from, m_prev = x0_{i-1}, m_prev2 = x0_{i-2}}:
        x0_i   = x_i - sigma_i * eps_i
        xc_i   = corr[i] . (last, m_prev, m_prev2, x0_i)        (i > 0; otherwise xc_i = x_i)
        x_{i+1} = pred[i] . (xc_i, x0_i, m_prev)
    """

    def __init__(self, num_inference_steps=50, flow_shift=5.0, num_train_timesteps=1000, solver_order=2):
        if solver_order != 2:
            raise NotImplementedError("the reference uses the scheduler default solver_order=2")
        self.sigmas, self.timesteps = flow_sigmas(num_inference_steps, flow_shift, num_train_timesteps)
        n = num_inference_steps
        s = self.sigmas
        self.corr: List[Optional[tuple]] = [None] * n
        self.pred: List[tuple] = [None] * n
        lower = 0
        prev_order = 1
        for i in range(n):
            if i > 0:
                self.corr[i] = self._corrector(s, i, prev_order)
            this_order = min(min(2, n - i), lower + 1)
            self.pred[i] = self._predictor(s, i, this_order)
            prev_order = this_order
            if lower < 2:
                lower += 1

    @staticmethod
    def _predictor(s, i, order):
        sig_t, sig_s = s[i + 1], s[i]
        alpha_t = 1 - sig_t
        h = _lam(sig_t) - _lam(sig_s)
        h_phi_1, _ = _phi(-h)
        B_h = h_phi_1
        c_x = sig_t / sig_s
        c_m0 = -alpha_t * h_phi_1
        c_m1 = 0.0
        if order == 2:
            rk = (_lam(s[i - 1]) - _lam(sig_s)) / h
            # pred_res = 0.5 * (m1 - m0) / rk
            k = -alpha_t * B_h * 0.5 / rk
            c_m1 += k
            c_m0 -= k
        return (c_x, c_m0, c_m1)

    @staticmethod
    def _corrector(s, i, order):
        sig_t, sig_s = s[i], s[i - 1]
        alpha_t = 1 - sig_t
        h = _lam(sig_t) - _lam(sig_s)
        hh = -h
        h_phi_1, h_phi_k = _phi(hh)
        B_h = h_phi_1
        c_last = sig_t / sig_s
        c_m0 = -alpha_t * h_phi_1  # m0 = x0_{i-1}
        c_m1 = 0.0                 # m1 = x0_{i-2}
        if order == 1:
            rho_t = 0.5
        else:
            rk = (_lam(s[i - 2]) - _lam(sig_s)) / h
            b1 = h_phi_k / B_h
            h_phi_k2 = h_phi_k / hh - 0.5
            b2 = h_phi_k2 * 2 / B_h
            # solve [[1, 1], [rk, 1]] rho = [b1, b2]
            det = 1.0 - rk
            rho0 = (b1 - b2) / det
            rho_t = (b2 - rk * b1) / det
            k = -alpha_t * B_h * rho0 / rk
            c_m1 += k
            c_m0 -= k
        kt = -alpha_t * B_h * rho_t  # times (x0_i - m0)
        return (c_last, c_m0 - kt, c_m1, kt)


class UniPCFlowSampler:
    """Stateful stepping on device tensors (fp32 latents), mirroring `scheduler.step(noise, t, latents)`."""

    def __init__(self, schedule: UniPCFlowSchedule, shape, device="cuda"):
        self.sch = schedule
        self.i = 0
        z = lambda: torch.zeros(shape, dtype=torch.float32, device=device)
        self.x0, self.m_prev, self.m_prev2, self.last, self.xc = z(), z(), z(), z(), z()

    def reset(self):
        self.i = 0

    def step(self, eps: torch.Tensor, sample: torch.Tensor) -> torch.Tensor:
        """eps: guided model output (fp32), sample: current latents (fp32, updated in place)."""
        i, sch = self.i, self.sch
        ops.axpby_n(self.x0, [sample, eps], [1.0, -float(sch.sigmas[i])])
        if i > 0:
            cl, c0, c1, ct = sch.corr[i]
            ops.axpby_n(self.xc, [self.last, self.m_prev, self.m_prev2, self.x0], [cl, c0, c1, ct])
            cur = self.xc
        else:
            cur = sample
        cx, c0, c1 = sch.pred[i]
        # last <- cur (the sample this predictor starts from), then x_{i+1}
        self.last.copy_(cur)
        ops.axpby_n(sample, [self.last, self.x0, self.m_prev], [cx, c0, c1])
        self.m_prev2, self.m_prev, self.x0 = self.m_prev, self.x0, self.m_prev2
        self.i += 1
        return sample


@torch.no_grad()
def denoise(transformer, latents: torch.Tensor, text_cond: torch.Tensor, text_uncond: torch.Tensor, *,
            num_inference_steps: int = 50, guidance_scale: float = 6.0, flow_shift: float = 5.0,
            batch_cfg: bool = True, step_callback=None) -> torch.Tensor:
    """50-step CFG sampling of one prompt batch: latents [B,16,T,H,W] fp32 noise -> denoised latents.
    cond and uncond run as one 2B batch (same arithmetic as the reference's two sequential forwards)."""
    dev = transformer.device
    sch = UniPCFlowSchedule(num_inference_steps, flow_shift)
    x = latents.to(dev, torch.float32).clone()
    B = x.shape[0]
    sampler = UniPCFlowSampler(sch, x.shape, dev)
    eps = torch.empty_like(x)
    if batch_cfg:
        text = torch.cat([text_cond, text_uncond], 0).to(dev)
        tstate = transformer.encode_text(text)
    else:
        st_c, st_u = transformer.encode_text(text_cond.to(dev)), transformer.encode_text(text_uncond.to(dev))
    tsteps = torch.from_numpy(sch.timesteps).to(dev)
    for i in range(num_inference_steps):
        t = tsteps[i].expand(B)
        xb = x.to(torch.bfloat16)
        if batch_cfg:
            out = transformer(torch.cat([xb, xb], 0), torch.cat([t, t], 0), None, return_dict=False, text_state=tstate)[0]
            cond, uncond = out[:B], out[B:]
        else:
            cond = transformer(xb, t, None, return_dict=False, text_state=st_c)[0]
            uncond = transformer(xb, t, None, return_dict=False, text_state=st_u)[0]
        ops.cfg_combine(cond, uncond, guidance_scale, out=eps)
        sampler.step(eps, x)
        if step_callback is not None:
            step_callback(i, x)
    return x
