"""ORACLE (test infrastructure): CPU restatement of the sampler the reference builds at
/root/reference/inference_t23d.py:65-70 --
    UniPCMultistepScheduler(prediction_type="flow_prediction", num_train_timesteps=1000,
                            use_flow_sigmas=True, flow_shift=args.flow_shift)
-- and of the CFG denoise loop WanPipeline runs around it (inference_t23d.py:94-103).

PARITY UNPINNED: `UniPCMultistepScheduler` lives in the un-vendored diffusers==0.33.1
(/root/reference/requirements.txt:20); this restates its published algorithm (UniPC, Zhao et al.
2023, B(h)=expm1(h) variant "bh2", solver_order=2, predict_x0, lower_order_final, final sigma 0)
with the tensor-style arithmetic of the original (scalars are fp32 torch tensors).
"""
from __future__ import annotations

import numpy as np
import torch


class UniPCFlowRef:
    def __init__(self, num_train_timesteps=1000, flow_shift=1.0, solver_order=2):
        self.num_train_timesteps = num_train_timesteps
        self.flow_shift = flow_shift
        self.solver_order = solver_order
        self.predict_x0 = True

    def set_timesteps(self, num_inference_steps: int):
        alphas = np.linspace(1, 1 / self.num_train_timesteps, num_inference_steps + 1)
        sigmas = 1.0 - alphas
        sigmas = np.flip(self.flow_shift * sigmas / (1 + (self.flow_shift - 1) * sigmas))[:-1].copy()
        timesteps = (sigmas * self.num_train_timesteps).copy()
        sigmas = np.concatenate([sigmas, [0.0]]).astype(np.float32)  # final_sigmas_type="zero"
        self.sigmas = torch.from_numpy(sigmas)
        self.timesteps = torch.from_numpy(timesteps).to(dtype=torch.int64)
        self.num_inference_steps = num_inference_steps
        self.model_outputs = [None] * self.solver_order
        self.lower_order_nums = 0
        self.last_sample = None
        self.step_index = 0
        self.this_order = 1

    @staticmethod
    def _alpha_sigma(sigma):
        return 1 - sigma, sigma

    def _lambda(self, sigma):
        a, s = self._alpha_sigma(sigma)
        return torch.log(a) - torch.log(s)

    def _Rb(self, rks, hh, order):
        h_phi_1 = torch.expm1(hh)
        h_phi_k = h_phi_1 / hh - 1
        B_h = torch.expm1(hh)
        R, b = [], []
        factorial_i = 1
        rks = torch.stack(rks)
        for i in range(1, order + 1):
            R.append(torch.pow(rks, i - 1))
            b.append(h_phi_k * factorial_i / B_h)
            factorial_i *= i + 1
            h_phi_k = h_phi_k / hh - 1 / factorial_i
        return torch.stack(R), torch.stack(b), h_phi_1, B_h

    def _predict(self, sample, order):
        m0 = self.model_outputs[-1]
        sigma_t, sigma_s0 = self.sigmas[self.step_index + 1], self.sigmas[self.step_index]
        alpha_t, sigma_t = self._alpha_sigma(sigma_t)
        lambda_t, lambda_s0 = self._lambda(self.sigmas[self.step_index + 1]), self._lambda(sigma_s0)
        h = lambda_t - lambda_s0
        rks, D1s = [], []
        for i in range(1, order):
            mi = self.model_outputs[-(i + 1)]
            rk = (self._lambda(self.sigmas[self.step_index - i]) - lambda_s0) / h
            rks.append(rk)
            D1s.append((mi - m0) / rk)
        rks.append(torch.tensor(1.0))
        R, b, h_phi_1, B_h = self._Rb(rks, -h, order)
        x_t_ = sigma_t / sigma_s0 * sample - alpha_t * h_phi_1 * m0
        if D1s:
            rhos_p = torch.tensor([0.5]) if order == 2 else torch.linalg.solve(R[:-1, :-1], b[:-1])
            pred_res = sum(r * d for r, d in zip(rhos_p, D1s))
        else:
            pred_res = 0
        return x_t_ - alpha_t * B_h * pred_res

    def _correct(self, model_t, last_sample, order):
        m0 = self.model_outputs[-1]
        sigma_t, sigma_s0 = self.sigmas[self.step_index], self.sigmas[self.step_index - 1]
        alpha_t, sigma_t = self._alpha_sigma(sigma_t)
        lambda_t, lambda_s0 = self._lambda(self.sigmas[self.step_index]), self._lambda(sigma_s0)
        h = lambda_t - lambda_s0
        rks, D1s = [], []
        for i in range(1, order):
            mi = self.model_outputs[-(i + 1)]
            rk = (self._lambda(self.sigmas[self.step_index - (i + 1)]) - lambda_s0) / h
            rks.append(rk)
            D1s.append((mi - m0) / rk)
        rks.append(torch.tensor(1.0))
        R, b, h_phi_1, B_h = self._Rb(rks, -h, order)
        rhos_c = torch.tensor([0.5]) if order == 1 else torch.linalg.solve(R, b)
        x_t_ = sigma_t / sigma_s0 * last_sample - alpha_t * h_phi_1 * m0
        corr_res = sum(r * d for r, d in zip(rhos_c[:-1], D1s)) if D1s else 0
        return x_t_ - alpha_t * B_h * (corr_res + rhos_c[-1] * (model_t - m0))

    def step(self, model_output: torch.Tensor, sample: torch.Tensor) -> torch.Tensor:
        use_corrector = self.step_index > 0 and self.last_sample is not None
        x0 = sample - self.sigmas[self.step_index] * model_output  # flow_prediction -> x0
        if use_corrector:
            sample = self._correct(x0, self.last_sample, self.this_order)
        for i in range(self.solver_order - 1):
            self.model_outputs[i] = self.model_outputs[i + 1]
        self.model_outputs[-1] = x0
        this_order = min(self.solver_order, len(self.timesteps) - self.step_index)  # lower_order_final
        self.this_order = min(this_order, self.lower_order_nums + 1)
        self.last_sample = sample
        prev = self._predict(sample, self.this_order)
        if self.lower_order_nums < self.solver_order:
            self.lower_order_nums += 1
        self.step_index += 1
        return prev


@torch.no_grad()
def denoise_loop(forward_fn, latents: torch.Tensor, text_cond, text_uncond, *, num_inference_steps=50,
                 guidance_scale=6.0, flow_shift=5.0, model_dtype=torch.float32):
    """WanPipeline.__call__ core loop (diffusers 0.33.1): two sequential B=1 forwards per step, CFG
    combine, scheduler.step; latents stay fp32 and are cast to the transformer dtype per call."""
    sch = UniPCFlowRef(flow_shift=flow_shift)
    sch.set_timesteps(num_inference_steps)
    latents = latents.float()
    for t in sch.timesteps:
        x = latents.to(model_dtype)
        ts = t.expand(latents.shape[0])
        cond = forward_fn(x, ts, text_cond)
        uncond = forward_fn(x, ts, text_uncond)
        noise = uncond + guidance_scale * (cond - uncond)
        latents = sch.step(noise.float(), latents)
    return latents
