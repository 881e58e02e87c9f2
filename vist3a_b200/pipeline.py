"""Device-resident denoise engine: the 50-step CFG sampling loop of one prompt with the two DiT
forwards of a step batched (B=2) and replayed from a CUDA graph.

Call-surface it replaces: the body of `pipe(prompt=..., num_inference_steps=50, guidance_scale=...,
output_type="latent")` between text encoding and latent output (/root/reference/inference_t23d.py:94-103).
A denoise step = cond forward + uncond forward + CFG combine + UniPC update.
"""
from __future__ import annotations

from types import SimpleNamespace
from typing import Optional

import torch

from . import ops
from .unipc import UniPCFlowSampler, UniPCFlowSchedule


class DenoiseEngine:
    def __init__(self, transformer, latent_shape, text_len: int, *, num_inference_steps: int = 50,
                 guidance_scale: float = 6.0, flow_shift: float = 5.0, use_graph: bool = True):
        self.tr = transformer
        self.dev = transformer.device
        self.B = latent_shape[0]
        self.shape = tuple(latent_shape)
        self.g = float(guidance_scale)
        self.n_steps = num_inference_steps
        self.sch = UniPCFlowSchedule(num_inference_steps, flow_shift)
        self.sampler = UniPCFlowSampler(self.sch, self.shape, self.dev)
        c = transformer.config
        D = c.num_attention_heads * c.attention_head_dim
        B2 = 2 * self.B
        dev = self.dev
        self.x = torch.zeros(self.shape, dtype=torch.float32, device=dev)            # latents (fp32, as the pipeline keeps them)
        self.xin = torch.zeros((B2,) + self.shape[1:], dtype=torch.bfloat16, device=dev)  # [cond ; uncond] model input
        self.t = torch.zeros((B2,), dtype=torch.float32, device=dev)
        self.eps = torch.zeros(self.shape, dtype=torch.float32, device=dev)
        self.text = SimpleNamespace(kv=torch.zeros((c.num_layers, B2 * text_len, 2 * D), dtype=torch.bfloat16, device=dev),
                                    B=B2, Lt=text_len)
        self.tsteps = torch.from_numpy(self.sch.timesteps).to(dev, torch.float32)
        self.out = None
        self.graph: Optional[torch.cuda.CUDAGraph] = None
        self.use_graph = use_graph

    # ---- per prompt
    def set_text(self, text_cond: torch.Tensor, text_uncond: torch.Tensor):
        """text_* [B, Lt, text_dim]: runs the text MLP + per-layer K/V projections once (constant over steps)."""
        text = torch.cat([text_cond, text_uncond], 0).to(self.dev, non_blocking=True)
        st = self.tr.encode_text(text)
        if st.kv.shape != self.text.kv.shape:
            raise ValueError(f"text shape {tuple(st.kv.shape)} does not match the engine's {tuple(self.text.kv.shape)}")
        self.text.kv.copy_(st.kv)

    def set_noise(self, noise: torch.Tensor):
        self.x.copy_(noise.to(self.dev, torch.float32, non_blocking=True))
        self.sampler.reset()

    # ---- one step
    def _forward(self):
        self.out = self.tr(self.xin, self.t, None, return_dict=False, text_state=self.text)[0]

    def _capture(self):
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(2):  # warm-up: workspaces, function attributes, RoPE tables
                self._forward()
        torch.cuda.current_stream().wait_stream(s)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            self._forward()
        self.graph = g

    def step(self, i: int):
        B = self.B
        self.xin[:B].copy_(self.x)
        self.xin[B:].copy_(self.x)
        self.t.fill_(float(self.sch.timesteps[i]))
        if self.use_graph:
            if self.graph is None:
                self._capture()
            self.graph.replay()
        else:
            self._forward()
        ops.cfg_combine(self.out[:B], self.out[B:], self.g, out=self.eps)
        self.sampler.step(self.eps, self.x)

    @torch.no_grad()
    def run(self, noise: torch.Tensor, text_cond: torch.Tensor, text_uncond: torch.Tensor) -> torch.Tensor:
        self.set_text(text_cond, text_uncond)
        self.set_noise(noise)
        for i in range(self.n_steps):
            self.step(i)
        return self.x
