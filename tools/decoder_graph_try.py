"""Can one stitched-decoder forward be captured in a CUDA graph, and what does a replay cost against the eager forward?  GPU box."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vist3a_b200.stitched_decoder import DecoderConfig, StitchVAE3DB200, random_state_dict  # noqa: E402

torch.cuda.set_device(0)
cfg = DecoderConfig()
m = StitchVAE3DB200.from_state_dict(random_state_dict(cfg, 0, "cuda"), cfg, "cuda")
g = torch.Generator(device="cuda").manual_seed(1)
lat = torch.randn(1, 16, 4, 64, 64, device="cuda", generator=g)
img = torch.rand(1, 3, 13, 448, 448, device="cuda", generator=g) * 2 - 1
for _ in range(3):
    o = m.forward_with_latent(lat, img)
torch.cuda.synchronize()


def timeit(fn, n=5):
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    s.record()
    for _ in range(n):
        fn()
    e.record()
    host = (time.perf_counter() - t0) / n * 1e3
    torch.cuda.synchronize()
    return s.elapsed_time(e) / n, host


print("eager: %.2f ms device-span per forward, %.2f ms of host time to enqueue it" % timeit(lambda: m.forward_with_latent(lat, img)))
ref = m.forward_with_latent(lat, img)
st = torch.cuda.Stream()
st.wait_stream(torch.cuda.current_stream())
gr = torch.cuda.CUDAGraph()
try:
    with torch.cuda.stream(st):
        with torch.cuda.graph(gr, stream=st):
            out = m.forward_with_latent(lat, img)
except Exception as ex:  # noqa: BLE001
    print("capture failed:", type(ex).__name__, str(ex)[:400])
    sys.exit(0)
torch.cuda.synchronize()
print("graph: %.2f ms device-span per replay, %.2f ms host" % timeit(gr.replay))
gm = out.gaussians
same = all(torch.equal(getattr(gm, f), getattr(ref.gaussians, f)) for f in ("means", "covariances", "harmonics", "opacities"))
print("replay output identical to the eager forward:", same)
