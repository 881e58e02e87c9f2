"""all_gather_into_tensor of the per-rank Gaussian records over NCCL: bandwidth and transport (run under torchrun, NCCL_DEBUG=INFO)."""
import os

import torch
import torch.distributed as dist

local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
world, rank = dist.get_world_size(), dist.get_rank()
n = 2609152 * 95  # fp32 values per rank at 13 views
src = torch.randn(n, device="cuda")
dst = torch.empty(world * n, device="cuda")
for _ in range(2):
    dist.all_gather_into_tensor(dst, src)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
dist.barrier()
e0.record()
for _ in range(5):
    dist.all_gather_into_tensor(dst, src)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
if rank == 0:
    print(f"all_gather of {n * 4 / 1e9:.2f} GB per rank x {world} ranks: {ms:.2f} ms = {(world - 1) * n * 4 / ms / 1e6:.0f} GB/s received per GPU; "
          f"p2p access 0->1: {torch.cuda.can_device_access_peer(0, 1)}")
dist.destroy_process_group()
