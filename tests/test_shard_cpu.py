"""Host logic of the multi-GPU path on CPU: prompt sharding and the Gaussian all-gather over a world_size-2 gloo group
(the same code runs over NCCL on the GPUs)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from vist3a_b200.stitched_decoder import Gaussians
from vist3a_b200.t23d import all_gather_gaussians, all_gather_gaussians_async, pack_gaussians, shard_prompts, unpack_gaussians


def _gauss(seed, n=37, d_sh=25):
    g = torch.Generator().manual_seed(seed)
    r = lambda *s: torch.randn(*s, generator=g)
    return Gaussians(means=r(1, n, 3), covariances=r(1, n, 3, 3), harmonics=r(1, n, 3, d_sh), opacities=r(1, n), scales=r(1, n, 3),
                     rotations=r(1, n, 4))


def test_shard_prompts_matches_reference_slicing():
    prompts = [f"p{i}" for i in range(11)]
    seen = []
    for r in range(4):
        part = shard_prompts(prompts, r, 4)
        assert part == prompts[r::4]
        seen += part
    assert sorted(seen) == sorted(prompts)           # every prompt exactly once
    assert shard_prompts(prompts[:2], 3, 4) == []    # ragged: more ranks than prompts
    with pytest.raises(ValueError):
        shard_prompts(prompts, 4, 4)


@pytest.mark.parametrize("cov", [False, True])
def test_pack_unpack_roundtrip(cov):
    g = _gauss(0)
    rec = pack_gaussians(g, cov)
    assert rec.shape == (1, 37, 86 + (9 if cov else 0))
    back = unpack_gaussians(rec, 25, cov)
    for k, v in back.items():
        assert torch.equal(v, getattr(g, k))


def _gauss_packed(seed, b=2, n=19, d_sh=25):
    """Gaussians whose fields are views of one field-major buffer, as the decoder produces them (ops.alloc_gaussian_fields)"""
    from vist3a_b200.ops import alloc_gaussian_fields

    o = alloc_gaussian_fields(b * n, d_sh, "cpu")
    o["packed"].copy_(torch.randn(o["packed"].shape, generator=torch.Generator().manual_seed(seed)))
    return Gaussians(means=o["means"].view(b, n, 3), covariances=o["covariances"].view(b, n, 3, 3), harmonics=o["harmonics"].view(b, n, 3, d_sh),
                     opacities=o["opacities"].view(b, n), scales=o["scales"].view(b, n, 3), rotations=o["rotations"].view(b, n, 4), packed=o["packed"])


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ok = True
        for make, cov in ((_gauss, True), (_gauss_packed, True), (_gauss_packed, False)):   # record path / packed buffer path
            got = all_gather_gaussians(make(100 + rank), with_covariances=cov)
            ok = ok and len(got) == world
            for r in range(world):
                want = make(100 + r)
                ok = ok and (("covariances" in got[r]) == cov)
                for k, v in got[r].items():
                    ok = ok and torch.equal(v, getattr(want, k))
        # ragged: voxelised fusion leaves a different number of Gaussians on every rank
        got = all_gather_gaussians(_gauss(200 + rank, n=11 + 5 * rank), with_covariances=True)
        for r in range(world):
            want = _gauss(200 + r, n=11 + 5 * r)
            for k, v in got[r].items():
                ok = ok and v.shape == getattr(want, k).shape and torch.equal(v, getattr(want, k))
        # fixed count (no count collective, no host read) and two gathers in flight at once, waited in order
        h1 = all_gather_gaussians_async(_gauss_packed(300 + rank), with_covariances=False, fixed_count=True)
        h2 = all_gather_gaussians_async(_gauss(400 + rank), with_covariances=True, fixed_count=True)
        for h, make, base in ((h1, _gauss_packed, 300), (h2, _gauss, 400)):
            got = h.wait()
            for r in range(world):
                want = make(base + r)
                for k, v in got[r].items():
                    ok = ok and torch.equal(v, getattr(want, k))
        q.put((rank, ok))
    finally:
        dist.destroy_process_group()


def test_all_gather_gaussians_gloo_world2():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]


def test_views_from_vae_mirrors_the_reference_resize():
    """inference_t23d.py:114-123: decode, then trilinear (align_corners=False) resize of H x W only"""
    import torch.nn.functional as F

    from vist3a_b200.t23d import views_from_vae

    class FakeVAE(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.w = torch.nn.Parameter(torch.ones(1, dtype=torch.float64))
            self.seen = None

        def decode(self, z, return_dict=True):
            assert return_dict is False
            self.seen = z.dtype
            t = 1 + 4 * (z.shape[2] - 1)
            g = torch.Generator().manual_seed(0)
            return (torch.rand(z.shape[0], 3, t, 8 * z.shape[3], 8 * z.shape[4], generator=g, dtype=z.dtype) * 2 - 1,)

    vae = FakeVAE()
    lat = torch.randn(1, 16, 2, 4, 4)
    out = views_from_vae(vae, lat, size=28)
    assert vae.seen == torch.float64 and out.dtype == torch.float32 and out.shape == (1, 3, 5, 28, 28)
    ref = F.interpolate(vae.decode(lat.double(), return_dict=False)[0].float(), (5, 28, 28), mode="trilinear", align_corners=False)
    assert torch.equal(out, ref)
    # the frame axis is not interpolated: every output frame depends on its own input frame only
    per_frame = F.interpolate(vae.decode(lat.double(), return_dict=False)[0].float()[:, :, 2], (28, 28), mode="bilinear", align_corners=False)
    assert torch.allclose(out[:, :, 2], per_frame, atol=1e-6)


def test_stitched_forward_mirrors_the_reference_surface():
    """StitchVAE3D.forward (models/stitched_model.py:140-163) = VAE encode + sample, then forward_with_latent; the engine takes the caller's
    VAE module and refuses without one (the Wan VAE is not part of it)."""
    import pytest

    from vist3a_b200.stitched_decoder import DecoderConfig, StitchVAE3DB200

    m = StitchVAE3DB200(DecoderConfig(), device="cpu")
    with pytest.raises(NotImplementedError, match="diffusion_vae"):
        m.forward(torch.zeros(1, 3, 5, 16, 16), torch.zeros(1, 3, 5, 14, 14))

    class Dist:
        def __init__(self, z):
            self.z = z

        def sample(self):
            return self.z

    class FakeVAE:
        def encode(self, images):
            self.seen = images
            b, _, t, h, w = images.shape
            return type("Out", (), {"latent_dist": Dist(torch.full((b, 16, 1 + (t - 1) // 4, h // 8, w // 8), 0.25))})()

    got = {}

    def fake_fwl(latent, feedforward_image, train=False):
        got.update(latent=latent, image=feedforward_image, train=train)
        return "decoded"

    m.diffusion_vae = FakeVAE()
    m.forward_with_latent = fake_fwl
    clip, views = torch.rand(2, 3, 5, 16, 16), torch.rand(2, 3, 5, 14, 14)
    assert m.forward(clip, views) == "decoded"
    assert m.diffusion_vae.seen is clip and got["image"] is views and got["train"] is False
    assert got["latent"].shape == (2, 16, 2, 2, 2) and float(got["latent"].mean()) == 0.25


def test_clone_output_keeps_the_field_major_layout():
    """stitched_decoder.clone_output (what a CUDA-graph replay of the decoder hands to the caller): every tensor is copied, the Gaussian fields of
    the copy are views of ONE new flat buffer at the offsets they had in the source (the layout all_gather_gaussians sends without packing),
    nested containers and non-tensor entries survive, and writing to the source afterwards does not reach the copy."""
    from vist3a_b200.stitched_decoder import EncoderOutput, clone_output

    B, N, d_sh = 1, 7, 4
    sizes = {"means": 3, "scales": 3, "rotations": 4, "opacities": 1, "harmonics": 3 * d_sh, "covariances": 9}
    flat = torch.arange(B * N * sum(sizes.values()), dtype=torch.float32)
    views, off = {}, 0
    for k, c in sizes.items():
        views[k] = flat[off:off + B * N * c].view(B, N, c)
        off += B * N * c
    g = Gaussians(means=views["means"], covariances=views["covariances"].view(B, N, 3, 3), harmonics=views["harmonics"].view(B, N, 3, d_sh),
                  opacities=views["opacities"].view(B, N), scales=views["scales"], rotations=views["rotations"], packed=flat)
    out = EncoderOutput(gaussians=g, pred_pose_enc_list=[torch.ones(1, 2, 9), torch.zeros(1, 2, 9)],
                        pred_context_pose={"extrinsic": torch.eye(4).expand(1, 2, 4, 4), "intrinsic": torch.eye(3).expand(1, 2, 3, 3)},
                        depth_dict={"depth": torch.rand(1, 2, 4, 4, 1)}, infos={"scene_scale": torch.tensor(2.0), "voxelize_ratio": 1.0},
                        last_pred_pose_enc=torch.ones(1, 2, 9))
    c = clone_output(out)
    assert c.gaussians.packed is not None and c.gaussians.packed.data_ptr() != flat.data_ptr()
    for f in ("means", "covariances", "harmonics", "opacities", "scales", "rotations"):
        a, b = getattr(g, f), getattr(c.gaussians, f)
        assert torch.equal(a, b) and b.shape == a.shape
        assert b.untyped_storage().data_ptr() == c.gaussians.packed.untyped_storage().data_ptr(), f
        assert b.storage_offset() == a.storage_offset() and b.stride() == a.stride(), f
    assert torch.equal(c.gaussians.packed, flat)
    assert c.infos["voxelize_ratio"] == 1.0 and torch.equal(c.infos["scene_scale"], out.infos["scene_scale"])
    assert len(c.pred_pose_enc_list) == 2 and c.pred_pose_enc_list[0].data_ptr() != out.pred_pose_enc_list[0].data_ptr()
    assert torch.equal(c.pred_context_pose["extrinsic"], out.pred_context_pose["extrinsic"]) and c.distill_infos is None
    flat.zero_()
    out.depth_dict["depth"].zero_()
    assert float(c.gaussians.means.sum()) > 0 and float(c.depth_dict["depth"].sum()) > 0
