"""3D-Gaussian PLY writer with the reference's attribute layout (host I/O):
AS/model/ply_export.py:12-74 -- x y z, zero normals, f_dc_0..2 (SH band 0), optional f_rest_*, opacity (as stored, not
logit), log(scale_0..2), rot_0..3 as (w, x, y, z).  Written as binary_little_endian float32 records, the format plyfile
emits for a structured float32 array; built with vectorised numpy instead of per-vertex tuples (2.6 M Gaussians in < 1 s).
"""
from __future__ import annotations

from pathlib import Path
from typing import List

import numpy as np
import torch


def ply_attributes(num_rest: int) -> List[str]:
    a = ["x", "y", "z", "nx", "ny", "nz"] + [f"f_dc_{i}" for i in range(3)] + [f"f_rest_{i}" for i in range(num_rest)]
    return a + ["opacity"] + [f"scale_{i}" for i in range(3)] + [f"rot_{i}" for i in range(4)]


def export_ply(means: torch.Tensor, scales: torch.Tensor, rotations: torch.Tensor, harmonics: torch.Tensor, opacities: torch.Tensor,
               path, shift_and_scale: bool = False, save_sh_dc_only: bool = True) -> Path:
    """means [N,3], scales [N,3], rotations [N,4] (xyzw), harmonics [N,3,d_sh], opacities [N]."""
    from scipy.spatial.transform import Rotation as R

    path = Path(path)
    means, scales = means.detach().float().cpu(), scales.detach().float().cpu()
    if shift_and_scale:
        means = means - means.median(dim=0).values
        f = means.abs().quantile(0.95, dim=0).max()
        means, scales = means / f, scales / f
    # the reference round-trips the quaternions through a rotation matrix (normalisation + canonical form), then stores wxyz
    q = R.from_matrix(R.from_quat(rotations.detach().float().cpu().numpy()).as_matrix()).as_quat()
    wxyz = np.stack((q[:, 3], q[:, 0], q[:, 1], q[:, 2]), axis=-1)
    harmonics = harmonics.detach().float().cpu()
    cols = [means.numpy(), np.zeros((means.shape[0], 3), np.float32), harmonics[..., 0].contiguous().numpy()]
    n_rest = 0
    if not save_sh_dc_only:
        rest = harmonics[..., 1:].flatten(start_dim=1).contiguous().numpy()
        n_rest = rest.shape[1]
        cols.append(rest)
    cols += [opacities.detach().float().cpu().numpy()[:, None], scales.log().numpy(), wxyz]
    rec = np.ascontiguousarray(np.concatenate(cols, axis=1).astype("<f4"))
    names = ply_attributes(n_rest)
    assert rec.shape[1] == len(names)
    header = "ply\nformat binary_little_endian 1.0\n" + f"element vertex {rec.shape[0]}\n" + "".join(f"property float {n}\n" for n in names) + "end_header\n"
    path.parent.mkdir(exist_ok=True, parents=True)
    with open(path, "wb") as f:
        f.write(header.encode("ascii"))
        f.write(rec.tobytes())
    return path


def read_ply(path):
    """Minimal reader of the files `export_ply` writes -> (attribute names, float32 [N, n_attr])."""
    raw = open(path, "rb").read()
    end = raw.index(b"end_header\n") + len(b"end_header\n")
    lines = raw[:end].decode("ascii").splitlines()
    n = int(next(l for l in lines if l.startswith("element vertex")).split()[-1])
    names = [l.split()[-1] for l in lines if l.startswith("property float")]
    return names, np.frombuffer(raw[end:], dtype="<f4").reshape(n, len(names))
