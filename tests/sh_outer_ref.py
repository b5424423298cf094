"""Plain PyTorch fp32 reference of the view-parallel SH-gradient reconstruction (csrc/preprocess_bwd.cu:
sh_outer_sum_kernel; include/fdgs.h: fdgs_sh_outer_sum).  TEST INFRASTRUCTURE: the checker of the CUDA kernel in the
-m gpu tests and the stand-in for it in the gloo (CPU) tests of fdgs/dist.py.

dL_dsh of one view is rank one per Gaussian (reference: backward.cu:144-481 computeColorFromSH_4D backward):
row[k] = w_k * dRGB with w_k = basis_k(dir) for the 16 spatial coefficients (k = 1 uses the l0m0 value: quirk,
backward.cu:190) and cos(2 pi n dt / duration) * basis_k(dir) for the n-th temporal block."""
import math

import torch

from gaussian_renderer import pyprep


def sh_weights(dirs, dir_t, D, D_t, M, duration, sh4d):
    """[n, M] basis weight of every coefficient for unit directions `dirs` and time offsets `dir_t` = t - timestamp."""
    n = dirs.shape[0]
    B = pyprep.sh_basis(dirs, D)
    nc = B.shape[1]
    w = torch.zeros(n, M, dtype=dirs.dtype, device=dirs.device)
    w[:, :nc] = B
    if sh4d and nc > 1:
        w[:, 1] = pyprep.C0
    if sh4d and D > 2 and D_t > 0:
        w[:, 16:16 + nc] = torch.cos(2 * math.pi * dir_t.double() / duration).to(dirs.dtype)[:, None] * B
        if D_t > 1:
            w[:, 32:32 + nc] = torch.cos(2 * math.pi * dir_t.double() * 2 / duration).to(dirs.dtype)[:, None] * B
    return w


def sh_outer_sum_ref(table, stride, meta_off, V, K, slot_of, means3D, ts, scales, scales_t, rotations, rotations_r,
                     scale_modifier, duration, rot_4d, gaussian_dim, force_sh_3d, D, D_t, M):
    P = means3D.shape[0]
    out = torch.zeros(P, M, 3, dtype=torch.float32, device=means3D.device)
    inu = slot_of >= 0
    sl = slot_of[inu].long()
    sh4d = not (gaussian_dim == 3 or force_sh_3d)
    table = table.reshape(-1, stride)
    for v in range(V):
        blk = table[v]
        f = blk[:3 * K].view(K, 3)[sl]
        timestamp, campos = blk[meta_off], blk[meta_off + 1:meta_off + 4]
        mean = means3D[inu]
        if rot_4d:
            _, delta = pyprep.conditional_covariance_and_offset(torch.cat([scales[inu], scales_t[inu]], 1), scale_modifier,
                                                                rotations[inu], rotations_r[inu], timestamp - ts[inu])
            mean = mean + delta
        d = mean - campos[None]
        d = d / d.norm(dim=1, keepdim=True)
        dir_t = (ts[inu][:, 0] - timestamp) if ts is not None and ts.numel() else torch.zeros_like(d[:, 0])
        w = sh_weights(d, dir_t, D, D_t, M, duration, sh4d)
        contrib = w[:, :, None] * f[:, None, :]
        contrib[(f == 0).all(1)] = 0          # not rendered by this view (a degenerate direction must not leak NaNs)
        out[inu] += contrib
    return out
