"""Device-agnostic PyTorch restatement of the reference's *Python preprocess* path.

The reference can compute the conditional 3D covariance, the mean shift, the temporal marginal
and the SH colours in PyTorch instead of in the rasterizer (`pipe.compute_cov3D_python`,
`pipe.convert_SHs_python`; reference: gaussian_renderer/__init__.py:73-81,98-111).  Its helpers
hard-code `device="cuda"` (utils/general_utils.py:66,84,103,136), so they cannot run on host
cores.  This module restates the same maths without a device assumption; it is used

  * by `render()` for the two python-preprocess flags (colour evaluation), and
  * by bench.py as the "reference Python/CPU preprocess timed on the host cores" comparator.

Formulas follow: utils/sh_utils.py:58-223 (eval_sh / eval_shfs_4d),
utils/general_utils.py:113-145 (build_rotation_4d / build_scaling_rotation_4d),
scene/gaussian_model.py:34-47,230-251 (conditional covariance, mean offset, marginal).
It is NOT on the product's hot path and is never used as a fallback for the CUDA kernels.
"""
import math

import torch

C0 = 0.28209479177387814
C1 = 0.4886025119029199
C2 = (1.0925484305920792, -1.0925484305920792, 0.31539156525252005, -1.0925484305920792, 0.5462742152960396)
C3 = (-0.5900435899266435, 2.890611442640554, -0.4570457994644658, 0.3731763325901154, -0.4570457994644658,
      1.445305721320277, -0.5900435899266435)


def sh_basis(dirs, deg):
    """[N, (deg+1)^2] real SH basis at unit directions `dirs` [N,3], degrees 0..3."""
    assert 0 <= deg <= 3
    x, y, z = dirs[:, 0], dirs[:, 1], dirs[:, 2]
    cols = [torch.full_like(x, C0)]
    if deg > 0:
        cols += [-C1 * y, C1 * z, -C1 * x]
    if deg > 1:
        xx, yy, zz, xy, yz, xz = x * x, y * y, z * z, x * y, y * z, x * z
        cols += [C2[0] * xy, C2[1] * yz, C2[2] * (2.0 * zz - xx - yy), C2[3] * xz, C2[4] * (xx - yy)]
    if deg > 2:
        cols += [C3[0] * y * (3 * xx - yy), C3[1] * xy * z, C3[2] * y * (4 * zz - xx - yy),
                 C3[3] * z * (2 * zz - 3 * xx - 3 * yy), C3[4] * x * (4 * zz - xx - yy), C3[5] * z * (xx - yy),
                 C3[6] * x * (xx - 3 * yy)]
    return torch.stack(cols, dim=1)


def eval_sh(deg, sh, dirs):
    """sh: [N, 3, K]; dirs: [N, 3] unit -> [N, 3]   (utils/sh_utils.py:58-113)"""
    B = sh_basis(dirs, deg)
    return (sh[:, :, : B.shape[1]] * B[:, None, :]).sum(-1)


def eval_shfs_4d(deg, deg_t, sh, dirs, dirs_t, l=math.pi):
    """4D SH with a Fourier series in time (utils/sh_utils.py:115-223).
    sh: [N, 3, K]; dirs: [N, 3]; dirs_t: [N, 1] -> [N, 3].  As in the reference's Python, the
    temporal terms apply for any `deg` (the CUDA path only applies them under deg > 2)."""
    B = sh_basis(dirs, deg)
    nb = B.shape[1]
    res = (sh[:, :, :nb] * B[:, None, :]).sum(-1)
    for n in range(1, deg_t + 1):
        tn = torch.cos(2 * math.pi * n * dirs_t / l)  # [N,1]
        res = res + tn * (sh[:, :, 16 * n: 16 * n + nb] * B[:, None, :]).sum(-1)
    return res


def rotation_4d(l, r):
    """[N,4,4] from left/right quaternions (utils/general_utils.py:113-133)."""
    ql = l / torch.norm(l, dim=-1, keepdim=True)
    qr = r / torch.norm(r, dim=-1, keepdim=True)
    a, b, c, d = ql.unbind(-1)
    p, q, rr, s = qr.unbind(-1)
    Ml = torch.stack([a, -b, -c, -d, b, a, -d, c, c, d, a, -b, d, -c, b, a]).view(4, 4, -1).permute(2, 0, 1)
    Mr = torch.stack([p, q, rr, s, -q, p, -s, rr, -rr, s, p, -q, -s, -rr, q, p]).view(4, 4, -1).permute(2, 0, 1)
    return (Ml @ Mr).flip(1, 2)


def scaling_rotation_4d(s, l, r):
    """L = R @ diag(s)  (utils/general_utils.py:135-145)."""
    return rotation_4d(l, r) * s[:, None, :]


def covariance_4d(scaling_xyzt, scaling_modifier, rot_l, rot_r):
    L = scaling_rotation_4d(scaling_modifier * scaling_xyzt, rot_l, rot_r)
    return L @ L.transpose(1, 2)


def conditional_covariance_and_offset(scaling_xyzt, scaling_modifier, rot_l, rot_r, dt):
    """(cov3D [N,6], mean_offset [N,3]) at dt = timestamp - t  (scene/gaussian_model.py:34-47)."""
    S = covariance_4d(scaling_xyzt, scaling_modifier, rot_l, rot_r)
    c11, c12, ct = S[:, :3, :3], S[:, 0:3, 3:4], S[:, 3:4, 3:4]
    cur = c11 - c12 @ c12.transpose(1, 2) / ct
    symm = torch.stack([cur[:, 0, 0], cur[:, 0, 1], cur[:, 0, 2], cur[:, 1, 1], cur[:, 1, 2], cur[:, 2, 2]], dim=1)
    offset = c12.squeeze(-1) / ct.squeeze(-1) * dt
    return symm, offset


def marginal_t(scaling_xyzt, scaling_modifier, rot_l, rot_r, t, timestamp, prefilter_var=-1.0):
    """exp(-0.5 (t - timestamp)^2 / Sigma_tt)  (scene/gaussian_model.py:230-242)."""
    sigma = covariance_4d(scaling_xyzt, scaling_modifier, rot_l, rot_r)[:, 3, 3].unsqueeze(1)
    if prefilter_var > 0.0:
        sigma = sigma + prefilter_var
    return torch.exp(-0.5 * (t - timestamp) ** 2 / sigma)


def python_preprocess(xyz, t, scaling_xyzt, rot_l, rot_r, opacity, features, campos, timestamp, time_duration,
                      sh_degree, sh_degree_t, scaling_modifier=1.0):
    """The whole python-preprocess branch of render() for a rot_4d / 4D-SH model, on whatever
    device the inputs live on (reference: gaussian_renderer/__init__.py:73-81,98-111,122-147).
    Returns (means3D, cov3D, opacity, colors, mask)."""
    dt = timestamp - t
    cov3D, delta = conditional_covariance_and_offset(scaling_xyzt, scaling_modifier, rot_l, rot_r, dt)
    means3D = xyz + delta
    marg = marginal_t(scaling_xyzt, scaling_modifier, rot_l, rot_r, t, timestamp)
    opac = opacity * marg
    shs_view = features.transpose(1, 2)
    dir_pp = means3D - campos[None, :]
    dir_pp = dir_pp / dir_pp.norm(dim=1, keepdim=True)
    rgb = eval_shfs_4d(sh_degree, sh_degree_t, shs_view, dir_pp, t - timestamp, time_duration)
    colors = torch.clamp_min(rgb + 0.5, 0.0)
    mask = marg[:, 0] > 0.05
    return means3D, cov3D, opac, colors, mask
