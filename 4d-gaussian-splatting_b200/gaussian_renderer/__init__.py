"""`gaussian_renderer` -- drop-in replacement for the reference package of the same name.

Put `4d-gaussian-splatting_b200/` ahead of the reference checkout on PYTHONPATH and the
reference's `train.py` (`from gaussian_renderer import render`, train.py:17) picks up this
module: `render()` keeps the reference's signature and returns the same 7-key dict
(reference: gaussian_renderer/__init__.py:19-194), but rasterizes through the B200-native
CUDA library (fdgs) instead of diff-gaussian-rasterization.

`render()` only duck-types its arguments (camera / Gaussian model / pipeline flags); it does not
import the reference's `scene` or `utils` packages, so it also works stand-alone (tests, bench).
"""
import math
import os

import torch
from torch.nn import functional as F

from .diff_gaussian_rasterization import GaussianRasterizationSettings, GaussianRasterizer
from . import pyprep

__all__ = ["render", "GaussianRasterizationSettings", "GaussianRasterizer"]


def _select(mask, *tensors):
    return tuple(None if t is None else t[mask] for t in tensors)


def _fused_prologue_ok(pc, pipe, override_color):
    """Can this frame take the raw-parameter entry (GaussianRasterizer.forward_raw)?  Yes when the model is the
    reference's GaussianModel in its all-CUDA configuration: raw parameter tensors present, the standard activations
    (scene/gaussian_model.py:44-60: exp / sigmoid / F.normalize) and no python-side preprocessing or colour override.
    Then exp, F.normalize x2, sigmoid and the torch.cat of get_features run inside the kernels.  FDGS_FUSED_PROLOGUE=0
    (or pipe.fused_prologue = False) forces the getter path."""
    if os.environ.get("FDGS_FUSED_PROLOGUE", "1") == "0" or not getattr(pipe, "fused_prologue", True):
        return False
    if override_color is not None or pipe.compute_cov3D_python or pipe.convert_SHs_python:
        return False
    need = ["_xyz", "_scaling", "_rotation", "_opacity", "_features_dc", "_features_rest"]
    if pc.gaussian_dim == 4:
        need += ["_t", "_scaling_t"]
    if pc.rot_4d:
        need += ["_rotation_r"]
    if not all(isinstance(getattr(pc, n, None), torch.Tensor) for n in need):
        return False
    return (getattr(pc, "scaling_activation", None) is torch.exp and getattr(pc, "opacity_activation", None) is torch.sigmoid
            and getattr(pc, "rotation_activation", None) is F.normalize
            and pc._features_dc.dim() == 3 and pc._features_dc.shape[1] == 1)


def _add_sky(rendered_image, alpha, viewpoint_camera, pc):
    """sky sphere of radius 60 sampled from the model's environment map (reference: :165-178)"""
    assert pc.env_map is not None
    R = 60
    rays_o, rays_d = viewpoint_camera.get_rays()
    od = (rays_o * rays_d).sum(-1)
    dd = (rays_d ** 2).sum(-1)
    delta = od ** 2 - dd * ((rays_o ** 2).sum(-1) - R ** 2)
    assert (delta > 0).all()
    t_inter = -od + torch.sqrt(delta) / dd
    hit = rays_o + rays_d * t_inter.unsqueeze(-1)
    tu = torch.atan2(hit[..., 1:2], hit[..., 0:1]) / (2 * torch.pi) + 0.5
    tv = torch.acos(hit[..., 2:3] / R) / torch.pi
    texcoord = torch.cat([tu, tv], dim=-1) * 2 - 1
    sky = F.grid_sample(pc.env_map[None], texcoord[None])[0]
    return rendered_image + (1 - alpha) * sky


def render(viewpoint_camera, pc, pipe, bg_color: torch.Tensor, scaling_modifier=1.0, override_color=None):
    """Render the scene from `viewpoint_camera` (background tensor must live on the GPU).

    Same contract as the reference's render(): returns {"render", "viewspace_points",
    "visibility_filter", "radii", "depth", "alpha", "flow"}.
    """
    xyz = pc.get_xyz
    dev = xyz.device
    # zero tensor whose .grad receives dL/d(screen-space mean) (reference: :27-31)
    screenspace_points = torch.zeros_like(xyz, dtype=xyz.dtype, requires_grad=True, device=dev) + 0
    try:
        screenspace_points.retain_grad()
    except Exception:
        pass

    env_map_res = getattr(pipe, "env_map_res", 0)
    duration = pc.time_duration[1] - pc.time_duration[0]
    settings = GaussianRasterizationSettings(
        image_height=int(viewpoint_camera.image_height),
        image_width=int(viewpoint_camera.image_width),
        tanfovx=math.tan(viewpoint_camera.FoVx * 0.5),
        tanfovy=math.tan(viewpoint_camera.FoVy * 0.5),
        bg=bg_color if not env_map_res else torch.zeros(3, device=dev),
        scale_modifier=scaling_modifier,
        viewmatrix=viewpoint_camera.world_view_transform,
        projmatrix=viewpoint_camera.full_proj_transform,
        sh_degree=pc.active_sh_degree,
        sh_degree_t=pc.active_sh_degree_t,
        campos=viewpoint_camera.camera_center,
        timestamp=viewpoint_camera.timestamp,
        time_duration=duration,
        rot_4d=pc.rot_4d,
        gaussian_dim=pc.gaussian_dim,
        force_sh_3d=pc.force_sh_3d,
        prefiltered=False,
        debug=pipe.debug,
    )
    rasterizer = GaussianRasterizer(raster_settings=settings)

    if _fused_prologue_ok(pc, pipe, override_color):
        # raw-parameter entry: no activation kernels, no SH concatenation (reference: :83-91,113-119 and the getters of
        # scene/gaussian_model.py:179-219)
        four_d = pc.gaussian_dim == 4
        rendered_image, radii, depth, alpha, flow, _covs = rasterizer.forward_raw(
            means3D=xyz, means2D=screenspace_points, opacity_logits=pc._opacity, features_dc=pc._features_dc,
            features_rest=pc._features_rest if pc._features_rest.numel() else None, log_scales=pc._scaling,
            rotations_raw=pc._rotation, flow_2d=torch.zeros_like(xyz[:, :2]), ts=pc._t if four_d else None,
            log_scales_t=pc._scaling_t if four_d else None, rotations_r_raw=pc._rotation_r if pc.rot_4d else None,
            prefilter_var=pc.prefilter_var if (four_d and pc.prefilter_var > 0.0) else -1.0)
        assert not env_map_res or pc.env_map is not None
        if env_map_res:
            rendered_image = _add_sky(rendered_image, alpha, viewpoint_camera, pc)
        return {"render": rendered_image, "viewspace_points": screenspace_points, "visibility_filter": radii > 0,
                "radii": radii, "depth": depth, "alpha": alpha, "flow": flow}

    means3D, means2D, opacity = xyz, screenspace_points, pc.get_opacity
    scales = scales_t = rotations = rotations_r = ts = cov3D_precomp = None
    prefilter_var = -1.0
    four_d = pc.gaussian_dim == 4
    py_cov = bool(pipe.compute_cov3D_python)
    marginal = None

    if py_cov:
        # covariance / mean shift / marginal in PyTorch (reference: :73-81)
        if pc.rot_4d:
            cov3D_precomp, delta_mean = pc.get_current_covariance_and_mean_offset(scaling_modifier,
                                                                                  viewpoint_camera.timestamp)
            means3D = means3D + delta_mean
        else:
            cov3D_precomp = pc.get_covariance(scaling_modifier)
        if four_d:
            marginal = pc.get_marginal_t(viewpoint_camera.timestamp)
            opacity = opacity * marginal
    else:
        # the rasterizer slices the 4D covariance itself (reference: :82-91)
        scales, rotations = pc.get_scaling, pc.get_rotation
        if four_d:
            scales_t, ts = pc.get_scaling_t, pc.get_t
            if pc.rot_4d:
                rotations_r = pc.get_rotation_r
            if pc.prefilter_var > 0.0:
                prefilter_var = pc.prefilter_var

    shs = colors_precomp = None
    if override_color is not None:
        colors_precomp = override_color
    elif pipe.convert_SHs_python:
        # SH -> RGB in PyTorch (reference: :98-111)
        feats = pc.get_features
        shs_view = feats.transpose(1, 2).view(-1, 3, pc.get_max_sh_channels)
        if py_cov:
            centres = means3D
        else:
            _, delta_mean = pc.get_current_covariance_and_mean_offset(scaling_modifier, viewpoint_camera.timestamp)
            centres = means3D + delta_mean
        dirs = (centres - viewpoint_camera.camera_center.repeat(feats.shape[0], 1)).detach()
        dirs = dirs / dirs.norm(dim=1, keepdim=True)
        if pc.gaussian_dim == 3 or pc.force_sh_3d:
            sh2rgb = pyprep.eval_sh(pc.active_sh_degree, shs_view, dirs)
        else:
            dir_t = (pc.get_t - viewpoint_camera.timestamp).detach()
            sh2rgb = pyprep.eval_shfs_4d(pc.active_sh_degree, pc.active_sh_degree_t, shs_view, dirs, dir_t, duration)
        colors_precomp = torch.clamp_min(sh2rgb + 0.5, 0.0)
    else:
        shs = pc.get_features
        if four_d and ts is None:
            ts = pc.get_t

    flow_2d = torch.zeros_like(xyz[:, :2])

    # python-side temporal prefilter (reference: :122-147)
    mask = None
    if py_cov and four_d:
        mask = marginal[:, 0] > 0.05
        (means2D, means3D, ts, shs, colors_precomp, opacity, scales, scales_t, rotations, rotations_r,
         cov3D_precomp, flow_2d) = _select(mask, means2D, means3D, ts, shs, colors_precomp, opacity, scales,
                                           scales_t, rotations, rotations_r, cov3D_precomp, flow_2d)

    rendered_image, radii, depth, alpha, flow, _covs = rasterizer(
        means3D=means3D, means2D=means2D, shs=shs, colors_precomp=colors_precomp, flow_2d=flow_2d,
        opacities=opacity, ts=ts, scales=scales, scales_t=scales_t, rotations=rotations,
        rotations_r=rotations_r, cov3D_precomp=cov3D_precomp, prefilter_var=prefilter_var)

    if env_map_res:
        rendered_image = _add_sky(rendered_image, alpha, viewpoint_camera, pc)

    if mask is not None:
        radii_all = radii.new_zeros(mask.shape)
        radii_all[mask] = radii
    else:
        radii_all = radii

    # Gaussians that were culled or have zero radius are excluded from the densification statistics
    return {"render": rendered_image,
            "viewspace_points": screenspace_points,
            "visibility_filter": radii_all > 0,
            "radii": radii_all,
            "depth": depth,
            "alpha": alpha,
            "flow": flow}
