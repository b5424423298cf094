"""Autograd boundary of the B200-native 4D Gaussian rasterizer.

Keeps the public surface of the reference's `gaussian_renderer/diff_gaussian_rasterization.py`
(reference: gaussian_renderer/diff_gaussian_rasterization.py:34-318) -- the names
`GaussianRasterizationSettings` (18 fields, same order), `GaussianRasterizer` (`forward`,
`markVisible`), `rasterize_gaussians`, the argument order, the validation errors, the 6-tuple
that comes back and the `debug` snapshot files -- so that `gaussian_renderer.render()` and the
reference's `train.py` run unchanged.

What differs is underneath: the reference JIT-compiles diff-gaussian-rasterization at import
(`:17-28`); here `_C` is the prebuilt sm_100a extension `fdgs_C.so` sitting on the C-ABI of
include/fdgs.h.  Importing this module fails loudly when that library has not been built --
there is no PyTorch or CPU fallback path.
"""
from typing import NamedTuple

import torch
from torch import nn

import fdgs

_C = fdgs.ext()

_SNAPSHOT_MSG = {
    "snapshot_fw.dump": "\nAn error occured in forward. Please forward snapshot_fw.dump for debugging.",
    "snapshot_bw.dump": "\nAn error occured in backward. Writing snapshot_bw.dump for debugging.\n",
}


def cpu_deep_copy_tuple(input_tuple):
    """CPU clones of every tensor in a tuple (reference: :30-32)."""
    return tuple(x.cpu().clone() if isinstance(x, torch.Tensor) else x for x in input_tuple)


def _invoke(fn, args, debug, dump_name):
    """Call into the extension; under `debug` dump the arguments on failure (reference: :122-131, :193-206)."""
    if not debug:
        return fn(*args)
    saved = cpu_deep_copy_tuple(args)  # copy before a failing kernel can corrupt them
    try:
        return fn(*args)
    except Exception:
        torch.save(saved, dump_name)
        print(_SNAPSHOT_MSG[dump_name])
        raise


def _notify_forward(radii):
    """View-parallel step (fdgs/dist.py): the radii of a view are known after its forward -- lets the step build the
    union of rendered Gaussians behind the backward pass."""
    from fdgs.dist import ViewParallelStep
    step = ViewParallelStep.current()
    if step is not None:
        step.on_forward(radii)


class GaussianRasterizationSettings(NamedTuple):
    """Per-view constants (reference: :227-245; names and order are the contract)."""
    image_height: int
    image_width: int
    tanfovx: float
    tanfovy: float
    bg: torch.Tensor
    scale_modifier: float
    viewmatrix: torch.Tensor
    projmatrix: torch.Tensor
    sh_degree: int
    sh_degree_t: int
    campos: torch.Tensor
    timestamp: float
    time_duration: float
    rot_4d: bool
    gaussian_dim: int
    force_sh_3d: bool
    prefiltered: bool
    debug: bool


class _RasterizeGaussians(torch.autograd.Function):
    """reference: :67-225.  Inputs, in order: means3D, means2D, sh, colors_precomp, flow_2d, opacities,
    ts, scales, scales_t, rotations, rotations_r, cov3Ds_precomp, prefilter_var, raster_settings."""

    @staticmethod
    def forward(ctx, means3D, means2D, sh, colors_precomp, flow_2d, opacities, ts, scales, scales_t,
                rotations, rotations_r, cov3Ds_precomp, prefilter_var, raster_settings):
        s = raster_settings
        # positional layout of _C.rasterize_gaussians (reference: :88-119, rasterize_points.h:18-49)
        args = (s.bg, means3D, colors_precomp, flow_2d, opacities, ts, scales, scales_t, rotations, rotations_r,
                s.scale_modifier, cov3Ds_precomp, prefilter_var, s.viewmatrix, s.projmatrix, s.tanfovx, s.tanfovy,
                s.image_height, s.image_width, sh, s.sh_degree, s.sh_degree_t, s.campos, s.timestamp,
                s.time_duration, s.rot_4d, s.gaussian_dim, s.force_sh_3d, s.prefiltered, s.debug)
        (num_rendered, color, flow, depth, T, radii, geom_buf, binning_buf, img_buf, covs_com,
         out_means3D) = _invoke(_C.rasterize_gaussians, args, s.debug, "snapshot_fw.dump")

        _notify_forward(radii)
        ctx.raster_settings = s
        ctx.num_rendered = num_rendered
        ctx.prefilter_var = prefilter_var
        ctx.save_for_backward(colors_precomp, means3D, out_means3D, scales, rotations, cov3Ds_precomp, radii, sh,
                              flow_2d, opacities, ts, scales_t, rotations_r, geom_buf, binning_buf, img_buf)
        ctx.mark_non_differentiable(radii)
        # outputs the loss does not touch arrive as None in backward() instead of zero images; the
        # library then skips their recurrences (the reference always blends all seven channels back)
        ctx.set_materialize_grads(False)
        ctx.image_shape = tuple(color.shape)
        # alpha = 1 - T (reference: :140)
        return color, radii, depth, 1 - T, flow, covs_com

    @staticmethod
    def backward(ctx, grad_out_color, grad_radii, grad_depth, grad_alpha, grad_flow, grad_covs_com):
        s = ctx.raster_settings
        (colors_precomp, means3D, out_means3D, scales, rotations, cov3Ds_precomp, radii, sh, flow_2d, opacities,
         ts, scales_t, rotations_r, geom_buf, binning_buf, img_buf) = ctx.saved_tensors
        if grad_out_color is None:
            grad_out_color = torch.zeros(ctx.image_shape, dtype=torch.float32, device=means3D.device)
        none = torch.empty(0, dtype=torch.float32, device=means3D.device)   # "no gradient" placeholder
        grad_depth = none if grad_depth is None else grad_depth
        grad_alpha = none if grad_alpha is None else grad_alpha
        grad_flow = none if grad_flow is None else grad_flow
        # positional layout of _C.rasterize_gaussians_backward (reference: :154-190, rasterize_points.h:51-89)
        args = (s.bg, means3D, out_means3D, radii, colors_precomp, flow_2d, opacities, ts, scales, scales_t,
                rotations, rotations_r, s.scale_modifier, cov3Ds_precomp, ctx.prefilter_var, s.viewmatrix,
                s.projmatrix, s.tanfovx, s.tanfovy, grad_out_color, grad_depth, grad_alpha, grad_flow, sh,
                s.sh_degree, s.sh_degree_t, s.campos, s.timestamp, s.time_duration, s.rot_4d, s.gaussian_dim,
                s.force_sh_3d, geom_buf, ctx.num_rendered, binning_buf, img_buf, s.debug)
        # View-parallel step (fdgs/dist.py, multi-GPU): the dense dL_dsh of this view is not produced; the backward
        # emits the [P,3] colour factor of its rank-one rows and the step rebuilds / sums the rows of all views after
        # the exchange.  No reference counterpart (its views are sequential on one GPU, train.py:104-166).
        from fdgs.dist import ViewParallelStep
        step = ViewParallelStep.current()
        if step is not None and step.wants_factors(sh, cov3Ds_precomp) and ctx.needs_input_grad[2]:
            out = _invoke(_C.rasterize_gaussians_backward_factors, args, s.debug, "snapshot_bw.dump")
            (g_means2D, g_colors, g_opacities, g_means3D, g_cov3D, _no_sh, g_flows, g_ts, g_scales, g_scales_t,
             g_rot, g_rot_r, factors) = out
            step.record_view(factors, s, dict(means3D=means3D, ts=ts, scales=scales, scales_t=scales_t,
                                              rotations=rotations, rotations_r=rotations_r))
            g_sh = None
        else:
            (g_means2D, g_colors, g_opacities, g_means3D, g_cov3D, g_sh, g_flows, g_ts, g_scales, g_scales_t,
             g_rot, g_rot_r) = _invoke(_C.rasterize_gaussians_backward, args, s.debug, "snapshot_bw.dump")

        def for_input(grad, inp):
            # inputs passed as the empty "not provided" placeholder take no gradient
            return grad if (inp.numel() > 0 and grad is not None) else None

        # one entry per forward input (reference: :208-223); prefilter_var / settings get None
        return (g_means3D, g_means2D, for_input(g_sh, sh), for_input(g_colors, colors_precomp),
                for_input(g_flows, flow_2d), g_opacities, for_input(g_ts, ts), for_input(g_scales, scales),
                for_input(g_scales_t, scales_t), for_input(g_rot, rotations), for_input(g_rot_r, rotations_r),
                for_input(g_cov3D, cov3Ds_precomp), None, None)


class _RasterizeGaussiansRaw(torch.autograd.Function):
    """Raw-parameter entry (no reference counterpart): the inputs are the optimiser's RAW
    parameters -- log-scales, un-normalised quaternions, opacity logits, and the SH row as the two tensors
    features_dc [P,1,3] / features_rest [P,M-1,3] -- and the gradients come back w.r.t. exactly those, so the
    exp / F.normalize x2 / sigmoid kernels and the 1.15 GB torch.cat of the reference's getters
    (scene/gaussian_model.py:179-219), with their autograd mirror images, disappear from the frame.
    Inputs, in order: means3D, means2D, features_dc, features_rest, flow_2d, opacity_logits, ts, log_scales,
    log_scales_t, rotations_raw, rotations_r_raw, prefilter_var, raster_settings."""

    @staticmethod
    def forward(ctx, means3D, means2D, sh_dc, sh_rest, flow_2d, opacities, ts, scales, scales_t, rotations,
                rotations_r, prefilter_var, raster_settings):
        s = raster_settings
        none = torch.Tensor([])
        args = (s.bg, means3D, none, flow_2d, opacities, ts, scales, scales_t, rotations, rotations_r,
                s.scale_modifier, none, prefilter_var, s.viewmatrix, s.projmatrix, s.tanfovx, s.tanfovy,
                s.image_height, s.image_width, sh_dc, s.sh_degree, s.sh_degree_t, s.campos, s.timestamp,
                s.time_duration, s.rot_4d, s.gaussian_dim, s.force_sh_3d, s.prefiltered, s.debug, sh_rest, True)
        (num_rendered, color, flow, depth, T, radii, geom_buf, binning_buf, img_buf, covs_com,
         out_means3D) = _invoke(_C.rasterize_gaussians_raw, args, s.debug, "snapshot_fw.dump")
        _notify_forward(radii)
        ctx.raster_settings = s
        ctx.num_rendered = num_rendered
        ctx.prefilter_var = prefilter_var
        ctx.save_for_backward(means3D, out_means3D, scales, rotations, radii, sh_dc, sh_rest, flow_2d, opacities, ts,
                              scales_t, rotations_r, geom_buf, binning_buf, img_buf)
        ctx.mark_non_differentiable(radii)
        ctx.set_materialize_grads(False)
        ctx.image_shape = tuple(color.shape)
        return color, radii, depth, 1 - T, flow, covs_com

    @staticmethod
    def backward(ctx, grad_out_color, grad_radii, grad_depth, grad_alpha, grad_flow, grad_covs_com):
        s = ctx.raster_settings
        (means3D, out_means3D, scales, rotations, radii, sh_dc, sh_rest, flow_2d, opacities, ts, scales_t, rotations_r,
         geom_buf, binning_buf, img_buf) = ctx.saved_tensors
        if grad_out_color is None:
            grad_out_color = torch.zeros(ctx.image_shape, dtype=torch.float32, device=means3D.device)
        none = torch.empty(0, dtype=torch.float32, device=means3D.device)
        grad_depth = none if grad_depth is None else grad_depth
        grad_alpha = none if grad_alpha is None else grad_alpha
        grad_flow = none if grad_flow is None else grad_flow
        from fdgs.dist import ViewParallelStep
        step = ViewParallelStep.current()
        factors = step is not None and step.wants_factors(sh_dc, none) and (ctx.needs_input_grad[2] or ctx.needs_input_grad[3])
        e = torch.Tensor([])
        args = (s.bg, means3D, out_means3D, radii, e, flow_2d, opacities, ts, scales, scales_t, rotations, rotations_r,
                s.scale_modifier, e, ctx.prefilter_var, s.viewmatrix, s.projmatrix, s.tanfovx, s.tanfovy, grad_out_color,
                grad_depth, grad_alpha, grad_flow, sh_dc, s.sh_degree, s.sh_degree_t, s.campos, s.timestamp,
                s.time_duration, s.rot_4d, s.gaussian_dim, s.force_sh_3d, geom_buf, ctx.num_rendered, binning_buf,
                img_buf, s.debug, sh_rest, True, bool(factors))
        (g_means2D, g_colors, g_opacities, g_means3D, g_cov3D, g_sh, g_flows, g_ts, g_scales, g_scales_t, g_rot, g_rot_r,
         sh_factors, g_sh_rest) = _invoke(_C.rasterize_gaussians_backward_raw, args, s.debug, "snapshot_bw.dump")
        if factors:
            step.record_view(sh_factors, s, dict(means3D=means3D, ts=ts, scales=scales, scales_t=scales_t,
                                                 rotations=rotations, rotations_r=rotations_r, raw=True))
            g_sh = g_sh_rest = None

        def for_input(grad, inp):
            return grad if (inp.numel() > 0 and grad is not None) else None

        return (g_means3D, g_means2D, for_input(g_sh, sh_dc), for_input(g_sh_rest, sh_rest), for_input(g_flows, flow_2d),
                g_opacities, for_input(g_ts, ts), for_input(g_scales, scales), for_input(g_scales_t, scales_t),
                for_input(g_rot, rotations), for_input(g_rot_r, rotations_r), None, None)


def rasterize_gaussians(means3D, means2D, sh, colors_precomp, flow_2d, opacities, ts, scales, scales_t,
                        rotations, rotations_r, cov3Ds_precomp, prefilter_var, raster_settings):
    """Functional entry point (reference: :34-65)."""
    return _RasterizeGaussians.apply(means3D, means2D, sh, colors_precomp, flow_2d, opacities, ts, scales,
                                     scales_t, rotations, rotations_r, cov3Ds_precomp, prefilter_var,
                                     raster_settings)


class GaussianRasterizer(nn.Module):
    """reference: :247-318"""

    def __init__(self, raster_settings):
        super().__init__()
        self.raster_settings = raster_settings

    def markVisible(self, positions):
        """bool[P]: view-space z > 0.2 (reference: :252-261, auxiliary.h:140-163)."""
        with torch.no_grad():
            return _C.mark_visible(positions, self.raster_settings.viewmatrix, self.raster_settings.projmatrix)

    def forward_raw(self, means3D, means2D, opacity_logits, features_dc, features_rest, log_scales, rotations_raw,
                    flow_2d=None, ts=None, log_scales_t=None, rotations_r_raw=None, prefilter_var=-1.0):
        """Same 6-tuple as forward(), from the optimiser's RAW parameters (see _RasterizeGaussiansRaw): activations and
        the SH concatenation happen inside the kernels, gradients arrive at the raw leaves directly."""
        if self.raster_settings.rot_4d and (rotations_r_raw is None or log_scales_t is None or ts is None):
            raise Exception(
                'Please provide exactly rotations_r and scales_t and ts if rot_4d and cov3D_precomp is None!')

        def opt(t):
            return torch.Tensor([]) if t is None else t

        return _RasterizeGaussiansRaw.apply(means3D, means2D, features_dc, opt(features_rest), opt(flow_2d), opacity_logits,
                                            opt(ts), log_scales, opt(log_scales_t), rotations_raw, opt(rotations_r_raw),
                                            prefilter_var, self.raster_settings)

    def forward(self, means3D, means2D, opacities, shs=None, colors_precomp=None, flow_2d=None, ts=None,
                scales=None, scales_t=None, rotations=None, rotations_r=None, cov3D_precomp=None,
                prefilter_var=-1.0):
        # the three argument checks of the reference (:271-280), same messages
        if (shs is None) == (colors_precomp is None):
            raise Exception('Please provide excatly one of either SHs or precomputed colors!')
        have_sr = scales is not None and rotations is not None
        any_sr = scales is not None or rotations is not None
        if (not have_sr and cov3D_precomp is None) or (any_sr and cov3D_precomp is not None):
            raise Exception('Please provide exactly one of either scale/rotation pair or precomputed 3D covariance!')
        if self.raster_settings.rot_4d and cov3D_precomp is None and (
                rotations_r is None or scales_t is None or ts is None):
            raise Exception(
                'Please provide exactly rotations_r and scales_t and ts if rot_4d and cov3D_precomp is None!')

        # "not provided" travels as an empty CPU tensor, i.e. a null pointer in C (reference: :282-300)
        def opt(t):
            return torch.Tensor([]) if t is None else t

        return rasterize_gaussians(means3D, means2D, opt(shs), opt(colors_precomp), opt(flow_2d), opacities,
                                   opt(ts), opt(scales), opt(scales_t), opt(rotations), opt(rotations_r),
                                   opt(cov3D_precomp), prefilter_var, self.raster_settings)
