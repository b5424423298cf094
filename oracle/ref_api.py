"""Python driver for the compiled, unmodified reference rasterizer (oracle/_ref/ref_rasterizer.so).

TEST / BASELINE INFRASTRUCTURE ONLY (tests/, bench.py --impl reference).  The reference's own
driver file (gaussian_renderer/diff_gaussian_rasterization.py) cannot be imported on the GPU box:
it JIT-compiles from /root/reference at import time (`:17-28`), and that tree does not exist
there.  This module performs the same calls in the same order on the prebuilt module: forward =
`_C.rasterize_gaussians(30 positional args)`, keep the three scratch buffers, backward =
`_C.rasterize_gaussians_backward(37 positional args)` (argument layouts:
diff_gaussian_rasterization.py:88-119 and :154-190).  Nothing of the product is on this path.
"""
import torch

import oracle_py


class _RefRasterize(torch.autograd.Function):
    @staticmethod
    def forward(ctx, means3D, means2D, sh, colors_precomp, flow_2d, opacities, ts, scales, scales_t, rotations,
                rotations_r, cov3Ds_precomp, prefilter_var, s):
        C = oracle_py.ref_module()
        out = C.rasterize_gaussians(
            s["bg"], means3D, colors_precomp, flow_2d, opacities, ts, scales, scales_t, rotations, rotations_r,
            s["scale_modifier"], cov3Ds_precomp, prefilter_var, s["viewmatrix"], s["projmatrix"], s["tanfovx"],
            s["tanfovy"], s["image_height"], s["image_width"], sh, s["sh_degree"], s["sh_degree_t"], s["campos"],
            s["timestamp"], s["time_duration"], s["rot_4d"], s["gaussian_dim"], s["force_sh_3d"], s["prefiltered"],
            s["debug"])
        num_rendered, color, flow, depth, T, radii, geom, binning, img, covs, out_means3D = out
        ctx.s, ctx.num_rendered, ctx.prefilter_var = s, num_rendered, prefilter_var
        ctx.save_for_backward(colors_precomp, means3D, out_means3D, scales, rotations, cov3Ds_precomp, radii, sh, flow_2d,
                              opacities, ts, scales_t, rotations_r, geom, binning, img)
        return color, radii, depth, 1 - T, flow, covs

    @staticmethod
    def backward(ctx, g_color, g_radii, g_depth, g_alpha, g_flow, g_covs):
        C = oracle_py.ref_module()
        s = ctx.s
        (colors_precomp, means3D, out_means3D, scales, rotations, cov3Ds_precomp, radii, sh, flow_2d, opacities, ts,
         scales_t, rotations_r, geom, binning, img) = ctx.saved_tensors
        (g_means2D, g_colors, g_opac, g_means3D, g_cov, g_sh, g_flows, g_ts, g_scales, g_scales_t, g_rot,
         g_rot_r) = C.rasterize_gaussians_backward(
            s["bg"], means3D, out_means3D, radii, colors_precomp, flow_2d, opacities, ts, scales, scales_t, rotations,
            rotations_r, s["scale_modifier"], cov3Ds_precomp, ctx.prefilter_var, s["viewmatrix"], s["projmatrix"],
            s["tanfovx"], s["tanfovy"], g_color, g_depth, g_alpha, g_flow, sh, s["sh_degree"], s["sh_degree_t"],
            s["campos"], s["timestamp"], s["time_duration"], s["rot_4d"], s["gaussian_dim"], s["force_sh_3d"], geom,
            ctx.num_rendered, binning, img, s["debug"])

        def pick(g, inp):
            return g if inp.numel() > 0 else None

        return (g_means3D, g_means2D, pick(g_sh, sh), pick(g_colors, colors_precomp), pick(g_flows, flow_2d), g_opac,
                pick(g_ts, ts), pick(g_scales, scales), pick(g_scales_t, scales_t), pick(g_rot, rotations),
                pick(g_rot_r, rotations_r), pick(g_cov, cov3Ds_precomp), None, None)


def rasterize(settings: dict, means3D, means2D, opacities, shs, flow_2d, ts, scales, scales_t, rotations, rotations_r,
              prefilter_var=-1.0):
    """The reference's GaussianRasterizer.forward for the SH + scale/rotation case (4D or 3D)."""
    e = torch.Tensor([])
    opt = lambda t: e if t is None else t
    return _RefRasterize.apply(means3D, means2D, opt(shs), e, opt(flow_2d), opacities, opt(ts), opt(scales), opt(scales_t),
                               opt(rotations), opt(rotations_r), e, prefilter_var, settings)
