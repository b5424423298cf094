"""GPU (-m gpu): INTEGRATION.md mode C -- the C-ABI driven directly, no torch types in the call: ctypes structs of raw
device pointers, a C callback allocator handing out the three scratch buffers, the CUDA stream as void*.  Results must
be bit-identical to the torch extension path (which sits on the same entry points)."""
import ctypes

import pytest
import torch

import helpers

pytestmark = pytest.mark.gpu
DEV = "cuda:0"

_fp = ctypes.POINTER(ctypes.c_float)
_ip = ctypes.POINTER(ctypes.c_int)


class FwdArgs(ctypes.Structure):       # include/fdgs.h: fdgs_forward_args, field for field
    _fields_ = [("P", ctypes.c_int), ("D", ctypes.c_int), ("D_t", ctypes.c_int), ("M", ctypes.c_int),
                ("background", _fp), ("width", ctypes.c_int), ("height", ctypes.c_int),
                ("means3D", _fp), ("shs", _fp), ("colors_precomp", _fp), ("flows_precomp", _fp), ("opacities", _fp),
                ("ts", _fp), ("scales", _fp), ("scales_t", _fp), ("scale_modifier", ctypes.c_float),
                ("rotations", _fp), ("rotations_r", _fp), ("cov3D_precomp", _fp), ("prefilter_var", ctypes.c_float),
                ("viewmatrix", _fp), ("projmatrix", _fp), ("cam_pos", _fp), ("timestamp", ctypes.c_float),
                ("time_duration", ctypes.c_float), ("rot_4d", ctypes.c_int), ("gaussian_dim", ctypes.c_int),
                ("force_sh_3d", ctypes.c_int), ("tan_fovx", ctypes.c_float), ("tan_fovy", ctypes.c_float),
                ("prefiltered", ctypes.c_int), ("debug", ctypes.c_int),
                ("out_means3D", _fp), ("out_color", _fp), ("out_flow", _fp), ("out_depth", _fp), ("out_T", _fp),
                ("radii", _ip), ("raw_params", ctypes.c_int), ("shs_rest", _fp)]


class FwdResult(ctypes.Structure):     # fdgs_forward_result
    _fields_ = [("num_rendered", ctypes.c_int), ("geom_buffer", ctypes.c_void_p), ("binning_buffer", ctypes.c_void_p),
                ("image_buffer", ctypes.c_void_p), ("geom_bytes", ctypes.c_size_t), ("binning_bytes", ctypes.c_size_t),
                ("image_bytes", ctypes.c_size_t), ("cov3D", _fp)]


class BwdArgs(ctypes.Structure):       # fdgs_backward_args
    _fields_ = [("P", ctypes.c_int), ("D", ctypes.c_int), ("D_t", ctypes.c_int), ("M", ctypes.c_int), ("R", ctypes.c_int),
                ("background", _fp), ("width", ctypes.c_int), ("height", ctypes.c_int),
                ("out_means3D", _fp), ("shs", _fp), ("colors_precomp", _fp), ("flows_2d", _fp), ("opacities", _fp),
                ("ts", _fp), ("scales", _fp), ("scales_t", _fp), ("scale_modifier", ctypes.c_float),
                ("rotations", _fp), ("rotations_r", _fp), ("cov3D_precomp", _fp), ("prefilter_var", ctypes.c_float),
                ("viewmatrix", _fp), ("projmatrix", _fp), ("campos", _fp), ("timestamp", ctypes.c_float),
                ("time_duration", ctypes.c_float), ("rot_4d", ctypes.c_int), ("gaussian_dim", ctypes.c_int),
                ("force_sh_3d", ctypes.c_int), ("tan_fovx", ctypes.c_float), ("tan_fovy", ctypes.c_float),
                ("radii", _ip), ("geom_buffer", ctypes.c_void_p), ("binning_buffer", ctypes.c_void_p),
                ("image_buffer", ctypes.c_void_p),
                ("dL_dpix", _fp), ("dL_depths", _fp), ("dL_masks", _fp), ("dL_dpix_flow", _fp), ("debug", ctypes.c_int),
                ("dL_dmean2D", _fp), ("dL_dconic", _fp), ("dL_dopacity", _fp), ("dL_dcolor", _fp), ("dL_dflows", _fp),
                ("dL_dmean3D", _fp), ("dL_dcov3D", _fp), ("dL_dsh", _fp), ("dL_dts", _fp), ("dL_dscale", _fp),
                ("dL_dscale_t", _fp), ("dL_drot", _fp), ("dL_drot_r", _fp), ("sh_factors", _fp),
                ("raw_params", ctypes.c_int), ("shs_rest", _fp), ("dL_dsh_rest", _fp)]


ALLOC_FN = ctypes.CFUNCTYPE(ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t)


def _p(t, typ=_fp):
    return ctypes.cast(ctypes.c_void_p(t.data_ptr()), typ) if (t is not None and t.numel()) else typ()


@pytest.mark.parametrize("name", ["small", "rotcam"])
def test_c_abi_through_ctypes_equals_extension_path(name):
    import fdgs
    lib = fdgs.lib()
    C = fdgs.ext()
    cfg, cam, sc, st = helpers.build(name, device=DEV)
    P, W, H = cfg["P"], cfg["W"], cfg["H"]
    f32 = dict(dtype=torch.float32, device=DEV)
    out = dict(means=torch.empty(P, 3, **f32), color=torch.empty(3, H, W, **f32), flow=torch.empty(2, H, W, **f32),
               depth=torch.empty(1, H, W, **f32), T=torch.empty(1, H, W, **f32),
               radii=torch.empty(P, dtype=torch.int32, device=DEV))
    a = FwdArgs()
    a.P, a.D, a.D_t, a.M = P, st["sh_degree"], st["sh_degree_t"], sc.shs.shape[1]
    a.background, a.width, a.height = _p(st["bg"]), W, H
    a.means3D, a.shs, a.flows_precomp, a.opacities = _p(sc.means3D), _p(sc.shs), _p(sc.flow_2d), _p(sc.opacities)
    a.ts, a.scales, a.scales_t, a.scale_modifier = _p(sc.ts), _p(sc.scales), _p(sc.scales_t), st["scale_modifier"]
    a.rotations, a.rotations_r, a.prefilter_var = _p(sc.rotations), _p(sc.rotations_r), -1.0
    a.viewmatrix, a.projmatrix, a.cam_pos = _p(st["viewmatrix"]), _p(st["projmatrix"]), _p(st["campos"])
    a.timestamp, a.time_duration, a.rot_4d, a.gaussian_dim = st["timestamp"], st["time_duration"], 1, 4
    a.tan_fovx, a.tan_fovy = st["tanfovx"], st["tanfovy"]
    a.out_means3D, a.out_color, a.out_flow, a.out_depth, a.out_T = (_p(out["means"]), _p(out["color"]), _p(out["flow"]),
                                                                    _p(out["depth"]), _p(out["T"]))
    a.radii = _p(out["radii"], _ip)

    keep = []      # the caller owns the scratch memory: the callback only allocates

    @ALLOC_FN
    def alloc(ctx, nbytes):
        t = torch.empty(max(int(nbytes), 1), dtype=torch.uint8, device=DEV)
        keep.append(t)
        return t.data_ptr()

    res = FwdResult()
    stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    lib.fdgs_forward.restype = ctypes.c_int
    rc = lib.fdgs_forward(ctypes.byref(a), alloc, None, alloc, None, alloc, None, stream, ctypes.byref(res))
    assert rc == 0, lib.fdgs_last_error()
    fw = C.rasterize_gaussians(*helpers.fwd_args(st, sc, cfg))
    torch.cuda.synchronize()
    assert res.num_rendered == fw[0]
    for k, i in (("color", 1), ("flow", 2), ("depth", 3), ("T", 4), ("radii", 5), ("means", 10)):
        assert torch.equal(out[k], fw[i]), k
    assert res.geom_bytes == lib.fdgs_geom_bytes(P) and res.image_bytes == lib.fdgs_image_bytes(W, H)

    # backward through the C-ABI with the scratch buffers the callback handed out
    gc, gd, ga, gf = helpers.pixel_grads(cfg, device=DEV)
    M = sc.shs.shape[1]
    g = dict(m2=torch.zeros(P, 3, **f32), conic=torch.zeros(P, 4, **f32), op=torch.zeros(P, **f32), col=torch.zeros(P, 3, **f32),
             fl=torch.zeros(P, 2, **f32), m3=torch.empty(P, 3, **f32), cov=torch.empty(P, 6, **f32), sh=torch.empty(P, M, 3, **f32),
             ts=torch.empty(P, **f32), sc=torch.empty(P, 3, **f32), sct=torch.empty(P, **f32), rot=torch.empty(P, 4, **f32),
             rotr=torch.empty(P, 4, **f32))
    b = BwdArgs()
    b.P, b.D, b.D_t, b.M, b.R = P, a.D, a.D_t, M, res.num_rendered
    b.background, b.width, b.height = a.background, W, H
    b.out_means3D, b.shs, b.flows_2d, b.opacities = _p(out["means"]), a.shs, a.flows_precomp, a.opacities
    b.ts, b.scales, b.scales_t, b.scale_modifier = a.ts, a.scales, a.scales_t, a.scale_modifier
    b.rotations, b.rotations_r, b.prefilter_var = a.rotations, a.rotations_r, -1.0
    b.viewmatrix, b.projmatrix, b.campos = a.viewmatrix, a.projmatrix, a.cam_pos
    b.timestamp, b.time_duration, b.rot_4d, b.gaussian_dim = a.timestamp, a.time_duration, 1, 4
    b.tan_fovx, b.tan_fovy, b.radii = a.tan_fovx, a.tan_fovy, a.radii
    b.geom_buffer, b.binning_buffer, b.image_buffer = res.geom_buffer, res.binning_buffer, res.image_buffer
    b.dL_dpix, b.dL_depths, b.dL_masks, b.dL_dpix_flow = _p(gc), _p(gd), _p(ga), _p(gf)
    b.dL_dmean2D, b.dL_dconic, b.dL_dopacity, b.dL_dcolor, b.dL_dflows = _p(g["m2"]), _p(g["conic"]), _p(g["op"]), _p(g["col"]), _p(g["fl"])
    b.dL_dmean3D, b.dL_dcov3D, b.dL_dsh, b.dL_dts = _p(g["m3"]), _p(g["cov"]), _p(g["sh"]), _p(g["ts"])
    b.dL_dscale, b.dL_dscale_t, b.dL_drot, b.dL_drot_r = _p(g["sc"]), _p(g["sct"]), _p(g["rot"]), _p(g["rotr"])
    lib.fdgs_backward.restype = ctypes.c_int
    rc = lib.fdgs_backward(ctypes.byref(b), stream)
    assert rc == 0, lib.fdgs_last_error()
    bw = C.rasterize_gaussians_backward(*helpers.bwd_args(st, sc, cfg, fw, (gc, gd, ga, gf)))
    torch.cuda.synchronize()
    for key, i in (("m2", 0), ("col", 1), ("op", 2), ("m3", 3), ("sh", 5), ("fl", 6), ("ts", 7), ("sc", 8), ("rot", 10)):
        assert helpers.l2_rel(helpers.to_np(g[key]).reshape(-1), helpers.to_np(bw[i]).reshape(-1)) < 1e-5, key
    # error path: a missing output pointer is reported, not dereferenced
    a.out_color = _fp()
    assert lib.fdgs_forward(ctypes.byref(a), alloc, None, alloc, None, alloc, None, stream, ctypes.byref(res)) == 1
    assert b"output" in lib.fdgs_last_error() or b"background" in lib.fdgs_last_error()
