"""BASELINE config 4 ("lego" shape) in miniature: the optimisation loop of the reference's train.py:84-249 over a
synthetic D-NeRF-like scene (100k points in (-1.3, 1.3)^3, 800x800, batch of 2 views, hyper-parameters of
configs/dnerf/lego.yaml:36-44), runnable through this package's render() or through the compiled reference
rasterizer.  Shared by tests/test_gpu_parity.py (loss curves must track) and bench.py --workload cfg4 (it/s).
TEST / BENCH INFRASTRUCTURE."""
import numpy as np
import torch

DEV = "cuda:0"


class _Pipe:
    convert_SHs_python = False
    compute_cov3D_python = False
    debug = False
    env_map_res = 0


def lego_setup(P, W, H, seed):
    import math
    from fdgs import synth
    g = torch.Generator().manual_seed(seed)
    cams = []
    for k in range(4):   # cameras on a circle of radius 4 looking at the origin, different timestamps
        ang = 2 * math.pi * k / 4 + 0.3
        eye = torch.tensor([4 * math.cos(ang), 0.6 * (k - 1.5), 4 * math.sin(ang)])
        fwd = torch.nn.functional.normalize(-eye, dim=0)
        right = torch.nn.functional.normalize(torch.linalg.cross(torch.tensor([0.0, 1.0, 0.0]), fwd), dim=0)
        up = torch.linalg.cross(fwd, right)
        R = torch.stack([right, up, fwd], 1)              # camera-to-world rotation (columns = camera axes)
        T = -(R.t() @ eye)
        cams.append(synth.make_camera(W, H, timestamp=0.2 + 0.2 * k, focal_scale=1.1, R=R, T=T))

    def params(scale):
        u = lambda *s: torch.rand(*s, generator=g)
        n = lambda *s: torch.randn(*s, generator=g)
        return dict(xyz=(u(P, 3) * 2.6 - 1.3), t=u(P, 1), log_s=torch.log(0.012 * scale * (0.5 + u(P, 3))),
                    log_st=torch.full((P, 1), math.log(math.sqrt(1.0 / 5))) + 0.1 * n(P, 1),     # gaussian_model.py:280-281
                    rot=torch.nn.functional.normalize(n(P, 4), dim=1), rot_r=torch.nn.functional.normalize(n(P, 4), dim=1),
                    op=torch.logit(torch.full((P, 1), 0.1)) + 0.5 * n(P, 1),                      # gaussian_model.py:286
                    sh=torch.cat([(u(P, 1, 3) - 0.5) / 0.28209479177387814, 0.02 * n(P, 47, 3)], 1))
    return cams, params(1.0), params(1.3)


class LegoModel:
    """The reference's GaussianModel getters (scene/gaussian_model.py:179-219) over raw parameters."""

    def __init__(self, raw):
        self.raw = raw
        self.active_sh_degree, self.active_sh_degree_t = 3, 2
        self.time_duration = [0.0, 1.0]
        self.rot_4d, self.gaussian_dim, self.force_sh_3d, self.prefilter_var = True, 4, False, -1.0
        self.get_max_sh_channels = 48

    get_xyz = property(lambda s: s.raw["xyz"])
    get_t = property(lambda s: s.raw["t"])
    get_scaling = property(lambda s: torch.exp(s.raw["log_s"]))
    get_scaling_t = property(lambda s: torch.exp(s.raw["log_st"]))
    get_rotation = property(lambda s: torch.nn.functional.normalize(s.raw["rot"]))
    get_rotation_r = property(lambda s: torch.nn.functional.normalize(s.raw["rot_r"]))
    get_opacity = property(lambda s: torch.sigmoid(s.raw["op"]))
    get_features = property(lambda s: s.raw["sh"])


def lego_render(model, cam, impl):
    """Both arms run the SAME Python prologue (getters, activations, settings tuple, zero screen-space / flow tensors --
    the reference's render(), gaussian_renderer/__init__.py:19-194); only the rasterizer underneath differs."""
    if impl == "ours":
        from gaussian_renderer import render
        return render(cam, model, _Pipe(), torch.zeros(3, device=DEV))["render"]
    return _ref_render(cam, model, torch.zeros(3, device=DEV))["render"]


def _ref_render(cam, pc, bg_color):
    """render() with the compiled reference rasterizer underneath (oracle/ref_api.py); mirrors
    4d-gaussian-splatting_b200/gaussian_renderer/__init__.py op for op for the all-CUDA branch."""
    import math
    import ref_api
    xyz = pc.get_xyz
    screenspace_points = torch.zeros_like(xyz, dtype=xyz.dtype, requires_grad=True, device=xyz.device) + 0
    try:
        screenspace_points.retain_grad()
    except Exception:
        pass
    st = dict(image_height=int(cam.image_height), image_width=int(cam.image_width), tanfovx=math.tan(cam.FoVx * 0.5),
              tanfovy=math.tan(cam.FoVy * 0.5), bg=bg_color, scale_modifier=1.0, viewmatrix=cam.world_view_transform,
              projmatrix=cam.full_proj_transform, sh_degree=pc.active_sh_degree, sh_degree_t=pc.active_sh_degree_t,
              campos=cam.camera_center, timestamp=cam.timestamp, time_duration=pc.time_duration[1] - pc.time_duration[0],
              rot_4d=pc.rot_4d, gaussian_dim=pc.gaussian_dim, force_sh_3d=pc.force_sh_3d, prefiltered=False, debug=False)
    opacity = pc.get_opacity
    scales, rotations = pc.get_scaling, pc.get_rotation
    scales_t, ts, rotations_r = pc.get_scaling_t, pc.get_t, pc.get_rotation_r
    shs = pc.get_features
    flow_2d = torch.zeros_like(xyz[:, :2])
    color, radii, depth, alpha, flow, _ = ref_api.rasterize(st, xyz, screenspace_points, opacity, shs, flow_2d, ts, scales,
                                                            scales_t, rotations, rotations_r)
    return {"render": color, "viewspace_points": screenspace_points, "visibility_filter": radii > 0, "radii": radii,
            "depth": depth, "alpha": alpha, "flow": flow}


def ssim_torch(img1, img2):
    """the reference's ssim() (utils/loss_utils.py:39-64) restated: five grouped 11x11 Gaussian convolutions"""
    import math
    import torch.nn.functional as F
    ch = img1.size(-3)
    g = torch.tensor([math.exp(-(x - 5) ** 2 / float(2 * 1.5 ** 2)) for x in range(11)])
    g = (g / g.sum()).unsqueeze(1)
    w = g.mm(g.t()).float().unsqueeze(0).unsqueeze(0).expand(ch, 1, 11, 11).contiguous().to(img1.device)
    mu1 = F.conv2d(img1, w, padding=5, groups=ch)
    mu2 = F.conv2d(img2, w, padding=5, groups=ch)
    mu1_sq, mu2_sq, mu1_mu2 = mu1.pow(2), mu2.pow(2), mu1 * mu2
    s1 = F.conv2d(img1 * img1, w, padding=5, groups=ch) - mu1_sq
    s2 = F.conv2d(img2 * img2, w, padding=5, groups=ch) - mu2_sq
    s12 = F.conv2d(img1 * img2, w, padding=5, groups=ch) - mu1_mu2
    C1, C2 = 0.01 ** 2, 0.03 ** 2
    return (((2 * mu1_mu2 + C1) * (2 * s12 + C2)) / ((mu1_sq + mu2_sq + C1) * (s1 + s2 + C2))).mean()


def ref_render(cam, pc, bg_color):
    """public name of _ref_render (bench.py train-step leg)"""
    return _ref_render(cam, pc, bg_color)


def make_optimizer(raw):
    """Adam over the raw parameters with the learning rates of configs/dnerf/lego.yaml:36-44 (eps 1e-15 as
    scene/gaussian_model.py:331-357)."""
    return torch.optim.Adam([{"params": [raw["xyz"]], "lr": 1.6e-4}, {"params": [raw["t"]], "lr": 1.6e-4},
                             {"params": [raw["sh"]], "lr": 2.5e-3}, {"params": [raw["op"]], "lr": 5e-2},
                             {"params": [raw["log_s"], raw["log_st"]], "lr": 5e-3},
                             {"params": [raw["rot"], raw["rot_r"]], "lr": 1e-3}], eps=1e-15)


_PYPREP = None


def _pyprep():
    """gaussian_renderer/pyprep.py by file path (importing the package would load the product's CUDA libraries)"""
    global _PYPREP
    if _PYPREP is None:
        import importlib.util
        import os
        path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "4d-gaussian-splatting_b200",
                            "gaussian_renderer", "pyprep.py")
        spec = importlib.util.spec_from_file_location("fdgs_pyprep_standalone_lego", path)
        _PYPREP = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(_PYPREP)
    return _PYPREP


def rigid_loss(model, impl, k=20):
    """train.py:132-152 (configs/dnerf/lego.yaml:58 lambda_rigid = 1): neighbours of every centre among all centres,
    weight exp(-100 d^2), penalise differences of the per-Gaussian velocity (mean offset over dt = 0.1).
    ours: the uniform-grid search (fdgs.knn); reference arm: the brute-force scan pointops2's knnquery performs
    (pointops2/src/knnquery/knnquery_cuda_kernel.cu:65-107 -- that extension cannot be built here, the same O(n^2)
    algorithm runs through csrc/knn.cu's brute_force kernel)."""
    from fdgs.knn import knn           # (with --rigid the reference arm maps libfdgs.so for the brute-force kNN kernel)
    pyprep = _pyprep()
    xyz = model.get_xyz
    idx, dist = knn(xyz[None].contiguous().detach(), xyz[None].contiguous().detach(), k, brute_force=(impl != "ours"))
    _, velocity = pyprep.conditional_covariance_and_offset(torch.cat([model.get_scaling, model.get_scaling_t], 1), 1.0,
                                                           model.raw["rot"], model.raw["rot_r"], 0.1)
    weight = torch.exp(-100 * dist)
    vel_dist = torch.norm(velocity[idx] - velocity[None, :, None], p=2, dim=-1)
    return (weight * vel_dist).sum() / k / xyz.shape[0]


def train_iteration(model, opt, cams, gts, it, batch, impl, device=DEV, lambda_rigid=0.0):
    """One iteration of train.py's loop: `batch` sequential views, L1 loss / batch, one Adam step."""
    total = torch.zeros((), device=device)
    for b in range(batch):
        k = (it * batch + b) % len(cams)
        img = lego_render(model, cams[k], "ours" if impl == "ours" else "ref")
        loss = (img - gts[k]).abs().mean()
        if lambda_rigid > 0:
            loss = loss + lambda_rigid * rigid_loss(model, impl)
        loss = loss / batch
        loss.backward()
        total += loss.detach()
    opt.step()
    opt.zero_grad(set_to_none=True)
    return total
