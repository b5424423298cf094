"""GPU (-m gpu): the view-parallel (multi-GPU) path.

  1. one GPU: the SH colour-factor backward + sh_outer_sum reconstruction equals the dense dL_dsh of the ordinary
     backward (one view), and the sum over several views equals accumulating the views' dL_dsh one after the other
     (what the reference's sequential loop does, train.py:104-166);  the reconstruction kernel against its PyTorch
     fp32 reference (tests/sh_outer_ref.py);
  2. two GPUs (skipped on a one-GPU box): N NCCL ranks through ViewParallelStep reproduce the gradients and the
     densification statistics of the sequential single-GPU loop over the same views.
"""
import os
import socket

import pytest
import torch

import helpers

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


class _Model:
    """GaussianModel stand-in over activated parameters; SH optionally split like the reference's
    features_dc / features_rest (scene/gaussian_model.py:210-214)."""

    def __init__(self, sc, split_sh=False):
        leaf = lambda t: t.clone().requires_grad_(True)
        self.get_xyz, self.get_opacity = leaf(sc.means3D), leaf(sc.opacities)
        self.get_scaling, self.get_scaling_t = leaf(sc.scales), leaf(sc.scales_t)
        self.get_rotation, self.get_rotation_r, self.get_t = leaf(sc.rotations), leaf(sc.rotations_r), leaf(sc.ts)
        self.sh_leaves = [leaf(sc.shs[:, :1].contiguous()), leaf(sc.shs[:, 1:].contiguous())] if split_sh else [leaf(sc.shs)]
        self.active_sh_degree, self.active_sh_degree_t = sc.sh_degree, sc.sh_degree_t
        self.time_duration = [0.0, sc.time_duration]
        self.rot_4d, self.gaussian_dim, self.force_sh_3d = sc.rot_4d, sc.gaussian_dim, sc.force_sh_3d
        self.prefilter_var = -1.0
        self.get_max_sh_channels = sc.shs.shape[1]

    @property
    def get_features(self):
        return self.sh_leaves[0] if len(self.sh_leaves) == 1 else torch.cat(self.sh_leaves, 1)

    def geometry(self):
        return [self.get_xyz, self.get_t, self.get_scaling, self.get_scaling_t, self.get_rotation, self.get_rotation_r,
                self.get_opacity]


class _Pipe:
    convert_SHs_python = False
    compute_cov3D_python = False
    debug = False
    env_map_res = 0


def _views(cfg, n, device):
    """n cameras around the configuration's base camera, each with its own timestamp."""
    from fdgs import synth
    out = []
    for v in range(n):
        pose = cfg.get("pose", (0.0, 0.0, 0.0, 0.0, 0.0, 0.0))
        R, T = synth.pose_from_euler(pose[0] + 2.0 * v, pose[1] - 1.0 * v, pose[2], pose[3] + 0.05 * v, pose[4], pose[5] - 0.03 * v)
        dur = cfg.get("time_duration", 1.0)
        out.append(synth.make_camera(cfg["W"], cfg["H"], timestamp=cfg.get("timestamp", 0.5) + 0.07 * dur * (v - 0.5 * n),
                                     negative_fov=cfg.get("negative_fov", False), R=R, T=T).to(device))
    return out


def _sequential(cfg, sc, cams, device, split_sh):
    """the reference's loop: render every view in turn, autograd accumulates (train.py:104-183)"""
    from gaussian_renderer import render
    from fdgs.dist import ViewBatchStats
    pc = _Model(sc, split_sh)
    stats = ViewBatchStats(sc.P, device)
    bg = torch.zeros(3, device=device)
    for k, cam in enumerate(cams):
        pkg = render(cam, pc, _Pipe(), bg)
        g = torch.Generator(device="cpu").manual_seed(1000 + k)
        G = torch.randn(3, cfg["H"], cfg["W"], generator=g).to(device)
        ((pkg["render"] * G).sum() / len(cams)).backward()
        stats.add_view(pkg["viewspace_points"].grad, pkg["radii"])
    return pc, stats


def _view_parallel(cfg, sc, cams, view_ids, device, split_sh, group=None, views_per_rank=None):
    from gaussian_renderer import render
    from fdgs.dist import ViewParallelStep
    pc = _Model(sc, split_sh)
    step = ViewParallelStep(sc.P, device, group=group)
    bg = torch.zeros(3, device=device)
    with step:
        for k in view_ids:
            pkg = render(cams[k], pc, _Pipe(), bg)
            g = torch.Generator(device="cpu").manual_seed(1000 + k)
            G = torch.randn(3, cfg["H"], cfg["W"], generator=g).to(device)
            ((pkg["render"] * G).sum() / len(cams)).backward()
            step.add_view_stats(pkg["viewspace_points"].grad, pkg["radii"])
    stats = step.finish(pc.geometry(), pc.sh_leaves, views_per_rank=views_per_rank)
    return pc, stats, step.info


def _cmp(pa, pb, sa, sb, tol_sh, tol_geo, tol_chain=None):
    """SH rows and blend-level tensors (opacity, screen-gradient norms): tight.  Covariance-chain tensors (xyz, t, scales,
    rotations): both sides run the same non-deterministic blend backward, whose 1e-7 noise the chain amplifies (see
    test_gpu_parity.py) -- tol_chain, which grows with the scene size."""
    tol_chain = tol_geo if tol_chain is None else tol_chain
    for j, (a, b) in enumerate(zip(pa.sh_leaves, pb.sh_leaves)):
        assert a.grad is not None and b.grad is not None
        err = helpers.l2_rel(helpers.to_np(a.grad), helpers.to_np(b.grad))
        if err > tol_sh:
            bad = ((a.grad - b.grad).reshape(a.shape[0], -1).abs().amax(1) > 1e-6 * float(b.grad.abs().max())).nonzero().view(-1)
            raise AssertionError("SH leaf %d: l2 %.3e > %.1e; %d rows differ, first %s, union rows among them: %d" % (
                j, err, tol_sh, bad.numel(), bad[:8].tolist(), int((sb.max_radii[bad] > 0).sum())))
    for i, (a, b) in enumerate(zip(pa.geometry(), pb.geometry())):
        err = helpers.l2_rel(helpers.to_np(a.grad), helpers.to_np(b.grad))
        assert err <= (tol_geo if i == 6 else tol_chain), (i, err)
    assert helpers.l2_rel(helpers.to_np(sa.grad_norm_sum), helpers.to_np(sb.grad_norm_sum)) <= tol_geo
    assert torch.equal(sa.visibility_count.view(-1), sb.visibility_count.view(-1))
    assert torch.equal(sa.max_radii, sb.max_radii)


@pytest.mark.parametrize("name,nviews,split", [("small", 1, False), ("rotcam", 1, True), ("dur10", 3, False),
                                               ("n3v", 4, True), ("sh3d", 2, False), ("mid", 3, True)])
def test_factor_mode_equals_dense_backward_one_gpu(name, nviews, split):
    """world = 1: the factor path (no dense dL_dsh) against ordinary autograd accumulation over the same views.
    Same backward kernels for everything but the SH rows, so the geometry matches to the blend's atomic noise; the SH
    rows are rebuilt with the same device functions and summed in view order."""
    cfg, cam, sc, st = helpers.build(name, device=DEV)
    cams = _views(cfg, nviews, DEV)
    ref, rs = _sequential(cfg, sc, cams, DEV, split)
    got, gs, info = _view_parallel(cfg, sc, cams, list(range(nviews)), DEV, split)
    assert info["views_total"] == nviews
    # the blend backward's RED order differs from run to run -> dL_dcolor (the factor) carries ~1e-7 noise
    _cmp(got, ref, gs, rs, tol_sh=2e-6, tol_geo=1e-4, tol_chain=1e-3 if cfg["P"] < 50000 else 5e-3)
    sh_ref = torch.cat([p.grad for p in ref.sh_leaves], 1)
    sh_got = torch.cat([p.grad for p in got.sh_leaves], 1)
    invisible = rs.max_radii <= 0
    assert float(sh_got[invisible].abs().sum()) == 0.0
    assert float(sh_ref.abs().max()) > 0


@pytest.mark.parametrize("name,nviews,slots,split", [("small", 1, 2, False), ("mid", 1, 2, True), ("mid", 3, 4, True)])
def test_factor_mode_with_empty_view_slots(name, nviews, slots, split):
    """A rank that renders fewer views than the largest shard carries EMPTY view blocks in the factor table (all-zero
    factors and meta): they must contribute nothing.  world = 1 with views_per_rank > rendered views reproduces that."""
    cfg, cam, sc, st = helpers.build(name, device=DEV)
    cams = _views(cfg, nviews, DEV)
    ref, rs = _sequential(cfg, sc, cams, DEV, split)
    got, gs, info = _view_parallel(cfg, sc, cams, list(range(nviews)), DEV, split, views_per_rank=slots)
    assert info["views_total"] == slots
    _cmp(got, ref, gs, rs, tol_sh=2e-6, tol_geo=1e-4, tol_chain=1e-3 if cfg["P"] < 50000 else 5e-3)


def test_sh_outer_sum_kernel_vs_torch_reference():
    """csrc/preprocess_bwd.cu: sh_outer_sum_kernel against tests/sh_outer_ref.py (plain PyTorch fp32) on random
    factors: vector and split-row stores, accumulate, M = 48 / 16 / 4, rows outside the union zeroed."""
    import fdgs
    import sh_outer_ref
    C = fdgs.ext()
    g = torch.Generator().manual_seed(12)
    for (M, D, D_t, split, rot_4d, gdim) in [(48, 3, 2, False, True, 4), (48, 3, 2, True, True, 4), (16, 3, 0, False, True, 4),
                                             (4, 1, 0, False, True, 4), (48, 3, 1, False, False, 4), (16, 3, 0, True, False, 3)]:
        cfg, cam, sc, st = helpers.build(dict(P=5000, W=64, H=64, seed=31 + M + D_t, M=M, sh_degree=D, sh_degree_t=D_t,
                                              rot_4d=rot_4d, gaussian_dim=gdim), device=DEV)
        P, V = sc.P, 3
        union = torch.rand(P, generator=g) < 0.4
        idx = torch.nonzero(union).squeeze(1)
        K = idx.numel()
        meta_off = (3 * K + 3) // 4 * 4
        stride = meta_off + 8
        table = torch.zeros(V, stride)
        for v in range(V):
            f = torch.randn(K, 3, generator=g) * (torch.rand(K, 1, generator=g) < 0.7)
            table[v, :3 * K] = f.reshape(-1)
            table[v, meta_off] = 0.2 + 0.25 * v
            table[v, meta_off + 1:meta_off + 4] = torch.tensor([0.1 * v, -0.2, 0.05 * v])
        slot = torch.where(union, torch.cumsum(union.to(torch.int32), 0, dtype=torch.int32) - 1, torch.full((P,), -1, dtype=torch.int32))
        table, slot = table.to(DEV), slot.to(torch.int32).to(DEV)
        outs = [torch.full((P, 1, 3), 7.0, device=DEV), torch.full((P, M - 1, 3), 7.0, device=DEV)] if split else \
            [torch.full((P, M, 3), 7.0, device=DEV)]
        args = (table, stride, meta_off, V, K, slot, idx.to(DEV), sc.means3D, sc.ts, sc.scales, sc.scales_t, sc.rotations, sc.rotations_r,
                1.0, sc.time_duration, rot_4d, gdim, False, D, D_t)
        C.sh_outer_sum(*args, outs, False)
        got = torch.cat(outs, 1)
        want = sh_outer_ref.sh_outer_sum_ref(table, stride, meta_off, V, K, slot, sc.means3D, sc.ts, sc.scales, sc.scales_t,
                                             sc.rotations, sc.rotations_r, 1.0, sc.time_duration, rot_4d, gdim, False, D, D_t, M)
        assert float(got[~union.to(DEV)].abs().sum()) == 0.0
        assert helpers.l2_rel(helpers.to_np(got), helpers.to_np(want)) < 2e-6, (M, D, D_t, split)
        assert helpers.max_rel(helpers.to_np(got), helpers.to_np(want)) < 2e-5, (M, D, D_t, split)
        C.sh_outer_sum(*args, outs, True)           # accumulate: exactly twice the sum
        assert torch.equal(torch.cat(outs, 1), got + got)


def test_second_device_in_the_same_process():
    """function attributes are per device -- every >48 KB-smem kernel must also launch on cuda:1."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import fdgs
    C = fdgs.ext()
    outs = []
    for dev in ("cuda:0", "cuda:1"):
        cfg, cam, sc, st = helpers.build("small", device=dev)
        fw = C.rasterize_gaussians(*helpers.fwd_args(st, sc, cfg))
        gc = helpers.pixel_grads(cfg, device=dev)
        e = torch.empty(0, device=dev)
        bw = C.rasterize_gaussians_backward(*helpers.bwd_args(st, sc, cfg, fw, (gc[0], e, e, e)))
        torch.cuda.synchronize(dev)
        outs.append((fw[1].cpu(), bw[5].cpu()))
    assert torch.equal(outs[0][0], outs[1][0])
    assert helpers.l2_rel(outs[0][1].numpy(), outs[1][1].numpy()) < 1e-5


# ---------------------------------------------------------------------------------------------------
# 2. N NCCL ranks == the sequential loop
# ---------------------------------------------------------------------------------------------------
def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _nccl_worker(rank, world, port, name, nviews, split, out):
    import torch.distributed as dist
    from fdgs.dist import shard_views
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dev = "cuda:%d" % rank
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", rank=rank, world_size=world)
    try:
        cfg, cam, sc, st = helpers.build(name, device=dev)
        cams = _views(cfg, nviews, dev)
        ref, rs = _sequential(cfg, sc, cams, dev, split)                 # every rank: the whole loop on its own GPU
        got, gs, info = _view_parallel(cfg, sc, cams, shard_views(nviews, rank, world), dev, split)
        torch.cuda.synchronize(dev)
        # cfg5's long time axis: the reference itself reproduces its chain gradients only to 5-30 % (profiles/)
        # tol_sh: the factors of the other rank's views carry that rank's blend-backward RED-order noise (~1e-7 per
        # entry; one of five runs of mid/3 views exceeded 2e-6 on the norm, the other four sat below it) -- a real
        # exchange error (a missing or misplaced view) is O(1)
        _cmp(got, ref, gs, rs, tol_sh=1e-5, tol_geo=1e-4, tol_chain=1e-3 if cfg["P"] < 50000 else (5e-3 if name != "cfg5" else 2.0))
        sh = torch.cat([p.grad for p in got.sh_leaves], 1)
        out[rank] = (info, sh.double().sum().item(), sh.cpu() if sc.P <= 20000 else None)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("name,nviews,split", [("small", 2, False), ("n3v", 4, True), ("mid", 3, True), ("cfg5", 4, False)])
def test_nccl_ranks_reproduce_the_sequential_loop(name, nviews, split):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (run under gpurun --gpus 2)")
    import torch.multiprocessing as mp
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_nccl_worker, args=(world, _free_port(), name, nviews, split, out), nprocs=world, join=True)
    assert len(out) == world
    assert out[0][0]["geometry_path"] == "rows"
    assert out[0][1] == out[1][1]                       # replicas bit-identical
    if out[0][2] is not None:
        assert torch.equal(out[0][2], out[1][2])
