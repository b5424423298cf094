"""CPU, world_size 2 over gloo: the host logic of the view-parallel (data-parallel) path --
view sharding, SUM all-reduce of the per-Gaussian gradients, and the reference's semantics for the
merged densification statistics (train.py:168-183): sum of per-view gradient norms, visibility
count = SUM, radii = MAX.  The renderer is replaced by a differentiable stand-in so that no GPU is
needed; the collective code under test is exactly what bench.py / training use on NCCL."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from fdgs import dist as fdist


def test_shard_views_partitions_the_batch():
    for n in (1, 2, 7, 8, 9, 24):
        for world in (1, 2, 3, 8):
            parts = [fdist.shard_views(n, r, world) for r in range(world)]
            flat = [v for p in parts for v in p]
            assert flat == list(range(n))
            assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1


def _fake_render(params, view):
    """Differentiable stand-in for render(): per-view weights, per-view visibility and radii."""
    g = torch.Generator().manual_seed(100 + view)
    P = params["xyz"].shape[0]
    w = torch.rand(P, 3, generator=g)
    vis = torch.rand(P, generator=g) > 0.4
    radii = (torch.randint(1, 50, (P,), generator=g) * vis).to(torch.int32)
    screen = torch.zeros(P, 3, requires_grad=True)
    img = ((params["xyz"] * w).sum(1) * vis + (screen[:, :2] * w[:, :2]).sum(1) * vis) * params["opacity"][:, 0]
    return {"render": img, "viewspace_points": screen, "radii": radii, "visibility_filter": radii > 0}


def _sequential_reference(P, num_views):
    """What the reference's single-GPU loop computes (train.py:104-183)."""
    torch.manual_seed(0)
    params = {"xyz": torch.randn(P, 3, requires_grad=True), "opacity": torch.rand(P, 1, requires_grad=True)}
    stats = fdist.ViewBatchStats(P, "cpu")
    for v in range(num_views):
        pkg = _fake_render(params, v)
        (pkg["render"].sum() / num_views).backward()
        stats.add_view(pkg["viewspace_points"].grad, pkg["radii"])
    return params, stats


def _worker(rank, world, port, P, num_views, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(0)      # identical replicas on every rank
        params = {"xyz": torch.randn(P, 3, requires_grad=True), "opacity": torch.rand(P, 1, requires_grad=True)}
        views = fdist.shard_views(num_views, rank, world)
        res = fdist.render_view_batch(lambda v: _fake_render(params, v), views, params,
                                      lambda pkg, v: pkg["render"].sum(), global_batch=num_views)
        st = res["stats"]
        out[rank] = (params["xyz"].grad.clone(), params["opacity"].grad.clone(), st.grad_norm_sum.clone(),
                     st.visibility_count.clone(), st.max_radii.clone(), float(res["loss"]))
    finally:
        dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("num_views", [2, 5])
def test_two_ranks_match_the_sequential_loop(num_views):
    P, world = 257, 2
    ref_params, ref_stats = _sequential_reference(P, num_views)
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), P, num_views, out), nprocs=world, join=True)
    assert len(out) == world
    for r in range(world):
        gx, go, gn, vc, mr, loss = out[r]
        assert torch.allclose(gx, ref_params["xyz"].grad, rtol=1e-5, atol=1e-6)
        assert torch.allclose(go, ref_params["opacity"].grad, rtol=1e-5, atol=1e-6)
        assert torch.allclose(gn, ref_stats.grad_norm_sum, rtol=1e-5, atol=1e-6)   # norm per view, then sum
        assert torch.equal(vc, ref_stats.visibility_count)                        # SUM
        assert torch.equal(mr, ref_stats.max_radii)                               # MAX
    assert out[0][5] == pytest.approx(out[1][5])


def _sparse_worker(rank, world, port, P, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(7 + rank)
        vis = torch.rand(P, generator=g) < 0.3                      # this rank's rendered Gaussians
        grads = [torch.randn(P, 3, generator=g) * vis[:, None], torch.randn(P, 5, 3, generator=g) * vis[:, None, None],
                 None, torch.randn(P, 1, generator=g) * vis[:, None]]
        dense = [None if x is None else x.clone() for x in grads]
        union = vis.to(torch.int32)
        dist.all_reduce(union, op=dist.ReduceOp.MAX)
        fdist.allreduce_gradients(grads, union_visible=union > 0)    # sparse exchange
        fdist.allreduce_gradients(dense)                            # dense exchange
        same = all(torch.equal(a, b) for a, b in zip(grads, dense) if a is not None)
        # above the density threshold the sparse call must take the dense path and still be right
        g2 = [x.clone() for x in dense if x is not None]
        fdist.allreduce_gradients(g2, union_visible=torch.ones(P, dtype=torch.bool), dense_above=0.5)
        out[rank] = (same, float((union > 0).float().mean()), all(torch.allclose(a, world * b) for a, b in zip(g2, [x for x in dense if x is not None])))
    finally:
        dist.destroy_process_group()


def test_sparse_union_allreduce_equals_dense():
    """The union-of-visibility exchange (one flat buffer over the rendered rows) gives bit-identical sums to
    the dense all-reduce: rows outside the union are zero on every rank."""
    world, P = 2, 1000
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_sparse_worker, args=(world, _free_port(), P, out), nprocs=world, join=True)
    assert len(out) == world
    for r in range(world):
        same, frac, dense_ok = out[r]
        assert same and dense_ok
        assert 0.3 < frac < 0.6


# ---------------------------------------------------------------------------------------------------
# SH colour-factor exchange (ViewParallelStep): the host logic -- union / slots / table layout / view order --
# on gloo, with the CUDA reconstruction kernel replaced by its PyTorch reference (tests/sh_outer_ref.py).
# ---------------------------------------------------------------------------------------------------
class _S:
    """the per-step constants the step reads from a view's raster settings"""
    def __init__(self, timestamp, campos):
        self.timestamp, self.campos = timestamp, campos
        self.scale_modifier, self.time_duration, self.rot_4d, self.gaussian_dim, self.force_sh_3d = 1.0, 1.0, True, 4, False
        self.sh_degree, self.sh_degree_t = 3, 2


def _factor_scene(P):
    g = torch.Generator().manual_seed(5)
    n = lambda *s: torch.randn(*s, generator=g)
    inputs = dict(means3D=n(P, 3) + torch.tensor([0.0, 0.0, 5.0]), ts=torch.rand(P, 1, generator=g),
                  scales=0.05 * torch.rand(P, 3, generator=g) + 0.01, scales_t=0.2 * torch.rand(P, 1, generator=g) + 0.1,
                  rotations=torch.nn.functional.normalize(n(P, 4), dim=1),
                  rotations_r=torch.nn.functional.normalize(n(P, 4), dim=1))
    return inputs


def _factor_view(P, v):
    g = torch.Generator().manual_seed(900 + v)
    vis = torch.rand(P, generator=g) < 0.12
    radii = (torch.randint(1, 40, (P,), generator=g) * vis).to(torch.int32)
    factors = torch.randn(P, 3, generator=g) * vis[:, None]
    geo = torch.randn(P, 3, generator=g) * vis[:, None]
    screen = torch.randn(P, 3, generator=g) * vis[:, None]
    return factors, radii, geo, screen, _S(0.1 + 0.2 * v, torch.tensor([0.1 * v, -0.05 * v, 0.02 * v]))


def _outer_ref_fn(inputs):
    import sh_outer_ref

    def fn(table, stride, meta_off, V, K, slot_of, outs, views):
        full = sh_outer_ref.sh_outer_sum_ref(table, stride, meta_off, V, K, slot_of, inputs["means3D"], inputs["ts"],
                                             inputs["scales"], inputs["scales_t"], inputs["rotations"],
                                             inputs["rotations_r"], 1.0, 1.0, True, 4, False, 3, 2, 48)
        off = 0
        for o in outs:
            o.copy_(full[:, off:off + o.shape[1]])
            off += o.shape[1]
    return fn


def _run_factor_step(P, view_ids, group_world, split):
    inputs = _factor_scene(P)
    xyz = torch.zeros(P, 3, requires_grad=True)
    shs = [torch.zeros(P, 1, 3, requires_grad=True), torch.zeros(P, 47, 3, requires_grad=True)] if split else \
        [torch.zeros(P, 48, 3, requires_grad=True)]
    step = fdist.ViewParallelStep(P, "cpu", outer_sum_fn=_outer_ref_fn(inputs))
    xyz.grad = torch.zeros(P, 3)
    for v in view_ids:
        factors, radii, geo, screen, S = _factor_view(P, v)
        xyz.grad += geo
        step.record_view(factors, S, inputs)
        step.add_view_stats(screen, radii)
    st = step.finish([xyz], shs, views_per_rank=None)
    return xyz.grad, torch.cat([p.grad for p in shs], 1), st, step.info


def _factor_worker(rank, world, port, P, num_views, split, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        gx, gsh, st, info = _run_factor_step(P, fdist.shard_views(num_views, rank, world), world, split)
        out[rank] = (gx.clone(), gsh.clone(), st.grad_norm_sum.clone(), st.visibility_count.clone(), st.max_radii.clone(), info)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("num_views,split", [(2, False), (4, True), (3, False)])
def test_sh_factor_exchange_matches_single_process(num_views, split):
    """2 ranks x (views / 2) == 1 process x all views: geometry rows, statistics and the rebuilt SH rows, including
    an uneven view split (3 views: the padded zero view of the rank with fewer views contributes nothing)."""
    P, world = 301, 2
    ref_x, ref_sh, ref_st, _ = _run_factor_step(P, list(range(num_views)), 1, split)
    assert float(ref_sh.abs().sum()) > 0
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_factor_worker, args=(world, _free_port(), P, num_views, split, out), nprocs=world, join=True)
    assert len(out) == world
    for r in range(world):
        gx, gsh, gn, vc, mr, info = out[r]
        assert torch.allclose(gx, ref_x, rtol=1e-6, atol=1e-6)
        assert torch.equal(gsh, ref_sh)            # same views in the same order through the same arithmetic
        assert torch.allclose(gn, ref_st.grad_norm_sum, rtol=1e-6, atol=1e-6)
        assert torch.equal(vc, ref_st.visibility_count) and torch.equal(mr, ref_st.max_radii)
        assert info["geometry_path"] == "rows" and info["views_total"] == 2 * ((num_views + 1) // 2)
    assert torch.equal(out[0][1], out[1][1])       # replicas stay bit-identical


def _dense_guard_worker(rank, world, port, P, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(3 + rank)
        vis = torch.rand(P, generator=g) < 0.3
        radii = vis.to(torch.int32) * 5
        dist.all_reduce(radii, op=dist.ReduceOp.MAX)
        # a gradient that is NOT confined to the rendered Gaussians (e.g. a rigidity loss over all of them)
        full = torch.randn(P, 3, generator=g)
        want = full.clone()
        dist.all_reduce(want)
        fdist.allreduce_gradients([full], union_radii=radii)
        out[rank] = torch.equal(full, want)
    finally:
        dist.destroy_process_group()


def test_sparse_exchange_detects_non_rasterizer_gradients():
    """rows outside the union are only zero for rasterizer gradients; a loss over ALL Gaussians must not be
    silently under-reduced -- the device-side guard routes the step to the dense all-reduce."""
    world, P = 2, 500
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_dense_guard_worker, args=(world, _free_port(), P, out), nprocs=world, join=True)
    assert all(out[r] for r in range(world))


def _repair_worker(rank, world, port, P, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(40 + rank)
        vis = torch.rand(P, generator=g) < 0.25
        radii = vis.to(torch.int32) * 7
        xyz = torch.zeros(P, 3, requires_grad=True)
        xyz.grad = torch.randn(P, 3, generator=g)            # NOT confined to the rendered rows (e.g. a rigidity loss)
        want = xyz.grad.clone()
        dist.all_reduce(want)
        step = fdist.ViewParallelStep(P, "cpu", sh_factors=False)
        step.add_view_stats(torch.randn(P, 3, generator=g) * vis[:, None], radii)
        step.finish([xyz], [])
        out[rank] = (torch.allclose(xyz.grad, want, rtol=1e-6, atol=1e-6), step.info["geometry_path"])
    finally:
        dist.destroy_process_group()


def test_view_parallel_step_repairs_non_rasterizer_gradients():
    """the optimistic sparse exchange of ViewParallelStep.finish(): when the device-side guard (read at the end) says
    some rank had gradient rows outside the union, the remaining rows are summed densely -- the result is the dense sum"""
    world, P = 2, 400
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_repair_worker, args=(world, _free_port(), P, out), nprocs=world, join=True)
    for r in range(world):
        ok, path = out[r]
        assert ok and path == "rows + dense repair"
