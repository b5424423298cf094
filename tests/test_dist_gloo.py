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
