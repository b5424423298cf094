"""Data-parallel rendering over views (frames) -- the only way this path shards.

The reference renders the `batch_size` views of one optimisation step sequentially on one GPU and
lets autograd sum the per-Gaussian gradients (reference: train.py:104-166, loss scaled by
1/batch_size at :162), then merges the densification statistics of the views (train.py:168-183).
Here the views of a step are spread over the ranks (one process per GPU, replicated Gaussians):
every rank renders its own views, then

  * one SUM all-reduce (NCCL over NVLink; gloo in the CPU tests) of the per-Gaussian parameter
    gradients -- the only data-path collective, a real exchange step of the algorithm -- restricted to
    the rows of Gaussians that some view of the step rendered (union of the ranks' visibility, which the
    MAX reduction of the radii below already provides), packed into one flat buffer;
  * the densification statistics are reduced with the reference's semantics:
    sum over views of ||d loss / d mean2D[:, :2]|| (norm per view first, train.py:164,173),
    visibility count = SUM (train.py:169), radii = MAX (train.py:171).

A single-view render never communicates.
"""
from typing import Dict, Iterable, List, Optional, Sequence

import torch
import torch.distributed as dist


def shard_views(num_views: int, rank: int, world: int) -> List[int]:
    """Indices of the views rank `rank` renders: contiguous blocks, sizes differing by at most one."""
    base, rem = divmod(num_views, world)
    start = rank * base + min(rank, rem)
    return list(range(start, start + base + (1 if rank < rem else 0)))


def allreduce_gradients(grads: Sequence[Optional[torch.Tensor]], group=None, average_over: Optional[int] = None,
                        union_visible: Optional[torch.Tensor] = None, dense_above: float = 0.6):
    """In-place SUM all-reduce of the per-Gaussian parameter gradients (None entries are skipped).

    `average_over` (the global batch size) reproduces the reference's loss / batch_size scaling when the
    caller did not already scale its local loss.

    `union_visible` (bool / int [P], IDENTICAL on every rank -- e.g. `ViewBatchStats.max_radii > 0` after its
    MAX reduction) enables the sparse exchange: a Gaussian that was rendered in no view of the step has an
    all-zero gradient row on every rank (the rasterizer writes zeros there), so only the rows of the union are
    packed into one flat buffer ([K x sum of row widths], one block per tensor), all-reduced with ONE
    collective and scattered back.  With 48 SH coefficients a row is 644 bytes and a view renders about a
    third of the Gaussians, so this moves a third of the bytes of the dense all-reduce.  Falls back to the dense
    exchange when the union covers more than `dense_above` of the Gaussians."""
    live = [g for g in grads if g is not None]
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        if average_over:
            for g in live:
                g.div_(average_over)
        return
    P = live[0].shape[0] if live else 0
    idx = None
    if union_visible is not None and live and all(g.shape[0] == P and g.is_contiguous() for g in live):
        idx = torch.nonzero(union_visible.reshape(-1) != 0).squeeze(1)
        if idx.numel() > dense_above * P:
            idx = None
    if idx is None:
        works = [dist.all_reduce(g, op=dist.ReduceOp.SUM, group=group, async_op=True) for g in live]
        for w in works:
            w.wait()
    elif live[0].is_cuda:
        # one gather kernel over all tensors -> one collective -> one scatter kernel (csrc/exchange.cu)
        import fdgs
        C = fdgs.ext()
        flat = C.pack_rows(live, idx)
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
        C.unpack_rows(flat, live, idx)
    else:
        # host tensors (the gloo tests): same exchange with torch indexing
        K = idx.numel()
        rows = [g.view(P, -1) for g in live]
        widths = [r.shape[1] for r in rows]
        flat = torch.empty(K * sum(widths), dtype=live[0].dtype, device=live[0].device)
        blocks, off = [], 0
        for r, c in zip(rows, widths):
            blk = flat[off:off + K * c].view(K, c)
            torch.index_select(r, 0, idx, out=blk)
            blocks.append(blk)
            off += K * c
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
        for r, blk in zip(rows, blocks):
            r.index_copy_(0, idx, blk)
    if average_over:
        for g in live:
            g.div_(average_over)


class ViewBatchStats:
    """Accumulates the reference's per-step densification statistics over the local views and
    reduces them across ranks (reference: train.py:168-183, scene/gaussian_model.py:579-589)."""

    def __init__(self, P: int, device):
        self.grad_norm_sum = torch.zeros(P, 1, device=device)      # sum_views ||viewspace grad[:, :2]||
        self.visibility_count = torch.zeros(P, device=device)      # sum_views (radii > 0)
        self.max_radii = torch.zeros(P, dtype=torch.int32, device=device)

    def add_view(self, viewspace_grad: torch.Tensor, radii: torch.Tensor):
        self.grad_norm_sum += torch.norm(viewspace_grad[:, :2], dim=-1, keepdim=True)
        self.visibility_count += (radii > 0).to(self.visibility_count.dtype)
        self.max_radii = torch.max(self.max_radii, radii.to(torch.int32))

    def reduce(self, group=None):
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            dist.all_reduce(self.grad_norm_sum, op=dist.ReduceOp.SUM, group=group)
            dist.all_reduce(self.visibility_count, op=dist.ReduceOp.SUM, group=group)
            dist.all_reduce(self.max_radii, op=dist.ReduceOp.MAX, group=group)
        return self


def render_view_batch(render_fn, views: Iterable, params: Dict[str, torch.Tensor], loss_fn, global_batch: int,
                      group=None) -> Dict[str, object]:
    """One optimisation step's worth of rendering, data-parallel over views.

    render_fn(view) -> the dict render() returns; loss_fn(pkg, view) -> scalar loss of that view.
    `views` are the views THIS rank owns (see shard_views).  After the call every tensor in `params`
    holds, in .grad, the batch gradient (sum over all ranks' views of d(loss/global_batch)), exactly
    what the reference's sequential loop leaves there, and the returned stats are the merged
    densification statistics."""
    first = next(iter(params.values()))
    stats = ViewBatchStats(first.shape[0], first.device)
    total = torch.zeros((), device=first.device)
    for view in views:
        pkg = render_fn(view)
        loss = loss_fn(pkg, view) / global_batch
        loss.backward()
        total += loss.detach()
        stats.add_view(pkg["viewspace_points"].grad, pkg["radii"])
    grads = []
    for p in params.values():
        if p.grad is None:
            p.grad = torch.zeros_like(p)
        grads.append(p.grad)
    stats.reduce(group=group)
    # Gaussians rendered by no view of the step carry zero gradients everywhere: exchange the union's rows only
    allreduce_gradients(grads, group=group, union_visible=stats.max_radii > 0)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(total, op=dist.ReduceOp.SUM, group=group)
    return {"loss": total, "stats": stats}
