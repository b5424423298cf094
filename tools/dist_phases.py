#!/usr/bin/env python3
"""Time the phases of the multi-GPU gradient exchange (torchrun, N ranks): dense all-reduce vs the
union-of-visibility exchange (nonzero / pack / all-reduce / unpack).  Synthetic tensors of the cfg3 shapes."""
import os
import sys
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "4d-gaussian-splatting_b200"))
from fdgs import dist as fdist  # noqa: E402

rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl")
dev = "cuda:%d" % local
P = 2_000_000
g = torch.Generator(device=dev).manual_seed(1 + rank)
vis = torch.rand(P, device=dev, generator=g) < 0.31
shapes = [(P, 3), (P, 1), (P, 48, 3), (P, 1), (P, 3), (P, 1), (P, 4), (P, 4), (P, 3)]
grads = [torch.randn(*s, device=dev) * vis.view(P, *([1] * (len(s) - 1))) for s in shapes]


def timed(fn, n=5):
    for _ in range(2):
        fn()
    dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


union = vis.to(torch.int32)
dist.all_reduce(union, op=dist.ReduceOp.MAX)
umask = union > 0
res = {}
res["dense 9 tensors"] = timed(lambda: fdist.allreduce_gradients(grads))
flat_all = torch.cat([x.view(-1) for x in grads])
res["dense 1 flat (%.2f GB)" % (flat_all.numel() * 4 / 1e9)] = timed(lambda: dist.all_reduce(flat_all))
res["sparse total (union %.2f)" % float(umask.float().mean())] = timed(lambda: fdist.allreduce_gradients(grads, union_visible=umask))
res["nonzero"] = timed(lambda: torch.nonzero(umask).squeeze(1))
idx = torch.nonzero(umask).squeeze(1)
K = idx.numel()
sh = grads[2].view(P, -1)
blk = torch.empty(K, sh.shape[1], device=dev)
res["index_select sh"] = timed(lambda: torch.index_select(sh, 0, idx, out=blk))
res["index_copy sh"] = timed(lambda: sh.index_copy_(0, idx, blk))
flat = torch.empty(K * 161, device=dev)
res["allreduce flat K*161 (%.2f GB)" % (flat.numel() * 4 / 1e9)] = timed(lambda: dist.all_reduce(flat))
if rank == 0:
    for k, v in res.items():
        print("%-40s %8.3f ms" % (k, v))
dist.destroy_process_group()
