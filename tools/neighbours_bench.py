#!/usr/bin/env python3
"""Timings of the training-step neighbours of the rasterizer next to the PyTorch / reference
paths they replace, with achieved HBM bandwidth against the measured peak (run under gpurun; markdown on stdout).

  fused L1 + SSIM (csrc/loss.cu)      vs  the reference's formula on cuDNN grouped convolutions + autograd
  fused multi-tensor Adam (optim.cu)  vs  torch.optim.Adam (foreach), dense; and the sparse (rendered rows) mode
  grid kNN, k = 20 (knn.cu)           vs  the brute-force scan pointops2's knnquery performs (same kernel family)
"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("4d-gaussian-splatting_b200", "oracle", "tests"):
    sys.path.insert(0, os.path.join(ROOT, p))
import lego  # noqa: E402
from fdgs.loss import l1_ssim_loss  # noqa: E402
from fdgs.optim import FusedAdam  # noqa: E402
from fdgs.knn import knn  # noqa: E402

dev = "cuda:0"
try:
    PEAK = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    PEAK = 6650.0


def timeit(fn, iters=20, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


print("# Training-step neighbours of the rasterizer: fused kernels vs the paths they replace\n")
print("GPU: %s; HBM peak used for the fractions: %.0f GB/s (MEASURED_PEAKS.json).\n" % (torch.cuda.get_device_name(0), PEAK))

# ---- loss ---------------------------------------------------------------------------------------------------------
H, W = 1014, 1352
g = torch.Generator().manual_seed(0)
gt = torch.rand(3, H, W, generator=g).to(dev)
img = (gt + 0.1 * torch.randn(3, H, W, generator=g).to(dev)).clamp(0, 1).requires_grad_(True)


def fused():
    img.grad = None
    l1_ssim_loss(img, gt, 0.2).backward()


def torch_path():
    img.grad = None
    loss = 0.8 * (img - gt).abs().mean() + 0.2 * (1.0 - lego.ssim_torch(img[None], gt[None]))
    loss.backward()


tf, tt = timeit(fused), timeit(torch_path)
N = 3 * H * W
alg = 5 * N * 4            # roofline against 5 planes (x, y in; dL/dx out; + the forward's two reads again in backward)
moved = 11 * N * 4         # what the two kernels move: fwd 2 in + 3 out, bwd 5 in + 1 out
print("## Fused L1 + SSIM, forward + backward, 3x%dx%d\n" % (H, W))
print("| path | ms | algorithmic 5 planes: GB/s (frac of peak) | moved 11 planes: GB/s (frac) |")
print("|---|---|---|---|")
print("| fdgs.loss.l1_ssim_loss (2 kernels) | %.3f | %.0f (%.3f) | %.0f (%.3f) |" % (tf, alg / tf / 1e6, alg / tf / 1e6 / PEAK, moved / tf / 1e6, moved / tf / 1e6 / PEAK))
print("| reference formula: 5 grouped conv2d (cuDNN) + elementwise + autograd | %.3f | %.0f (%.3f) | - |" % (tt, alg / tt / 1e6, alg / tt / 1e6 / PEAK))
print("\nspeed-up %.1fx; the fused kernels are bound by their shared-memory filter passes (22 taps x 8 maps per pixel), not by HBM.\n" % (tt / tf))

# ---- Adam ---------------------------------------------------------------------------------------------------------
P = 2_000_000
shapes = [(P, 3), (P, 1), (P, 3), (P, 1), (P, 4), (P, 4), (P, 1), (P, 1, 3), (P, 47, 3)]
lrs = [1.6e-4, 1.6e-4, 5e-3, 5e-3, 1e-3, 1e-3, 5e-2, 2.5e-3, 2.5e-3 / 20]
pa = [torch.randn(*s, device=dev).requires_grad_(True) for s in shapes]
pb = [p.detach().clone().requires_grad_(True) for p in pa]
grads = [torch.randn(*s, device=dev) for s in shapes]
oa = torch.optim.Adam([{"params": [p], "lr": lr} for p, lr in zip(pa, lrs)], lr=0.0, eps=1e-15)
ob = FusedAdam([{"params": [p], "lr": lr} for p, lr in zip(pb, lrs)], eps=1e-15)
for p, q, gr in zip(pa, pb, grads):
    p.grad, q.grad = gr, gr.clone()
vis = torch.rand(P, device=dev) < 0.31
rows = torch.nonzero(vis).squeeze(1)
ta = timeit(lambda: oa.step(), 10, 3)
tb = timeit(lambda: ob.step(), 10, 3)
ts_ = timeit(lambda: ob.step(rows=rows), 10, 3)
elems = sum(int(torch.tensor(s[1:]).prod()) for s in shapes) * P
bytes_dense = 28 * elems
print("## Adam step over the 9 parameter groups of a 2M-Gaussian model (161 floats per Gaussian)\n")
print("| path | ms | GB/s at 28 B per element (frac of peak) |")
print("|---|---|---|")
print("| torch.optim.Adam (foreach), dense | %.3f | %.0f (%.2f) |" % (ta, bytes_dense / ta / 1e6, bytes_dense / ta / 1e6 / PEAK))
print("| fdgs.optim.FusedAdam, dense (one launch) | %.3f | %.0f (%.2f) |" % (tb, bytes_dense / tb / 1e6, bytes_dense / tb / 1e6 / PEAK))
print("| fdgs.optim.FusedAdam, sparse: %d rendered rows (%.0f %%) | %.3f | %.0f (%.2f) |" % (
    rows.numel(), 100.0 * rows.numel() / P, ts_, bytes_dense * rows.numel() / P / ts_ / 1e6, bytes_dense * rows.numel() / P / ts_ / 1e6 / PEAK))
del pa, pb, grads, oa, ob
torch.cuda.empty_cache()

# ---- kNN ----------------------------------------------------------------------------------------------------------
print("\n## k = 20 nearest neighbours of a cloud among itself (rigidity loss, train.py:132-152)\n")
print("| points | grid search ms | brute force ms | speed-up |")
print("|---|---|---|---|")
for n in (20_000, 100_000, 300_000):
    x = (torch.rand(n, 3, device=dev) * 2.6 - 1.3)[None].contiguous()
    tg = timeit(lambda: knn(x, x, 20), 5, 2)
    tb_ = timeit(lambda: knn(x, x, 20, brute_force=True), 2 if n > 100_000 else 3, 1)
    print("| %d | %.3f | %.1f | %.0fx |" % (n, tg, tb_, tb_ / tg))

# ---- view-parallel SH reconstruction (csrc/preprocess_bwd.cu: sh_outer_dir_kernel + sh_outer_sum_kernel) ---------------
import fdgs  # noqa: E402
import helpers  # noqa: E402
C = fdgs.ext()
cfg, cam, sc, st = helpers.build(dict(P=2_000_000, W=64, H=64, seed=5), device=dev)
print("\n## Rebuilding + summing the dL_dsh rows of V views from 3-float colour factors (2M Gaussians, 48 coefficients)\n")
print("| views V | union rows K | ms | bytes written + read | GB/s (frac of peak) |")
print("|---|---|---|---|---|")
for V, frac in ((1, 0.31), (2, 0.33), (8, 0.40)):
    union = torch.rand(sc.P, device=dev) < frac
    cs = torch.cumsum(union, 0, dtype=torch.int32)
    K = int(cs[-1])
    slot, idx = C.union_maps(union.to(torch.int32), cs, K)
    meta_off = (3 * K + 3) // 4 * 4
    stride = meta_off + 8
    table = torch.randn(V, stride, device=dev) * (torch.rand(V, stride, device=dev) < 0.8)
    for v in range(V):
        table[v, meta_off] = 0.3 + 0.05 * v
        table[v, meta_off + 1:meta_off + 4] = torch.tensor([0.1 * v, -0.2, 0.05 * v], device=dev)
    out = torch.empty(sc.P, 48, 3, device=dev)
    fn = lambda: C.sh_outer_sum(table, stride, meta_off, V, K, slot, idx, sc.means3D, sc.ts, sc.scales, sc.scales_t, sc.rotations,
                                sc.rotations_r, 1.0, 1.0, True, 4, False, 3, 2, [out], False)
    t = timeit(fn, 10, 3)
    nbytes = sc.P * 576 + V * K * (12 + 32 + 32) + sc.P * 4 + K * 8
    print("| %d | %d | %.3f | %.2f GB | %.0f (%.2f) |" % (V, K, t, nbytes / 1e9, nbytes / t / 1e6, nbytes / t / 1e6 / PEAK))
