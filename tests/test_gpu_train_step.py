"""GPU (-m gpu): the training-step neighbours of the rasterizer:
fused L1 + SSIM against the reference's formula in plain PyTorch fp32 (utils/loss_utils.py:18-64), fused Adam against
torch.optim.Adam(eps=1e-15) (scene/gaussian_model.py:357), grid kNN against the brute-force scan of
pointops2/src/knnquery/knnquery_cuda_kernel.cu:65-107."""
import math

import numpy as np
import pytest
import torch
import torch.nn.functional as F

import helpers

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


# ---- reference formulas, restated (utils/loss_utils.py:24-64) -------------------------------------------------------
def _window(channel, device):
    g = torch.tensor([math.exp(-(x - 5) ** 2 / float(2 * 1.5 ** 2)) for x in range(11)])
    g = (g / g.sum()).unsqueeze(1)
    w2 = g.mm(g.t()).float().unsqueeze(0).unsqueeze(0)
    return w2.expand(channel, 1, 11, 11).contiguous().to(device)


def _ssim_ref(img1, img2):
    ch = img1.size(-3)
    w = _window(ch, img1.device)
    mu1 = F.conv2d(img1, w, padding=5, groups=ch)
    mu2 = F.conv2d(img2, w, padding=5, groups=ch)
    mu1_sq, mu2_sq, mu1_mu2 = mu1.pow(2), mu2.pow(2), mu1 * mu2
    s1 = F.conv2d(img1 * img1, w, padding=5, groups=ch) - mu1_sq
    s2 = F.conv2d(img2 * img2, w, padding=5, groups=ch) - mu2_sq
    s12 = F.conv2d(img1 * img2, w, padding=5, groups=ch) - mu1_mu2
    C1, C2 = 0.01 ** 2, 0.03 ** 2
    return (((2 * mu1_mu2 + C1) * (2 * s12 + C2)) / ((mu1_sq + mu2_sq + C1) * (s1 + s2 + C2))).mean()


@pytest.mark.parametrize("shape", [(3, 67, 101), (3, 256, 256), (1, 40, 33), (3, 1014, 1352)])
@pytest.mark.parametrize("lam", [0.2, 0.0, 1.0])
def test_fused_l1_ssim_matches_the_pytorch_formula(shape, lam):
    from fdgs.loss import l1_ssim_loss
    g = torch.Generator().manual_seed(shape[1] * 7 + shape[2])
    gt = torch.rand(*shape, generator=g)
    img = (gt + 0.15 * torch.randn(*shape, generator=g)).clamp(0, 1.2)
    img[:, : shape[1] // 3] = gt[:, : shape[1] // 3]                  # a region with x == y exactly (sign(0) = 0)
    gt, img = gt.to(DEV), img.to(DEV)
    a = img.clone().requires_grad_(True)
    b = img.clone().requires_grad_(True)
    # reference: train.py:115-117 with torch.backends' fp32 convolution (no TF32)
    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        ref = (1.0 - lam) * (a - gt).abs().mean() + lam * (1.0 - _ssim_ref(a[None], gt[None]))
        (3.0 * ref).backward()
    finally:
        torch.backends.cudnn.allow_tf32 = old
    loss, l1, ssim = l1_ssim_loss(b, gt, lam, return_terms=True)
    (3.0 * loss).backward()
    assert abs(float(loss) - float(ref)) <= 1e-5 * max(1.0, abs(float(ref)))
    assert abs(float(l1) - float((img - gt).abs().mean())) < 1e-6
    n = a.grad.abs().max().item()
    assert (b.grad - a.grad).abs().max().item() <= 1e-5 * n + 1e-9, ((b.grad - a.grad).abs().max().item(), n)
    assert helpers.l2_rel(helpers.to_np(b.grad), helpers.to_np(a.grad)) < 1e-5


def test_loss_module_mirrors_reference_names():
    from fdgs import loss as L
    g = torch.Generator().manual_seed(1)
    x, y = torch.rand(3, 64, 80, generator=g).to(DEV), torch.rand(3, 64, 80, generator=g).to(DEV)
    assert abs(float(L.l1_loss(x, y)) - float((x - y).abs().mean())) < 1e-6
    assert abs(float(L.ssim(x, y)) - float(_ssim_ref(x[None], y[None]))) < 1e-5
    assert abs(float(L.ssim(x, x)) - 1.0) < 1e-6


# ---- Adam ---------------------------------------------------------------------------------------------------------------
def _adam_setup(P, seed):
    g = torch.Generator().manual_seed(seed)
    shapes = [(P, 3), (P, 1), (P, 48, 3), (P, 4), (P, 1, 3)]
    lrs = [1.6e-4, 5e-2, 2.5e-3 / 20, 1e-3, 2.5e-3]
    params = [torch.randn(*s, generator=g).to(DEV) for s in shapes]
    return params, lrs, g


def test_fused_adam_dense_matches_torch_adam():
    """K steps of the fused kernel vs torch.optim.Adam (eps 1e-15, the reference's setting): same update op for op."""
    from fdgs.optim import FusedAdam
    P, K = 5003, 6
    params, lrs, g = _adam_setup(P, 3)
    pa = [p.clone().requires_grad_(True) for p in params]
    pb = [p.clone().requires_grad_(True) for p in params]
    oa = torch.optim.Adam([{"params": [p], "lr": lr} for p, lr in zip(pa, lrs)], lr=0.0, eps=1e-15)
    ob = FusedAdam([{"params": [p], "lr": lr} for p, lr in zip(pb, lrs)], lr=0.0, eps=1e-15)
    for k in range(K):
        for a, b in zip(pa, pb):
            gr = (torch.randn(a.shape, generator=g) * (torch.rand(a.shape[0], generator=g) < 0.6).view(-1, *([1] * (a.dim() - 1)))).to(DEV)
            a.grad, b.grad = gr.clone(), gr.clone()
        oa.step()
        ob.step()
    exact = 0
    for a, b, lr in zip(pa, pb, lrs):
        # every step moves a parameter by <= lr; the two implementations may differ by an ulp of the moments per step
        d = (a - b).abs()
        i = int(d.argmax())
        assert float(d.max()) <= 1e-5 * lr * K + 2e-7 * float(a.abs().max()), (
            "lr", lr, "max diff", float(d.max()), "at", i, float(a.flatten()[i]), float(b.flatten()[i]))
        exact += int(torch.equal(a, b))
    for a, b in zip(pa, pb):   # optimizer state too
        assert torch.allclose(oa.state[a]["exp_avg"], ob.state[b]["exp_avg"], rtol=2e-6, atol=2e-7)      # |g| ~ 1: an ulp of g
        assert torch.allclose(oa.state[a]["exp_avg_sq"], ob.state[b]["exp_avg_sq"], rtol=2e-6, atol=1e-9)
    print("fused Adam: %d of %d tensors bit-identical to torch.optim.Adam after %d steps" % (exact, len(pa), K))


def test_fused_adam_sparse_rows_and_zero_grad():
    from fdgs.optim import FusedAdam
    P = 4001
    params, lrs, g = _adam_setup(P, 5)
    pd = [p.clone().requires_grad_(True) for p in params]
    ps = [p.clone().requires_grad_(True) for p in params]
    od = FusedAdam([{"params": [p], "lr": lr} for p, lr in zip(pd, lrs)], eps=1e-15)
    osp = FusedAdam([{"params": [p], "lr": lr} for p, lr in zip(ps, lrs)], eps=1e-15)
    vis = torch.rand(P, generator=g) < 0.3
    rows = torch.nonzero(vis).squeeze(1).to(DEV)
    for a, b in zip(pd, ps):
        gr = (torch.randn(a.shape, generator=g) * vis.view(-1, *([1] * (a.dim() - 1)))).to(DEV)
        a.grad, b.grad = gr.clone(), gr.clone()
    od.step()
    osp.step(rows=rows, zero_grad=True)
    v = vis.to(DEV)
    for a, b, p0 in zip(pd, ps, params):
        assert torch.equal(a[v], b[v])                  # listed rows: the dense update, bit for bit (first step)
        assert torch.equal(b[~v], p0[~v])               # other rows untouched
        assert float(b.grad.abs().sum()) == 0.0         # consumed gradients cleared


# ---- kNN ----------------------------------------------------------------------------------------------------------------
def _knn_numpy(x, k):
    """reference: knnquery_cuda_kernel.cu:65-107 (squared distances, strict < keeps the lower index among equals)"""
    d2 = ((x[:, None, :] - x[None, :, :]) ** 2)
    d2 = (d2[..., 0] + d2[..., 1]) + d2[..., 2]
    idx = np.argsort(d2, axis=1, kind="stable")[:, :k]
    return idx, np.take_along_axis(d2, idx, 1)


@pytest.mark.parametrize("n,kind", [(2000, "uniform"), (20000, "uniform"), (20000, "clustered"), (100000, "lego"), (12, "tiny")])
def test_grid_knn_equals_brute_force(n, kind):
    from fdgs.knn import knn
    g = torch.Generator().manual_seed(n + len(kind))
    if kind == "clustered":       # dense blobs + sparse background: empty cells, long shell walks
        c = torch.randn(8, 3, generator=g) * 3
        x = torch.cat([c[torch.randint(0, 8, (n - n // 10,), generator=g)] + 0.05 * torch.randn(n - n // 10, 3, generator=g),
                       (torch.rand(n // 10, 3, generator=g) - 0.5) * 40])
    elif kind == "lego":          # the cfg4 initialisation: U(-1.3, 1.3)^3 (scene/dataset_readers.py:329)
        x = torch.rand(n, 3, generator=g) * 2.6 - 1.3
    else:
        x = torch.rand(n, 3, generator=g) * torch.tensor([4.0, 1.0, 0.25])
    x = x.to(DEV)
    k = 20
    idx, d2 = knn(x[None], x[None], k)
    bidx, bd2 = knn(x[None], x[None], k, brute_force=True)
    assert idx.shape == (1, n, k) and idx.dtype == torch.int64
    if n >= k:
        assert torch.equal(d2, bd2)
        assert torch.equal(idx, bidx)
        assert torch.equal(idx[0, :, 0], torch.arange(n, device=DEV))       # the point itself first
        assert bool((d2[0, :, 1:] >= d2[0, :, :-1]).all())
    else:
        assert torch.equal(idx[0, :, :n], bidx[0, :, :n]) and torch.equal(d2[0, :, :n], bd2[0, :, :n])
        assert bool((d2[0, :, n:] == 1e10).all())
    if n <= 2000:
        ni, nd = _knn_numpy(x.cpu().numpy().astype(np.float32), min(k, n))
        assert (idx[0, :, :min(k, n)].cpu().numpy() == ni).all()


def test_rigid_loss_with_grid_knn_matches_brute_force():
    """train.py:132-152: weight = exp(-100 dist), Lrigid = sum(weight * |v_i - v_j|) / k / n -- same value with either search"""
    from fdgs.knn import knn
    g = torch.Generator().manual_seed(8)
    n, k = 30000, 20
    xyz = (torch.rand(n, 3, generator=g) * 2.6 - 1.3).to(DEV)
    vel = (0.1 * torch.randn(n, 3, generator=g)).to(DEV).requires_grad_(True)
    out = []
    for bf in (False, True):
        idx, dist = knn(xyz[None].contiguous(), xyz[None].contiguous(), k, brute_force=bf)
        weight = torch.exp(-100 * dist)
        vd = torch.norm(vel[idx] - vel[None, :, None], p=2, dim=-1)
        out.append((weight * vd).sum() / k / n)
    assert float(out[0]) == float(out[1]) and float(out[0]) > 0


def test_dist_cuda2_equals_exact_three_nearest():
    """simple-knn's distCUDA2 (simple-knn/simple_knn.cu:139-178): mean squared distance to the 3 nearest other points"""
    from fdgs.knn import distCUDA2
    g = torch.Generator().manual_seed(21)
    x = torch.rand(3000, 3, generator=g) * 2.6 - 1.3
    x[10] = x[4]                                             # an exact duplicate: distance 0 counts, the point itself does not
    xd = x.to(DEV)
    got = distCUDA2(xd).cpu().numpy()
    xn = x.numpy().astype(np.float32)
    d = xn[:, None, :] - xn[None, :, :]
    d2 = (d[..., 0] * d[..., 0] + d[..., 1] * d[..., 1]) + d[..., 2] * d[..., 2]
    np.fill_diagonal(d2, np.inf)
    best = np.sort(d2, axis=1)[:, :3]
    want = (best[:, 0] + best[:, 1] + best[:, 2]) / np.float32(3.0)
    assert np.allclose(got, want, rtol=1e-6, atol=0)
    assert got[10] <= got.mean() and np.isfinite(got).all()
