"""Fused photometric loss of the training step (csrc/loss.cu over the C-ABI).

Mirrors the reference's `utils/loss_utils.py` (`l1_loss` :18-19, `ssim` :39-64) and the way train.py:115-117 combines
them, `loss = (1 - lambda_dssim) * Ll1 + lambda_dssim * (1 - ssim)`, in ONE autograd node: two kernels instead of
five grouped convolutions + ~10 elementwise kernels forward and as many backward.  CUDA only; there is no fallback.
"""
import torch

import fdgs


class _L1Ssim(torch.autograd.Function):
    @staticmethod
    def forward(ctx, image, gt, lambda_dssim):
        C = fdgs.ext()
        sums, maps = C.l1_ssim_forward(image, gt)
        n = float(image.numel())
        ctx.save_for_backward(image, gt, maps)
        ctx.lambda_dssim = float(lambda_dssim)
        l1 = sums[0] / n
        ssim = sums[1] / n
        loss = (1.0 - lambda_dssim) * l1 + lambda_dssim * (1.0 - ssim)
        l1, ssim = l1.to(torch.float32), ssim.to(torch.float32)
        ctx.mark_non_differentiable(l1, ssim)
        return loss.to(torch.float32), l1, ssim

    @staticmethod
    def backward(ctx, g_loss, g_l1, g_ssim):
        image, gt, maps = ctx.saved_tensors
        if g_loss is None:
            return None, None, None
        dx = fdgs.ext().l1_ssim_backward(image, gt, maps, g_loss.reshape(1), ctx.lambda_dssim)
        return dx, None, None


def l1_ssim_loss(image, gt, lambda_dssim=0.2, return_terms=False):
    """(1 - lambda) * mean|image - gt| + lambda * (1 - SSIM(image, gt)); images [3,H,W] (or [C,H,W]) float32 on the GPU.
    Gradient w.r.t. `image` only (the ground truth is data)."""
    loss, l1, ssim = _L1Ssim.apply(image, gt.detach(), float(lambda_dssim))
    return (loss, l1, ssim) if return_terms else loss


def l1_loss(network_output, gt):
    """reference: utils/loss_utils.py:18-19"""
    return l1_ssim_loss(network_output, gt, 0.0)


def ssim(img1, img2, window_size=11, size_average=True):
    """reference: utils/loss_utils.py:39-47 (window 11, sigma 1.5, mean over the whole map).  1 - l1_ssim_loss(.., 1)."""
    if window_size != 11 or not size_average:
        raise NotImplementedError("fdgs.loss.ssim: only window_size=11, size_average=True (what train.py uses)")
    return 1.0 - l1_ssim_loss(img1, img2, 1.0)
