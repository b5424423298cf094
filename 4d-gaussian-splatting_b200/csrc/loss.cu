// loss.cu -- fused photometric loss of the training step: (1 - lambda) * L1 + lambda * (1 - SSIM), forward and backward.
//
// Replaces, in the reference, the step right after the rasterizer:
//   l1_loss (utils/loss_utils.py:18-19), ssim / _ssim (:39-64: five grouped 11x11 Gaussian conv2d's over the image,
//   the ground truth and their products, ~10 elementwise kernels) and their autograd backward (five more
//   convolutions), combined as train.py:115-117.
// Here: two kernels.
//   ssim_fwd_kernel   one CTA per 32x32 pixel tile and channel: stages the tile + 5-pixel halo of both images in shared
//                     memory once, runs the separable 11-tap Gaussian (sigma 1.5, zero padding like conv2d's
//                     padding = 5) over x, y, x^2, y^2, xy, evaluates the SSIM map and the three partial derivatives
//                     dS/dmu_x, dS/dE[x^2], dS/dE[xy] per pixel, writes those three planes and accumulates the two
//                     loss sums (sum |x - y|, sum SSIM) with one double atomicAdd per CTA each.
//   ssim_bwd_kernel   dL/dx = -lambda/N * (G*a + 2 x (G*b) + y (G*c)) + (1 - lambda)/N * sign(x - y): the adjoint of the
//                     windowed means is the same symmetric separable filter applied to the derivative planes.
// HBM traffic: forward reads 2 and writes 3 planes per channel, backward reads 5 and writes 1 -- 11 plane passes
// (181 MB at 3 x 1352 x 1014) against ~60 for the reference's op-by-op path; the kernels are bound by the shared-memory
// filter passes, not by HBM.  dL/dx is the dL_dpix the blend backward consumes (fdgs_backward_args.dL_dpix).
#include "fdgs_internal.h"

namespace fdgs {
namespace {

constexpr int LT = 32;            // tile edge (pixels)
constexpr int LR = 5;             // filter radius: window_size 11 (reference: utils/loss_utils.py:39)
constexpr int LW = LT + 2 * LR;   // staged edge
constexpr int LTHREADS = 256;

struct GaussWin {
    float g[2 * LR + 1];
};

// x, y: [C,H,W]; maps: [3,C,H,W] (a = dS/dmu_x, b = dS/dE[xx], c = dS/dE[xy]); sums[0] += sum|x-y|, sums[1] += sum SSIM
__global__ void __launch_bounds__(LTHREADS) ssim_fwd_kernel(const float* __restrict__ x, const float* __restrict__ y, int H, int W,
                                                           const GaussWin win, float* __restrict__ maps, size_t plane_stride,
                                                           double* __restrict__ sums) {
    __shared__ float sx[LW][LW + 1], sy[LW][LW + 1];
    __shared__ float hz[5][LW][LT + 1];          // horizontal pass of x, y, xx, yy, xy
    __shared__ double red[2][LTHREADS / 32];
    const int ch = blockIdx.z;
    const int x0 = blockIdx.x * LT, y0 = blockIdx.y * LT;
    const size_t HW = (size_t)H * W;
    const float* px = x + ch * HW;
    const float* py = y + ch * HW;
    for (int i = threadIdx.x; i < LW * LW; i += LTHREADS) {
        const int r = i / LW, c = i - r * LW;
        const int gy = y0 + r - LR, gx = x0 + c - LR;
        const bool in = gy >= 0 && gy < H && gx >= 0 && gx < W;   // zero padding (F.conv2d padding = 5)
        sx[r][c] = in ? px[(size_t)gy * W + gx] : 0.f;
        sy[r][c] = in ? py[(size_t)gy * W + gx] : 0.f;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < LW * LT; i += LTHREADS) {
        const int r = i / LT, c = i - r * LT;
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f, a4 = 0.f;
#pragma unroll
        for (int k = 0; k <= 2 * LR; ++k) {
            const float u = sx[r][c + k], v = sy[r][c + k], w = win.g[k];
            a0 = fmaf(w, u, a0);
            a1 = fmaf(w, v, a1);
            a2 = fmaf(w, u * u, a2);
            a3 = fmaf(w, v * v, a3);
            a4 = fmaf(w, u * v, a4);
        }
        hz[0][r][c] = a0; hz[1][r][c] = a1; hz[2][r][c] = a2; hz[3][r][c] = a3; hz[4][r][c] = a4;
    }
    __syncthreads();
    double l1 = 0.0, ss = 0.0;
    for (int i = threadIdx.x; i < LT * LT; i += LTHREADS) {
        const int r = i / LT, c = i - r * LT;
        const int gy = y0 + r, gx = x0 + c;
        if (gy >= H || gx >= W) continue;
        float m1 = 0.f, m2 = 0.f, e11 = 0.f, e22 = 0.f, e12 = 0.f;
#pragma unroll
        for (int k = 0; k <= 2 * LR; ++k) {
            const float w = win.g[k];
            m1 = fmaf(w, hz[0][r + k][c], m1);
            m2 = fmaf(w, hz[1][r + k][c], m2);
            e11 = fmaf(w, hz[2][r + k][c], e11);
            e22 = fmaf(w, hz[3][r + k][c], e22);
            e12 = fmaf(w, hz[4][r + k][c], e12);
        }
        // reference: utils/loss_utils.py:50-60
        const float C1 = 0.01f * 0.01f, C2 = 0.03f * 0.03f;
        const float m11 = m1 * m1, m22 = m2 * m2, m12 = m1 * m2;
        const float s11 = e11 - m11, s22 = e22 - m22, s12 = e12 - m12;
        const float A1 = 2.f * m12 + C1, A2 = 2.f * s12 + C2, B1 = m11 + m22 + C1, B2 = s11 + s22 + C2;
        const float inv = 1.f / (B1 * B2);
        const float S = A1 * A2 * inv;
        // S as a function of (mu_x, E[xx], E[xy]) with the ground-truth moments fixed
        const float da = 2.f * m2 * (A2 - A1) * inv + 2.f * m1 * S * (1.f / B2 - 1.f / B1);
        const float db = -S / B2;
        const float dc = 2.f * A1 * inv;
        const size_t o = ch * HW + (size_t)gy * W + gx;
        maps[o] = da;
        maps[plane_stride + o] = db;
        maps[2 * plane_stride + o] = dc;
        l1 += (double)fabsf(sx[r + LR][c + LR] - sy[r + LR][c + LR]);
        ss += (double)S;
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        l1 += __shfl_xor_sync(0xffffffffu, l1, d);
        ss += __shfl_xor_sync(0xffffffffu, ss, d);
    }
    if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = l1; red[1][threadIdx.x >> 5] = ss; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double t0 = 0.0, t1 = 0.0;
        for (int w = 0; w < LTHREADS / 32; ++w) { t0 += red[0][w]; t1 += red[1][w]; }
        atomicAdd(&sums[0], t0);
        atomicAdd(&sums[1], t1);
    }
}

// dL_dx[C,H,W] = k_ssim * (G*a + 2 x (G*b) + y (G*c)) + k_l1 * sign(x - y)
__global__ void __launch_bounds__(LTHREADS) ssim_bwd_kernel(const float* __restrict__ x, const float* __restrict__ y, int H, int W,
                                                           const GaussWin win, const float* __restrict__ maps, size_t plane_stride,
                                                           const float* __restrict__ grad_scale, float k_l1, float k_ssim,
                                                           float* __restrict__ dL_dx) {
    __shared__ float sm[3][LW][LW + 1];
    __shared__ float hz[3][LW][LT + 1];
    const int ch = blockIdx.z;
    const int x0 = blockIdx.x * LT, y0 = blockIdx.y * LT;
    const size_t HW = (size_t)H * W;
    for (int i = threadIdx.x; i < LW * LW; i += LTHREADS) {
        const int r = i / LW, c = i - r * LW;
        const int gy = y0 + r - LR, gx = x0 + c - LR;
        const bool in = gy >= 0 && gy < H && gx >= 0 && gx < W;
        const size_t o = ch * HW + (size_t)(in ? gy : 0) * W + (in ? gx : 0);
#pragma unroll
        for (int m = 0; m < 3; ++m) sm[m][r][c] = in ? maps[m * plane_stride + o] : 0.f;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < LW * LT; i += LTHREADS) {
        const int r = i / LT, c = i - r * LT;
        float a0 = 0.f, a1 = 0.f, a2 = 0.f;
#pragma unroll
        for (int k = 0; k <= 2 * LR; ++k) {
            const float w = win.g[k];
            a0 = fmaf(w, sm[0][r][c + k], a0);
            a1 = fmaf(w, sm[1][r][c + k], a1);
            a2 = fmaf(w, sm[2][r][c + k], a2);
        }
        hz[0][r][c] = a0; hz[1][r][c] = a1; hz[2][r][c] = a2;
    }
    __syncthreads();
    const float gs = grad_scale ? grad_scale[0] : 1.f;   // upstream d(total)/d(loss), a device scalar (no host sync)
    for (int i = threadIdx.x; i < LT * LT; i += LTHREADS) {
        const int r = i / LT, c = i - r * LT;
        const int gy = y0 + r, gx = x0 + c;
        if (gy >= H || gx >= W) continue;
        float ca = 0.f, cb = 0.f, cc = 0.f;
#pragma unroll
        for (int k = 0; k <= 2 * LR; ++k) {
            const float w = win.g[k];
            ca = fmaf(w, hz[0][r + k][c], ca);
            cb = fmaf(w, hz[1][r + k][c], cb);
            cc = fmaf(w, hz[2][r + k][c], cc);
        }
        const size_t o = ch * HW + (size_t)gy * W + gx;
        const float xv = x[o], yv = y[o];
        const float d = xv - yv;
        const float sgn = (d > 0.f) ? 1.f : ((d < 0.f) ? -1.f : 0.f);   // torch.abs backward: sign, 0 at 0
        dL_dx[o] = gs * (k_ssim * (ca + 2.f * xv * cb + yv * cc) + k_l1 * sgn);
    }
}

GaussWin make_window() {
    // reference: utils/loss_utils.py:24-26 gaussian(11, 1.5): float32 exp values normalised by their float32 sum
    GaussWin w;
    float v[2 * LR + 1];
    float s = 0.f;
    for (int i = 0; i <= 2 * LR; ++i) {
        v[i] = (float)exp(-(double)((i - LR) * (i - LR)) / (2.0 * 1.5 * 1.5));
        s += v[i];
    }
    for (int i = 0; i <= 2 * LR; ++i) w.g[i] = v[i] / s;
    return w;
}

}  // namespace

cudaError_t launch_l1_ssim_fwd(const float* x, const float* y, int C, int H, int W, float* maps, double* sums, cudaStream_t stream) {
    if (C <= 0 || H <= 0 || W <= 0) return cudaSuccess;
    cudaError_t e = cudaMemsetAsync(sums, 0, 2 * sizeof(double), stream);
    if (e != cudaSuccess) return e;
    dim3 grid((W + LT - 1) / LT, (H + LT - 1) / LT, C);
    ssim_fwd_kernel<<<grid, LTHREADS, 0, stream>>>(x, y, H, W, make_window(), maps, (size_t)C * H * W, sums);
    return cudaGetLastError();
}

cudaError_t launch_l1_ssim_bwd(const float* x, const float* y, int C, int H, int W, const float* maps, const float* grad_scale,
                               float lambda_dssim, float* dL_dx, cudaStream_t stream) {
    if (C <= 0 || H <= 0 || W <= 0) return cudaSuccess;
    const float n = (float)((double)C * H * W);
    dim3 grid((W + LT - 1) / LT, (H + LT - 1) / LT, C);
    ssim_bwd_kernel<<<grid, LTHREADS, 0, stream>>>(x, y, H, W, make_window(), maps, (size_t)C * H * W, grad_scale,
                                                   (1.f - lambda_dssim) / n, -lambda_dssim / n, dL_dx);
    return cudaGetLastError();
}

}  // namespace fdgs
