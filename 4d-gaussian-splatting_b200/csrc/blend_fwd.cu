// blend_fwd.cu -- per-tile front-to-back alpha blending (forward).
//
// Replaces the reference's renderCUDA<3> forward kernel (forward.cu:501-626).  Same tile
// decomposition (16x16 pixels per CTA, one thread per pixel), same per-pixel arithmetic in the
// same order (results are bit-identical to the reference kernel as compiled by nvcc), but a
// different execution structure:
//
//   * the tile's work list is a contiguous array of 64-byte instance records (binning.cu), so it
//     is streamed into shared memory with cp.async.bulk (TMA 1-D) through a 3-stage
//     full/empty mbarrier ring -- no __syncthreads in the loop, no index chasing, and the colour
//     / depth / flow payload arrives with the geometry instead of being re-read from global
//     memory per contributing pixel (reference: forward.cu:600-604);
//   * each warp owns an 8x4 pixel footprint and first culls every batch of 32 instances against
//     that footprint (one instance per lane, one ballot), then walks only the surviving
//     instances.  The cull box is the exact alpha >= 1/255 extent of the Gaussian, so skipped
//     (pixel, Gaussian) pairs are pairs the reference would have rejected anyway;
//   * pairs whose power is below log(1/(255*opacity)) skip expf() (same argument).
#include "fdgs_internal.h"

namespace fdgs {
namespace {

constexpr int BF_THREADS = 256;
constexpr int BF_WARPS = BF_THREADS / 32;
constexpr int BF_BATCH = 128;
constexpr int BF_STAGES = 3;

struct __align__(128) BlendFwdSmem {
    StageRec recs[BF_STAGES][BF_BATCH];
    uint64_t full[BF_STAGES];
    uint64_t empty[BF_STAGES];
    int done_warps;
};

__global__ void __launch_bounds__(BF_THREADS) blend_fwd_kernel(const BlendFwdParams p) {
    __shared__ BlendFwdSmem sm;
    const int tile = blockIdx.y * p.grid_x + blockIdx.x;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // 8x4 footprint per warp: warps tile the 16x16 block as 2 (x) by 4 (y)
    const int wx0 = blockIdx.x * TILE_X + (warp & 1) * 8;
    const int wy0 = blockIdx.y * TILE_Y + (warp >> 1) * 4;
    const int pix_x = wx0 + (lane & 7), pix_y = wy0 + (lane >> 3);
    const bool inside = pix_x < p.W && pix_y < p.H;
    const float pxf = (float)pix_x, pyf = (float)pix_y;
    const float bx0 = (float)wx0, bx1 = (float)(wx0 + 7), by0 = (float)wy0, by1 = (float)(wy0 + 3);

    const uint2 range = p.ranges[tile];
    const int n = (int)(range.y - range.x);
    const int nb = (n + BF_BATCH - 1) / BF_BATCH;
    const StageRec* src = p.recs + range.x;

    if (threadIdx.x == 0) {
        for (int s = 0; s < BF_STAGES; ++s) {
            mbar_init(&sm.full[s], 1);
            mbar_init(&sm.empty[s], BF_WARPS);
        }
        sm.done_warps = 0;
        mbar_fence_init();
    }
    __syncthreads();

    int issued = 0;
    if (threadIdx.x == 0) {
        for (; issued < nb && issued < BF_STAGES; ++issued) {
            const int cnt = min(BF_BATCH, n - issued * BF_BATCH);
            mbar_expect_tx(&sm.full[issued], (uint32_t)cnt * kStageRecBytes);
            bulk_g2s(&sm.recs[issued][0], src + (size_t)issued * BF_BATCH, (uint32_t)cnt * kStageRecBytes, &sm.full[issued]);
        }
    }

    float T = 1.0f;
    float C0 = 0.f, C1 = 0.f, C2 = 0.f, F0 = 0.f, F1 = 0.f, Dp = 0.f;
    uint32_t last_contributor = 0;
    bool done = !inside;
    bool warp_done = false;
    volatile int* done_warps = &sm.done_warps;

    int b = 0;
    for (; b < nb; ++b) {
        const int s = b % BF_STAGES;
        const uint32_t ph = (uint32_t)(b / BF_STAGES) & 1u;
        // Wait for batch b -- or for the whole tile to finish.  Once every warp is done the producer
        // stops refilling (below), so a warp that ran ahead must not wait for a batch that will never
        // come; lane 0 polls both conditions and the warp leaves together.
        bool stop = false;
        if (lane == 0) {
            while (!mbar_try_wait(&sm.full[s], ph)) {
                if (*done_warps == BF_WARPS) { stop = true; break; }
            }
            if (!stop && *done_warps == BF_WARPS) stop = true;
        }
        if (__any_sync(0xffffffffu, stop)) break;
        mbar_wait(&sm.full[s], ph);   // already complete: gives every lane the acquire on the batch
        if (!warp_done) {
            const int cnt = min(BF_BATCH, n - b * BF_BATCH);
            const StageRec* st = sm.recs[s];
            for (int r0 = 0; r0 < cnt; r0 += 32) {
                const int j = r0 + lane;
                bool rel = false;
                if (j < cnt) {
                    const float4 a = st[j].q0;
                    const float4 c = st[j].q1;
                    const float4 e = st[j].q3;
                    rel = rect_may_contribute(a.x, a.y, c.x, c.y, c.z, a.z, e.z, e.w, bx0, bx1, by0, by1);
                }
                uint32_t m = __ballot_sync(0xffffffffu, rel);
                while (m) {
                    const int k = __ffs(m) - 1;
                    m &= m - 1;
                    const StageRec* g = st + (r0 + k);
                    const float4 q0 = g->q0;   // x, y, pmin, -
                    const float4 q1 = g->q1;   // A, B, C, opacity
                    // reference: forward.cu:578-581 (contraction as compiled)
                    const float dx = fsub(q0.x, pxf);
                    const float dy = fsub(q0.y, pyf);
                    const float power =
                        ffma(ffma(dx, fmul(dx, q1.x), fmul(dy, fmul(dy, q1.z))), -0.5f, -fmul(dy, fmul(dx, q1.y)));
                    if (done || power > 0.0f || power < q0.z) continue;
                    const float alpha = fminf(fmul(q1.w, expf(power)), 0.99f);
                    if (alpha < 1.0f / 255.0f) continue;
                    const float test_T = fmul(T, fsub(1.0f, alpha));
                    if (test_T < 0.0001f) {
                        done = true;
                        continue;
                    }
                    const float4 q2 = g->q2;   // r, g, b, depth
                    const float4 q3 = g->q3;   // flow x, flow y, ex, ey
                    C0 = ffma(T, fmul(alpha, q2.x), C0);
                    C1 = ffma(T, fmul(alpha, q2.y), C1);
                    C2 = ffma(T, fmul(alpha, q2.z), C2);
                    F0 = ffma(T, fmul(alpha, q3.x), F0);
                    F1 = ffma(T, fmul(alpha, q3.y), F1);
                    Dp = ffma(T, fmul(alpha, q2.w), Dp);
                    T = test_T;
                    last_contributor = (uint32_t)(b * BF_BATCH + r0 + k + 1);
                }
                if (__all_sync(0xffffffffu, done)) {
                    warp_done = true;
                    if (lane == 0) atomicAdd(&sm.done_warps, 1);
                    break;
                }
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&sm.empty[s]);
        if (threadIdx.x == 0 && issued < nb) {
            // refill this stage with batch `issued` once every warp has released it
            bool go = true;
            while (!mbar_try_wait(&sm.empty[s], ph)) {
                if (*done_warps == BF_WARPS) { go = false; break; }
            }
            if (go) {
                const int cnt = min(BF_BATCH, n - issued * BF_BATCH);
                mbar_expect_tx(&sm.full[s], (uint32_t)cnt * kStageRecBytes);
                bulk_g2s(&sm.recs[s][0], src + (size_t)issued * BF_BATCH, (uint32_t)cnt * kStageRecBytes, &sm.full[s]);
                ++issued;
            }
        }
    }
    // drain bulk copies that were issued but never consumed (early exit) before the CTA retires
    if (threadIdx.x == 0) {
        for (int j = b; j < issued; ++j) mbar_wait(&sm.full[j % BF_STAGES], (uint32_t)(j / BF_STAGES) & 1u);
    }

    if (inside) {
        const int pix_id = pix_y * p.W + pix_x;
        const int HW = p.H * p.W;
        p.final_T[pix_id] = T;
        p.out_T[pix_id] = T;
        p.n_contrib[pix_id] = last_contributor;
        p.out_color[0 * HW + pix_id] = ffma(T, p.background[0], C0);
        p.out_color[1 * HW + pix_id] = ffma(T, p.background[1], C1);
        p.out_color[2 * HW + pix_id] = ffma(T, p.background[2], C2);
        p.out_flow[0 * HW + pix_id] = F0;
        p.out_flow[1 * HW + pix_id] = F1;
        p.out_depth[pix_id] = Dp;
    }
}

}  // namespace

cudaError_t launch_blend_fwd(const BlendFwdParams& p, cudaStream_t stream) {
    dim3 grid(p.grid_x, p.grid_y, 1);
    blend_fwd_kernel<<<grid, BF_THREADS, 0, stream>>>(p);
    return cudaGetLastError();
}

}  // namespace fdgs
