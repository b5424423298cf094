// blend_bwd.cu -- per-tile back-to-front alpha blending (backward).
//
// Replaces the reference's renderCUDA<3> backward kernel (backward.cu:926-1137): same per-pixel
// back-to-front recurrences, same threshold decisions as the forward (power > 0, alpha < 1/255,
// the per-pixel n_contrib cut), but
//
//   * the 12 fp32 atomicAdd per contributing (pixel, Gaussian) PAIR of the reference
//     (backward.cu:1076,1091,1124-1134) become one warp reduce-scatter (18 shuffles for 12
//     values) per (warp, Gaussian) followed by a single RED instruction with 12 active lanes --
//     a 32x..(32/cull) reduction in atomic traffic, and none at all for Gaussians whose
//     alpha >= 1/255 box misses the warp's 8x4 pixel footprint (culled before any math);
//   * the work list is streamed with cp.async.bulk through the same mbarrier ring as the
//     forward, walking the tile's contiguous record range from the back, starting at the deepest
//     position any pixel of the tile actually reached (max n_contrib) instead of the list end.
#include "fdgs_internal.h"

namespace fdgs {
namespace {

constexpr int BB_THREADS = 256;
constexpr int BB_WARPS = BB_THREADS / 32;
constexpr int BB_BATCH = 128;
constexpr int BB_STAGES = 3;

struct __align__(128) BlendBwdSmem {
    InstRec recs[BB_STAGES][BB_BATCH];
    uint64_t full[BB_STAGES];
    uint64_t empty[BB_STAGES];
    unsigned int nmax;
};

__global__ void __launch_bounds__(BB_THREADS, 3) blend_bwd_kernel(const BlendBwdParams p) {
    __shared__ BlendBwdSmem sm;
    const int tile = blockIdx.y * p.grid_x + blockIdx.x;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int wx0 = blockIdx.x * TILE_X + (warp & 1) * 8;
    const int wy0 = blockIdx.y * TILE_Y + (warp >> 1) * 4;
    const int pix_x = wx0 + (lane & 7), pix_y = wy0 + (lane >> 3);
    const bool inside = pix_x < p.W && pix_y < p.H;
    const float pxf = (float)pix_x, pyf = (float)pix_y;
    const float bx0 = (float)wx0, bx1 = (float)(wx0 + 7), by0 = (float)wy0, by1 = (float)(wy0 + 3);
    const int HW = p.H * p.W;
    const int pix_id = pix_y * p.W + pix_x;

    const uint2 range = p.ranges[tile];
    const int n_list = (int)(range.y - range.x);
    const unsigned int my_last = inside ? p.n_contrib[pix_id] : 0u;   // backward.cu:982
    const unsigned int wmax = __reduce_max_sync(0xffffffffu, my_last);

    if (threadIdx.x == 0) {
        for (int s = 0; s < BB_STAGES; ++s) {
            mbar_init(&sm.full[s], 1);
            mbar_init(&sm.empty[s], BB_WARPS);
        }
        sm.nmax = 0;
        mbar_fence_init();
    }
    __syncthreads();
    if (lane == 0 && wmax) atomicMax(&sm.nmax, wmax);
    __syncthreads();
    // entries at list positions >= max n_contrib are skipped by every pixel (backward.cu:1040)
    const int n = min(n_list, (int)sm.nmax);
    const int nb = (n + BB_BATCH - 1) / BB_BATCH;
    const InstRec* src = p.recs + range.x;

    int issued = 0;
    if (threadIdx.x == 0) {
        for (; issued < nb && issued < BB_STAGES; ++issued) {
            const int hi = n - issued * BB_BATCH, lo = max(0, hi - BB_BATCH);
            mbar_expect_tx(&sm.full[issued], (uint32_t)(hi - lo) * 64u);
            bulk_g2s(&sm.recs[issued][0], src + lo, (uint32_t)(hi - lo) * 64u, &sm.full[issued]);
        }
    }

    // per-pixel state, backward.cu:976-1011
    const float T_final = inside ? p.final_T[pix_id] : 0.f;
    float T = T_final;
    float acc_c0 = 0.f, acc_c1 = 0.f, acc_c2 = 0.f, acc_f0 = 0.f, acc_f1 = 0.f, acc_d = 0.f, acc_m = 0.f;
    float gp0 = 0.f, gp1 = 0.f, gp2 = 0.f, gf0 = 0.f, gf1 = 0.f, gd = 0.f, gm = 0.f;
    if (inside) {
        gp0 = p.dL_dpix[0 * HW + pix_id];
        gp1 = p.dL_dpix[1 * HW + pix_id];
        gp2 = p.dL_dpix[2 * HW + pix_id];
        gf0 = p.dL_dpix_flow[0 * HW + pix_id];
        gf1 = p.dL_dpix_flow[1 * HW + pix_id];
        gd = p.dL_depths[pix_id];
        gm = p.dL_masks[pix_id];
    }
    const float bg_dot_dpixel = p.background[0] * gp0 + p.background[1] * gp1 + p.background[2] * gp2;
    const float ddelx_dx = 0.5f * (float)p.W, ddely_dy = 0.5f * (float)p.H;

    // Reduce-scatter target of this lane: after the shuffles lane (grp = lane>>3, i = lane&7 < 3)
    // owns value 3*grp + i of { dcolor r,g,b | dmean2D x,y,z | dconic x,y,w | dflow x,y, dopacity }.
    const int grp = lane >> 3, li = lane & 7;
    float* red_base;
    int red_mul, red_off;
    if (grp == 0) { red_base = p.dL_dcolor; red_mul = 3; red_off = li; }
    else if (grp == 1) { red_base = p.dL_dmean2D; red_mul = 3; red_off = li; }
    else if (grp == 2) { red_base = p.dL_dconic; red_mul = 4; red_off = (li == 2) ? 3 : li; }
    else if (li < 2) { red_base = p.dL_dflows; red_mul = 2; red_off = li; }
    else { red_base = p.dL_dopacity; red_mul = 1; red_off = 0; }
    const bool red_lane = li < 3;
    const bool hi16 = (lane & 16) != 0, hi8 = (lane & 8) != 0;

    for (int b = 0; b < nb; ++b) {
        const int s = b % BB_STAGES;
        const uint32_t ph = (uint32_t)(b / BB_STAGES) & 1u;
        const int hi = n - b * BB_BATCH, lo = max(0, hi - BB_BATCH), cnt = hi - lo;
        mbar_wait(&sm.full[s], ph);
        if ((unsigned)lo < wmax) {
            const InstRec* st = sm.recs[s];
            for (int r0 = ((cnt - 1) >> 5) << 5; r0 >= 0; r0 -= 32) {
                const int j = r0 + lane;
                bool rel = false;
                if (j < cnt && (unsigned)(lo + j) < wmax) {
                    const float4 a = st[j].q0;
                    const float4 c = st[j].q1;
                    const float4 e = st[j].q3;
                    rel = rect_may_contribute(a.x, a.y, c.x, c.y, c.z, a.z, e.z, e.w, bx0, bx1, by0, by1);
                }
                uint32_t m = __ballot_sync(0xffffffffu, rel);
                while (m) {
                    const int k = 31 - __clz(m);
                    m &= ~(1u << k);
                    const InstRec* g = st + (r0 + k);
                    const unsigned int pos = (unsigned)(lo + r0 + k);   // "contributor" index
                    const float4 q0 = g->q0;
                    const float4 q1 = g->q1;
                    const float dx = fsub(q0.x, pxf);
                    const float dy = fsub(q0.y, pyf);
                    const float power =
                        ffma(ffma(dx, fmul(dx, q1.x), fmul(dy, fmul(dy, q1.z))), -0.5f, -fmul(dy, fmul(dx, q1.y)));
                    bool contrib = (pos < my_last) && !(power > 0.0f) && !(power < q0.z);
                    float G = 0.f, alpha = 0.f;
                    if (contrib) {
                        G = expf(power);
                        alpha = fminf(fmul(q1.w, G), 0.99f);
                        contrib = !(alpha < 1.0f / 255.0f);
                    }
                    if (!__any_sync(0xffffffffu, contrib)) continue;

                    float v[12];
#pragma unroll
                    for (int i = 0; i < 12; ++i) v[i] = 0.f;
                    if (contrib) {
                        const float4 q2 = g->q2;
                        const float4 q3 = g->q3;
                        // 1/(1-alpha): alpha <= 0.99, so the approximate reciprocal (1 ulp) is safe; the
                        // reference divides twice here (backward.cu:1056,1113)
                        const float om = 1.f - alpha;
                        float rom;
                        asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rom) : "f"(om));
                        T = T * rom;
                        const float dchannel_dcolor = alpha * T;
                        // colour / flow / depth / mask recurrences, backward.cu:1062-1102
                        float dL_dalpha = (q2.x - acc_c0) * gp0 + (q2.y - acc_c1) * gp1 + (q2.z - acc_c2) * gp2;
                        dL_dalpha += (q3.x - acc_f0) * gf0 + (q3.y - acc_f1) * gf1;
                        dL_dalpha += (q2.w - acc_d) * gd;
                        dL_dalpha += (1.0f - acc_m) * gm;
                        dL_dalpha *= T;
                        // accumulators as seen by the next (nearer) Gaussian
                        acc_c0 = alpha * q2.x + om * acc_c0;
                        acc_c1 = alpha * q2.y + om * acc_c1;
                        acc_c2 = alpha * q2.z + om * acc_c2;
                        acc_f0 = alpha * q3.x + om * acc_f0;
                        acc_f1 = alpha * q3.y + om * acc_f1;
                        acc_d = alpha * q2.w + om * acc_d;
                        acc_m = alpha + om * acc_m;
                        // background term, backward.cu:1110-1113
                        dL_dalpha += (-T_final * rom) * bg_dot_dpixel;

                        const float dL_dG = q1.w * dL_dalpha;
                        const float gdx = G * dx, gdy = G * dy;
                        const float dG_ddelx = -gdx * q1.x - gdy * q1.y;
                        const float dG_ddely = -gdy * q1.z - gdx * q1.y;
                        v[0] = dchannel_dcolor * gp0;
                        v[1] = dchannel_dcolor * gp1;
                        v[2] = dchannel_dcolor * gp2;
                        v[3] = dL_dG * dG_ddelx * ddelx_dx;
                        v[4] = dL_dG * dG_ddely * ddely_dy;
                        v[5] = gd * dchannel_dcolor;
                        v[6] = -0.5f * gdx * dx * dL_dG;
                        v[7] = -0.5f * gdx * dy * dL_dG;
                        v[8] = -0.5f * gdy * dy * dL_dG;
                        v[9] = dchannel_dcolor * gf0;
                        v[10] = dchannel_dcolor * gf1;
                        v[11] = G * dL_dalpha;
                    }
                    // warp reduce-scatter: 12 values x 32 lanes -> 12 lanes own one sum each
                    float w[6];
#pragma unroll
                    for (int i = 0; i < 6; ++i) {
                        const float send = hi16 ? v[i] : v[i + 6];
                        const float keep = hi16 ? v[i + 6] : v[i];
                        w[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
                    }
                    float u[3];
#pragma unroll
                    for (int i = 0; i < 3; ++i) {
                        const float send = hi8 ? w[i] : w[i + 3];
                        const float keep = hi8 ? w[i + 3] : w[i];
                        u[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
                    }
#pragma unroll
                    for (int i = 0; i < 3; ++i) {
                        u[i] += __shfl_xor_sync(0xffffffffu, u[i], 4);
                        u[i] += __shfl_xor_sync(0xffffffffu, u[i], 2);
                        u[i] += __shfl_xor_sync(0xffffffffu, u[i], 1);
                    }
                    if (red_lane) {
                        const float val = (li == 0) ? u[0] : (li == 1) ? u[1] : u[2];
                        const unsigned int gid = __float_as_uint(q0.w);
                        atomicAdd(red_base + (size_t)gid * red_mul + red_off, val);
                    }
                }
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&sm.empty[s]);
        if (threadIdx.x == 0 && issued < nb) {
            mbar_wait(&sm.empty[s], ph);
            const int hi2 = n - issued * BB_BATCH, lo2 = max(0, hi2 - BB_BATCH);
            mbar_expect_tx(&sm.full[s], (uint32_t)(hi2 - lo2) * 64u);
            bulk_g2s(&sm.recs[s][0], src + lo2, (uint32_t)(hi2 - lo2) * 64u, &sm.full[s]);
            ++issued;
        }
    }
}

}  // namespace

cudaError_t launch_blend_bwd(const BlendBwdParams& p, cudaStream_t stream) {
    dim3 grid(p.grid_x, p.grid_y, 1);
    blend_bwd_kernel<<<grid, BB_THREADS, 0, stream>>>(p);
    return cudaGetLastError();
}

}  // namespace fdgs
