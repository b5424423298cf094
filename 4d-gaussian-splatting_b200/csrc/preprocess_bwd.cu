// preprocess_bwd.cu -- per-Gaussian backward preprocessing (one thread per Gaussian).
//
// Fuses the reference's two backward preprocessing kernels -- computeCov2DCUDA
// (backward.cu:486-617) and preprocessCUDA<3> (backward.cu:839-923) with its helpers
// computeColorFromSH (:20-139), computeColorFromSH_4D (:144-481), computeCov3D (:621-684),
// computeCov3D_conditional (:689-834) -- into one pass, so dL_dcov3D / dL_dmeans never make a
// round trip through HBM between them.
//
// B200 design:
//   * every output row is written by this kernel (zeros for Gaussians that were not rendered), so
//     the host never zero-fills the eight parameter-gradient tensors (the reference memsets
//     716 B per Gaussian first, rasterize_points.cu:201-213);
//   * SH rows (12*M bytes) are streamed in with cp.async.bulk like the forward, and the
//     12*M-byte dL_dsh rows are staged in the same shared-memory rows and written out by the whole
//     CTA as fully coalesced 16-byte stores (the reference writes 12-byte pieces at a 12*M-byte
//     stride per thread).
// Bug-compatible with the reference where the bugs are observable:
// dL_dsh[1] uses l0m0 in the 4D variant (backward.cu:190), the temporal derivative has no minus
// sign (:303,:384) and is overwritten, not accumulated, for deg_t > 1 (:403).
#include "fdgs_internal.h"

namespace fdgs {
namespace {

constexpr int PB_THREADS = 128;

__device__ __forceinline__ float3 dnormvdv3(float3 v, float3 dv) {   // auxiliary.h:108-118
    const float sum2 = v.x * v.x + v.y * v.y + v.z * v.z;
    const float invsum32 = 1.0f / sqrtf(sum2 * sum2 * sum2);
    float3 o;
    o.x = ((+sum2 - v.x * v.x) * dv.x - v.y * v.x * dv.y - v.z * v.x * dv.z) * invsum32;
    o.y = (-v.x * v.y * dv.x + (sum2 - v.y * v.y) * dv.y - v.z * v.y * dv.z) * invsum32;
    o.z = (-v.x * v.z * dv.x - v.y * v.z * dv.y + (sum2 - v.z * v.z) * dv.z) * invsum32;
    return o;
}

struct ShDeriv {
    float l[16], dx[16], dy[16], dz[16];
};

// basis values and their derivatives w.r.t. the normalised direction (backward.cu:172-263)
__device__ __forceinline__ void sh_basis_deriv(float x, float y, float z, int deg, ShDeriv& S) {
#pragma unroll
    for (int i = 0; i < 16; ++i) S.l[i] = S.dx[i] = S.dy[i] = S.dz[i] = 0.f;
    S.l[0] = kSH_C0;
    if (deg > 0) {
        S.l[1] = -kSH_C1 * y; S.dy[1] = -kSH_C1;
        S.l[2] = kSH_C1 * z;  S.dz[2] = kSH_C1;
        S.l[3] = -kSH_C1 * x; S.dx[3] = -kSH_C1;
        if (deg > 1) {
            const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
            S.l[4] = kSH_C2[0] * xy; S.dx[4] = kSH_C2[0] * y; S.dy[4] = kSH_C2[0] * x;
            S.l[5] = kSH_C2[1] * yz; S.dy[5] = kSH_C2[1] * z; S.dz[5] = kSH_C2[1] * y;
            S.l[6] = kSH_C2[2] * (2.0f * zz - xx - yy);
            S.dx[6] = -2.f * kSH_C2[2] * x; S.dy[6] = -2.f * kSH_C2[2] * y; S.dz[6] = 4.f * kSH_C2[2] * z;
            S.l[7] = kSH_C2[3] * xz; S.dx[7] = kSH_C2[3] * z; S.dz[7] = kSH_C2[3] * x;
            S.l[8] = kSH_C2[4] * (xx - yy); S.dx[8] = 2.f * kSH_C2[4] * x; S.dy[8] = -2.f * kSH_C2[4] * y;
            if (deg > 2) {
                S.l[9] = kSH_C3[0] * y * (3.f * xx - yy);
                S.dx[9] = kSH_C3[0] * y * 6.f * x; S.dy[9] = kSH_C3[0] * (3.f * xx - 3.f * yy);
                S.l[10] = kSH_C3[1] * xy * z;
                S.dx[10] = kSH_C3[1] * yz; S.dy[10] = kSH_C3[1] * xz; S.dz[10] = kSH_C3[1] * xy;
                S.l[11] = kSH_C3[2] * y * (4.f * zz - xx - yy);
                S.dx[11] = -kSH_C3[2] * y * 2.f * x; S.dy[11] = kSH_C3[2] * (4.f * zz - xx - 3.f * yy);
                S.dz[11] = kSH_C3[2] * y * 8.f * z;
                S.l[12] = kSH_C3[3] * z * (2.f * zz - 3.f * xx - 3.f * yy);
                S.dx[12] = -kSH_C3[3] * z * 6.f * x; S.dy[12] = -kSH_C3[3] * z * 6.f * y;
                S.dz[12] = kSH_C3[3] * (6.f * zz - 3.f * xx - 3.f * yy);
                S.l[13] = kSH_C3[4] * x * (4.f * zz - xx - yy);
                S.dx[13] = kSH_C3[4] * (4.f * zz - 3.f * xx - yy); S.dy[13] = -kSH_C3[4] * x * 2.f * y;
                S.dz[13] = kSH_C3[4] * x * 8.f * z;
                S.l[14] = kSH_C3[5] * z * (xx - yy);
                S.dx[14] = kSH_C3[5] * z * 2.f * x; S.dy[14] = -kSH_C3[5] * z * 2.f * y;
                S.dz[14] = kSH_C3[5] * (xx - yy);
                S.l[15] = kSH_C3[6] * x * (xx - 3.f * yy);
                S.dx[15] = kSH_C3[6] * (3.f * xx - 3.f * yy); S.dy[15] = -kSH_C3[6] * x * 6.f * y;
            }
        }
    }
}

// 4x4 helpers, glm convention: A[c][r], (A*B)[c][r] = sum_k A[k][r] * B[c][k]
__device__ __forceinline__ void mat4_mul(const float A[4][4], const float B[4][4], float C[4][4]) {
#pragma unroll
    for (int c = 0; c < 4; ++c)
#pragma unroll
        for (int r = 0; r < 4; ++r)
            C[c][r] = A[0][r] * B[c][0] + A[1][r] * B[c][1] + A[2][r] * B[c][2] + A[3][r] * B[c][3];
}


// 3x3 helpers, same convention
__device__ __forceinline__ void mat3_mul(const float A[3][3], const float B[3][3], float C[3][3]) {
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
        for (int r = 0; r < 3; ++r) C[c][r] = A[0][r] * B[c][0] + A[1][r] * B[c][1] + A[2][r] * B[c][2];
}

template <bool BULK>
__global__ void __launch_bounds__(PB_THREADS) preprocess_bwd_kernel(const PreprocessBwdParams a) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ unsigned char row_live[PB_THREADS];

    const int idx = blockIdx.x * PB_THREADS + threadIdx.x;
    const bool in_range = idx < a.P;
    const bool vis = in_range && (a.radii[idx] > 0) && (a.tiles_touched[idx] != 0u);
    const int row_floats = 3 * a.M;
    const bool has_sh = (a.shs != nullptr) && a.M > 0;
    float* my_row = reinterpret_cast<float*>(smem_raw) + (size_t)threadIdx.x * a.sh_row_stride_floats;

    if (BULK) {
        if (threadIdx.x == 0) {
            mbar_init(&bar, 1);
            mbar_fence_init();
        }
        row_live[threadIdx.x] = vis ? 1 : 0;
        const int nvis = __syncthreads_count(vis);
        if (threadIdx.x == 0 && nvis > 0) mbar_expect_tx(&bar, (uint32_t)nvis * (uint32_t)row_floats * 4u);
        if (vis) bulk_g2s(my_row, a.shs + (size_t)idx * row_floats, (uint32_t)row_floats * 4u, &bar);
    }

    // gradients produced for this Gaussian (zeros unless rendered)
    float g_mean[3] = {0.f, 0.f, 0.f};
    float g_cov[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    float g_ts = 0.f, g_scale[3] = {0.f, 0.f, 0.f}, g_scale_t = 0.f;
    float g_rot[4] = {0.f, 0.f, 0.f, 0.f}, g_rotr[4] = {0.f, 0.f, 0.f, 0.f};
    float dRGB[3] = {0.f, 0.f, 0.f};
    float tw1 = 0.f, tw2 = 0.f;   // temporal weights of SH blocks 1 and 2 (0 = block inactive)
    ShDeriv SD;
    bool sh4d = false;

    if (vis) {
        const float* V = a.viewmatrix;
        const float mx = a.means3D[3 * idx + 0], my = a.means3D[3 * idx + 1], mz = a.means3D[3 * idx + 2];
        const float* c3 = a.cov3D + 6 * idx;

        // ---------------- computeCov2DCUDA, backward.cu:486-617 ----------------
        {
            const float dcx = a.dL_dconic[4 * idx + 0], dcy = a.dL_dconic[4 * idx + 1], dcz = a.dL_dconic[4 * idx + 3];
            Proj2D Pj;
            build_T(V, mx, my, mz, a.focal_x, a.focal_y, a.tan_fovx, a.tan_fovy, Pj);
            const float limx = fmul(1.3f, a.tan_fovx), limy = fmul(1.3f, a.tan_fovy);
            const float x_grad_mul = (Pj.txtz < -limx || Pj.txtz > limx) ? 0.f : 1.f;
            const float y_grad_mul = (Pj.tytz < -limy || Pj.tytz > limy) ? 0.f : 1.f;
            float ca, cb, cc;
            cov2d_from_T(Pj, c3, ca, cb, cc);
            const float A = ca + 0.3f, B = cb, C = cc + 0.3f;
            const float denom = A * C - B * B;
            float dL_da = 0.f, dL_db = 0.f, dL_dc = 0.f;
            const float denom2inv = 1.0f / ((denom * denom) + 0.0000001f);
            const float T00 = Pj.T00, T01 = Pj.T01, T02 = Pj.T02, T10 = Pj.T10, T11 = Pj.T11, T12 = Pj.T12;
            if (denom2inv != 0) {
                dL_da = denom2inv * (-C * C * dcx + 2 * B * C * dcy + (denom - A * C) * dcz);
                dL_dc = denom2inv * (-A * A * dcz + 2 * A * B * dcy + (denom - A * C) * dcx);
                dL_db = denom2inv * 2 * (B * C * dcx - (denom + 2 * B * B) * dcy + A * B * dcz);
                g_cov[0] = (T00 * T00 * dL_da + T00 * T10 * dL_db + T10 * T10 * dL_dc);
                g_cov[3] = (T01 * T01 * dL_da + T01 * T11 * dL_db + T11 * T11 * dL_dc);
                g_cov[5] = (T02 * T02 * dL_da + T02 * T12 * dL_db + T12 * T12 * dL_dc);
                g_cov[1] = 2 * T00 * T01 * dL_da + (T00 * T11 + T01 * T10) * dL_db + 2 * T10 * T11 * dL_dc;
                g_cov[2] = 2 * T00 * T02 * dL_da + (T00 * T12 + T02 * T10) * dL_db + 2 * T10 * T12 * dL_dc;
                g_cov[4] = 2 * T02 * T01 * dL_da + (T01 * T12 + T02 * T11) * dL_db + 2 * T11 * T12 * dL_dc;
            }
            // Vrk (symmetric): V0 = (c0,c1,c2), V1 = (c1,c3,c4), V2 = (c2,c4,c5)
            const float r0a = T00 * c3[0] + T01 * c3[1] + T02 * c3[2];
            const float r0b = T00 * c3[1] + T01 * c3[3] + T02 * c3[4];
            const float r0c = T00 * c3[2] + T01 * c3[4] + T02 * c3[5];
            const float r1a = T10 * c3[0] + T11 * c3[1] + T12 * c3[2];
            const float r1b = T10 * c3[1] + T11 * c3[3] + T12 * c3[4];
            const float r1c = T10 * c3[2] + T11 * c3[4] + T12 * c3[5];
            const float dL_dT00 = 2 * r0a * dL_da + r1a * dL_db;
            const float dL_dT01 = 2 * r0b * dL_da + r1b * dL_db;
            const float dL_dT02 = 2 * r0c * dL_da + r1c * dL_db;
            const float dL_dT10 = 2 * r1a * dL_dc + r0a * dL_db;
            const float dL_dT11 = 2 * r1b * dL_dc + r0b * dL_db;
            const float dL_dT12 = 2 * r1c * dL_dc + r0c * dL_db;
            // W[c][r] = view[4*r + c]   (backward.cu:525-528)
            const float dL_dJ00 = V[0] * dL_dT00 + V[4] * dL_dT01 + V[8] * dL_dT02;
            const float dL_dJ02 = V[2] * dL_dT00 + V[6] * dL_dT01 + V[10] * dL_dT02;
            const float dL_dJ11 = V[1] * dL_dT10 + V[5] * dL_dT11 + V[9] * dL_dT12;
            const float dL_dJ12 = V[2] * dL_dT10 + V[6] * dL_dT11 + V[10] * dL_dT12;
            const float tz = 1.f / Pj.tz, tz2 = tz * tz, tz3 = tz2 * tz;
            const float h_x = a.focal_x, h_y = a.focal_y;
            const float dL_dtx = x_grad_mul * -h_x * tz2 * dL_dJ02;
            const float dL_dty = y_grad_mul * -h_y * tz2 * dL_dJ12;
            const float dL_dtz = -h_x * tz2 * dL_dJ00 - h_y * tz2 * dL_dJ11 + (2 * h_x * Pj.tx) * tz3 * dL_dJ02 +
                                 (2 * h_y * Pj.ty) * tz3 * dL_dJ12;
            // transformVec4x3Transpose, plus the depth gradient riding in dL_dmean2D.z (:611)
            const float vx = dL_dtx, vy = dL_dty, vz = dL_dtz + a.dL_dmean2D[3 * idx + 2];
            g_mean[0] = V[0] * vx + V[1] * vy + V[2] * vz;
            g_mean[1] = V[4] * vx + V[5] * vy + V[6] * vz;
            g_mean[2] = V[8] * vx + V[9] * vy + V[10] * vz;
        }

        // ---------------- preprocessCUDA (backward), backward.cu:877-894 ----------------
        {
            const float* Pm = a.projmatrix;
            const float m_hw = Pm[3] * mx + Pm[7] * my + Pm[11] * mz + Pm[15];
            const float m_w = 1.0f / (m_hw + 0.0000001f);
            const float mul1 = (Pm[0] * mx + Pm[4] * my + Pm[8] * mz + Pm[12]) * m_w * m_w;
            const float mul2 = (Pm[1] * mx + Pm[5] * my + Pm[9] * mz + Pm[13]) * m_w * m_w;
            const float d2x = a.dL_dmean2D[3 * idx + 0], d2y = a.dL_dmean2D[3 * idx + 1];
            g_mean[0] += (Pm[0] * m_w - Pm[3] * mul1) * d2x + (Pm[1] * m_w - Pm[3] * mul2) * d2y;
            g_mean[1] += (Pm[4] * m_w - Pm[7] * mul1) * d2x + (Pm[5] * m_w - Pm[7] * mul2) * d2y;
            g_mean[2] += (Pm[8] * m_w - Pm[11] * mul1) * d2x + (Pm[9] * m_w - Pm[11] * mul2) * d2y;
        }

        // ---------------- SH backward, backward.cu:20-139 / :144-481 ----------------
        if (has_sh) {
            sh4d = !((a.gaussian_dim == 3) || a.force_sh_3d);
            const unsigned cl = a.clamped[idx];
            dRGB[0] = (cl & 1u) ? 0.f : a.dL_dcolor[3 * idx + 0];
            dRGB[1] = (cl & 2u) ? 0.f : a.dL_dcolor[3 * idx + 1];
            dRGB[2] = (cl & 4u) ? 0.f : a.dL_dcolor[3 * idx + 2];
            const float3 dir_orig = make_float3(mx - a.campos[0], my - a.campos[1], mz - a.campos[2]);
            const float len = sqrtf(dir_orig.x * dir_orig.x + dir_orig.y * dir_orig.y + dir_orig.z * dir_orig.z);
            const float x = dir_orig.x / len, y = dir_orig.y / len, z = dir_orig.z / len;
            sh_basis_deriv(x, y, z, a.D, SD);
            float dt1_dt = 0.f, dt2_dt = 0.f;
            if (sh4d && a.D > 2 && a.D_t > 0) {
                const float dir_t = a.ts[idx] - a.timestamp;
                const double w = 2 * FDGS_MY_PI * (double)dir_t / (double)a.time_duration;
                tw1 = (float)cos(w);
                dt1_dt = (float)(sin(w) * 2 * FDGS_MY_PI / (double)a.time_duration);
                if (a.D_t > 1) {
                    const double w2 = 2 * FDGS_MY_PI * (double)dir_t * 2 / (double)a.time_duration;
                    tw2 = (float)cos(w2);
                    dt2_dt = (float)(sin(w2) * 2 * FDGS_MY_PI * 2 / (double)a.time_duration);
                }
            }
            const bool blk1 = sh4d && a.D > 2 && a.D_t > 0, blk2 = blk1 && a.D_t > 1;
            if (BULK) mbar_wait(&bar, 0);
            // s_k = sh[k] . dL_dRGB ; direction / time gradients are weighted sums of s_k
            float ddx = 0.f, ddy = 0.f, ddz = 0.f, dtt = 0.f;
            const float* grow = a.shs + (size_t)idx * row_floats;
            const int ncoef = (a.D + 1) * (a.D + 1);
#pragma unroll 1
            for (int blk = 0; blk < 3; ++blk) {
                if (blk == 1 && !blk1) break;
                if (blk == 2 && !blk2) break;
                const float twt = (blk == 0) ? 1.f : (blk == 1 ? tw1 : tw2);
                float sx = 0.f, sy = 0.f, sz = 0.f, sl = 0.f;
#pragma unroll
                for (int k = 0; k < 16; ++k) {
                    if (k >= ncoef) break;
                    const int c = blk * 16 + k;
                    float s0, s1, s2;
                    if (BULK) { s0 = my_row[3 * c + 0]; s1 = my_row[3 * c + 1]; s2 = my_row[3 * c + 2]; }
                    else { s0 = __ldg(grow + 3 * c + 0); s1 = __ldg(grow + 3 * c + 1); s2 = __ldg(grow + 3 * c + 2); }
                    const float sk = s0 * dRGB[0] + s1 * dRGB[1] + s2 * dRGB[2];
                    sx += SD.dx[k] * sk;
                    sy += SD.dy[k] * sk;
                    sz += SD.dz[k] * sk;
                    sl += SD.l[k] * sk;
                }
                ddx += twt * sx;
                ddy += twt * sy;
                ddz += twt * sz;
                if (blk == 1) dtt = dt1_dt * sl;
                if (blk == 2) dtt = dt2_dt * sl;   // overwrites (backward.cu:403)
            }
            const float3 dm = dnormvdv3(dir_orig, make_float3(ddx, ddy, ddz));
            g_mean[0] += dm.x;
            g_mean[1] += dm.y;
            g_mean[2] += dm.z;
            if (sh4d) g_ts += dtt;
        }

        // ---------------- covariance backward ----------------
        if (a.has_scales) {
            const float mod = a.scale_modifier;
            if (a.rot_4d) {
                // backward.cu:689-834
                const float t = a.ts[idx];
                const float dt = fsub(a.timestamp, t);
                const float4 rot = reinterpret_cast<const float4*>(a.rotations)[idx];
                const float4 rotr = reinterpret_cast<const float4*>(a.rotations_r)[idx];
                const float sc[4] = {fmul(mod, a.scales[3 * idx + 0]), fmul(mod, a.scales[3 * idx + 1]),
                                     fmul(mod, a.scales[3 * idx + 2]), fmul(mod, a.scales_t[idx])};
                Sigma4 S;
                float R4[4][4];
                build_M4<true>(sc[0], sc[1], sc[2], sc[3], rot, rotr, S.M, R4);
                sigma_from_M(S);
                const float cov_t = S.s33;
                const float ctp = (a.prefilter_var > 0.0f) ? fadd(a.prefilter_var, cov_t) : cov_t;
                const float marginal = marginal_from(dt, cov_t, a.prefilter_var);
                if ((double)marginal > 0.05) {
                    const float c12[3] = {S.s03, S.s13, S.s23};
                    const float* d = g_cov;
                    float dc12[3];
                    dc12[0] = -(d[0] * c12[0] + d[1] * c12[1] * 0.5f + d[2] * c12[2] * 0.5f) * 2.0f / cov_t;
                    dc12[1] = -(d[1] * c12[0] * 0.5f + d[3] * c12[1] + d[4] * c12[2] * 0.5f) * 2.0f / cov_t;
                    dc12[2] = -(d[2] * c12[0] * 0.5f + d[4] * c12[1] * 0.5f + d[5] * c12[2]) * 2.0f / cov_t;
                    float dcovt = (c12[0] * c12[0] * d[0] + c12[0] * c12[1] * d[1] + c12[0] * c12[2] * d[2] +
                                   c12[1] * c12[1] * d[3] + c12[1] * c12[2] * d[4] + c12[2] * c12[2] * d[5]) /
                                  (cov_t * cov_t);
                    // opacity -> marginal chain (:769-774)
                    const float dop = a.dL_dopacity[idx];
                    const float dmarg = dop * a.opacities[idx];
                    a.dL_dopacity[idx] = dop * marginal;
                    const float dmarg_dcovt = marginal * dt * dt / 2 / (ctp * ctp);
                    const float dmarg_dt = marginal * dt / ctp;
                    dcovt += dmarg_dcovt * dmarg;
                    float dL_dt = dmarg * dmarg_dt;
                    // mean-shift chain (:777-782)
                    dc12[0] += g_mean[0] / cov_t * dt;
                    dc12[1] += g_mean[1] / cov_t * dt;
                    dc12[2] += g_mean[2] / cov_t * dt;
                    const float ddot = g_mean[0] * c12[0] + g_mean[1] * c12[1] + g_mean[2] * c12[2];
                    dcovt += -ddot / (cov_t * cov_t) * dt;
                    dL_dt += -ddot / cov_t;
                    g_ts += dL_dt;

                    float dS[4][4];
                    dS[0][0] = d[0]; dS[1][1] = d[3]; dS[2][2] = d[5]; dS[3][3] = dcovt;
                    dS[0][1] = dS[1][0] = 0.5f * d[1];
                    dS[0][2] = dS[2][0] = 0.5f * d[2];
                    dS[1][2] = dS[2][1] = 0.5f * d[4];
                    dS[0][3] = dS[3][0] = 0.5f * dc12[0];
                    dS[1][3] = dS[3][1] = 0.5f * dc12[1];
                    dS[2][3] = dS[3][2] = 0.5f * dc12[2];
                    float M2[4][4], dM[4][4];
#pragma unroll
                    for (int c = 0; c < 4; ++c)
#pragma unroll
                        for (int r = 0; r < 4; ++r) M2[c][r] = 2.0f * S.M[c][r];
                    mat4_mul(M2, dS, dM);   // dL_dM = 2 * M * dL_dSigma
                    // dL_dscale_i = sum_c R[c][i] * dL_dM[c][i],  R = M / s
                    float N[4][4];   // N[i][j] = s_i * dL_dM[j][i]   (= scaled dL_dMt)
                    float gs[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        float acc = 0.f;
#pragma unroll
                        for (int c = 0; c < 4; ++c) {
                            acc += R4[c][i] * dM[c][i];
                            N[i][c] = sc[i] * dM[c][i];
                        }
                        gs[i] = acc;
                    }
                    g_scale[0] = gs[0]; g_scale[1] = gs[1]; g_scale[2] = gs[2]; g_scale_t = gs[3];
                    const float qa = rot.x, qb = rot.y, qc = rot.z, qd = rot.w;
                    const float qp = rotr.x, qq = rotr.y, qr = rotr.z, qs = rotr.w;
                    const float Ml[4][4] = {{qa, qb, -qc, qd}, {-qb, qa, qd, qc}, {qc, -qd, qa, qb}, {-qd, -qc, -qb, qa}};
                    const float Mr[4][4] = {{qp, qq, -qr, -qs}, {-qq, qp, qs, -qr}, {qr, -qs, qp, -qq}, {qs, qr, qq, qp}};
                    float X[4][4], Y[4][4];
                    mat4_mul(N, Mr, X);    // dL_dml_t = dL_dMt * M_r
                    mat4_mul(Ml, N, Y);    // dL_dmr_t = M_l * dL_dMt
                    g_rot[0] = X[0][0] + X[1][1] + X[2][2] + X[3][3];
                    g_rot[1] = -X[0][1] + X[1][0] - X[2][3] + X[3][2];
                    g_rot[2] = X[0][2] - X[1][3] - X[2][0] + X[3][1];
                    g_rot[3] = -X[0][3] - X[1][2] + X[2][1] + X[3][0];
                    g_rotr[0] = Y[0][0] + Y[1][1] + Y[2][2] + Y[3][3];
                    g_rotr[1] = -Y[0][1] + Y[1][0] + Y[2][3] - Y[3][2];
                    g_rotr[2] = Y[0][2] + Y[1][3] - Y[2][0] - Y[3][1];
                    g_rotr[3] = Y[0][3] - Y[1][2] + Y[2][1] - Y[3][0];
                }
            } else {
                // backward.cu:621-684 (and no marginal-opacity gradient for non-rot 4D, :917-919)
                const float4 q = reinterpret_cast<const float4*>(a.rotations)[idx];
                const float sc[3] = {mod * a.scales[3 * idx + 0], mod * a.scales[3 * idx + 1], mod * a.scales[3 * idx + 2]};
                const float r = q.x, x = q.y, y = q.z, z = q.w;
                float R[3][3] = {{1.f - 2.f * (y * y + z * z), 2.f * (x * y - r * z), 2.f * (x * z + r * y)},
                                 {2.f * (x * y + r * z), 1.f - 2.f * (x * x + z * z), 2.f * (y * z - r * x)},
                                 {2.f * (x * z - r * y), 2.f * (y * z + r * x), 1.f - 2.f * (x * x + y * y)}};
                float M2[3][3];
#pragma unroll
                for (int c = 0; c < 3; ++c)
#pragma unroll
                    for (int rr = 0; rr < 3; ++rr) M2[c][rr] = 2.0f * sc[rr] * R[c][rr];
                const float* d = g_cov;
                const float dS[3][3] = {{d[0], 0.5f * d[1], 0.5f * d[2]}, {0.5f * d[1], d[3], 0.5f * d[4]},
                                        {0.5f * d[2], 0.5f * d[4], d[5]}};
                float dM[3][3];
                mat3_mul(M2, dS, dM);
                float Nt[3][3];   // dL_dMt after scaling: Nt[i][j] = s_i * dL_dM[j][i]
#pragma unroll
                for (int i = 0; i < 3; ++i) {
                    float acc = 0.f;
#pragma unroll
                    for (int c = 0; c < 3; ++c) {
                        acc += R[c][i] * dM[c][i];
                        Nt[i][c] = sc[i] * dM[c][i];
                    }
                    g_scale[i] = acc;
                }
                g_rot[0] = 2 * z * (Nt[0][1] - Nt[1][0]) + 2 * y * (Nt[2][0] - Nt[0][2]) + 2 * x * (Nt[1][2] - Nt[2][1]);
                g_rot[1] = 2 * y * (Nt[1][0] + Nt[0][1]) + 2 * z * (Nt[2][0] + Nt[0][2]) + 2 * r * (Nt[1][2] - Nt[2][1]) -
                           4 * x * (Nt[2][2] + Nt[1][1]);
                g_rot[2] = 2 * x * (Nt[1][0] + Nt[0][1]) + 2 * r * (Nt[2][0] - Nt[0][2]) + 2 * z * (Nt[1][2] + Nt[2][1]) -
                           4 * y * (Nt[2][2] + Nt[0][0]);
                g_rot[3] = 2 * r * (Nt[0][1] - Nt[1][0]) + 2 * x * (Nt[2][0] + Nt[0][2]) + 2 * y * (Nt[1][2] + Nt[2][1]) -
                           4 * z * (Nt[1][1] + Nt[0][0]);
            }
        }
    }

    // ---------------- outputs: every row of every overwritten tensor ----------------
    if (in_range) {
#pragma unroll
        for (int i = 0; i < 3; ++i) a.dL_dmean3D[3 * idx + i] = g_mean[i];
#pragma unroll
        for (int i = 0; i < 6; ++i) a.dL_dcov3D[6 * idx + i] = g_cov[i];
        a.dL_dts[idx] = g_ts;
#pragma unroll
        for (int i = 0; i < 3; ++i) a.dL_dscale[3 * idx + i] = g_scale[i];
        a.dL_dscale_t[idx] = g_scale_t;
        reinterpret_cast<float4*>(a.dL_drot)[idx] = make_float4(g_rot[0], g_rot[1], g_rot[2], g_rot[3]);
        reinterpret_cast<float4*>(a.dL_drot_r)[idx] = make_float4(g_rotr[0], g_rotr[1], g_rotr[2], g_rotr[3]);
    }

    if (a.dL_dsh != nullptr && a.M > 0) {
        // dL_dsh[k] = weight_k * dL_dRGB; weight 0 beyond the active degree
        const int ncoef = (a.D + 1) * (a.D + 1);
        auto weight = [&](int c) -> float {
            const int blk = c >> 4, k = c & 15;
            if (!vis || k >= ncoef || c >= a.M) return 0.f;
            if (blk == 0) return (sh4d && k == 1) ? SD.l[0] : SD.l[k];   // quirk: backward.cu:190
            if (!sh4d || !(a.D > 2)) return 0.f;
            if (blk == 1) return (a.D_t > 0) ? tw1 * SD.l[k] : 0.f;
            if (blk == 2) return (a.D_t > 1) ? tw2 * SD.l[k] : 0.f;
            return 0.f;
        };
        if (BULK) {
            // stage my row in shared memory (all lanes: zeros if not rendered), then the CTA writes the
            // contiguous [rows_in_block, 3M] slab with coalesced 16-byte stores
            if (vis && has_sh) {
                for (int c = 0; c < a.M; ++c) {
                    const float w = weight(c);
                    my_row[3 * c + 0] = w * dRGB[0];
                    my_row[3 * c + 1] = w * dRGB[1];
                    my_row[3 * c + 2] = w * dRGB[2];
                }
            }
            __syncthreads();
            const int rows_here = min(PB_THREADS, a.P - blockIdx.x * PB_THREADS);
            const int q_per_row = row_floats / 4;
            float4* dst = reinterpret_cast<float4*>(a.dL_dsh + (size_t)blockIdx.x * PB_THREADS * row_floats);
            const int total = rows_here * q_per_row;
            for (int f = threadIdx.x; f < total; f += PB_THREADS) {
                const int r = f / q_per_row, q = f - r * q_per_row;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (row_live[r])
                    v = reinterpret_cast<const float4*>(reinterpret_cast<float*>(smem_raw) +
                                                        (size_t)r * a.sh_row_stride_floats)[q];
                dst[f] = v;
            }
        } else if (in_range) {
            float* drow = a.dL_dsh + (size_t)idx * row_floats;
            for (int c = 0; c < a.M; ++c) {
                const float w = (vis && has_sh) ? weight(c) : 0.f;
                drow[3 * c + 0] = w * dRGB[0];
                drow[3 * c + 1] = w * dRGB[1];
                drow[3 * c + 2] = w * dRGB[2];
            }
        }
    }
}

}  // namespace

cudaError_t launch_preprocess_bwd(const PreprocessBwdParams& p, cudaStream_t stream) {
    if (p.P <= 0) return cudaSuccess;
    const int blocks = (p.P + PB_THREADS - 1) / PB_THREADS;
    const bool bulk = p.sh_bulk_ok && p.shs != nullptr && p.M > 0 && p.dL_dsh != nullptr;
    if (bulk) {
        const size_t smem = (size_t)PB_THREADS * p.sh_row_stride_floats * sizeof(float);
        static bool attr_set = false;
        if (!attr_set) {
            cudaError_t e = cudaFuncSetAttribute(preprocess_bwd_kernel<true>,
                                                 cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
            if (e != cudaSuccess) return e;
            attr_set = true;
        }
        preprocess_bwd_kernel<true><<<blocks, PB_THREADS, smem, stream>>>(p);
    } else {
        preprocess_bwd_kernel<false><<<blocks, PB_THREADS, 0, stream>>>(p);
    }
    return cudaGetLastError();
}

namespace {
// reference: rasterizer_impl.cu:54-67 checkFrustum + auxiliary.h:140-163
__global__ void mark_visible_kernel(int P, const float* __restrict__ means, const float* __restrict__ V,
                                    unsigned char* __restrict__ present) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= P) return;
    const float z = xform_row(V[2], V[6], V[10], V[14], means[3 * idx], means[3 * idx + 1], means[3 * idx + 2]);
    present[idx] = (z <= 0.2f) ? 0 : 1;
}
}  // namespace

cudaError_t launch_mark_visible(int P, const float* means3D, const float* viewmatrix, unsigned char* present,
                                cudaStream_t stream) {
    if (P <= 0) return cudaSuccess;
    mark_visible_kernel<<<(P + 255) / 256, 256, 0, stream>>>(P, means3D, viewmatrix, present);
    return cudaGetLastError();
}

}  // namespace fdgs
