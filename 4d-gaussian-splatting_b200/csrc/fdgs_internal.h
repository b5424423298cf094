// fdgs_internal.h -- kernel parameter blocks and launch prototypes (library-private).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <atomic>
#include "fdgs_common.cuh"

namespace fdgs {

// ---- per-device launch configuration -------------------------------------------------------------
// Function attributes (the > 48 KB dynamic shared-memory opt-in) belong to the (function, device) pair, so a
// process that renders on several GPUs has to set them once on EACH device; the flag word is per call site, one
// bit per device ordinal, updated atomically (two racing threads both set the idempotent attribute -- harmless).
struct PerDeviceOnce {
    std::atomic<unsigned long long> done{0ull};
};
template <class Kernel>
inline cudaError_t ensure_dynamic_smem(Kernel kernel, int bytes, PerDeviceOnce& once) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    const unsigned long long bit = 1ull << (dev & 63);
    if (dev < 64 && (once.done.load(std::memory_order_acquire) & bit)) return cudaSuccess;
    e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e == cudaSuccess && dev < 64) once.done.fetch_or(bit, std::memory_order_release);
    return e;
}
// SM count of the current device (148 on a B200), cached per device ordinal
int device_sm_count();

// ---- forward preprocess ---------------------------------------------------------------------
struct PreprocessFwdParams {
    int P, D, D_t, M;
    const float* means3D;
    const float* ts;
    const float* scales;
    const float* scales_t;
    float scale_modifier;
    const float* rotations;
    const float* rotations_r;
    const float* opacities;
    const float* shs;
    const float* cov3D_precomp;
    float prefilter_var;
    const float* colors_precomp;
    const float* viewmatrix;
    const float* projmatrix;
    const float* cam_pos;
    float timestamp, time_duration;
    int rot_4d, gaussian_dim, force_sh_3d;
    int W, H;
    float tan_fovx, tan_fovy, focal_x, focal_y;
    int grid_x, grid_y;
    int prefiltered;
    int sh_bulk_ok;             // SH rows can be streamed with cp.async.bulk (16-byte aligned rows)
    int sh_row_stride_floats;   // padded shared-memory row stride (multiple of 4 floats)
    int raw_params;             // scales / scales_t are log-scales, rotations un-normalised, opacities logits
    int quat_norm_mode;         // summation order of the quaternion norm (fdgs_common.cuh: quat_norm)
    const float* shs_rest;      // non-NULL: SH row split as shs = [P,1,3] (dc) | shs_rest = [P,M-1,3]
    // outputs
    const float* flows;     // [P,2] or NULL
    float* out_means3D;
    int* radii;
    float* cov3D;           // [P,6]
    InstRec* grec;          // [P] the 64-byte record every tile instance of the Gaussian copies
    uint8_t* clamped;       // bit ch set = channel ch was clamped at 0
    uint32_t* tiles_touched;
    uint4* binrec;          // [P] compact bin record of EVERY Gaussian (see launch_bin_count)
    int tile_cull;          // != 0: list a Gaussian only in the tiles its alpha >= 1/255 footprint reaches (fdgs_set_tile_cull)
};
cudaError_t launch_preprocess_fwd(const PreprocessFwdParams& p, cudaStream_t stream);
cudaError_t launch_debug_activate(int n, const float* log_s, const float* logit, const float* quat, int mode, float* s_out,
                                  float* o_out, float* q_out, cudaStream_t stream);

// ---- binning -----------------------------------------------------------------------------------
// counting sort of the (Gaussian, tile) instances by tile, BIN_CTAS persistent CTAs (binning.cu)
int bin_ctas();                                 // rows of the [bin_ctas()][num_tiles] count matrix
// binrec[i] = { x0 | y0 << 16, x1 | y1 << 16, depth bits, tile mask }: tile rectangle (empty = not listed anywhere) of
// Gaussian i; for rectangles of up to 32 tiles bit j of the mask = tile j (row-major) takes an instance
cudaError_t launch_bin_count(int P, const uint4* binrec, int grid_x, int grid_y, uint32_t* matrix, cudaStream_t stream);
// matrix -> per-CTA column prefixes; tile_offset / ranges; info[0] = total instances R, info[1] = largest tile
cudaError_t launch_tile_scan(int num_tiles, uint32_t* matrix, uint32_t* tile_total, uint32_t* tile_offset, uint2* ranges,
                             uint32_t* info, cudaStream_t stream);
cudaError_t launch_bin_scatter(int P, const uint4* binrec, int grid_x, int grid_y, uint32_t* matrix,
                               const uint32_t* tile_offset, uint64_t* keys, cudaStream_t stream);
// sorts every tile's keys and writes the 64-byte instance records + the sorted index list
int tile_sort_pack_kernel_count(int max_count);
cudaError_t launch_tile_sort_pack(int num_tiles, int max_count, int R, const uint2* ranges, uint64_t* keys,
                                  const InstRec* grec, StageRec* recs, uint32_t* point_list, cudaStream_t stream);
cudaError_t launch_unpack_grec(int P, const InstRec* grec, const int* radii, float* depths, float* means2D,
                               float* conic_opacity, float* rgb, cudaStream_t stream);

// ---- blend ---------------------------------------------------------------------------------------
struct BlendFwdParams {
    int W, H, grid_x, grid_y;
    const uint2* ranges;
    const StageRec* recs;
    const float* background;
    float* final_T;        // [H*W]  (scratch copy used by the backward)
    uint32_t* n_contrib;   // [H*W]
    float* out_color;      // [3,H,W]
    float* out_flow;       // [2,H,W]
    float* out_depth;      // [H,W]
    float* out_T;          // [H,W]
};
cudaError_t launch_blend_fwd(const BlendFwdParams& p, cudaStream_t stream);

struct BlendBwdParams {
    int W, H, grid_x, grid_y;
    const uint2* ranges;
    const StageRec* recs;
    const float* background;
    const float* final_T;
    const uint32_t* n_contrib;
    const float* dL_dpix;
    const float* dL_depths;
    const float* dL_masks;
    const float* dL_dpix_flow;
    float* dL_dmean2D;    // [P,3]
    float* dL_dconic;     // [P,4]
    float* dL_dopacity;   // [P]
    float* dL_dcolor;     // [P,3]
    float* dL_dflows;     // [P,2]
};
cudaError_t launch_blend_bwd(const BlendBwdParams& p, cudaStream_t stream);
// true when launch_blend_bwd(p) leaves RAW moment sums (W10, W01 | W20, W11, W02 | W00 of w = G dL/dalpha
// about each Gaussian's mean) in dL_dmean2D.xy / dL_dconic / dL_dopacity instead of the final
// gradients; launch_preprocess_bwd finishes them (PreprocessBwdParams::blend_raw)
bool blend_bwd_is_raw(const BlendBwdParams& p);

// ---- backward preprocess ------------------------------------------------------------------------
struct PreprocessBwdParams {
    int P, D, D_t, M;
    const float* means3D;   // shifted means (out_means3D of the forward)
    const int* radii;
    const float* shs;
    const float* ts;
    const float* opacities;
    const uint8_t* clamped;
    const uint32_t* tiles_touched;
    const float* scales;
    const float* scales_t;
    const float* rotations;
    const float* rotations_r;
    float scale_modifier;
    const float* cov3D;     // precomputed or the forward's
    float prefilter_var;
    const float* viewmatrix;
    const float* projmatrix;
    float focal_x, focal_y, tan_fovx, tan_fovy;
    const float* campos;
    float timestamp, time_duration;
    int rot_4d, gaussian_dim, force_sh_3d;
    int has_scales;         // scales != NULL in the reference's sense (backward.cu:908)
    int sh_bulk_ok;         // SH rows are 16-byte aligned and 16-byte multiples: stream with cp.async.bulk
    int sh_row_stride_floats;
    int raw_params;         // as in the forward: gradients then come out w.r.t. the RAW parameters
    int quat_norm_mode;
    const float* shs_rest;  // split SH input (see PreprocessFwdParams)
    float* dL_dsh_rest;     // [P,M-1,3] when shs_rest is given (dL_dsh is then [P,1,3])
    float* dL_dmean2D;      // in/out when blend_raw
    float* dL_dconic;       // in/out when blend_raw
    float* dL_dopacity;     // in/out
    const float* dL_dcolor;
    const InstRec* grec;    // per-Gaussian records of the forward (conic, blended opacity)
    int blend_raw;          // dL_dmean2D.xy / dL_dconic hold raw moment sums (blend_bwd v2)
    int W, H;
    float* dL_dmean3D;
    float* dL_dcov3D;
    float* dL_dsh;          // [P,M,3], or NULL when sh_factors is set
    float* sh_factors;      // [P,3] clamp-masked colour gradient (view-parallel mode) instead of the dL_dsh rows
    float* dL_dts;
    float* dL_dscale;
    float* dL_dscale_t;
    float* dL_drot;
    float* dL_drot_r;
};
cudaError_t launch_preprocess_bwd(const PreprocessBwdParams& p, cudaStream_t stream);
int preprocess_bwd_kernel_count(const PreprocessBwdParams& p);

// ---- view-parallel SH gradient (preprocess_bwd.cu: sh_outer_sum_kernel) -----------------------------------------
struct ShSumParams {
    int P, V, K;
    const float* table;       // [V][view_stride] floats: K rows of 3 (clamp-masked colour gradient), then metadata
    long long view_stride;    // floats per view block
    long long meta_off;       // offset of the view's metadata inside its block: timestamp, campos x, y, z
    const int* slot_of;       // [P] row of the Gaussian inside a view block, -1 = rendered by no view
    const long long* union_idx;   // [K] the inverse map: Gaussian of every row (sorted)
    float* dirs;              // [V][K][8] scratch: direction, temporal weights and colour factor per (view, row)
    const float* means3D;     // replicated Gaussian parameters, as handed to the rasterizer
    const float* ts;
    const float* scales;
    const float* scales_t;
    const float* rotations;
    const float* rotations_r;
    float scale_modifier, time_duration;
    int rot_4d, gaussian_dim, force_sh_3d, D, D_t, M;
    float* out0;              // [P, m0, 3]
    int m0;                   // coefficients 0 .. m0-1 of a row live in out0 ...
    float* out1;              // ... the rest in out1 [P, M - m0, 3] (NULL when m0 == M)
    int accumulate;           // add to the tensors' contents instead of overwriting them
};
cudaError_t launch_sh_outer_sum(const ShSumParams& p, cudaStream_t stream);
// flag[0] |= 1 if any row r with mask[r] == 0 of any of the n tensors has a non-zero element
cudaError_t launch_rows_zero_check(int n, const float* const* tensors, const int* widths, long long P, const int* mask_radii,
                                   int* flag, cudaStream_t stream);

cudaError_t launch_union_maps(long long P, const int* radii, const int* cs, int* slot_of, long long* idx, cudaStream_t stream);
cudaError_t launch_view_stats(long long P, const float* vgrad, int vstride, const int* radii, float* grad_norm_sum,
                              float* visibility_count, int* max_radii, cudaStream_t stream);
cudaError_t launch_pack_rows(bool unpack, int n, float* const* tensors, const int* widths, const long long* block_off,
                             const long long* idx, long long K, float* flat, cudaStream_t stream);

// ---- training-step neighbours of the rasterizer -----------------------------------------
// fused (1 - lambda) L1 + lambda (1 - SSIM): loss.cu
cudaError_t launch_l1_ssim_fwd(const float* x, const float* y, int C, int H, int W, float* maps, double* sums, cudaStream_t stream);
cudaError_t launch_l1_ssim_bwd(const float* x, const float* y, int C, int H, int W, const float* maps, const float* grad_scale,
                               float lambda_dssim, float* dL_dx, cudaStream_t stream);
// fused multi-tensor Adam, dense or over listed rows: optim.cu
cudaError_t launch_adam(int n, float* const* params, float* const* grads, float* const* m, float* const* v, const int* widths,
                        const float* step_sizes, long long rows, const long long* row_idx, float w1, float beta2, float w2,
                        float eps, float sqrt_bc2, int zero_grad, cudaStream_t stream);
// uniform-grid k nearest neighbours: knn.cu
size_t knn_scratch_bytes(int n);
cudaError_t launch_knn(int n, int k, const float* xyz, char* scratch, int* idx, float* dist2, cudaStream_t stream);
cudaError_t launch_knn_bruteforce(int n, int k, const float* xyz, int* idx, float* dist2, cudaStream_t stream);

cudaError_t launch_mark_visible(int P, const float* means3D, const float* viewmatrix, unsigned char* present,
                                cudaStream_t stream);

}  // namespace fdgs
