// fdgs_internal.h -- kernel parameter blocks and launch prototypes (library-private).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "fdgs_common.cuh"

namespace fdgs {

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
    // outputs
    float* out_means3D;
    int* radii;
    float* means2D;         // float2[P]
    float* depths;
    float* cov3D;           // [P,6]
    float* rgb;             // [P,3]
    float* conic_opacity;   // float4[P]
    uint8_t* clamped;       // bit ch set = channel ch was clamped at 0
    uint32_t* tiles_touched;
};
cudaError_t launch_preprocess_fwd(const PreprocessFwdParams& p, cudaStream_t stream);

// ---- binning -----------------------------------------------------------------------------------
size_t scan_temp_bytes(int P);
cudaError_t launch_scan(void* temp, size_t temp_bytes, const uint32_t* in, uint32_t* out, int P,
                        cudaStream_t stream);
cudaError_t launch_emit_keys(int P, const float* means2D, const float* depths, const uint32_t* offsets,
                             const int* radii, int grid_x, int grid_y, uint64_t* keys, uint32_t* vals,
                             cudaStream_t stream);
size_t sort_temp_bytes(int R);
cudaError_t launch_sort_pairs(void* temp, size_t temp_bytes, const uint64_t* keys_in, uint64_t* keys_out,
                              const uint32_t* vals_in, uint32_t* vals_out, int R, int end_bit,
                              cudaStream_t stream);
// builds the 64-byte instance records in tile-sorted order and the per-tile [start,end) ranges
cudaError_t launch_pack_instances(int R, const uint64_t* keys_sorted, const uint32_t* point_list,
                                  const float* means2D, const float* conic_opacity, const float* rgb,
                                  const float* depths, const float* flows, InstRec* recs, uint2* ranges,
                                  cudaStream_t stream);

// ---- blend ---------------------------------------------------------------------------------------
struct BlendFwdParams {
    int W, H, grid_x, grid_y;
    const uint2* ranges;
    const InstRec* recs;
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
    const InstRec* recs;
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
    const float* dL_dmean2D;
    const float* dL_dconic;
    float* dL_dopacity;     // in/out
    const float* dL_dcolor;
    float* dL_dmean3D;
    float* dL_dcov3D;
    float* dL_dsh;
    float* dL_dts;
    float* dL_dscale;
    float* dL_dscale_t;
    float* dL_drot;
    float* dL_drot_r;
};
cudaError_t launch_preprocess_bwd(const PreprocessBwdParams& p, cudaStream_t stream);

cudaError_t launch_mark_visible(int P, const float* means3D, const float* viewmatrix, unsigned char* present,
                                cudaStream_t stream);

}  // namespace fdgs
