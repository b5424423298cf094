/*
 * fdgs.h -- C-ABI of the B200-native differentiable 4D Gaussian rasterizer ("fdgs").
 *
 * This is the drop-in boundary of the hot path.  Every entry point replaces one
 * method of the reference's L0 C++ API `CudaRasterizer::Rasterizer`
 * (reference: diff-gaussian-rasterization/cuda_rasterizer/rasterizer.h:20-113):
 *
 *   fdgs_forward       <- Rasterizer::forward      (rasterizer.h:31-65,  rasterizer_impl.cu:199-364)
 *   fdgs_backward      <- Rasterizer::backward     (rasterizer.h:67-112, rasterizer_impl.cu:368-496)
 *   fdgs_mark_visible  <- Rasterizer::markVisible  (rasterizer.h:24-29,  rasterizer_impl.cu:142-154)
 *
 * Differences from the reference interface, all forced by "plain C, no C++ types":
 *   - the three `std::function<char*(size_t)>` scratch allocators become a C
 *     function pointer + opaque context (fdgs_alloc_fn);
 *   - the ~40 positional arguments become two POD structs of raw device
 *     pointers and scalars (field names = the reference's argument names);
 *   - every call takes the CUDA stream to launch on (the reference uses the
 *     legacy default stream implicitly) and returns an int status instead of
 *     throwing; fdgs_last_error() gives the message for the calling thread.
 *
 * All pointers are DEVICE pointers unless stated otherwise.  Nullable inputs
 * (colors_precomp, cov3D_precomp, shs, scales ...) follow the reference: NULL
 * means "not provided" (reference: rasterize_points.cu:110-141 passes the
 * data pointer of an empty tensor, i.e. NULL).
 *
 * The library never allocates or frees device memory itself; the three scratch
 * buffers are obtained through the callbacks and are opaque to the caller
 * (their layout is private and differs from the reference's GeometryState /
 * BinningState / ImageState; they only have to be handed back unchanged to
 * fdgs_backward, like the reference's geomBuffer/binningBuffer/imgBuffer).
 */
#ifndef FDGS_H_INCLUDED
#define FDGS_H_INCLUDED

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FDGS_VERSION 2

/* status codes */
#define FDGS_OK 0
#define FDGS_ERR_INVALID_ARG 1
#define FDGS_ERR_CUDA 2
#define FDGS_ERR_ALLOC 3
#define FDGS_ERR_UNSUPPORTED 4

/* Scratch allocator callback: must return a device pointer to at least
 * `bytes` bytes (128-byte aligned or better), or NULL on failure.
 * Replaces std::function<char*(size_t)> (reference: rasterizer.h:32-34,
 * rasterize_points.cu:28-34 `resizeFunctional`). */
typedef char* (*fdgs_alloc_fn)(void* ctx, size_t bytes);

/* Inputs/outputs of one forward pass.  Field names follow
 * Rasterizer::forward (reference: rasterizer.h:31-65). */
typedef struct fdgs_forward_args {
    int P;                 /* number of Gaussians                              */
    int D;                 /* active spatial SH degree (0..3)                  */
    int D_t;               /* active temporal SH degree (0..2)                 */
    int M;                 /* SH coefficients per Gaussian (row length / 3)    */
    const float* background;      /* [3]                                       */
    int width, height;
    const float* means3D;         /* [P,3]                                     */
    const float* shs;             /* [P,M,3] or NULL                           */
    const float* colors_precomp;  /* [P,3] or NULL                             */
    const float* flows_precomp;   /* [P,2]                                     */
    const float* opacities;       /* [P]                                       */
    const float* ts;              /* [P] or NULL                               */
    const float* scales;          /* [P,3] or NULL                             */
    const float* scales_t;        /* [P] or NULL                               */
    float scale_modifier;
    const float* rotations;       /* [P,4] or NULL                             */
    const float* rotations_r;     /* [P,4] or NULL                             */
    const float* cov3D_precomp;   /* [P,6] or NULL                             */
    float prefilter_var;
    const float* viewmatrix;      /* [16], column-major (transposed torch)     */
    const float* projmatrix;      /* [16], column-major                        */
    const float* cam_pos;         /* [3]                                       */
    float timestamp;
    float time_duration;
    int rot_4d;
    int gaussian_dim;
    int force_sh_3d;
    float tan_fovx, tan_fovy;
    int prefiltered;
    int debug;             /* !=0: synchronise + check after every stage       */
    /* outputs */
    float* out_means3D;    /* [P,3]  written for every Gaussian (shifted mean) */
    float* out_color;      /* [3,H,W]                                          */
    float* out_flow;       /* [2,H,W]                                          */
    float* out_depth;      /* [1,H,W]                                          */
    float* out_T;          /* [1,H,W] final transmittance                      */
    int* radii;            /* [P]                                              */
    /* Raw-parameter entry (version 2).  The reference activates the optimiser's raw
     * parameters in PyTorch before every render (scene/gaussian_model.py:179-219: exp, sigmoid, F.normalize, and a
     * torch.cat of features_dc / features_rest); with raw_params != 0 the kernels do it themselves:
     *   scales, scales_t   are LOG-scales            (exp applied)
     *   rotations(_r)      are un-normalised         (q / max(||q||, 1e-12))
     *   opacities          are logits                (sigmoid)
     * and, independently, a non-NULL shs_rest splits the SH row like the reference model's two tensors:
     *   shs = features_dc [P,1,3], shs_rest = features_rest [P,M-1,3]  (no 1.15 GB concatenation per frame). */
    int raw_params;
    const float* shs_rest;
} fdgs_forward_args;

/* Scratch handles returned by fdgs_forward (device pointers obtained through
 * the callbacks) plus what the torch shim needs to build `covs3D_com`. */
typedef struct fdgs_forward_result {
    int num_rendered;      /* R = number of (tile, Gaussian) instances         */
    char* geom_buffer;
    char* binning_buffer;
    char* image_buffer;
    size_t geom_bytes, binning_bytes, image_bytes;
    const float* cov3D;    /* [P,6] inside geom_buffer (reference: rasterize_points.cu:144-147) */
} fdgs_forward_result;

/* Inputs/outputs of one backward pass.  Field names follow
 * Rasterizer::backward (reference: rasterizer.h:67-112). */
typedef struct fdgs_backward_args {
    int P, D, D_t, M, R;
    const float* background;
    int width, height;
    const float* out_means3D;     /* shifted means written by the forward      */
    const float* shs;
    const float* colors_precomp;
    const float* flows_2d;
    const float* opacities;
    const float* ts;
    const float* scales;
    const float* scales_t;
    float scale_modifier;
    const float* rotations;
    const float* rotations_r;
    const float* cov3D_precomp;
    float prefilter_var;
    const float* viewmatrix;
    const float* projmatrix;
    const float* campos;
    float timestamp;
    float time_duration;
    int rot_4d;
    int gaussian_dim;
    int force_sh_3d;
    float tan_fovx, tan_fovy;
    const int* radii;
    const char* geom_buffer;
    const char* binning_buffer;
    const char* image_buffer;
    const float* dL_dpix;         /* [3,H,W]                                   */
    const float* dL_depths;       /* [1,H,W]  or NULL = no gradient (all zero) */
    const float* dL_masks;        /* [1,H,W]  grad of alpha = 1 - T, or NULL   */
    const float* dL_dpix_flow;    /* [2,H,W]  or NULL                          */
    int debug;
    /* Outputs.  The five "blend" gradients must be ZERO on entry (they are
     * accumulated into, like the reference: rasterize_points.cu:201-213); the
     * others are fully overwritten for every Gaussian (zeros where a Gaussian
     * was not rendered), so they may be uninitialised on entry. */
    float* dL_dmean2D;     /* [P,3]   accumulated                              */
    float* dL_dconic;      /* [P,2,2] accumulated (x, y, -, w)                 */
    float* dL_dopacity;    /* [P]     accumulated, then scaled by marginal_t   */
    float* dL_dcolor;      /* [P,3]   accumulated                              */
    float* dL_dflows;      /* [P,2]   accumulated                              */
    float* dL_dmean3D;     /* [P,3]   overwritten                              */
    float* dL_dcov3D;      /* [P,6]   overwritten                              */
    float* dL_dsh;         /* [P,M,3] overwritten                              */
    float* dL_dts;         /* [P]     overwritten                              */
    float* dL_dscale;      /* [P,3]   overwritten                              */
    float* dL_dscale_t;    /* [P]     overwritten                              */
    float* dL_drot;        /* [P,4]   overwritten                              */
    float* dL_drot_r;      /* [P,4]   overwritten                              */
    /* View-parallel mode (version 2; multi-GPU, see fdgs_sh_outer_sum below): when `sh_factors` is non-NULL and
     * dL_dsh is NULL the 12*M-byte dL_dsh rows are NOT written; instead sh_factors[P,3] receives the clamp-masked
     * colour gradient of every Gaussian (zeros where this view did not render it) -- the only view-dependent factor
     * of the rank-one dL_dsh row.  The direction / time terms of the SH backward still go to dL_dmean3D / dL_dts. */
    float* sh_factors;     /* [P,3]   overwritten, or NULL                     */
    /* Raw-parameter entry, as in fdgs_forward_args: the SAME values as in the forward.  With raw_params != 0 the
     * gradients dL_dscale, dL_dscale_t, dL_drot, dL_drot_r and dL_dopacity come out w.r.t. the RAW parameters (chain
     * rule through exp / normalize / sigmoid applied in the kernel).  With shs_rest the SH gradient is split the same
     * way: dL_dsh = [P,1,3], dL_dsh_rest = [P,M-1,3]. */
    int raw_params;
    const float* shs_rest;
    float* dL_dsh_rest;
} fdgs_backward_args;

/* Library / build identification. */
int fdgs_version(void);
/* Message of the last error on the calling thread ("" if none). */
const char* fdgs_last_error(void);

/* Sizes of the private scratch buffers (bytes), for callers that want to
 * pre-allocate (reference: required<GeometryState>(P) etc.,
 * rasterizer_impl.h:67-73). */
size_t fdgs_geom_bytes(int P);
size_t fdgs_image_bytes(int width, int height);
size_t fdgs_binning_bytes(int num_rendered, int width, int height);

/* Forward.  `stream` is a cudaStream_t passed as void*.  Performs exactly one
 * host synchronisation on `stream` (to learn num_rendered and size the binning
 * buffer), like the reference (rasterizer_impl.cu:302). */
int fdgs_forward(const fdgs_forward_args* args,
                 fdgs_alloc_fn geom_alloc, void* geom_ctx,
                 fdgs_alloc_fn binning_alloc, void* binning_ctx,
                 fdgs_alloc_fn image_alloc, void* image_ctx,
                 void* stream,
                 fdgs_forward_result* result);

/* Backward.  Fully asynchronous on `stream`. */
int fdgs_backward(const fdgs_backward_args* args, void* stream);

/* present[i] = (view-space z of means3D[i] > 0.2)
 * (reference: rasterizer_impl.cu:54-67, auxiliary.h:140-163). */
int fdgs_mark_visible(int P, const float* means3D, const float* viewmatrix,
                      const float* projmatrix, unsigned char* present, void* stream);

/* Multi-GPU gradient exchange helpers (new: the reference has no multi-GPU path; it sums the gradients of
 * sequential views, train.py:104-166).  pack: flat[block_off[t] + r*width[t] + c] = tensors[t][idx[r]*width[t] + c]
 * for r < K; unpack: the inverse scatter.  `tensors` / `widths` / `block_off` are HOST arrays of n <=
 * FDGS_MAX_PACK entries, `idx` (sorted row indices, int64) and `flat` are device pointers. */
#define FDGS_MAX_PACK 16
int fdgs_pack_rows(int n, const float* const* tensors, const int* widths, const long long* block_off,
                   const long long* idx, long long K, float* flat, void* stream);
int fdgs_unpack_rows(int n, float* const* tensors, const int* widths, const long long* block_off,
                     const long long* idx, long long K, const float* flat, void* stream);

/* View-parallel SH gradient (multi-GPU; new -- the reference sums the dL_dsh of sequential views in autograd,
 * train.py:104-166).  After the colour factors of all V views of a step have been all-gathered for the K Gaussians
 * of the union (table = V blocks of `view_stride` floats: K rows of 3 floats, then at `meta_off` the view's
 * timestamp and camera position), rebuilds every view's rank-one dL_dsh row from the replicated Gaussian
 * parameters and sums them in view order: out[i] = sum_v w(dir_v(i), t_v) x factors_v[slot_of[i]], zeros for
 * slot_of[i] < 0.  Bit-identical to accumulating the views' dL_dsh tensors one after the other.  The row may be
 * split over two tensors (out0 = first m0 coefficients, out1 = the rest; out1 NULL and m0 = M for one tensor). */
typedef struct fdgs_sh_sum_args {
    int P, V, K;
    const float* table;
    long long view_stride;
    long long meta_off;
    const int* slot_of;           /* [P]                                       */
    const long long* union_idx;   /* [K] Gaussian index of every union row (ascending): the inverse of slot_of */
    float* dir_scratch;           /* [V*K*8] floats of scratch                 */
    const float* means3D;         /* [P,3] the rasterizer's inputs (replicated) */
    const float* ts;              /* [P] or NULL                               */
    const float* scales;          /* [P,3] (needed when rot_4d)                */
    const float* scales_t;        /* [P]                                       */
    const float* rotations;       /* [P,4]                                     */
    const float* rotations_r;     /* [P,4]                                     */
    float scale_modifier;
    float time_duration;
    int rot_4d, gaussian_dim, force_sh_3d;
    int D, D_t, M;
    float* out0;
    int m0;
    float* out1;
    int accumulate;
} fdgs_sh_sum_args;
int fdgs_sh_outer_sum(const fdgs_sh_sum_args* args, void* stream);

/* flag[0] |= 1 if a row r with radii[r] <= 0 of any of the n tensors ([P, widths[i]], device) holds a non-zero
 * element: the guard of the sparse (union-rows-only) gradient exchange.  `tensors` / `widths` are HOST arrays. */
int fdgs_check_rows_zero(int n, const float* const* tensors, const int* widths, long long P, const int* radii,
                         int* flag, void* stream);

/* Densification statistics of ONE view, accumulated in one pass (reference: train.py:164-183 -- viewspace gradient norm,
 * visibility, max radius per Gaussian): grad_norm_sum[i] += ||viewspace_grad[i, 0:2]||, visibility_count[i] += radii[i] > 0,
 * max_radii[i] = max(max_radii[i], radii[i]).  grad_stride = floats per row of viewspace_grad (3 for means2D.grad). */
int fdgs_view_stats(long long P, const float* viewspace_grad, int grad_stride, const int* radii, float* grad_norm_sum,
                    float* visibility_count, int* max_radii, void* stream);

/* ---- the callers either side of the rasterizer in a training step ---------------------- */

/* Fused photometric loss (reference: utils/loss_utils.py:18-64 l1_loss + ssim, combined as train.py:115-117):
 *   loss = (1 - lambda) * mean|x - y| + lambda * (1 - mean(SSIM(x, y))),  11x11 Gaussian window, sigma 1.5, zero padding.
 * forward : sums[0] = sum|x - y|, sums[1] = sum SSIM map (device doubles; loss = combine on the device or host),
 *           maps[3,C,H,W] = per-pixel partial derivatives kept for the backward.
 * backward: dL_dx[C,H,W] = grad_scale[0] * d loss / d x   (grad_scale: device scalar, NULL = 1). */
int fdgs_l1_ssim_forward(const float* x, const float* y, int C, int H, int W, float* maps, double* sums, void* stream);
int fdgs_l1_ssim_backward(const float* x, const float* y, int C, int H, int W, const float* maps, const float* grad_scale,
                          float lambda_dssim, float* dL_dx, void* stream);

/* Fused multi-tensor Adam step (reference: torch.optim.Adam(eps=1e-15) over the parameter groups of
 * scene/gaussian_model.py:331-357, stepped at train.py:248-249).  n <= FDGS_MAX_PACK tensors [P, widths[i]] (device),
 * one learning rate each; `step` is the 1-based step count (bias correction).  beta1 / beta2 / eps are DOUBLES like the
 * Python floats torch receives: the weights 1 - beta are formed in double and then rounded to fp32 (1 - 0.999 in fp32
 * arithmetic is 1.3e-5 off 0.001f).  rows = NULL: every row (torch's dense
 * semantics); otherwise only the `num_rows` listed rows are touched (sparse Adam over the rendered Gaussians).
 * zero_grad != 0 clears the gradient elements it consumed. */
int fdgs_adam_step(int n, float* const* params, float* const* grads, float* const* exp_avg, float* const* exp_avg_sq,
                   const int* widths, const float* lrs, long long P, const long long* rows, long long num_rows,
                   long long step, double beta1, double beta2, double eps, int zero_grad, void* stream);

/* k nearest neighbours (k <= 32) of every point among the same n points, squared distances ascending, the point
 * itself first (reference: pointops2 knnquery, pointops2/src/knnquery/knnquery_cuda_kernel.cu:65-107, as called by
 * utils/general_utils.py:170-184 for the rigidity loss, train.py:132-152).  Uniform-grid search; `scratch` must hold
 * fdgs_knn_scratch_bytes(n) bytes.  brute_force != 0 runs the reference's O(n^2) scan instead (validation). */
size_t fdgs_knn_scratch_bytes(int n);
int fdgs_knn(int n, int k, const float* xyz, char* scratch, int* idx, float* dist2, int brute_force, void* stream);

/* Union bookkeeping of the exchange: cs = inclusive prefix sum of (radii > 0) over the P Gaussians (device, int32).
 * Writes slot_of[i] = cs[i] - 1 for union Gaussians, -1 otherwise, and the inverse map idx[slot] = i (idx: K =
 * cs[P-1] entries, int64). */
int fdgs_union_maps(long long P, const int* radii, const int* cs, int* slot_of, long long* idx, void* stream);

/* Test/diagnostic hooks (used by tests/ and bench.py only): copy private
 * per-Gaussian / per-instance state out of the scratch buffers into plain
 * caller-provided device arrays so that parity tests can compare them with the
 * oracle bit for bit.  Any output pointer may be NULL. */
int fdgs_debug_export_geom(const char* geom_buffer, int P,
                           float* depths, float* means2D /*[P,2]*/,
                           float* conic_opacity /*[P,4]*/, float* rgb /*[P,3]*/,
                           unsigned char* clamped /*[P,3]*/, unsigned int* tiles_touched,
                           void* stream);
int fdgs_debug_export_binning(const char* binning_buffer, const char* image_buffer,
                              int num_rendered, int width, int height,
                              unsigned int* point_list /*[R]*/,
                              unsigned int* ranges /*[tiles,2]*/,
                              unsigned int* n_contrib /*[H*W]*/,
                              void* stream);

/* Test hook of the raw-parameter entry: applies the in-kernel activations to plain arrays of n elements (log_s, logit:
 * [n]; quat: [n,4]; mode = quaternion-norm summation order, 0 = the library's default). */
int fdgs_debug_activate(int n, const float* log_s, const float* logit, const float* quat, int mode, float* s_out,
                        float* o_out, float* q_out, void* stream);

/* Tile lists.  The reference lists a Gaussian in every tile of the square of half-width ceil(3 sigma_max) around its
 * centre (auxiliary.h:46-59 getRect, rasterizer_impl.cu:71-112), although only the pixels where
 * alpha = min(0.99, opacity * exp(power)) reaches 1/255 ever blend it (forward.cu:590).
 *   mode 1 (default): a Gaussian is listed only in the tiles that ellipse can reach (exact conservative test per tile).
 *                     The instances dropped are instances no pixel would have blended: the images, radii, every other
 *                     output and every gradient are unchanged; the PRIVATE scratch state differs from the
 *                     reference's (shorter tile lists, smaller num_rendered, n_contrib counted in the shorter lists).
 *   mode 0:           the reference's tile lists exactly (point_list, ranges, num_rendered, n_contrib bit-identical).
 * Process-wide; returns the previous mode.  The environment variable FDGS_TILE_CULL sets the initial mode. */
int fdgs_set_tile_cull(int mode);

/* Measurement hooks (bench.py): per-stage device time with CUDA events recorded on the launch
 * stream, and a count of the kernels this library launched.  The reference has no equivalent
 * (it times whole iterations from Python, train.py:57-58,89,185). */
#define FDGS_NUM_STAGES 8
/* stage ids: 0 preprocess_fwd, 1 scan, 2 emit_keys, 3 sort, 4 pack_instances, 5 blend_fwd,
 *            6 blend_bwd, 7 preprocess_bwd */
int fdgs_profile_enable(int on);
/* Synchronises the recorded events, adds up the milliseconds and call counts per stage since the
 * last read, then resets the accumulators. */
int fdgs_profile_read(double ms[FDGS_NUM_STAGES], long long calls[FDGS_NUM_STAGES]);
/* number of CUDA kernels launched by this library since it was loaded */
long long fdgs_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* FDGS_H_INCLUDED */
