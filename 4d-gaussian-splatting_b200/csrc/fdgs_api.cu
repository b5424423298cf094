// fdgs_api.cu -- C-ABI entry points (include/fdgs.h) and scratch-buffer carving.
//
// Host-side orchestration of one forward / backward pass; replaces
// CudaRasterizer::Rasterizer::forward / backward / markVisible
// (reference: rasterizer_impl.cu:199-364, :368-496, :142-154) and the GeometryState /
// ImageState / BinningState chunk carving (rasterizer_impl.cu:156-195, rasterizer_impl.h:21-73).
#include <math.h>
#include <string.h>
#include <atomic>
#include <cstdlib>
#include <mutex>
#include <string>
#include <vector>
#include <nvtx3/nvToolsExt.h>   // header-only NVTX v3: ranges cost nothing unless a profiler is attached
#include "../../include/fdgs.h"
#include "fdgs_internal.h"

namespace {

thread_local std::string g_last_error;

int fail(int code, const std::string& msg) {
    g_last_error = msg;
    return code;
}

#define FDGS_CUDA(expr, what)                                                                         \
    do {                                                                                              \
        cudaError_t _e = (expr);                                                                      \
        if (_e != cudaSuccess)                                                                        \
            return fail(FDGS_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(_e));            \
    } while (0)

// ---- measurement hooks (include/fdgs.h: fdgs_profile_*, fdgs_launch_count) -----------------------
struct ProfEvent {
    int stage;
    cudaEvent_t a, b;
};
std::mutex g_prof_mu;
bool g_prof_on = false;
std::vector<ProfEvent> g_prof_events;
constexpr size_t kMaxProfEvents = 4096;
std::atomic<long long> g_kernel_launches{0};
// FDGS_TRACE=1 + debug=true: print every stage to stderr before and after its synchronisation
const bool g_trace = getenv("FDGS_TRACE") != nullptr;
// summation order of the quaternion norm in the raw-parameter entry (fdgs_common.cuh: quat_norm); 0 = ATen's order
const int g_quat_norm_mode = getenv("FDGS_NORMALIZE_MODE") ? atoi(getenv("FDGS_NORMALIZE_MODE")) : 0;

// tile lists: 1 = only the tiles a Gaussian's alpha >= 1/255 footprint reaches, 0 = the reference's 3-sigma square
// (include/fdgs.h: fdgs_set_tile_cull)
std::atomic<int> g_tile_cull{getenv("FDGS_TILE_CULL") ? atoi(getenv("FDGS_TILE_CULL")) : 1};

struct StageTimer {
    cudaEvent_t a = nullptr, b = nullptr;
    int stage;
    cudaStream_t stream;
    StageTimer(int stage_, cudaStream_t s, const char* name) : stage(stage_), stream(s) {
        nvtxRangePushA(name);   // one NVTX range per pipeline stage (nsys / ncu --nvtx)
        if (g_prof_on) {
            cudaEventCreate(&a);
            cudaEventCreate(&b);
            cudaEventRecord(a, stream);
        }
    }
    bool open = true;
    ~StageTimer() {
        if (open) nvtxRangePop();   // error return between start and stop
    }
    void stop(int kernels) {
        nvtxRangePop();
        open = false;
        g_kernel_launches += kernels;
        if (a) {
            cudaEventRecord(b, stream);
            std::lock_guard<std::mutex> lk(g_prof_mu);
            if (g_prof_events.size() < kMaxProfEvents) {
                g_prof_events.push_back({stage, a, b});
            } else {   // nobody is reading (fdgs_profile_read): do not grow without bound
                cudaEventDestroy(a);
                cudaEventDestroy(b);
            }
        }
    }
};

// FDGS_STAGE(stage id, kernels launched, launch expression, name)
#define FDGS_STAGE(sid, nk, expr, what)                                                               \
    do {                                                                                              \
        StageTimer _t(sid, stream, what);                                                                \
        FDGS_CUDA(expr, what);                                                                        \
        _t.stop(nk);                                                                                  \
        if (debug) {                                                                                  \
            if (g_trace) fprintf(stderr, "[fdgs] %s launched, synchronising\n", what);               \
            cudaError_t _s = cudaStreamSynchronize(stream);                                           \
            if (g_trace) fprintf(stderr, "[fdgs] %s done: %s\n", what, cudaGetErrorString(_s));      \
            if (_s != cudaSuccess)                                                                    \
                return fail(FDGS_ERR_CUDA, std::string(what) + " (debug sync): " + cudaGetErrorString(_s)); \
        }                                                                                             \
    } while (0)

// bump allocator over one chunk; every sub-array 128-byte aligned (like the reference's obtain())
struct Carver {
    char* base;
    size_t off;
    explicit Carver(char* b) : base(b), off(0) {}
    template <typename T>
    T* take(size_t count) {
        off = (off + 127) & ~(size_t)127;
        T* p = base ? reinterpret_cast<T*>(base + off) : nullptr;
        off += count * sizeof(T);
        return p;
    }
    size_t total() const { return off + 128; }
};

struct GeomState {
    float* cov3D;            // [P,6]  (first: exposed as covs3D_com)
    fdgs::InstRec* grec;     // [P]    per-Gaussian 64-byte record
    uint8_t* clamped;        // [P]
    uint32_t* tiles_touched; // [P]
    uint4* binrec;           // [P] tile rectangle + depth bits of every Gaussian, for the binning passes
    size_t bytes;
    static GeomState carve(char* base, int P) {
        GeomState g;
        Carver c(base);
        const size_t p = (size_t)(P > 0 ? P : 0);
        g.cov3D = c.take<float>(6 * p);   // first: 128-byte aligned [P,6] view for covs3D_com
        g.grec = c.take<fdgs::InstRec>(p);
        g.clamped = c.take<uint8_t>(p);
        g.tiles_touched = c.take<uint32_t>(p);
        g.binrec = c.take<uint4>(p);
        g.bytes = c.total();
        return g;
    }
};

struct ImageState {
    float* final_T;
    uint32_t* n_contrib;
    uint2* ranges;           // [tiles] start/end of every tile's instance list
    uint32_t* bin_matrix;    // [bin_ctas][tiles] per-CTA tile histograms -> column prefixes
    uint32_t* tile_total;    // [tiles]
    uint32_t* tile_offset;   // [tiles]
    uint32_t* bin_info;      // [4] total instances, largest tile, overflow flag, spare
    size_t tiles;
    size_t bytes;
    static ImageState carve(char* base, int W, int H) {
        ImageState s;
        Carver c(base);
        const size_t N = (size_t)W * H;
        s.tiles = (size_t)((W + fdgs::TILE_X - 1) / fdgs::TILE_X) * ((H + fdgs::TILE_Y - 1) / fdgs::TILE_Y);
        s.final_T = c.take<float>(N);
        s.n_contrib = c.take<uint32_t>(N);
        s.ranges = c.take<uint2>(s.tiles);
        s.bin_matrix = c.take<uint32_t>(s.tiles * (size_t)fdgs::bin_ctas());
        s.tile_total = c.take<uint32_t>(s.tiles);
        s.tile_offset = c.take<uint32_t>(s.tiles);
        s.bin_info = c.take<uint32_t>(4);
        s.bytes = c.total();
        return s;
    }
};

struct BinningState {
    fdgs::StageRec* recs;    // [R] instance records in tile-sorted order (80-byte stride, see fdgs_common.cuh)
    uint32_t* point_list;    // [R] sorted Gaussian indices (the reference's point_list)
    uint64_t* keys;          // [R] depth_bits<<32 | index, grouped by tile
    size_t bytes;
    static BinningState carve(char* base, int R) {
        BinningState b;
        Carver c(base);
        const size_t r = (size_t)(R > 0 ? R : 0);
        b.recs = c.take<fdgs::StageRec>(r);
        b.point_list = c.take<uint32_t>(r);
        b.keys = c.take<uint64_t>(r);
        b.bytes = c.total();
        return b;
    }
};

char* align128(char* p) {
    return reinterpret_cast<char*>((reinterpret_cast<uintptr_t>(p) + 127) & ~(uintptr_t)127);
}

// `whole_blocks`: the forward reads a row in whole 16-coefficient blocks (RowSmem::load_block, 12 float4 each), so
// its staged path needs M % 16 == 0; the backward bounds its accesses by the row length and only needs M % 4 == 0.
void sh_staging(const float* shs, int M, bool whole_blocks, int* bulk_ok, int* stride_floats) {
    *bulk_ok = 0;
    *stride_floats = 0;
    if (!shs || M <= 0) return;
    if ((M % (whole_blocks ? 16 : 4)) != 0 || (reinterpret_cast<uintptr_t>(shs) % 16) != 0) return;
    int q = (3 * M) / 4 + 1;   // 16-byte units incl. padding
    if ((q & 1) == 0) ++q;     // odd stride in 16-byte units -> conflict-free LDS.128 / STS.128
    if ((size_t)q * 16 * 128 > 200 * 1024) return;
    *bulk_ok = 1;
    *stride_floats = q * 4;
}

}  // namespace

extern "C" {

int fdgs_version(void) { return FDGS_VERSION; }

const char* fdgs_last_error(void) { return g_last_error.c_str(); }

size_t fdgs_geom_bytes(int P) { return GeomState::carve(nullptr, P).bytes + 128; }
size_t fdgs_image_bytes(int width, int height) { return ImageState::carve(nullptr, width, height).bytes + 128; }
size_t fdgs_binning_bytes(int num_rendered, int, int) { return BinningState::carve(nullptr, num_rendered).bytes + 128; }

int fdgs_forward(const fdgs_forward_args* a, fdgs_alloc_fn geom_alloc, void* geom_ctx, fdgs_alloc_fn binning_alloc,
                 void* binning_ctx, fdgs_alloc_fn image_alloc, void* image_ctx, void* stream_v,
                 fdgs_forward_result* res) {
    g_last_error.clear();
    if (!a || !res || !geom_alloc || !binning_alloc || !image_alloc) return fail(FDGS_ERR_INVALID_ARG, "null argument");
    memset(res, 0, sizeof(*res));
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_v);
    const bool debug = a->debug != 0;
    const int P = a->P, W = a->width, H = a->height;
    if (P < 0 || W <= 0 || H <= 0) return fail(FDGS_ERR_INVALID_ARG, "bad P / width / height");
    // the blend kernel writes every pixel of every image even when nothing is rendered (P == 0: background only)
    if (!a->out_color || !a->out_flow || !a->out_depth || !a->out_T || !a->background)
        return fail(FDGS_ERR_INVALID_ARG, "background and the output images must be set");
    if (P > 0) {
        if (!a->means3D || !a->opacities || !a->viewmatrix || !a->projmatrix || !a->cam_pos || !a->background)
            return fail(FDGS_ERR_INVALID_ARG, "means3D/opacities/viewmatrix/projmatrix/cam_pos/background must be set");
        if (!a->out_means3D || !a->out_color || !a->out_flow || !a->out_depth || !a->out_T || !a->radii)
            return fail(FDGS_ERR_INVALID_ARG, "output pointers must be set");
        if (!a->cov3D_precomp) {
            if (!a->scales || !a->rotations) return fail(FDGS_ERR_INVALID_ARG, "scales/rotations or cov3D_precomp required");
            if (a->rot_4d && (!a->rotations_r || !a->scales_t || !a->ts))
                return fail(FDGS_ERR_INVALID_ARG, "rot_4d needs rotations_r, scales_t and ts");
            if (!a->rot_4d && a->gaussian_dim == 4 && (!a->scales_t || !a->ts))
                return fail(FDGS_ERR_INVALID_ARG, "gaussian_dim 4 needs scales_t and ts");
        }
        if (!a->colors_precomp) {
            // reference: rasterizer_impl.cu:255-258 (NUM_CHANNELS is 3 here, so SHs are required)
            if (!a->shs) return fail(FDGS_ERR_INVALID_ARG, "either shs or colors_precomp required");
            const bool sh3d = a->gaussian_dim == 3 || a->force_sh_3d;
            int need = (a->D + 1) * (a->D + 1);
            if (!sh3d && a->D > 2 && a->D_t > 0) need = 16 * (a->D_t + 1);
            if (a->D < 0 || a->D > 3 || a->D_t < 0 || a->D_t > 2 || a->M < need)
                return fail(FDGS_ERR_INVALID_ARG, "SH degree / coefficient count mismatch");
            if (!sh3d && !a->ts) return fail(FDGS_ERR_INVALID_ARG, "4D SH needs ts");
        }
    }
    const int grid_x = (W + fdgs::TILE_X - 1) / fdgs::TILE_X, grid_y = (H + fdgs::TILE_Y - 1) / fdgs::TILE_Y;
    // tile coordinates travel as 16-bit fields of the per-Gaussian bin record (preprocess_fwd.cu / binning.cu)
    if (grid_x > 65535 || grid_y > 65535) return fail(FDGS_ERR_UNSUPPORTED, "image larger than 1048560 pixels on a side");
    const size_t N = (size_t)W * H;

    // scratch: geometry + image
    const size_t geom_bytes = fdgs_geom_bytes(P);
    char* geom_raw = geom_alloc(geom_ctx, geom_bytes);
    if (!geom_raw) return fail(FDGS_ERR_ALLOC, "geometry buffer allocation failed");
    GeomState geom = GeomState::carve(align128(geom_raw), P);
    const size_t img_bytes = fdgs_image_bytes(W, H);
    char* img_raw = image_alloc(image_ctx, img_bytes);
    if (!img_raw) return fail(FDGS_ERR_ALLOC, "image buffer allocation failed");
    ImageState img = ImageState::carve(align128(img_raw), W, H);
    res->geom_buffer = geom_raw;
    res->geom_bytes = geom_bytes;
    res->image_buffer = img_raw;
    res->image_bytes = img_bytes;
    res->cov3D = geom.cov3D;

    int num_rendered = 0;
    if (P > 0) {
        fdgs::PreprocessFwdParams pp;
        memset(&pp, 0, sizeof(pp));
        pp.P = P; pp.D = a->D; pp.D_t = a->D_t; pp.M = a->M;
        pp.means3D = a->means3D; pp.ts = a->ts; pp.scales = a->scales; pp.scales_t = a->scales_t;
        pp.scale_modifier = a->scale_modifier; pp.rotations = a->rotations; pp.rotations_r = a->rotations_r;
        pp.opacities = a->opacities; pp.shs = a->shs; pp.cov3D_precomp = a->cov3D_precomp;
        pp.prefilter_var = a->prefilter_var; pp.colors_precomp = a->colors_precomp;
        pp.viewmatrix = a->viewmatrix; pp.projmatrix = a->projmatrix; pp.cam_pos = a->cam_pos;
        pp.timestamp = a->timestamp; pp.time_duration = a->time_duration;
        pp.rot_4d = a->rot_4d; pp.gaussian_dim = a->gaussian_dim; pp.force_sh_3d = a->force_sh_3d;
        pp.W = W; pp.H = H; pp.tan_fovx = a->tan_fovx; pp.tan_fovy = a->tan_fovy;
        // reference: rasterizer_impl.cu:235-236
        pp.focal_y = H / (2.0f * a->tan_fovy);
        pp.focal_x = W / (2.0f * a->tan_fovx);
        pp.grid_x = grid_x; pp.grid_y = grid_y; pp.prefiltered = a->prefiltered;
        sh_staging(a->colors_precomp ? nullptr : a->shs, a->M, true, &pp.sh_bulk_ok, &pp.sh_row_stride_floats);
        pp.raw_params = a->raw_params; pp.quat_norm_mode = g_quat_norm_mode;
        pp.shs_rest = a->colors_precomp ? nullptr : a->shs_rest;
        if (pp.shs_rest && a->M < 2) return fail(FDGS_ERR_INVALID_ARG, "split SH rows need M >= 2");
        if (a->raw_params && a->cov3D_precomp) return fail(FDGS_ERR_INVALID_ARG, "raw_params needs scales / rotations, not cov3D_precomp");
        pp.flows = a->flows_precomp;
        pp.out_means3D = a->out_means3D; pp.radii = a->radii; pp.cov3D = geom.cov3D; pp.grec = geom.grec;
        pp.clamped = geom.clamped; pp.tiles_touched = geom.tiles_touched; pp.binrec = geom.binrec;
        pp.tile_cull = g_tile_cull.load();
        FDGS_STAGE(0, 1, fdgs::launch_preprocess_fwd(pp, stream), "preprocess_fwd");
    }
    // binning: per-tile counts, offsets, ranges and the total instance count R
    FDGS_STAGE(1, (P > 0 ? 1 : 0) + 2,
               [&]() {
                   cudaError_t e = fdgs::launch_bin_count(P, geom.binrec, grid_x, grid_y, img.bin_matrix, stream);
                   if (e != cudaSuccess) return e;
                   return fdgs::launch_tile_scan((int)img.tiles, img.bin_matrix, img.tile_total, img.tile_offset, img.ranges,
                                                 img.bin_info, stream);
               }(),
               "bin_count + tile_scan");
    // the one host synchronisation of the forward (reference: rasterizer_impl.cu:302)
    int bin_info[4] = {0, 0, 0, 0};
    FDGS_CUDA(cudaMemcpyAsync(bin_info, img.bin_info, sizeof(bin_info), cudaMemcpyDeviceToHost, stream), "num_rendered copy");
    FDGS_CUDA(cudaStreamSynchronize(stream), "num_rendered sync");
    num_rendered = bin_info[0];
    // bin_info[2]: the 64-bit total computed on the device exceeded 2^31 - 1 (a wrapped 32-bit total would look valid)
    if (num_rendered < 0 || bin_info[2] != 0) return fail(FDGS_ERR_UNSUPPORTED, "more than 2^31 - 1 tile instances");
    res->num_rendered = num_rendered;

    const size_t bin_bytes = fdgs_binning_bytes(num_rendered, W, H);
    char* bin_raw = binning_alloc(binning_ctx, bin_bytes);
    if (!bin_raw) return fail(FDGS_ERR_ALLOC, "binning buffer allocation failed");
    BinningState bin = BinningState::carve(align128(bin_raw), num_rendered);
    res->binning_buffer = bin_raw;
    res->binning_bytes = bin_bytes;

    if (num_rendered > 0) {
        FDGS_STAGE(2, 1, fdgs::launch_bin_scatter(P, geom.binrec, grid_x, grid_y, img.bin_matrix, img.tile_offset, bin.keys, stream),
                   "bin_scatter");
        FDGS_STAGE(3, fdgs::tile_sort_pack_kernel_count(bin_info[1]),
                   fdgs::launch_tile_sort_pack((int)img.tiles, bin_info[1], num_rendered, img.ranges, bin.keys, geom.grec, bin.recs,
                                               bin.point_list, stream),
                   "tile_sort_pack");
    }
    fdgs::BlendFwdParams bp;
    bp.W = W; bp.H = H; bp.grid_x = grid_x; bp.grid_y = grid_y;
    bp.ranges = img.ranges; bp.recs = bin.recs; bp.background = a->background;
    bp.final_T = img.final_T; bp.n_contrib = img.n_contrib;
    bp.out_color = a->out_color; bp.out_flow = a->out_flow; bp.out_depth = a->out_depth; bp.out_T = a->out_T;
    (void)N;
    FDGS_STAGE(5, 1, fdgs::launch_blend_fwd(bp, stream), "blend_fwd");
    return FDGS_OK;
}

int fdgs_backward(const fdgs_backward_args* a, void* stream_v) {
    g_last_error.clear();
    if (!a) return fail(FDGS_ERR_INVALID_ARG, "null argument");
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_v);
    const bool debug = a->debug != 0;
    const int P = a->P, W = a->width, H = a->height, R = a->R;
    if (P <= 0) return FDGS_OK;
    if (W <= 0 || H <= 0 || R < 0) return fail(FDGS_ERR_INVALID_ARG, "bad width / height / R");
    if (!a->geom_buffer || !a->image_buffer || (R > 0 && !a->binning_buffer))
        return fail(FDGS_ERR_INVALID_ARG, "scratch buffers of the forward pass required");
    // dL_depths / dL_masks / dL_dpix_flow may be NULL: no upstream gradient for that image
    if (!a->dL_dpix) return fail(FDGS_ERR_INVALID_ARG, "colour image gradient required");
    if (!a->background || !a->radii || !a->out_means3D || !a->viewmatrix || !a->projmatrix || !a->campos)
        return fail(FDGS_ERR_INVALID_ARG, "background / radii / out_means3D / viewmatrix / projmatrix / campos required");
    if (a->sh_factors && a->dL_dsh) return fail(FDGS_ERR_INVALID_ARG, "sh_factors and dL_dsh are mutually exclusive");
    if (!a->dL_dmean2D || !a->dL_dconic || !a->dL_dopacity || !a->dL_dcolor || !a->dL_dflows || !a->dL_dmean3D ||
        !a->dL_dcov3D || !a->dL_dts || !a->dL_dscale || !a->dL_dscale_t || !a->dL_drot || !a->dL_drot_r)
        return fail(FDGS_ERR_INVALID_ARG, "gradient outputs required");
    const int grid_x = (W + fdgs::TILE_X - 1) / fdgs::TILE_X, grid_y = (H + fdgs::TILE_Y - 1) / fdgs::TILE_Y;
    GeomState geom = GeomState::carve(align128(const_cast<char*>(a->geom_buffer)), P);
    ImageState img = ImageState::carve(align128(const_cast<char*>(a->image_buffer)), W, H);
    BinningState bin = BinningState::carve(a->binning_buffer ? align128(const_cast<char*>(a->binning_buffer)) : nullptr, R);

    bool blend_raw = false;
    if (R > 0) {
        fdgs::BlendBwdParams bp;
        bp.W = W; bp.H = H; bp.grid_x = grid_x; bp.grid_y = grid_y;
        bp.ranges = img.ranges; bp.recs = bin.recs; bp.background = a->background;
        bp.final_T = img.final_T; bp.n_contrib = img.n_contrib;
        bp.dL_dpix = a->dL_dpix; bp.dL_depths = a->dL_depths; bp.dL_masks = a->dL_masks; bp.dL_dpix_flow = a->dL_dpix_flow;
        bp.dL_dmean2D = a->dL_dmean2D; bp.dL_dconic = a->dL_dconic; bp.dL_dopacity = a->dL_dopacity;
        bp.dL_dcolor = a->dL_dcolor; bp.dL_dflows = a->dL_dflows;
        blend_raw = fdgs::blend_bwd_is_raw(bp);
        FDGS_STAGE(6, 1, fdgs::launch_blend_bwd(bp, stream), "blend_bwd");
    }
    fdgs::PreprocessBwdParams pb;
    memset(&pb, 0, sizeof(pb));
    pb.P = P; pb.D = a->D; pb.D_t = a->D_t; pb.M = a->M;
    pb.means3D = a->out_means3D; pb.radii = a->radii; pb.shs = a->shs; pb.ts = a->ts; pb.opacities = a->opacities;
    pb.clamped = geom.clamped; pb.tiles_touched = geom.tiles_touched;
    pb.scales = a->scales; pb.scales_t = a->scales_t; pb.rotations = a->rotations; pb.rotations_r = a->rotations_r;
    pb.scale_modifier = a->scale_modifier;
    pb.cov3D = a->cov3D_precomp ? a->cov3D_precomp : geom.cov3D;
    pb.prefilter_var = a->prefilter_var; pb.viewmatrix = a->viewmatrix; pb.projmatrix = a->projmatrix;
    pb.focal_y = H / (2.0f * a->tan_fovy);
    pb.focal_x = W / (2.0f * a->tan_fovx);
    pb.tan_fovx = a->tan_fovx; pb.tan_fovy = a->tan_fovy; pb.campos = a->campos;
    pb.timestamp = a->timestamp; pb.time_duration = a->time_duration;
    pb.rot_4d = a->rot_4d; pb.gaussian_dim = a->gaussian_dim; pb.force_sh_3d = a->force_sh_3d;
    pb.has_scales = (a->scales != nullptr) ? 1 : 0;
    pb.grec = geom.grec; pb.blend_raw = blend_raw ? 1 : 0; pb.W = W; pb.H = H;
    sh_staging(a->shs, a->M, false, &pb.sh_bulk_ok, &pb.sh_row_stride_floats);
    if (a->dL_dsh && (reinterpret_cast<uintptr_t>(a->dL_dsh) % 16) != 0) pb.sh_bulk_ok = 0;
    pb.raw_params = a->raw_params; pb.quat_norm_mode = g_quat_norm_mode;
    pb.shs_rest = a->shs ? a->shs_rest : nullptr; pb.dL_dsh_rest = a->dL_dsh_rest;
    if (pb.shs_rest && a->dL_dsh && !a->dL_dsh_rest) return fail(FDGS_ERR_INVALID_ARG, "split SH rows need dL_dsh_rest");
    if (!pb.shs_rest && a->dL_dsh_rest) return fail(FDGS_ERR_INVALID_ARG, "dL_dsh_rest without shs_rest");
    pb.dL_dmean2D = a->dL_dmean2D; pb.dL_dconic = a->dL_dconic; pb.dL_dopacity = a->dL_dopacity; pb.dL_dcolor = a->dL_dcolor;
    pb.dL_dmean3D = a->dL_dmean3D; pb.dL_dcov3D = a->dL_dcov3D; pb.dL_dsh = a->dL_dsh; pb.sh_factors = a->sh_factors;
    pb.dL_dts = a->dL_dts;
    pb.dL_dscale = a->dL_dscale; pb.dL_dscale_t = a->dL_dscale_t; pb.dL_drot = a->dL_drot; pb.dL_drot_r = a->dL_drot_r;
    if (pb.has_scales && pb.rot_4d && (!a->rotations_r || !a->scales_t || !a->ts || !a->opacities))
        return fail(FDGS_ERR_INVALID_ARG, "rot_4d backward needs rotations_r, scales_t, ts, opacities");
    FDGS_STAGE(7, fdgs::preprocess_bwd_kernel_count(pb), fdgs::launch_preprocess_bwd(pb, stream), "preprocess_bwd");
    return FDGS_OK;
}

int fdgs_mark_visible(int P, const float* means3D, const float* viewmatrix, const float*, unsigned char* present,
                      void* stream_v) {
    g_last_error.clear();
    if (P < 0 || (P > 0 && (!means3D || !viewmatrix || !present))) return fail(FDGS_ERR_INVALID_ARG, "null argument");
    FDGS_CUDA(fdgs::launch_mark_visible(P, means3D, viewmatrix, present, reinterpret_cast<cudaStream_t>(stream_v)),
              "mark_visible");
    return FDGS_OK;
}

static int pack_common(bool unpack, int n, float* const* tensors, const int* widths, const long long* block_off,
                       const long long* idx, long long K, float* flat, void* stream_v) {
    g_last_error.clear();
    if (n < 0 || n > FDGS_MAX_PACK || K < 0) return fail(FDGS_ERR_INVALID_ARG, "bad tensor count / row count");
    if (n == 0 || K == 0) return FDGS_OK;
    if (!tensors || !widths || !block_off || !idx || !flat) return fail(FDGS_ERR_INVALID_ARG, "null argument");
    for (int i = 0; i < n; ++i)
        if (!tensors[i] || widths[i] <= 0 || block_off[i] < 0) return fail(FDGS_ERR_INVALID_ARG, "bad tensor table entry");
    FDGS_CUDA(fdgs::launch_pack_rows(unpack, n, tensors, widths, block_off, idx, K, flat, reinterpret_cast<cudaStream_t>(stream_v)),
              unpack ? "unpack_rows" : "pack_rows");
    g_kernel_launches += 1;
    return FDGS_OK;
}

int fdgs_pack_rows(int n, const float* const* tensors, const int* widths, const long long* block_off, const long long* idx,
                   long long K, float* flat, void* stream) {
    return pack_common(false, n, const_cast<float* const*>(tensors), widths, block_off, idx, K, flat, stream);
}

int fdgs_unpack_rows(int n, float* const* tensors, const int* widths, const long long* block_off, const long long* idx,
                     long long K, const float* flat, void* stream) {
    return pack_common(true, n, tensors, widths, block_off, idx, K, const_cast<float*>(flat), stream);
}

int fdgs_sh_outer_sum(const fdgs_sh_sum_args* a, void* stream_v) {
    g_last_error.clear();
    if (!a) return fail(FDGS_ERR_INVALID_ARG, "null argument");
    if (a->P <= 0) return FDGS_OK;
    if (a->V < 0 || a->K < 0 || a->M <= 0 || a->m0 <= 0 || a->m0 > a->M || a->D < 0 || a->D > 3 || a->D_t < 0 || a->D_t > 2)
        return fail(FDGS_ERR_INVALID_ARG, "bad V / K / M / m0 / degrees");
    if (!a->slot_of || !a->means3D || !a->out0 || (a->m0 < a->M && !a->out1) ||
        (a->V > 0 && a->K > 0 && (!a->table || !a->union_idx || !a->dir_scratch)))
        return fail(FDGS_ERR_INVALID_ARG, "slot_of / union_idx / dir_scratch / means3D / table / outputs required");
    if (a->rot_4d && (!a->scales || !a->scales_t || !a->rotations || !a->rotations_r || !a->ts))
        return fail(FDGS_ERR_INVALID_ARG, "rot_4d needs scales, scales_t, rotations, rotations_r, ts");
    const bool sh4d = !(a->gaussian_dim == 3 || a->force_sh_3d);
    if (sh4d && !a->ts) return fail(FDGS_ERR_INVALID_ARG, "4D SH needs ts");
    if (a->V > 0 && (a->meta_off < 3ll * a->K || a->view_stride < a->meta_off + 4))
        return fail(FDGS_ERR_INVALID_ARG, "view block too small for K rows + metadata");
    fdgs::ShSumParams p;
    p.P = a->P; p.V = a->V; p.K = a->K; p.table = a->table; p.view_stride = a->view_stride; p.meta_off = a->meta_off;
    p.slot_of = a->slot_of; p.union_idx = a->union_idx; p.dirs = a->dir_scratch; p.means3D = a->means3D; p.ts = a->ts; p.scales = a->scales; p.scales_t = a->scales_t;
    p.rotations = a->rotations; p.rotations_r = a->rotations_r; p.scale_modifier = a->scale_modifier;
    p.time_duration = a->time_duration; p.rot_4d = a->rot_4d; p.gaussian_dim = a->gaussian_dim;
    p.force_sh_3d = a->force_sh_3d; p.D = a->D; p.D_t = a->D_t; p.M = a->M;
    p.out0 = a->out0; p.m0 = a->m0; p.out1 = (a->m0 < a->M) ? a->out1 : nullptr; p.accumulate = a->accumulate;
    FDGS_CUDA(fdgs::launch_sh_outer_sum(p, reinterpret_cast<cudaStream_t>(stream_v)), "sh_outer_sum");
    g_kernel_launches += 1;
    return FDGS_OK;
}

int fdgs_union_maps(long long P, const int* radii, const int* cs, int* slot_of, long long* idx, void* stream_v) {
    g_last_error.clear();
    if (P < 0) return fail(FDGS_ERR_INVALID_ARG, "bad P");
    if (P == 0) return FDGS_OK;
    if (!radii || !cs || !slot_of || !idx) return fail(FDGS_ERR_INVALID_ARG, "null argument");
    FDGS_CUDA(fdgs::launch_union_maps(P, radii, cs, slot_of, idx, reinterpret_cast<cudaStream_t>(stream_v)), "union_maps");
    g_kernel_launches += 1;
    return FDGS_OK;
}

int fdgs_view_stats(long long P, const float* viewspace_grad, int grad_stride, const int* radii, float* grad_norm_sum,
                    float* visibility_count, int* max_radii, void* stream_v) {
    g_last_error.clear();
    if (P < 0 || grad_stride < 2) return fail(FDGS_ERR_INVALID_ARG, "bad P / gradient row stride");
    if (P == 0) return FDGS_OK;
    if (!viewspace_grad || !radii || !grad_norm_sum || !visibility_count || !max_radii) return fail(FDGS_ERR_INVALID_ARG, "null argument");
    FDGS_CUDA(fdgs::launch_view_stats(P, viewspace_grad, grad_stride, radii, grad_norm_sum, visibility_count, max_radii,
                                      reinterpret_cast<cudaStream_t>(stream_v)), "view_stats");
    g_kernel_launches += 1;
    return FDGS_OK;
}

int fdgs_check_rows_zero(int n, const float* const* tensors, const int* widths, long long P, const int* radii, int* flag,
                         void* stream_v) {
    g_last_error.clear();
    if (n < 0 || n > FDGS_MAX_PACK || P < 0) return fail(FDGS_ERR_INVALID_ARG, "bad tensor count / row count");
    if (n == 0 || P == 0) return FDGS_OK;
    if (!tensors || !widths || !radii || !flag) return fail(FDGS_ERR_INVALID_ARG, "null argument");
    for (int i = 0; i < n; ++i)
        if (!tensors[i] || widths[i] <= 0) return fail(FDGS_ERR_INVALID_ARG, "bad tensor table entry");
    FDGS_CUDA(fdgs::launch_rows_zero_check(n, tensors, widths, P, radii, flag, reinterpret_cast<cudaStream_t>(stream_v)),
              "rows_zero_check");
    g_kernel_launches += 1;
    return FDGS_OK;
}

int fdgs_l1_ssim_forward(const float* x, const float* y, int C, int H, int W, float* maps, double* sums, void* stream_v) {
    g_last_error.clear();
    if (C <= 0 || H <= 0 || W <= 0) return fail(FDGS_ERR_INVALID_ARG, "bad C / H / W");
    if (!x || !y || !maps || !sums) return fail(FDGS_ERR_INVALID_ARG, "null argument");
    FDGS_CUDA(fdgs::launch_l1_ssim_fwd(x, y, C, H, W, maps, sums, reinterpret_cast<cudaStream_t>(stream_v)), "l1_ssim_forward");
    g_kernel_launches += 1;
    return FDGS_OK;
}

int fdgs_l1_ssim_backward(const float* x, const float* y, int C, int H, int W, const float* maps, const float* grad_scale,
                          float lambda_dssim, float* dL_dx, void* stream_v) {
    g_last_error.clear();
    if (C <= 0 || H <= 0 || W <= 0) return fail(FDGS_ERR_INVALID_ARG, "bad C / H / W");
    if (!x || !y || !maps || !dL_dx) return fail(FDGS_ERR_INVALID_ARG, "null argument");
    FDGS_CUDA(fdgs::launch_l1_ssim_bwd(x, y, C, H, W, maps, grad_scale, lambda_dssim, dL_dx, reinterpret_cast<cudaStream_t>(stream_v)),
              "l1_ssim_backward");
    g_kernel_launches += 1;
    return FDGS_OK;
}

int fdgs_adam_step(int n, float* const* params, float* const* grads, float* const* exp_avg, float* const* exp_avg_sq,
                   const int* widths, const float* lrs, long long P, const long long* rows, long long num_rows, long long step,
                   double beta1, double beta2, double eps, int zero_grad, void* stream_v) {
    g_last_error.clear();
    if (n < 0 || n > FDGS_MAX_PACK || P < 0 || step < 1 || (rows && num_rows < 0))
        return fail(FDGS_ERR_INVALID_ARG, "bad tensor count / row count / step");
    if (n == 0 || P == 0 || (rows && num_rows == 0)) return FDGS_OK;
    if (!params || !grads || !exp_avg || !exp_avg_sq || !widths || !lrs) return fail(FDGS_ERR_INVALID_ARG, "null argument");
    float step_sizes[FDGS_MAX_PACK];
    // bias corrections in double on the host, like torch.optim.Adam's single-tensor path
    const double bc1 = 1.0 - pow(beta1, (double)step), bc2 = 1.0 - pow(beta2, (double)step);
    for (int i = 0; i < n; ++i) {
        if (!params[i] || !grads[i] || !exp_avg[i] || !exp_avg_sq[i] || widths[i] <= 0)
            return fail(FDGS_ERR_INVALID_ARG, "bad tensor table entry");
        step_sizes[i] = (float)((double)lrs[i] / bc1);
    }
    FDGS_CUDA(fdgs::launch_adam(n, params, grads, exp_avg, exp_avg_sq, widths, step_sizes, rows ? num_rows : P, rows, (float)(1.0 - beta1),
                                (float)beta2, (float)(1.0 - beta2), (float)eps, (float)sqrt(bc2), zero_grad, reinterpret_cast<cudaStream_t>(stream_v)),
              "adam_step");
    g_kernel_launches += 1;
    return FDGS_OK;
}

size_t fdgs_knn_scratch_bytes(int n) { return fdgs::knn_scratch_bytes(n > 0 ? n : 0); }

int fdgs_knn(int n, int k, const float* xyz, char* scratch, int* idx, float* dist2, int brute_force, void* stream_v) {
    g_last_error.clear();
    if (n < 0 || k < 1 || k > 32) return fail(FDGS_ERR_INVALID_ARG, "bad n / k (1 <= k <= 32)");
    if (n == 0) return FDGS_OK;
    if (!xyz || !idx || !dist2 || (!brute_force && !scratch)) return fail(FDGS_ERR_INVALID_ARG, "null argument");
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_v);
    if (brute_force) FDGS_CUDA(fdgs::launch_knn_bruteforce(n, k, xyz, idx, dist2, stream), "knn (brute force)");
    else FDGS_CUDA(fdgs::launch_knn(n, k, xyz, scratch, idx, dist2, stream), "knn");
    g_kernel_launches += brute_force ? 1 : 7;
    return FDGS_OK;
}

int fdgs_debug_activate(int n, const float* log_s, const float* logit, const float* quat, int mode, float* s_out, float* o_out,
                        float* q_out, void* stream_v) {
    g_last_error.clear();
    if (n < 0 || (n > 0 && (!log_s || !logit || !quat || !s_out || !o_out || !q_out))) return fail(FDGS_ERR_INVALID_ARG, "null argument");
    FDGS_CUDA(fdgs::launch_debug_activate(n, log_s, logit, quat, mode, s_out, o_out, q_out, reinterpret_cast<cudaStream_t>(stream_v)),
              "debug_activate");
    return FDGS_OK;
}

int fdgs_profile_enable(int on) {
    std::lock_guard<std::mutex> lk(g_prof_mu);
    g_prof_on = on != 0;
    return FDGS_OK;
}

int fdgs_profile_read(double ms[FDGS_NUM_STAGES], long long calls[FDGS_NUM_STAGES]) {
    std::vector<ProfEvent> ev;
    {
        std::lock_guard<std::mutex> lk(g_prof_mu);
        ev.swap(g_prof_events);
    }
    for (int i = 0; i < FDGS_NUM_STAGES; ++i) {
        if (ms) ms[i] = 0.0;
        if (calls) calls[i] = 0;
    }
    for (auto& e : ev) {
        cudaEventSynchronize(e.b);
        float t = 0.f;
        cudaEventElapsedTime(&t, e.a, e.b);
        if (e.stage >= 0 && e.stage < FDGS_NUM_STAGES) {
            if (ms) ms[e.stage] += t;
            if (calls) calls[e.stage] += 1;
        }
        cudaEventDestroy(e.a);
        cudaEventDestroy(e.b);
    }
    return FDGS_OK;
}

long long fdgs_launch_count(void) { return g_kernel_launches.load(); }

int fdgs_set_tile_cull(int mode) { return g_tile_cull.exchange(mode ? 1 : 0); }

int fdgs_debug_export_geom(const char* geom_buffer, int P, float* depths, float* means2D, float* conic_opacity,
                           float* rgb, unsigned char* clamped, unsigned int* tiles_touched, void* stream_v) {
    g_last_error.clear();
    if (!geom_buffer || P <= 0) return fail(FDGS_ERR_INVALID_ARG, "null argument");
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_v);
    GeomState g = GeomState::carve(align128(const_cast<char*>(geom_buffer)), P);
    const size_t p = (size_t)P;
    // rows the forward did not render read as zeros (their records are never written)
    FDGS_CUDA(fdgs::launch_unpack_grec(P, g.grec, reinterpret_cast<const int*>(g.tiles_touched), depths, means2D,
                                       conic_opacity, rgb, stream),
              "export records");
    if (clamped) FDGS_CUDA(cudaMemcpyAsync(clamped, g.clamped, p, cudaMemcpyDeviceToDevice, stream), "export clamped");
    if (tiles_touched)
        FDGS_CUDA(cudaMemcpyAsync(tiles_touched, g.tiles_touched, p * 4, cudaMemcpyDeviceToDevice, stream), "export tiles");
    return FDGS_OK;
}

int fdgs_debug_export_binning(const char* binning_buffer, const char* image_buffer, int R, int W, int H,
                              unsigned int* point_list, unsigned int* ranges, unsigned int* n_contrib, void* stream_v) {
    g_last_error.clear();
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_v);
    if (!image_buffer) return fail(FDGS_ERR_INVALID_ARG, "null argument");
    ImageState img = ImageState::carve(align128(const_cast<char*>(image_buffer)), W, H);
    const size_t tiles = (size_t)((W + fdgs::TILE_X - 1) / fdgs::TILE_X) * ((H + fdgs::TILE_Y - 1) / fdgs::TILE_Y);
    if (point_list && R > 0) {
        if (!binning_buffer) return fail(FDGS_ERR_INVALID_ARG, "null binning buffer");
        BinningState b = BinningState::carve(align128(const_cast<char*>(binning_buffer)), R);
        FDGS_CUDA(cudaMemcpyAsync(point_list, b.point_list, (size_t)R * 4, cudaMemcpyDeviceToDevice, stream),
                  "export point_list");
    }
    if (ranges) FDGS_CUDA(cudaMemcpyAsync(ranges, img.ranges, tiles * 8, cudaMemcpyDeviceToDevice, stream), "export ranges");
    if (n_contrib)
        FDGS_CUDA(cudaMemcpyAsync(n_contrib, img.n_contrib, (size_t)W * H * 4, cudaMemcpyDeviceToDevice, stream),
                  "export n_contrib");
    return FDGS_OK;
}

}  // extern "C"
