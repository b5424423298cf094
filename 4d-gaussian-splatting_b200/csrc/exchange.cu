// exchange.cu -- pack / unpack of per-Gaussian gradient rows for the multi-GPU exchange.
//
// The reference has no multi-GPU path; its batch loop sums the per-Gaussian gradients of sequential views
// (train.py:104-166).  Here views are spread over ranks (fdgs/dist.py) and the sum is one NCCL all-reduce.
// A Gaussian rendered by no view of the step has an all-zero gradient row on every rank, so only the rows
// of the union of rendered Gaussians are exchanged: pack_rows gathers, for a sorted index list, the rows of up
// to FDGS_MAX_PACK tensors into one flat buffer (one block of K rows per tensor, block starts 16-byte
// aligned), the buffer is all-reduced with a single collective, unpack_rows scatters the sums back.
// Pure HBM streaming: 16-byte vector accesses where a row is a multiple of 16 bytes (the SH rows, 89 % of
// the bytes), scalar otherwise; one launch covers all tensors (blockIdx.y = tensor).
#include "../../include/fdgs.h"
#include "fdgs_internal.h"

namespace fdgs {
namespace {

struct PackTable {
    float* tensor[FDGS_MAX_PACK];       // [P, width] row-major
    long long block_off[FDGS_MAX_PACK];  // float offset of the tensor's block inside the flat buffer
    int width[FDGS_MAX_PACK];
    int vec4[FDGS_MAX_PACK];            // rows and block are 16-byte aligned multiples
    int n;
};

template <bool UNPACK>
__global__ void __launch_bounds__(256) pack_rows_kernel(const PackTable tb, const long long* __restrict__ idx, long long K,
                                                        float* __restrict__ flat) {
    const int t = blockIdx.y;
    const int w = tb.width[t];
    float* __restrict__ ten = tb.tensor[t];
    float* __restrict__ blk = flat + tb.block_off[t];
    const long long stride = (long long)gridDim.x * blockDim.x;
    if (tb.vec4[t]) {
        const int q = w >> 2;   // float4 per row
        const long long total = K * q;
        float4* t4 = reinterpret_cast<float4*>(ten);
        float4* b4 = reinterpret_cast<float4*>(blk);
        for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
            const long long r = i / q;
            const int c = (int)(i - r * q);
            const long long g = idx[r] * q + c;
            if (UNPACK) t4[g] = b4[i];
            else b4[i] = t4[g];
        }
    } else {
        const long long total = K * w;
        for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
            const long long r = i / w;
            const int c = (int)(i - r * w);
            const long long g = idx[r] * w + c;
            if (UNPACK) ten[g] = blk[i];
            else blk[i] = ten[g];
        }
    }
}

// Guard of the sparse exchange: rows of Gaussians outside the union are assumed to be zero on every rank (true for
// the rasterizer's gradients, false for e.g. a rigidity loss over all Gaussians, reference train.py:132-152).  One
// pass over the (small) geometry gradients; the caller reads the flag together with the union size.
struct CheckTable {
    const float* tensor[FDGS_MAX_PACK];
    int width[FDGS_MAX_PACK];
};
__global__ void __launch_bounds__(256) rows_zero_check_kernel(const CheckTable tb, long long P, const int* __restrict__ radii,
                                                              int* __restrict__ flag) {
    const int t = blockIdx.y;
    const int w = tb.width[t];
    const float* __restrict__ ten = tb.tensor[t];
    bool bad = false;
    for (long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x; r < P; r += (long long)gridDim.x * blockDim.x) {
        if (radii[r] > 0) continue;
        for (int c = 0; c < w; ++c) bad |= (ten[r * w + c] != 0.f);
    }
    if (__any_sync(0xffffffffu, bad) && (threadIdx.x & 31) == 0) atomicOr(flag, 1);
}

// union bookkeeping of the exchange in one pass: given the inclusive scan `cs` of (radii > 0),
// slot_of[i] = cs[i] - 1 for union Gaussians (else -1) and idx[slot] = i, the inverse map
__global__ void __launch_bounds__(256) union_maps_kernel(long long P, const int* __restrict__ radii, const int* __restrict__ cs,
                                                         int* __restrict__ slot_of, long long* __restrict__ idx) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P) return;
    if (radii[i] > 0) {
        const int sl = cs[i] - 1;
        slot_of[i] = sl;
        idx[sl] = i;
    } else {
        slot_of[i] = -1;
    }
}

}  // namespace

// Per-view densification statistics in one pass (reference: train.py:164-183, scene/gaussian_model.py:579-589):
//   grad_norm_sum[i] += ||viewspace_grad[i, :2]||,  visibility_count[i] += (radii[i] > 0),  max_radii[i] = max(., radii[i])
static __global__ void __launch_bounds__(256) view_stats_kernel(long long P, const float* __restrict__ vgrad, int vstride,
                                                         const int* __restrict__ radii, float* __restrict__ grad_norm_sum,
                                                         float* __restrict__ visibility_count, int* __restrict__ max_radii) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P) return;
    const float gx = vgrad[i * vstride + 0], gy = vgrad[i * vstride + 1];
    const int r = radii[i];
    grad_norm_sum[i] += sqrtf(gx * gx + gy * gy);
    if (r > 0) visibility_count[i] += 1.0f;
    if (r > max_radii[i]) max_radii[i] = r;
}

cudaError_t launch_view_stats(long long P, const float* vgrad, int vstride, const int* radii, float* grad_norm_sum,
                              float* visibility_count, int* max_radii, cudaStream_t stream) {
    if (P <= 0) return cudaSuccess;
    view_stats_kernel<<<(unsigned)((P + 255) / 256), 256, 0, stream>>>(P, vgrad, vstride, radii, grad_norm_sum, visibility_count, max_radii);
    return cudaGetLastError();
}

cudaError_t launch_union_maps(long long P, const int* radii, const int* cs, int* slot_of, long long* idx, cudaStream_t stream) {
    if (P <= 0) return cudaSuccess;
    union_maps_kernel<<<(unsigned)((P + 255) / 256), 256, 0, stream>>>(P, radii, cs, slot_of, idx);
    return cudaGetLastError();
}

cudaError_t launch_rows_zero_check(int n, const float* const* tensors, const int* widths, long long P, const int* mask_radii,
                                   int* flag, cudaStream_t stream) {
    if (n <= 0 || P <= 0) return cudaSuccess;
    CheckTable tb;
    for (int i = 0; i < n; ++i) {
        tb.tensor[i] = tensors[i];
        tb.width[i] = widths[i];
    }
    long long blocks = (P + 255) / 256;
    if (blocks > device_sm_count() * 8) blocks = device_sm_count() * 8;
    rows_zero_check_kernel<<<dim3((unsigned)blocks, (unsigned)n, 1), 256, 0, stream>>>(tb, P, mask_radii, flag);
    return cudaGetLastError();
}

cudaError_t launch_pack_rows(bool unpack, int n, float* const* tensors, const int* widths, const long long* block_off,
                             const long long* idx, long long K, float* flat, cudaStream_t stream) {
    if (n <= 0 || K <= 0) return cudaSuccess;
    PackTable tb;
    tb.n = n;
    long long widest = 1;
    for (int i = 0; i < n; ++i) {
        tb.tensor[i] = tensors[i];
        tb.width[i] = widths[i];
        tb.block_off[i] = block_off[i];
        tb.vec4[i] = (widths[i] % 4 == 0) && (reinterpret_cast<uintptr_t>(tensors[i]) % 16 == 0) &&
                     (reinterpret_cast<uintptr_t>(flat + block_off[i]) % 16 == 0);
        const long long work = K * (tb.vec4[i] ? widths[i] / 4 : widths[i]);
        if (work > widest) widest = work;
    }
    long long blocks = (widest + 255) / 256;
    if (blocks > device_sm_count() * 16) blocks = device_sm_count() * 16;   // grid-stride beyond 16 CTAs per SM
    dim3 grid((unsigned)blocks, (unsigned)n, 1);
    if (unpack) pack_rows_kernel<true><<<grid, 256, 0, stream>>>(tb, idx, K, flat);
    else pack_rows_kernel<false><<<grid, 256, 0, stream>>>(tb, idx, K, flat);
    return cudaGetLastError();
}

}  // namespace fdgs
