// binning.cu -- tile binning: offsets scan, key emission, radix sort, instance packing.
//
// Replaces, in the reference: cub::DeviceScan::InclusiveSum (rasterizer_impl.cu:298),
// duplicateWithKeys (:71-112), cub::DeviceRadixSort::SortPairs (:325-330), the ranges memset +
// identifyTileRanges (:117-139, :332-339).
//
// The sorted order is the contract ("tile assignment bit-exact"): instances are ordered by
// (tile id, depth bits) with ties broken by emission order = Gaussian index order, exactly what
// the reference's stable LSD sort of (tile<<32 | depth_bits) produces.  We sort the same keys but
// only over the live bits [0, 32 + ceil(log2(#tiles))) -- the upper key bits are zero, so the
// result is identical to the reference's 64-bit sort (which hard-codes bit = 32,
// rasterizer_impl.cu:322).
//
// New relative to the reference: after the sort a "pack" pass gathers everything the blend
// kernels need into one 64-byte record per instance, in sorted order, so that a tile's work list
// is a contiguous byte range (streamed by cp.async.bulk in blend_fwd.cu / blend_bwd.cu) instead
// of an index list that every tile has to chase through four per-Gaussian arrays.
#include <cub/cub.cuh>
#include "fdgs_internal.h"

namespace fdgs {

size_t scan_temp_bytes(int P) {
    size_t bytes = 0;
    cub::DeviceScan::InclusiveSum(nullptr, bytes, (const uint32_t*)nullptr, (uint32_t*)nullptr, P);
    return bytes;
}

cudaError_t launch_scan(void* temp, size_t temp_bytes, const uint32_t* in, uint32_t* out, int P,
                        cudaStream_t stream) {
    return cub::DeviceScan::InclusiveSum(temp, temp_bytes, in, out, P, stream);
}

namespace {

// reference: rasterizer_impl.cu:71-112 duplicateWithKeys
__global__ void __launch_bounds__(256) emit_keys_kernel(int P, const float2* __restrict__ means2D,
                                                        const float* __restrict__ depths,
                                                        const uint32_t* __restrict__ offsets,
                                                        const int* __restrict__ radii, int grid_x, int grid_y,
                                                        uint64_t* __restrict__ keys, uint32_t* __restrict__ vals) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= P) return;
    const int radius = radii[idx];
    if (radius <= 0) return;
    uint32_t off = (idx == 0) ? 0u : offsets[idx - 1];
    const float2 p = means2D[idx];
    int x0, y0, x1, y1;
    get_rect(p.x, p.y, radius, grid_x, grid_y, x0, y0, x1, y1);
    const uint64_t depth_bits = (uint64_t)__float_as_uint(depths[idx]);
    for (int y = y0; y < y1; ++y) {
        for (int x = x0; x < x1; ++x) {
            const uint64_t key = ((uint64_t)(uint32_t)(y * grid_x + x) << 32) | depth_bits;
            keys[off] = key;
            vals[off] = (uint32_t)idx;
            ++off;
        }
    }
}

__global__ void __launch_bounds__(256)
pack_instances_kernel(int R, const uint64_t* __restrict__ keys, const uint32_t* __restrict__ point_list,
                      const float2* __restrict__ means2D, const float4* __restrict__ conic_opacity,
                      const float* __restrict__ rgb, const float* __restrict__ depths,
                      const float* __restrict__ flows, InstRec* __restrict__ recs, uint2* __restrict__ ranges) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= R) return;
    const uint64_t key = keys[r];
    const uint32_t tile = (uint32_t)(key >> 32);
    // reference: rasterizer_impl.cu:117-139 identifyTileRanges
    if (r == 0) {
        ranges[tile].x = 0;
    } else {
        const uint32_t prev = (uint32_t)(keys[r - 1] >> 32);
        if (prev != tile) {
            ranges[prev].y = (uint32_t)r;
            ranges[tile].x = (uint32_t)r;
        }
    }
    if (r == R - 1) ranges[tile].y = (uint32_t)R;

    const uint32_t g = point_list[r];
    const float2 xy = means2D[g];
    const float4 co = conic_opacity[g];
    const float o = co.w;
    // Box of pixels where alpha = min(0.99, o*exp(power)) can reach 1/255 (the reference's
    // cut-off, forward.cu:590).  Conservative: used only to SKIP work, never to change a result.
    float ex, ey, pmin;
    const float det = co.x * co.z - co.y * co.y;
    if (o < 0.00392156886f) {   // alpha <= o < 1/255 everywhere: never contributes
        ex = ey = -INFINITY;
        pmin = INFINITY;
    } else if (!(o <= 3.0e38f) || !(co.x > 0.f) || !(co.z > 0.f) || !(det > 0.f) || !(det <= 3.0e38f)) {
        ex = ey = INFINITY;   // odd inputs (NaN/inf opacity, non-PD conic): no culling
        pmin = -INFINITY;
    } else {
        const float qs = logf(255.0f * o) + 0.02f;   // contribute only where q(d) <= qs
        ex = sqrtf(2.0f * qs * co.z / det) * 1.001f + 0.01f;
        ey = sqrtf(2.0f * qs * co.x / det) * 1.001f + 0.01f;
        pmin = -qs;
        if (!(ex <= 3.0e38f) || !(ey <= 3.0e38f)) { ex = ey = INFINITY; pmin = -INFINITY; }
    }
    InstRec rec;
    rec.q0 = make_float4(xy.x, xy.y, pmin, __uint_as_float(g));
    rec.q1 = co;
    rec.q2 = make_float4(rgb[3 * g + 0], rgb[3 * g + 1], rgb[3 * g + 2], depths[g]);
    const float fx = flows ? flows[2 * g + 0] : 0.f, fy = flows ? flows[2 * g + 1] : 0.f;
    rec.q3 = make_float4(fx, fy, ex, ey);
    float4* dst = reinterpret_cast<float4*>(recs + r);
    dst[0] = rec.q0;
    dst[1] = rec.q1;
    dst[2] = rec.q2;
    dst[3] = rec.q3;
}

}  // namespace

cudaError_t launch_emit_keys(int P, const float* means2D, const float* depths, const uint32_t* offsets,
                             const int* radii, int grid_x, int grid_y, uint64_t* keys, uint32_t* vals,
                             cudaStream_t stream) {
    if (P <= 0) return cudaSuccess;
    emit_keys_kernel<<<(P + 255) / 256, 256, 0, stream>>>(P, reinterpret_cast<const float2*>(means2D), depths,
                                                          offsets, radii, grid_x, grid_y, keys, vals);
    return cudaGetLastError();
}

size_t sort_temp_bytes(int R) {
    size_t bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, bytes, (const uint64_t*)nullptr, (uint64_t*)nullptr,
                                    (const uint32_t*)nullptr, (uint32_t*)nullptr, R);
    return bytes;
}

cudaError_t launch_sort_pairs(void* temp, size_t temp_bytes, const uint64_t* keys_in, uint64_t* keys_out,
                              const uint32_t* vals_in, uint32_t* vals_out, int R, int end_bit,
                              cudaStream_t stream) {
    if (R <= 0) return cudaSuccess;
    return cub::DeviceRadixSort::SortPairs(temp, temp_bytes, keys_in, keys_out, vals_in, vals_out, R, 0, end_bit,
                                           stream);
}

cudaError_t launch_pack_instances(int R, const uint64_t* keys_sorted, const uint32_t* point_list,
                                  const float* means2D, const float* conic_opacity, const float* rgb,
                                  const float* depths, const float* flows, InstRec* recs, uint2* ranges,
                                  cudaStream_t stream) {
    if (R <= 0) return cudaSuccess;
    pack_instances_kernel<<<(R + 255) / 256, 256, 0, stream>>>(
        R, keys_sorted, point_list, reinterpret_cast<const float2*>(means2D),
        reinterpret_cast<const float4*>(conic_opacity), rgb, depths, flows, recs, ranges);
    return cudaGetLastError();
}

}  // namespace fdgs
