// mma_test.cu -- precision of the split-tf32 tensor-core reduction used by blend_bwd v2.
// One warp: values x[8 cols][32 px] (random magnitudes), weights g[32 px]; compares sum_px g*x computed with
// hi/lo split HMMA (as in the kernel) against double precision.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>
#include <math.h>
#include <stdlib.h>
__device__ __forceinline__ void mma_tf32(float (&d)[4], uint32_t a0, uint32_t a2, uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a0), "r"(0u), "r"(a2), "r"(0u), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t hi_of(float x) { return __float_as_uint(x) & 0xffffe000u; }
__device__ __forceinline__ uint32_t lo_of(float x, uint32_t hi, int mask) {
    uint32_t r = __float_as_uint(x - __uint_as_float(hi));
    return mask ? (r & 0xffffe000u) : r;
}
__global__ void k(const float* x /*[8][36]*/, const float* g /*[32]*/, float* out /*[8]*/, int mode) {
    const int lane = threadIdx.x, fg = lane >> 2, ft = lane & 3;
    float c[4] = {0, 0, 0, 0};
    for (int s = 0; s < 4; ++s) {
        float ga = g[8 * s + ft], gb = g[8 * s + ft + 4];
        if ((fg & 3) != 0) { ga = 0.f; gb = 0.f; }
        uint32_t gah = hi_of(ga), gbh = hi_of(gb);
        uint32_t a0 = (fg < 4) ? gah : lo_of(ga, gah, mode & 1), a2 = (fg < 4) ? gbh : lo_of(gb, gbh, mode & 1);
        float x0 = x[fg * 36 + 8 * s + ft], x1 = x[fg * 36 + 8 * s + ft + 4];
        uint32_t x0h = hi_of(x0), x1h = hi_of(x1);
        if (mode & 2) { mma_tf32(c, a0, a2, __float_as_uint(x0), __float_as_uint(x1)); }   // raw (unmasked) hi operand
        else mma_tf32(c, a0, a2, x0h, x1h);
        mma_tf32(c, a0, a2, lo_of(x0, x0h, mode & 1), lo_of(x1, x1h, mode & 1));
    }
    c[0] += __shfl_xor_sync(0xffffffffu, c[0], 16);
    c[1] += __shfl_xor_sync(0xffffffffu, c[1], 16);
    if (fg == 0) { out[2 * ft] = c[0]; out[2 * ft + 1] = c[1]; }
}
int main() {
    float hx[8 * 36], hg[32], *dx, *dg, *dout, hout[8];
    cudaMalloc(&dx, sizeof(hx)); cudaMalloc(&dg, sizeof(hg)); cudaMalloc(&dout, sizeof(hout));
    for (int mode = 0; mode < 4; ++mode) {
        double worst = 0, sum2 = 0; int n = 0;
        srand(1);
        for (int trial = 0; trial < 2000; ++trial) {
            for (int i = 0; i < 8 * 36; ++i) hx[i] = (rand() % 3 == 0) ? 0.f : (float)(rand() / (double)RAND_MAX) * expf((rand() % 12) - 6.f);
            for (int i = 0; i < 32; ++i) hg[i] = (float)(rand() / (double)RAND_MAX * 2 - 1);
            cudaMemcpy(dx, hx, sizeof(hx), cudaMemcpyHostToDevice); cudaMemcpy(dg, hg, sizeof(hg), cudaMemcpyHostToDevice);
            k<<<1, 32>>>(dx, dg, dout, mode);
            cudaMemcpy(hout, dout, sizeof(hout), cudaMemcpyDeviceToHost);
            for (int col = 0; col < 8; ++col) {
                double ref = 0, mag = 0;
                for (int p = 0; p < 32; ++p) { ref += (double)hg[p] * hx[col * 36 + p]; mag += fabs((double)hg[p] * hx[col * 36 + p]); }
                if (mag == 0) continue;
                double e = fabs(hout[col] - ref) / mag;
                worst = fmax(worst, e); sum2 += e * e; ++n;
            }
        }
        printf("mode %d (lo masked=%d, raw hi=%d): max err / sum|terms| = %.3e  rms = %.3e  (%s)\n", mode, mode & 1, (mode >> 1) & 1, worst, sqrt(sum2 / n), cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
