// ubench.cu -- instruction-throughput microbenchmarks that decide the blend-kernel design on B200:
// FFMA vs FFMA2 (fma.rn.f32x2), FMNMX/FSEL (alu pipe), SHFL, MUFU.EX2, LDS, mma.sync tf32 m16n8k8,
// and an FFMA+SHFL mix.  Prints warp-instructions per clock per SM for a full-occupancy launch.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench tools/ubench.cu && ./ubench
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>

#define ITERS 4096
#define UNROLL 8

__device__ __forceinline__ unsigned long long pk(float a, float b) {
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
    return r;
}

template <int MODE>
__global__ void __launch_bounds__(256) bench(float* out, float seed) {
    float a[UNROLL], b = seed, c = seed * 0.5f;
    unsigned long long a2[UNROLL], b2 = pk(seed, seed), c2 = pk(c, c);
#pragma unroll
    for (int i = 0; i < UNROLL; ++i) { a[i] = seed + i + threadIdx.x; a2[i] = pk(a[i], a[i] + 1.f); }
    __shared__ float sm[1024];
    sm[threadIdx.x] = seed; sm[threadIdx.x + 256] = seed; sm[threadIdx.x + 512] = seed; sm[threadIdx.x + 768] = seed;
    __syncthreads();
    float d[4] = {0.f, 0.f, 0.f, 0.f};
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < UNROLL; ++i) {
            if (MODE == 0) a[i] = fmaf(a[i], b, c);                                  // FFMA
            if (MODE == 1) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(a2[i]) : "l"(b2), "l"(c2));   // FFMA2
            if (MODE == 2) a[i] = fminf(a[i], b + (float)i);                         // FMNMX (alu)
            if (MODE == 3) a[i] = __shfl_xor_sync(0xffffffffu, a[i], 1 + (i & 15));  // SHFL
            if (MODE == 4) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));  // MUFU
            if (MODE == 5) a[i] += sm[(threadIdx.x + (int)a[i]) & 1023];             // LDS (+FADD, F2I)
            if (MODE == 6) {                                                         // FFMA + SHFL 3:1
                a[i] = fmaf(a[i], b, c);
                if ((i & 3) == 3) a[i] = __shfl_xor_sync(0xffffffffu, a[i], 4);
            }
            if (MODE == 7) {                                                         // FFMA + FMNMX 1:1
                a[i] = fmaf(a[i], b, c);
                a[i] = fminf(a[i], 1e30f);
            }
            if (MODE == 8) {                                                         // mma.sync tf32 m16n8k8
                uint32_t A0 = __float_as_uint(a[i]), B0 = __float_as_uint(b);
                asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                             : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                             : "r"(A0), "r"(A0), "r"(A0), "r"(A0), "r"(B0), "r"(B0));
            }
            if (MODE == 9) {                                                         // FFMA2 + FMNMX 1:1
                asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(a2[i]) : "l"(b2), "l"(c2));
                a[i] = fminf(a[i], 1e30f);
            }
            if (MODE == 10) asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(a2[i]) : "l"(b2));   // FMUL2
            if (MODE == 11) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(a2[i]) : "l"(b2));   // FADD2
            if (MODE == 12) a[i] = a[i] + b;                                          // FADD
            if (MODE == 13) {                                                         // mma 8 independent accumulators? (same d: dependent)
                uint32_t A0 = __float_as_uint(a[i]), B0 = __float_as_uint(b);
                float e0 = a[i], e1 = b, e2 = c, e3 = seed;
                asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                             : "+f"(e0), "+f"(e1), "+f"(e2), "+f"(e3)
                             : "r"(A0), "r"(A0), "r"(A0), "r"(A0), "r"(B0), "r"(B0));
                a[i] = e0 + e3;
            }
        }
    }
    float s = d[0] + d[1] + d[2] + d[3];
#pragma unroll
    for (int i = 0; i < UNROLL; ++i) {
        s += a[i];
        float lo, hi;
        asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(a2[i]));
        s += lo + hi;
    }
    if (s == 123.456f) out[0] = s;
}

template <int MODE>
void run(const char* name, double instr_per_iter, int clock_khz, int sms) {
    float* out;
    cudaMalloc(&out, 4);
    const int blocks = sms * 8;   // 8 CTAs x 256 thr = 64 warps per SM
    bench<MODE><<<blocks, 256>>>(out, 1.0001f);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    bench<MODE><<<blocks, 256>>>(out, 1.0001f);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double warp_instr = (double)blocks * 8 * ITERS * UNROLL * instr_per_iter;
    const double clocks = ms * 1e-3 * clock_khz * 1e3;
    printf("%-28s %8.3f ms   %.3f warp-instr/clk/SM (at %d MHz nominal)  err=%s\n", name, ms, warp_instr / clocks / sms,
           clock_khz / 1000, cudaGetErrorString(cudaGetLastError()));
    cudaFree(out);
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    int khz = 0;
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    printf("%s  SMs=%d  clock=%d kHz\n", p.name, p.multiProcessorCount, khz);
    const int sms = p.multiProcessorCount;
    run<0>("FFMA", 1, khz, sms);
    run<1>("FFMA2 (f32x2)", 1, khz, sms);
    run<10>("FMUL2", 1, khz, sms);
    run<11>("FADD2", 1, khz, sms);
    run<12>("FADD", 1, khz, sms);
    run<2>("FMNMX (+FADD)", 2, khz, sms);
    run<3>("SHFL.BFLY", 1, khz, sms);
    run<4>("MUFU.EX2", 1, khz, sms);
    run<5>("LDS (+F2I,FADD,IADD,LOP)", 1, khz, sms);
    run<6>("FFMA x4 + SHFL x1", 1.25, khz, sms);
    run<7>("FFMA + FMNMX", 2, khz, sms);
    run<9>("FFMA2 + FMNMX", 2, khz, sms);
    run<8>("mma.tf32.m16n8k8 (dep chain)", 1, khz, sms);
    run<13>("mma.tf32.m16n8k8 (indep)+FADD", 1, khz, sms);
    return 0;
}
