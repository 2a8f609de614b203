// Micro-benchmark (development aid, not part of the product): issue rate of the packed FP32
// instruction FFMA2 (PTX fma.rn.f32x2, new on sm_100) against scalar FFMA, same flops.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2 ffma2.cu && ./ffma2
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
    unsigned long long ra = *reinterpret_cast<unsigned long long*>(&a), rb = *reinterpret_cast<unsigned long long*>(&b),
                       rc = *reinterpret_cast<unsigned long long*>(&c), rd;
    asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(rc));
    return *reinterpret_cast<float2*>(&rd);
}
__device__ __forceinline__ float2 fadd2(float2 a, float2 b) {
    unsigned long long ra = *reinterpret_cast<unsigned long long*>(&a), rb = *reinterpret_cast<unsigned long long*>(&b), rd;
    asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(rd) : "l"(ra), "l"(rb));
    return *reinterpret_cast<float2*>(&rd);
}

template <int MODE>
__global__ void __launch_bounds__(256) k(float2* out, float g0, int iters) {
    float2 acc[16];
    for (int e = 0; e < 16; e++) acc[e] = make_float2(threadIdx.x * 1e-3f + e, e * 0.5f);
    float2 x = make_float2(1.0001f, 0.9999f);
    const float2 g2 = make_float2(g0, g0);
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int e = 0; e < 16; e++) {
            if (MODE == 0) { acc[e].x = __fmaf_rn(g0, x.x, acc[e].x); acc[e].y = __fmaf_rn(g0, x.y, acc[e].y); }
            else if (MODE == 1) acc[e] = ffma2(g2, x, acc[e]);
            else if (MODE == 2) acc[e] = fadd2(x, acc[e]);
            else { acc[e].x = __fadd_rn(x.x, acc[e].x); acc[e].y = __fadd_rn(x.y, acc[e].y); }
        }
    }
    float2 s = make_float2(0.f, 0.f);
    for (int e = 0; e < 16; e++) { s.x += acc[e].x; s.y += acc[e].y; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
void run(const char* name, int warps_per_sm) {
    int sms = 148, iters = 4096;
    int threads = 256, blocks = sms * warps_per_sm * 32 / threads;
    float2* out; cudaMalloc(&out, sizeof(float2) * blocks * threads);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    k<MODE><<<blocks, threads>>>(out, 1.0000001f, 64);
    cudaEventRecord(a);
    k<MODE><<<blocks, threads>>>(out, 1.0000001f, iters);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    double flop = 2.0 * 2 * 16 * (double)iters * blocks * threads;   // (fma = 2 flop) x 2 lanes of the pair x 16 accumulators
    if (MODE >= 2) flop /= 2;
    printf("%-14s warps/SM %2d: %8.3f ms  %7.2f TFLOP/s  %6.1f Gpair-op/s/SM-lane\n", name, warps_per_sm, ms, flop / ms / 1e9, 0.0);
    cudaFree(out);
}

int main() {
    for (int w : {8, 16, 32}) {
        run<0>("FFMA scalar", w);
        run<1>("FFMA2 packed", w);
        run<3>("FADD scalar", w);
        run<2>("FADD2 packed", w);
    }
    return 0;
}
