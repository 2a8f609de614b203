// scan.cuh -- exclusive prefix sum of uint32 counts on the device (n inputs -> n+1 outputs, the
// last one being the total).  Used to turn per-chunk hit counts and per-window frame counts into
// write offsets so that candidates and frames come out in reference order without a sort.
#pragma once
#include "common.cuh"

#if defined(__CUDACC__)
namespace snrx {

constexpr int kScanThreads = 256;
constexpr int kScanPerThread = 4;
constexpr int kScanBlock = kScanThreads * kScanPerThread;   // 1024 items per block

__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t* total_out) {
    // v: per-thread value; returns exclusive prefix inside the block, *total_out = block sum (all threads)
    __shared__ uint32_t warp_sums[kScanThreads / 32];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    uint32_t inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { uint32_t t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
    if (lane == 31) warp_sums[wid] = inc;
    __syncthreads();
    uint32_t ws = (lane < kScanThreads / 32) ? warp_sums[lane] : 0u;
    uint32_t winc = ws;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { uint32_t t = __shfl_up_sync(0xffffffffu, winc, o); if (lane >= o) winc += t; }
    uint32_t wbase = __shfl_sync(0xffffffffu, winc - ws, wid);
    uint32_t total = __shfl_sync(0xffffffffu, winc, kScanThreads / 32 - 1);
    __syncthreads();
    *total_out = total;
    return wbase + inc - v;
}

// pass 1: block sums
__global__ void __launch_bounds__(kScanThreads) k_scan_sums(const uint32_t* __restrict__ in, uint32_t n,
                                                            uint32_t* __restrict__ sums) {
    const uint32_t base = blockIdx.x * kScanBlock + threadIdx.x * kScanPerThread;
    uint32_t s = 0;
#pragma unroll
    for (int k = 0; k < kScanPerThread; k++) if (base + k < n) s += in[base + k];
    uint32_t total;
    block_exclusive_scan(s, &total);
    if (threadIdx.x == 0) sums[blockIdx.x] = total;
}

// single-block scan of up to kScanBlock values (in place allowed); writes total to out[n]
__global__ void __launch_bounds__(kScanThreads) k_scan_small(const uint32_t* __restrict__ in, uint32_t n,
                                                             uint32_t* __restrict__ out) {
    const uint32_t base = threadIdx.x * kScanPerThread;
    uint32_t v[kScanPerThread];
    uint32_t s = 0;
#pragma unroll
    for (int k = 0; k < kScanPerThread; k++) { v[k] = (base + k < n) ? in[base + k] : 0u; s += v[k]; }
    uint32_t total;
    uint32_t pre = block_exclusive_scan(s, &total);
#pragma unroll
    for (int k = 0; k < kScanPerThread; k++) { if (base + k < n) out[base + k] = pre; pre += v[k]; }
    if (threadIdx.x == 0) out[n] = total;
}

// pass 3: per-block exclusive scan plus the scanned block offset; block 0 thread 0 writes out[n]
__global__ void __launch_bounds__(kScanThreads) k_scan_apply(const uint32_t* __restrict__ in, uint32_t n,
                                                             const uint32_t* __restrict__ block_offsets,
                                                             uint32_t n_blocks, uint32_t* __restrict__ out) {
    const uint32_t base = blockIdx.x * kScanBlock + threadIdx.x * kScanPerThread;
    uint32_t v[kScanPerThread];
    uint32_t s = 0;
#pragma unroll
    for (int k = 0; k < kScanPerThread; k++) { v[k] = (base + k < n) ? in[base + k] : 0u; s += v[k]; }
    uint32_t total;
    uint32_t pre = block_exclusive_scan(s, &total) + block_offsets[blockIdx.x];
#pragma unroll
    for (int k = 0; k < kScanPerThread; k++) { if (base + k < n) out[base + k] = pre; pre += v[k]; }
    if (blockIdx.x == 0 && threadIdx.x == 0) out[n] = block_offsets[n_blocks];
}

// Host helper: out[0..n] = exclusive scan of in[0..n-1].  `scratch` must hold
// scan_scratch_items(n) uint32.  Returns the number of kernels launched.
inline size_t scan_scratch_items(size_t n) {
    size_t total = 0;
    while (n > (size_t)kScanBlock) { n = (n + kScanBlock - 1) / kScanBlock; total += n + 1 + n + 1; }
    return total + 8;
}

inline int exclusive_scan(const uint32_t* in, uint32_t n, uint32_t* out, uint32_t* scratch, cudaStream_t st) {
    if (n <= (uint32_t)kScanBlock) {
        k_scan_small<<<1, kScanThreads, 0, st>>>(in, n, out);
        return 1;
    }
    const uint32_t nb = (n + kScanBlock - 1) / kScanBlock;
    uint32_t* sums = scratch;              // nb items
    uint32_t* offs = scratch + nb + 1;     // nb + 1 items
    k_scan_sums<<<nb, kScanThreads, 0, st>>>(in, n, sums);
    int launches = 1 + exclusive_scan(sums, nb, offs, scratch + 2 * (nb + 1), st);
    k_scan_apply<<<nb, kScanThreads, 0, st>>>(in, n, offs, nb, out);
    return launches + 1;
}

}  // namespace snrx
#endif
