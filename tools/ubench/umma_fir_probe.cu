// umma_fir_probe.cu -- MEASUREMENT ONLY (not part of libsnoutrx.so): the polyphase FIR of the channelizer "recast as a dense
// contraction" on the tensor cores (north_star's condition for using them), against the FP32 FMA form the product kernel runs.
//
// The recast.  For one input phase rho the two branches rho, rho + 24 are ONE 16-tap FIR over the decimated
// sequence X[c] = x[24 c - rho]:  v_b[m] = sum_p h_b[p] X[m + 16 - 2 p - b],  b = 0, 1,  p = 0..7.  With ROWS = blocks of 8 output
// times the data matrix is Hankel, A[j][k] = X[8 j + k], k = 0..31, and it is never materialised: a K-major no-swizzle UMMA
// descriptor addresses it IN PLACE in the contiguous fp16 sequence -- core-matrix row pitch 16 B = 8 samples (one row block),
// leading-dimension offset 16 B (the next 8 K-elements overlap the next row), stride offset 128 B (8 row blocks).  B is the
// [N = 16 x K = 32] Toeplitz tap matrix B[(e, b)][k] = h_b[p] where k = e + 16 - 2 p - b.  fp16 hi/lo split of samples and
// taps, 3 products, FP32 accumulation in TMEM: per rho and per re / im 6 tcgen05.mma of M128 N16 K16 for 1024 output times.
//
// One CTA iteration = 1024 output times x 6 of the 24 phases (the other 18 are the same work): 72 MMAs, 192 TMEM columns.
// Prints accuracy against float64 and SM clocks per output time (scaled to all 24 phases) for
//   (a) tensor: make + split + store the sequences, mma, tcgen05.ld;   (b) tensor: mma + tcgen05.ld only;
//   (c) FMA: the same FIR as 128 packed FFMA2 per (phase, 8 output times) from registers, as pfb_fir_thread does it.
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../../snout_b200/csrc/pfb.cuh"      // smem_u32, mbar_*, f2_fma

using namespace snrx;

constexpr int kRho = 6;                        // phases per iteration
constexpr int kRows = 128;                     // row blocks = TMEM lanes; 8 output times each
constexpr int kSeq = 8 * kRows + 24;           // samples of one sequence (1048)
constexpr int kSeqBytes = ((kSeq * 2 + 15) / 16) * 16;          // fp16
constexpr int kParts = 4;                      // re hi, re lo, im hi, im lo
constexpr int kBBytes = 16 * 32 * 2;           // one tap matrix (hi or lo): 2 row groups x 4 core matrices x 128 B
constexpr int kSmem = kRho * kParts * kSeqBytes + kRho * 2 * kBBytes + 64;
constexpr uint32_t kTmemCols = 256;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { fprintf(stderr, "%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e)); exit(1); } } while (0)

// sample c of phase i of a tile: made from the seed of the 8-sample group it belongs to -- ONE FFMA per component (in the product
// a sample costs one shared-memory load): the seed is scaled per phase (4 multiplies per group and phase), the coefficients depend
// on the position inside the group only and fold into immediates
__host__ __device__ inline float phase_scale(int i) { return 1.0f + 0.125f * (float)i; }
__host__ __device__ inline float gen_re(const float (&s)[4], int i, int c) { (void)i; const int l = c & 7; return fmaf(s[l & 3], 0.25f + 0.0625f * (float)((l * 3) % 13), -0.375f + 0.03125f * (float)((l * 5) % 7)); }
__host__ __device__ inline float gen_im(const float (&s)[4], int i, int c) { (void)i; const int l = c & 7; return fmaf(s[(l + 1) & 3], -0.5f + 0.0625f * (float)((l * 5) % 11), 0.25f - 0.03125f * (float)((l * 3) % 5)); }
__host__ __device__ inline float tap(int i, int b, int p) { return (0.9f - 0.1f * (float)p) * ((b ? 0.7f : 1.0f) + 0.01f * (float)i) * ((p & 1) ? -1.0f : 1.0f) * 0.25f; }

__host__ __device__ inline int b_off(int n, int k) { return (n & 7) * 16 + (n >> 3) * 512 + (k >> 3) * 128 + (k & 7) * 2; }

__device__ __forceinline__ uint64_t desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((addr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | ((uint64_t)1 << 46);
}
constexpr uint32_t kIdesc = (1u << 4) | ((uint32_t)(16 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);   // f16 x f16 -> f32, M128 N16

__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t accumulate) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem_d), "l"(da), "l"(db),
                 "r"(kIdesc), "r"(accumulate)
                 : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                   "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; i++) v[i] = __uint_as_float(r[i]);
}

// seeds: [tiles_in][131] float4 (one per 8-sample group of a sequence: 128 + 3 tail groups); bmat: [kRho][2][kBBytes]
template <bool MAKE_EVERY_TILE>
__global__ void __launch_bounds__(128, 1) k_fir_tensor(const float4* __restrict__ seeds, int tiles_in, const uint8_t* __restrict__ bmat, int n_tiles,
                                                       float* __restrict__ out0, float* __restrict__ checksum) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* seq = smem;                                            // [kRho][kParts][kSeqBytes]
    uint8_t* bm = smem + kRho * kParts * kSeqBytes;                 // [kRho][2][kBBytes]
    uint64_t* bar = reinterpret_cast<uint64_t*>(bm + kRho * 2 * kBBytes);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < kRho * 2 * kBBytes / 16; i += 128) reinterpret_cast<uint4*>(bm)[i] = reinterpret_cast<const uint4*>(bmat)[i];
    if (tid == 0) mbar_init(bar, 1);
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(tmem_slot)), "r"(kTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    const uint32_t tmem = *tmem_slot;

    float acc = 0.f;
    uint32_t parity = 0;
    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        if (MAKE_EVERY_TILE || t == (int)blockIdx.x) {
            // ---- make + split + store: group g of 8 samples (thread tid; threads 0..2 also the three tail groups)
            for (int g = tid; g < kSeq / 8; g += 128) {
                const float4 sd = __ldg(seeds + (size_t)(t % tiles_in) * (kSeq / 8) + g);
                const float s4[4] = {sd.x, sd.y, sd.z, sd.w};
#pragma unroll
                for (int i = 0; i < kRho; i++) {
                    uint32_t w[kParts][4];
                    const float ps = phase_scale(i);
                    const float s4i[4] = {s4[0] * ps, s4[1] * ps, s4[2] * ps, s4[3] * ps};
#pragma unroll
                    for (int j = 0; j < 4; j++) {
                        const int c = 2 * j;                          // position inside the group of 8
                        const float r0 = gen_re(s4i, i, c), r1 = gen_re(s4i, i, c + 1), i0 = gen_im(s4i, i, c), i1 = gen_im(s4i, i, c + 1);
                        const float rh0 = __uint_as_float(__float_as_uint(r0) & 0xFFFFE000u), rh1 = __uint_as_float(__float_as_uint(r1) & 0xFFFFE000u);
                        const float ih0 = __uint_as_float(__float_as_uint(i0) & 0xFFFFE000u), ih1 = __uint_as_float(__float_as_uint(i1) & 0xFFFFE000u);
                        const __half2 a = __floats2half2_rn(rh0, rh1), b = __floats2half2_rn(r0 - rh0, r1 - rh1);
                        const __half2 cc = __floats2half2_rn(ih0, ih1), d = __floats2half2_rn(i0 - ih0, i1 - ih1);
                        w[0][j] = *reinterpret_cast<const uint32_t*>(&a); w[1][j] = *reinterpret_cast<const uint32_t*>(&b);
                        w[2][j] = *reinterpret_cast<const uint32_t*>(&cc); w[3][j] = *reinterpret_cast<const uint32_t*>(&d);
                    }
#pragma unroll
                    for (int p = 0; p < kParts; p++)
                        *reinterpret_cast<uint4*>(seq + (i * kParts + p) * kSeqBytes + 16 * g) = make_uint4(w[p][0], w[p][1], w[p][2], w[p][3]);
                }
            }
            asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
        }
        asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
        __syncthreads();
        if (tid == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
            const uint32_t s0 = smem_u32(seq), b0 = smem_u32(bm);
#pragma unroll
            for (int i = 0; i < kRho; i++)
#pragma unroll
                for (int c = 0; c < 2; c++) {                       // re, im
                    const uint32_t hi = s0 + (i * kParts + 2 * c) * kSeqBytes, lo = hi + kSeqBytes;
                    const uint32_t bh = b0 + i * 2 * kBBytes, bl = bh + kBBytes;
#pragma unroll
                    for (int term = 0; term < 3; term++)
#pragma unroll
                        for (int ks = 0; ks < 2; ks++)
                            umma_f16(tmem + (uint32_t)((i * 2 + c) * 16), desc((term == 1 ? lo : hi) + ks * 32, 16, 128),
                                     desc((term == 2 ? bl : bh) + ks * 256, 128, 512), (term | ks) ? 1u : 0u);
                }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_u32(bar)) : "memory");
        }
        mbar_wait(bar, parity);
        parity ^= 1u;
        asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
#pragma unroll
        for (int c = 0; c < kRho * 2; c++) {
            float r[16];
            tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)(16 * c), r);
#pragma unroll
            for (int i = 0; i < 16; i++) acc += r[i];
            if (t == 0 && out0) {
#pragma unroll
                for (int i = 0; i < 16; i++) out0[((size_t)tid * kRho * 2 + c) * 16 + i] = r[i];
            }
        }
        asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
        __syncthreads();
    }
    if (acc == 123.456f) checksum[0] = acc;
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem), "r"(kTmemCols) : "memory");
}

// FP32 form: thread = row block j; per phase 24 samples made in registers, 8 times x 2 branches x 8 taps = 128 packed FFMA2
__global__ void __launch_bounds__(128) k_fir_fma(const float4* __restrict__ seeds, int tiles_in, int n_tiles, float* __restrict__ out0,
                                                 float* __restrict__ checksum) {
    float acc = 0.f;
    const int j = threadIdx.x;
    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        const float4* sp = seeds + (size_t)(t % tiles_in) * (kSeq / 8) + j;
        float s4[3][4];
#pragma unroll
        for (int q = 0; q < 3; q++) { const float4 sd = __ldg(sp + q); s4[q][0] = sd.x; s4[q][1] = sd.y; s4[q][2] = sd.z; s4[q][3] = sd.w; }
#pragma unroll 1
        for (int i = 0; i < kRho; i++) {
            const float ps = phase_scale(i);
            float2 x[24];
#pragma unroll
            for (int k = 0; k < 24; k++) {
                const float si[4] = {s4[k >> 3][0] * ps, s4[k >> 3][1] * ps, s4[k >> 3][2] * ps, s4[k >> 3][3] * ps};
                x[k] = make_float2(gen_re(si, i, k), gen_im(si, i, k));
            }
            float h[2][8];
#pragma unroll
            for (int b = 0; b < 2; b++)
#pragma unroll
                for (int p = 0; p < 8; p++) h[b][p] = tap(i, b, p);
#pragma unroll
            for (int b = 0; b < 2; b++)
#pragma unroll
                for (int e = 0; e < 8; e++) {
                    float2 v = make_float2(0.f, 0.f);
#pragma unroll
                    for (int p = 0; p < 8; p++) v = f2_fma(make_float2(h[b][p], h[b][p]), x[e + 16 - 2 * p - b], v);
                    acc += v.x + v.y;
                    if (t == 0 && out0) { out0[((size_t)j * kRho * 2 + 2 * i) * 16 + 2 * e + b] = v.x; out0[((size_t)j * kRho * 2 + 2 * i + 1) * 16 + 2 * e + b] = v.y; }
                }
        }
    }
    if (acc == 123.456f) checksum[0] = acc;
}

int main(int argc, char** argv) {
    const int n_tiles = argc > 1 ? atoi(argv[1]) : 3840;           // 3840 x 1024 = 3.93 M output times = one capture-second (6 of 24 phases)
    const int tiles_in = 16, groups = kSeq / 8;
    std::vector<float> seeds((size_t)tiles_in * groups * 4);
    uint32_t s = 777u;
    for (auto& x : seeds) { s = s * 1664525u + 1013904223u; x = ((int)(s >> 8) - (1 << 23)) / (float)(1 << 23) * 1.5f; }
    std::vector<uint8_t> b((size_t)kRho * 2 * kBBytes, 0);
    for (int i = 0; i < kRho; i++)
        for (int n = 0; n < 16; n++) {                             // column n = 2 e + b of D
            const int e = n >> 1, br = n & 1;
            for (int p = 0; p < 8; p++) {
                const int k = e + 16 - 2 * p - br;
                const float h = tap(i, br, p);
                const __half hh = __float2half_rn(h), hl = __float2half_rn(h - __half2float(hh));
                *reinterpret_cast<__half*>(&b[(size_t)(i * 2) * kBBytes + b_off(n, k)]) = hh;
                *reinterpret_cast<__half*>(&b[(size_t)(i * 2 + 1) * kBBytes + b_off(n, k)]) = hl;
            }
        }
    float4* d_seeds; uint8_t* d_b; float *d_out_t, *d_out_f, *d_ck;
    const size_t out_n = (size_t)kRows * kRho * 2 * 16;
    CK(cudaMalloc(&d_seeds, seeds.size() * 4)); CK(cudaMalloc(&d_b, b.size())); CK(cudaMalloc(&d_out_t, out_n * 4)); CK(cudaMalloc(&d_out_f, out_n * 4));
    CK(cudaMalloc(&d_ck, 4));
    CK(cudaMemcpy(d_seeds, seeds.data(), seeds.size() * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(d_b, b.data(), b.size(), cudaMemcpyHostToDevice));
    CK(cudaMemset(d_out_t, 0, out_n * 4));
    CK(cudaFuncSetAttribute(k_fir_tensor<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem));
    CK(cudaFuncSetAttribute(k_fir_tensor<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem));
    int sms = 0; CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    auto time_it = [&](auto launch) {
        launch(); CK(cudaDeviceSynchronize());
        float best = 1e30f;
        for (int rep = 0; rep < 5; rep++) {
            CK(cudaEventRecord(e0)); launch(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
            float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); best = fminf(best, ms);
        }
        CK(cudaGetLastError());
        return best;
    };
    const int ctas = 2;                                            // per SM: 63 KB of shared memory and 256 TMEM columns each
    const float t_full = time_it([&] { k_fir_tensor<true><<<ctas * sms, 128, kSmem>>>(d_seeds, tiles_in, d_b, n_tiles, d_out_t, d_ck); });
    const float t_mma = time_it([&] { k_fir_tensor<false><<<ctas * sms, 128, kSmem>>>(d_seeds, tiles_in, d_b, n_tiles, nullptr, d_ck); });
    const float t_fma = time_it([&] { k_fir_fma<<<16 * sms, 128>>>(d_seeds, tiles_in, n_tiles, d_out_f, d_ck); });
    std::vector<float> ot(out_n), of(out_n);
    CK(cudaMemcpy(ot.data(), d_out_t, out_n * 4, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(of.data(), d_out_f, out_n * 4, cudaMemcpyDeviceToHost));
    double err_t = 0, err_f = 0, ms2 = 0; size_t cnt = 0;
    for (int j = 0; j < kRows; j++)
        for (int i = 0; i < kRho; i++)
            for (int c = 0; c < 2; c++)
                for (int n = 0; n < 16; n++) {
                    const int e = n >> 1, br = n & 1;
                    double v = 0;
                    for (int p = 0; p < 8; p++) {
                        const int cc = 8 * j + e + 16 - 2 * p - br, g = cc >> 3;
                        const float ps = phase_scale(i);
                        const float s4[4] = {seeds[(size_t)g * 4] * ps, seeds[(size_t)g * 4 + 1] * ps, seeds[(size_t)g * 4 + 2] * ps, seeds[(size_t)g * 4 + 3] * ps};
                        v += (double)tap(i, br, p) * (double)(c ? gen_im(s4, i, cc) : gen_re(s4, i, cc));
                    }
                    const size_t o = ((size_t)j * kRho * 2 + 2 * i + c) * 16 + n;
                    ms2 += v * v; cnt++;
                    err_t = fmax(err_t, fabs(ot[o] - v)); err_f = fmax(err_f, fabs(of[o] - v));
                }
    const double rms = sqrt(ms2 / cnt), mt = (double)n_tiles * 8 * kRows / 1e6, scale = 24.0 / kRho;
    printf("{\"output_times\": %.0f, \"phases_per_iteration\": %d, \"tensor_max_err_over_rms\": %.3e, \"fma_max_err_over_rms\": %.3e, "
           "\"sm_clocks_per_output_time_all_24_phases\": {\"tensor_make_split_mma_read\": %.2f, \"tensor_mma_read\": %.2f, \"fma_fir\": %.2f}, "
           "\"us_per_Mtime_all_24_phases\": {\"tensor_full\": %.2f, \"tensor_mma_read\": %.2f, \"fma_fir\": %.2f}, \"sms\": %d}\n",
           mt * 1e6, kRho, err_t / rms, err_f / rms, scale * t_full * 1e-3 * 1.965e9 * sms / (mt * 1e6), scale * t_mma * 1e-3 * 1.965e9 * sms / (mt * 1e6),
           scale * t_fma * 1e-3 * 1.965e9 * sms / (mt * 1e6), scale * t_full * 1e3 / mt, scale * t_mma * 1e3 / mt, scale * t_fma * 1e3 / mt, sms);
    return 0;
}
