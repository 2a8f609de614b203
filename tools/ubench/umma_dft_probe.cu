// umma_dft_probe.cu -- MEASUREMENT ONLY (not part of libsnoutrx.so): the 48 -> 40-bin transform of the BLE channelizer on the
// 5th-generation tensor cores (tcgen05.mma, accumulators in TMEM) against the FP32 FFT the product kernel uses, on the same
// inputs.  It answers north_star's "tensor cores only if ncu shows it beats the FMA path" with numbers:
//
//   tensor path, per tile of 128 output times (one 128-thread CTA):
//     split   every thread (= output time) splits its 96 FP32 branch outputs (48 complex) into fp16 hi + fp16 lo
//             (hi = the value with the low 13 mantissa bits cleared: exact in fp16; lo = fp16(v - hi)) and stores both as
//             the A operand [M = 128 x K = 96] in the canonical K-major no-swizzle core-matrix layout (8 rows x 16 B);
//     mma     ONE thread issues 18 tcgen05.mma.kind::f16 (M128 N80 K16): D = Ahi Bhi + Alo Bhi + Ahi Blo over 6 K steps,
//             B = the [N = 80 x K = 96] real form of exp(+j 2 pi q r / 48) for the 40 used bins, hi/lo split, in shared
//             memory once per CTA; FP32 accumulation in 80 TMEM columns; tcgen05.commit -> mbarrier;
//     read    each warp tcgen05.ld's its 32 lanes x 80 columns and folds them into a checksum (tile 0 is also written out).
//   FMA path: one lane per output time runs Idft3xQ<48> (csrc/fft.cuh: 3 x 16-point + radix-3 combination, packed FFMA2),
//             the code k_pfb_ble runs, and folds the same 40 bins.
//
// Usage:  umma_dft_probe [n_tiles]     prints max |tensor - float64 DFT| / rms, and microseconds per 10^6 output times for
//         (a) tensor: split + mma + read, (b) tensor: mma + read only (A written once), (c) FMA FFT.
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../../snout_b200/csrc/fft.cuh"
#include "../../snout_b200/csrc/pfb.cuh"      // ble_channel_of_q, smem_u32, mbar_*

using namespace snrx;

constexpr int kM = 128, kK = 96, kN = 80;
// The 96 inputs of an output time are MADE in registers from one coalesced 16-byte load (in the product they are the FIR's
// accumulators and never come from memory): x[k] = s[k & 3] * ck + dk, one FFMA each, ck / dk compile-time constants.
__host__ __device__ inline float gen_c(int k) { return 0.25f + 0.03125f * (float)((k * 7) % 23); }
__host__ __device__ inline float gen_d(int k) { return -0.5f + 0.0625f * (float)((k * 5) % 17); }
__host__ __device__ inline float gen_x(const float (&s)[4], int k) { return fmaf(s[k & 3], gen_c(k), gen_d(k)); }
constexpr int kLbo = 128;                      // bytes between core matrices along K
constexpr int kSboA = (kK / 8) * kLbo;         // bytes between 8-row groups of A (1536)
constexpr int kSboB = (kK / 8) * kLbo;
constexpr int kABytes = (kM / 8) * kSboA;      // 24576 per hi / lo
constexpr int kBBytes = (kN / 8) * kSboB;      // 15360 per hi / lo
constexpr int kSmem = 2 * kABytes + 2 * kBBytes + 64;
constexpr uint32_t kTmemCols = 128;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { fprintf(stderr, "%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e)); exit(1); } } while (0)

__host__ __device__ inline int q_of_used(int u) { return u < 21 ? u : u + 8; }   // the 40 bins that carry a channel: q = 0..20, 29..47

// byte offset of element (row, k) of a K-major no-swizzle operand: core matrix = 8 rows x 16 bytes
__host__ __device__ inline int core_off(int row, int k, int sbo) { return (row & 7) * 16 + (row >> 3) * sbo + (k >> 3) * kLbo + (k & 7) * 2; }

__device__ __forceinline__ uint64_t smem_desc(uint32_t addr, int sbo) {
    uint64_t d = 0;
    d |= (uint64_t)((addr >> 4) & 0x3FFF);                 // start address
    d |= (uint64_t)((kLbo >> 4) & 0x3FFF) << 16;           // leading-dimension byte offset (K direction)
    d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;            // stride byte offset (M / N direction)
    d |= (uint64_t)1 << 46;                                // descriptor version (Blackwell)
    return d;                                              // base offset 0, swizzle none
}
constexpr uint32_t kIdesc = (1u << 4) | ((uint32_t)(kN >> 3) << 17) | ((uint32_t)(kM >> 4) << 24);   // f16 x f16 -> f32, K-major A and B

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

// v: [tiles_in][128] float4 seeds (gen_x makes the 96 inputs: re, im of branch r at k = 2r, 2r + 1); bmat: [2][kBBytes] fp16 hi / lo already in core layout
template <bool SPLIT_EVERY_TILE>
__global__ void __launch_bounds__(128, 2) k_dft_tensor(const float* __restrict__ v, int tiles_in, const uint8_t* __restrict__ bmat,
                                                       int n_tiles, float* __restrict__ out0, float* __restrict__ checksum) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* a_hi = smem;
    uint8_t* a_lo = smem + kABytes;
    uint8_t* b_hi = smem + 2 * kABytes;
    uint8_t* b_lo = b_hi + kBBytes;
    uint64_t* bar = reinterpret_cast<uint64_t*>(b_lo + kBBytes);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);
    const int tid = threadIdx.x, warp = tid >> 5;

    for (int i = tid; i < 2 * kBBytes / 16; i += 128) reinterpret_cast<uint4*>(b_hi)[i] = reinterpret_cast<const uint4*>(bmat)[i];
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
        // ---- split: this thread's output time = row tid of the tile
        if (SPLIT_EVERY_TILE || t == (int)blockIdx.x) {
            const float4 sd = __ldg(reinterpret_cast<const float4*>(v) + (size_t)(t % tiles_in) * kM + tid);
            const float seed[4] = {sd.x, sd.y, sd.z, sd.w};
#pragma unroll
            for (int g = 0; g < kK / 8; g++) {
                float x[8];
#pragma unroll
                for (int j = 0; j < 8; j++) x[j] = gen_x(seed, 8 * g + j);
                uint32_t hi[4], lo[4];
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    const float h0 = __uint_as_float(__float_as_uint(x[2 * j]) & 0xFFFFE000u), h1 = __uint_as_float(__float_as_uint(x[2 * j + 1]) & 0xFFFFE000u);
                    const __half2 hh = __floats2half2_rn(h0, h1);                       // exact
                    const __half2 ll = __floats2half2_rn(x[2 * j] - h0, x[2 * j + 1] - h1);
                    hi[j] = *reinterpret_cast<const uint32_t*>(&hh);
                    lo[j] = *reinterpret_cast<const uint32_t*>(&ll);
                }
                const int off = core_off(tid, 8 * g, kSboA);
                *reinterpret_cast<uint4*>(a_hi + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
                *reinterpret_cast<uint4*>(a_lo + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
            }
            asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");          // generic-proxy stores before the tensor core reads them
        }
        asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
        __syncthreads();
        // ---- mma: one thread
        if (tid == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
            const uint32_t ah = smem_u32(a_hi), al = smem_u32(a_lo), bh = smem_u32(b_hi), bl = smem_u32(b_lo);
#pragma unroll
            for (int term = 0; term < 3; term++) {
                const uint32_t a0 = term == 1 ? al : ah, b0 = term == 2 ? bl : bh;
#pragma unroll
                for (int ks = 0; ks < kK / 16; ks++)
                    umma_f16(tmem, smem_desc(a0 + ks * 2 * kLbo, kSboA), smem_desc(b0 + ks * 2 * kLbo, kSboB), (term | ks) ? 1u : 0u);
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_u32(bar)) : "memory");
        }
        // ---- read: every warp its 32 TMEM lanes
        mbar_wait(bar, parity);
        parity ^= 1u;
        asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
#pragma unroll
        for (int c = 0; c < kN / 16; c++) {
            float r[16];
            tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)(16 * c), r);
#pragma unroll
            for (int i = 0; i < 16; i++) acc += r[i];
            if (t == 0 && out0) {
#pragma unroll
                for (int i = 0; i < 16; i++) out0[(size_t)tid * kN + 16 * c + i] = r[i];
            }
        }
        asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
        __syncthreads();                                                            // TMEM and A may be overwritten
    }
    if (acc == 123.456f) checksum[0] = acc;                                         // keeps the loads alive
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem), "r"(kTmemCols) : "memory");
}

// the FP32 path of the product kernel: one lane per output time
__global__ void __launch_bounds__(128) k_dft_fma(const float* __restrict__ v, int tiles_in, int n_tiles, float* __restrict__ out0,
                                                 float* __restrict__ checksum) {
    float acc = 0.f;
    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        const float4 sd = __ldg(reinterpret_cast<const float4*>(v) + (size_t)(t % tiles_in) * kM + threadIdx.x);
        const float seed[4] = {sd.x, sd.y, sd.z, sd.w};
        cf in[48], y[48];
#pragma unroll
        for (int r = 0; r < 48; r++) in[r] = cf{gen_x(seed, 2 * r), gen_x(seed, 2 * r + 1)};
        Idft3xQ<48>::run(in, y);
#pragma unroll
        for (int u = 0; u < 40; u++) {
            const int q = q_of_used(u);
            acc += y[q].r + y[q].i;
            if (t == 0 && out0) { out0[(size_t)threadIdx.x * kN + 2 * u] = y[q].r; out0[(size_t)threadIdx.x * kN + 2 * u + 1] = y[q].i; }
        }
    }
    if (acc == 123.456f) checksum[0] = acc;
}

int main(int argc, char** argv) {
    const int n_tiles = argc > 1 ? atoi(argv[1]) : 30720;           // 30720 x 128 = 3.93 M output times = one capture-second
    const int tiles_in = 64;                                       // seeds only: the probe is about the transform
    std::vector<float> v((size_t)tiles_in * kM * 4);
    uint32_t s = 12345u;
    for (auto& x : v) { s = s * 1664525u + 1013904223u; x = ((int)(s >> 8) - (1 << 23)) / (float)(1 << 23) * 1.7f; }
    // B operand: rows n = 2u + c (bin u of the 40, re / im), columns k = 2r + c'
    std::vector<uint8_t> b(2 * kBBytes, 0);
    for (int n = 0; n < kN; n++)
        for (int k = 0; k < kK; k++) {
            const int q = q_of_used(n >> 1), r = k >> 1;
            const double th = 2.0 * M_PI * (double)((q * r) % 48) / 48.0;
            const double val = (n & 1) ? ((k & 1) ? cos(th) : sin(th)) : ((k & 1) ? -sin(th) : cos(th));
            const __half h = __float2half_rn((float)val);
            const __half l = __float2half_rn((float)(val - (double)__half2float(h)));
            const int off = core_off(n, k, kSboB);
            *reinterpret_cast<__half*>(&b[off]) = h;
            *reinterpret_cast<__half*>(&b[kBBytes + off]) = l;
        }
    float *d_v, *d_out_t, *d_out_f, *d_ck;
    uint8_t* d_b;
    CK(cudaMalloc(&d_v, v.size() * 4)); CK(cudaMalloc(&d_b, b.size())); CK(cudaMalloc(&d_out_t, kM * kN * 4)); CK(cudaMalloc(&d_out_f, kM * kN * 4));
    CK(cudaMalloc(&d_ck, 4));
    CK(cudaMemcpy(d_v, v.data(), v.size() * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(d_b, b.data(), b.size(), cudaMemcpyHostToDevice));
    CK(cudaMemset(d_out_t, 0, kM * kN * 4));
    CK(cudaFuncSetAttribute(k_dft_tensor<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem));
    CK(cudaFuncSetAttribute(k_dft_tensor<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem));
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
    const float t_full = time_it([&] { k_dft_tensor<true><<<2 * sms, 128, kSmem>>>(d_v, tiles_in, d_b, n_tiles, d_out_t, d_ck); });
    const float t_mma = time_it([&] { k_dft_tensor<false><<<2 * sms, 128, kSmem>>>(d_v, tiles_in, d_b, n_tiles, nullptr, d_ck); });
    const float t_fma = time_it([&] { k_dft_fma<<<16 * sms, 128>>>(d_v, tiles_in, n_tiles, d_out_f, d_ck); });
    // ---- accuracy of tile 0 against the float64 definition
    std::vector<float> ot(kM * kN), of(kM * kN);
    CK(cudaMemcpy(ot.data(), d_out_t, ot.size() * 4, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(of.data(), d_out_f, of.size() * 4, cudaMemcpyDeviceToHost));
    double err_t = 0, err_f = 0, ms2 = 0;
    for (int m = 0; m < kM; m++)
        for (int u = 0; u < 40; u++) {
            double re = 0, im = 0;
            const int q = q_of_used(u);
            for (int r = 0; r < 48; r++) {
                const float sd[4] = {v[(size_t)m * 4], v[(size_t)m * 4 + 1], v[(size_t)m * 4 + 2], v[(size_t)m * 4 + 3]};
                const double th = 2.0 * M_PI * (double)((q * r) % 48) / 48.0, a = gen_x(sd, 2 * r), bb = gen_x(sd, 2 * r + 1);
                re += a * cos(th) - bb * sin(th);
                im += a * sin(th) + bb * cos(th);
            }
            ms2 += re * re + im * im;
            err_t = fmax(err_t, fmax(fabs(ot[m * kN + 2 * u] - re), fabs(ot[m * kN + 2 * u + 1] - im)));
            err_f = fmax(err_f, fmax(fabs(of[m * kN + 2 * u] - re), fabs(of[m * kN + 2 * u + 1] - im)));
        }
    const double rms = sqrt(ms2 / (kM * 40));
    const double mt = (double)n_tiles * kM / 1e6;
    printf("{\"output_times\": %.0f, \"tensor_max_err_over_rms\": %.3e, \"fma_max_err_over_rms\": %.3e, "
           "\"tensor_split_mma_read_us_per_Mtime\": %.2f, \"tensor_mma_read_us_per_Mtime\": %.2f, \"fma_fft_us_per_Mtime\": %.2f, "
           "\"sm_clocks_per_output_time\": {\"tensor_full\": %.2f, \"tensor_mma_read\": %.2f, \"fma_fft\": %.2f}, \"sms\": %d}\n",
           mt * 1e6, err_t / rms, err_f / rms, t_full * 1e3 / mt, t_mma * 1e3 / mt, t_fma * 1e3 / mt,
           t_full * 1e-3 * 1.965e9 * sms / (mt * 1e6), t_mma * 1e-3 * 1.965e9 * sms / (mt * 1e6), t_fma * 1e-3 * 1.965e9 * sms / (mt * 1e6), sms);
    return 0;
}
