// pfb_zb.cuh -- wideband channelizer for the 16 IEEE 802.15.4 channels, fused with the
// quadrature (FM) demodulator.
//
// Same filterbank definition as pfb.cuh (no reference counterpart), evaluated as the full
// 96-branch bank because the Zigbee centres 2405 + 5 i MHz fall on odd and even 1 MHz bins:
//   v_r[m] = sum_{p<L/96} h[r + 96 p] x[24 m - r - 96 p]                 r = 0..95
//   y_k[m] = (-j)^(k m) sum_r v_r[m] exp(+j 2 pi k r / 96)               k in the 16 Zigbee bins
// The 96-point inverse DFT is three 32-point transforms (r = 3 r2 + r1) combined only for the 16
// needed bins.  The discriminator f[n] = arg(y'[n] conj(y'[n-1])) (analog.quadrature_demod_cf(1),
// top_block.py:73) is taken on the un-rotated outputs and the constant factor (-j)^k applied to
// the product -- a swap/negation, so the value is bit-identical to demodulating the rotated
// stream y' (which is what the oracle does with the engine's debug stream).
#pragma once
#include "pfb.cuh"
#include "zb.cuh"

namespace snrx {

constexpr int kZbChunkT = 8;                              // output times per FIR thread (4 branches each)
constexpr int kZbFirThreads = 24 * (kTileT / kZbChunkT);  // 384
constexpr int kZbVStride = 129;                           // float2 per V row (odd: conflict free)

SNRX_HD constexpr int zb_bin_of_slot(int c) { return ((5 * c - 35) % 96 + 96) % 96; }   // channel 11 + c

template <int NT> struct PfbZbGeom {
    using G = PfbGeom<NT, kZbChunkT>;
    static constexpr int kSmemBytes = (G::kXsLen + 96 * kZbVStride) * 8 + 5 * 16 * 8 + 260 * 4;
    static_assert(kSmemBytes <= 227 * 1024, "tile does not fit in shared memory");
};

template <int C>
SNRX_HD void pfb_dft96_combine(const cf* f0, const cf* f1, const cf* f2, cf (&y)[16]);

// 96-point inverse DFT of one output time restricted to the 16 Zigbee bins (un-rotated).
SNRX_HD void pfb_dft96_zb(const float2* vcol /* &V[0][m] */, cf (&y)[16]) {
    cf f0[32], f1[32], f2[32];
    {
        cf v[96];
#pragma unroll
        for (int r = 0; r < 96; r++) { const float2 t = vcol[r * kZbVStride]; v[r].r = t.x; v[r].i = t.y; }
        IdftPow2<32, 3>::run(v, f0);
        IdftPow2<32, 3>::run(v + 1, f1);
        IdftPow2<32, 3>::run(v + 2, f2);
    }
    pfb_dft96_combine<0>(f0, f1, f2, y);
}

template <int C>
SNRX_HD void pfb_dft96_combine(const cf* f0, const cf* f1, const cf* f2, cf (&y)[16]) {
    constexpr int k = zb_bin_of_slot(C), k1 = k % 32;
    const cf b = twmul<k, 96>(f1[k1]);
    const cf c = twmul<2 * k, 96>(f2[k1]);
    y[C].r = f_add(f_add(f0[k1].r, b.r), c.r);
    y[C].i = f_add(f_add(f0[k1].i, b.i), c.i);
    if constexpr (C + 1 < 16) pfb_dft96_combine<C + 1>(f0, f1, f2, y);
}

// discriminator of slot C from un-rotated y[m] (cur) and y[m+1] (nxt)
template <int C, class Tab>
SNRX_HD float zb_disc_g(const cf& cur, const cf& nxt, Tab tab) {
    // d = nxt * conj(cur), then times (-j)^k
    const float re = f_add(f_mul(nxt.r, cur.r), f_mul(nxt.i, cur.i));
    const float im = f_sub(f_mul(nxt.i, cur.r), f_mul(nxt.r, cur.i));
    constexpr int rot = zb_bin_of_slot(C) & 3;
    const float rr = rot == 0 ? re : rot == 1 ? im : rot == 2 ? -re : -im;
    const float ii = rot == 0 ? im : rot == 1 ? -re : rot == 2 ? -im : re;
    return tab_atan2_g(ii, rr, tab);
}
template <int C>
SNRX_HD float zb_disc(const cf& cur, const cf& nxt, const float* tab) { return zb_disc_g<C>(cur, nxt, AtanTab{tab}); }

#if defined(__CUDACC__)
template <int C, bool EDGE = true, class Tab = AtanTab>
__device__ __forceinline__ void zb_disc_all(const cf (&y)[16], const float2* next_warp_first, int lane, Tab tab, float (&out)[16]);

struct PfbZbArgs {
    const float2* x; uint64_t stride; int64_t n_in; int32_t n_out; int32_t n_tiles; int32_t tile0;
    const float* taps_rho;     // [24][NT]
    float* f; size_t f_stride; // [cap][16][f_stride]
    const float* atan_tab;
    float2* dbg_cf;            // [cap][16][n_out] rotated channel streams, or null
};

template <int NT, bool DEBUG>
__global__ void __launch_bounds__(kZbFirThreads, 1) k_pfb_zb(PfbZbArgs a) {
    using G = PfbGeom<NT, kZbChunkT>;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2* xs = reinterpret_cast<float2*>(smem_raw);
    float2* V = xs + G::kXsLen;                                        // [96][kZbVStride]
    float2* edge = V + 96 * kZbVStride;                                // [5][16]
    float* tab = reinterpret_cast<float*>(edge + 5 * 16);              // [257]

    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int tile = a.tile0 + (int)(blockIdx.x % a.n_tiles);
    const int cap = blockIdx.x / a.n_tiles;
    const float2* xcap = a.x + (size_t)cap * a.stride;
    const int g_first = kTileStride * tile;                            // tiles advance by 127 samples (see pfb.cuh)

    pfb_stage_tile<G, kZbChunkT, kZbFirThreads>(xs, xcap, (int64_t)kPfbD * g_first - G::kHist, a.n_in, tid);
    for (int i = tid; i < 257; i += kZbFirThreads) tab[i] = a.atan_tab[i];
    // 12 warps: warp = (rho group of 8) x (4 chunks of 8 output times)
    const int rho = 8 * (wid % 3) + (lane & 7);
    const int q = 4 * (wid / 3) + (lane >> 3);                         // 0..15
    float g[NT];
#pragma unroll
    for (int d = 0; d < NT; d++) g[d] = __ldg(a.taps_rho + rho * NT + d);
    cp_async_commit_wait_all();
    __syncthreads();

    {
        float2 acc[4][kZbChunkT];
        pfb_fir_thread<NT, 4, kZbChunkT>(xs + fir_base<NT, kZbChunkT>(rho, q), rho <= 12 ? 8 : 0, g, acc);
#pragma unroll
        for (int br = 0; br < 4; br++)
#pragma unroll
            for (int e = 0; e < kZbChunkT; e++) V[(rho + 24 * br) * kZbVStride + kZbChunkT * q + e] = acc[br][e];
    }
    __syncthreads();

    if (wid < 4) {
        cf y[16];
        const int m = tid;                                              // 0..127
        const int mg = g_first + m;
        pfb_dft96_zb(V + m, y);
        if (DEBUG && a.dbg_cf && m < kTileStride && mg < a.n_out) {
#pragma unroll
            for (int c = 0; c < 16; c++) {
                const int rot = (zb_bin_of_slot(c) * (mg & 3)) & 3;
                const float rr = rot == 0 ? y[c].r : rot == 1 ? y[c].i : rot == 2 ? -y[c].r : -y[c].i;
                const float ii = rot == 0 ? y[c].i : rot == 1 ? -y[c].r : rot == 2 ? -y[c].i : y[c].r;
                a.dbg_cf[((size_t)cap * 16 + c) * (size_t)a.n_out + mg] = make_float2(rr, ii);
            }
        }
        if (lane == 0) {
#pragma unroll
            for (int c = 0; c < 16; c++) edge[wid * 16 + c] = make_float2(y[c].r, y[c].i);
        }
        asm volatile("bar.sync 1, 128;\n" ::);
        float out[16];
        zb_disc_all<0>(y, edge + ((wid + 1) & 3) * 16, lane, AtanTab{tab}, out);
        const int n = mg + 1;                                           // f[n] pairs y[n] with y[n-1]
        if (m < kTileStride && n < a.n_out) {
#pragma unroll
            for (int c = 0; c < 16; c++) a.f[((size_t)cap * 16 + c) * a.f_stride + n] = out[c];
        }
        if (mg == 0) {
#pragma unroll
            for (int c = 0; c < 16; c++) a.f[((size_t)cap * 16 + c) * a.f_stride] = 0.0f;   // x[-1] = 0 -> atan2(0,0)
        }
    }
}

// EDGE: lane 31's successor is the first sample of the next warp (k_pfb_zb); the one-warp tile never stores lane 31's value
// (its successor belongs to the next tile), so there the fix-up -- a predicated address computation and shared-memory load
// per channel -- is left out
template <int C, bool EDGE, class Tab>
__device__ __forceinline__ void zb_disc_all(const cf (&y)[16], const float2* next_warp_first, int lane, Tab tab, float (&out)[16]) {
    cf nxt;
    if (EDGE) {
        nxt.r = __shfl_down_sync(0xffffffffu, y[C].r, 1);
        nxt.i = __shfl_down_sync(0xffffffffu, y[C].i, 1);
        if (lane == 31) { const float2 t = next_warp_first[C]; nxt.r = t.x; nxt.i = t.y; }
    } else {
        // lane 31 pairs with lane 0 (any ordinary sample will do: its value is dropped).  With shfl_down it would pair with
        // ITSELF: y conj(y) has a zero imaginary part, 0 / |y|^2 sends the IEEE division of tab_atan2 down its slow path, and the
        // whole warp waits for one lane's subroutine call in every channel of every tile (measured: the front end 24 % slower)
        nxt.r = __shfl_sync(0xffffffffu, y[C].r, (lane + 1) & 31);
        nxt.i = __shfl_sync(0xffffffffu, y[C].i, (lane + 1) & 31);
    }
    out[C] = zb_disc_g<C>(y[C], nxt, tab);
    if constexpr (C + 1 < 16) zb_disc_all<C + 1, EDGE, Tab>(y, next_warp_first, lane, tab, out);
}

// ---------------------------------------------------------------------------------------------------
// One-warp tiles (same shape as k_pfb_ble, pfb.cuh): a warp computes 32 channel-rate samples of all 16
// channels end to end and writes the 31 discriminator values whose predecessor it holds; tiles advance by 31
// samples, so they are independent -- no CTA-wide barrier, 17 KB of shared memory, ~12 warps per SM.
//   phase 0  bulk (TMA) staging of the skewed input tile;
//   phase 1  three passes gi = 0..2: FIR of the 32 branches r = gi + 3 r2 (lane (rl, chunk) owns rho = gi + 3 rl,
//            i.e. r2 = rl + 8 a, a = 0..3, and 8 output times), an 8 KB transpose through shared memory,
//            one 32-point inverse DFT per lane (= output time), of which only the 16 bins that carry a channel
//            are ever used: the radix-3 combination y_k += W96^(k gi) F_gi[k mod 32] is accumulated pass by pass;
//   phase 2  successor sample by warp shuffle, cross product, table atan2, coalesced store of f.
// Arithmetic per output is the same sequence of operations as k_pfb_zb (pfb_fir_thread, IdftPow2<32>,
// f0 + W f1 + W^2 f2), so tests/emu's statement of the tile covers both.
template <int NT> struct PfbZbWarpGeom {
    static constexpr int kT = 32, kStride = 31, kThreads = 32;
    using G = PfbGeom<NT, kChunkT, kT>;
    static constexpr int kXsBytes = ((G::kXsLen * 8 + 15) / 16) * 16;
    static constexpr int kVBytes = 2 * 8 * 32 * 16;                    // two float4 planes [8][32]: (a0, a1) and (a2, a3)
    static constexpr int kSmemBytes = kXsBytes + kVBytes + 16;
    static constexpr int kCtasPerSm = (228 * 1024) / (kSmemBytes + 1024) < 12 ? (228 * 1024) / (kSmemBytes + 1024) : 12;
};

template <int GI, int C>
SNRX_HD void pfb_zb_accumulate(const cf (&f)[32], cf (&y)[16]) {
    constexpr int k = zb_bin_of_slot(C), k1 = k % 32;
    if constexpr (GI == 0) {
        y[C] = f[k1];
    } else {
        const cf t = twmul<GI * k, 96>(f[k1]);
        y[C].r = f_add(y[C].r, t.r);
        y[C].i = f_add(y[C].i, t.i);
    }
    if constexpr (C + 1 < 16) pfb_zb_accumulate<GI, C + 1>(f, y);
}

struct PfbZbWarpArgs {
    const float2* x; uint64_t stride; int64_t n_in; int32_t n_out; int32_t n_tiles;
    const float4* taps_pass;   // [3][NT/4][8] float4, as PfbBleArgs::taps_pass
    float* f; size_t f_stride; // [cap][16][f_stride]
    const float2* atan_pairs;  // [256] AtanTabPairs, global (L1 resident)
    float2* dbg_cf;            // [cap][16][n_out] rotated channel streams, or null
};

#ifndef SNRX_PFB_ZB_UNROLL
#define SNRX_PFB_ZB_UNROLL 1          // 1 = the pass loop stays a loop; 3 = unrolled (round 2's first version)
#endif
constexpr int kZbPassUnroll = SNRX_PFB_ZB_UNROLL;
template <int NT, bool DEBUG>
__global__ void __launch_bounds__(32, PfbZbWarpGeom<NT>::kCtasPerSm) k_pfb_zb_warp(PfbZbWarpArgs a) {
    using B = PfbZbWarpGeom<NT>;
    using G = typename B::G;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2* xs = reinterpret_cast<float2*>(smem_raw);
    float4* V = reinterpret_cast<float4*>(smem_raw + B::kXsBytes);
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw + B::kXsBytes + B::kVBytes);

    const int lane = threadIdx.x;
    const int tile = (int)blockIdx.x;                      // grid = (tiles, captures): no division ahead of the bulk copies
    const int cap = (int)blockIdx.y;
    const float2* xcap = a.x + (size_t)cap * a.stride;
    const int g_first = B::kStride * tile;
    {
        const int64_t x0 = (int64_t)kPfbD * g_first - G::kHist;
        if (x0 >= 0 && x0 + G::kTileIn <= a.n_in) {
            pfb_stage_tile_bulk<G, kChunkT, true>(xs, xcap + x0, bar, lane);     // the elected lane initialises the barrier too
            __syncwarp();
            mbar_wait(bar, 0);
        } else {
            pfb_stage_tile<G, kChunkT, B::kThreads>(xs, xcap, x0, a.n_in, lane);
            cp_async_commit_wait_all();
            __syncwarp();
        }
    }

    cf y[16];
    {
        const int rl = lane & 7, c = lane >> 3;
        // The three passes are ONE loop body (not unrolled): FIR and 32-point transform are the same code for every gi, only
        // the 16 twiddles of the accumulation differ.  Unrolled, the kernel is 2656 instructions = 42 KB of straight-line code
        // that every one-warp CTA runs through exactly once -- more than the 32 KB instruction cache of the SM, and ncu showed
        // 11 % of the warp time as no_instruction; rolled it is a third of that.  Same operations in the same order.
#pragma unroll kZbPassUnroll
        for (int gi = 0; gi < 3; gi++) {
            const int rho = gi + 3 * rl;
            float g[NT];
            const float4* gp = a.taps_pass + gi * (NT / 4) * 8 + rl;
#pragma unroll
            for (int d = 0; d < NT / 4; d++) {
                const float4 t = __ldg(gp + 8 * d);
                g[4 * d] = t.x; g[4 * d + 1] = t.y; g[4 * d + 2] = t.z; g[4 * d + 3] = t.w;
            }
            {
                float2 acc[4][kChunkT];
                pfb_fir_thread<NT, 4, kChunkT>(xs + fir_base<NT, kChunkT>(rho, c), rho <= 12 ? 8 : 0, g, acc);
#pragma unroll
                for (int e = 0; e < kChunkT; e++) {
                    V[v_pos(rl, 8 * c + e)] = make_float4(acc[0][e].x, acc[0][e].y, acc[1][e].x, acc[1][e].y);
                    V[256 + v_pos(rl, 8 * c + e)] = make_float4(acc[2][e].x, acc[2][e].y, acc[3][e].x, acc[3][e].y);
                }
            }
            __syncwarp();
            cf v32[32], f[32];
#pragma unroll
            for (int r = 0; r < 8; r++) {
                const float4 t0 = V[v_pos(r, lane)], t1 = V[256 + v_pos(r, lane)];
                v32[r].r = t0.x; v32[r].i = t0.y; v32[r + 8].r = t0.z; v32[r + 8].i = t0.w;
                v32[r + 16].r = t1.x; v32[r + 16].i = t1.y; v32[r + 24].r = t1.z; v32[r + 24].i = t1.w;
            }
            __syncwarp();                                                  // V is rewritten by the next pass
            IdftPow2<32, 1>::run(v32, f);
            if (gi == 0) pfb_zb_accumulate<0, 0>(f, y);
            else if (gi == 1) pfb_zb_accumulate<1, 0>(f, y);
            else pfb_zb_accumulate<2, 0>(f, y);
        }
    }

    const int mg = g_first + lane;
    // (lane 31's sample belongs to the next tile -- except in the capture's last tile, which has no successor)
    if (DEBUG && a.dbg_cf && (lane < B::kStride || tile == a.n_tiles - 1) && mg < a.n_out) {
#pragma unroll
        for (int c = 0; c < 16; c++) {
            const int rot = (zb_bin_of_slot(c) * (mg & 3)) & 3;
            const float rr = rot == 0 ? y[c].r : rot == 1 ? y[c].i : rot == 2 ? -y[c].r : -y[c].i;
            const float ii = rot == 0 ? y[c].i : rot == 1 ? -y[c].r : rot == 2 ? -y[c].i : y[c].r;
            a.dbg_cf[((size_t)cap * 16 + c) * (size_t)a.n_out + mg] = make_float2(rr, ii);
        }
    }
    float out[16];
    zb_disc_all<0, false>(y, nullptr /* lane 31's successor is not in the tile: its value is never stored */, lane, AtanTabPairs{a.atan_pairs}, out);
    const int n = mg + 1;                                                   // f[n] pairs y[n] with y[n-1]
    if (lane < B::kStride && n < a.n_out) {
#pragma unroll
        for (int c = 0; c < 16; c++) a.f[((size_t)cap * 16 + c) * a.f_stride + n] = out[c];
    }
    if (mg == 0) {
#pragma unroll
        for (int c = 0; c < 16; c++) a.f[((size_t)cap * 16 + c) * a.f_stride] = 0.0f;       // x[-1] = 0 -> atan2(0,0)
    }
}

// ------------------------------------------------------------------------------------ host side
inline int zb_wideband_init(ZbState& s, const snrx_config_t& cfg, const double* proto, uint32_t max_caps, uint32_t max_out,
                            std::string& err) {
    const int L = (int)cfg.pfb_taps, NT = L / 24;
    s.wb_nt = NT;
    { const char* e = getenv("SNRX_ZB_PFB"); s.wb_cta_kernel = e && std::string(e) == "cta"; }   // A/B switch: the 384-thread tile kernel
    std::vector<float> flat(L), rho(L);
    for (int n = 0; n < L; n++) flat[n] = (float)proto[n];
    for (int r = 0; r < 24; r++) for (int d = 0; d < NT; d++) rho[r * NT + d] = flat[r + 24 * d];
    ZCK(cudaMalloc((void**)&s.d_wb_taps_rho, sizeof(float) * L));
    ZCK(cudaMalloc((void**)&s.d_wb_taps_flat, sizeof(float) * L));
    ZCK(cudaMemcpy(s.d_wb_taps_rho, rho.data(), sizeof(float) * L, cudaMemcpyHostToDevice));
    ZCK(cudaMemcpy(s.d_wb_taps_flat, flat.data(), sizeof(float) * L, cudaMemcpyHostToDevice));
    {
        std::vector<float> pass(L);
        for (int gi = 0; gi < 3; gi++) for (int d4 = 0; d4 < NT / 4; d4++) for (int rl = 0; rl < 8; rl++) for (int k = 0; k < 4; k++)
            pass[((gi * (NT / 4) + d4) * 8 + rl) * 4 + k] = flat[(gi + 3 * rl) + 24 * (4 * d4 + k)];
        ZCK(cudaMalloc((void**)&s.d_wb_taps_pass, sizeof(float) * L));
        ZCK(cudaMemcpy(s.d_wb_taps_pass, pass.data(), sizeof(float) * L, cudaMemcpyHostToDevice));
    }
    ZCK(cudaFuncSetAttribute(k_pfb_zb_warp<16, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, PfbZbWarpGeom<16>::kSmemBytes));
    ZCK(cudaFuncSetAttribute(k_pfb_zb_warp<16, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, PfbZbWarpGeom<16>::kSmemBytes));
    ZCK(cudaFuncSetAttribute(k_pfb_zb_warp<32, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, PfbZbWarpGeom<32>::kSmemBytes));
    ZCK(cudaFuncSetAttribute(k_pfb_zb_warp<32, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, PfbZbWarpGeom<32>::kSmemBytes));
    ZCK(cudaFuncSetAttribute(k_pfb_zb<16, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, PfbZbGeom<16>::kSmemBytes));
    ZCK(cudaFuncSetAttribute(k_pfb_zb<16, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, PfbZbGeom<16>::kSmemBytes));
    ZCK(cudaFuncSetAttribute(k_pfb_zb<32, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, PfbZbGeom<32>::kSmemBytes));
    ZCK(cudaFuncSetAttribute(k_pfb_zb<32, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, PfbZbGeom<32>::kSmemBytes));
    if (cfg.flags & SNRX_F_KEEP_STREAMS)
        ZCK(cudaMalloc((void**)&s.d_wb_cf, sizeof(float2) * (size_t)max_caps * 16 * max_out));
    return SNRX_OK;
}

inline int zb_wideband_front(ZbState& s, const snrx_config_t& cfg, const float2* x, uint32_t n_captures, uint64_t n_samples,
                             uint64_t stride, uint32_t n_out, cudaStream_t st, int& launches, std::string& err) {
    const bool dbg = (cfg.flags & SNRX_F_KEEP_STREAMS) != 0;
    if (!s.wb_cta_kernel) {
        PfbZbWarpArgs a;
        a.x = x; a.stride = stride; a.n_in = (int64_t)n_samples; a.n_out = (int32_t)n_out;
        a.n_tiles = (int32_t)std::max<uint32_t>(1u, (n_out - 1 + 30) / 31);
        a.taps_pass = reinterpret_cast<const float4*>(s.d_wb_taps_pass);
        a.f = s.d_f; a.f_stride = s.stride; a.atan_pairs = s.d_atan_pairs; a.dbg_cf = s.d_wb_cf;
        const dim3 grid((unsigned)a.n_tiles, n_captures);
        if (s.wb_nt == 16) {
            if (dbg) k_pfb_zb_warp<16, true><<<grid, 32, PfbZbWarpGeom<16>::kSmemBytes, st>>>(a);
            else k_pfb_zb_warp<16, false><<<grid, 32, PfbZbWarpGeom<16>::kSmemBytes, st>>>(a);
        } else {
            if (dbg) k_pfb_zb_warp<32, true><<<grid, 32, PfbZbWarpGeom<32>::kSmemBytes, st>>>(a);
            else k_pfb_zb_warp<32, false><<<grid, 32, PfbZbWarpGeom<32>::kSmemBytes, st>>>(a);
        }
        launches++;
        ZCK(cudaGetLastError());
        return SNRX_OK;
    }
    PfbZbArgs a;
    a.x = x; a.stride = stride; a.n_in = (int64_t)n_samples; a.n_out = (int32_t)n_out;
    a.n_tiles = (int32_t)std::max<uint32_t>(1u, (n_out - 1 + kTileStride - 1) / kTileStride); a.tile0 = 0;
    a.taps_rho = s.d_wb_taps_rho;
    a.f = s.d_f; a.f_stride = s.stride; a.atan_tab = s.d_atan; a.dbg_cf = s.d_wb_cf;
    const dim3 grid((unsigned)a.n_tiles * n_captures);
    if (s.wb_nt == 16) {
        if (dbg) k_pfb_zb<16, true><<<grid, kZbFirThreads, PfbZbGeom<16>::kSmemBytes, st>>>(a);
        else k_pfb_zb<16, false><<<grid, kZbFirThreads, PfbZbGeom<16>::kSmemBytes, st>>>(a);
    } else {
        if (dbg) k_pfb_zb<32, true><<<grid, kZbFirThreads, PfbZbGeom<32>::kSmemBytes, st>>>(a);
        else k_pfb_zb<32, false><<<grid, kZbFirThreads, PfbZbGeom<32>::kSmemBytes, st>>>(a);
    }
    launches++;
    ZCK(cudaGetLastError());
    return SNRX_OK;
}
#endif  // __CUDACC__

}  // namespace snrx
