// ble_front.cuh -- narrow-band BLE front end: 4 Msps cf32 of ONE channel -> int8 grid -> slicer bits.
//
// Reference: the HackRF delivers interleaved int8 I,Q (rx_callback, btle_rx.c:489-498) and the
// receiver decides every bit as (I0*Q1 - I1*Q0) > 0 on adjacent samples (btle_rx.c:1357-1361,
// 1385-1392).  Here the capture arrives as cf32; it is put on the int8 grid first,
//   q = clamp(rint(x * scale), -128, 127),
// (products of two such values and their difference are exact in FP32, so the FP32
// cross product below takes the same decision as the reference's `int` arithmetic).
//
// HBM-bound streaming kernel: 8 bytes in, 1 bit out per sample.  One warp turns 128 consecutive
// samples (1 KiB, four coalesced 256-byte loads) into four words of the channel's bit stream (BitsLayout:
// natural sample order) with four __ballot_sync; the successor of a word's last sample comes from lane 0
// of the next word by warp shuffle.
#pragma once
#include "common.cuh"

namespace snrx {

SNRX_HD float quant_exact(float x, float scale) {
    // same value as the oracle's rintf(x*scale) followed by the clamp (ble_oracle_quantize)
    float t = f_add(f_mul(x, scale), 12582912.0f);
    t = fminf(fmaxf(t, 12582912.0f - 128.0f), 12582912.0f + 127.0f);
    return f_sub(t, 12582912.0f);
}

SNRX_HD bool slicer_bit(float i0, float q0, float i1, float q1) {
    return f_fma(i0, q1, -f_mul(i1, q0)) > 0.0f;      // exact: |terms| <= 2^14
}

#if defined(__CUDACC__)
struct NbArgs {
    const float2* x;          // [n_captures][stride] cf32
    uint64_t stride;          // samples between captures
    int64_t n;                // samples per capture
    int32_t n_groups;         // ceil(n / 128)
    uint32_t n_captures;
    uint32_t blocks_per_cap;  // the grid is n_captures * blocks_per_cap blocks: a block stays inside one capture (no 64-bit division per item)
    float scale;
    uint32_t* bits;
    BitsLayout lay;           // n_channels == 1
    int8_t* dbg_q8;           // [cap][n][2] or null
};

template <bool DEBUG>
__global__ void __launch_bounds__(256) k_ble_slice_nb(NbArgs a) {
    const int lane = threadIdx.x & 31;
    const uint32_t warps_per_block = blockDim.x >> 5;
    // one 32-bit division per block; the item loop below has none (ncu of the first version, which split a 64-bit item index
    // per item: 185 instructions per 128 samples, ALU pipe 58 % busy, issue slots 67 % -- not the HBM-bound kernel it should be)
    const uint32_t cap = blockIdx.x / a.blocks_per_cap;
    const float2* xc = a.x + (size_t)cap * a.stride;
    for (int32_t grp = (int32_t)((blockIdx.x % a.blocks_per_cap) * warps_per_block + (threadIdx.x >> 5)); grp < a.n_groups;
         grp += (int32_t)(a.blocks_per_cap * warps_per_block)) {
        const int64_t n0 = (int64_t)grp * 128 + lane;             // this lane's sample of word 0
        float2 t[5];                                              // [k]: sample n0 + 32 k; [4] only matters in lane 0
        if ((int64_t)grp * 128 + 160 <= a.n) {                    // the whole item and its successor sample lie inside the capture
            const float2* p = xc + n0;
#pragma unroll
            for (int k = 0; k < 4; k++) t[k] = __ldcs(p + 32 * k);    // streaming loads: every byte is used once
            t[4] = make_float2(0.f, 0.f);
            if (lane == 0) t[4] = __ldcs(p + 128);
        } else {
#pragma unroll
            for (int k = 0; k < 5; k++) {
                const int64_t n = n0 + 32 * k;
                t[k] = make_float2(0.f, 0.f);                     // beyond the capture: zero fill
                if ((k < 4 || lane == 0) && n < a.n) t[k] = __ldcs(xc + n);
            }
        }
        float I[5], Q[5];
#pragma unroll
        for (int k = 0; k < 5; k++) {
            I[k] = quant_exact(t[k].x, a.scale); Q[k] = quant_exact(t[k].y, a.scale);
            const int64_t n = n0 + 32 * k;
            if (DEBUG && a.dbg_q8 && k < 4 && n < a.n) {
                const size_t o = ((size_t)cap * (size_t)a.n + (size_t)n) * 2;
                a.dbg_q8[o] = (int8_t)I[k]; a.dbg_q8[o + 1] = (int8_t)Q[k];
            }
        }
        uint32_t w[4];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            // successor of sample (word k, lane): (k, lane + 1), for lane 31 (k + 1, 0) -- lane 0 offers its sample of the NEXT
            // word, every other lane its sample of this one: one shuffle per component
            const float i1 = __shfl_sync(0xffffffffu, lane == 0 ? I[k + 1] : I[k], (lane + 1) & 31);
            const float q1 = __shfl_sync(0xffffffffu, lane == 0 ? Q[k + 1] : Q[k], (lane + 1) & 31);
            w[k] = __ballot_sync(0xffffffffu, slicer_bit(I[k], Q[k], i1, q1));
        }
        if (lane < 4) {
            const uint32_t out = lane == 0 ? w[0] : lane == 1 ? w[1] : lane == 2 ? w[2] : w[3];
            a.bits[a.lay.index(cap, 0, kBitsLeadWords + 4 * grp + lane)] = out;
        }
    }
}
#endif

}  // namespace snrx
