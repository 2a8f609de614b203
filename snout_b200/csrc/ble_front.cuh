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
// samples (1 KiB, two 16-byte loads per lane) into the four phase words of one 32-symbol slot
// group with four __ballot_sync.
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
    const float4* x;          // [n_captures][stride] cf32 viewed as float4 (2 samples)
    uint64_t stride;          // samples between captures (even)
    int64_t n;                // samples per capture
    int32_t n_groups;         // ceil(n / 128)
    uint32_t n_captures;
    float scale;
    uint32_t* bits;
    BitsLayout lay;           // n_channels == 1
    int8_t* dbg_q8;           // [cap][n][2] or null
};

template <bool DEBUG>
__global__ void __launch_bounds__(256) k_ble_slice_nb(NbArgs a) {
    const int lane = threadIdx.x & 31;
    const uint32_t warps_per_block = blockDim.x >> 5;
    const uint64_t n_items = (uint64_t)a.n_captures * (uint64_t)a.n_groups;
    for (uint64_t item = (uint64_t)blockIdx.x * warps_per_block + (threadIdx.x >> 5); item < n_items;
         item += (uint64_t)gridDim.x * warps_per_block) {
        const uint32_t cap = (uint32_t)(item / (uint64_t)a.n_groups);
        const int32_t grp = (int32_t)(item % (uint64_t)a.n_groups);
        const float4* xc = a.x + (size_t)cap * (a.stride / 2);
        const int64_t n0 = (int64_t)grp * 128 + 4 * lane;        // first of this lane's 4 samples
        const float2* x2 = reinterpret_cast<const float2*>(xc);
        float4 v0 = make_float4(0.f, 0.f, 0.f, 0.f), v1 = v0;
        if (n0 + 4 <= a.n) {                                      // streaming loads: every byte is used once
            v0 = __ldcs(xc + n0 / 2);
            v1 = __ldcs(xc + n0 / 2 + 1);
        } else {                                                  // ragged end of the capture: zero fill
            float2 t;
            if (n0 + 0 < a.n) { t = __ldg(x2 + n0 + 0); v0.x = t.x; v0.y = t.y; }
            if (n0 + 1 < a.n) { t = __ldg(x2 + n0 + 1); v0.z = t.x; v0.w = t.y; }
            if (n0 + 2 < a.n) { t = __ldg(x2 + n0 + 2); v1.x = t.x; v1.y = t.y; }
        }
        float I[5], Q[5];
        I[0] = quant_exact(v0.x, a.scale); Q[0] = quant_exact(v0.y, a.scale);
        I[1] = quant_exact(v0.z, a.scale); Q[1] = quant_exact(v0.w, a.scale);
        I[2] = quant_exact(v1.x, a.scale); Q[2] = quant_exact(v1.y, a.scale);
        I[3] = quant_exact(v1.z, a.scale); Q[3] = quant_exact(v1.w, a.scale);
        I[4] = __shfl_down_sync(0xffffffffu, I[0], 1);
        Q[4] = __shfl_down_sync(0xffffffffu, Q[0], 1);
        if (lane == 31) {                                         // first sample of the next group
            const int64_t nn = n0 + 4;
            float2 t = make_float2(0.f, 0.f);
            if (nn < a.n) t = __ldg(x2 + nn);
            I[4] = quant_exact(t.x, a.scale); Q[4] = quant_exact(t.y, a.scale);
        }
        if (DEBUG && a.dbg_q8) {
#pragma unroll
            for (int k = 0; k < 4; k++) {
                if (n0 + k < a.n) {
                    const size_t o = ((size_t)cap * (size_t)a.n + (size_t)(n0 + k)) * 2;
                    a.dbg_q8[o] = (int8_t)I[k]; a.dbg_q8[o + 1] = (int8_t)Q[k];
                }
            }
        }
        uint32_t w[4];
#pragma unroll
        for (int j = 0; j < 4; j++) w[j] = __ballot_sync(0xffffffffu, slicer_bit(I[j], Q[j], I[j + 1], Q[j + 1]));
        if (lane < 4) {
            const uint32_t out = lane == 0 ? w[0] : lane == 1 ? w[1] : lane == 2 ? w[2] : w[3];
            a.bits[a.lay.index(cap, 0, lane, kBitsLeadWords + grp)] = out;
        }
    }
}
#endif

}  // namespace snrx
