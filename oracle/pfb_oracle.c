/* TEST INFRASTRUCTURE ONLY (oracle) -- never linked into, imported by or
 * executed from the product path.
 *
 * CPU statement of the wideband channelizer.  The reference has NO channelizer
 * (Snout retunes one channel at a time: snout/util/btle.py:62,
 * snout/core/radio.py:415), so there is no reference code to follow here; the
 * definition below IS the contract (DESIGN.md "Channelizer"):
 *
 *   y_k[m] = (-j)^(k m) * sum_{n=0}^{L-1} h[n] * exp(+j 2 pi k n / M) * x[D m - n]
 *
 * i.e. mix bin k (k/M cycles per input sample) to DC, low-pass with the real
 * prototype h (L taps), keep every D-th sample; M = 96, D = 24, x[i] = 0 outside
 * the buffer.
 *
 *   pfb_oracle_direct  evaluates the definition term by term in double precision
 *                      (slow, used by the tolerance tests);
 *   pfb_oracle_fast    is the usual polyphase + FFT factorisation in float with
 *                      OpenMP over time, i.e. what a reasonable CPU implementation
 *                      would run; it is the wideband CPU baseline of bench.py and
 *                      is itself checked against pfb_oracle_direct.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define PFB_M 96
#define PFB_D 24

/* out[b * n_out + (m - m0)] for m in [m0, m1), interleaved cf32 */
void pfb_oracle_direct(const float* x, int64_t n_in, const double* h, int L,
                       const int* bins, int nb, int64_t m0, int64_t m1, float* out) {
    const int64_t n_out = m1 - m0;
    double cw[PFB_M], sw[PFB_M];
    for (int i = 0; i < PFB_M; i++) { cw[i] = cos(2.0 * M_PI * i / PFB_M); sw[i] = sin(2.0 * M_PI * i / PFB_M); }
#pragma omp parallel for schedule(static)
    for (int64_t m = m0; m < m1; m++) {
        for (int b = 0; b < nb; b++) {
            const int k = bins[b];
            double ar = 0.0, ai = 0.0;
            for (int n = 0; n < L; n++) {
                int64_t i = (int64_t)PFB_D * m - n;
                if (i < 0 || i >= n_in) continue;
                int ph = (int)(((int64_t)k * n) % PFB_M);
                double xr = x[2 * i], xi = x[2 * i + 1];
                double wr = h[n] * cw[ph], wi = h[n] * sw[ph];
                ar += wr * xr - wi * xi;
                ai += wr * xi + wi * xr;
            }
            int rot = (int)(((int64_t)k * (m & 3)) & 3);       /* (-j)^(k m) */
            double yr, yi;
            switch (rot) {
                case 0: yr = ar; yi = ai; break;
                case 1: yr = ai; yi = -ar; break;
                case 2: yr = -ar; yi = -ai; break;
                default: yr = -ai; yi = ar; break;
            }
            out[2 * (b * n_out + (m - m0))] = (float)yr;
            out[2 * (b * n_out + (m - m0)) + 1] = (float)yi;
        }
    }
}

/* ---- small mixed-radix inverse DFT (sign +) for N = 96 = 3 * 32 ------------- */
typedef struct { float r, i; } cf;

static void idft_pow2(cf* a, int n, const cf* tw /* exp(+j2pi k/n), k<n/2 */) {
    /* iterative radix-2 DIT, n power of two */
    for (int i = 1, j = 0; i < n; i++) {
        int bit = n >> 1;
        for (; j & bit; bit >>= 1) j ^= bit;
        j ^= bit;
        if (i < j) { cf t = a[i]; a[i] = a[j]; a[j] = t; }
    }
    for (int len = 2; len <= n; len <<= 1) {
        int step = n / len;
        for (int i = 0; i < n; i += len)
            for (int k = 0; k < len / 2; k++) {
                cf w = tw[k * step];
                cf u = a[i + k], v = a[i + k + len / 2];
                cf t = { v.r * w.r - v.i * w.i, v.r * w.i + v.i * w.r };
                a[i + k].r = u.r + t.r; a[i + k].i = u.i + t.i;
                a[i + k + len / 2].r = u.r - t.r; a[i + k + len / 2].i = u.i - t.i;
            }
    }
}

typedef struct { cf tw32[16]; cf tw96[PFB_M]; } idft96_plan;

static void idft96_init(idft96_plan* p) {
    for (int k = 0; k < 16; k++) { p->tw32[k].r = (float)cos(2.0 * M_PI * k / 32); p->tw32[k].i = (float)sin(2.0 * M_PI * k / 32); }
    for (int k = 0; k < PFB_M; k++) { p->tw96[k].r = (float)cos(2.0 * M_PI * k / PFB_M); p->tw96[k].i = (float)sin(2.0 * M_PI * k / PFB_M); }
}

/* Y[k] = sum_r v[r] exp(+j 2 pi k r / 96): r = 3 r1 + r2, three 32-point transforms + radix-3 */
static void idft96(const idft96_plan* p, const cf* v, cf* Y) {
    cf sub[3][32];
    for (int r2 = 0; r2 < 3; r2++) {
        for (int r1 = 0; r1 < 32; r1++) sub[r2][r1] = v[3 * r1 + r2];
        idft_pow2(sub[r2], 32, p->tw32);
    }
    for (int k = 0; k < PFB_M; k++) {
        int k1 = k & 31;
        cf acc = sub[0][k1];
        for (int r2 = 1; r2 < 3; r2++) {
            cf w = p->tw96[(k * r2) % PFB_M];
            cf s = sub[r2][k1];
            acc.r += s.r * w.r - s.i * w.i;
            acc.i += s.r * w.i + s.i * w.r;
        }
        Y[k] = acc;
    }
}

/* ---- the fast CPU statement: what a careful CPU implementation of the same channelizer runs --------------------------
 * float polyphase FIR + pruned mixed-radix transform, OpenMP over blocks of 16 output times:
 *   FIR    per output time and tap block p the 96 branches read 96 CONSECUTIVE input samples (backwards), so with the taps
 *          stored reversed and duplicated (re, im) the inner loop is one element-wise multiply-add over 192 floats;
 *   DFT    16 output times are transformed at once, structure-of-arrays, so every butterfly is a 16-lane vector
 *          operation; when only even bins are wanted (the 40 BLE channels) the input is folded to a 48-point transform
 *          (3 x 16), else 96 points (3 x 32); the radix-3 combination is evaluated for the wanted bins only.
 * This is the wideband CPU baseline of bench.py (cpu_baseline / --impl reference); it is checked against
 * pfb_oracle_direct in tests/test_oracle_pfb.py. */
#define PFB_LANES 16
typedef float lane_t[PFB_LANES];

static void soa_idft_pow2(lane_t* ar, lane_t* ai, int n, const cf* tw /* exp(+j 2 pi k / n), k < n/2 */) {
    for (int i = 1, j = 0; i < n; i++) {
        int bit = n >> 1;
        for (; j & bit; bit >>= 1) j ^= bit;
        j ^= bit;
        if (i < j) {
            lane_t t;
            memcpy(t, ar[i], sizeof t); memcpy(ar[i], ar[j], sizeof t); memcpy(ar[j], t, sizeof t);
            memcpy(t, ai[i], sizeof t); memcpy(ai[i], ai[j], sizeof t); memcpy(ai[j], t, sizeof t);
        }
    }
    for (int len = 2; len <= n; len <<= 1) {
        const int step = n / len, half = len / 2;
        for (int i = 0; i < n; i += len)
            for (int k = 0; k < half; k++) {
                const float wr = tw[k * step].r, wi = tw[k * step].i;
                float* ur = ar[i + k]; float* ui = ai[i + k];
                float* vr = ar[i + k + half]; float* vi = ai[i + k + half];
#pragma omp simd
                for (int l = 0; l < PFB_LANES; l++) {
                    const float tr = vr[l] * wr - vi[l] * wi, ti = vr[l] * wi + vi[l] * wr;
                    vr[l] = ur[l] - tr; vi[l] = ui[l] - ti;
                    ur[l] += tr; ui[l] += ti;
                }
            }
    }
}

void pfb_oracle_fast(const float* x, int64_t n_in, const float* hf, int L,
                     const int* bins, int nb, int64_t m0, int64_t m1, float* out) {
    const int64_t n_out = m1 - m0;
    const int P = L / PFB_M;
    int all_even = 1;
    for (int b = 0; b < nb; b++) if (bins[b] & 1) all_even = 0;
    const int N = all_even ? PFB_M / 2 : PFB_M, Q = N / 3;          /* transform length, sub-transform length */
    /* taps reversed inside each block of 96 and duplicated for (re, im): h2[p][2 s + c] = h[96 p + 95 - s] */
    float* h2 = (float*)malloc(sizeof(float) * 2 * (size_t)L);
    for (int p = 0; p < P; p++)
        for (int s2 = 0; s2 < PFB_M; s2++) h2[2 * (PFB_M * p + s2)] = h2[2 * (PFB_M * p + s2) + 1] = hf[PFB_M * p + (PFB_M - 1 - s2)];
    cf twq[16], *tww = (cf*)malloc(sizeof(cf) * 3 * (size_t)nb);     /* sub-transform twiddles; radix-3 factors per wanted bin */
    for (int k = 0; k < Q / 2; k++) { twq[k].r = (float)cos(2.0 * M_PI * k / Q); twq[k].i = (float)sin(2.0 * M_PI * k / Q); }
    for (int b = 0; b < nb; b++) {
        const int kk = all_even ? bins[b] / 2 : bins[b];
        for (int r2 = 0; r2 < 3; r2++) {
            tww[3 * b + r2].r = (float)cos(2.0 * M_PI * ((kk * r2) % N) / N);
            tww[3 * b + r2].i = (float)sin(2.0 * M_PI * ((kk * r2) % N) / N);
        }
    }
#pragma omp parallel
    {
        lane_t* sr = (lane_t*)aligned_alloc(64, sizeof(lane_t) * PFB_M);   /* sub[r2][r1] = a[3 r1 + r2], [r2 * Q + r1] */
        lane_t* si = (lane_t*)aligned_alloc(64, sizeof(lane_t) * PFB_M);
        float v[2 * PFB_M] __attribute__((aligned(64)));
#pragma omp for schedule(static)
        for (int64_t blk = m0; blk < m1; blk += PFB_LANES) {
            const int nl = (int)(m1 - blk < PFB_LANES ? m1 - blk : PFB_LANES);
            for (int l = 0; l < PFB_LANES; l++) {
                const int64_t m = blk + l;
                memset(v, 0, sizeof v);
                if (l < nl) {
                    const int64_t hi = (int64_t)PFB_D * m, lo = hi - (int64_t)PFB_M * P + 1;
                    if (lo >= 0 && hi < n_in) {
                        for (int p = 0; p < P; p++) {
                            const float* xf = x + 2 * (hi - (int64_t)PFB_M * p - (PFB_M - 1));
                            const float* hp = h2 + 2 * PFB_M * p;
#pragma omp simd
                            for (int i = 0; i < 2 * PFB_M; i++) v[i] += hp[i] * xf[i];
                        }
                    } else {
                        for (int p = 0; p < P; p++)
                            for (int s2 = 0; s2 < PFB_M; s2++) {
                                const int64_t i = hi - (int64_t)PFB_M * p - (PFB_M - 1) + s2;
                                if (i < 0 || i >= n_in) continue;
                                v[2 * s2] += h2[2 * (PFB_M * p + s2)] * x[2 * i];
                                v[2 * s2 + 1] += h2[2 * (PFB_M * p + s2)] * x[2 * i + 1];
                            }
                    }
                }
                /* v[2 s] holds branch r = 95 - s; fold r and r + 48 when only even bins are wanted; sort into sub-transforms */
                for (int r = 0; r < N; r++) {
                    float re = v[2 * (PFB_M - 1 - r)], im = v[2 * (PFB_M - 1 - r) + 1];
                    if (all_even) { re += v[2 * (PFB_M - 1 - r - N)]; im += v[2 * (PFB_M - 1 - r - N) + 1]; }
                    sr[(r % 3) * Q + r / 3][l] = re;
                    si[(r % 3) * Q + r / 3][l] = im;
                }
            }
            for (int r2 = 0; r2 < 3; r2++) soa_idft_pow2(sr + r2 * Q, si + r2 * Q, Q, twq);
            for (int b = 0; b < nb; b++) {
                const int k = bins[b], kk = all_even ? k / 2 : k, k1 = kk % Q;
                const cf w1 = tww[3 * b + 1], w2 = tww[3 * b + 2];
                float yr[PFB_LANES], yi[PFB_LANES];
#pragma omp simd
                for (int l = 0; l < PFB_LANES; l++) {
                    yr[l] = sr[k1][l] + sr[Q + k1][l] * w1.r - si[Q + k1][l] * w1.i + sr[2 * Q + k1][l] * w2.r - si[2 * Q + k1][l] * w2.i;
                    yi[l] = si[k1][l] + sr[Q + k1][l] * w1.i + si[Q + k1][l] * w1.r + sr[2 * Q + k1][l] * w2.i + si[2 * Q + k1][l] * w2.r;
                }
                float* o = out + 2 * ((int64_t)b * n_out + (blk - m0));
                for (int l = 0; l < nl; l++) {
                    const int rot = (int)(((int64_t)k * ((blk + l) & 3)) & 3);        /* (-j)^(k m) */
                    const float a = yr[l], c = yi[l];
                    o[2 * l] = rot == 0 ? a : rot == 1 ? c : rot == 2 ? -a : -c;
                    o[2 * l + 1] = rot == 0 ? c : rot == 1 ? -a : rot == 2 ? -c : a;
                }
            }
        }
        free(sr); free(si);
    }
    free(h2); free(tww);
}

void pfb_oracle_set_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

int pfb_oracle_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

double pfb_oracle_now(void) {
    struct timespec t;
    clock_gettime(CLOCK_MONOTONIC, &t);
    return (double)t.tv_sec + 1e-9 * (double)t.tv_nsec;
}
