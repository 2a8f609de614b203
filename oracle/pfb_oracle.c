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

/* float polyphase + FFT channelizer; hf = prototype taps as float, L multiple of 96 */
void pfb_oracle_fast(const float* x, int64_t n_in, const float* hf, int L,
                     const int* bins, int nb, int64_t m0, int64_t m1, float* out) {
    const int64_t n_out = m1 - m0;
    const int P = L / PFB_M;
    idft96_plan plan;
    idft96_init(&plan);
#pragma omp parallel for schedule(static)
    for (int64_t m = m0; m < m1; m++) {
        cf v[PFB_M], Y[PFB_M];
        for (int r = 0; r < PFB_M; r++) {
            float ar = 0.0f, ai = 0.0f;
            for (int p = 0; p < P; p++) {
                int64_t i = (int64_t)PFB_D * m - r - (int64_t)PFB_M * p;
                if (i < 0 || i >= n_in) continue;
                float c = hf[r + PFB_M * p];
                ar += c * x[2 * i];
                ai += c * x[2 * i + 1];
            }
            v[r].r = ar; v[r].i = ai;
        }
        idft96(&plan, v, Y);
        for (int b = 0; b < nb; b++) {
            const int k = bins[b];
            cf y = Y[k];
            int rot = (int)(((int64_t)k * (m & 3)) & 3);
            float yr, yi;
            switch (rot) {
                case 0: yr = y.r; yi = y.i; break;
                case 1: yr = y.i; yi = -y.r; break;
                case 2: yr = -y.r; yi = -y.i; break;
                default: yr = -y.i; yi = y.r; break;
            }
            out[2 * (b * n_out + (m - m0))] = yr;
            out[2 * (b * n_out + (m - m0)) + 1] = yi;
        }
    }
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
