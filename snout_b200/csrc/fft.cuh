// fft.cuh -- fully unrolled in-register inverse DFTs (kernel exp(+j 2 pi k r / N)) for the
// channelizer: N = 3 * 2^p (48 for the even-bin BLE bank, 96 for the full bank).
//
// Everything is resolved at compile time: the recursion is template recursion, all array
// indices are constants after unrolling (so the arrays live in registers) and the twiddles are
// constexpr values that end up as FFMA immediates.  Radix-2 decimation in time below a single
// radix-3 stage; butterflies use the 6-FMA form  out0 = E + w O, out1 = 2E - out0.
#pragma once
#include "common.cuh"

namespace snrx {

struct cf { float r, i; };

namespace detail {
constexpr double kPi = 3.14159265358979323846264338327950288;

constexpr double cx_sin_small(double x) {   // |x| <= pi/4, Taylor to x^19
    double x2 = x * x, term = x, sum = x;
    for (int k = 1; k <= 9; ++k) { term *= -x2 / ((2 * k) * (2 * k + 1)); sum += term; }
    return sum;
}
constexpr double cx_cos_small(double x) {
    double x2 = x * x, term = 1.0, sum = 1.0;
    for (int k = 1; k <= 10; ++k) { term *= -x2 / ((2 * k - 1) * (2 * k)); sum += term; }
    return sum;
}
// cos / sin of 2*pi*num/den with exact octant reduction
constexpr double cx_cos_frac(int num, int den) {
    num %= den; if (num < 0) num += den;
    // reduce to the first octant using symmetries, working on the fraction num/den of a turn
    // angle = 2 pi num / den
    if (8 * num <= den) return cx_cos_small(2 * kPi * num / den);
    if (4 * num <= den) return cx_sin_small(2 * kPi * (den - 4 * num) / (4.0 * den));      // cos(a) = sin(pi/2 - a)
    if (2 * num <= den) return -cx_cos_frac(den - 2 * num, 2 * den) ;                       // cos(a) = -cos(pi - a)
    return cx_cos_frac(den - num, den);                                                      // cos(a) = cos(2pi - a)
}
constexpr double cx_sin_frac(int num, int den) {
    // sin(a) = cos(a - pi/2) = cos(2 pi (num/den - 1/4)) = cos(2 pi (4 num - den) / (4 den))
    return cx_cos_frac(4 * num - den, 4 * den);
}
}  // namespace detail

template <int NUM, int DEN> struct Tw {
    static constexpr float c = (float)detail::cx_cos_frac(NUM, DEN);
    static constexpr float s = (float)detail::cx_sin_frac(NUM, DEN);
};

// out0 = e + w*o ; out1 = e - w*o   with w = exp(+j 2 pi K / N)
template <int K, int N>
SNRX_HD void bfly(const cf& e, const cf& o, cf& out0, cf& out1) {
    constexpr int k = ((K % N) + N) % N;
    if constexpr (k == 0) {
        out0.r = f_add(e.r, o.r); out0.i = f_add(e.i, o.i);
        out1.r = f_sub(e.r, o.r); out1.i = f_sub(e.i, o.i);
    } else if constexpr (4 * k == N) {          // w = +j
        out0.r = f_sub(e.r, o.i); out0.i = f_add(e.i, o.r);
        out1.r = f_add(e.r, o.i); out1.i = f_sub(e.i, o.r);
    } else {
        constexpr float c = Tw<k, N>::c, s = Tw<k, N>::s;
        float r0 = f_fma(c, o.r, f_fma(-s, o.i, e.r));
        float i0 = f_fma(c, o.i, f_fma(s, o.r, e.i));
        out0.r = r0; out0.i = i0;
        out1.r = f_fma(2.0f, e.r, -r0);
        out1.i = f_fma(2.0f, e.i, -i0);
    }
}

// N-point inverse DFT, N a power of two: in[0], in[S], in[2S], ... -> out[0..N-1] natural order
template <int N, int S>
struct IdftPow2 {
    SNRX_HD static void run(const cf* in, cf* out) {
        cf e[N / 2], o[N / 2];
        IdftPow2<N / 2, 2 * S>::run(in, e);
        IdftPow2<N / 2, 2 * S>::run(in + S, o);
        step<0>(e, o, out);
    }
    template <int K>
    SNRX_HD static void step(const cf* e, const cf* o, cf* out) {
        bfly<K, N>(e[K], o[K], out[K], out[K + N / 2]);
        if constexpr (K + 1 < N / 2) step<K + 1>(e, o, out);
    }
};
template <int S>
struct IdftPow2<1, S> {
    SNRX_HD static void run(const cf* in, cf* out) { out[0] = in[0]; }
};

// complex multiply by exp(+j 2 pi K / N)
template <int K, int N>
SNRX_HD cf twmul(const cf& x) {
    constexpr int k = ((K % N) + N) % N;
    if constexpr (k == 0) {
        return x;
    } else if constexpr (4 * k == N) {
        return cf{-x.i, x.r};
    } else if constexpr (2 * k == N) {
        return cf{-x.r, -x.i};
    } else if constexpr (4 * k == 3 * N) {
        return cf{x.i, -x.r};
    } else {
        constexpr float c = Tw<k, N>::c, s = Tw<k, N>::s;
        return cf{f_fma(c, x.r, f_mul(-s, x.i)), f_fma(c, x.i, f_mul(s, x.r))};
    }
}

// N = 3 * Q inverse DFT: r = 3*r2 + r1, k = k1 + Q*k2
//   Y[k1 + Q k2] = sum_{r1} W3^{k2 r1} * W_N^{k1 r1} * F_{r1}[k1],  F_{r1} = IDFT_Q(in[3 r2 + r1])
template <int N>
struct Idft3xQ {
    static constexpr int Q = N / 3;
    template <int K1>
    SNRX_HD static void combine(const cf* f0, const cf* f1, const cf* f2, cf* out) {
        constexpr float h = 0.86602540378443864676f;   // sqrt(3)/2
        cf a = f0[K1];
        cf b = twmul<K1, N>(f1[K1]);
        cf c = twmul<2 * K1, N>(f2[K1]);
        cf s{f_add(b.r, c.r), f_add(b.i, c.i)};
        cf d{f_sub(b.r, c.r), f_sub(b.i, c.i)};
        out[K1].r = f_add(a.r, s.r);
        out[K1].i = f_add(a.i, s.i);
        cf m{f_fma(-0.5f, s.r, a.r), f_fma(-0.5f, s.i, a.i)};
        // k2 = 1: a + w b + w^2 c, w = exp(+j 2pi/3) = -1/2 + j h  ->  m + j h d
        out[K1 + Q].r = f_fma(-h, d.i, m.r);
        out[K1 + Q].i = f_fma(h, d.r, m.i);
        out[K1 + 2 * Q].r = f_fma(h, d.i, m.r);
        out[K1 + 2 * Q].i = f_fma(-h, d.r, m.i);
        if constexpr (K1 + 1 < Q) combine<K1 + 1>(f0, f1, f2, out);
    }
    SNRX_HD static void run(const cf* in, cf* out) {
        cf f0[Q], f1[Q], f2[Q];
        IdftPow2<Q, 3>::run(in, f0);
        IdftPow2<Q, 3>::run(in + 1, f1);
        IdftPow2<Q, 3>::run(in + 2, f2);
        combine<0>(f0, f1, f2, out);
    }
};

}  // namespace snrx
