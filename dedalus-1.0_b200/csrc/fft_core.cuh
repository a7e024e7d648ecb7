// FP64 complex FFT building blocks for tiles of pencils staged in shared memory.
//
// A tile holds `np` pencils of length N as tile[row * ld + c] (row = position along the
// transform axis, c = pencil).  Two in-place transforms are provided:
//   * tile_fft_dif: natural-order input  -> digit-reversed ("scrambled") output
//   * tile_fft_dit: scrambled input      -> natural-order output
// so that inverse(DIF) -> pointwise work at scrambled positions -> forward(DIT) needs no
// reordering pass at all.  pos_of_index / index_of_pos translate between the two orders.
//
// Every stage is a loop over work items (pencil, butterfly) strided by the CTA size, one
// __syncthreads() per stage.  The per-item bodies are plain functions so that the host
// emulator (tests/host/emul_fft.cpp) can execute exactly the same code without a GPU.
#pragma once

#ifdef __CUDACC__
#include <cuda_runtime.h>
#define DDL_HD __host__ __device__ __forceinline__
#define DDL_D __device__ __forceinline__
typedef double2 cplx;
#else
#define DDL_HD inline
struct alignas(16) cplx { double x, y; };
#endif

namespace ddl {

DDL_HD cplx mk(double x, double y) { cplx r; r.x = x; r.y = y; return r; }
DDL_HD cplx operator+(cplx a, cplx b) { return mk(a.x + b.x, a.y + b.y); }
DDL_HD cplx operator-(cplx a, cplx b) { return mk(a.x - b.x, a.y - b.y); }
DDL_HD cplx cmul(cplx a, cplx b) { return mk(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
DDL_HD cplx conj(cplx a) { return mk(a.x, -a.y); }
DDL_HD cplx scal(cplx a, double s) { return mk(a.x * s, a.y * s); }
// multiply by exp(DIR * i*pi/2): DIR=-1 -> -i (forward), DIR=+1 -> +i (inverse)
template <int DIR> DDL_HD cplx mul_i(cplx a) { return DIR < 0 ? mk(a.y, -a.x) : mk(-a.y, a.x); }

// ---------------------------------------------------------------------------------------
// Factorisation of N (powers of two): radix-8 stages first, then radix-4, at most one radix-2.
// ---------------------------------------------------------------------------------------
constexpr int ilog2(int n) { return n <= 1 ? 0 : 1 + ilog2(n / 2); }
constexpr bool is_pow2(int n) { return n > 0 && (n & (n - 1)) == 0; }

template <int N> struct Fac {
    static_assert(is_pow2(N) && N >= 2, "transform sizes must be powers of two");
    static constexpr int E = ilog2(N);
    // e = 3a + 2b (+1 only for N == 2)
    static constexpr int N4 = (E == 1) ? 0 : (E % 3 == 0 ? 0 : (E % 3 == 2 ? 1 : 2));
    static constexpr int N8 = (E == 1) ? 0 : (E - 2 * N4) / 3;
    static constexpr int N2 = (E == 1) ? 1 : 0;
    static constexpr int S = N8 + N4 + N2;
    static constexpr int radix(int s) { return s < N8 ? 8 : (s < N8 + N4 ? 4 : 2); }
    // number of independent sub-problems before stage s
    static constexpr int P(int s) { return s == 0 ? 1 : P(s - 1) * radix(s - 1); }
    static constexpr int RMAX = N8 ? 8 : (N4 ? 4 : 2);
    static constexpr int T = N / RMAX;   // natural thread count per pencil
};

// position (scrambled order) of natural index k, and the inverse map
template <int N> DDL_HD int pos_of_index(int k) {
    int p = 0, rem = N;
#pragma unroll
    for (int s = 0; s < Fac<N>::S; ++s) {
        const int R = Fac<N>::radix(s);
        rem /= R;
        p += (k % R) * rem;
        k /= R;
    }
    return p;
}
template <int N> DDL_HD int index_of_pos(int p) {
    int k = 0, rem = N, w = 1;
#pragma unroll
    for (int s = 0; s < Fac<N>::S; ++s) {
        const int R = Fac<N>::radix(s);
        rem /= R;
        k += (p / rem) * w;
        p %= rem;
        w *= R;
    }
    return k;
}

// ---------------------------------------------------------------------------------------
// Small DFTs, in place, natural order out:  v[r] = sum_j v[j] * exp(DIR * 2*pi*i * j*r / R)
// ---------------------------------------------------------------------------------------
template <int DIR> DDL_HD void dft2(cplx& a, cplx& b) { cplx t = a - b; a = a + b; b = t; }

template <int DIR> DDL_HD void dft4(cplx& a, cplx& b, cplx& c, cplx& d) {
    cplx t0 = a + c, t1 = a - c, t2 = b + d, t3 = mul_i<DIR>(b - d);
    a = t0 + t2; c = t0 - t2; b = t1 + t3; d = t1 - t3;
}

template <int DIR> DDL_HD void dft8(cplx (&v)[8]) {
    const double h = 0.70710678118654752440;
    // even / odd 4-point transforms
    dft4<DIR>(v[0], v[2], v[4], v[6]);
    dft4<DIR>(v[1], v[3], v[5], v[7]);
    // twiddle odd outputs by w8^r, w8 = exp(DIR*i*pi/4)
    cplx o1 = v[3], o2 = v[5], o3 = v[7];
    // after dft4 in place on (1,3,5,7): v[1]=O0, v[3]=O1, v[5]=O2, v[7]=O3; same for evens
    cplx O0 = v[1];
    cplx O1 = DIR < 0 ? mk((o1.x + o1.y) * h, (o1.y - o1.x) * h) : mk((o1.x - o1.y) * h, (o1.x + o1.y) * h);
    cplx O2 = mul_i<DIR>(o2);
    cplx O3 = DIR < 0 ? mk((o3.y - o3.x) * h, -(o3.x + o3.y) * h) : mk(-(o3.x + o3.y) * h, (o3.x - o3.y) * h);
    cplx E0 = v[0], E1 = v[2], E2 = v[4], E3 = v[6];
    v[0] = E0 + O0; v[4] = E0 - O0;
    v[1] = E1 + O1; v[5] = E1 - O1;
    v[2] = E2 + O2; v[6] = E2 - O2;
    v[3] = E3 + O3; v[7] = E3 - O3;
}

template <int R, int DIR> DDL_HD void dftR(cplx (&v)[R]) {
    if constexpr (R == 2) dft2<DIR>(v[0], v[1]);
    else if constexpr (R == 4) dft4<DIR>(v[0], v[1], v[2], v[3]);
    else dft8<DIR>(v);
}

// twiddle factors: table tw[m] = exp(-2*pi*i*m/N), m in [0,N).  v[r] *= w^(r*step) with
// w = tw (forward) or conj(tw) (inverse); few table loads, the rest by products.
template <int R, int DIR> DDL_HD void apply_twiddles(cplx (&v)[R], int step, const cplx* __restrict__ tw) {
    if (step == 0) return;
    cplx w1 = tw[step];
    if (DIR > 0) w1 = conj(w1);
    v[1] = cmul(v[1], w1);
    if constexpr (R >= 4) {
        cplx w2 = tw[2 * step];
        if (DIR > 0) w2 = conj(w2);
        v[2] = cmul(v[2], w2);
        v[3] = cmul(v[3], cmul(w1, w2));
        if constexpr (R == 8) {
            cplx w4 = tw[4 * step];
            if (DIR > 0) w4 = conj(w4);
            v[4] = cmul(v[4], w4);
            v[5] = cmul(v[5], cmul(w4, w1));
            v[6] = cmul(v[6], cmul(w4, w2));
            v[7] = cmul(v[7], cmul(w4, cmul(w1, w2)));
        }
    }
}

// ---------------------------------------------------------------------------------------
// One work item of stage s: pencil c, butterfly w in [0, N/R).
// DIF: load, DFT, twiddle, store (same positions).  DIT (transposed graph): load, twiddle,
// DFT, store.  Positions: q*M + b + j*Q with M = N/P(s), Q = M/R, q = w / Q, b = w % Q.
// ---------------------------------------------------------------------------------------
template <int N, int S_IDX, int DIR, bool DIT>
DDL_HD void stage_item(cplx* tile, int ld, int c, int w, const cplx* __restrict__ tw) {
    constexpr int R = Fac<N>::radix(S_IDX);
    constexpr int P = Fac<N>::P(S_IDX);
    constexpr int M = N / P;
    constexpr int Q = M / R;
    const int q = w / Q, b = w % Q;
    cplx* base = tile + (size_t)(q * M + b) * ld + c;
    cplx v[R];
#pragma unroll
    for (int j = 0; j < R; ++j) v[j] = base[(size_t)j * Q * ld];
    if constexpr (DIT) {
        if constexpr (Q > 1) apply_twiddles<R, DIR>(v, b * P, tw);
        dftR<R, DIR>(v);
    } else {
        dftR<R, DIR>(v);
        if constexpr (Q > 1) apply_twiddles<R, DIR>(v, b * P, tw);
    }
#pragma unroll
    for (int j = 0; j < R; ++j) base[(size_t)j * Q * ld] = v[j];
}

// ---------------------------------------------------------------------------------------
// Runtime factorisation: any length whose prime factors are <= DDL_RT_MAXR.  The reference
// transforms through FFTW, which takes any N (its own samples run 450 x 450, 48 x 2 x 48,
// 100 x 100, 30 x 10 grids); the compile-time kernels above cover the powers of two, every
// other length goes through the same stage structure with the radices in a table:
// 8 / 4 / 2 for the power-of-two part, hand-written 3- and 5-point butterflies, and a direct
// O(R^2) DFT for any other prime factor (compatibility path: correctness over speed).
// ---------------------------------------------------------------------------------------
#define DDL_RT_MAXS 16
#define DDL_RT_MAXR 64

struct RtFac {
    int n;                      // transform length
    int S;                      // number of stages
    int rmax;                   // largest radix (sizes the CTA)
    int radix[DDL_RT_MAXS];
};

// host side: fills f, returns false when n has a prime factor > DDL_RT_MAXR
inline bool rt_factor(int n, RtFac& f) {
    f.n = n; f.S = 0; f.rmax = 1;
    if (n < 1) return false;
    int rem = n;
    auto push = [&](int r) { f.radix[f.S++] = r; if (r > f.rmax) f.rmax = r; rem /= r; };
    while (rem % 8 == 0) push(8);
    while (rem % 4 == 0) push(4);
    while (rem % 2 == 0) push(2);
    for (int p = 3; p <= DDL_RT_MAXR && rem > 1; p += 2)
        while (rem % p == 0) { if (f.S >= DDL_RT_MAXS) return false; push(p); }
    if (rem != 1) return false;
    if (f.S == 0) push(1);      // n == 1: one trivial stage
    return f.S <= DDL_RT_MAXS;
}

DDL_HD int pos_of_index_rt(const RtFac& f, int k) {
    int p = 0, rem = f.n;
    for (int s = 0; s < f.S; ++s) {
        const int R = f.radix[s];
        rem /= R;
        p += (k % R) * rem;
        k /= R;
    }
    return p;
}
DDL_HD int index_of_pos_rt(const RtFac& f, int p) {
    int k = 0, rem = f.n, w = 1;
    for (int s = 0; s < f.S; ++s) {
        const int R = f.radix[s];
        rem /= R;
        k += (p / rem) * w;
        p %= rem;
        w *= R;
    }
    return k;
}

template <int DIR> DDL_HD void dft3(cplx& a, cplx& b, cplx& c) {
    const double s = 0.86602540378443864676;            // sin(2 pi / 3)
    const cplx t = b + c, d = scal(mul_i<DIR>(b - c), s);
    const cplx m = mk(a.x - 0.5 * t.x, a.y - 0.5 * t.y);
    a = a + t; b = m + d; c = m - d;
}

template <int DIR> DDL_HD void dft5(cplx (&v)[5]) {
    const double c1 = 0.30901699437494742410, c2 = -0.80901699437494742410;    // cos(2 pi/5), cos(4 pi/5)
    const double s1 = 0.95105651629515357212, s2 = 0.58778525229247312917;     // sin(2 pi/5), sin(4 pi/5)
    const cplx a1 = v[1] + v[4], a2 = v[2] + v[3], b1 = v[1] - v[4], b2 = v[2] - v[3];
    const cplx m1 = mk(v[0].x + c1 * a1.x + c2 * a2.x, v[0].y + c1 * a1.y + c2 * a2.y);
    const cplx m2 = mk(v[0].x + c2 * a1.x + c1 * a2.x, v[0].y + c2 * a1.y + c1 * a2.y);
    const cplx d1 = mul_i<DIR>(mk(s1 * b1.x + s2 * b2.x, s1 * b1.y + s2 * b2.y));
    const cplx d2 = mul_i<DIR>(mk(s2 * b1.x - s1 * b2.x, s2 * b1.y - s1 * b2.y));
    v[0] = v[0] + a1 + a2;
    v[1] = m1 + d1; v[4] = m1 - d1;
    v[2] = m2 + d2; v[3] = m2 - d2;
}

template <int DIR> DDL_HD cplx tw_at(const cplx* __restrict__ tw, int m) { return DIR > 0 ? conj(tw[m]) : tw[m]; }

// stage item with a compile-time radix and runtime geometry (M = n / P, Q = M / R)
template <int R, int DIR, bool DIT>
DDL_HD void stage_item_g(cplx* tile, int ld, int c, int w, int M, int Q, int P, const cplx* __restrict__ tw) {
    const int q = w / Q, b = w % Q;
    cplx* base = tile + (size_t)(q * M + b) * ld + c;
    cplx v[R];
#pragma unroll
    for (int j = 0; j < R; ++j) v[j] = base[(size_t)j * Q * ld];
    const int step = b * P;
    if (DIT && step) {
#pragma unroll
        for (int j = 1; j < R; ++j) v[j] = cmul(v[j], tw_at<DIR>(tw, j * step));
    }
    if constexpr (R == 2) dft2<DIR>(v[0], v[1]);
    else if constexpr (R == 3) dft3<DIR>(v[0], v[1], v[2]);
    else if constexpr (R == 4) dft4<DIR>(v[0], v[1], v[2], v[3]);
    else if constexpr (R == 5) dft5<DIR>(v);
    else dft8<DIR>(v);
    if (!DIT && step) {
#pragma unroll
        for (int j = 1; j < R; ++j) v[j] = cmul(v[j], tw_at<DIR>(tw, j * step));
    }
#pragma unroll
    for (int j = 0; j < R; ++j) base[(size_t)j * Q * ld] = v[j];
}

// any other prime radix: direct DFT, X[r] = sum_j v[j] W^(j r), W = exp(DIR 2 pi i / R) = tw[n / R]
template <int DIR, bool DIT>
DDL_HD void stage_item_prime(cplx* tile, int ld, int c, int w, int n, int R, int M, int Q, int P, const cplx* __restrict__ tw) {
    const int q = w / Q, b = w % Q;
    cplx* base = tile + (size_t)(q * M + b) * ld + c;
    cplx v[DDL_RT_MAXR], o[DDL_RT_MAXR];
    const int step = b * P, unit = n / R;
    for (int j = 0; j < R; ++j) {
        v[j] = base[(size_t)j * Q * ld];
        if (DIT && step && j) v[j] = cmul(v[j], tw_at<DIR>(tw, j * step));
    }
    for (int r = 0; r < R; ++r) {
        cplx acc = v[0];
        int e = 0;
        for (int j = 1; j < R; ++j) {
            e += r; if (e >= R) e -= R;                  // (j r) mod R
            acc = acc + cmul(v[j], tw_at<DIR>(tw, e * unit));
        }
        o[r] = acc;
    }
    for (int r = 0; r < R; ++r) {
        if (!DIT && step && r) o[r] = cmul(o[r], tw_at<DIR>(tw, r * step));
        base[(size_t)r * Q * ld] = o[r];
    }
}

template <int DIR, bool DIT>
DDL_HD void stage_item_rt(cplx* tile, int ld, int c, int w, int n, int R, int M, int Q, int P, const cplx* __restrict__ tw) {
    switch (R) {
        case 1: break;
        case 2: stage_item_g<2, DIR, DIT>(tile, ld, c, w, M, Q, P, tw); break;
        case 3: stage_item_g<3, DIR, DIT>(tile, ld, c, w, M, Q, P, tw); break;
        case 4: stage_item_g<4, DIR, DIT>(tile, ld, c, w, M, Q, P, tw); break;
        case 5: stage_item_g<5, DIR, DIT>(tile, ld, c, w, M, Q, P, tw); break;
        case 8: stage_item_g<8, DIR, DIT>(tile, ld, c, w, M, Q, P, tw); break;
        default: stage_item_prime<DIR, DIT>(tile, ld, c, w, n, R, M, Q, P, tw); break;
    }
}

}  // namespace ddl
