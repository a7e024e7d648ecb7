// Host emulation of csrc/fft_core.cuh: runs the exact per-item stage code sequentially and
// compares against a naive long-double DFT.  Built and run by tests/test_host_emulation.py.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "fft_core.cuh"
using namespace ddl;

template <int N, int DIR, bool DIT, int S_IDX> struct RunStages {
    static void go(cplx* tile, int ld, int np, const cplx* tw) {
        constexpr int R = Fac<N>::radix(S_IDX);
        for (int i = 0; i < np * (N / R); ++i) stage_item<N, S_IDX, DIR, DIT>(tile, ld, i % np, i / np, tw);
        if constexpr (!DIT) { if constexpr (S_IDX + 1 < Fac<N>::S) RunStages<N, DIR, DIT, S_IDX + 1>::go(tile, ld, np, tw); }
        else { if constexpr (S_IDX > 0) RunStages<N, DIR, DIT, S_IDX - 1>::go(tile, ld, np, tw); }
    }
};

template <int N, int DIR> double check() {
    const int np = 3, ld = 5;
    std::vector<cplx> tw(N), tile(N * ld), x(N * np);
    const long double PI = acosl(-1.0L);
    for (int m = 0; m < N; ++m) { tw[m].x = (double)cosl(-2 * PI * m / N); tw[m].y = (double)sinl(-2 * PI * m / N); }
    for (auto& v : x) { v.x = drand48() - 0.5; v.y = drand48() - 0.5; }
    std::vector<long double> re(N * np), im(N * np);
    for (int c = 0; c < np; ++c)
        for (int k = 0; k < N; ++k) {
            long double sr = 0, si = 0;
            for (int n = 0; n < N; ++n) {
                long double a = DIR * 2 * PI * (long double)((long long)k * n % N) / N;
                sr += x[n * np + c].x * cosl(a) - x[n * np + c].y * sinl(a);
                si += x[n * np + c].x * sinl(a) + x[n * np + c].y * cosl(a);
            }
            re[k * np + c] = sr; im[k * np + c] = si;
        }
    double err = 0, nrm = 0;
    // DIF: natural in, scrambled out
    for (int n = 0; n < N; ++n) for (int c = 0; c < np; ++c) tile[n * ld + c] = x[n * np + c];
    RunStages<N, DIR, false, 0>::go(tile.data(), ld, np, tw.data());
    for (int k = 0; k < N; ++k) for (int c = 0; c < np; ++c) {
        int p = pos_of_index<N>(k);
        if (index_of_pos<N>(p) != k) { printf("N=%d bad index maps\n", N); exit(1); }
        cplx v = tile[p * ld + c];
        err += (double)((v.x - re[k * np + c]) * (v.x - re[k * np + c]) + (v.y - im[k * np + c]) * (v.y - im[k * np + c]));
        nrm += (double)(re[k * np + c] * re[k * np + c] + im[k * np + c] * im[k * np + c]);
    }
    double e1 = sqrt(err / nrm);
    // DIT: scrambled in, natural out
    err = 0;
    for (int n = 0; n < N; ++n) for (int c = 0; c < np; ++c) tile[pos_of_index<N>(n) * ld + c] = x[n * np + c];
    RunStages<N, DIR, true, Fac<N>::S - 1>::go(tile.data(), ld, np, tw.data());
    for (int k = 0; k < N; ++k) for (int c = 0; c < np; ++c) {
        cplx v = tile[k * ld + c];
        err += (double)((v.x - re[k * np + c]) * (v.x - re[k * np + c]) + (v.y - im[k * np + c]) * (v.y - im[k * np + c]));
    }
    double e2 = sqrt(err / nrm);
    printf("N=%4d DIR=%+d  dif %.2e  dit %.2e\n", N, DIR, e1, e2);
    return e1 > e2 ? e1 : e2;
}

// runtime-length (mixed radix) stages: the same check with the radices from rt_factor
template <int DIR> double check_rt(int N) {
    RtFac f;
    if (!rt_factor(N, f)) { printf("N=%d not factorable\n", N); exit(1); }
    const int np = 3, ld = 5;
    std::vector<cplx> tw(N), tile(N * ld), x(N * np);
    const long double PI = acosl(-1.0L);
    for (int m = 0; m < N; ++m) { tw[m].x = (double)cosl(-2 * PI * m / N); tw[m].y = (double)sinl(-2 * PI * m / N); }
    for (auto& v : x) { v.x = drand48() - 0.5; v.y = drand48() - 0.5; }
    std::vector<long double> re(N * np), im(N * np);
    for (int c = 0; c < np; ++c)
        for (int k = 0; k < N; ++k) {
            long double sr = 0, si = 0;
            for (int n = 0; n < N; ++n) {
                long double a = DIR * 2 * PI * (long double)((long long)k * n % N) / N;
                sr += x[n * np + c].x * cosl(a) - x[n * np + c].y * sinl(a);
                si += x[n * np + c].x * sinl(a) + x[n * np + c].y * cosl(a);
            }
            re[k * np + c] = sr; im[k * np + c] = si;
        }
    auto stage = [&](int s, bool dit) {
        int P = 1;
        for (int t = 0; t < s; ++t) P *= f.radix[t];
        const int R = f.radix[s], M = N / P, Q = M / R;
        for (int i = 0; i < np * (N / R); ++i) {
            if (dit) stage_item_rt<DIR, true>(tile.data(), ld, i % np, i / np, N, R, M, Q, P, tw.data());
            else stage_item_rt<DIR, false>(tile.data(), ld, i % np, i / np, N, R, M, Q, P, tw.data());
        }
    };
    double err = 0, nrm = 0;
    for (int n = 0; n < N; ++n) for (int c = 0; c < np; ++c) tile[n * ld + c] = x[n * np + c];
    for (int s = 0; s < f.S; ++s) stage(s, false);
    for (int k = 0; k < N; ++k) for (int c = 0; c < np; ++c) {
        const int p = pos_of_index_rt(f, k);
        if (index_of_pos_rt(f, p) != k) { printf("N=%d bad index maps\n", N); exit(1); }
        cplx v = tile[p * ld + c];
        err += (double)((v.x - re[k * np + c]) * (v.x - re[k * np + c]) + (v.y - im[k * np + c]) * (v.y - im[k * np + c]));
        nrm += (double)(re[k * np + c] * re[k * np + c] + im[k * np + c] * im[k * np + c]);
    }
    const double e1 = sqrt(err / nrm);
    err = 0;
    for (int n = 0; n < N; ++n) for (int c = 0; c < np; ++c) tile[pos_of_index_rt(f, n) * ld + c] = x[n * np + c];
    for (int s = f.S - 1; s >= 0; --s) stage(s, true);
    for (int k = 0; k < N; ++k) for (int c = 0; c < np; ++c) {
        cplx v = tile[k * ld + c];
        err += (double)((v.x - re[k * np + c]) * (v.x - re[k * np + c]) + (v.y - im[k * np + c]) * (v.y - im[k * np + c]));
    }
    const double e2 = sqrt(err / nrm);
    printf("N=%4d DIR=%+d  rt (%d stages, max radix %d)  dif %.2e  dit %.2e\n", N, DIR, f.S, f.rmax, e1, e2);
    return e1 > e2 ? e1 : e2;
}

int main() {
    double worst = 0;
    for (int N : {2, 3, 6, 9, 10, 12, 15, 22, 30, 45, 48, 49, 50, 61, 100, 450, 2 * 3 * 5 * 7 * 11}) {
        const double a = check_rt<-1>(N), b = check_rt<+1>(N);
        if (a > worst) worst = a;
        if (b > worst) worst = b;
    }
    RtFac bad;
    if (rt_factor(2 * 67, bad) || rt_factor(0, bad)) { printf("rt_factor accepted an unsupported length\n"); return 1; }
#define CHK(N) { double a = check<N, -1>(), b = check<N, +1>(); if (a > worst) worst = a; if (b > worst) worst = b; }
    CHK(2) CHK(4) CHK(8) CHK(16) CHK(32) CHK(64) CHK(128) CHK(256) CHK(512) CHK(1024) CHK(2048)
    printf("worst %.3e\n", worst);
    return worst < 5e-15 ? 0 : 1;
}
