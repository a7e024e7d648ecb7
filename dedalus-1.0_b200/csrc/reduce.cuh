// Device reductions: one-pass spectral invariants of a state and the reduction machinery they
// (and the CFL capture of the x passes) share.
//
// Replaces the volume-average tasks of dedalus/analysis/volume_average.py:71-331 (ekin, emag,
// ux2..bz2, temp2, enstrophy, energy_dissipation, divergence, divergence_sum, mag_div,
// mag_div_sum, vort_cenk), which in the reference are one or more full-array numpy passes EACH
// (deriv temporaries, abs, sum): here every invariant of a state comes out of ONE sweep that
// reads each component once (16 B per mode and component, nothing written but the partials).
//
// Determinism: a launch has a fixed grid, every thread walks a fixed index sequence, partials
// are combined in a fixed order (warp shuffle tree -> shared memory -> second kernel), so two
// runs on the same input give the same bits.
#pragma once
#include "pointwise.cuh"

namespace ddl {

enum { RED_SUM = 0, RED_MAX = 1 };

#define DDL_RED_MAXBLOCKS (148 * 8)
#define DDL_RED_MAXV 32

template <int OP> DDL_HD double red_identity() { return 0.0; }     // sums; maxima of non-negative values
template <int OP> DDL_HD double red_op(double a, double b) {
    if (OP == RED_SUM) return a + b;
    return (b > a || b != b) ? b : a;                                // NaN propagates
}

#if DDL_DEVICE_BUILD
// F: struct with `static constexpr int NR` and `void operator()(long long i, double* acc) const`
template <class F, int OP>
__global__ void __launch_bounds__(256) reduce_kernel(const __grid_constant__ F f, long long count, double* __restrict__ partial) {
    constexpr int NR = F::NR;
    static_assert(NR <= DDL_RED_MAXV, "too many reduced values");
    double acc[NR];
#pragma unroll
    for (int r = 0; r < NR; ++r) acc[r] = red_identity<OP>();
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += stride) f(i, acc);
    __shared__ double sm[8][NR];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int r = 0; r < NR; ++r) {
        double v = acc[r];
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) v = red_op<OP>(v, __shfl_xor_sync(0xffffffffu, v, off));
        if (lane == 0) sm[warp][r] = v;
    }
    __syncthreads();
    if (threadIdx.x < NR) {
        double v = sm[0][threadIdx.x];
        for (int w = 1; w < 8; ++w) v = red_op<OP>(v, sm[w][threadIdx.x]);
        partial[(long long)blockIdx.x * NR + threadIdx.x] = v;
    }
}

// one CTA of 256 threads: warp w combines the block partials of values w, w + 8, ...
template <int OP>
__global__ void __launch_bounds__(256) reduce_final_kernel(const double* __restrict__ partial, int nblocks, int nr,
                                                           double* __restrict__ out, int accumulate) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int r = warp; r < nr; r += 8) {
        double v = red_identity<OP>();
        for (int b = lane; b < nblocks; b += 32) v = red_op<OP>(v, partial[(long long)b * nr + r]);
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) v = red_op<OP>(v, __shfl_xor_sync(0xffffffffu, v, off));
        if (lane == 0) out[r] = accumulate ? red_op<OP>(out[r], v) : v;
    }
}
#endif

// out[0..NR) = reduction over i in [0, count) of what f accumulates; `partial` is device scratch of
// DDL_RED_MAXBLOCKS * NR doubles; accumulate: combine with the values already in `out`
template <class F, int OP>
int launch_reduce(const F& f, long long count, double* partial, double* out, int accumulate, ddl_stream_t stream,
                  const char* name = "reduce") {
    constexpr int NR = F::NR;
#if DDL_DEVICE_BUILD
    const int threads = 256;
    long long blocks = (count + threads - 1) / threads;
    if (blocks > DDL_RED_MAXBLOCKS) blocks = DDL_RED_MAXBLOCKS;
    if (blocks < 1) blocks = 1;
    prof_begin(name, stream);
    reduce_kernel<F, OP><<<(unsigned)blocks, threads, 0, stream>>>(f, count, partial);
    prof_end(stream);
    DDL_CUDA_CHECK(cudaGetLastError());
    prof_begin("reduce_final", stream);
    reduce_final_kernel<OP><<<1, 256, 0, stream>>>(partial, (int)blocks, NR, out, accumulate);
    prof_end(stream);
    DDL_CUDA_CHECK(cudaGetLastError());
#else
    (void)partial;
    prof_begin(name, stream);
    prof_begin("reduce_final", stream);
    double acc[NR];
    for (int r = 0; r < NR; ++r) acc[r] = red_identity<OP>();
    for (long long i = 0; i < count; ++i) f(i, acc);
    for (int r = 0; r < NR; ++r) out[r] = accumulate ? red_op<OP>(out[r], acc[r]) : acc[r];
#endif
    return 0;
}

// number of non-zero entries OUTSIDE the dealias mask, per array (ddl_reduce_outside_mask): the check that lets a buffer the
// caller has written regain its "zero outside the mask" status (retained-modes-only sweeps, fused stage kernel); reads the
// masked-out entries only
struct OutsideF {
    static constexpr int NR = DDL_MAXC;
    KGeom g;
    const cplx* arr[DDL_MAXC];
    int narr;
    DDL_HD void operator()(long long i, double* acc) const {
        int ia, ib, ic;
        split3(i, g.dim, ia, ib, ic);
        bool keep = g.keep[1][ib] && g.keep[2][ic];
        if (g.keep[0]) keep = keep && g.keep[0][ia];
        if (!keep) {
            for (int f = 0; f < narr; ++f) {
                const cplx v = arr[f][i];
                if (v.x != 0.0 || v.y != 0.0) acc[f] += 1.0;          // NaN counts
            }
        }
    }
};

// ------------------------------------------------------------------------------------------
// Spectral invariants of one state (indices: include/ddl.h DDL_INV_*).  ND = 2 / 3; the state is
// u (ND components) followed by NB components of a second group: 0 (hydro), 1 (T), ND (B).
// Weight w = 1 on the kx = 0 plane, 2 elsewhere: the half-complex layout stores kx >= 0 only
// (volume_average.py:84-97).
// ------------------------------------------------------------------------------------------
#define DDL_NINV_ 24

template <int ND, int NB>
struct InvariantsF {
    static constexpr int NR = DDL_NINV_;
    static constexpr int NCOMP = ND + NB;
    const cplx* S[ND + NB];
    KGeom g;                    // full sweep
    int compact;                // retained modes only (the state vanishes outside the mask)
    int cdim[3];
    long long fstride[3];
    const int* ftab[3];
    const double* kvc[3];
    // level -> component index (x=0, y=1, z=2; -1 = absent): 3-D levels are (y, z, x), 2-D (-, x, y)
    static constexpr int axl(int l) { return ND == 3 ? (l == 0 ? 1 : (l == 1 ? 2 : 0)) : l - 1; }
    static constexpr int XLEVEL = ND == 3 ? 2 : 1;      // level of the half-complex axis

    DDL_HD void operator()(long long i, double* acc) const {
        int j[3];               // FULL (stored) index per level
        double kk[3] = {0.0, 0.0, 0.0};
        long long fi;
        if (compact) {
            int c[3];
            split3(i, cdim, c[0], c[1], c[2]);
            fi = 0;
#pragma unroll
            for (int l = 0; l < 3; ++l) {
                j[l] = ftab[l] ? ftab[l][c[l]] : c[l];
                fi += (long long)j[l] * fstride[l];
                if (axl(l) >= 0) kk[axl(l) < 0 ? 0 : axl(l)] = kvc[l][c[l]];
            }
        } else {
            split3(i, g.dim, j[0], j[1], j[2]);
            fi = i;
#pragma unroll
            for (int l = 0; l < 3; ++l)
                if (axl(l) >= 0) kk[axl(l) < 0 ? 0 : axl(l)] = g.kv[l][j[l]];
        }
        const double w = (j[XLEVEL] == 0) ? 1.0 : 2.0;
        cplx v[NCOMP];
#pragma unroll
        for (int c = 0; c < NCOMP; ++c) v[c] = S[c][fi];
        // k^2 in the reference's summation order (representations.py:434-436): y, z, x / x, y
        const double k2 = (ND == 3) ? (kk[1] * kk[1] + kk[2] * kk[2] + kk[0] * kk[0]) : (kk[0] * kk[0] + kk[1] * kk[1]);

        double eu = 0.0;
        cplx du = mk(0.0, 0.0);
#pragma unroll
        for (int c = 0; c < ND; ++c) {
            const double m = v[c].x * v[c].x + v[c].y * v[c].y;
            eu += m;
            acc[14 + c] += w * m;
            du = du + scal(v[c], kk[c]);
        }
        acc[0] += w * 0.5 * eu;
        acc[22] += w * (du.x * du.x + du.y * du.y);         // |k.u|^2: compressive part of sum |k|^2 |u|^2 = |k.u|^2 + |k x u|^2
        acc[2] += sqrt(du.x * du.x + du.y * du.y);          // |i k.u|, unweighted (volume_average.py:287-295)
        acc[8] += w * (-du.y);                               // i k.u = (-Im, Re)
        acc[9] += w * du.x;
        // centroid wavenumber (volume_average.py:201-212): en[0,0] = 0 zeroes the first TWO stored
        // axes' origin, i.e. the single k = 0 mode in 2-D but the whole (ky,kz) = 0 line in 3-D
        const bool origin = (ND == 3) ? (j[0] == 0 && j[1] == 0) : (j[1] == 0 && j[2] == 0);
        if (!origin) {
            const double k2n = (k2 == 0.0) ? 1.0 : k2;
            acc[12] += k2n * sqrt(k2n) * 0.5 * eu;
            acc[13] += k2n * 0.5 * eu;
        }
        // vorticity i k x u
        if (ND == 3) {
            const cplx wx = lin2(kk[1], v[2], -kk[2], v[1]);
            const cplx wy = lin2(kk[2], v[0], -kk[0], v[2]);
            const cplx wz = lin2(kk[0], v[1], -kk[1], v[0]);
            acc[4] += w * 0.5 * (wx.x * wx.x + wx.y * wx.y + wy.x * wy.x + wy.y * wy.y + wz.x * wz.x + wz.y * wz.y);
            // Re(u . conj(i k x u)) = sum_c Im-type cross terms: Re(a conj(i b)) = a.y b.x - a.x b.y ... written out
            acc[6] += w * ((v[0].y * wx.x - v[0].x * wx.y) + (v[1].y * wy.x - v[1].x * wy.y) + (v[2].y * wz.x - v[2].x * wz.y));
        } else {
            const cplx wz = lin2(kk[0], v[1], -kk[1], v[0]);
            acc[4] += w * 0.5 * (wz.x * wz.x + wz.y * wz.y);
        }
        if (NB == ND) {
            double eb = 0.0, ub = 0.0;
            cplx db = mk(0.0, 0.0);
#pragma unroll
            for (int c = 0; c < ND; ++c) {
                const cplx b = v[ND + c];
                const double m = b.x * b.x + b.y * b.y;
                eb += m;
                acc[14 + ND + c] += w * m;
                db = db + scal(b, kk[c]);
                ub += v[c].x * b.x + v[c].y * b.y;
            }
            acc[1] += w * 0.5 * eb;
            acc[3] += sqrt(db.x * db.x + db.y * db.y);
            acc[23] += w * (db.x * db.x + db.y * db.y);
            acc[10] += w * (-db.y);
            acc[11] += w * db.x;
            acc[7] += w * ub;
            if (ND == 3) {
                const cplx jx = lin2(kk[1], v[5], -kk[2], v[4]);
                const cplx jy = lin2(kk[2], v[3], -kk[0], v[5]);
                const cplx jz = lin2(kk[0], v[4], -kk[1], v[3]);
                acc[5] += w * 0.5 * (jx.x * jx.x + jx.y * jx.y + jy.x * jy.x + jy.y * jy.y + jz.x * jz.x + jz.y * jz.y);
                // magnetic helicity <A.B>, A = i k x B / k^2 (Coulomb gauge), k = 0 excluded
                if (k2 != 0.0)
                    acc[20] += w * ((jx.x * v[3].y - jx.y * v[3].x) + (jy.x * v[4].y - jy.y * v[4].x) + (jz.x * v[5].y - jz.y * v[5].x)) / k2;
            } else {
                const cplx jz = lin2(kk[0], v[3], -kk[1], v[2]);
                acc[5] += w * 0.5 * (jz.x * jz.x + jz.y * jz.y);
            }
        } else if (NB == 1) {
            const cplx t = v[ND];
            const double m = t.x * t.x + t.y * t.y;
            acc[1] += w * 0.5 * m;
            acc[14 + ND] += w * m;
            acc[21] += w * k2 * m;                           // sum_i |k_i T|^2 (volume_average.py:262-271)
        }
    }
};

}  // namespace ddl
