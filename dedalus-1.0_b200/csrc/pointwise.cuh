// Pointwise spectral kernels: RK stage updates with in-register integrating factors,
// spectral assembly (derivatives, curl, solenoidal projection), dealias mask, ik-derivative.
//
// Replaces dedalus/time_stepping/forward_step_cy_{2d,3d}.pyx (euler/etd1/etd2rk1/etd2rk2),
// dedalus/data_objects/dealias_cy_{2d,3d}.pyx, representations.py:419-425 (deriv) and the
// numpy passes of physics.py:180-195,374-416,588-599.
#pragma once
#include <cmath>

#include "ddl_common.cuh"
#include "physics_ops.cuh"

namespace ddl {

// A k-space array seen as three nested levels (level 2 fastest):
//   3-D: [ny][nz][nx/2+1] -> levels (y, z, x);   2-D: [1][nx/2+1][ny] -> levels (-, x, y)
struct KGeom {
    int dim[3];                 // full extents per level
    const double* kv[3];        // wavenumber value per full index (NULL for an absent level)
    const unsigned char* keep[3];  // 1 = survives dealiasing
    int ax[3];                  // level -> component index (x=0,y=1,z=2), -1 = absent
    int twod;
};

DDL_HD void split3(long long i, const int* dim, int& a, int& b, int& c) {
    c = (int)(i % dim[2]);
    long long r = i / dim[2];
    b = (int)(r % dim[1]);
    a = (int)(r / dim[1]);
}

// sum k^2 in the reference's order (representations.py:434-436: dict order y,z,x / x,y)
DDL_HD double ksq(const KGeom& g, int a, int b, int c) {
    double k2 = 0.0;
    if (g.kv[0]) { double v = g.kv[0][a]; k2 += v * v; }
    { double v = g.kv[1][b]; k2 += v * v; }
    { double v = g.kv[2][c]; k2 += v * v; }
    return k2;
}

DDL_HD double ipow(double x, int n) {
    if (n == 1) return x;
    if (n == 2) return x * x;
    return pow(x, (double)n);
}

// phi functions of the ETD schemes, branch structure of forward_step_cy_3d.pyx:50-59 /
// forward_step_cy_2d.pyx:50-60 (|Z| < 0.5: Taylor to Z^14/14!; else closed forms).
DDL_HD void phi_funcs(double Z, int twod, double& f0, double& f1, double& f2) {
    if (Z < 0.5 && Z > -0.5) {
        // 1/j!, j = 0..14
        const double r[15] = {1.0, 1.0, 0.5, 1.0 / 6, 1.0 / 24, 1.0 / 120, 1.0 / 720, 1.0 / 5040, 1.0 / 40320,
                              1.0 / 362880, 1.0 / 3628800, 1.0 / 39916800, 1.0 / 479001600, 1.0 / 6227020800.0,
                              1.0 / 87178291200.0};
        double a1 = r[14], a2 = r[14], a0 = r[14];
#pragma unroll
        for (int j = 13; j >= 2; --j) { a2 = a2 * Z + r[j]; }
#pragma unroll
        for (int j = 13; j >= 1; --j) { a1 = a1 * Z + r[j]; }
        f1 = a1; f2 = a2;
        if (twod) {
#pragma unroll
            for (int j = 13; j >= 0; --j) { a0 = a0 * Z + r[j]; }
            f0 = a0;
        } else {
            f0 = exp(Z);
        }
    } else {
        f0 = exp(Z);
        f1 = (f0 - 1.0) / Z;
        f2 = (f1 - 1.0) / Z;
    }
}

struct StageArgs {
    KGeom g;
    cplx* out[DDL_MAXC];
    const cplx* start[DDL_MAXC];
    const cplx* d1[DDL_MAXC];
    const cplx* d2[DDL_MAXC];
    cplx* total[DDL_MAXC];
    double coeff[DDL_MAXC];
    int ncomp, kind, vo, first, last;
    double dt, wdiv;
    // compact sweep: visit only the retained modes (every operand is known to vanish outside
    // the dealias mask, so the update there is 0 -> 0)
    int compact;
    int cdim[3];
    long long fstride[3];
    const int* ftab[3];
    const double* kvc[3];
};

enum { SK_EULER = 0, SK_ETD1 = 1, SK_ETD2RK1 = 2, SK_ETD2RK2 = 3, SK_RK4 = 4, SK_CN = 5 };

DDL_HD cplx etd1_step(cplx s, cplx d, double Z, double f0, double f1, double dt) {
    if (Z == 0.0) return mk(s.x + dt * d.x, s.y + dt * d.y);
    const double a = f1 * dt;
    return mk(s.x * f0 + d.x * a, s.y * f0 + d.y * a);
}

// One component of one mode: the update of stage kind `kind` given the start value s, the
// derivative(s) and the phi functions of Z.  Shared by the plain sweep (StageF) and by the sweep
// fused with the spectral assembly (AssembleStageF), so that both are the same arithmetic.
//   EULER / ETD1 / CN: d1 = derivative;  ETD2RK1/2: d1 = first, d2 = second derivative;
//   RK4: d1 = k_i, `total` = running weighted sum (read unless first, written unless last)
DDL_HD cplx stage_apply(int kind, cplx s, cplx d1, cplx d2, cplx* total, int first, int last, double wdiv, double Z,
                        double f0, double f1, double f2, double dt, double IF) {
    if (kind == SK_EULER) return mk(s.x + dt * d1.x, s.y + dt * d1.y);
    if (kind == SK_ETD1) return etd1_step(s, d1, Z, f0, f1, dt);
    if (kind == SK_ETD2RK1) {
        const double w = (Z == 0.0) ? dt / 2.0 : f2 * dt;
        return mk(s.x + (d2.x - d1.x) * w, s.y + (d2.y - d1.y) * w);
    }
    if (kind == SK_ETD2RK2) {
        if (Z == 0.0) return mk(s.x + dt * d2.x, s.y + dt * d2.y);
        const double w2 = 2.0 * f2 * dt, w1 = f1 * dt;
        return mk(s.x * f0 + (d2.x - d1.x) * w2 + d1.x * w1, s.y * f0 + (d2.y - d1.y) * w2 + d1.y * w1);
    }
    if (kind == SK_RK4) {
        cplx t = mk(d1.x / wdiv, d1.y / wdiv);
        if (!first) { const cplx o = *total; t = mk(o.x + t.x, o.y + t.y); }
        if (!last) *total = t;
        return etd1_step(s, last ? t : d1, Z, f0, f1, dt);
    }
    // SK_CN
    const double top = 1.0 / dt - 0.5 * IF, bottom = 1.0 / dt + 0.5 * IF;
    const double r1 = top / bottom, r2 = 1.0 / bottom;
    return mk(r1 * s.x + r2 * d1.x, r1 * s.y + r2 * d1.y);
}

struct StageF {
    StageArgs a;
    DDL_HD void operator()(long long i) const {
        int ia, ib, ic;
        double k2;
        if (a.compact) {
            int j[3];
            split3(i, a.cdim, j[0], j[1], j[2]);
            long long fi = 0;
            k2 = 0.0;
#pragma unroll
            for (int l = 0; l < 3; ++l) {
                fi += (long long)(a.ftab[l] ? a.ftab[l][j[l]] : j[l]) * a.fstride[l];
                if (a.kvc[l]) { const double v = a.kvc[l][j[l]]; k2 += v * v; }
            }
            i = fi;
        } else {
            split3(i, a.g.dim, ia, ib, ic);
            k2 = ksq(a.g, ia, ib, ic);
        }
        const double pw = ipow(k2, a.vo);
        double lastc = -1.0, Z = 0.0, f0 = 1.0, f1 = 1.0, f2 = 0.5;
        for (int c = 0; c < a.ncomp; ++c) {
            const double co = a.coeff[c];
            if (a.kind != SK_EULER && co != lastc) {
                lastc = co;
                Z = -(co * pw) * a.dt;
                if (Z != 0.0 && a.kind != SK_CN) phi_funcs(Z, a.g.twod, f0, f1, f2);
            }
            const bool two = (a.kind == SK_ETD2RK1 || a.kind == SK_ETD2RK2);
            const cplx d2 = two ? a.d2[c][i] : mk(0.0, 0.0);
            a.out[c][i] = stage_apply(a.kind, a.start[c][i], a.d1[c][i], d2, a.kind == SK_RK4 ? &a.total[c][i] : nullptr,
                                      a.first, a.last, a.wdiv, Z, f0, f1, f2, a.dt, co * pw);
        }
    }
};

// euler / etd1 / etd2rk1 / etd2rk2 with the integrating factor as an ARRAY: the literal signature of the reference's Cython
// kernels (forward_step_cy_3d.pyx:17-124, _2d:21-120: start, output, deriv1[, deriv2], intfactor, dt; Z = intfactor * dt),
// for callers whose factor is not c (k^2)^n of static wavenumbers (shearing box) or who bind the kernels one to one
struct StageArrayF {
    const cplx* start;
    cplx* out;
    const cplx* d1;
    const cplx* d2;
    const double* intfactor;    // NULL: Euler forms (the reference's `integrating_factor is None` branches)
    int kind, twod;
    double dt;
    DDL_HD void operator()(long long i) const {
        double Z = 0.0, f0 = 1.0, f1 = 1.0, f2 = 0.5;
        if (intfactor && kind != SK_EULER) {
            Z = intfactor[i] * dt;
            if (Z != 0.0) phi_funcs(Z, twod, f0, f1, f2);
        }
        const bool two = (kind == SK_ETD2RK1 || kind == SK_ETD2RK2);
        out[i] = stage_apply(kind, start[i], d1[i], two ? d2[i] : mk(0.0, 0.0), nullptr, 0, 0, 1.0, Z, f0, f1, f2, dt, 0.0);
    }
};

// dealias_23 with the reference's own signature (dealias_cy_2d.pyx:13-41, dealias_cy_3d.pyx:13-46): wavenumber VALUES per
// axis, ky optionally dense over (kx, ky) / (ky, kx) -- the shearing box's second branch -- and the Nyquist values
struct DealiasArrayF {
    cplx* data;
    const double* kx; const double* ky; const double* kz;
    int ndim, ky_dense;
    int dim[3];                 // 3-D: (ny, nz, nkx); 2-D: (1, nkx, ny)
    double cut[3];              // 2/3 k_nyquist for x, y, z
    DDL_HD static bool out(double k, double c) { return k >= c || k <= -c; }
    DDL_HD void operator()(long long i) const {
        int a, b, c;
        split3(i, dim, a, b, c);
        bool zero;
        if (ndim == 3) {        // data[y][z][x]
            const double kyv = ky_dense ? ky[(long long)a * dim[2] + c] : ky[a];
            zero = out(kx[c], cut[0]) || out(kyv, cut[1]) || out(kz[b], cut[2]);
        } else {                // data[x][y]
            const double kyv = ky_dense ? ky[(long long)b * dim[2] + c] : ky[c];
            zero = out(kx[b], cut[0]) || out(kyv, cut[1]);
        }
        if (zero) data[i] = mk(0.0, 0.0);
    }
};

// The stage update at the modes OUTSIDE the dealias mask, where every derivative of the fused pipeline vanishes: what a state
// that carries content there (hydro never dealiases its state, SURVEY F7: the Nyquist-row entries of the reference's 2-D
// Taylor-Green field) still has to undergo -- the integrating factor alone.  Same arithmetic as StageF with d1 = d2 = 0, so
// the result equals the full sweep's bit for bit; touches the masked-out entries only.
struct OutsideStageF {
    StageArgs a;
    DDL_HD void operator()(long long i) const {
        int ia, ib, ic;
        split3(i, a.g.dim, ia, ib, ic);
        bool keep = a.g.keep[1][ib] && a.g.keep[2][ic];
        if (a.g.keep[0]) keep = keep && a.g.keep[0][ia];
        if (keep) return;
        const double pw = ipow(ksq(a.g, ia, ib, ic), a.vo);
        const cplx zero = mk(0.0, 0.0);
        for (int c = 0; c < a.ncomp; ++c) {
            const double co = a.coeff[c];
            double Z = 0.0, f0 = 1.0, f1 = 1.0, f2 = 0.5;
            if (a.kind != SK_EULER) {
                Z = -(co * pw) * a.dt;
                if (Z != 0.0 && a.kind != SK_CN) phi_funcs(Z, a.g.twod, f0, f1, f2);
            }
            cplx tot = zero;
            a.out[c][i] = stage_apply(a.kind, a.start[c][i], zero, zero, a.kind == SK_RK4 ? &tot : nullptr, 1, 1, a.wdiv, Z, f0, f1, f2,
                                      a.dt, co * pw);
        }
    }
};

// zero the masked-out modes of up to DDL_MAXF arrays, touching only those entries
struct MaskF {
    KGeom g;
    cplx* arr[DDL_MAXF];
    int narr;
    DDL_HD void operator()(long long i) const {
        int ia, ib, ic;
        split3(i, g.dim, ia, ib, ic);
        bool keep = g.keep[1][ib] && g.keep[2][ic];
        if (g.keep[0]) keep = keep && g.keep[0][ia];
        if (!keep) {
            for (int f = 0; f < narr; ++f) arr[f][i] = mk(0.0, 0.0);
        }
    }
};

struct DerivF {
    KGeom g;
    const cplx* in;
    cplx* out;
    int level;
    DDL_HD void operator()(long long i) const {
        int idx[3];
        split3(i, g.dim, idx[0], idx[1], idx[2]);
        const double k = g.kv[level][idx[level]];
        cplx v = in[i];
        out[i] = mk(-v.y * k, v.x * k);
    }
};

// divergence spectra theta = i k.F of the velocity (and magnetic) field, the extra inverse-transform inputs
// of the advective-form policies (physics_ops.cuh AdvOf)
struct ThetaF {
    KGeom g;
    const cplx* U[3];
    const cplx* B[3];
    cplx* thu;
    cplx* thb;          // NULL unless the state carries B
    int nd;
    DDL_HD void operator()(long long i) const {
        int idx[3];
        split3(i, g.dim, idx[0], idx[1], idx[2]);
        double kk[3] = {0.0, 0.0, 0.0};
#pragma unroll
        for (int l = 0; l < 3; ++l)
            if (g.ax[l] >= 0) kk[g.ax[l]] = g.kv[l][idx[l]];
        cplx du = mk(0.0, 0.0), db = mk(0.0, 0.0);
        for (int c = 0; c < nd; ++c) {
            du = du + scal(U[c][i], kk[c]);
            if (thb) db = db + scal(B[c][i], kk[c]);
        }
        thu[i] = mk(-du.y, du.x);
        if (thb) thb[i] = mk(-db.y, db.x);
    }
};

// spectral assembly over the retained modes: compact product arrays -> full deriv arrays
template <class PHYS>
struct AssembleF {
    const cplx* P[PHYS::NO];            // compact [c0][c1][pitch]
    const cplx* S[PHYS::NS ? PHYS::NS : 1];   // state (full layout)
    cplx* D[PHYS::NC];                  // deriv (full layout)
    int cdim[3];                        // compact extents per level
    long long cstride[3];               // compact strides per level
    long long fstride[3];               // full strides per level
    const int* ftab[3];                 // compact -> full index per level (NULL = identity)
    const double* kvc[3];               // wavenumber per COMPACT index per level
    int ax[3];
    PhysConst pc;
    DDL_HD void operator()(long long i) const {
        int j[3];
        split3(i, cdim, j[0], j[1], j[2]);
        long long ci = 0, fi = 0;
        double kk[3] = {0.0, 0.0, 0.0};
#pragma unroll
        for (int l = 0; l < 3; ++l) {
            ci += j[l] * cstride[l];
            fi += (long long)(ftab[l] ? ftab[l][j[l]] : j[l]) * fstride[l];
            if (ax[l] >= 0) kk[ax[l]] = kvc[l][j[l]];
        }
        cplx p[PHYS::NO], s[PHYS::NS ? PHYS::NS : 1], d[PHYS::NC];
#pragma unroll
        for (int f = 0; f < PHYS::NO; ++f) p[f] = P[f][ci];
#pragma unroll
        for (int f = 0; f < PHYS::NS; ++f) s[f] = S[f][fi];
        PHYS::assemble(p, s, d, kk[0], kk[1], kk[2], pc);
#pragma unroll
        for (int f = 0; f < PHYS::NC; ++f) D[f][fi] = d[f];
    }
};

// Spectral assembly fused with a stage update of any integrator: the derivative k is formed in
// registers from the product spectra and consumed at once (ETD1 / Euler / RK4 / CN: as the
// derivative; ETD2RK1/2: as the SECOND derivative, the first is read from deriv1), so it is not
// written unless a later stage needs it (kout, RK2's k1).  Saves the 16 B write + 16 B read per mode
// and component, and a launch.  Retained modes only: every operand must vanish outside the mask.
// HOIST: every operand of the mode (product spectra, start values, running totals, first derivatives) is loaded before any
// arithmetic; without it the per-component loads sit behind the branches and divisions of the previous component's update and
// a mode pays one memory round trip per component (seven for 3-D MHD in the SASS) instead of one or two.
template <class PHYS, bool HOIST = false>
struct AssembleStageF {
    AssembleF<PHYS> a;
    const cplx* y[PHYS::NC];        // start
    const cplx* d1[PHYS::NC];       // ETD2RK1/2: first derivative
    cplx* kout[PHYS::NC];           // optional: store k
    cplx* total[PHYS::NC];          // RK4
    cplx* out[PHYS::NC];
    double coeff[PHYS::NC];
    int kind, vo, first, last, twod, has_d1, has_kout;
    double dt, wdiv;
    DDL_HD void operator()(long long i) const {
        int j[3];
        split3(i, a.cdim, j[0], j[1], j[2]);
        long long ci = 0, fi = 0;
        double kk[3] = {0.0, 0.0, 0.0};
        double k2 = 0.0;
#pragma unroll
        for (int l = 0; l < 3; ++l) {
            ci += j[l] * a.cstride[l];
            fi += (long long)(a.ftab[l] ? a.ftab[l][j[l]] : j[l]) * a.fstride[l];
            if (a.ax[l] >= 0) { const double v = a.kvc[l][j[l]]; kk[a.ax[l]] = v; k2 += v * v; }   // StageF's summation order
        }
        cplx p[PHYS::NO], s[PHYS::NS ? PHYS::NS : 1], d[PHYS::NC];
#pragma unroll
        for (int f = 0; f < PHYS::NO; ++f) p[f] = a.P[f][ci];
#pragma unroll
        for (int f = 0; f < PHYS::NS; ++f) s[f] = a.S[f][fi];
        cplx yv[PHYS::NC], tv[PHYS::NC], dv[PHYS::NC];
        if constexpr (HOIST) {
            const bool rd_total = kind == SK_RK4 && !first;
#pragma unroll
            for (int c = 0; c < PHYS::NC; ++c) yv[c] = y[c][fi];
#pragma unroll
            for (int c = 0; c < PHYS::NC; ++c) tv[c] = rd_total ? total[c][fi] : mk(0.0, 0.0);
#pragma unroll
            for (int c = 0; c < PHYS::NC; ++c) dv[c] = has_d1 ? d1[c][fi] : mk(0.0, 0.0);
        }
        PHYS::assemble(p, s, d, kk[0], kk[1], kk[2], a.pc);
        const double pw = ipow(k2, vo);
        double lastc = -1.0, Z = 0.0, f0 = 1.0, f1 = 1.0, f2 = 0.5;
#pragma unroll
        for (int c = 0; c < PHYS::NC; ++c) {
            const double co = coeff[c];
            if (kind != SK_EULER && co != lastc) {
                lastc = co;
                Z = -(co * pw) * dt;
                if (Z != 0.0 && kind != SK_CN) phi_funcs(Z, twod, f0, f1, f2);
            }
            const cplx kc = d[c];
            if (has_kout) kout[c][fi] = kc;
            if constexpr (HOIST) {
                // same calls as below on the values loaded above; the running total goes back where stage_apply would have left it
                const cplx first_d = has_d1 ? dv[c] : kc;
                out[c][fi] = stage_apply(kind, yv[c], first_d, kc, kind == SK_RK4 ? &tv[c] : nullptr, first, last, wdiv,
                                         Z, f0, f1, f2, dt, co * pw);
                if (kind == SK_RK4 && !last) total[c][fi] = tv[c];
            } else {
                const cplx first_d = has_d1 ? d1[c][fi] : kc;
                out[c][fi] = stage_apply(kind, y[c][fi], first_d, kc, kind == SK_RK4 ? &total[c][fi] : nullptr, first, last, wdiv,
                                         Z, f0, f1, f2, dt, co * pw);
            }
        }
    }
};

#if DDL_DEVICE_BUILD
template <class F>
__global__ void __launch_bounds__(256, 4) items_kernel(const __grid_constant__ F f, long long count) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += stride) f(i);
}
#endif

#if DDL_DEVICE_BUILD
// the same sweep with MINB resident CTAs per SM instead of four: 65536 / (256 MINB) registers per thread for loads in flight
template <class F, int MINB>
__global__ void __launch_bounds__(256, MINB) items_kernel_b(const __grid_constant__ F f, long long count) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += stride) f(i);
}
#endif

template <class F, int MINB>
int launch_items_b(const F& f, long long count, ddl_stream_t stream, const char* name) {
    if (count <= 0) return 0;
#if DDL_DEVICE_BUILD
    const int threads = 256;
    long long blocks = (count + threads - 1) / threads;
    const long long cap = 148LL * 4 * MINB;
    if (blocks > cap) blocks = cap;
    prof_begin(name, stream);
    items_kernel_b<F, MINB><<<(unsigned)blocks, threads, 0, stream>>>(f, count);
    prof_end(stream);
    DDL_CUDA_CHECK(cudaGetLastError());
#else
    prof_begin(name, stream);
    for (long long i = 0; i < count; ++i) f(i);
#endif
    return 0;
}

template <class F>
int launch_items(const F& f, long long count, ddl_stream_t stream, const char* name = "pointwise") {
    if (count <= 0) return 0;
#if DDL_DEVICE_BUILD
    const int threads = 256;
    long long blocks = (count + threads - 1) / threads;
    const long long cap = 148LL * 16;
    if (blocks > cap) blocks = cap;
    prof_begin(name, stream);
    items_kernel<F><<<(unsigned)blocks, threads, 0, stream>>>(f, count);
    prof_end(stream);
    DDL_CUDA_CHECK(cudaGetLastError());
#else
    prof_begin(name, stream);       // the emulation counts launches like the device build (tests/ launch accounting)
    for (long long i = 0; i < count; ++i) f(i);
#endif
    return 0;
}

}  // namespace ddl
