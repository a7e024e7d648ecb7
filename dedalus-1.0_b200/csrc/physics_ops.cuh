// Real-space products and spectral assembly for the three physics classes.
//
// Conservative ("flux") formulations, algebraically identical to the reference's advective
// forms on the dealiased, solenoidal subspace (F_min table):
//   hydro      : inverse u (ndim)          -> forward u_i u_j (sym.)                 physics.py:527-599
//   Boussinesq : inverse u, T              -> forward u_i u_j, u_j T                 physics.py:664-712
//   MHD        : inverse u, B              -> forward u_i u_j - B_i B_j/(4 pi rho0), u x B   physics.py:770-819
// Assembly in k-space:  du/dt = P[-i k_j T_ij (+ g alpha T e_dir)],  P = I - k k/k^2  (k^2(0) := 1,
// physics.py:588-599, 407-416);  dT/dt = -i k_j (u_j T) - beta u_dir;  dB/dt = i k x E (curlX, :374-405).
#pragma once
#include "fft_core.cuh"

namespace ddl {

struct PhysConst {
    double inv_fpr;    // 1 / (4 pi rho0)                      (physics.py:791)
    double g_alpha;    // g * alpha_t                           (physics.py:691-694)
    double beta;       // stratification                        (physics.py:706-708)
    int bdir;          // buoyancy direction as component index (x=0, y=1, z=2)
};

enum PhysicsId { PH_HYDRO = 0, PH_BOUSSINESQ = 1, PH_MHD = 2 };

DDL_HD cplx mul_mi(cplx a) { return mk(a.y, -a.x); }   // * (-i)
DDL_HD cplx mul_pi(cplx a) { return mk(-a.y, a.x); }   // * (+i)
DDL_HD cplx lin3(double a, cplx x, double b, cplx y, double c, cplx z) {
    return mk(a * x.x + b * y.x + c * z.x, a * x.y + b * y.y + c * z.y);
}
DDL_HD cplx lin2(double a, cplx x, double b, cplx y) { return mk(a * x.x + b * y.x, a * x.y + b * y.y); }

// solenoidal projection of (Nx,Ny,Nz) given real k; k2 already has the k=0 -> 1 substitution
DDL_HD void project3(cplx& nx, cplx& ny, cplx& nz, double kx, double ky, double kz, double k2) {
    cplx kn = lin3(kx, nx, ky, ny, kz, nz);
    double s = 1.0 / k2;
    kn = scal(kn, s);
    nx = nx - scal(kn, kx); ny = ny - scal(kn, ky); nz = nz - scal(kn, kz);
}
DDL_HD void project2(cplx& nx, cplx& ny, double kx, double ky, double k2) {
    cplx kn = scal(lin2(kx, nx, ky, ny), 1.0 / k2);
    nx = nx - scal(kn, kx); ny = ny - scal(kn, ky);
}
DDL_HD double k2nz3(double kx, double ky, double kz) {
    double k2 = ky * ky + kz * kz + kx * kx;     // reference summation order y, z, x
    return k2 == 0.0 ? 1.0 : k2;
}
DDL_HD double k2nz2(double kx, double ky) {
    double k2 = kx * kx + ky * ky;               // 2-D dict order x, y
    return k2 == 0.0 ? 1.0 : k2;
}

// ------------------------------------------------------------------ 3-D hydro
struct Hydro3C {
    static constexpr int NI = 3, NO = 6, NS = 0, NC = 3, NDIM = 3, NG1 = 0;
    DDL_HD static void apply(const double* in, double* o, const PhysConst&) {
        const double u = in[0], v = in[1], w = in[2];
        o[0] = u * u; o[1] = u * v; o[2] = u * w; o[3] = v * v; o[4] = v * w; o[5] = w * w;
    }
    DDL_HD static void assemble(const cplx* P, const cplx*, cplx* D, double kx, double ky, double kz, const PhysConst&) {
        cplx nx = mul_mi(lin3(kx, P[0], ky, P[1], kz, P[2]));
        cplx ny = mul_mi(lin3(kx, P[1], ky, P[3], kz, P[4]));
        cplx nz = mul_mi(lin3(kx, P[2], ky, P[4], kz, P[5]));
        project3(nx, ny, nz, kx, ky, kz, k2nz3(kx, ky, kz));
        D[0] = nx; D[1] = ny; D[2] = nz;
    }
};

// ------------------------------------------------------------------ 3-D Boussinesq
struct Bouss3C {
    static constexpr int NI = 4, NO = 9, NS = 4, NC = 4, NDIM = 3, NG1 = 1;
    DDL_HD static void apply(const double* in, double* o, const PhysConst&) {
        const double u = in[0], v = in[1], w = in[2], T = in[3];
        o[0] = u * u; o[1] = u * v; o[2] = u * w; o[3] = v * v; o[4] = v * w; o[5] = w * w;
        o[6] = u * T; o[7] = v * T; o[8] = w * T;
    }
    DDL_HD static void assemble(const cplx* P, const cplx* S, cplx* D, double kx, double ky, double kz, const PhysConst& pc) {
        cplx n[3];
        n[0] = mul_mi(lin3(kx, P[0], ky, P[1], kz, P[2]));
        n[1] = mul_mi(lin3(kx, P[1], ky, P[3], kz, P[4]));
        n[2] = mul_mi(lin3(kx, P[2], ky, P[4], kz, P[5]));
        n[pc.bdir] = n[pc.bdir] + scal(S[3], pc.g_alpha);
        project3(n[0], n[1], n[2], kx, ky, kz, k2nz3(kx, ky, kz));
        D[0] = n[0]; D[1] = n[1]; D[2] = n[2];
        D[3] = mul_mi(lin3(kx, P[6], ky, P[7], kz, P[8])) - scal(S[pc.bdir], pc.beta);
    }
};

// ------------------------------------------------------------------ 3-D MHD
struct MHD3C {
    static constexpr int NI = 6, NO = 9, NS = 0, NC = 6, NDIM = 3, NG1 = 3;
    DDL_HD static void apply(const double* in, double* o, const PhysConst& pc) {
        const double u = in[0], v = in[1], w = in[2], a = in[3], b = in[4], c = in[5];
        const double f = pc.inv_fpr;
        o[0] = u * u - f * (a * a); o[1] = u * v - f * (a * b); o[2] = u * w - f * (a * c);
        o[3] = v * v - f * (b * b); o[4] = v * w - f * (b * c); o[5] = w * w - f * (c * c);
        o[6] = v * c - w * b; o[7] = w * a - u * c; o[8] = u * b - v * a;     // E = u x B
    }
    DDL_HD static void assemble(const cplx* P, const cplx*, cplx* D, double kx, double ky, double kz, const PhysConst&) {
        cplx nx = mul_mi(lin3(kx, P[0], ky, P[1], kz, P[2]));
        cplx ny = mul_mi(lin3(kx, P[1], ky, P[3], kz, P[4]));
        cplx nz = mul_mi(lin3(kx, P[2], ky, P[4], kz, P[5]));
        project3(nx, ny, nz, kx, ky, kz, k2nz3(kx, ky, kz));
        D[0] = nx; D[1] = ny; D[2] = nz;
        D[3] = mul_pi(lin2(ky, P[8], -kz, P[7]));
        D[4] = mul_pi(lin2(kz, P[6], -kx, P[8]));
        D[5] = mul_pi(lin2(kx, P[7], -ky, P[6]));
    }
};

// ------------------------------------------------------------------ 3-D, traceless flux: one forward transform fewer
// The momentum equation sees the flux only through P[-i k_j T_ij].  An isotropic part of T is a pressure: -i k_j (delta_ij s) = -i k_i s
// is parallel to k and the projector removes it mode by mode, whatever s is.  With s = T_zz the tensor T' = T - delta T_zz has
// T'_zz = 0 identically, so that component is never formed, transformed or read: 5 momentum products instead of 6 (the classic
// saving of pseudospectral Navier-Stokes codes).  Same derivative as Hydro3C / Bouss3C / MHD3C up to the rounding of the
// projection (measured: 1e-16 relative, tests/test_gpu_widen.py::test_traceless_flux_policies_agree); used by the one-rank
// ddl_rhs / ddl_rhs_stage (ddl_set_option("traceless_flux", 0) restores the six-product policies, which the slab-decomposed
// phases keep: their field counts are part of the exchange layout).
//   products: 0: T_xx - T_zz   1: T_xy   2: T_xz   3: T_yy - T_zz   4: T_yz   then the policy's other products
struct Hydro3T {
    static constexpr int NI = 3, NO = 5, NS = 0, NC = 3, NDIM = 3, NG1 = 0;
    DDL_HD static void apply(const double* in, double* o, const PhysConst&) {
        const double u = in[0], v = in[1], w = in[2];
        const double tzz = w * w;
        o[0] = u * u - tzz; o[1] = u * v; o[2] = u * w; o[3] = v * v - tzz; o[4] = v * w;
    }
    DDL_HD static void momentum(const cplx* P, cplx& nx, cplx& ny, cplx& nz, double kx, double ky, double kz) {
        nx = mul_mi(lin3(kx, P[0], ky, P[1], kz, P[2]));
        ny = mul_mi(lin3(kx, P[1], ky, P[3], kz, P[4]));
        nz = mul_mi(lin2(kx, P[2], ky, P[4]));
    }
    DDL_HD static void assemble(const cplx* P, const cplx*, cplx* D, double kx, double ky, double kz, const PhysConst&) {
        cplx nx, ny, nz;
        momentum(P, nx, ny, nz, kx, ky, kz);
        project3(nx, ny, nz, kx, ky, kz, k2nz3(kx, ky, kz));
        D[0] = nx; D[1] = ny; D[2] = nz;
    }
};

struct Bouss3T {
    static constexpr int NI = 4, NO = 8, NS = 4, NC = 4, NDIM = 3, NG1 = 1;
    DDL_HD static void apply(const double* in, double* o, const PhysConst& pc) {
        Hydro3T::apply(in, o, pc);
        const double T = in[3];
        o[5] = in[0] * T; o[6] = in[1] * T; o[7] = in[2] * T;
    }
    DDL_HD static void assemble(const cplx* P, const cplx* S, cplx* D, double kx, double ky, double kz, const PhysConst& pc) {
        cplx n[3];
        Hydro3T::momentum(P, n[0], n[1], n[2], kx, ky, kz);
        n[pc.bdir] = n[pc.bdir] + scal(S[3], pc.g_alpha);
        project3(n[0], n[1], n[2], kx, ky, kz, k2nz3(kx, ky, kz));
        D[0] = n[0]; D[1] = n[1]; D[2] = n[2];
        D[3] = mul_mi(lin3(kx, P[5], ky, P[6], kz, P[7])) - scal(S[pc.bdir], pc.beta);
    }
};

struct MHD3T {
    static constexpr int NI = 6, NO = 8, NS = 0, NC = 6, NDIM = 3, NG1 = 3;
    DDL_HD static void apply(const double* in, double* o, const PhysConst& pc) {
        const double u = in[0], v = in[1], w = in[2], a = in[3], b = in[4], c = in[5];
        const double f = pc.inv_fpr;
        const double tzz = w * w - f * (c * c);
        o[0] = (u * u - f * (a * a)) - tzz; o[1] = u * v - f * (a * b); o[2] = u * w - f * (a * c);
        o[3] = (v * v - f * (b * b)) - tzz; o[4] = v * w - f * (b * c);
        o[5] = v * c - w * b; o[6] = w * a - u * c; o[7] = u * b - v * a;     // E = u x B
    }
    DDL_HD static void assemble(const cplx* P, const cplx*, cplx* D, double kx, double ky, double kz, const PhysConst&) {
        cplx nx, ny, nz;
        Hydro3T::momentum(P, nx, ny, nz, kx, ky, kz);
        project3(nx, ny, nz, kx, ky, kz, k2nz3(kx, ky, kz));
        D[0] = nx; D[1] = ny; D[2] = nz;
        D[3] = mul_pi(lin2(ky, P[7], -kz, P[6]));
        D[4] = mul_pi(lin2(kz, P[5], -kx, P[7]));
        D[5] = mul_pi(lin2(kx, P[6], -ky, P[5]));
    }
};

// ------------------------------------------------------------------ 2-D
struct Hydro2C {
    static constexpr int NI = 2, NO = 3, NS = 0, NC = 2, NDIM = 2, NG1 = 0;
    DDL_HD static void apply(const double* in, double* o, const PhysConst&) {
        o[0] = in[0] * in[0]; o[1] = in[0] * in[1]; o[2] = in[1] * in[1];
    }
    DDL_HD static void assemble(const cplx* P, const cplx*, cplx* D, double kx, double ky, double, const PhysConst&) {
        cplx nx = mul_mi(lin2(kx, P[0], ky, P[1]));
        cplx ny = mul_mi(lin2(kx, P[1], ky, P[2]));
        project2(nx, ny, kx, ky, k2nz2(kx, ky));
        D[0] = nx; D[1] = ny;
    }
};

struct Bouss2C {
    static constexpr int NI = 3, NO = 5, NS = 3, NC = 3, NDIM = 2, NG1 = 1;
    DDL_HD static void apply(const double* in, double* o, const PhysConst&) {
        const double u = in[0], v = in[1], T = in[2];
        o[0] = u * u; o[1] = u * v; o[2] = v * v; o[3] = u * T; o[4] = v * T;
    }
    DDL_HD static void assemble(const cplx* P, const cplx* S, cplx* D, double kx, double ky, double, const PhysConst& pc) {
        cplx n[2];
        n[0] = mul_mi(lin2(kx, P[0], ky, P[1]));
        n[1] = mul_mi(lin2(kx, P[1], ky, P[2]));
        n[pc.bdir] = n[pc.bdir] + scal(S[2], pc.g_alpha);
        project2(n[0], n[1], kx, ky, k2nz2(kx, ky));
        D[0] = n[0]; D[1] = n[1];
        D[2] = mul_mi(lin2(kx, P[3], ky, P[4])) - scal(S[pc.bdir], pc.beta);
    }
};

struct MHD2C {
    static constexpr int NI = 4, NO = 4, NS = 0, NC = 4, NDIM = 2, NG1 = 2;
    DDL_HD static void apply(const double* in, double* o, const PhysConst& pc) {
        const double u = in[0], v = in[1], a = in[2], b = in[3];
        const double f = pc.inv_fpr;
        o[0] = u * u - f * (a * a); o[1] = u * v - f * (a * b); o[2] = v * v - f * (b * b);
        o[3] = u * b - v * a;                                                  // E_z
    }
    DDL_HD static void assemble(const cplx* P, const cplx*, cplx* D, double kx, double ky, double, const PhysConst&) {
        cplx nx = mul_mi(lin2(kx, P[0], ky, P[1]));
        cplx ny = mul_mi(lin2(kx, P[1], ky, P[2]));
        project2(nx, ny, kx, ky, k2nz2(kx, ky));
        D[0] = nx; D[1] = ny;
        D[2] = mul_pi(scal(P[3], ky));       // dBx/dt =  d_y E_z
        D[3] = mul_mi(scal(P[3], kx));       // dBy/dt = -d_x E_z
    }
};


// ------------------------------------------------------------------ advective-form variants
// The conservative forms above equal the reference's advective forms only for SOLENOIDAL u (and B):
//   d_j(u_i u_j) = (u.grad) u_i + u_i (div u),   d_j(u_j T) = u.grad T + T (div u),
//   ((curl B) x B)_i = d_j(B_i B_j) - B_i (div B) - d_i(B^2/2).
// The reference (physics.py:197-228, 527-599, 664-712, 770-819) evaluates the advective forms whatever the
// state, and never removes a compressive part the caller put there (e.g. the kz axis of turb_new's 3-D
// fields, init_cond.py:334-341).  For such states the host layer switches to these policies: the
// divergences theta_u = div u (theta_B = div B) are extra inverse-transform inputs (their spectra i k.u are
// written by ThetaF into caller-provided scratch arrays that follow the state in the pointer list), the
// products u_i theta_u (- B_i theta_B / 4 pi rho0, T theta_u) extra forward transforms, and the assembly
// adds them back.  Same results as the reference for ANY state; ~1.4x the transforms of the solenoidal path.
template <class BASE, int ND, int NB>
struct AdvOf {
    // inputs: BASE inputs, theta_u [, theta_B when NB == ND]; outputs: BASE outputs, c_i (ND) [, T theta when NB == 1]
    static constexpr int NTH = (NB == ND) ? 2 : 1;
    static constexpr int NI = BASE::NI + NTH, NO = BASE::NO + ND + (NB == 1 ? 1 : 0);
    static constexpr int NS = BASE::NS, NC = BASE::NC, NDIM = ND, NG1 = BASE::NG1;
    DDL_HD static void apply(const double* in, double* o, const PhysConst& pc) {
        BASE::apply(in, o, pc);
        const double thu = in[BASE::NI];
#pragma unroll
        for (int i = 0; i < ND; ++i) o[BASE::NO + i] = in[i] * thu;
        if (NB == ND) {
            const double thb = in[BASE::NI + 1] * pc.inv_fpr;
#pragma unroll
            for (int i = 0; i < ND; ++i) o[BASE::NO + i] -= in[ND + i] * thb;
        }
        if (NB == 1) o[BASE::NO + ND] = in[ND] * thu;
    }
    DDL_HD static void assemble(const cplx* P, const cplx* S, cplx* D, double kx, double ky, double kz, const PhysConst& pc) {
        BASE::assemble(P, S, D, kx, ky, kz, pc);
        // the momentum correction c goes through the projector, which is linear: D_u += P[c]
        cplx c[3] = {mk(0.0, 0.0), mk(0.0, 0.0), mk(0.0, 0.0)};
#pragma unroll
        for (int i = 0; i < ND; ++i) c[i] = P[BASE::NO + i];
        if (ND == 3) project3(c[0], c[1], c[2], kx, ky, kz, k2nz3(kx, ky, kz));
        else project2(c[0], c[1], kx, ky, k2nz2(kx, ky));
#pragma unroll
        for (int i = 0; i < ND; ++i) D[i] = D[i] + c[i];
        if (NB == 1) D[ND] = D[ND] + P[BASE::NO + ND];
    }
};
typedef AdvOf<Hydro3C, 3, 0> Hydro3A;
typedef AdvOf<Bouss3C, 3, 1> Bouss3A;
typedef AdvOf<MHD3C, 3, 3> MHD3A;
typedef AdvOf<Hydro2C, 2, 0> Hydro2A;
typedef AdvOf<Bouss2C, 2, 1> Bouss2A;
typedef AdvOf<MHD2C, 2, 2> MHD2A;

// identity policy for the plain real<->complex transforms
struct PhysNone {
    static constexpr int NI = 1, NO = 1, NS = 0, NC = 1, NDIM = 0, NG1 = 0;
    DDL_HD static void apply(const double* in, double* o, const PhysConst&) { o[0] = in[0]; }
};

}  // namespace ddl
