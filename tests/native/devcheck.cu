// devcheck -- a torch-free, Python-free device check of libddl_b200.so through its C ABI (include/ddl.h).
//
// Test infrastructure.  Purpose: (1) prove the C ABI is usable from plain C/C++ (what a cgo / ctypes /
// Cython binding of the reference would call), (2) check the device-only parts of the reductions and
// of the CFL capture against straightforward evaluations, in a few seconds of GPU time:
//   A  ddl_reduce_invariants (retained-only and full sweeps) vs a host loop over the downloaded state
//   B  ddl_reduce_max_square / ddl_rhs_capture_max vs ddl_backward of every component + host max
//   C  ddl_rhs: every x-pass variant and the generic tile kernels against each other
//   D  ddl_step_array / ddl_dealias_array (the Cython kernels' own signatures) and the shearing-box transform (ddl_set_shear)
//   T  (size argument >= 256) CUDA-event timings of the RHS per x-pass variant and of the reductions
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -I include tests/native/devcheck.cu \
//        -L dedalus-1.0_b200/dedalus/_lib -lddl_b200 -o tests/native/_build/devcheck
//   g++ -x c++ -DDEVCHECK_EMUL ... -lddl_emul      (same checks against the host-emulation build)
//
//   devcheck [n_check=64] [n_time=0] [outfile] [option=value ...]   e.g.  devcheck 0 512 out.txt reps=5 xfused_variant=1
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "ddl.h"

#ifndef DEVCHECK_EMUL
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { say("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(3); } } while (0)
#endif

static FILE* g_out = nullptr;
static int g_fail = 0;
static int g_vmask = 0x3f;    // x-pass variants the sweeps of sections C and T visit (bit v = variant v; argument vmask=...)
static int g_quick = 0;        // quick=1: section C (compared through 24 sums) and the variant / RK4-step timings only (kernel experiments on a
                               // short GPU budget); quick=2: the same with the element-by-element comparison
static int g_assemble_variant = 3;
static int g_strided_two = 1;
static int g_variant = 0;      // x-pass variant in force outside the variant sweeps (set by an xfused_variant=V argument)
template <class... A> static void say(const char* fmt, A... a) {
    printf(fmt, a...); fflush(stdout);
    if (g_out) { fprintf(g_out, fmt, a...); fflush(g_out); }
}
static void say(const char* s) { say("%s", s); }
#define DDL(x) do { int rc_ = (x); if (rc_ != 0) { say("ddl error %d: %s  [%s]\n", rc_, ddl_last_error(), #x); exit(2); } } while (0)

// ---------------------------------------------------------------- memory + timing shims
static void* dmalloc(size_t bytes) {
#ifndef DEVCHECK_EMUL
    void* p; CK(cudaMalloc(&p, bytes)); return p;
#else
    return malloc(bytes);
#endif
}
static void dfree(void* p) {
#ifndef DEVCHECK_EMUL
    cudaFree(p);
#else
    free(p);
#endif
}
static void d2h(void* h, const void* d, size_t bytes) {
#ifndef DEVCHECK_EMUL
    CK(cudaMemcpy(h, d, bytes, cudaMemcpyDeviceToHost));
#else
    memcpy(h, d, bytes);
#endif
}
static void dzero(void* d, size_t bytes) {
#ifndef DEVCHECK_EMUL
    CK(cudaMemset(d, 0, bytes));
#else
    memset(d, 0, bytes);
#endif
}
static void dsync() {
#ifndef DEVCHECK_EMUL
    CK(cudaDeviceSynchronize());
#endif
}

// x-space noise: a hash of the index, in [-1, 1)
#ifndef DEVCHECK_EMUL
__host__ __device__
#endif
static inline double hash01(unsigned long long i, unsigned long long seed) {
    unsigned long long z = i * 0x9E3779B97F4A7C15ull + seed * 0xBF58476D1CE4E5B9ull + 0x94D049BB133111EBull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z ^= z >> 31;
    return (double)(z >> 11) * (2.0 / 9007199254740992.0) - 1.0;
}
#ifndef DEVCHECK_EMUL
__global__ void fill_kernel(double* x, long long n, unsigned long long seed) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) x[i] = hash01(i, seed);
}
#endif
static void fill(double* x, long long n, unsigned long long seed) {
#ifndef DEVCHECK_EMUL
    fill_kernel<<<148 * 8, 256>>>(x, n, seed);
    CK(cudaGetLastError());
#else
    for (long long i = 0; i < n; ++i) x[i] = hash01(i, seed);
#endif
}

struct Timer {
#ifndef DEVCHECK_EMUL
    cudaEvent_t a, b;
    Timer() { cudaEventCreate(&a); cudaEventCreate(&b); }
    void start() { cudaEventRecord(a, 0); }
    double stop_ms() { cudaEventRecord(b, 0); cudaEventSynchronize(b); float ms; cudaEventElapsedTime(&ms, a, b); return ms; }
#else
    void start() {}
    double stop_ms() { return 0.0; }
#endif
};

// ---------------------------------------------------------------- a cubic MHD problem
struct Problem {
    int n; long long nk, nx3;
    std::vector<double> k;          // wavenumbers, fftfreq order with the Nyquist positive (representations.py:204-233)
    std::vector<uint8_t> keep;      // 2/3 rule: |k| < 2/3 kny survives (dealias_cy_3d.pyx:34-36)
    std::vector<double> kx; std::vector<uint8_t> keepx;
    ddl_plan* plan = nullptr;
    void* state[8]; void* deriv[6]; void* deriv2[6];
    void* work = nullptr; size_t work_bytes = 0;
    double* xbuf = nullptr;         // one x-space array
    double* dout = nullptr;         // device scratch for reduction outputs
    ddl_phys_params prm;

    explicit Problem(int n_, bool second_deriv) : n(n_) {
        nk = (long long)n * n * (n / 2 + 1); nx3 = (long long)n * n * n;
        k.resize(n); keep.resize(n); kx.resize(n / 2 + 1); keepx.resize(n / 2 + 1);
        const double kny = n / 2.0;
        for (int i = 0; i < n; ++i) { k[i] = (i <= n / 2) ? i : i - n; keep[i] = std::fabs(k[i]) < 2.0 / 3.0 * kny; }
        for (int i = 0; i <= n / 2; ++i) { kx[i] = i; keepx[i] = std::fabs(kx[i]) < 2.0 / 3.0 * kny; }
        int64_t shape[3] = {n, n, n};
        DDL(ddl_plan_create(&plan, 3, shape, kx.data(), k.data(), k.data(), keepx.data(), keep.data(), keep.data()));
        work_bytes = ddl_rhs_workspace_bytes(plan, DDL_MHD);
        size_t wb = ddl_workspace_bytes(plan, 1, 1);
        if (wb > work_bytes) work_bytes = wb;
        work = dmalloc(work_bytes);
        for (int c = 0; c < 6; ++c) { state[c] = dmalloc(nk * 16); deriv[c] = dmalloc(nk * 16); deriv2[c] = second_deriv ? dmalloc(nk * 16) : nullptr; }
        state[6] = state[7] = nullptr;
        xbuf = (double*)dmalloc(nx3 * 8);
        dout = (double*)dmalloc(64 * 8);
        prm.rho0 = 1.0; prm.g = 1.0; prm.alpha_t = 1.0; prm.beta = 1.0; prm.boussinesq_dir = 2; prm.reserved = 0;
        // state: noise -> forward (normalised, dealiased, Hermitian-consistent)
        for (int c = 0; c < 6; ++c) {
            fill(xbuf, nx3, 17 + c);
            DDL(ddl_forward(plan, xbuf, state[c], work, work_bytes, nullptr));
        }
        dsync();
    }
    ~Problem() {
        for (int c = 0; c < 6; ++c) { dfree(state[c]); dfree(deriv[c]); if (deriv2[c]) dfree(deriv2[c]); }
        dfree(work); dfree(xbuf); dfree(dout);
        ddl_plan_destroy(plan);
    }
};

static void verdict(const char* what, double err, double tol) {
    const bool ok = err <= tol;      // NaN fails
    if (!ok) g_fail++;
    say("  %-58s err %.3e  tol %.1e  %s\n", what, err, tol, ok ? "ok" : "FAIL");
}
static double rel(double a, double b) { return std::fabs(a - b) / (std::fabs(b) > 1e-300 ? std::fabs(b) : 1.0); }

static double max_rel_diff(Problem& P, void* const* a, void* const* b) {
    std::vector<double> ha(P.nk * 2), hb(P.nk * 2);
    double num = 0, den = 0;
    for (int c = 0; c < 6; ++c) {
        d2h(ha.data(), a[c], P.nk * 16); d2h(hb.data(), b[c], P.nk * 16);
        for (long long i = 0; i < P.nk * 2; ++i) { const double d = ha[i] - hb[i]; num += d * d; den += hb[i] * hb[i]; }
    }
    return (num == num && den > 0) ? std::sqrt(num / den) : NAN;
}

// quick mode: two sets of six spectra compared through the 24 sums of ddl_reduce_invariants (one sweep each on the device)
// instead of element by element on the host (13 GB of downloads per comparison at 512^3)
static double invariants_diff(Problem& P, void* const* a, void* const* b) {
    double ia[DDL_NINV], ib[DDL_NINV];
    void* a8[8] = {a[0], a[1], a[2], a[3], a[4], a[5], nullptr, nullptr};
    void* b8[8] = {b[0], b[1], b[2], b[3], b[4], b[5], nullptr, nullptr};
    DDL(ddl_reduce_invariants(P.plan, DDL_MHD, a8, 0, P.dout, nullptr)); dsync(); d2h(ia, P.dout, sizeof ia);
    DDL(ddl_reduce_invariants(P.plan, DDL_MHD, b8, 0, P.dout, nullptr)); dsync(); d2h(ib, P.dout, sizeof ib);
    double e = 0;
    for (int j = 0; j < DDL_NINV; ++j) {
        if (ia[j] != ia[j]) return NAN;
        e = std::fmax(e, std::fabs(ia[j] - ib[j]) / (std::fabs(ib[j]) + std::fabs(ib[0])));
    }
    return e;
}

static void check(int n) {
    say("== devcheck: MHD %d^3, %s\n", n, ddl_version());
    Problem P(n, true);
    const long long nk = P.nk;
    const int nh = n / 2 + 1;
    double brute[2] = {0, 0};
    double got[2];
    if (!g_quick) {
    // ---------------- A: invariants
    double inv_c[DDL_NINV], inv_f[DDL_NINV];
    DDL(ddl_reduce_invariants(P.plan, DDL_MHD, P.state, DDL_STAGE_RETAINED_ONLY, P.dout, nullptr));
    dsync(); d2h(inv_c, P.dout, sizeof inv_c);
    DDL(ddl_reduce_invariants(P.plan, DDL_MHD, P.state, 0, P.dout, nullptr));
    dsync(); d2h(inv_f, P.dout, sizeof inv_f);
    std::vector<std::vector<double>> h(6, std::vector<double>(nk * 2));
    for (int c = 0; c < 6; ++c) d2h(h[c].data(), P.state[c], nk * 16);
    double ref[DDL_NINV] = {0};
    for (int iy = 0; iy < n; ++iy) for (int iz = 0; iz < n; ++iz) for (int ix = 0; ix < nh; ++ix) {
        const long long i = ((long long)iy * n + iz) * nh + ix;
        const double kk[3] = {P.kx[ix], P.k[iy], P.k[iz]};
        const double w = ix == 0 ? 1.0 : 2.0;
        double re[6], im[6];
        for (int c = 0; c < 6; ++c) { re[c] = h[c][2 * i]; im[c] = h[c][2 * i + 1]; }
        double eu = 0, eb = 0, dur = 0, dui = 0, dbr = 0, dbi = 0, ub = 0;
        for (int c = 0; c < 3; ++c) {
            eu += re[c] * re[c] + im[c] * im[c]; eb += re[3 + c] * re[3 + c] + im[3 + c] * im[3 + c];
            dur += kk[c] * re[c]; dui += kk[c] * im[c]; dbr += kk[c] * re[3 + c]; dbi += kk[c] * im[3 + c];
            ub += re[c] * re[3 + c] + im[c] * im[3 + c];
            ref[DDL_INV_MSQ + c] += w * (re[c] * re[c] + im[c] * im[c]);
            ref[DDL_INV_MSQ + 3 + c] += w * (re[3 + c] * re[3 + c] + im[3 + c] * im[3 + c]);
        }
        ref[DDL_INV_EKIN] += w * 0.5 * eu; ref[DDL_INV_E2] += w * 0.5 * eb;
        ref[DDL_INV_DIV2] += w * (dur * dur + dui * dui); ref[DDL_INV_MAG_DIV2] += w * (dbr * dbr + dbi * dbi);
        ref[DDL_INV_DIV_SUM] += std::sqrt(dur * dur + dui * dui); ref[DDL_INV_MAG_DIV_SUM] += std::sqrt(dbr * dbr + dbi * dbi);
        ref[DDL_INV_HEL_CROSS] += w * ub;
        auto cross2 = [&](const double* r, const double* m) {      // |k x v|^2
            const double wxr = kk[1] * r[2] - kk[2] * r[1], wxi = kk[1] * m[2] - kk[2] * m[1];
            const double wyr = kk[2] * r[0] - kk[0] * r[2], wyi = kk[2] * m[0] - kk[0] * m[2];
            const double wzr = kk[0] * r[1] - kk[1] * r[0], wzi = kk[0] * m[1] - kk[1] * m[0];
            return wxr * wxr + wxi * wxi + wyr * wyr + wyi * wyi + wzr * wzr + wzi * wzi;
        };
        ref[DDL_INV_ENSTROPHY] += w * 0.5 * cross2(re, im);
        ref[DDL_INV_CURRENT2] += w * 0.5 * cross2(re + 3, im + 3);
    }
    say("A  ddl_reduce_invariants\n");
    const int idx[] = {DDL_INV_EKIN, DDL_INV_E2, DDL_INV_DIV_SUM, DDL_INV_MAG_DIV_SUM, DDL_INV_ENSTROPHY, DDL_INV_CURRENT2,
                       DDL_INV_HEL_CROSS, DDL_INV_MSQ, DDL_INV_MSQ + 1, DDL_INV_MSQ + 2, DDL_INV_MSQ + 3, DDL_INV_MSQ + 4, DDL_INV_MSQ + 5,
                       DDL_INV_DIV2, DDL_INV_MAG_DIV2};
    double e_c = 0, e_f = 0;
    for (int j : idx) { e_c = std::fmax(e_c, rel(inv_c[j], ref[j])); e_f = std::fmax(e_f, rel(inv_f[j], ref[j])); }
    if (inv_c[0] != inv_c[0]) e_c = NAN;
    if (inv_f[0] != inv_f[0]) e_f = NAN;
    say("  ekin %.15e (host %.15e)  emag %.15e  enstrophy %.15e\n", inv_c[0], ref[0], inv_c[1], inv_c[4]);
    verdict("retained-only sweep vs host loop (15 entries)", e_c, 1e-11);
    verdict("full sweep vs host loop", e_f, 1e-11);
    double e_cf = 0;
    for (int j = 0; j < DDL_NINV; ++j) e_cf = std::fmax(e_cf, std::fabs(inv_c[j] - inv_f[j]) / (std::fabs(inv_f[j]) + std::fabs(ref[0])));
    verdict("retained-only vs full sweep (all 24 entries)", e_cf, 1e-11);
    // bit reproducibility
    double again[DDL_NINV];
    DDL(ddl_reduce_invariants(P.plan, DDL_MHD, P.state, 0, P.dout, nullptr));
    dsync(); d2h(again, P.dout, sizeof again);
    verdict("full sweep twice: identical bits", memcmp(again, inv_f, sizeof again) ? 1.0 : 0.0, 0.0);

    // ---------------- B: max_square
    say("B  CFL maxima\n");
    {
        std::vector<double> hx(P.nx3);
        void* tmp = dmalloc(nk * 16);
        for (int c = 0; c < 6; ++c) {
#ifndef DEVCHECK_EMUL
            CK(cudaMemcpy(tmp, P.state[c], nk * 16, cudaMemcpyDeviceToDevice));
#else
            memcpy(tmp, P.state[c], nk * 16);
#endif
            DDL(ddl_backward(P.plan, tmp, P.xbuf, P.work, P.work_bytes, nullptr));
            dsync(); d2h(hx.data(), P.xbuf, P.nx3 * 8);
            double m = 0; for (double v : hx) m = std::fmax(m, v * v);
            brute[c / 3] = std::fmax(brute[c / 3], m);
        }
        dfree(tmp);
    }
    for (int fast = 1; fast >= 0; --fast) {
        ddl_set_option("fast_kernels", fast);
        DDL(ddl_reduce_max_square(P.plan, DDL_MHD, &P.prm, P.state, P.work, P.work_bytes, 0, P.dout, nullptr));
        dsync(); d2h(got, P.dout, sizeof got);
        say("  max u^2 %.15e (brute %.15e)  max B^2 %.15e (brute %.15e)\n", got[0], brute[0], got[1], brute[1]);
        verdict(fast ? "ddl_reduce_max_square, specialised kernels" : "ddl_reduce_max_square, generic tile kernels",
                std::fmax(rel(got[0], brute[0]), rel(got[1], brute[1])), 1e-12);
    }
    ddl_set_option("fast_kernels", 1);
    }   // !g_quick

    // ---------------- C: RHS variants
    say("C  ddl_rhs variants\n");
    ddl_set_option("fast_kernels", 0);
    DDL(ddl_rhs(P.plan, DDL_MHD, &P.prm, P.state, P.deriv2, P.work, P.work_bytes, DDL_RHS_ZERO_FILL, nullptr));
    dsync();
    ddl_set_option("fast_kernels", 1);
    for (int v = 0; v <= 15; ++v) {
        if (!(g_vmask >> v & 1)) continue;
        ddl_set_option("traceless_flux", v == 0);       // the launch variants exist for the six-product policies only
        ddl_set_option("xfused_variant", v);
        for (int c = 0; c < 6; ++c) dzero(P.deriv[c], nk * 16);
        dzero(P.dout, 16);
        DDL(ddl_rhs_capture_max(P.plan, (v == 0 || v == 4 || v == 5) ? P.dout : nullptr));     // variants 0, 4 and 5 (persistent) also with the capture on
        DDL(ddl_rhs(P.plan, DDL_MHD, &P.prm, P.state, P.deriv, P.work, P.work_bytes, DDL_RHS_ZERO_FILL, nullptr));
        DDL(ddl_rhs_capture_max(P.plan, nullptr));
        dsync();
        char label[96]; snprintf(label, sizeof label, g_quick == 1 ? "x-pass variant %d vs generic tile kernels (24 sums)" : "x-pass variant %d vs generic tile kernels (rel L2)", v);
        verdict(label, g_quick == 1 ? invariants_diff(P, P.deriv, P.deriv2) : max_rel_diff(P, P.deriv, P.deriv2), 1e-12);
        if (!g_quick && (v == 0 || v == 4 || v == 5)) {
            d2h(got, P.dout, sizeof got);
            verdict("maxima captured inside that RHS", std::fmax(rel(got[0], brute[0]), rel(got[1], brute[1])), 1e-12);
        }
    }
    ddl_set_option("xfused_variant", g_variant);
    ddl_set_option("traceless_flux", 1);
    for (int tv = 1; tv <= 2; ++tv) {
        // the y / z passes as two register butterflies around one trip through shared memory (csrc/fast_two.cuh)
        ddl_set_option("strided_two", tv);
        for (int c = 0; c < 6; ++c) dzero(P.deriv[c], nk * 16);
        DDL(ddl_rhs(P.plan, DDL_MHD, &P.prm, P.state, P.deriv, P.work, P.work_bytes, DDL_RHS_ZERO_FILL, nullptr));
        dsync();
        char label[96]; snprintf(label, sizeof label, g_quick == 1 ? "strided_two %d vs generic tile kernels (24 sums)" : "strided_two %d vs generic tile kernels (rel L2)", tv);
        verdict(label, g_quick == 1 ? invariants_diff(P, P.deriv, P.deriv2) : max_rel_diff(P, P.deriv, P.deriv2), 1e-12);
    }
    ddl_set_option("strided_two", g_strided_two);
    {
        // six momentum products instead of the default five (traceless flux, csrc/physics_ops.cuh): same derivative to round-off
        ddl_set_option("traceless_flux", 0);
        for (int c = 0; c < 6; ++c) dzero(P.deriv[c], nk * 16);
        DDL(ddl_rhs(P.plan, DDL_MHD, &P.prm, P.state, P.deriv, P.work, P.work_bytes, DDL_RHS_ZERO_FILL, nullptr));
        dsync();
        verdict(g_quick == 1 ? "traceless_flux 0 vs generic tile kernels, traceless (24 sums)" : "traceless_flux 0 vs generic tile kernels, traceless (rel L2)",
                g_quick == 1 ? invariants_diff(P, P.deriv, P.deriv2) : max_rel_diff(P, P.deriv, P.deriv2), 1e-12);
        ddl_set_option("traceless_flux", 1);
    }
    for (int chunk : {1, 3}) {
        if (g_quick) break;
        ddl_set_option("rhs_plane_chunk", chunk);
        for (int c = 0; c < 6; ++c) dzero(P.deriv[c], nk * 16);
        DDL(ddl_rhs(P.plan, DDL_MHD, &P.prm, P.state, P.deriv, P.work, P.work_bytes, DDL_RHS_ZERO_FILL, nullptr));
        dsync();
        char label[96]; snprintf(label, sizeof label, "rhs_plane_chunk %d vs generic tile kernels (rel L2)", chunk);
        verdict(label, max_rel_diff(P, P.deriv, P.deriv2), 1e-12);
    }
    ddl_set_option("rhs_plane_chunk", 0);
}

// ---------------------------------------------------------------- D: the one-to-one kernels and the shearing-box phase hook
static void h2d(void* d, const void* h, size_t bytes) {
#ifndef DEVCHECK_EMUL
    CK(cudaMemcpy(d, h, bytes, cudaMemcpyHostToDevice));
#else
    memcpy(d, h, bytes);
#endif
}

static void check_seams(int n) {
    say("D  kernels with the reference's own signatures, shearing-box transform (%d^3)\n", n);
    const int nh = n / 2 + 1;
    const long long nk = (long long)n * n * nh, nx3 = (long long)n * n * n;
    // ---- ddl_step_array (forward_step_cy_3d.pyx etd1 / etd2rk2) vs the formulas on the host
    {
        const long long cnt = 4096;
        std::vector<double> s(2 * cnt), d1(2 * cnt), d2(2 * cnt), IF(cnt), o(2 * cnt);
        for (long long i = 0; i < 2 * cnt; ++i) { s[i] = hash01(i, 1); d1[i] = hash01(i, 2); d2[i] = hash01(i, 3); }
        for (long long i = 0; i < cnt; ++i) IF[i] = (i % 7 == 0) ? 0.0 : -std::fabs(hash01(i, 4)) * ((i % 5 == 1) ? 0.4 : 40.0);
        void *ds = dmalloc(16 * cnt), *dd1 = dmalloc(16 * cnt), *dd2 = dmalloc(16 * cnt), *dout = dmalloc(16 * cnt);
        double* dIF = (double*)dmalloc(8 * cnt);
        h2d(ds, s.data(), 16 * cnt); h2d(dd1, d1.data(), 16 * cnt); h2d(dd2, d2.data(), 16 * cnt); h2d(dIF, IF.data(), 8 * cnt);
        const double dt = 0.02;
        double worst = 0;
        for (int kind : {DDL_ETD1, DDL_ETD2RK2}) {
            DDL(ddl_step_array(kind, 3, cnt, ds, dout, dd1, dd2, dIF, dt, nullptr));
            dsync(); d2h(o.data(), dout, 16 * cnt);
            for (long long i = 0; i < cnt; ++i) {
                const double Z = IF[i] * dt;
                if (std::fabs(Z) < 0.5 && Z != 0.0) continue;          // the series branch is pinned by the CPU goldens
                for (int c = 0; c < 2; ++c) {
                    const double S = s[2 * i + c], A = d1[2 * i + c], B = d2[2 * i + c];
                    double want;
                    if (Z == 0.0) want = S + dt * (kind == DDL_ETD1 ? A : B);
                    else {
                        const double f0 = std::exp(Z), f1 = (f0 - 1.0) / Z, f2 = (f1 - 1.0) / Z;
                        want = kind == DDL_ETD1 ? S * f0 + A * f1 * dt : S * f0 + (B - A) * 2.0 * f2 * dt + A * f1 * dt;
                    }
                    worst = std::fmax(worst, std::fabs(o[2 * i + c] - want) / (std::fabs(want) + 1.0));
                }
            }
        }
        verdict("ddl_step_array (etd1, etd2rk2) vs the closed forms", worst, 1e-14);
        dfree(ds); dfree(dd1); dfree(dd2); dfree(dout); dfree(dIF);
    }
    // ---- shearing box: a plane wave cos(a x + b y) transforms to a single coefficient at ky0 = b + S t a (S t a an integer)
    {
        std::vector<double> k(n), kx(nh);
        std::vector<uint8_t> keep(n), keepx(nh), full(n, 1);
        const double kny = n / 2.0;
        for (int i = 0; i < n; ++i) { k[i] = (i <= n / 2) ? i : i - n; keep[i] = std::fabs(k[i]) < 2.0 / 3.0 * kny; }
        for (int i = 0; i < nh; ++i) { kx[i] = i; keepx[i] = std::fabs(kx[i]) < 2.0 / 3.0 * kny; }
        int64_t shape[3] = {n, n, n};
        ddl_plan* pl = nullptr;
        DDL(ddl_plan_create(&pl, 3, shape, kx.data(), k.data(), k.data(), keepx.data(), full.data(), keep.data()));
        const size_t wb = ddl_workspace_bytes(pl, 1, 1);
        void* work = dmalloc(wb);
        double* x = (double*)dmalloc(nx3 * 8);
        void* kk = dmalloc(nk * 16);
        const int a = 3, b = 2;
        const double S = 1.5, t = 2.0 / (S * a) * 1.0;           // S t a = 2: the wave sits at stored ky index b + 2
        const double two_pi = 6.283185307179586476925286766559, dy = two_pi / n;
        std::vector<double> hx(nx3);
        for (int iz = 0; iz < n; ++iz) for (int iy = 0; iy < n; ++iy) for (int ix = 0; ix < n; ++ix)
            hx[((long long)iz * n + iy) * n + ix] = std::cos(a * (two_pi * ix / n) + b * (iy * dy));
        h2d(x, hx.data(), nx3 * 8);
        DDL(ddl_set_shear(pl, 1, S, t, dy));
        DDL(ddl_forward(pl, x, kk, work, wb, nullptr));
        dsync();
        std::vector<double> hk(2 * nk);
        d2h(hk.data(), kk, nk * 16);
        const long long hit = ((long long)(b + 2) * n + 0) * nh + a;       // [ky][kz][kx]
        double off = 0;
        for (long long i = 0; i < nk; ++i) if (i != hit) off = std::fmax(off, std::hypot(hk[2 * i], hk[2 * i + 1]));
        verdict("sheared forward of a plane wave: the one coefficient", std::hypot(hk[2 * hit] - 0.5, hk[2 * hit + 1]), 1e-13);
        verdict("sheared forward of a plane wave: everything else", off, 1e-13);
        DDL(ddl_backward(pl, kk, x, work, wb, nullptr));
        dsync();
        std::vector<double> hb(nx3);
        d2h(hb.data(), x, nx3 * 8);
        double e = 0;
        for (long long i = 0; i < nx3; ++i) e = std::fmax(e, std::fabs(hb[i] - hx[i]));
        verdict("sheared backward returns the wave", e, 1e-12);
        // ---- ddl_dealias_array, dense-ky branch, vs a host loop
        std::vector<double> kyd((size_t)n * nh), data(2 * nk);
        for (int iy = 0; iy < n; ++iy) for (int ix = 0; ix < nh; ++ix) kyd[(size_t)iy * nh + ix] = k[iy] - 0.37 * kx[ix];
        for (long long i = 0; i < 2 * nk; ++i) data[i] = hash01(i, 9);
        double *dkx = (double*)dmalloc(8 * nh), *dky = (double*)dmalloc(8 * (size_t)n * nh), *dkz = (double*)dmalloc(8 * n);
        h2d(dkx, kx.data(), 8 * nh); h2d(dky, kyd.data(), 8 * (size_t)n * nh); h2d(dkz, k.data(), 8 * n); h2d(kk, data.data(), 16 * nk);
        int64_t kshape[3] = {n, n, nh};
        const double knyq[3] = {kny, kny, kny};
        DDL(ddl_dealias_array(3, kshape, kk, dkx, dky, dkz, 1, knyq, nullptr));
        dsync(); d2h(hk.data(), kk, nk * 16);
        long long bad = 0, zeroed = 0;
        const double cut = 2.0 / 3.0 * kny;
        for (int iy = 0; iy < n; ++iy) for (int iz = 0; iz < n; ++iz) for (int ix = 0; ix < nh; ++ix) {
            const long long i = ((long long)iy * n + iz) * nh + ix;
            const double ky = kyd[(size_t)iy * nh + ix];
            const bool z = kx[ix] >= cut || kx[ix] <= -cut || ky >= cut || ky <= -cut || k[iz] >= cut || k[iz] <= -cut;
            zeroed += z;
            const double wr = z ? 0.0 : data[2 * i], wi = z ? 0.0 : data[2 * i + 1];
            if (hk[2 * i] != wr || hk[2 * i + 1] != wi) bad++;
        }
        verdict("ddl_dealias_array, dense ky: entries differing from the host loop", (double)bad, 0.0);
        if (zeroed == 0 || zeroed == nk) { g_fail++; say("  dealias_array: degenerate mask  FAIL\n"); }
        dfree(dkx); dfree(dky); dfree(dkz); dfree(work); dfree(x); dfree(kk);
        ddl_plan_destroy(pl);
    }
    // ---- ddl_stage over the retained modes + ddl_stage_outside over the rest == ddl_stage over everything (bit for bit), for a
    //      state with content outside the mask and a derivative without
    {
        Problem P(n, false);
        std::vector<double> junk(2 * nk);
        d2h(junk.data(), P.state[0], nk * 16);
        for (long long i = 0; i < nk; ++i) { junk[2 * i] += 0.25 * hash01(i, 31); junk[2 * i + 1] += 0.25 * hash01(i, 32); }   // everywhere, also outside
        h2d(P.state[0], junk.data(), nk * 16);
        void* start[1] = {P.state[0]}; void* d1[1] = {P.state[1]};          // state[1] is dealiased: a legitimate derivative
        void* full[1] = {P.deriv[0]}; void* split[1] = {P.deriv[1]};
        const double coeff[1] = {0.37}, dt = 0.013;
        double worst = 0;
        for (int kind : {DDL_ETD1, DDL_EULER}) {
            dzero(full[0], nk * 16); dzero(split[0], nk * 16);
            DDL(ddl_stage(P.plan, kind, 1, start, full, d1, nullptr, coeff, 1, dt, 0, nullptr));
            DDL(ddl_stage(P.plan, kind, 1, start, split, d1, nullptr, coeff, 1, dt, DDL_STAGE_RETAINED_ONLY, nullptr));
            DDL(ddl_stage_outside(P.plan, kind, 1, start, split, coeff, 1, dt, nullptr));
            dsync();
            std::vector<double> a(2 * nk), b(2 * nk);
            d2h(a.data(), full[0], nk * 16); d2h(b.data(), split[0], nk * 16);
            long long bad = 0;
            for (long long i = 0; i < 2 * nk; ++i) if (a[i] != b[i]) bad++;
            worst = std::fmax(worst, (double)bad);
        }
        verdict("retained sweep + ddl_stage_outside vs the full sweep (entries differing)", worst, 0.0);
    }
}

// one RK4 step the way the Python integrator issues it once the state is dealiased (time_step.py RK4._advance_fused):
// four ddl_rhs_stage calls, the spectral assembly fused with the stage update, y advanced in place
static void rk4_fused_step(Problem& P, void* const* tmp, void* const* total, const double* coeff, double dt, void* const* final_out = nullptr) {
    ddl_stage_fuse f;
    memset(&f, 0, sizeof f);
    f.total = total; f.coeff = coeff; f.visc_order = 1; f.kind = DDL_FUSE_RK4;
    const int flags = DDL_STAGE_RETAINED_ONLY * 0;      // ddl_rhs flags: state already dealiased, nothing to zero-fill
    struct { void* const* in; void* const* out; double wdiv, h; int first, last; } st[4] = {
        {P.state, tmp, 6., dt / 2., 1, 0}, {tmp, tmp, 3., dt / 2., 0, 0}, {tmp, tmp, 3., dt, 0, 0}, {tmp, final_out ? final_out : P.state, 6., dt, 0, 1}};
    for (int i = 0; i < 4; ++i) {
        f.y = P.state; f.out = st[i].out; f.wdiv = st[i].wdiv; f.dt_step = st[i].h; f.first = st[i].first; f.last = st[i].last;
        DDL(ddl_rhs_stage(P.plan, DDL_MHD, &P.prm, st[i].in, P.work, P.work_bytes, flags, &f, nullptr));
    }
}

static void timing(int n, int reps) {
    say("== timing: MHD %d^3 (CUDA events, ms per call, best of %d after 1 warm-up)\n", n, reps);
    Problem P(n, false);
    Timer t;
    ddl_set_option("traceless_flux", 0);       // the x-pass launch variants: six-product policies, like the earlier records of this sweep
    for (int v = 0; v <= 15; ++v) {
        if (!(g_vmask >> v & 1)) continue;
        ddl_set_option("xfused_variant", v);
        double best = 1e30;
        for (int r = 0; r <= reps; ++r) {
            t.start();
            DDL(ddl_rhs(P.plan, DDL_MHD, &P.prm, P.state, P.deriv, P.work, P.work_bytes, 0, nullptr));
            const double ms = t.stop_ms();
            if (r > 0 && ms < best) best = ms;
        }
        ddl_profile_enable(1);
        DDL(ddl_rhs(P.plan, DDL_MHD, &P.prm, P.state, P.deriv, P.work, P.work_bytes, 0, nullptr));
        static char pbuf[1 << 14];
        DDL(ddl_profile_report(pbuf, sizeof pbuf));
        ddl_profile_enable(0);
        const char* xf = strstr(pbuf, "\"x_fused\"");
        say("  ddl_rhs, x-pass variant %d: %.3f ms   %.40s\n", v, best, xf ? xf : "");
    }
    ddl_set_option("xfused_variant", g_variant);
    ddl_set_option("traceless_flux", 1);
    for (int tv = 0; tv <= 2; ++tv) {
        ddl_set_option("strided_two", tv);
        double best = 1e30;
        for (int r = 0; r <= reps; ++r) {
            t.start();
            DDL(ddl_rhs(P.plan, DDL_MHD, &P.prm, P.state, P.deriv, P.work, P.work_bytes, 0, nullptr));
            const double ms = t.stop_ms();
            if (r > 0 && ms < best) best = ms;
        }
        ddl_profile_enable(1);
        DDL(ddl_rhs(P.plan, DDL_MHD, &P.prm, P.state, P.deriv, P.work, P.work_bytes, 0, nullptr));
        static char pbuf2[1 << 14];
        DDL(ddl_profile_report(pbuf2, sizeof pbuf2));
        ddl_profile_enable(0);
        say("  ddl_rhs, strided_two %d: %.3f ms   %s\n", tv, best, pbuf2);
    }
    ddl_set_option("strided_two", g_strided_two);
    for (int tf = 0; tf <= 1; ++tf) {
        ddl_set_option("traceless_flux", tf);
        double best = 1e30;
        for (int r = 0; r <= reps; ++r) {
            t.start();
            DDL(ddl_rhs(P.plan, DDL_MHD, &P.prm, P.state, P.deriv, P.work, P.work_bytes, 0, nullptr));
            const double ms = t.stop_ms();
            if (r > 0 && ms < best) best = ms;
        }
        ddl_profile_enable(1);
        DDL(ddl_rhs(P.plan, DDL_MHD, &P.prm, P.state, P.deriv, P.work, P.work_bytes, 0, nullptr));
        static char pbuf3[1 << 14];
        DDL(ddl_profile_report(pbuf3, sizeof pbuf3));
        ddl_profile_enable(0);
        say("  ddl_rhs, traceless_flux %d (%d product fields): %.3f ms   %s\n", tf, tf ? 8 : 9, best, pbuf3);
    }
    ddl_set_option("traceless_flux", 1);
    // opt-in L2-residency experiment: y_inv -> x -> y_fwd over chunks of z-planes (include/ddl.h "rhs_plane_chunk")
    for (int chunk : {1, 2, 4, 8}) {
        if (g_quick) break;
        ddl_set_option("rhs_plane_chunk", chunk);
        double best = 1e30;
        for (int r = 0; r <= reps; ++r) {
            t.start();
            DDL(ddl_rhs(P.plan, DDL_MHD, &P.prm, P.state, P.deriv, P.work, P.work_bytes, 0, nullptr));
            const double ms = t.stop_ms();
            if (r > 0 && ms < best) best = ms;
        }
        say("  ddl_rhs, rhs_plane_chunk %d: %.3f ms\n", chunk, best);
    }
    ddl_set_option("rhs_plane_chunk", 0);
    for (int mode = 0; mode < 3; ++mode) {
        if (g_quick) break;
        double best = 1e30;
        for (int r = 0; r <= reps; ++r) {
            t.start();
            if (mode == 0) DDL(ddl_reduce_invariants(P.plan, DDL_MHD, P.state, DDL_STAGE_RETAINED_ONLY, P.dout, nullptr));
            if (mode == 1) DDL(ddl_reduce_invariants(P.plan, DDL_MHD, P.state, 0, P.dout, nullptr));
            if (mode == 2) DDL(ddl_reduce_max_square(P.plan, DDL_MHD, &P.prm, P.state, P.work, P.work_bytes, 0, P.dout, nullptr));
            const double ms = t.stop_ms();
            if (r > 0 && ms < best) best = ms;
        }
        say("  %s: %.3f ms\n", mode == 0 ? "invariants, retained-only sweep" : mode == 1 ? "invariants, full sweep" : "max_square (inverse half + capture)", best);
    }
    // the bench's step, natively: RK4 with fused stages (the deriv arrays double as `total` and the stage state)
    {
        void* tmp[6]; void* total[6];
        for (int c = 0; c < 6; ++c) { total[c] = P.deriv[c]; tmp[c] = dmalloc(P.nk * 16); dzero(tmp[c], P.nk * 16); dzero(total[c], P.nk * 16); }
        const double coeff[6] = {1e-3, 1e-3, 1e-3, 1e-3, 1e-3, 1e-3};
        double inv[DDL_NINV];
        DDL(ddl_reduce_invariants(P.plan, DDL_MHD, P.state, 0, P.dout, nullptr)); dsync(); d2h(inv, P.dout, sizeof inv);
        const double dt = 0.2 * (2 * 3.14159265358979323846 / n) / std::sqrt(2.0 * inv[0] + 1e-300);      // ~ CFL 0.2 on the rms speed
        // the fused assembly + stage kernel in its launch variants (ddl_set_option("assemble_variant", v)): same step from the same
        // state into a side buffer, compared with variant 0, and timed
        {
            void* ref6[6]; void* out6[6];
            for (int c = 0; c < 6; ++c) { ref6[c] = dmalloc(P.nk * 16); out6[c] = dmalloc(P.nk * 16); }
            for (int av = 0; av <= 4; ++av) {
                ddl_set_option("assemble_variant", av);
                double b = 1e30;
                for (int r = 0; r <= reps; ++r) {
                    t.start();
                    rk4_fused_step(P, tmp, total, coeff, dt, av == 0 ? ref6 : out6);
                    const double ms = t.stop_ms();
                    if (r > 0 && ms < b) b = ms;
                }
                ddl_profile_enable(1);
                rk4_fused_step(P, tmp, total, coeff, dt, av == 0 ? ref6 : out6);
                static char pb[1 << 16];
                DDL(ddl_profile_report(pb, sizeof pb));
                ddl_profile_enable(0);
                const char* as = strstr(pb, "\"assemble_stage\"");
                say("  RK4 step, assemble_variant %d: %.3f ms   %.48s\n", av, b, as ? as : "");
                if (av > 0) { char label[96]; snprintf(label, sizeof label, "assemble_variant %d vs 0: state after the step (rel L2)", av); verdict(label, g_quick == 1 ? invariants_diff(P, out6, ref6) : max_rel_diff(P, out6, ref6), 1e-14); }
            }
            ddl_set_option("assemble_variant", g_assemble_variant);
            for (int c = 0; c < 6; ++c) { dfree(ref6[c]); dfree(out6[c]); }
        }
        double best = 1e30;
        for (int r = 0; r <= reps; ++r) {
            t.start();
            rk4_fused_step(P, tmp, total, coeff, dt);
            const double ms = t.stop_ms();
            if (r > 0 && ms < best) best = ms;
        }
        double after[DDL_NINV];
        DDL(ddl_reduce_invariants(P.plan, DDL_MHD, P.state, 0, P.dout, nullptr)); dsync(); d2h(after, P.dout, sizeof after);
        say("  RK4 step, fused stages (4 x ddl_rhs_stage): %.3f ms = %.3e mode-stage updates/s   (ekin %.6e -> %.6e, div2 %.1e)\n",
            best, 4.0 * P.nk / (best * 1e-3), inv[0], after[0], after[DDL_INV_DIV2]);
        if (!(after[0] == after[0]) || after[0] > 4 * inv[0]) { g_fail++; say("  RK4 step: energy not sane  FAIL\n"); }
        // per-kernel CUDA-event times of one such step (ddl_profile_*)
        ddl_profile_enable(1);
        rk4_fused_step(P, tmp, total, coeff, dt);
        static char buf[1 << 16];
        DDL(ddl_profile_report(buf, sizeof buf));
        ddl_profile_enable(0);
        say("  per-kernel totals of one step: %s\n", buf);
        for (int c = 0; c < 6; ++c) dfree(tmp[c]);
    }
    dsync();
}

// ---------------------------------------------------------------- E: ddl_copy_boxes, the retained box of a spectrum across PCIe
static void* hpinned(size_t bytes) {
#ifndef DEVCHECK_EMUL
    void* p; CK(cudaMallocHost(&p, bytes)); return p;
#else
    return malloc(bytes);
#endif
}
static void hfree(void* p) {
#ifndef DEVCHECK_EMUL
    cudaFreeHost(p);
#else
    free(p);
#endif
}
static void check_boxes(int n, bool time_it) {
    say("E  ddl_copy_boxes: retained box of a (ky, kz, kx) spectrum between pinned host memory and the device (%d^3)\n", n);
    Problem P(n, false);
    const int nh = n / 2 + 1;
    const long long nk = P.nk;
    // runs of kept indices per axis, as the host layer derives them from the plan's masks
    std::vector<std::pair<int, int>> ry, rx;
    for (int i = 0; i < n;) { if (!P.keep[i]) { ++i; continue; } int j = i; while (j < n && P.keep[j]) ++j; ry.push_back({i, j}); i = j; }
    for (int i = 0; i < nh;) { if (!P.keepx[i]) { ++i; continue; } int j = i; while (j < nh && P.keepx[j]) ++j; rx.push_back({i, j}); i = j; }
    std::vector<int64_t> boxes;
    long long kept = 0;
    for (auto& a : ry) for (auto& b : ry) for (auto& c : rx) {
        boxes.insert(boxes.end(), {a.first, a.second, b.first, b.second, c.first, c.second});
        kept += (long long)(a.second - a.first) * (b.second - b.first) * (c.second - c.first);
    }
    const int nbox = (int)boxes.size() / 6;
    int64_t shape[3] = {n, n, nh};
    double* full = (double*)hpinned(nk * 16);
    double* part = (double*)hpinned(nk * 16);
    d2h(full, P.state[0], nk * 16);
    memset(part, 0, nk * 16);
    DDL(ddl_copy_boxes(part, P.state[0], shape, nbox, boxes.data(), 16, 0, nullptr));
    dsync();
    verdict("device -> host, boxes only: equals the full download (bytes differing)", memcmp(full, part, nk * 16) ? 1.0 : 0.0, 0.0);
    void* dev = dmalloc(nk * 16);
    dzero(dev, nk * 16);
    DDL(ddl_copy_boxes(dev, full, shape, nbox, boxes.data(), 16, 1, nullptr));
    dsync();
    d2h(part, dev, nk * 16);
    verdict("host -> device into a zeroed array: equals the spectrum", memcmp(full, part, nk * 16) ? 1.0 : 0.0, 0.0);
    say("  %d boxes, %.1f %% of the array\n", nbox, 100.0 * kept / nk);
    if (time_it) {
        Timer t;
        double best_d = 1e30, best_u = 1e30, best_f = 1e30;
        for (int r = 0; r < 4; ++r) {
            t.start(); DDL(ddl_copy_boxes(part, P.state[0], shape, nbox, boxes.data(), 16, 0, nullptr)); double ms = t.stop_ms(); if (ms < best_d) best_d = ms;
            t.start(); DDL(ddl_copy_boxes(dev, full, shape, nbox, boxes.data(), 16, 1, nullptr)); ms = t.stop_ms(); if (ms < best_u) best_u = ms;
#ifndef DEVCHECK_EMUL
            t.start(); CK(cudaMemcpyAsync(part, P.state[0], nk * 16, cudaMemcpyDeviceToHost, 0)); ms = t.stop_ms(); if (ms < best_f) best_f = ms;
#endif
        }
        say("  one component: boxes D2H %.3f ms (%.1f GB/s), boxes H2D %.3f ms (%.1f GB/s), full array D2H %.3f ms (%.1f GB/s)\n",
            best_d, kept * 16 / best_d / 1e6, best_u, kept * 16 / best_u / 1e6, best_f, nk * 16 / best_f / 1e6);
    }
    dfree(dev); hfree(full); hfree(part);
}

int main(int argc, char** argv) {
    // devcheck [n_check=64] [n_time=0] [outfile] [option=value ...]      options: ddl_set_option names, reps=R
    int n_check = 64, n_time = 0, reps = 3, pos = 0;
    for (int i = 1; i < argc; ++i) {
        const char* eq = strchr(argv[i], '=');
        if (eq) {
            char name[64]; snprintf(name, sizeof name, "%.*s", (int)(eq - argv[i]), argv[i]);
            const int val = (int)strtol(eq + 1, nullptr, 0);
            if (!strcmp(name, "reps")) reps = val;
            else if (!strcmp(name, "vmask")) g_vmask = val;
            else if (!strcmp(name, "quick")) g_quick = val;
            else { if (!strcmp(name, "xfused_variant")) g_variant = val; if (!strcmp(name, "assemble_variant")) g_assemble_variant = val; if (!strcmp(name, "strided_two")) g_strided_two = val; DDL(ddl_set_option(name, val)); }
        } else if (pos == 0) { n_check = atoi(argv[i]); pos++; }
        else if (pos == 1) { n_time = atoi(argv[i]); pos++; }
        else if (pos == 2) { g_out = fopen(argv[i], "w"); pos++; }
    }
    if (n_check > 0) { check(n_check); if (!g_quick) { check_seams(n_check < 64 ? n_check : 64); check_boxes(n_check, false); } }
    if (n_time > 0) { timing(n_time, reps); if (!g_quick) check_boxes(n_time, true); }
    say("devcheck: %s (%d failure%s)\n", g_fail ? "FAILED" : "all ok", g_fail, g_fail == 1 ? "" : "s");
    if (g_out) fclose(g_out);
    return g_fail ? 1 : 0;
}
