// C ABI (include/ddl.h): plan, transforms, fused RHS pipelines, stage updates.
#include <chrono>
#include <cstdarg>
#include <cmath>
#include <string>
#include <utility>
#include <vector>

#include "../../include/ddl.h"
#include "pointwise.cuh"
#include "reduce.cuh"
#include "tile_kernel.cuh"
#include "fast_kernels.cuh"
#include "fast_two.cuh"
#include "xfused_kernel.cuh"

namespace ddl {

// ---------------------------------------------------------------- errors / memory
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

// ---------------------------------------------------------------- launch accounting
struct ProfEvent {
    const char* name;
#if DDL_DEVICE_BUILD
    cudaEvent_t a, b;
#else
    double a, b;        // host emulation (tests only): wall-clock seconds, so that the report carries the labels of the path taken
#endif
};
#if !DDL_DEVICE_BUILD
static double prof_now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
#endif
static long long g_launches = 0;
static bool g_prof_on = false;
static std::vector<ProfEvent> g_prof;

void prof_begin(const char* name, ddl_stream_t stream) {
    ++g_launches;
#if DDL_DEVICE_BUILD
    if (!g_prof_on) return;
    ProfEvent e;
    e.name = name ? name : "?";
    cudaEventCreate(&e.a);
    cudaEventCreate(&e.b);
    cudaEventRecord(e.a, stream);
    g_prof.push_back(e);
#else
    (void)stream;
    if (!g_prof_on) return;
    ProfEvent e;
    e.name = name ? name : "?";
    e.a = e.b = prof_now();
    if (!g_prof.empty() && g_prof.back().b == g_prof.back().a) g_prof.back().b = e.a;   // emulated launchers without a prof_end
    g_prof.push_back(e);
#endif
}
void prof_end(ddl_stream_t stream) {
#if DDL_DEVICE_BUILD
    if (g_prof_on && !g_prof.empty()) cudaEventRecord(g_prof.back().b, stream);
#else
    (void)stream;
    if (g_prof_on && !g_prof.empty()) g_prof.back().b = prof_now();
#endif
}

void* dev_alloc(size_t bytes) {
#if DDL_DEVICE_BUILD
    void* p = nullptr;
    if (cudaMalloc(&p, bytes ? bytes : 16) != cudaSuccess) return nullptr;
    return p;
#else
    return malloc(bytes ? bytes : 16);
#endif
}
void dev_free(void* p) {
    if (!p) return;
#if DDL_DEVICE_BUILD
    cudaFree(p);
#else
    free(p);
#endif
}
int dev_upload(void* dst, const void* src, size_t bytes) {
#if DDL_DEVICE_BUILD
    return cudaMemcpy(dst, src, bytes, cudaMemcpyHostToDevice) == cudaSuccess ? 0 : -2;
#else
    memcpy(dst, src, bytes);
    return 0;
#endif
}

// ---------------------------------------------------------------- per-length kernels
#define DDL_DECL(N) int run_tile_##N(int, int, int, const TileParams&, int, ddl_stream_t);
DDL_DECL(8) DDL_DECL(16) DDL_DECL(32) DDL_DECL(64) DDL_DECL(128) DDL_DECL(256) DDL_DECL(512) DDL_DECL(1024)
DDL_DECL(2048) DDL_DECL(0)

// largest tile the generic kernel may ask for (227 KB of shared memory per CTA on sm_100)
static const size_t TILE_SMEM_MAX = 227 * 1024;

// Lengths with a translation unit of their own run the compile-time kernels; every other length
// (the reference transforms any N through FFTW) runs the runtime-length instantiation.
static int run_tile(int N, int mode, int dir, int phys, TileParams& p, int nthreads, ddl_stream_t s) {
    if ((size_t)N * p.ld * sizeof(cplx) > TILE_SMEM_MAX) {
        set_error("transform length %d with %d pencils per tile needs %zu bytes of shared memory (limit %zu)", N, p.ld,
                  (size_t)N * p.ld * sizeof(cplx), TILE_SMEM_MAX);
        return -1;
    }
    if (!p.sh_on) {        // the shearing-box phase hook lives in the runtime-length instantiation only
        switch (N) {
#define DDL_CASE(N) case N: return run_tile_##N(mode, dir, phys, p, nthreads, s);
            DDL_CASE(8) DDL_CASE(16) DDL_CASE(32) DDL_CASE(64) DDL_CASE(128) DDL_CASE(256) DDL_CASE(512)
            DDL_CASE(1024) DDL_CASE(2048)
        }
    }
    if (!rt_factor(N, p.rt)) {
        set_error("unsupported transform length %d (prime factors up to %d)", N, DDL_RT_MAXR);
        return -1;
    }
    return run_tile_0(mode, dir, phys, p, nthreads, s);
}

// threads one pencil keeps busy: butterflies of the widest stage
static int pencil_threads(int N) {
    if (N >= 8 && !(N & (N - 1))) return N / 8;
    RtFac f;
    if (!rt_factor(N, f)) return 1;
    int rmin = f.radix[0];
    for (int i = 1; i < f.S; ++i) if (f.radix[i] < rmin) rmin = f.radix[i];
    const int t = N / (rmin > 0 ? rmin : 1);
    return t < 1 ? 1 : t;
}

#define DDL_DECLX(N) int run_xfused_##N(int, const XFusedParams&, int, int, ddl_stream_t);
DDL_DECLX(8) DDL_DECLX(16) DDL_DECLX(32) DDL_DECLX(64) DDL_DECLX(128) DDL_DECLX(256) DDL_DECLX(512) DDL_DECLX(1024)
DDL_DECLX(2048)
static int g_xfused_variant = 0;
static int g_plane_chunk = 0;     // ddl_set_option("rhs_plane_chunk", n): see rhs_impl
static int run_xfused(int N, int phys, const XFusedParams& p, int n_outer, ddl_stream_t s) {
    switch (N) {
#define DDL_CASEX(N) case N: return run_xfused_##N(phys, p, n_outer, g_xfused_variant, s);
        DDL_CASEX(8) DDL_CASEX(16) DDL_CASEX(32) DDL_CASEX(64) DDL_CASEX(128) DDL_CASEX(256) DDL_CASEX(512)
        DDL_CASEX(1024) DDL_CASEX(2048)
    }
    return 1;
}

// two-stage strided pass (fast_two.cuh), both builds; ddl_set_option("strided_two", 1 | 2) routes the plain-row passes of the
// lengths it covers (256, 512) through it (1: stage-0 twiddles generated in registers, 2: loaded from the table), 0 = the
// three-stage strided_fast.  Measured at 512^3 MHD on a B200 (profiles/r2/devcheck_b200_strided_two.txt), ms per launch 0 / 1 / 2:
// z_inv 1.04 / 0.93 / 1.19, y_inv 1.50 / 1.28 / 1.63, y_fwd 2.13 / 1.93 / 2.49, z_fwd 1.31 / 1.24 / 1.57: 1 is the default.
#define DDL_DECLT(N) int run_two_strided_##N(int, int, const TwoParams&, int, int, const char*, ddl_stream_t);
DDL_DECLT(8) DDL_DECLT(16) DDL_DECLT(32) DDL_DECLT(64) DDL_DECLT(128) DDL_DECLT(256) DDL_DECLT(512) DDL_DECLT(1024)
DDL_DECLT(2048)
static int g_strided_two = 1;
static int run_two_strided(int N, int dir, const TwoParams& p, int nf, int n_outer, const char* name, ddl_stream_t s) {
    switch (N) {
#define DDL_CASET(N) case N: return run_two_strided_##N(dir, g_strided_two, p, nf, n_outer, name, s);
        DDL_CASET(8) DDL_CASET(16) DDL_CASET(32) DDL_CASET(64) DDL_CASET(128) DDL_CASET(256) DDL_CASET(512)
        DDL_CASET(1024) DDL_CASET(2048)
    }
    return 1;
}

#if DDL_DEVICE_BUILD
#define DDL_DECLF(N) int run_fast_strided_##N(int, const FastParams&, int, int, const char*, ddl_stream_t);
DDL_DECLF(8) DDL_DECLF(16) DDL_DECLF(32) DDL_DECLF(64) DDL_DECLF(128) DDL_DECLF(256) DDL_DECLF(512) DDL_DECLF(1024)
DDL_DECLF(2048)
static int run_fast_strided(int N, int dir, const FastParams& p, int nf, int n_outer, const char* name, ddl_stream_t s) {
    switch (N) {
#define DDL_CASEF(N) case N: return run_fast_strided_##N(dir, p, nf, n_outer, name, s);
        DDL_CASEF(8) DDL_CASEF(16) DDL_CASEF(32) DDL_CASEF(64) DDL_CASEF(128) DDL_CASEF(256) DDL_CASEF(512)
        DDL_CASEF(1024) DDL_CASEF(2048)
    }
    return 1;
}
#endif

// generic kernels only (set by ddl_set_option("fast_kernels", 0); used by the tests to compare)
static int g_use_fast = 1;

template <class T>
static T* upload_vec(const std::vector<T>& v) {
    T* d = (T*)dev_alloc(v.size() * sizeof(T));
    if (d && !v.empty()) dev_upload(d, v.data(), v.size() * sizeof(T));
    return d;
}

}  // namespace ddl

using namespace ddl;

// one periodic axis: wavenumbers, mask, compaction tables, twiddles
struct Axis {
    int n = 1;          // x-space length
    int nk = 1;         // stored spectral length (n, or n/2+1 for the half-complex axis)
    int m = 0;          // retained modes: |index| <= m
    int cnt = 1;        // retained count: 2m+1 (full axis) or m+1 (half axis)
    bool half = false;
    bool full = false;   // every row kept, Nyquist included (shearing box, ky axis only)
    double* kv = nullptr;           // [nk] wavenumber per stored index
    unsigned char* keep = nullptr;  // [nk]
    double* kvc = nullptr;          // [cnt] wavenumber per compact index
    int* c2f = nullptr;             // [cnt] compact -> stored index
    int* f2c = nullptr;             // [nk] stored -> compact or -1
    int* f2f = nullptr;             // [nk] stored -> stored or -1
    cplx* tw = nullptr;             // [n] exp(-2 pi i m / n)
    std::vector<int> h_c2f;         // host copy of c2f
};

// The rows of the ky axis this rank owns (k-space slab, representations.py:231-233); with one
// rank it is the whole axis.
struct YSlab {
    int nyl = 1, ky0 = 0;           // stored rows [ky0, ky0 + nyl)
    int cyl = 1, cy0 = 0;           // retained rows: compact indices [cy0, cy0 + cyl)
    const double* kv = nullptr;     // [nyl]
    const unsigned char* keep = nullptr;
    const double* kvc = nullptr;    // [cyl]
    int* c2f = nullptr;             // [cyl] local compact -> local stored row
};

struct ddl_plan {
    int ndim = 0;
    Axis ax, ay, az;
    // slab decomposition (3-D): x-space split along z, k-space along ky (FFTW-MPI transposed
    // layout, dedalus/utils/fftw/_fftw.pyx:114-148, representations.py:180-186)
    int nranks = 1, rank = 0;
    int nzl = 1, z0 = 0;
    YSlab yl;
    int layout = 0;                 // ky ownership: 0 block (reference), 1 cyclic
    std::vector<int> cyl_of;        // retained ky rows per rank
    unsigned char* y_owner = nullptr;   // [cy] owning rank of every x-side row position
    int* ypos = nullptr;            // [ny] stored ky row -> x-side row position (owner-major), -1 = pruned
    long long ntot = 1;
    KGeom geom;          // local k-array geometry
    long long nmodes = 0;
    int shear_on = 0;               // shearing box (ddl_set_shear): phase factors in the x passes of ddl_forward / ddl_backward
    double sh_S = 0.0, sh_t = 0.0, sh_dy = 0.0;
    double* cfl_out = nullptr;      // CFL capture target of the x passes (ddl_rhs_capture_max), caller-owned
    double* red_partial = nullptr;  // block partials of the reductions (reduce.cuh)
    std::vector<void*> owned;
};

static int build_axis(ddl_plan* pl, Axis& a, int n, bool half, const double* kv, const uint8_t* keep, bool allow_full = false) {
    a.n = n; a.half = half; a.nk = half ? n / 2 + 1 : n;
    {
        RtFac f;
        if (n < 2 || n > 2048 || !rt_factor(n, f)) {
            set_error("axis length %d unsupported (2..2048 with prime factors up to %d)", n, DDL_RT_MAXR);
            return -1;
        }
    }
    int m = -1;
    for (int j = 0; j < a.nk; ++j) {
        int mi = (j <= n / 2) ? j : n - j;
        if (keep[j] && mi > m) m = mi;
    }
    if (m < 0) { set_error("dealias mask removes every mode"); return -1; }
    // a full axis keeps every row, the Nyquist one included: no pruning along it (the shearing box, whose ky mask
    // depends on kx and time and is applied by the caller, representations.py:627-642); only the ky axis may be full
    const bool full = !half && 2 * m >= n && n % 2 == 0;
    if (full && !allow_full) { set_error("dealias mask keeps the Nyquist mode; unsupported on this axis"); return -1; }
    if (half && 2 * m >= n) { set_error("dealias mask keeps the Nyquist mode; unsupported"); return -1; }      // odd n has none
    for (int j = 0; j < a.nk; ++j) {
        int mi = (j <= n / 2) ? j : n - j;
        if ((keep[j] != 0) != (mi <= m)) { set_error("dealias mask is not of the form |k index| <= m"); return -1; }
    }
    a.m = m; a.cnt = half ? m + 1 : (full ? n : 2 * m + 1); a.full = full;
    std::vector<double> kvh(kv, kv + a.nk), kvc(a.cnt);
    std::vector<unsigned char> kp(keep, keep + a.nk);
    std::vector<int> c2f(a.cnt), f2c(a.nk, -1), f2f(a.nk, -1);
    for (int j = 0; j < a.cnt; ++j) {
        int f = (j <= m) ? j : j + (a.nk - a.cnt);
        c2f[j] = f; f2c[f] = j; f2f[f] = f; kvc[j] = kv[f];
    }
    a.h_c2f = c2f;
    std::vector<cplx> tw(n);
    const long double PI = acosl(-1.0L);
    for (int q = 0; q < n; ++q) { tw[q].x = (double)cosl(-2 * PI * q / n); tw[q].y = (double)sinl(-2 * PI * q / n); }
    a.kv = upload_vec(kvh); a.keep = upload_vec(kp); a.kvc = upload_vec(kvc);
    a.c2f = upload_vec(c2f); a.f2c = upload_vec(f2c); a.f2f = upload_vec(f2f); a.tw = upload_vec(tw);
    void* all[] = {a.kv, a.keep, a.kvc, a.c2f, a.f2c, a.f2f, a.tw};
    for (void* p : all) { if (!p) { set_error("device allocation failed"); return -2; } pl->owned.push_back(p); }
    return 0;
}

extern "C" int ddl_plan_create_slab(ddl_plan** out, int ndim, const int64_t* shape_x, const double* kx, const double* ky,
                                    const double* kz, const uint8_t* keepx, const uint8_t* keepy, const uint8_t* keepz,
                                    int nranks, int rank, int layout) {
    const double* kv_host_y = ky;
    if (layout != 0 && layout != 1) { set_error("ky layout must be 0 (block) or 1 (cyclic)"); return -1; }
    if (!out || (ndim != 2 && ndim != 3)) { set_error("Must use either 2 or 3 dimensions."); return -1; }
    if (nranks < 1 || rank < 0 || rank >= nranks) { set_error("bad rank %d of %d", rank, nranks); return -1; }
    if (nranks > 1 && ndim != 3) { set_error("slab decomposition is 3-D only (2-D grids run as replicas)"); return -1; }
    ddl_plan* pl = new ddl_plan();
    pl->ndim = ndim;
    int rc = 0;
    if (ndim == 3) {
        rc = build_axis(pl, pl->az, (int)shape_x[0], false, kz, keepz);
        if (!rc) rc = build_axis(pl, pl->ay, (int)shape_x[1], false, ky, keepy, true);
        if (!rc) rc = build_axis(pl, pl->ax, (int)shape_x[2], true, kx, keepx);
    } else {
        rc = build_axis(pl, pl->ay, (int)shape_x[0], false, ky, keepy, true);
        if (!rc) rc = build_axis(pl, pl->ax, (int)shape_x[1], true, kx, keepx);
    }
    if (rc) { ddl_plan_destroy(pl); return rc; }
    const Axis& Y = pl->ay;
    if (nranks > 1 && ((nranks & (nranks - 1)) || pl->az.n % nranks || Y.n % nranks)) {
        set_error("slab decomposition needs a power-of-two rank count dividing nz=%d and ny=%d (got %d)", pl->az.n, Y.n, nranks);
        ddl_plan_destroy(pl);
        return -1;
    }
    pl->nranks = nranks; pl->rank = rank;
    pl->nzl = (ndim == 3 ? pl->az.n : 1) / nranks; pl->z0 = rank * pl->nzl;
    // Ownership of the stored ky rows.  layout 0 ("block", the reference's): rank r owns rows
    // [r*nyl, (r+1)*nyl).  layout 1 ("cyclic"): rank r owns rows r, r+P, r+2P, ... - under 2/3
    // dealiasing the retained rows sit at both ends of the FFT-ordered axis, so block slabs leave
    // the middle ranks without work (8 ranks, ny = 512: 64,64,43,0,0,42,64,64 retained rows)
    // while cyclic ownership gives every rank 42 or 43.
    YSlab& L = pl->yl;
    pl->layout = layout;
    L.nyl = Y.nk / nranks; L.ky0 = layout == 0 ? rank * L.nyl : rank;
    auto owner_of = [&](int i) { return layout == 0 ? i / L.nyl : i % nranks; };
    auto local_of = [&](int i) { return layout == 0 ? i % L.nyl : i / nranks; };
    pl->cyl_of.assign(nranks, 0);
    for (int j = 0; j < Y.cnt; ++j) pl->cyl_of[owner_of(Y.h_c2f[j])]++;
    std::vector<int> cy0_of(nranks, 0);
    for (int r = 1; r < nranks; ++r) cy0_of[r] = cy0_of[r - 1] + pl->cyl_of[r - 1];
    // x-side row order: owner-major (rank 0's retained rows, then rank 1's, ...), each rank's rows
    // in stored order; block layout: identical to the compact order
    std::vector<int> ypos(Y.nk, -1), lc2f, fill(nranks, 0);
    std::vector<unsigned char> owner(Y.cnt > 0 ? Y.cnt : 1);
    std::vector<double> lkv(L.nyl), lkvc;
    std::vector<unsigned char> lkeep(L.nyl);
    for (int i = 0; i < Y.nk; ++i) {
        const int r = owner_of(i);
        const bool kept = i <= Y.m || i >= Y.nk - Y.m;
        if (r == rank) { lkv[local_of(i)] = kv_host_y[i]; lkeep[local_of(i)] = kept ? 1 : 0; }
        if (!kept) continue;
        const int pos = cy0_of[r] + fill[r]++;
        ypos[i] = pos; owner[pos] = (unsigned char)r;
        if (r == rank) { lc2f.push_back(local_of(i)); lkvc.push_back(kv_host_y[i]); }
    }
    L.cyl = (int)lc2f.size();
    L.cy0 = cy0_of[rank];
    pl->y_owner = upload_vec(owner);
    pl->ypos = upload_vec(ypos);
    if (lkvc.empty()) lkvc.push_back(0.0);
    if (lc2f.empty()) lc2f.push_back(0);
    double* d_kv = upload_vec(lkv); unsigned char* d_keep = upload_vec(lkeep); double* d_kvc = upload_vec(lkvc);
    L.c2f = upload_vec(lc2f);
    void* mine[] = {pl->y_owner, pl->ypos, d_kv, d_keep, d_kvc, L.c2f};
    for (void* q : mine) {
        if (!q) { set_error("device allocation failed"); ddl_plan_destroy(pl); return -2; }
        pl->owned.push_back(q);
    }
    L.kv = d_kv; L.keep = d_keep; L.kvc = d_kvc;
    KGeom& g = pl->geom;
    if (ndim == 3) {
        pl->ntot = (long long)pl->ax.n * pl->ay.n * pl->az.n;
        g.dim[0] = L.nyl; g.kv[0] = L.kv; g.keep[0] = L.keep; g.ax[0] = 1;
        g.dim[1] = pl->az.nk; g.kv[1] = pl->az.kv; g.keep[1] = pl->az.keep; g.ax[1] = 2;
        g.dim[2] = pl->ax.nk; g.kv[2] = pl->ax.kv; g.keep[2] = pl->ax.keep; g.ax[2] = 0;
        g.twod = 0;
    } else {
        pl->ntot = (long long)pl->ax.n * pl->ay.n;
        g.dim[0] = 1; g.kv[0] = nullptr; g.keep[0] = nullptr; g.ax[0] = -1;
        g.dim[1] = pl->ax.nk; g.kv[1] = pl->ax.kv; g.keep[1] = pl->ax.keep; g.ax[1] = 0;
        g.dim[2] = pl->ay.nk; g.kv[2] = pl->ay.kv; g.keep[2] = pl->ay.keep; g.ax[2] = 1;
        g.twod = 1;
    }
    pl->nmodes = (long long)g.dim[0] * g.dim[1] * g.dim[2];
    pl->red_partial = (double*)dev_alloc((size_t)DDL_RED_MAXBLOCKS * DDL_RED_MAXV * sizeof(double));
    if (!pl->red_partial) { set_error("device allocation failed"); ddl_plan_destroy(pl); return -2; }
    pl->owned.push_back(pl->red_partial);
    *out = pl;
    return 0;
}

extern "C" int ddl_plan_create(ddl_plan** out, int ndim, const int64_t* shape_x, const double* kx, const double* ky,
                               const double* kz, const uint8_t* keepx, const uint8_t* keepy, const uint8_t* keepz) {
    return ddl_plan_create_slab(out, ndim, shape_x, kx, ky, kz, keepx, keepy, keepz, 1, 0, 0);
}

extern "C" int ddl_plan_destroy(ddl_plan* pl) {
    if (!pl) return 0;
    for (void* p : pl->owned) dev_free(p);
    delete pl;
    return 0;
}

// ---------------------------------------------------------------- workspace layout
// pitch of the retained-kx axis in the 3-D workspace arrays: a multiple of 8 complex (128 B)
// so that every CX-wide row segment of the strided passes is a whole number of cache lines
static long long kx_pitch(const ddl_plan* pl) { return pl->ndim == 3 ? (pl->ax.cnt + 7) / 8 * 8 : pl->ax.cnt; }

// Per-field sizes (complex elements) of the pipeline arrays of THIS rank, 3-D:
//   ks : k-side pencils  [peer][cyl][nzl][CX]   z-pass output (inverse) / input (forward); peer-blocked rows
//   xs : x-side pencils  [cy][nzl][CX]          y-pass input (inverse) / output (forward)
//   b  : half-transformed lines [nzl][y][CX]    y-pass output / x-pass input and output
//   e  : retained product spectra [cyl][cz][CX]
struct SlabSizes { long long ks, xs, b, e; };
static SlabSizes slab_sizes(const ddl_plan* pl) {
    SlabSizes s;
    const long long CX = kx_pitch(pl);
    s.ks = (long long)pl->yl.cyl * pl->az.n * CX;
    s.xs = (long long)pl->ay.cnt * pl->nzl * CX;
    s.b = (long long)pl->nzl * pl->ay.n * CX;
    s.e = (long long)pl->yl.cyl * pl->az.cnt * CX;
    return s;
}

struct WsLayout {
    long long r0, r1, r2;   // region sizes in cplx elements
};
static WsLayout ws_layout(const ddl_plan* pl, int ni, int no) {
    WsLayout w;
    const long long CX = kx_pitch(pl), CY = pl->ay.cnt;
    if (pl->ndim == 3) {
        const SlabSizes s = slab_sizes(pl);
        const int nmax = ni > no ? ni : no;
        const long long a = s.ks > s.xs ? s.ks : s.xs;            // one rank: ks == xs
        w.r0 = nmax * a;                                          // k-side / x-side pencils (inverse), again (forward)
        long long b = (long long)ni * s.b, e = (long long)no * s.e;
        w.r1 = b > e ? b : e;                                     // B, later E
        w.r2 = (long long)no * s.b;                               // C
    } else {
        const long long ny = pl->ay.n;
        w.r0 = (long long)ni * CX * ny;                           // A [f][kx_c][y]
        w.r1 = (long long)no * CX * ny;                           // C [f][kx_c][y]
        w.r2 = (long long)no * CX * CY;                           // E [f][kx_c][ky_c]
    }
    return w;
}

extern "C" size_t ddl_workspace_bytes(const ddl_plan* pl, int n_in, int n_out) {
    WsLayout w = ws_layout(pl, n_in, n_out);
    return (size_t)(w.r0 + w.r1 + w.r2) * sizeof(cplx);
}

// ni / no: inverse / forward transforms, code: kernel policy (tile_inst.cu), nc: state components.
// The *_ADV ids are the advective-form policies for non-solenoidal states: ni = nc + (1 or 2) divergence
// spectra, which live in caller-provided scratch arrays after the state in the pointer list.
static void phys_counts(int ndim, int physics, int& ni, int& no, int& code, int* nc_out = nullptr) {
    const bool adv = physics >= DDL_HYDRO_ADV;
    const int base = adv ? physics - DDL_HYDRO_ADV : physics;
    int nc;
    if (ndim == 3) {
        if (base == DDL_HYDRO) { nc = 3; no = 6; code = 3; }
        else if (base == DDL_BOUSSINESQ) { nc = 4; no = 9; code = 4; }
        else { nc = 6; no = 9; code = 5; }
    } else {
        if (base == DDL_HYDRO) { nc = 2; no = 3; code = 0; }
        else if (base == DDL_BOUSSINESQ) { nc = 3; no = 5; code = 1; }
        else { nc = 4; no = 4; code = 2; }
    }
    ni = nc;
    if (adv) {
        ni += (base == DDL_MHD) ? 2 : 1;
        no += ndim + (base == DDL_BOUSSINESQ ? 1 : 0);
        code = (ndim == 3 ? 9 : 6) + base;
    }
    if (nc_out) *nc_out = nc;
}

// theta = i k.u (and i k.B) into the scratch arrays that follow the state (advective-form policies only)
static int phase_theta(ddl_plan* pl, int physics, void* const* state, ddl_stream_t st) {
    if (physics < DDL_HYDRO_ADV) return 0;
    const int base = physics - DDL_HYDRO_ADV, nd = pl->ndim;
    int ni, no, code, nc;
    phys_counts(nd, physics, ni, no, code, &nc);
    ThetaF f;
    memset(&f, 0, sizeof(f));
    f.g = pl->geom; f.nd = nd;
    for (int c = 0; c < nd; ++c) f.U[c] = (const cplx*)state[c];
    f.thu = (cplx*)state[nc];
    if (base == DDL_MHD) {
        for (int c = 0; c < nd; ++c) f.B[c] = (const cplx*)state[nd + c];
        f.thb = (cplx*)state[nc + 1];
    }
    return launch_items(f, pl->nmodes, st, "theta");
}

extern "C" size_t ddl_rhs_workspace_bytes(const ddl_plan* pl, int physics) {
    int ni, no, code;
    phys_counts(pl->ndim, physics, ni, no, code);
    return ddl_workspace_bytes(pl, ni, no);
}

// ---------------------------------------------------------------- pass builders
static int pick_c2c_group(int N, long long inner_len) {
    int g = 4096 / N;
    if (g < 1) g = 1;
    if (g > 32) g = 32;
    if (g > inner_len) g = (int)inner_len;
    return g;
}
// CTA size of the generic tile kernel (72 registers per thread: at most 896 threads fit an SM)
static int round32(int t) { t = (t + 31) / 32 * 32; return t < 64 ? 64 : (t > DDL_TILE_MAX_THREADS ? DDL_TILE_MAX_THREADS : t); }

// complex pass of nf fields along an axis of length N
// retained rows of a pruned axis (m < 0: all rows present); compact: 0 stored in place, 1 stored
// contiguously in FFT order, 2 stored at the positions of the side's n_tab (owner-major order)
struct RowSpec { int m; int compact; };
static const RowSpec ALL_ROWS = {-1, 0};

// peer mode (slab exchange fused into the pass): the output blocks are given by a device table
// of nf x nblk pointers into the peers' arenas
struct PeerOut { void* const* tab; const unsigned char* own; int nblk; int mask_all; long long off; };

static int pass_c2c(const char* name, int N, int dir, int nf, const void* const* in, void* const* out, const TileSide& si,
                    const TileSide& so, RowSpec ri, RowSpec ro, int inner_len, int n_outer, double scale, const cplx* tw,
                    ddl_stream_t st, const PeerOut* peer = nullptr) {
    if (n_outer <= 0 || nf <= 0 || inner_len <= 0) return 0;     // a rank may own no retained ky row
#if DDL_DEVICE_BUILD
    const bool staged_opt_in = ddl::g_strided_staged != 0;        // an explicit opt-in to the staged three-stage kernel wins
#else
    const bool staged_opt_in = false;
#endif
    if (g_strided_two && !staged_opt_in && g_use_fast && !peer && si.s_inner == 1 && so.s_inner == 1 && si.s_n != 1 && so.s_n != 1 && nf <= DDL_MAXF &&
        !si.split && !so.split && ri.compact != 2 && ro.compact != 2) {
        TwoParams t;
        memset(&t, 0, sizeof(t));
        bool distinct = true;       // the kernel reads through the non-coherent path: never for a pass that writes its own input
        for (int i = 0; i < nf; ++i) { t.in[i] = (const cplx*)in[i]; t.out[i] = (cplx*)out[i]; distinct = distinct && in[i] != (const void*)out[i]; }
        t.si.s_n = si.s_n; t.si.s_outer = si.s_outer; t.si.outer_tab = si.outer_tab; t.si.m = ri.m; t.si.compact = ri.compact;
        t.so.s_n = so.s_n; t.so.s_outer = so.s_outer; t.so.outer_tab = so.outer_tab; t.so.m = ro.m; t.so.compact = ro.compact;
        t.inner_len = inner_len; t.scale = scale; t.tw = tw;
        const int rc = distinct ? run_two_strided(N, dir, t, nf, n_outer, name, st) : 1;
        if (rc <= 0) return rc;
    }
#if DDL_DEVICE_BUILD
    const bool pow2 = !(si.split & (si.split - 1)) && !(so.split & (so.split - 1));
    if (g_use_fast && si.s_inner == 1 && so.s_inner == 1 && si.s_n != 1 && so.s_n != 1 && nf <= DDL_MAXF && pow2) {
        FastParams f;
        memset(&f, 0, sizeof(f));
        for (int i = 0; i < nf; ++i) { f.in[i] = (const cplx*)in[i]; f.out[i] = (cplx*)out[i]; }
        auto conv = [](FastSide& d, const TileSide& s, RowSpec r) {
            d.s_n = s.s_n; d.s_outer = s.s_outer; d.outer_tab = s.outer_tab; d.m = r.m; d.compact = r.compact;
            d.row_tab = (r.compact == 2) ? s.n_tab : nullptr;
            d.split_shift = 31; d.split_mask = 0x7fffffff; d.s_blk = 0;
            if (s.split) { int sh = 0; while ((1 << sh) < s.split) ++sh; d.split_shift = sh; d.split_mask = s.split - 1; d.s_blk = s.s_blk; }
        };
        conv(f.si, si, ri); conv(f.so, so, ro);
        if (peer) {
            f.so.peer_tab = (cplx* const*)peer->tab; f.so.own_tab = peer->own; f.so.nblk = peer->nblk; f.so.peer_off = peer->off;
            if (peer->mask_all) f.so.split_mask = 0x7fffffff;
        }
        f.inner_len = inner_len; f.scale = scale; f.tw = tw;
        int rc = run_fast_strided(N, dir, f, nf, n_outer, name, st);
        if (rc <= 0) return rc;
    }
    if (peer) { set_error("peer-store passes need the specialised strided kernel (axis length %d)", N); return -1; }
#else
    (void)ri; (void)ro;
    if (peer) { set_error("peer-store passes need the CUDA build"); return -1; }
#endif
    TileParams p;
    memset(&p, 0, sizeof(p));
    for (int f = 0; f < nf; ++f) { p.in[f] = in[f]; p.out[f] = out[f]; }
    p.si = si; p.so = so; p.nf_in = p.nf_out = nf; p.nft = 1;
    p.G = pick_c2c_group(N, inner_len);
    p.inner_len = inner_len; p.n_outer = n_outer; p.kn = 0;
    p.ld = (si.s_n == 1 || so.s_n == 1) ? (p.G | 1) : p.G;
    p.scale = scale; p.tw = tw; p.name = name;
    return run_tile(N, TM_C2C, dir, 0, p, round32(p.G * pencil_threads(N)), st);
}

// pair-mode pass (C2R / R2C / FUSED) over real lines
static int pass_pair(const char* name, int N, int mode, int phys, int ni, int no, const void* const* in, void* const* out,
                     const TileSide& si, const TileSide& so, int n_lines, int n_outer, int kn, double scale,
                     const cplx* tw, const PhysConst& pc, ddl_stream_t st, double* cfl = nullptr, const ddl_plan* shear = nullptr) {
    if (n_outer <= 0) return 0;
    TileParams p;
    memset(&p, 0, sizeof(p));
    if (shear && shear->shear_on) {
        p.sh_on = 1; p.sh_S = shear->sh_S; p.sh_t = shear->sh_t; p.sh_dy = shear->sh_dy; p.sh_kx = shear->ax.kv;
    }
    for (int f = 0; f < ni; ++f) p.in[f] = in[f];
    for (int f = 0; f < no; ++f) p.out[f] = out[f];
    p.si = si; p.so = so; p.nf_in = ni; p.nf_out = no;
    p.nft = ni > no ? ni : no;
    int g = 6144 / (N * p.nft);
    if (g < 1) g = 1;
    if (g > 16) g = 16;
    const int pairs = (n_lines + 1) / 2;
    if (g > pairs) g = pairs;
    p.G = g;
    p.inner_len = n_lines; p.n_outer = n_outer; p.kn = kn;
    p.ld = (g * p.nft) | 1;
    p.scale = scale; p.tw = tw; p.pc = pc; p.name = name; p.cfl = cfl;
    return run_tile(N, mode, 0, phys, p, round32(g * p.nft * pencil_threads(N)), st);
}

static TileSide side(long long s_n, long long s_inner, long long s_outer, const int* n_tab, const int* outer_tab,
                     int split = 0, long long s_blk = 0) {
    TileSide s; s.s_n = s_n; s.s_inner = s_inner; s.s_outer = s_outer; s.n_tab = n_tab; s.outer_tab = outer_tab;
    s.split = split; s.s_blk = s_blk;
    return s;
}

#define DDL_TRY(x) do { int _rc = (x); if (_rc) return _rc; } while (0)

static PhysConst phys_const(const ddl_phys_params* prm) {
    PhysConst pc;
    pc.inv_fpr = 1.0 / (4.0 * 3.14159265358979323846 * prm->rho0);
    pc.g_alpha = prm->g * prm->alpha_t;
    pc.beta = prm->beta;
    pc.bdir = prm->boussinesq_dir;
    return pc;
}

// ---------------------------------------------------------------- 3-D pipeline phases
// One RHS / transform is  zinv -> [exchange] -> yinv -> x pass -> yfwd -> [exchange] -> zfwd.
// With one rank the exchanges are the identity (the k-side and x-side pencil layouts coincide);
// with P ranks the host layer moves peer block s of every k-side array to rank s (and back) with
// an all-to-all between the calls (dedalus/data_objects/slab.py).

// k-side layout of this rank: stored row z of compact local ky row j at
//   (z / nzl) * (cyl*nzl*CX) + j * (nzl*CX) + (z % nzl) * CX
static TileSide kside(const ddl_plan* pl) {
    const long long CX = kx_pitch(pl);
    if (pl->nranks == 1) return side(CX, 1, (long long)pl->az.n * CX, nullptr, nullptr);
    return side(CX, 1, (long long)pl->nzl * CX, nullptr, nullptr, pl->nzl, (long long)pl->yl.cyl * pl->nzl * CX);
}

// z pass, inverse: k[kyl][kz][kx] (retained modes) -> k-side pencils
static int phase_zinv(ddl_plan* pl, int nf, const void* const* kin, void* const* S, ddl_stream_t st) {
    const Axis &X = pl->ax, &Z = pl->az;
    const long long KP = X.nk;
    return pass_c2c("z_inv", Z.n, +1, nf, kin, S, side(KP, 1, (long long)Z.n * KP, Z.f2f, pl->yl.c2f), kside(pl),
                    RowSpec{Z.m, 0}, ALL_ROWS, X.cnt, pl->yl.cyl, 1.0, Z.tw, st);
}
// z pass, inverse, with the exchange fused in: row z of local compact ky row j goes straight to
//   tab[f][z / nzl] + j*(nzl*CX) + (z % nzl)*CX     (tab[f][s] = rank s's x-side field f + cy0_me*nzl*CX)
static int phase_zinv_peer(ddl_plan* pl, int nf, const void* const* kin, void* const* tab, ddl_stream_t st) {
    const Axis &X = pl->ax, &Z = pl->az;
    const long long KP = X.nk, CX = kx_pitch(pl);
    PeerOut po = {tab, nullptr, pl->nranks, 0, 0};
    std::vector<void*> dummy(nf, nullptr);
    TileSide so = side(CX, 1, (long long)pl->nzl * CX, nullptr, nullptr, pl->nzl, 0);
    return pass_c2c("z_inv", Z.n, +1, nf, kin, dummy.data(), side(KP, 1, (long long)Z.n * KP, Z.f2f, pl->yl.c2f), so,
                    RowSpec{Z.m, 0}, ALL_ROWS, X.cnt, pl->yl.cyl, 1.0, Z.tw, st, &po);
}
// y pass, forward, with the exchange fused in: compact ky row j of local plane zl goes to
//   tab[f][owner(j)] + j*(nzl*CX) + zl*CX     (tab[f][s] = rank s's k-side field f + (me*cyl_s - cy0_s)*nzl*CX)
// planes [z0, z0 + nzc) of the local slab only (chunked so that it overlaps the x pass of the next chunk)
static int phase_yfwd_peer(ddl_plan* pl, int nf, const void* const* Cin, void* const* tab, int z0, int nzc, ddl_stream_t st) {
    const Axis &X = pl->ax, &Y = pl->ay;
    const long long CX = kx_pitch(pl), nzl = pl->nzl;
    PeerOut po = {tab, pl->y_owner, pl->nranks, 1, (long long)z0 * CX};
    std::vector<void*> dummy(nf, nullptr);
    std::vector<const void*> cin(nf);
    for (int f = 0; f < nf; ++f) cin[f] = (const cplx*)Cin[f] + (long long)z0 * Y.n * CX;
    return pass_c2c("y_fwd", Y.n, -1, nf, cin.data(), dummy.data(), side(CX, 1, (long long)Y.n * CX, nullptr, nullptr),
                    side(nzl * CX, 1, CX, pl->ypos, nullptr), ALL_ROWS, RowSpec{Y.m, pl->layout ? 2 : 1}, X.cnt, nzc, 1.0, Y.tw, st, &po);
}
// y pass, inverse: x-side pencils A[cy][nzl][CX] -> B[nzl][y][CX]
static int phase_yinv(ddl_plan* pl, int nf, const void* const* A, void* const* B, ddl_stream_t st) {
    const Axis &X = pl->ax, &Y = pl->ay;
    const long long CX = kx_pitch(pl), nzl = pl->nzl;
    return pass_c2c("y_inv", Y.n, +1, nf, A, B, side(nzl * CX, 1, CX, pl->ypos, nullptr), side(CX, 1, (long long)Y.n * CX, nullptr, nullptr),
                    RowSpec{Y.m, pl->layout ? 2 : 1}, ALL_ROWS, X.cnt, (int)nzl, 1.0, Y.tw, st);
}
// the same two passes for local planes [z0, z0 + nzc) with CHUNK-LOCAL b / c arrays Bc, Cc[nzc][y][CX] (rhs_plane_chunk)
static int phase_yinv_planes(ddl_plan* pl, int nf, const void* const* A, void* const* Bc, int z0, int nzc, ddl_stream_t st) {
    const Axis &X = pl->ax, &Y = pl->ay;
    const long long CX = kx_pitch(pl), nzl = pl->nzl;
    std::vector<const void*> a(nf);
    for (int f = 0; f < nf; ++f) a[f] = (const cplx*)A[f] + (long long)z0 * CX;
    return pass_c2c("y_inv", Y.n, +1, nf, a.data(), Bc, side(nzl * CX, 1, CX, pl->ypos, nullptr), side(CX, 1, (long long)Y.n * CX, nullptr, nullptr),
                    RowSpec{Y.m, pl->layout ? 2 : 1}, ALL_ROWS, X.cnt, nzc, 1.0, Y.tw, st);
}
static int phase_yfwd_planes(ddl_plan* pl, int nf, const void* const* Cc, void* const* D, int z0, int nzc, ddl_stream_t st) {
    const Axis &X = pl->ax, &Y = pl->ay;
    const long long CX = kx_pitch(pl), nzl = pl->nzl;
    std::vector<void*> d(nf);
    for (int f = 0; f < nf; ++f) d[f] = (cplx*)D[f] + (long long)z0 * CX;
    return pass_c2c("y_fwd", Y.n, -1, nf, Cc, d.data(), side(CX, 1, (long long)Y.n * CX, nullptr, nullptr), side(nzl * CX, 1, CX, pl->ypos, nullptr),
                    ALL_ROWS, RowSpec{Y.m, pl->layout ? 2 : 1}, X.cnt, nzc, 1.0, Y.tw, st);
}
// y pass, forward: C[nzl][y][CX] -> x-side pencils D[cy][nzl][CX]
static int phase_yfwd(ddl_plan* pl, int nf, const void* const* Cin, void* const* D, ddl_stream_t st) {
    const Axis &X = pl->ax, &Y = pl->ay;
    const long long CX = kx_pitch(pl), nzl = pl->nzl;
    return pass_c2c("y_fwd", Y.n, -1, nf, Cin, D, side(CX, 1, (long long)Y.n * CX, nullptr, nullptr), side(nzl * CX, 1, CX, pl->ypos, nullptr),
                    ALL_ROWS, RowSpec{Y.m, pl->layout ? 2 : 1}, X.cnt, (int)nzl, 1.0, Y.tw, st);
}
// z pass, forward: k-side pencils -> E[cyl][kz_c][CX] (compact products) or k[kyl][kz][kx] (full_out)
static int phase_zfwd(ddl_plan* pl, int nf, const void* const* R, void* const* dst, bool full_out, ddl_stream_t st) {
    const Axis &X = pl->ax, &Z = pl->az;
    const long long KP = X.nk, CX = kx_pitch(pl);
    TileSide so = full_out ? side(KP, 1, (long long)Z.n * KP, Z.f2f, pl->yl.c2f) : side(CX, 1, (long long)Z.cnt * CX, Z.f2c, nullptr);
    return pass_c2c("z_fwd", Z.n, -1, nf, R, dst, kside(pl), so, ALL_ROWS, RowSpec{Z.m, full_out ? 0 : 1}, X.cnt, pl->yl.cyl,
                    1.0, Z.tw, st);
}
// x pass with the real-space products: B[f][nzl][y][CX] -> C[f][nzl][y][CX]
static int phase_xfused(ddl_plan* pl, int code, int ni, int no, const void* const* Bin, void* const* Cin, const PhysConst& pc,
                        ddl_stream_t st, int z0 = 0, int nzc = -1) {
    const Axis &X = pl->ax, &Y = pl->ay;
    const long long KXP = kx_pitch(pl);
    const double sc = 1.0 / (double)pl->ntot;
    if (nzc < 0) nzc = pl->nzl;
    if (nzc <= 0) return 0;
    std::vector<const void*> Bv(ni);
    std::vector<void*> Cv(no);
    for (int f = 0; f < ni; ++f) Bv[f] = (const cplx*)Bin[f] + (long long)z0 * Y.n * KXP;
    for (int f = 0; f < no; ++f) Cv[f] = (cplx*)Cin[f] + (long long)z0 * Y.n * KXP;
    const void* const* B = Bv.data();
    void* const* Cout = Cv.data();
    if (g_use_fast && Y.n % 2 == 0 && ni <= DDL_XF_MAXI && no <= DDL_XF_MAXO) {
        XFusedParams xp;
        memset(&xp, 0, sizeof(xp));
        for (int f = 0; f < ni; ++f) xp.in[f] = (const cplx*)B[f];
        for (int f = 0; f < no; ++f) xp.out[f] = (cplx*)Cout[f];
        xp.pitch = KXP; xp.s_outer = (long long)Y.n * KXP; xp.n_lines = Y.n; xp.kn = X.cnt;
        xp.scale = sc; xp.tw = X.tw; xp.pc = pc; xp.cfl = pl->cfl_out;
        const int rcx = run_xfused(X.n, code, xp, nzc, st);
        if (rcx <= 0) return rcx;
    }
    TileSide s = side(1, KXP, (long long)Y.n * KXP, nullptr, nullptr);
    return pass_pair("x_fused", X.n, TM_FUSED, code, ni, no, B, Cout, s, s, Y.n, nzc, X.cnt, sc, X.tw, pc, st, pl->cfl_out);
}
// plain x passes of the transform API: B[nzl][y][CX] -> x[nzl][y][nx] and back (normalised)
static int phase_xc2r(ddl_plan* pl, const void* B, double* x, ddl_stream_t st) {
    const Axis &X = pl->ax, &Y = pl->ay;
    const void* in[1] = {B};
    void* out[1] = {x};
    PhysConst pc = {};
    return pass_pair("x_c2r", X.n, TM_C2R, 0, 1, 1, in, out, side(1, kx_pitch(pl), (long long)Y.n * kx_pitch(pl), nullptr, nullptr),
                     side(1, X.n, (long long)Y.n * X.n, nullptr, nullptr), Y.n, pl->nzl, X.cnt, 1.0, X.tw, pc, st, nullptr, pl);
}
static int phase_xr2c(ddl_plan* pl, const double* x, void* Cout, ddl_stream_t st) {
    const Axis &X = pl->ax, &Y = pl->ay;
    const void* in[1] = {x};
    void* out[1] = {Cout};
    PhysConst pc = {};
    return pass_pair("x_r2c", X.n, TM_R2C, 0, 1, 1, in, out, side(1, X.n, (long long)Y.n * X.n, nullptr, nullptr),
                     side(1, kx_pitch(pl), (long long)Y.n * kx_pitch(pl), nullptr, nullptr), Y.n, pl->nzl, X.cnt,
                     1.0 / (double)pl->ntot, X.tw, pc, st, nullptr, pl);
}

// ---------------------------------------------------------------- 2-D passes
// inverse: k[kx][ky] -> A[kx_c][y];  forward tail: C[kx_c][y] -> E[kx_c][ky_c] or k[kx][ky]
static int inverse_head_2d(ddl_plan* pl, int nf, const void* const* kin, cplx* r0, void** heads, ddl_stream_t st) {
    const Axis &X = pl->ax, &Y = pl->ay;
    const long long CX = kx_pitch(pl), ny = Y.n;
    std::vector<void*> A(nf);
    for (int f = 0; f < nf; ++f) A[f] = r0 + f * CX * ny;
    DDL_TRY(pass_c2c("y_inv", Y.n, +1, nf, kin, A.data(), side(1, ny, 0, Y.f2f, nullptr), side(1, ny, 0, nullptr, nullptr),
                     RowSpec{Y.m, 0}, ALL_ROWS, X.cnt, 1, 1.0, Y.tw, st));
    for (int f = 0; f < nf; ++f) heads[f] = A[f];
    return 0;
}
static int forward_tail_2d(ddl_plan* pl, int nf, void* const* Cin, void* const* dst, bool full_out, ddl_stream_t st) {
    const Axis &X = pl->ax, &Y = pl->ay;
    const long long ny = Y.n, CY = Y.cnt;
    TileSide so = full_out ? side(1, ny, 0, Y.f2f, nullptr) : side(1, CY, 0, Y.f2c, nullptr);
    return pass_c2c("y_fwd", Y.n, -1, nf, Cin, dst, side(1, ny, 0, nullptr, nullptr), so, ALL_ROWS,
                    RowSpec{Y.m, full_out ? 0 : 1}, X.cnt, 1, 1.0, Y.tw, st);
}

static int mask_arrays(ddl_plan* pl, int n, void* const* arr, ddl_stream_t st) {
    MaskF f;
    f.g = pl->geom;
    int done = 0;
    while (done < n) {
        int c = n - done < DDL_MAXF ? n - done : DDL_MAXF;
        for (int i = 0; i < c; ++i) f.arr[i] = (cplx*)arr[done + i];
        f.narr = c;
        DDL_TRY(launch_items(f, pl->nmodes, st, "mask"));
        done += c;
    }
    return 0;
}

static int check_ws(const ddl_plan* pl, int ni, int no, void* work, size_t bytes) {
    if (!work || bytes < ddl_workspace_bytes(pl, ni, no)) {
        set_error("workspace too small: %zu < %zu bytes", bytes, ddl_workspace_bytes(pl, ni, no));
        return -1;
    }
    return 0;
}
static int need_one_rank(const ddl_plan* pl, const char* what) {
    if (pl->nranks != 1) {
        set_error("%s is the one-rank entry point; a slab-decomposed plan (%d ranks) goes through the ddl_slab_* phases", what, pl->nranks);
        return -1;
    }
    return 0;
}

// shearing-box plans (full ky axis) take the generic kernels for every pass
struct FastGuard {
    int saved;
    explicit FastGuard(bool off) : saved(g_use_fast) { if (off) g_use_fast = 0; }
    ~FastGuard() { g_use_fast = saved; }
};

// ---------------------------------------------------------------- transforms
extern "C" int ddl_dealias(ddl_plan* pl, void* k, void* stream) {
    void* arr[1] = {k};
    return mask_arrays(pl, 1, arr, (ddl_stream_t)stream);
}

extern "C" int ddl_dealias_array(int ndim, const int64_t* kshape, void* data, const double* kx, const double* ky, const double* kz,
                                 int ky_dense, const double* knyquist, void* stream) {
    if (ndim != 2 && ndim != 3) { set_error("Must use either 2 or 3 dimensions."); return -1; }
    if (!kshape || !data || !kx || !ky || !knyquist || (ndim == 3 && !kz)) { set_error("ddl_dealias_array: NULL argument"); return -1; }
    DealiasArrayF f;
    memset(&f, 0, sizeof(f));
    f.data = (cplx*)data; f.kx = kx; f.ky = ky; f.kz = kz; f.ndim = ndim; f.ky_dense = ky_dense;
    if (ndim == 3) {            // k-space axes (ky, kz, kx); knyquist in the same order (representations.py:208-211)
        f.dim[0] = (int)kshape[0]; f.dim[1] = (int)kshape[1]; f.dim[2] = (int)kshape[2];
        f.cut[0] = 2. / 3. * knyquist[2]; f.cut[1] = 2. / 3. * knyquist[0]; f.cut[2] = 2. / 3. * knyquist[1];
    } else {                    // (kx, ky)
        f.dim[0] = 1; f.dim[1] = (int)kshape[0]; f.dim[2] = (int)kshape[1];
        f.cut[0] = 2. / 3. * knyquist[0]; f.cut[1] = 2. / 3. * knyquist[1];
    }
    return launch_items(f, (long long)f.dim[0] * f.dim[1] * f.dim[2], (ddl_stream_t)stream, "dealias_array");
}

// Host <-> device copy of the listed boxes of a 3-D array (both sides in the same full layout): what a k-space component whose
// modes outside the dealias mask are known to be zero needs to cross PCIe -- the retained box (two ky ranges x two kz ranges x
// [0, kn) along kx) is 30 % of the array under the 2/3 rule.  One cudaMemcpy3DAsync per box; nothing is touched outside them.
extern "C" int ddl_copy_boxes(void* dst, const void* src, const int64_t* shape, int nbox, const int64_t* boxes, int elem_bytes,
                              int to_device, void* stream) {
    if (!dst || !src || !shape || (nbox > 0 && !boxes) || elem_bytes <= 0) { set_error("ddl_copy_boxes: NULL argument"); return -1; }
    for (int b = 0; b < nbox; ++b) {
        const int64_t* q = boxes + 6 * b;
        for (int d = 0; d < 3; ++d)
            if (q[2 * d] < 0 || q[2 * d + 1] < q[2 * d] || q[2 * d + 1] > shape[d]) { set_error("ddl_copy_boxes: box %d outside the array", b); return -1; }
    }
    const size_t pitch = (size_t)shape[2] * elem_bytes;
    for (int b = 0; b < nbox; ++b) {
        const int64_t* q = boxes + 6 * b;
        const size_t n0 = q[1] - q[0], n1 = q[3] - q[2], wb = (size_t)(q[5] - q[4]) * elem_bytes;
        if (!n0 || !n1 || !wb) continue;
#if DDL_DEVICE_BUILD
        cudaMemcpy3DParms c;
        memset(&c, 0, sizeof(c));
        c.srcPtr = make_cudaPitchedPtr(const_cast<void*>(src), pitch, pitch, (size_t)shape[1]);
        c.dstPtr = make_cudaPitchedPtr(dst, pitch, pitch, (size_t)shape[1]);
        c.srcPos = c.dstPos = make_cudaPos((size_t)q[4] * elem_bytes, (size_t)q[2], (size_t)q[0]);
        c.extent = make_cudaExtent(wb, n1, n0);
        c.kind = to_device ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToHost;
        DDL_CUDA_CHECK(cudaMemcpy3DAsync(&c, (cudaStream_t)stream));
#else
        (void)to_device; (void)stream;
        for (size_t i = 0; i < n0; ++i)
            for (size_t j = 0; j < n1; ++j) {
                const size_t off = ((size_t)(q[0] + i) * shape[1] + (q[2] + j)) * pitch + (size_t)q[4] * elem_bytes;
                memcpy((char*)dst + off, (const char*)src + off, wb);
            }
#endif
    }
    return 0;
}

extern "C" int ddl_backward(ddl_plan* pl, void* k, double* x, void* work, size_t work_bytes, void* stream) {
    ddl_stream_t st = (ddl_stream_t)stream;
    DDL_TRY(need_one_rank(pl, "ddl_backward"));
    FastGuard fg(pl->shear_on != 0);
    DDL_TRY(check_ws(pl, 1, 1, work, work_bytes));
    DDL_TRY(ddl_dealias(pl, k, stream));
    WsLayout w = ws_layout(pl, 1, 1);
    cplx* r0 = (cplx*)work; cplx* r1 = r0 + w.r0;
    const void* kin[1] = {k};
    if (pl->ndim == 3) {
        void* A[1] = {r0};
        void* B[1] = {r1};
        DDL_TRY(phase_zinv(pl, 1, kin, A, st));
        DDL_TRY(phase_yinv(pl, 1, A, B, st));
        return phase_xc2r(pl, r1, x, st);
    }
    const Axis &X = pl->ax, &Y = pl->ay;
    void* head[1];
    DDL_TRY(inverse_head_2d(pl, 1, kin, r0, head, st));
    void* xo[1] = {x};
    PhysConst pc = {};
    return pass_pair("x_c2r", X.n, TM_C2R, 0, 1, 1, head, xo, side(Y.n, 1, 0, nullptr, nullptr), side(1, X.n, 0, nullptr, nullptr),
                     Y.n, 1, X.cnt, 1.0, X.tw, pc, st, nullptr, pl);
}

extern "C" int ddl_forward(ddl_plan* pl, const double* x, void* k, void* work, size_t work_bytes, void* stream) {
    ddl_stream_t st = (ddl_stream_t)stream;
    DDL_TRY(need_one_rank(pl, "ddl_forward"));
    FastGuard fg(pl->shear_on != 0);
    DDL_TRY(check_ws(pl, 1, 1, work, work_bytes));
    WsLayout w = ws_layout(pl, 1, 1);
    cplx* r0 = (cplx*)work; cplx* r2 = r0 + w.r0 + w.r1;
    void* dst[1] = {k};
    if (pl->ndim == 3) {
        void* Cb[1] = {r2};
        void* D[1] = {r0};
        DDL_TRY(phase_xr2c(pl, x, r2, st));
        DDL_TRY(phase_yfwd(pl, 1, Cb, D, st));
        DDL_TRY(phase_zfwd(pl, 1, D, dst, true, st));
        return ddl_dealias(pl, k, stream);
    }
    const Axis &X = pl->ax, &Y = pl->ay;
    const void* xi[1] = {x};
    void* Cb[1] = {(void*)(r0 + w.r0)};
    PhysConst pc = {};
    DDL_TRY(pass_pair("x_r2c", X.n, TM_R2C, 0, 1, 1, xi, Cb, side(1, X.n, 0, nullptr, nullptr), side(Y.n, 1, 0, nullptr, nullptr),
                      Y.n, 1, X.cnt, 1.0 / (double)pl->ntot, X.tw, pc, st, nullptr, pl));
    DDL_TRY(forward_tail_2d(pl, 1, Cb, dst, true, st));
    return ddl_dealias(pl, k, stream);
}

extern "C" int ddl_deriv(ddl_plan* pl, const void* k_in, void* k_out, int axis, void* stream) {
    DerivF f;
    f.g = pl->geom; f.in = (const cplx*)k_in; f.out = (cplx*)k_out; f.level = -1;
    for (int l = 0; l < 3; ++l) if (pl->geom.ax[l] == axis) f.level = l;
    if (f.level < 0) { set_error("deriv: bad axis %d", axis); return -1; }
    return launch_items(f, pl->nmodes, (ddl_stream_t)stream, "deriv");
}

// ---------------------------------------------------------------- fused RHS
template <class PHYS>
static long long fill_assemble(ddl_plan* pl, AssembleF<PHYS>& f, void* const* E, void* const* state, void* const* deriv,
                               const PhysConst& pc) {
    const Axis &X = pl->ax, &Y = pl->ay, &Z = pl->az;
    const long long CX = X.cnt, KXP = kx_pitch(pl);
    long long count;
    if (pl->ndim == 3) {
        const long long CZ = Z.cnt, CYL = pl->yl.cyl;
        count = CYL * CZ * CX;
        f.cdim[0] = (int)CYL; f.cdim[1] = (int)CZ; f.cdim[2] = (int)CX;
        f.cstride[0] = CZ * KXP; f.cstride[1] = KXP; f.cstride[2] = 1;
        f.fstride[0] = (long long)Z.n * X.nk; f.fstride[1] = X.nk; f.fstride[2] = 1;
        f.ftab[0] = pl->yl.c2f; f.ftab[1] = Z.c2f; f.ftab[2] = nullptr;
        f.kvc[0] = pl->yl.kvc; f.kvc[1] = Z.kvc; f.kvc[2] = X.kvc;
        f.ax[0] = 1; f.ax[1] = 2; f.ax[2] = 0;
    } else {
        const long long CY = Y.cnt;
        count = CX * CY;
        f.cdim[0] = 1; f.cdim[1] = (int)CX; f.cdim[2] = (int)CY;
        f.cstride[0] = 0; f.cstride[1] = CY; f.cstride[2] = 1;
        f.fstride[0] = 0; f.fstride[1] = Y.n; f.fstride[2] = 1;
        f.ftab[0] = nullptr; f.ftab[1] = nullptr; f.ftab[2] = Y.c2f;
        f.kvc[0] = nullptr; f.kvc[1] = X.kvc; f.kvc[2] = Y.kvc;
        f.ax[0] = -1; f.ax[1] = 0; f.ax[2] = 1;
    }
    for (int i = 0; i < PHYS::NO; ++i) f.P[i] = (const cplx*)E[i];
    for (int i = 0; i < PHYS::NS; ++i) f.S[i] = (const cplx*)state[i];
    for (int i = 0; i < PHYS::NC; ++i) f.D[i] = deriv ? (cplx*)deriv[i] : nullptr;
    f.pc = pc;
    return count;
}

template <class PHYS>
static int assemble(ddl_plan* pl, void* const* E, void* const* state, void* const* deriv, const PhysConst& pc, ddl_stream_t st) {
    AssembleF<PHYS> f;
    const long long count = fill_assemble<PHYS>(pl, f, E, state, deriv, pc);
    return launch_items(f, count, st, "assemble");
}

// ddl_set_option("assemble_variant", v): 0 = loads where the arithmetic needs them, four CTAs per SM; 1 = every operand of a
// mode loaded up front (AssembleStageF<PHYS, true>), four CTAs per SM; 2, 3 = the same with three / two CTAs per SM (85 / 128
// registers per thread, so that all of a mode's loads can be in flight at once); 4 = variant 0 with two CTAs per SM.
// Measured at 512^3 MHD RK4 on a B200 (profiles/r2/devcheck_b200_xpass_pack.txt, ..._shapes.txt), ms per launch, same bits out:
// 0: 2.04   1: 3.52 (64 registers: 704 B of spills)   2: 2.68   3: 1.77   4: 2.68.  In variant 0 a mode pays one memory round
// trip for the nine product spectra and one more per component (start value, running total); variant 3 has all 21 loads of a
// mode in flight at once: 10.9 GB in 1.87 ms under ncu = 5.8 TB/s, 0.90 of the measured copy bandwidth (variant 0: 0.80).
static int g_assemble_variant = 3;
// ddl_set_option("traceless_flux", 0 | 1): the one-rank 3-D RHS with 5 momentum products (T - delta T_zz) instead of 6
static int g_traceless_flux = 1;

template <class PHYS, bool HOIST>
static int assemble_rk4_v(ddl_plan* pl, void* const* E, void* const* state, const PhysConst& pc, const ddl_stage_fuse* fu,
                          ddl_stream_t st, int minb) {
    AssembleStageF<PHYS, HOIST> f;
    const long long count = fill_assemble<PHYS>(pl, f.a, E, state, nullptr, pc);
    for (int c = 0; c < PHYS::NC; ++c) {
        f.y[c] = (const cplx*)fu->y[c]; f.out[c] = (cplx*)fu->out[c];
        f.total[c] = fu->total ? (cplx*)fu->total[c] : nullptr;
        f.d1[c] = fu->deriv1 ? (const cplx*)fu->deriv1[c] : nullptr;
        f.kout[c] = fu->k_out ? (cplx*)fu->k_out[c] : nullptr;
        f.coeff[c] = fu->coeff ? fu->coeff[c] : 0.0;
    }
    f.kind = fu->kind; f.has_d1 = fu->deriv1 != nullptr; f.has_kout = fu->k_out != nullptr;
    f.vo = fu->visc_order; f.first = fu->first; f.last = fu->last; f.twod = pl->geom.twod;
    f.dt = fu->dt_step; f.wdiv = fu->wdiv;
    if (minb == 3) return launch_items_b<AssembleStageF<PHYS, HOIST>, 3>(f, count, st, "assemble_stage");
    if (minb == 2) return launch_items_b<AssembleStageF<PHYS, HOIST>, 2>(f, count, st, "assemble_stage");
    return launch_items(f, count, st, "assemble_stage");
}

template <class PHYS>
static int assemble_rk4(ddl_plan* pl, void* const* E, void* const* state, const PhysConst& pc, const ddl_stage_fuse* fu,
                        ddl_stream_t st) {
    const int v = g_assemble_variant;
    if (v <= 0) return assemble_rk4_v<PHYS, false>(pl, E, state, pc, fu, st, 4);
    if (v == 4) return assemble_rk4_v<PHYS, false>(pl, E, state, pc, fu, st, 2);
    return assemble_rk4_v<PHYS, true>(pl, E, state, pc, fu, st, v == 2 ? 3 : (v == 3 ? 2 : 4));
}

static int assemble_rk4_any(ddl_plan* pl, int code, void* const* E, void* const* state, const PhysConst& pc,
                            const ddl_stage_fuse* fu, ddl_stream_t st) {
    if (code > 5 && code < 12) { set_error("the fused stage update exists for the solenoidal policies only (use ddl_rhs + ddl_stage)"); return -1; }
    switch (code) {
        case 12: return assemble_rk4<Hydro3T>(pl, E, state, pc, fu, st);
        case 13: return assemble_rk4<Bouss3T>(pl, E, state, pc, fu, st);
        case 14: return assemble_rk4<MHD3T>(pl, E, state, pc, fu, st);
        case 0: return assemble_rk4<Hydro2C>(pl, E, state, pc, fu, st);
        case 1: return assemble_rk4<Bouss2C>(pl, E, state, pc, fu, st);
        case 2: return assemble_rk4<MHD2C>(pl, E, state, pc, fu, st);
        case 3: return assemble_rk4<Hydro3C>(pl, E, state, pc, fu, st);
        case 4: return assemble_rk4<Bouss3C>(pl, E, state, pc, fu, st);
        default: return assemble_rk4<MHD3C>(pl, E, state, pc, fu, st);
    }
}

static int assemble_any(ddl_plan* pl, int code, void* const* E, void* const* state, void* const* deriv, const PhysConst& pc,
                        ddl_stream_t st) {
    switch (code) {
        case 0: return assemble<Hydro2C>(pl, E, state, deriv, pc, st);
        case 1: return assemble<Bouss2C>(pl, E, state, deriv, pc, st);
        case 2: return assemble<MHD2C>(pl, E, state, deriv, pc, st);
        case 3: return assemble<Hydro3C>(pl, E, state, deriv, pc, st);
        case 4: return assemble<Bouss3C>(pl, E, state, deriv, pc, st);
        case 5: return assemble<MHD3C>(pl, E, state, deriv, pc, st);
        case 6: return assemble<Hydro2A>(pl, E, state, deriv, pc, st);
        case 7: return assemble<Bouss2A>(pl, E, state, deriv, pc, st);
        case 8: return assemble<MHD2A>(pl, E, state, deriv, pc, st);
        case 9: return assemble<Hydro3A>(pl, E, state, deriv, pc, st);
        case 10: return assemble<Bouss3A>(pl, E, state, deriv, pc, st);
        case 12: return assemble<Hydro3T>(pl, E, state, deriv, pc, st);
        case 13: return assemble<Bouss3T>(pl, E, state, deriv, pc, st);
        case 14: return assemble<MHD3T>(pl, E, state, deriv, pc, st);
        default: return assemble<MHD3A>(pl, E, state, deriv, pc, st);
    }
}

static int check_fuse(const ddl_stage_fuse* fu) {
    if (!fu || !fu->y || !fu->out) { set_error("fused stage: incomplete descriptor"); return -1; }
    if (fu->kind < DDL_EULER || fu->kind > DDL_FUSE_CN) { set_error("fused stage: bad kind %d", fu->kind); return -1; }
    if (fu->kind == DDL_FUSE_RK4 && !fu->total) { set_error("fused RK4 stage needs the total arrays"); return -1; }
    if ((fu->kind == DDL_ETD2RK1 || fu->kind == DDL_ETD2RK2) && !fu->deriv1) { set_error("fused ETD2 stage needs deriv1"); return -1; }
    return 0;
}

static int check_physics(const ddl_plan* pl, int physics, const ddl_phys_params* prm) {
    if (physics < 0 || physics > DDL_MHD_ADV) { set_error("unknown physics id %d", physics); return -1; }
    if ((physics == DDL_BOUSSINESQ || physics == DDL_BOUSSINESQ_ADV) && (prm->boussinesq_dir < 0 || prm->boussinesq_dir >= pl->ndim)) {
        set_error("boussinesq_direction component %d not present in %d-D", prm->boussinesq_dir, pl->ndim);
        return -1;
    }
    return 0;
}

static int rhs_impl(ddl_plan* pl, int physics, const ddl_phys_params* prm, void* const* state, void* const* deriv,
                    void* work, size_t work_bytes, int flags, const ddl_stage_fuse* fuse, void* stream, bool front_only = false) {
    ddl_stream_t st = (ddl_stream_t)stream;
    int ni, no, code, ncomp;
    DDL_TRY(need_one_rank(pl, "ddl_rhs"));
    DDL_TRY(check_physics(pl, physics, prm));
    phys_counts(pl->ndim, physics, ni, no, code, &ncomp);
    DDL_TRY(check_ws(pl, ni, no, work, work_bytes));
    // one rank, 3-D, solenoidal policy: the traceless-flux twin (physics_ops.cuh) -- T_zz is never formed or transformed, one
    // product field fewer through the x pass, the forward y and z passes and the assembly (the workspace was sized for `no`)
    if (g_traceless_flux && pl->ndim == 3 && code >= 3 && code <= 5) { code += 9; no -= 1; }
    const PhysConst pc = phys_const(prm);
    if (flags & DDL_RHS_DEALIAS_STATE) DDL_TRY(mask_arrays(pl, ncomp, state, st));
    if ((flags & DDL_RHS_ZERO_FILL) && deriv) DDL_TRY(mask_arrays(pl, ncomp, deriv, st));
    DDL_TRY(phase_theta(pl, physics, state, st));     // advective-form policies: divergence spectra into the scratch arrays

    WsLayout w = ws_layout(pl, ni, no);
    cplx* r0 = (cplx*)work; cplx* r1 = r0 + w.r0; cplx* r2 = r1 + w.r1;
    const Axis &X = pl->ax, &Y = pl->ay;
    std::vector<void*> A(ni), B(ni), C(no), D(no), E(no);
    if (pl->ndim == 3) {
        const SlabSizes s = slab_sizes(pl);
        for (int f = 0; f < ni; ++f) { A[f] = r0 + f * s.ks; B[f] = r1 + f * s.b; }
        for (int f = 0; f < no; ++f) { C[f] = r2 + f * s.b; D[f] = r0 + f * s.ks; E[f] = r1 + f * s.e; }
        DDL_TRY(phase_zinv(pl, ni, (const void* const*)state, A.data(), st));
        // Opt-in experiment (ddl_set_option("rhs_plane_chunk", n), default off; measured in round 2: 12.2-18.9 ms per RHS against 10.6 ms, profiles/r2/devcheck_b200_r2_first.txt): y_inv -> x -> y_fwd over chunks of n
        // z-planes with chunk-sized b / c arrays that every chunk reuses, so that the half-transformed lines (22 of the 52 GB a
        // 512^3 MHD stage moves) can live in the 126 MB L2 instead of crossing HBM twice.  15 fields x 1.44 MB per plane: n <= 4.
        const long long cb = (long long)g_plane_chunk * Y.n * kx_pitch(pl);
        if (g_plane_chunk > 0 && !front_only && (long long)(ni + no) * cb <= w.r1 && (long long)no * s.xs <= w.r2) {
            std::vector<void*> Bc(ni), Cc(no);
            for (int f = 0; f < ni; ++f) Bc[f] = r1 + f * cb;
            for (int f = 0; f < no; ++f) { Cc[f] = r1 + (ni + f) * cb; D[f] = r2 + f * s.xs; }
            for (int z0 = 0; z0 < pl->nzl; z0 += g_plane_chunk) {
                const int nzc = pl->nzl - z0 < g_plane_chunk ? pl->nzl - z0 : g_plane_chunk;
                DDL_TRY(phase_yinv_planes(pl, ni, A.data(), Bc.data(), z0, nzc, st));
                DDL_TRY(phase_xfused(pl, code, ni, no, Bc.data(), Cc.data(), pc, st, 0, nzc));
                DDL_TRY(phase_yfwd_planes(pl, no, Cc.data(), D.data(), z0, nzc, st));
            }
            DDL_TRY(phase_zfwd(pl, no, D.data(), E.data(), false, st));
        } else {
            DDL_TRY(phase_yinv(pl, ni, A.data(), B.data(), st));
            DDL_TRY(phase_xfused(pl, code, ni, no, B.data(), C.data(), pc, st));
            if (front_only) return 0;
            DDL_TRY(phase_yfwd(pl, no, C.data(), D.data(), st));
            DDL_TRY(phase_zfwd(pl, no, D.data(), E.data(), false, st));
        }
    } else {
        const double sc = 1.0 / (double)pl->ntot;
        DDL_TRY(inverse_head_2d(pl, ni, (const void* const*)state, r0, A.data(), st));
        const long long per = (long long)X.cnt * Y.n, pere = (long long)X.cnt * Y.cnt;
        for (int f = 0; f < no; ++f) { C[f] = r1 + f * per; E[f] = r2 + f * pere; }
        TileSide s = side(Y.n, 1, 0, nullptr, nullptr);
        DDL_TRY(pass_pair("x_fused", X.n, TM_FUSED, code, ni, no, A.data(), C.data(), s, s, Y.n, 1, X.cnt, sc, X.tw, pc, st, pl->cfl_out));
        if (front_only) return 0;
        DDL_TRY(forward_tail_2d(pl, no, C.data(), E.data(), false, st));
    }
    if (fuse) return assemble_rk4_any(pl, code, E.data(), state, pc, fuse, st);
    return assemble_any(pl, code, E.data(), state, deriv, pc, st);
}

extern "C" int ddl_rhs(ddl_plan* pl, int physics, const ddl_phys_params* prm, void* const* state, void* const* deriv,
                       void* work, size_t work_bytes, int flags, void* stream) {
    return rhs_impl(pl, physics, prm, state, deriv, work, work_bytes, flags, nullptr, stream);
}

extern "C" int ddl_rhs_stage(ddl_plan* pl, int physics, const ddl_phys_params* prm, void* const* state, void* work,
                           size_t work_bytes, int flags, const ddl_stage_fuse* fuse, void* stream) {
    DDL_TRY(check_fuse(fuse));
    return rhs_impl(pl, physics, prm, state, nullptr, work, work_bytes, flags, fuse, stream);
}

// ---------------------------------------------------------------- slab phase API (include/ddl.h)
extern "C" int ddl_slab_info(const ddl_plan* pl, int64_t* out) {
    const SlabSizes s = slab_sizes(pl);
    out[0] = pl->nranks; out[1] = pl->rank; out[2] = pl->nzl; out[3] = pl->yl.nyl; out[4] = pl->yl.cyl; out[5] = pl->yl.cy0;
    out[6] = pl->ay.cnt; out[7] = pl->az.cnt; out[8] = kx_pitch(pl); out[9] = pl->ax.cnt;
    out[10] = s.ks; out[11] = s.xs; out[12] = s.b; out[13] = s.e; out[14] = pl->z0; out[15] = pl->yl.ky0;
    out[16] = pl->layout;
    return 0;
}
extern "C" int ddl_slab_rows(const ddl_plan* pl, int64_t* cyl_of_rank) {
    for (int r = 0; r < pl->nranks; ++r) cyl_of_rank[r] = pl->cyl_of[r];
    return 0;
}
static int need_3d(const ddl_plan* pl) {
    if (pl->ndim != 3) { set_error("the slab phase API is 3-D only"); return -1; }
    return 0;
}
extern "C" int ddl_slab_theta(ddl_plan* pl, int physics, void* const* state, void* stream) {
    if (physics < 0 || physics > DDL_MHD_ADV) { set_error("unknown physics id %d", physics); return -1; }
    return phase_theta(pl, physics, state, (ddl_stream_t)stream);
}
extern "C" int ddl_slab_zinv(ddl_plan* pl, int nf, void* const* k_in, void* const* ks_out, void* stream) {
    DDL_TRY(need_3d(pl));
    FastGuard fg(pl->ay.full);
    return phase_zinv(pl, nf, (const void* const*)k_in, ks_out, (ddl_stream_t)stream);
}
extern "C" int ddl_slab_zinv_peer(ddl_plan* pl, int nf, void* const* k_in, void* const* peer_tab, void* stream) {
    DDL_TRY(need_3d(pl));
    return phase_zinv_peer(pl, nf, (const void* const*)k_in, peer_tab, (ddl_stream_t)stream);
}
static int check_planes(const ddl_plan* pl, int z0, int nzc) {
    if (z0 < 0 || nzc < 0 || z0 + nzc > pl->nzl) { set_error("plane range [%d, %d) outside the local slab of %d planes", z0, z0 + nzc, pl->nzl); return -1; }
    return 0;
}
extern "C" int ddl_slab_yfwd_peer(ddl_plan* pl, int nf, void* const* c_in, void* const* peer_tab, int z0, int nzc, void* stream) {
    DDL_TRY(need_3d(pl));
    DDL_TRY(check_planes(pl, z0, nzc));
    return phase_yfwd_peer(pl, nf, (const void* const*)c_in, peer_tab, z0, nzc, (ddl_stream_t)stream);
}
extern "C" int ddl_slab_yfwd_planes(ddl_plan* pl, int nf, void* const* c_in, void* const* xs_out, int z0, int nzc, void* stream) {
    DDL_TRY(need_3d(pl));
    DDL_TRY(check_planes(pl, z0, nzc));
    FastGuard fg(pl->ay.full);
    const long long CX = kx_pitch(pl);
    std::vector<const void*> cin(nf);
    for (int f = 0; f < nf; ++f) cin[f] = (const cplx*)c_in[f] + (long long)z0 * pl->ay.n * CX;      // phase_yfwd_planes takes chunk-local inputs
    return phase_yfwd_planes(pl, nf, cin.data(), xs_out, z0, nzc, (ddl_stream_t)stream);
}
extern "C" int ddl_slab_xfused_planes(ddl_plan* pl, int physics, const ddl_phys_params* prm, void* const* b_in, void* const* c_out,
                                      int z0, int nzc, void* stream) {
    int ni, no, code;
    DDL_TRY(need_3d(pl));
    DDL_TRY(check_physics(pl, physics, prm));
    DDL_TRY(check_planes(pl, z0, nzc));
    phys_counts(3, physics, ni, no, code);
    return phase_xfused(pl, code, ni, no, (const void* const*)b_in, c_out, phys_const(prm), (ddl_stream_t)stream, z0, nzc);
}
extern "C" int ddl_slab_yinv(ddl_plan* pl, int nf, void* const* xs_in, void* const* b_out, void* stream) {
    DDL_TRY(need_3d(pl));
    FastGuard fg(pl->ay.full);
    return phase_yinv(pl, nf, (const void* const*)xs_in, b_out, (ddl_stream_t)stream);
}
extern "C" int ddl_slab_xfused(ddl_plan* pl, int physics, const ddl_phys_params* prm, void* const* b_in, void* const* c_out,
                               void* stream) {
    int ni, no, code;
    DDL_TRY(need_3d(pl));
    DDL_TRY(check_physics(pl, physics, prm));
    phys_counts(3, physics, ni, no, code);
    return phase_xfused(pl, code, ni, no, (const void* const*)b_in, c_out, phys_const(prm), (ddl_stream_t)stream);
}
extern "C" int ddl_slab_xc2r(ddl_plan* pl, const void* b_in, double* x_out, void* stream) {
    DDL_TRY(need_3d(pl));
    return phase_xc2r(pl, b_in, x_out, (ddl_stream_t)stream);
}
extern "C" int ddl_slab_xr2c(ddl_plan* pl, const double* x_in, void* c_out, void* stream) {
    DDL_TRY(need_3d(pl));
    return phase_xr2c(pl, x_in, c_out, (ddl_stream_t)stream);
}
extern "C" int ddl_slab_yfwd(ddl_plan* pl, int nf, void* const* c_in, void* const* xs_out, void* stream) {
    DDL_TRY(need_3d(pl));
    FastGuard fg(pl->ay.full);
    return phase_yfwd(pl, nf, (const void* const*)c_in, xs_out, (ddl_stream_t)stream);
}
extern "C" int ddl_slab_zfwd(ddl_plan* pl, int nf, void* const* ks_in, void* const* out, int full_out, void* stream) {
    DDL_TRY(need_3d(pl));
    FastGuard fg(pl->ay.full);
    return phase_zfwd(pl, nf, (const void* const*)ks_in, out, full_out != 0, (ddl_stream_t)stream);
}
extern "C" int ddl_slab_assemble_stage(ddl_plan* pl, int physics, const ddl_phys_params* prm, void* const* e_in, void* const* state,
                                     const ddl_stage_fuse* fuse, void* stream) {
    int ni, no, code;
    DDL_TRY(need_3d(pl));
    DDL_TRY(check_physics(pl, physics, prm));
    DDL_TRY(check_fuse(fuse));
    phys_counts(3, physics, ni, no, code);
    return assemble_rk4_any(pl, code, e_in, state, phys_const(prm), fuse, (ddl_stream_t)stream);
}
extern "C" int ddl_slab_assemble(ddl_plan* pl, int physics, const ddl_phys_params* prm, void* const* e_in, void* const* state,
                                 void* const* deriv, void* stream) {
    int ni, no, code;
    DDL_TRY(need_3d(pl));
    DDL_TRY(check_physics(pl, physics, prm));
    phys_counts(3, physics, ni, no, code);
    return assemble_any(pl, code, e_in, state, deriv, phys_const(prm), (ddl_stream_t)stream);
}

// ---------------------------------------------------------------- reductions (include/ddl.h)
extern "C" int ddl_rhs_capture_max(ddl_plan* pl, double* out2) {
    pl->cfl_out = out2;
    return 0;
}

static int zero_doubles(double* p, int n, ddl_stream_t st) {
#if DDL_DEVICE_BUILD
    DDL_CUDA_CHECK(cudaMemsetAsync(p, 0, n * sizeof(double), st));
#else
    (void)st;
    memset(p, 0, n * sizeof(double));
#endif
    return 0;
}

extern "C" int ddl_reduce_max_square(ddl_plan* pl, int physics, const ddl_phys_params* prm, void* const* state, void* work,
                                     size_t work_bytes, int flags, double* out2, void* stream) {
    if (!out2) { set_error("ddl_reduce_max_square: out2 is NULL"); return -1; }
    DDL_TRY(need_one_rank(pl, "ddl_reduce_max_square"));
    DDL_TRY(zero_doubles(out2, 2, (ddl_stream_t)stream));
    double* saved = pl->cfl_out;
    pl->cfl_out = out2;
    const int rc = rhs_impl(pl, physics, prm, state, nullptr, work, work_bytes, flags & DDL_RHS_DEALIAS_STATE, nullptr, stream, true);
    pl->cfl_out = saved;
    return rc;
}

template <int ND, int NB>
static int invariants_t(ddl_plan* pl, void* const* state, int flags, double* out, ddl_stream_t st) {
    InvariantsF<ND, NB> f;
    memset(&f, 0, sizeof(f));
    for (int c = 0; c < ND + NB; ++c) f.S[c] = (const cplx*)state[c];
    f.g = pl->geom;
    long long count = pl->nmodes;
    if (flags & DDL_STAGE_RETAINED_ONLY) {
        const Axis &X = pl->ax, &Y = pl->ay, &Z = pl->az;
        f.compact = 1;
        if (ND == 3) {
            f.cdim[0] = pl->yl.cyl; f.cdim[1] = Z.cnt; f.cdim[2] = X.cnt;
            f.fstride[0] = (long long)Z.n * X.nk; f.fstride[1] = X.nk; f.fstride[2] = 1;
            f.ftab[0] = pl->yl.c2f; f.ftab[1] = Z.c2f; f.ftab[2] = nullptr;
            f.kvc[0] = pl->yl.kvc; f.kvc[1] = Z.kvc; f.kvc[2] = X.kvc;
        } else {
            f.cdim[0] = 1; f.cdim[1] = X.cnt; f.cdim[2] = Y.cnt;
            f.fstride[0] = 0; f.fstride[1] = Y.n; f.fstride[2] = 1;
            f.ftab[0] = nullptr; f.ftab[1] = nullptr; f.ftab[2] = Y.c2f;
            f.kvc[0] = nullptr; f.kvc[1] = X.kvc; f.kvc[2] = Y.kvc;
        }
        count = (long long)f.cdim[0] * f.cdim[1] * f.cdim[2];
    }
    if (count <= 0) return zero_doubles(out, DDL_NINV, st);      // a rank may own no retained ky row
    return launch_reduce<InvariantsF<ND, NB>, RED_SUM>(f, count, pl->red_partial, out, 0, st, "invariants");
}

extern "C" int ddl_reduce_outside_mask(ddl_plan* pl, int n, void* const* arrays, double* out, void* stream) {
    if (!pl || !arrays || !out) { set_error("ddl_reduce_outside_mask: NULL argument"); return -1; }
    if (n < 1 || n > DDL_MAXC) { set_error("ddl_reduce_outside_mask: %d arrays (1..%d)", n, DDL_MAXC); return -1; }
    OutsideF f;
    memset(&f, 0, sizeof(f));
    f.g = pl->geom; f.narr = n;
    for (int i = 0; i < n; ++i) f.arr[i] = (const cplx*)arrays[i];
    return launch_reduce<OutsideF, RED_SUM>(f, pl->nmodes, pl->red_partial, out, 0, (ddl_stream_t)stream, "outside_mask");
}

extern "C" int ddl_reduce_invariants(ddl_plan* pl, int physics, void* const* state, int flags, double* out, void* stream) {
    static_assert(DDL_NINV == DDL_NINV_, "include/ddl.h and reduce.cuh disagree on the invariant count");
    ddl_stream_t st = (ddl_stream_t)stream;
    if (!out || !state) { set_error("ddl_reduce_invariants: NULL argument"); return -1; }
    if (physics < 0 || physics > DDL_MHD_ADV) { set_error("unknown physics id %d", physics); return -1; }
    if (physics >= DDL_HYDRO_ADV) physics -= DDL_HYDRO_ADV;
    if (pl->ndim == 3) {
        if (physics == DDL_HYDRO) return invariants_t<3, 0>(pl, state, flags, out, st);
        if (physics == DDL_BOUSSINESQ) return invariants_t<3, 1>(pl, state, flags, out, st);
        return invariants_t<3, 3>(pl, state, flags, out, st);
    }
    if (physics == DDL_HYDRO) return invariants_t<2, 0>(pl, state, flags, out, st);
    if (physics == DDL_BOUSSINESQ) return invariants_t<2, 1>(pl, state, flags, out, st);
    return invariants_t<2, 2>(pl, state, flags, out, st);
}

// ---------------------------------------------------------------- stage updates
extern "C" int ddl_step_array(int kind, int ndim, long long count, const void* start, void* out, const void* deriv1, const void* deriv2,
                              const double* intfactor, double dt, void* stream) {
    if (kind < DDL_EULER || kind > DDL_ETD2RK2) { set_error("ddl_step_array: kind %d is not euler / etd1 / etd2rk1 / etd2rk2", kind); return -1; }
    if (ndim != 2 && ndim != 3) { set_error("Must use either 2 or 3 dimensions."); return -1; }
    if (!start || !out || !deriv1 || ((kind == DDL_ETD2RK1 || kind == DDL_ETD2RK2) && !deriv2)) {
        set_error("ddl_step_array: NULL array"); return -1;
    }
    if (count <= 0) return 0;
    StageArrayF f;
    f.start = (const cplx*)start; f.out = (cplx*)out; f.d1 = (const cplx*)deriv1; f.d2 = (const cplx*)deriv2;
    f.intfactor = intfactor; f.kind = kind; f.twod = ndim == 2; f.dt = dt;
    return launch_items(f, count, (ddl_stream_t)stream, "stage_array");
}

static int fill_stage(ddl_plan* pl, StageArgs& a, int ncomp, const double* coeff, int vo, int flags, long long& count) {
    if (ncomp < 1 || ncomp > DDL_MAXC) { set_error("ncomp %d out of range 1..%d", ncomp, DDL_MAXC); return -1; }
    memset(&a, 0, sizeof(a));
    a.g = pl->geom; a.ncomp = ncomp; a.vo = vo;
    for (int c = 0; c < ncomp; ++c) a.coeff[c] = coeff ? coeff[c] : 0.0;
    count = pl->nmodes;
    if (flags & DDL_STAGE_RETAINED_ONLY) {
        const Axis &X = pl->ax, &Y = pl->ay, &Z = pl->az;
        a.compact = 1;
        if (pl->ndim == 3) {
            a.cdim[0] = pl->yl.cyl; a.cdim[1] = Z.cnt; a.cdim[2] = X.cnt;
            a.fstride[0] = (long long)Z.n * X.nk; a.fstride[1] = X.nk; a.fstride[2] = 1;
            a.ftab[0] = pl->yl.c2f; a.ftab[1] = Z.c2f; a.ftab[2] = nullptr;
            a.kvc[0] = pl->yl.kvc; a.kvc[1] = Z.kvc; a.kvc[2] = X.kvc;
        } else {
            a.cdim[0] = 1; a.cdim[1] = X.cnt; a.cdim[2] = Y.cnt;
            a.fstride[0] = 0; a.fstride[1] = Y.n; a.fstride[2] = 1;
            a.ftab[0] = nullptr; a.ftab[1] = nullptr; a.ftab[2] = Y.c2f;
            a.kvc[0] = nullptr; a.kvc[1] = X.kvc; a.kvc[2] = Y.kvc;
        }
        count = (long long)a.cdim[0] * a.cdim[1] * a.cdim[2];
    }
    return 0;
}

extern "C" int ddl_stage(ddl_plan* pl, int kind, int ncomp, void* const* start, void* const* out, void* const* d1,
                         void* const* d2, const double* coeff, int visc_order, double dt, int flags, void* stream) {
    StageF f;
    long long count;
    DDL_TRY(fill_stage(pl, f.a, ncomp, coeff, visc_order, flags, count));
    if (kind < DDL_EULER || kind > DDL_ETD2RK2) { set_error("bad stage kind %d", kind); return -1; }
    if ((kind == DDL_ETD2RK1 || kind == DDL_ETD2RK2) && !d2) { set_error("stage kind %d needs deriv2", kind); return -1; }
    f.a.kind = kind; f.a.dt = dt;
    for (int c = 0; c < ncomp; ++c) {
        f.a.start[c] = (const cplx*)start[c]; f.a.out[c] = (cplx*)out[c]; f.a.d1[c] = (const cplx*)d1[c];
        f.a.d2[c] = d2 ? (const cplx*)d2[c] : nullptr;
    }
    return launch_items(f, count, (ddl_stream_t)stream, "stage");
}

// kind: DDL_EULER .. DDL_ETD2RK2, DDL_FUSE_RK4, DDL_FUSE_CN (the kinds of ddl_rhs_stage)
extern "C" int ddl_stage_outside(ddl_plan* pl, int kind, int ncomp, void* const* start, void* const* out, const double* coeff,
                                 int visc_order, double dt, void* stream) {
    OutsideStageF f;
    long long count;
    DDL_TRY(fill_stage(pl, f.a, ncomp, coeff, visc_order, 0, count));
    if (kind < DDL_EULER || kind > DDL_FUSE_CN) { set_error("bad stage kind %d", kind); return -1; }
    f.a.kind = kind; f.a.dt = dt; f.a.wdiv = 1.0;
    for (int c = 0; c < ncomp; ++c) { f.a.start[c] = (const cplx*)start[c]; f.a.out[c] = (cplx*)out[c]; }
    return launch_items(f, count, (ddl_stream_t)stream, "stage_outside");
}

extern "C" int ddl_rk4_stage(ddl_plan* pl, int ncomp, void* const* y, void* const* k, void* const* total, void* const* out,
                             const double* coeff, int visc_order, double wdiv, double dt_step, int first, int last,
                             int flags, void* stream) {
    StageF f;
    long long count;
    DDL_TRY(fill_stage(pl, f.a, ncomp, coeff, visc_order, flags, count));
    f.a.kind = SK_RK4; f.a.dt = dt_step; f.a.wdiv = wdiv; f.a.first = first; f.a.last = last;
    for (int c = 0; c < ncomp; ++c) {
        f.a.start[c] = (const cplx*)y[c]; f.a.out[c] = (cplx*)out[c]; f.a.d1[c] = (const cplx*)k[c];
        f.a.total[c] = (cplx*)total[c];
    }
    return launch_items(f, count, (ddl_stream_t)stream, "stage");
}

extern "C" int ddl_cn_step(ddl_plan* pl, int ncomp, void* const* y, void* const* k, const double* coeff, int visc_order,
                           double dt, int flags, void* stream) {
    StageF f;
    long long count;
    DDL_TRY(fill_stage(pl, f.a, ncomp, coeff, visc_order, flags, count));
    f.a.kind = SK_CN; f.a.dt = dt;
    for (int c = 0; c < ncomp; ++c) { f.a.start[c] = (const cplx*)y[c]; f.a.out[c] = (cplx*)y[c]; f.a.d1[c] = (const cplx*)k[c]; }
    return launch_items(f, count, (ddl_stream_t)stream, "stage");
}

extern "C" int ddl_sync(void* stream) {
#if DDL_DEVICE_BUILD
    DDL_CUDA_CHECK(cudaStreamSynchronize((cudaStream_t)stream));
#else
    (void)stream;
#endif
    return 0;
}

// shearing box: phase factors of the following ddl_forward / ddl_backward calls on this plan (include/ddl.h)
extern "C" int ddl_set_shear(ddl_plan* pl, int enable, double shear_rate, double time, double dy) {
    if (!pl) { set_error("ddl_set_shear: NULL plan"); return -1; }
    pl->shear_on = enable ? 1 : 0;
    pl->sh_S = shear_rate; pl->sh_t = time; pl->sh_dy = dy;
    return 0;
}

namespace ddl { extern int g_p2p_timeout_s; }
#if DDL_DEVICE_BUILD
namespace ddl { int g_peer_pass_ctas = 0; int g_strided_staged = 0; int g_persist_stagger_ns = 0; extern int g_push_tma; }
#endif
extern "C" int ddl_set_option(const char* name, int value) {
#if DDL_DEVICE_BUILD
    if (name && !strcmp(name, "peer_pass_ctas")) { ddl::g_peer_pass_ctas = value < 0 ? 0 : value; return 0; }
    if (name && !strcmp(name, "strided_staged")) { ddl::g_strided_staged = value != 0; return 0; }
    if (name && !strcmp(name, "push_tma")) { ddl::g_push_tma = value != 0; return 0; }
    if (name && !strcmp(name, "persist_stagger_ns")) { ddl::g_persist_stagger_ns = value < 0 ? 0 : value; return 0; }
#else
    if (name && (!strcmp(name, "peer_pass_ctas") || !strcmp(name, "strided_staged") || !strcmp(name, "push_tma") ||
                 !strcmp(name, "persist_stagger_ns"))) return 0;
#endif
    if (name && !strcmp(name, "fast_kernels")) { g_use_fast = value; return 0; }
    if (name && !strcmp(name, "p2p_timeout_s")) { ddl::g_p2p_timeout_s = value < 0 ? 0 : value; return 0; }
    if (name && !strcmp(name, "xfused_variant")) { g_xfused_variant = value; return 0; }
    if (name && !strcmp(name, "assemble_variant")) { g_assemble_variant = value; return 0; }
    if (name && !strcmp(name, "strided_two")) { g_strided_two = value; return 0; }
    if (name && !strcmp(name, "traceless_flux")) { g_traceless_flux = value != 0; return 0; }
    if (name && !strcmp(name, "rhs_plane_chunk")) { g_plane_chunk = value < 0 ? 0 : value; return 0; }
    set_error("unknown option %s", name ? name : "(null)");
    return -1;
}

extern "C" long long ddl_launch_count(void) { return g_launches; }

extern "C" int ddl_profile_enable(int on) {
    g_prof_on = on != 0;
    return 0;
}

// Aggregate the recorded launches by kernel label into `buf` as JSON
// {"label": {"n": launches, "ms": total device ms}, ...} and clear the record.
extern "C" int ddl_profile_report(char* buf, size_t nbuf) {
    std::vector<std::pair<std::string, std::pair<long long, double>>> agg;
#if DDL_DEVICE_BUILD
    DDL_CUDA_CHECK(cudaDeviceSynchronize());
    for (auto& e : g_prof) {
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e.a, e.b);
        cudaEventDestroy(e.a);
        cudaEventDestroy(e.b);
        bool found = false;
        for (auto& a : agg) if (a.first == e.name) { a.second.first++; a.second.second += ms; found = true; break; }
        if (!found) agg.push_back({e.name, {1, (double)ms}});
    }
#else
    if (!g_prof.empty() && g_prof.back().b == g_prof.back().a) g_prof.back().b = prof_now();
    for (auto& e : g_prof) {
        const double ms = (e.b - e.a) * 1e3;
        bool found = false;
        for (auto& a : agg) if (a.first == e.name) { a.second.first++; a.second.second += ms; found = true; break; }
        if (!found) agg.push_back({e.name, {1, ms}});
    }
#endif
    g_prof.clear();
    std::string out = "{";
    for (size_t i = 0; i < agg.size(); ++i) {
        char tmp[256];
        snprintf(tmp, sizeof(tmp), "%s\"%s\": {\"n\": %lld, \"ms\": %.6f}", i ? ", " : "", agg[i].first.c_str(),
                 agg[i].second.first, agg[i].second.second);
        out += tmp;
    }
    out += "}";
    if (out.size() + 1 > nbuf) { set_error("profile buffer too small"); return -1; }
    memcpy(buf, out.c_str(), out.size() + 1);
    return 0;
}

extern "C" const char* ddl_last_error(void) { return g_err; }
extern "C" const char* ddl_version(void) {
#if DDL_DEVICE_BUILD
    return "ddl-b200 0.1 (sm_100a)";
#else
    return "ddl-b200 0.1 (host emulation, tests only)";
#endif
}
