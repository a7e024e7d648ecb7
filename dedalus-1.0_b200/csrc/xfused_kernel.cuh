// Specialised fused x-pass of the 3-D RHS (the dominant kernel of the step):
//
//   half-spectra of a PAIR of adjacent x-lines  --pack-->  inverse FFT (DIF, natural -> scrambled)
//     -->  real-space products at every grid point  -->  forward FFT (DIT, scrambled -> natural)
//     --unpack-->  retained half-spectra of the product fields
//
// for all NI input and NO output fields of one line pair per CTA; the real-space fields never
// leave the SM.  Same mathematics as tile_kernel.cuh TM_FUSED (which stays as the generic
// implementation and as the comparison in the tests); what differs is the mapping:
//   * factorisation N = R0 x R1 (x R2) x 2 with radix-16/8/4 register butterflies, so a
//     512-point pencil crosses shared memory twice per transform instead of three times;
//   * the last inverse stage (radix 2), the products and the first forward stage (radix 2) are
//     one phase over 8-byte halves (one real line each), so the product fields are born in
//     registers and stored once;
//   * the first inverse stage reads the Hermitian-packed pencil straight from global memory
//     (both lines, mirrored half included) - no staging pass;
//   * the last forward stage leaves Z[k] and its mirror Z[N-k] in registers of lanes a and
//     Q0-a of the same warp: the Hermitian unpack is a warp shuffle, the spectra go registers
//     -> global (no shared-memory pass);
//   * a pencil belongs to ONE group of N/R0 threads (a warp for N = 512) for all of its outer
//     stages, so those stages synchronise with __syncwarp / a named barrier; only the product
//     phase is bracketed by CTA-wide barriers;
//   * XOR-swizzled pencils: every 16-byte and 8-byte access pattern of every phase is
//     bank-conflict free for N <= 512 (profiles/ carries the model and the measured counters);
//   * twiddles of a butterfly are powers of ONE table entry, generated in registers
//     (the FP64 pipe has the headroom, the load/store pipe does not).
// Replaces, per RHS, the x-direction halves of the reference's 15+21 transforms and every
// real-space numpy pass between them (physics.py:197-228, 309-354; representations.py:318-357).
#pragma once
#include "ddl_common.cuh"
#include "physics_ops.cuh"

namespace ddl {

#define DDL_XF_MAXI 8
#define DDL_XF_MAXO 12

struct XFusedParams {
    const cplx* in[DDL_XF_MAXI];    // [outer][line][pitch] half-spectra, kn retained modes per line
    cplx* out[DDL_XF_MAXO];
    long long pitch;                // elements between consecutive lines (same for in and out)
    long long s_outer;              // elements between consecutive outer planes
    int n_lines;                    // lines per outer plane (even)
    int kn;                         // retained non-negative modes along x (kn - 1 < N/2)
    double scale;                   // forward normalisation 1/N_total
    const cplx* tw;                 // exp(-2 pi i m / N)
    PhysConst pc;
    double* cfl;                    // CFL capture (kernels instantiated with CFL = true): cfl[0] = max(cfl[0], max_{x,i} u_i(x)^2),
                                    // cfl[1] likewise for the second group (B or T); fields.py:153-157 max_square
};

// Outer (shared-memory) stages of the length-N transform; the innermost radix-2 stage is fused
// with the products.
template <int N> struct XFac {
    static constexpr bool ok = false;
    static constexpr int S = 0;
    static constexpr int radix(int) { return 1; }
};
template <> struct XFac<128> { static constexpr bool ok = true; static constexpr int S = 2; static constexpr int radix(int s) { return 8; } };
template <> struct XFac<256> { static constexpr bool ok = true; static constexpr int S = 2; static constexpr int radix(int s) { return s == 0 ? 16 : 8; } };
template <> struct XFac<512> { static constexpr bool ok = true; static constexpr int S = 2; static constexpr int radix(int s) { return 16; } };
template <> struct XFac<1024> { static constexpr bool ok = true; static constexpr int S = 3; static constexpr int radix(int s) { return s == 0 ? 16 : (s == 1 ? 8 : 4); } };
template <int N> constexpr int xfac_P(int s) { return s == 0 ? 1 : xfac_P<N>(s - 1) * XFac<N>::radix(s - 1); }

// 16-byte element e of a pencil lives at swizzled slot xsw(e): a permutation inside aligned
// groups of 8 elements (one 128-byte bank row), linear over GF(2).  With bits 3-5 and 6-8 folded
// into the low three bits every access pattern of every stage (strides N/R0, 2R, 2, ...; 16- and
// 8-byte accesses) is conflict-free for N = 128 ... 1024 (profiles/xfused_banks.py).
template <int N> DDL_HD int xsw(int e) {
    return e ^ ((e >> 3) & 7) ^ ((e >> 6) & 7);
}

// ---------------------------------------------------------------------------------------
// radix-16 butterfly as 4 x 4.  Input v[j] natural; output X[r] is left in v[reg16(r)].
// ---------------------------------------------------------------------------------------
DDL_HD constexpr int reg16(int r) { return (r >> 2) + 4 * (r & 3); }

template <int DIR> DDL_HD cplx mul_w16(cplx a, int e) {
    // a * exp(DIR * 2 pi i e / 16) for the exponents the 4x4 split needs
    const double c1 = 0.92387953251128675613, s1 = 0.38268343236508977173, h = 0.70710678118654752440;
    switch (e) {
        case 1: return DIR < 0 ? mk(a.x * c1 + a.y * s1, a.y * c1 - a.x * s1) : mk(a.x * c1 - a.y * s1, a.y * c1 + a.x * s1);
        case 2: return DIR < 0 ? mk((a.x + a.y) * h, (a.y - a.x) * h) : mk((a.x - a.y) * h, (a.x + a.y) * h);
        case 3: return DIR < 0 ? mk(a.x * s1 + a.y * c1, a.y * s1 - a.x * c1) : mk(a.x * s1 - a.y * c1, a.y * s1 + a.x * c1);
        case 4: return mul_i<DIR>(a);
        case 6: return DIR < 0 ? mk((a.y - a.x) * h, -(a.x + a.y) * h) : mk(-(a.x + a.y) * h, (a.x - a.y) * h);
        case 9: return DIR < 0 ? mk(-a.x * c1 - a.y * s1, a.x * s1 - a.y * c1) : mk(a.y * s1 - a.x * c1, -a.x * s1 - a.y * c1);
        default: return a;
    }
}

template <int DIR> DDL_HD void dft16(cplx (&v)[16]) {
    // inner DFT4 over i2 for each j: elements j, j+4, j+8, j+12 -> output m at v[j + 4m]
#pragma unroll
    for (int j = 0; j < 4; ++j) dft4<DIR>(v[j], v[j + 4], v[j + 8], v[j + 12]);
    // twiddle W16^(j m)
#pragma unroll
    for (int j = 1; j < 4; ++j)
#pragma unroll
        for (int m = 1; m < 4; ++m) v[j + 4 * m] = mul_w16<DIR>(v[j + 4 * m], j * m);
    // outer DFT4 over j for each m: output n at v[n + 4m]  (X[m + 4n])
#pragma unroll
    for (int m = 0; m < 4; ++m) dft4<DIR>(v[4 * m], v[4 * m + 1], v[4 * m + 2], v[4 * m + 3]);
}

template <int R> DDL_HD constexpr int xreg(int r) { return R == 16 ? reg16(r) : r; }

template <int R, int DIR> DDL_HD void xdft(cplx (&v)[R]) {
    if constexpr (R == 16) dft16<DIR>(v);
    else dftR<R, DIR>(v);
}

DDL_HD cplx csq(cplx a) { return mk(a.x * a.x - a.y * a.y, 2.0 * a.x * a.y); }

// v[idx(r)] *= w^r, r = 1..R-1; the powers come from two chains (odd / even) of products with
// w^2, so only three twiddles are live at a time
template <int R, bool OUT_ORDER>
DDL_HD void xtwiddle(cplx (&v)[R], cplx w1) {
    auto at = [&](int r) -> cplx& { return v[OUT_ORDER ? xreg<R>(r) : r]; };
    at(1) = cmul(at(1), w1);
    if constexpr (R > 2) {
        const cplx w2 = csq(w1);
        cplx wo = w1, we = w2;
        at(2) = cmul(at(2), we);
#pragma unroll
        for (int r = 3; r < R; r += 2) {
            wo = cmul(wo, w2);
            at(r) = cmul(at(r), wo);
            if (r + 1 < R) {
                we = cmul(we, w2);
                at(r + 1) = cmul(at(r + 1), we);
            }
        }
    }
}

// Stages whose sub-problems have length 2R (Q == 2) have only the twiddles w^r = exp(-2 pi i r / 2R)
// (butterflies with b == 1; b == 0 has none): compile-time constants, no generation, no load.
// Value (cos, sin) of exp(+2 pi i r / 2R); the forward sign conjugates.
template <int R> DDL_HD cplx xconst_twiddle(int r) {
    constexpr double c16 = 0.9238795325112867561282, s16 = 0.3826834323650897717285, h = 0.7071067811865475244008;
    constexpr double c32a = 0.9807852804032304491262, s32a = 0.1950903220161282678483;
    constexpr double c32b = 0.8314696123025452370788, s32b = 0.5555702330196022247428;
    if constexpr (R == 16) {          // exp(2 pi i r / 32)
        switch (r) {
            case 1: return mk(c32a, s32a);   case 2: return mk(c16, s16);     case 3: return mk(c32b, s32b);
            case 4: return mk(h, h);         case 5: return mk(s32b, c32b);   case 6: return mk(s16, c16);
            case 7: return mk(s32a, c32a);   case 8: return mk(0.0, 1.0);     case 9: return mk(-s32a, c32a);
            case 10: return mk(-s16, c16);   case 11: return mk(-s32b, c32b); case 12: return mk(-h, h);
            case 13: return mk(-c32b, s32b); case 14: return mk(-c16, s16);   case 15: return mk(-c32a, s32a);
            default: return mk(1.0, 0.0);
        }
    } else {                          // R == 8: exp(2 pi i r / 16)
        switch (r) {
            case 1: return mk(c16, s16);  case 2: return mk(h, h);    case 3: return mk(s16, c16);  case 4: return mk(0.0, 1.0);
            case 5: return mk(-s16, c16); case 6: return mk(-h, h);   case 7: return mk(-c16, s16);
            default: return mk(1.0, 0.0);
        }
    }
}

// v[idx(r)] *= exp(DIR * 2 pi i r / 2R) with constant operands
template <int R, int DIR, bool OUT_ORDER>
DDL_HD void xtwiddle_const(cplx (&v)[R]) {
#pragma unroll
    for (int r = 1; r < R; ++r) {
        cplx w = xconst_twiddle<R>(r);
        if (DIR < 0) w.y = -w.y;
        cplx& a = v[OUT_ORDER ? xreg<R>(r) : r];
        if (2 * r == R) a = mul_i<DIR>(a);            // w = +-i
        else a = cmul(a, w);
    }
}

#if DDL_DEVICE_BUILD
#define DDL_XF_ITEMS(i, count, NT) for (int i = threadIdx.x; i < (count); i += (NT))
#define DDL_XF_THREADS(t, NT) for (int t = threadIdx.x, _once = 1; _once; _once = 0)
#define DDL_LDG(p) __ldg(p)
#else
#define DDL_XF_ITEMS(i, count, NT) for (int i = 0; i < (count); ++i)
#define DDL_XF_THREADS(t, NT) for (int t = 0; t < (NT); ++t)
#define DDL_LDG(p) (*(p))
#endif

// barrier among the TP threads that own one pencil (emulation: phases run to completion in order)
template <int TP>
DDL_BODY void xgroup_sync(int group) {
#if DDL_DEVICE_BUILD
    if constexpr (TP <= 32) __syncwarp();
    else asm volatile("bar.sync %0, %1;" ::"r"(group + 1), "r"(TP) : "memory");
#else
    (void)group;
#endif
}

// Position of butterfly input/output j of an in-place stage: xsw(base + j*Q) = sb ^ xsw(j*Q)
// (the swizzle is linear over GF(2) and base, j*Q occupy disjoint bits), sb = xsw(base).
template <int N, int S_IDX, int DIR, bool DIT, int TP>
DDL_BODY void xstage(cplx* T, int lane, const cplx* __restrict__ tw) {
    constexpr int R = XFac<N>::radix(S_IDX);
    constexpr int P = xfac_P<N>(S_IDX);
    constexpr int M = N / P, Q = M / R;
    constexpr int ITEMS = N / R;
#pragma unroll
    for (int w0 = 0; w0 < ITEMS; w0 += TP) {
        const int w = w0 + lane;
        const int b = w / P, q = w % P;          // q fastest across lanes: conflict-free with the swizzle
        const int sb = xsw<N>(q * M + b);
        cplx v[R];
#pragma unroll
        for (int j = 0; j < R; ++j) v[j] = T[sb ^ xsw<N>(j * Q)];
        constexpr bool CONST_TW = (Q == 2) && (R == 16 || R == 8);
        if constexpr (DIT) {
            if (b != 0) {
                if constexpr (CONST_TW) xtwiddle_const<R, DIR, false>(v);
                else {
                    cplx w1 = DDL_LDG(&tw[b * P]);
                    if (DIR > 0) w1 = conj(w1);
                    xtwiddle<R, false>(v, w1);
                }
            }
            xdft<R, DIR>(v);
        } else {
            xdft<R, DIR>(v);
            if (b != 0) {
                if constexpr (CONST_TW) xtwiddle_const<R, DIR, true>(v);
                else {
                    cplx w1 = DDL_LDG(&tw[b * P]);
                    if (DIR > 0) w1 = conj(w1);
                    xtwiddle<R, true>(v, w1);
                }
            }
        }
#pragma unroll
        for (int r = 0; r < R; ++r) T[sb ^ xsw<N>(r * Q)] = v[xreg<R>(r)];
    }
}

// complex load that returns 0 when `ok` is false, as ONE predicated instruction (no branch region: the loads of a pencil
// stay independent of each other and of every predicate but their own)
DDL_BODY cplx xldg_if(const cplx* __restrict__ ptr, bool ok) {
#if DDL_DEVICE_BUILD
    cplx r;
    asm("{\n\t.reg .pred q;\n\tsetp.ne.s32 q, %3, 0;\n\tmov.f64 %0, 0d0000000000000000;\n\tmov.f64 %1, 0d0000000000000000;\n\t"
        "@q ld.global.nc.v2.f64 {%0, %1}, [%2];\n\t}"
        : "=d"(r.x), "=d"(r.y) : "l"(ptr), "r"((int)ok));
    return r;
#else
    return ok ? *ptr : mk(0.0, 0.0);
#endif
}

// Hermitian-packed first-stage inputs of one pencil, v[j] = Z[a + j Q0] with Z = A + iB of the two lines at pa, pa + pitch
// (pa already offset by a):  Z[e] = A[e] + i B[e] (e < kn),  Z[N-e] = conj(A[e]) + i conj(B[e]),  Z[0] = (Re A[0], Re B[0]),  else 0.
// Input j is element e = a + j Q0: e < N/2 for j < R0/2 (read index e), e >= N/2 otherwise (mirror, read index N - e >= 1; e == N/2
// reads nothing since kn <= N/2) - which of the two is known at compile time, so every input is one predicate and a constant offset
// from one of four pointers.
//   PACK 1: all loads from global memory into registers, predicated, none inside a branch region.
//   PACK 2: only the direct half is loaded (each retained mode once instead of twice); the mirrored half comes from the partner
//           lane through the pencil's own shared-memory slot T, which stage 0 overwrites afterwards.  Every thread of the
//           pencil's group must call (kn = 0 for a group without work): two group barriers inside.
template <int N, int R0, int PACK, int TP>
DDL_BODY void xpack_inputs(cplx (&v)[R0], const cplx* __restrict__ pa, long long pitch, int a, int kn, cplx* T, int group) {
    constexpr int Q0 = N / R0, H = R0 / 2;
    const cplx* __restrict__ pb = pa + pitch;
#if DDL_DEVICE_BUILD
    if constexpr (PACK == 2) {
        cplx za[H], zb[H];
#pragma unroll
        for (int j = 0; j < H; ++j) {
            const bool ok = a + j * Q0 < kn;
            za[j] = xldg_if(pa + j * Q0, ok);
            zb[j] = xldg_if(pb + j * Q0, ok);
        }
#pragma unroll
        for (int j = 0; j < H; ++j) {
            if (a + j * Q0 < kn) { T[a + j * Q0] = za[j]; T[N / 2 + a + j * Q0] = zb[j]; }
            if (j == 0 && a == 0) { za[j].y = -0.0; zb[j].y = 0.0; }      // exact: x - (+0) = x, (-0) + x = x
            v[j] = mk(za[j].x - zb[j].y, za[j].y + zb[j].x);
        }
        xgroup_sync<TP>(group);
#pragma unroll
        for (int j = H; j < R0; ++j) {
            const int idx = N - a - j * Q0;
            cplx ma = mk(0.0, 0.0), mb = mk(0.0, 0.0);
            if (idx < kn) { ma = T[idx]; mb = T[N / 2 + idx]; }
            v[j] = mk(ma.x + mb.y, mb.x - ma.y);
        }
        xgroup_sync<TP>(group);
        return;
    }
#endif
    (void)T; (void)group;
    const cplx* __restrict__ ma = pa + (N - 2 * a);
    const cplx* __restrict__ mb = pb + (N - 2 * a);
#pragma unroll
    for (int j = 0; j < R0; ++j) {
        const bool mir = j >= H;
        const bool ok = (mir ? N - a - j * Q0 : a + j * Q0) < kn;
        cplx za = xldg_if(mir ? ma - j * Q0 : pa + j * Q0, ok);
        cplx zb = xldg_if(mir ? mb - j * Q0 : pb + j * Q0, ok);
        if (j == 0 && a == 0) { za.y = -0.0; zb.y = 0.0; }
        v[j] = mir ? mk(za.x + zb.y, zb.x - za.y) : mk(za.x - zb.y, za.y + zb.x);
    }
}

// running maximum of non-negative values; NaN wins (numpy's max propagates it)
DDL_HD double xmax_nn(double m, double a) { return (a > m || a != a) ? a : m; }

// One CTA: G line pairs (lines 2*(bx*G+g), +1 of outer plane by), NT threads.
// CFL: also reduce max_{x,i} u_i(x)^2 and max_{x,i} B_i(x)^2 (T^2) over the grid points of this CTA into p.cfl -
// the real-space fields exist only here, so the time-step limit (physics.py:601-610,821-836) costs
// no transform of its own.
// KNC > 0: the retained-mode count is the compile-time constant KNC (the 2/3 rule: N/3 + 1) instead of p.kn, so the
// zero tests of the Hermitian pack and the bounds of the unpack fold at compile time (launch variant 3, below).
// PACK: how the first inverse stage gets its Hermitian-packed inputs.  0: one branch region per input (a memory round trip
// each once the compiler has scheduled them: nine per pencil in the SASS of <512, MHD3C>); 1, 2: xpack_inputs above, with the
// stage-0 twiddle fetched before the inputs instead of after the butterfly.
template <int N, class PHYS, int NT, int G, bool CFL = false, int KNC = 0, int PACK = 0>
DDL_BODY void xfused_block(const XFusedParams& p, cplx* tile, int bx, int by) {
    constexpr int NI = PHYS::NI, NO = PHYS::NO;
    constexpr int NS = NI > NO ? NI : NO;     // pencil slots per line pair
    constexpr int R0 = XFac<N>::radix(0);
    constexpr int Q0 = N / R0;
    constexpr int TP = Q0;                    // threads that own a pencil: one stage-0 butterfly each
    constexpr int GPR = NT / TP;              // pencils in flight per round
    static_assert(NT % TP == 0, "CTA size must be a multiple of the pencil group");
    constexpr int S = XFac<N>::S;
    const cplx* __restrict__ tw = p.tw;
    const int kn = KNC > 0 ? KNC : p.kn;
    const long long plane = (long long)by * p.s_outer;
    const int pair0 = bx * G;

    // PACK != 0: the stage-0 twiddle of a thread, tw[a], is the same for every pencil it touches in either direction.  It is
    // fetched once, together with the first inputs, and parked in TP entries behind the tile: every thread writes and later
    // reads entry a itself (the other groups write the same value there), so no barrier is involved and the three uses cost
    // a shared-memory load instead of a trip to L2 after each butterfly.
    cplx* wtab = tile + (size_t)G * NS * N;
    if constexpr (PACK != 0) {
        DDL_XF_THREADS(t, NT) wtab[t % TP] = DDL_LDG(&tw[t % TP]);
    }

    // ================= inverse: per pencil, stage 0 (global -> registers -> shared) then the
    //                   remaining outer stages, synchronised inside the pencil's own group
#pragma unroll 1
    for (int pen0 = 0; pen0 < G * NI; pen0 += GPR) {
        DDL_XF_THREADS(t, NT) {
            const int pen = pen0 + t / TP, a = t % TP;
            if constexpr (PACK != 0) {
                // every thread of a pencil group goes through xpack_inputs (PACK 2 synchronises the group inside); a group
                // without a pencil, or past the last line, loads nothing (kn = 0)
                const bool act = pen < G * NI;
                const int g = act ? pen / NI : 0, f = act ? pen % NI : 0;
                const int l0 = 2 * (pair0 + g);
                cplx* T = tile + (g * NS + f) * N;
                cplx v[R0];
                xpack_inputs<N, R0, PACK, TP>(v, p.in[f] + plane + (long long)l0 * p.pitch + a, p.pitch, a,
                                              (act && l0 < p.n_lines) ? kn : 0, T, t / TP);
                if (act) {
                    xdft<R0, +1>(v);
                    if (a != 0) xtwiddle<R0, true>(v, conj(wtab[a]));
                    const int sb = xsw<N>(a);
#pragma unroll
                    for (int r = 0; r < R0; ++r) T[sb ^ xsw<N>(r * Q0)] = v[xreg<R0>(r)];
                }
            } else if (pen < G * NI) {
                const int g = pen / NI, f = pen % NI;
                const int l0 = 2 * (pair0 + g);
                cplx v[R0];
                if (l0 < p.n_lines) {
                    // Hermitian-packed pencil Z = A + iB of the two lines:
                    //   Z[e] = A[e] + i B[e] (e < kn),  Z[N-e] = conj(A[e]) + i conj(B[e]),  else 0
                    const cplx* __restrict__ A = p.in[f] + plane + (long long)l0 * p.pitch;
                    const cplx* __restrict__ B = A + p.pitch;
#pragma unroll
                    for (int j = 0; j < R0; ++j) {
                        const int e = a + j * Q0;
                        cplx z = mk(0.0, 0.0);
                        if (e < kn) {
                            const cplx za = DDL_LDG(&A[e]), zb = DDL_LDG(&B[e]);
                            z = (e == 0) ? mk(za.x, zb.x) : mk(za.x - zb.y, za.y + zb.x);
                        } else if (N - e < kn) {
                            const cplx za = DDL_LDG(&A[N - e]), zb = DDL_LDG(&B[N - e]);
                            z = mk(za.x + zb.y, zb.x - za.y);
                        }
                        v[j] = z;
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < R0; ++j) v[j] = mk(0.0, 0.0);
                }
                xdft<R0, +1>(v);
                if (a != 0) xtwiddle<R0, true>(v, conj(DDL_LDG(&tw[a])));
                cplx* T = tile + (g * NS + f) * N;
                const int sb = xsw<N>(a);
#pragma unroll
                for (int r = 0; r < R0; ++r) T[sb ^ xsw<N>(r * Q0)] = v[xreg<R0>(r)];
            }
        }
        DDL_XF_THREADS(t, NT) xgroup_sync<TP>(t / TP);
        if constexpr (S >= 2) {
            DDL_XF_THREADS(t, NT) {
                const int pen = pen0 + t / TP;
                if (pen < G * NI) xstage<N, 1, +1, false, TP>(tile + ((pen / NI) * NS + pen % NI) * N, t % TP, tw);
            }
        }
        if constexpr (S >= 3) {
            DDL_XF_THREADS(t, NT) xgroup_sync<TP>(t / TP);
            DDL_XF_THREADS(t, NT) {
                const int pen = pen0 + t / TP;
                if (pen < G * NI) xstage<N, 2, +1, false, TP>(tile + ((pen / NI) * NS + pen % NI) * N, t % TP, tw);
            }
        }
        static_assert(S <= 3, "at most three outer stages");
    }
    DDL_SYNC();

    // ================= innermost radix-2 (inverse) -> products -> innermost radix-2 (forward);
    //                   one item = one real line (8-byte half) of one position pair
    {
        double* td = reinterpret_cast<double*>(tile);
        double m0 = 0.0, m1 = 0.0;
        DDL_XF_ITEMS(i, G * N, NT) {
            const int c = i & 1, wp = (i >> 1) % (N / 2), g = (i >> 1) / (N / 2);
            const int s0 = 2 * xsw<N>(2 * wp) + c, s1 = s0 ^ 2;     // xsw(2wp+1) = xsw(2wp) ^ 1
            double* T = td + (size_t)g * NS * N * 2;
            double u0[NI], u1[NI], o0[NO], o1[NO];
#pragma unroll
            for (int f = 0; f < NI; ++f) {
                const double x0 = T[f * 2 * N + s0], x1 = T[f * 2 * N + s1];
                u0[f] = x0 + x1; u1[f] = x0 - x1;
            }
            if constexpr (CFL) {
                double a0 = 0.0, a1 = 0.0, b0 = 0.0, b1 = 0.0;
#pragma unroll
                for (int f = 0; f < NI; ++f) {
                    if (f < PHYS::NDIM) { a0 = xmax_nn(a0, u0[f] * u0[f]); a1 = xmax_nn(a1, u1[f] * u1[f]); }
                    else if (f < PHYS::NDIM + PHYS::NG1) { b0 = xmax_nn(b0, u0[f] * u0[f]); b1 = xmax_nn(b1, u1[f] * u1[f]); }
                }
                m0 = xmax_nn(xmax_nn(m0, a0), a1);
                m1 = xmax_nn(xmax_nn(m1, b0), b1);
            }
            PHYS::apply(u0, o0, p.pc);
            PHYS::apply(u1, o1, p.pc);
#pragma unroll
            for (int f = 0; f < NO; ++f) {
                T[f * 2 * N + s0] = o0[f] + o1[f];
                T[f * 2 * N + s1] = o0[f] - o1[f];
            }
        }
        if constexpr (CFL) {
#if DDL_DEVICE_BUILD
            // every thread of the CTA is here (the item loop is uniform in its exit): full-warp tree,
            // then one atomic per warp; non-negative doubles order like their bit patterns
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
                m0 = xmax_nn(m0, __shfl_xor_sync(0xffffffffu, m0, off));
                m1 = xmax_nn(m1, __shfl_xor_sync(0xffffffffu, m1, off));
            }
            if ((threadIdx.x & 31) == 0) {
                atomicMax(reinterpret_cast<unsigned long long*>(p.cfl), (unsigned long long)__double_as_longlong(m0));
                atomicMax(reinterpret_cast<unsigned long long*>(p.cfl) + 1, (unsigned long long)__double_as_longlong(m1));
            }
#else
            p.cfl[0] = xmax_nn(p.cfl[0], m0);
            p.cfl[1] = xmax_nn(p.cfl[1], m1);
#endif
        }
    }
    DDL_SYNC();

    // ================= forward: per pencil, outer stages S-1..1 (shared), then stage 0
    //                   (shared -> registers), Hermitian unpack and store of the retained modes
    const double h = 0.5 * p.scale;
#pragma unroll 1
    for (int pen0 = 0; pen0 < G * NO; pen0 += GPR) {
        if constexpr (S >= 3) {
            DDL_XF_THREADS(t, NT) {
                const int pen = pen0 + t / TP;
                if (pen < G * NO) xstage<N, 2, -1, true, TP>(tile + ((pen / NO) * NS + pen % NO) * N, t % TP, tw);
            }
            DDL_XF_THREADS(t, NT) xgroup_sync<TP>(t / TP);
        }
        if constexpr (S >= 2) {
            DDL_XF_THREADS(t, NT) {
                const int pen = pen0 + t / TP;
                if (pen < G * NO) xstage<N, 1, -1, true, TP>(tile + ((pen / NO) * NS + pen % NO) * N, t % TP, tw);
            }
            DDL_XF_THREADS(t, NT) xgroup_sync<TP>(t / TP);
        }
#if DDL_DEVICE_BUILD
        constexpr bool SHUFFLE_UNPACK = (TP <= 32);
#else
        constexpr bool SHUFFLE_UNPACK = false;
#endif
        DDL_XF_THREADS(t, NT) {
            const int pen = pen0 + t / TP, a = t % TP;
            const bool act = pen < G * NO;
            const int g = act ? pen / NO : 0, f = act ? pen % NO : 0;
            cplx* T = tile + (g * NS + f) * N;
            const int sb = xsw<N>(a);
            cplx v[R0];
            if (act) {
#pragma unroll
                for (int j = 0; j < R0; ++j) v[j] = T[sb ^ xsw<N>(j * Q0)];
                if (a != 0) xtwiddle<R0, false>(v, PACK != 0 ? wtab[a] : DDL_LDG(&tw[a]));
                xdft<R0, -1>(v);
            } else {
#pragma unroll
                for (int j = 0; j < R0; ++j) v[j] = mk(0.0, 0.0);
            }
            if constexpr (SHUFFLE_UNPACK) {
#if DDL_DEVICE_BUILD
                // Z[k], k = a + r Q0, is output r of lane a; its mirror Z[N-k] is output R0-1-r of
                // lane Q0-a of the same group (lane 0: own output R0-r).  A[k] = (Z[k] + conj Z[N-k])/2,
                // B[k] = (Z[k] - conj Z[N-k]) / 2i.
                const int l0 = 2 * (pair0 + g);
                const bool st = act && l0 < p.n_lines;
                cplx* __restrict__ dst = p.out[f] + plane + (long long)l0 * p.pitch;
                const int src = (threadIdx.x & ~(TP - 1) & 31) | ((TP - a) & (TP - 1));
                const int rmax = (kn - 1) / Q0;          // uniform
#pragma unroll
                for (int r = 0; r < R0 / 2; ++r) {
                    if (r <= rmax) {
                        const cplx Zk = v[xreg<R0>(r)];
                        const cplx mine = v[xreg<R0>(R0 - 1 - r)];
                        cplx Zm;
                        Zm.x = __shfl_sync(0xffffffffu, mine.x, src);
                        Zm.y = __shfl_sync(0xffffffffu, mine.y, src);
                        if (a == 0) Zm = v[xreg<R0>((R0 - r) % R0)];
                        const int k = a + r * Q0;
                        if (st && k < kn) {
                            dst[k] = mk((Zk.x + Zm.x) * h, (Zk.y - Zm.y) * h);
                            dst[p.pitch + k] = mk((Zk.y + Zm.y) * h, (Zm.x - Zk.x) * h);
                        }
                    }
                }
#endif
            } else {
                if (act) {
#pragma unroll
                    for (int r = 0; r < R0; ++r) T[sb ^ xsw<N>(r * Q0)] = v[xreg<R0>(r)];
                }
            }
        }
        if constexpr (!SHUFFLE_UNPACK) {
            DDL_XF_THREADS(t, NT) xgroup_sync<TP>(t / TP);
            DDL_XF_THREADS(t, NT) {
                const int pen = pen0 + t / TP, a = t % TP;
                if (pen < G * NO) {
                    const int g = pen / NO, f = pen % NO;
                    const int l0 = 2 * (pair0 + g);
                    if (l0 < p.n_lines) {
                        const cplx* T = tile + (g * NS + f) * N;
                        cplx* __restrict__ dst = p.out[f] + plane + (long long)l0 * p.pitch;
                        for (int k = a; k < kn; k += TP) {
                            const cplx Zk = T[xsw<N>(k)], Zm = T[xsw<N>((N - k) % N)];
                            dst[k] = mk((Zk.x + Zm.x) * h, (Zk.y - Zm.y) * h);
                            dst[p.pitch + k] = mk((Zk.y + Zm.y) * h, (Zm.x - Zk.x) * h);
                        }
                    }
                }
            }
            // the next round reuses no slot of this one: no barrier needed here
        }
    }
}

// Launch variants (ddl_set_option("xfused_variant", v); the default is the measured best):
//   0: 6 warps per CTA, 3 CTAs per SM   1: 9 warps per CTA, 2 CTAs per SM   2: 6 warps, 2 CTAs (more registers)
//   3: shape 0 with the retained-mode count of the 2/3 rule (N/3 + 1) as a compile-time constant; falls back to 0 for
//      any other mask.  Measured (profiles/r2/devcheck_b200_xfused_variants.txt): 4.30 ms against 4.12 ms for variant 0 although it has
//      336 fewer static instructions (cuobjdump: 2 296 -> 1 960 for <512, MHD3C>): opt-in.
//   4: persistent CTAs, next pair staged by cp.async.bulk + mbarrier (xfused_persist.cuh): 4.92 ms, opt-in
//   5: persistent CTAs of shape 0, no staging; warps without a second-round pencil start the next pair (xfused_rot.cuh)
//   6, 8: shape 0 with the branch-free input packs of xpack_inputs (registers only / mirrored half through shared memory):
//      one to three memory round trips per pencil instead of nine, 4.26 / 4.35 ms -- the pass is bound by the shared-memory /
//      L1 data pipe (70 % busy), not by that latency: opt-in
template <int N, class PHYS, int V> struct XFusedCfg {
    static constexpr int NS = PHYS::NI > PHYS::NO ? PHYS::NI : PHYS::NO;
    static constexpr int G = (N >= 512) ? 1 : 512 / N;
    static constexpr int TP = N / XFac<N>::radix(0);
    static constexpr int NT = TP > 32 ? 9 * TP : (V == 1 ? 288 : 192);
    static constexpr int PACK = V == 6 ? 1 : (V == 8 ? 2 : 0);
    static constexpr size_t SMEM = ((size_t)G * NS * N + (PACK != 0 ? TP : 0)) * sizeof(cplx);      // PACK != 0: + the parked stage-0 twiddles
    static constexpr int WANT = (V == 0 || V >= 6) ? 3 : 2;
    static constexpr int MINB = (SMEM * WANT <= 222 * 1024 && NT * WANT <= 1024) ? WANT : ((SMEM * 2 <= 222 * 1024 && NT * 2 <= 1024) ? 2 : 1);
    // register budget that still lets MINB CTAs share an SM: each of the four sub-partitions owns
    // 16 K registers and holds ceil(warps / 4) of the resident warps (allocation unit: 8 per thread)
    static constexpr int WPS = (NT / 32 * MINB + 3) / 4;
    static constexpr int MAXREG = (16384 / (WPS * 32) / 8 * 8) > 255 ? 255 : (16384 / (WPS * 32) / 8 * 8);
};

#if DDL_DEVICE_BUILD
template <int N, class PHYS, int V, bool CFL, int KNC = 0>
__global__ void __maxnreg__((XFusedCfg<N, PHYS, V>::MAXREG))
xfused_kernel(const __grid_constant__ XFusedParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    xfused_block<N, PHYS, XFusedCfg<N, PHYS, V>::NT, XFusedCfg<N, PHYS, V>::G, CFL, KNC, XFusedCfg<N, PHYS, V>::PACK>(p, reinterpret_cast<cplx*>(smem_raw), blockIdx.x, blockIdx.y);
}
#endif

// returns 0 on success, negative on error
template <int N, class PHYS, int V, bool CFL = false, int KNC = 0>
int launch_xfused_v(const XFusedParams& p, int n_outer, ddl_stream_t stream) {
    using Cfg = XFusedCfg<N, PHYS, V>;
    const int pairs = (p.n_lines + 1) / 2;
    const int gx = (pairs + Cfg::G - 1) / Cfg::G;
#if DDL_DEVICE_BUILD
    auto kern = xfused_kernel<N, PHYS, V, CFL, KNC>;
    static DeviceOnce once;
    if (once.get([&]() -> int {
            DDL_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM));
            DDL_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
            return 1;
        }) < 0) return -2;
    dim3 grid(gx, n_outer, 1);
    prof_begin("x_fused", stream);
    kern<<<grid, Cfg::NT, Cfg::SMEM, stream>>>(p);
    prof_end(stream);
    DDL_CUDA_CHECK(cudaGetLastError());
#else
    prof_begin("x_fused", stream);
    cplx* tile = (cplx*)malloc(Cfg::SMEM);
    for (int by = 0; by < n_outer; ++by)
        for (int bx = 0; bx < gx; ++bx) xfused_block<N, PHYS, Cfg::NT, Cfg::G, CFL, KNC, Cfg::PACK>(p, tile, bx, by);
    free(tile);
#endif
    return 0;
}

#if DDL_DEVICE_BUILD
template <int N, class PHYS, bool CFL> int launch_xfused_persist(const XFusedParams& p, int n_outer, ddl_stream_t stream);   // xfused_persist.cuh
template <int N, class PHYS, bool CFL> int launch_xfused_rot(const XFusedParams& p, int n_outer, ddl_stream_t stream);       // xfused_rot.cuh
#endif

template <int N, class PHYS>
int launch_xfused(const XFusedParams& p, int n_outer, int variant, ddl_stream_t stream) {
#if DDL_DEVICE_BUILD
    if (variant == 4) {
        // persistent CTAs with the next pair's lines staged by the bulk-copy engine (xfused_persist.cuh); 1 = not applicable
        const int rc = p.cfl ? launch_xfused_persist<N, PHYS, true>(p, n_outer, stream) : launch_xfused_persist<N, PHYS, false>(p, n_outer, stream);
        if (rc != 1) return rc;
    }
    if (variant == 5) {
        // persistent CTAs of the default shape; the warps without a second-round pencil start the next pair (xfused_rot.cuh)
        const int rc = p.cfl ? launch_xfused_rot<N, PHYS, true>(p, n_outer, stream) : launch_xfused_rot<N, PHYS, false>(p, n_outer, stream);
        if (rc != 1) return rc;
    }
#endif
    if (p.cfl) return launch_xfused_v<N, PHYS, 0, true>(p, n_outer, stream);     // capture: default CTA shape only
    if (variant == 3 && p.kn == N / 3 + 1) return launch_xfused_v<N, PHYS, 0, false, N / 3 + 1>(p, n_outer, stream);
    // 6: branch-free register pack; 8: direct half from global memory, mirrored half through the pencil's shared-memory slot
    // (xpack_inputs).  Both were also measured with the retained count at compile time (4.25 / 4.46 ms) and 8 in the CTA shapes
    // 3 x 3, 4 x 3 and 9 x 2 warps (4.26 / 4.34 / 4.80 ms; the original pack in 3 x 3 and 4 x 3: 6.23 / 5.98 ms) -- those
    // instantiations were removed after the measurement (profiles/r2/devcheck_b200_xpass_pack.txt, ..._shapes.txt)
    if (variant == 6) return launch_xfused_v<N, PHYS, 6>(p, n_outer, stream);
    if (variant == 8) return launch_xfused_v<N, PHYS, 8>(p, n_outer, stream);
#if DDL_DEVICE_BUILD
    if (variant == 1) return launch_xfused_v<N, PHYS, 1>(p, n_outer, stream);
    if (variant == 2) return launch_xfused_v<N, PHYS, 2>(p, n_outer, stream);
#else
    (void)variant;
#endif
    return launch_xfused_v<N, PHYS, 0>(p, n_outer, stream);
}

// default CTA shape only (plus its CFL-capture twin): for the advective-form policies, which are off the measured path
template <int N, class PHYS>
int launch_xfused_basic(const XFusedParams& p, int n_outer, ddl_stream_t stream) {
    if (p.cfl) return launch_xfused_v<N, PHYS, 0, true>(p, n_outer, stream);
    return launch_xfused_v<N, PHYS, 0>(p, n_outer, stream);
}

}  // namespace ddl
