// Two-stage strided pass (y / z axes) for plain rows: N = R0 x R1 with both factors done as register butterflies, so a pencil
// crosses shared memory ONCE (one write, one read, one barrier) instead of twice, and every thread keeps 32 independent 16-byte
// loads in flight.  Same mathematics as strided_fast (fast_kernels.cuh) and the generic tile kernel: a decimation-in-frequency
// split  n = a + R1 j,  k = r + R0 s:
//     y_a[r] = w_N^(a r) * sum_j x[a + R1 j] w_R0^(j r)          (stage 0: global -> registers -> shared)
//     X[r + R0 s] = sum_a y_a[r] w_R1^(a s)                       (stage 1: shared -> registers -> global)
// Outputs land on their natural rows directly (k is computed, nothing is scrambled); rows the 2/3 rule prunes are predicated by
// formula on either side, never read or written.  A thread owns one pencil column c of CX for all its butterflies: N/R1/TPP
// stage-0 butterflies (TPP = N/R1 threads per pencil) and one stage-1 butterfly.  The 32-point butterfly is two 16-point ones plus
// constant twiddles (decimation in time inside the registers).
//
// Written as a block body like the x pass (phases over the threads of a CTA), so that the host emulation runs the same index
// logic and arithmetic (tests only); the slab extensions (peer-blocked rows, row tables, peer stores) stay with strided_fast.
// Replaces, like strided_fast, the y and z halves of the reference's FFTW plans (dedalus/utils/fftw/_fftw.pyx:246-309) with the
// dealias mask of dedalus/data_objects/dealias_cy_3d.pyx:13-46 fused into the row predicates.
#pragma once
#include "ddl_common.cuh"
#include "xfused_kernel.cuh"

namespace ddl {

template <int N> struct TwoFac { static constexpr bool ok = false; static constexpr int R0 = 1, R1 = 1; };
template <> struct TwoFac<256> { static constexpr bool ok = true; static constexpr int R0 = 16, R1 = 16; };
template <> struct TwoFac<512> { static constexpr bool ok = true; static constexpr int R0 = 16, R1 = 32; };

struct TwoSide {
    long long s_n, s_outer;   // element strides of the transform axis and of the outer index
    const int* outer_tab;     // outer index -> stored outer index (NULL = identity)
    int m;                    // retained rows |index| <= m, or -1 = every row present
    int compact;              // retained rows stored contiguously (workspace) instead of in place (state)
};

struct TwoParams {
    const cplx* in[DDL_MAXF];
    cplx* out[DDL_MAXF];
    TwoSide si, so;
    int inner_len;
    double scale;
    const cplx* tw;           // exp(-2 pi i m / N)
};

// stored row of logical row r and whether the row exists (fast_kernels.cuh fast_row without branches: the row of a butterfly
// input depends on the thread, and a branch per row would split the warp 32 times per pass)
template <int N> DDL_HD int two_row(int r, int m, int compact, bool& keep) {
    const bool hi = r >= N - m;
    keep = (m < 0) | (r <= m) | hi;
    return r - ((hi & (m >= 0) & (compact != 0)) ? (N - 2 * m - 1) : 0);
}

// complex store that happens only when `ok`, as ONE predicated instruction (see xldg_if)
DDL_BODY void xstg_if(cplx* ptr, cplx v, bool ok) {
#if DDL_DEVICE_BUILD
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.s32 q, %3, 0;\n\t@q st.global.v2.f64 [%0], {%1, %2};\n\t}"
                 :: "l"(ptr), "d"(v.x), "d"(v.y), "r"((int)ok) : "memory");
#else
    if (ok) *ptr = v;
#endif
}

// in-place 32-point DFT, v natural in; output X[s] is left in v[two_reg32(s)]:
// even / odd halves by dft16 (outputs at reg16 positions), odd half times exp(DIR 2 pi i t / 32), radix-2 combination
DDL_HD constexpr int two_reg32(int s) { return s < 16 ? reg16(s) : 16 + reg16(s - 16); }

template <int DIR> DDL_HD void dft32(cplx (&v)[32]) {
    cplx e[16], o[16];
#pragma unroll
    for (int m = 0; m < 16; ++m) { e[m] = v[2 * m]; o[m] = v[2 * m + 1]; }
    dft16<DIR>(e);
    dft16<DIR>(o);
    xtwiddle_const<16, DIR, true>(o);
#pragma unroll
    for (int t = 0; t < 16; ++t) {
        const cplx E = e[reg16(t)], O = o[reg16(t)];
        v[reg16(t)] = E + O;
        v[16 + reg16(t)] = E - O;
    }
}

template <int R, int DIR> DDL_HD void two_dft(cplx (&v)[R]) {
    if constexpr (R == 32) dft32<DIR>(v);
    else dft16<DIR>(v);
}
template <int R> DDL_HD constexpr int two_reg(int s) { return R == 32 ? two_reg32(s) : reg16(s); }

// one CTA: CX adjacent pencils (inner indices bx*CX ...) of outer index by, field bz; NT = CX * N / R1 threads
// TWL: the 15 stage-0 twiddles w^(a r) of a butterfly are loaded from the table (true) or generated in registers as powers of
// w^a by two product chains (false, like the x pass)
template <int N, int DIR, int CX, bool TWL>
DDL_BODY void strided_two_block(const TwoParams& p, cplx* tile, int bx, int by, int bz) {
    constexpr int R0 = TwoFac<N>::R0, R1 = TwoFac<N>::R1;
    static_assert(R0 == 16 && R0 * R1 == N, "stage 0 is a 16-point butterfly");
    constexpr int TPP = N / R1;               // threads per pencil = stage-1 butterflies per pencil (= R0)
    constexpr int NB0 = R1 / TPP;             // stage-0 butterflies per thread (a = a0 + TPP h)
    constexpr int NT = CX * TPP;
    const cplx* __restrict__ tw = p.tw;
    const cplx* __restrict__ in = p.in[bz];
    cplx* __restrict__ out = p.out[bz];
    const unsigned sn_i = (unsigned)p.si.s_n, sn_o = (unsigned)p.so.s_n;      // N * s_n < 2^31 elements (checked by the launcher)
    const long long ib0 = (long long)(p.si.outer_tab ? p.si.outer_tab[by] : by) * p.si.s_outer;
    const long long ob0 = (long long)(p.so.outer_tab ? p.so.outer_tab[by] : by) * p.so.s_outer;

    // ---- stage 0: rows a + R1 j (j = 0..15) of butterfly a, for a = a0, a0 + TPP, ...: all loads first, then the butterflies
    DDL_XF_THREADS(t, NT) {
        const int c = t % CX, a0 = t / CX;
        const int inner = bx * CX + c;
        const bool live = inner < p.inner_len;
        const cplx* __restrict__ src = in + ib0 + inner;
        cplx v[NB0][R0];
#pragma unroll
        for (int h = 0; h < NB0; ++h)
#pragma unroll
            for (int j = 0; j < R0; ++j) {
                bool keep;
                const int row = two_row<N>(a0 + TPP * h + R1 * j, p.si.m, p.si.compact, keep);
                v[h][j] = xldg_if(&src[(unsigned)row * sn_i], live && keep);
            }
#pragma unroll
        for (int h = 0; h < NB0; ++h) {
            const int a = a0 + TPP * h;
            dft16<DIR>(v[h]);
            if (a != 0) {
                if constexpr (TWL) {
#pragma unroll
                    for (int r = 1; r < R0; ++r) {
                        cplx w = DDL_LDG(&tw[a * r]);
                        if (DIR > 0) w = conj(w);
                        v[h][reg16(r)] = cmul(v[h][reg16(r)], w);
                    }
                } else {
                    cplx w1 = DDL_LDG(&tw[a]);
                    if (DIR > 0) w1 = conj(w1);
                    xtwiddle<R0, true>(v[h], w1);
                }
            }
#pragma unroll
            for (int r = 0; r < R0; ++r) tile[(r * R1 + a) * CX + c] = v[h][reg16(r)];
        }
    }
    DDL_SYNC();
    // ---- stage 1: butterfly r over a = 0..R1-1; output rows r + R0 s
    DDL_XF_THREADS(t, NT) {
        const int c = t % CX, r = t / CX;
        const int inner = bx * CX + c;
        const bool live = inner < p.inner_len;
        cplx* __restrict__ dst = out + ob0 + inner;
        cplx u[R1];
#pragma unroll
        for (int a = 0; a < R1; ++a) u[a] = tile[(r * R1 + a) * CX + c];
        two_dft<R1, DIR>(u);
        const double sc = p.scale;
#pragma unroll
        for (int s = 0; s < R1; ++s) {
            bool keep;
            const int row = two_row<N>(r + R0 * s, p.so.m, p.so.compact, keep);
            xstg_if(&dst[(unsigned)row * sn_o], scal(u[two_reg<R1>(s)], sc), live && keep);
        }
    }
}

template <int N> struct TwoCfg {
    static constexpr int CX = 8;                                   // 128-byte row segments
    static constexpr int NT = CX * (N / TwoFac<N>::R1);
    static constexpr size_t SMEM = (size_t)N * CX * sizeof(cplx);
    static constexpr int MINB = (SMEM * 4 <= 222 * 1024) ? 4 : ((SMEM * 3 <= 222 * 1024) ? 3 : 2);      // 512: 3 x 64 KB, 168 registers; 256: 4 x 32 KB, 128 registers
};

#if DDL_DEVICE_BUILD
template <int N, int DIR, bool TWL>
__global__ void __launch_bounds__(TwoCfg<N>::NT, TwoCfg<N>::MINB)
strided_two(const __grid_constant__ TwoParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    strided_two_block<N, DIR, TwoCfg<N>::CX, TWL>(p, reinterpret_cast<cplx*>(smem_raw), blockIdx.x, blockIdx.y, blockIdx.z);
}
#endif

// returns 0 on success, negative on error
template <int N, int DIR, bool TWL>
int launch_strided_two(const TwoParams& p, int nf, int n_outer, const char* name, ddl_stream_t stream) {
    using Cfg = TwoCfg<N>;
    const int gx = (p.inner_len + Cfg::CX - 1) / Cfg::CX;
    if (p.si.s_n <= 0 || p.so.s_n <= 0 || p.si.s_n * N >= (1LL << 31) || p.so.s_n * N >= (1LL << 31)) return 1;      // not this kernel's case
#if DDL_DEVICE_BUILD
    auto kern = strided_two<N, DIR, TWL>;
    static DeviceOnce once;
    if (once.get([&]() -> int {
            DDL_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM));
            return 1;
        }) < 0) return -2;
    dim3 grid(gx, n_outer, nf);
    prof_begin(name, stream);
    kern<<<grid, Cfg::NT, Cfg::SMEM, stream>>>(p);
    prof_end(stream);
    DDL_CUDA_CHECK(cudaGetLastError());
#else
    prof_begin(name, stream);
    cplx* tile = (cplx*)malloc(Cfg::SMEM);
    for (int bz = 0; bz < nf; ++bz)
        for (int by = 0; by < n_outer; ++by)
            for (int bx = 0; bx < gx; ++bx) strided_two_block<N, DIR, Cfg::CX, TWL>(p, tile, bx, by, bz);
    free(tile);
#endif
    return 0;
}

}  // namespace ddl
