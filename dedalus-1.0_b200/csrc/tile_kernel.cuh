// Generic tiled pencil-FFT kernel (all axis passes of the 2-D / 3-D transforms and the fused
// "inverse -> real-space products -> forward" middle pass).
//
// One CTA owns a tile of pencils along the transform axis, staged in shared memory as
// tile[row*ld + c].  Four modes:
//   TM_C2C   : complex pass along a strided or contiguous axis (pruned rows via index tables)
//   TM_C2R   : half-complex spectra of a PAIR of adjacent real lines packed into one complex
//              pencil (z = a + i b), inverse transform, store the two real lines
//   TM_R2C   : the mirror (forward)
//   TM_FUSED : Hermitian pack -> inverse (DIF) -> PHYS::apply at every grid point ->
//              forward (DIT) -> unpack; the real-space fields never touch HBM
// Replaces, per transform, fftw rPlan execute + "/= N" + dealias passes of the reference
// (dedalus/data_objects/representations.py:318-357, dealias_cy_3d.pyx:13-46).
#pragma once
#include <cmath>
#include "ddl_common.cuh"
#include "physics_ops.cuh"

namespace ddl {

enum TileMode { TM_C2C = 0, TM_C2R = 1, TM_R2C = 2, TM_FUSED = 3 };

struct TileSide {
    long long s_n, s_inner, s_outer;   // element strides: transform axis, pencil (inner) index, outer index
    const int* n_tab;                  // spectral side: logical row -> stored row, -1 = absent (zero / skip); NULL = identity
    const int* outer_tab;              // outer index -> stored outer index; NULL = identity
    // slab exchange buffers: rows come in `split` consecutive rows per peer block (0 = plain):
    // stored row r sits at (r / split) * s_blk + (r % split) * s_n
    int split;
    long long s_blk;
};

DDL_HD long long row_off(const TileSide& s, int r) {
    return s.split ? (long long)(r / s.split) * s.s_blk + (long long)(r % s.split) * s.s_n : (long long)r * s.s_n;
}

struct TileParams {
    const void* in[DDL_MAXF];
    void* out[DDL_MAXF];
    TileSide si, so;
    int nf_in, nf_out;    // fields read / written (TM_C2C: one field per blockIdx.z)
    int nft;              // pencil slots per group inside the tile (>= max(nf_in, nf_out))
    int G;                // TM_C2C: pencils per tile; pair modes: line PAIRS per tile
    int inner_len;        // TM_C2C: number of pencils; pair modes: number of real lines
    int n_outer;
    int kn;               // pair modes: stored non-negative modes along the transform axis
    int ld;               // tile leading dimension
    double scale;
    const cplx* tw;
    PhysConst pc;
    const char* name;     // label for launch accounting / profiling (host side only)
    RtFac rt;             // N == 0 instantiation (runtime length): the factorisation of the transform length
    // shearing box (representations.py:558-740), pair modes of the N == 0 instantiation only: the spectrum of the line
    // at y = l * sh_dy is multiplied by exp(-+ i ((S kx) y) t) between the x pass and the y pass (rev_np / fwd_np :700-740)
    int sh_on;
    double sh_S, sh_t, sh_dy;
    const double* sh_kx;  // kx value per stored non-negative mode
    double* cfl;          // TM_FUSED: optional CFL capture, cfl[0] = max(cfl[0], max_{x,i} u_i(x)^2), cfl[1]: second group
                          // (B or T) -- fields.py:153-157 max_square without a transform of its own; NULL = off
};

template <int N, int S_IDX, int DIR, bool DIT>
DDL_BODY void stage_all(cplx* tile, int ld, int nfa, int nft, int G, const cplx* __restrict__ tw) {
    constexpr int R = Fac<N>::radix(S_IDX);
    const int items = nfa * G * (N / R);
    DDL_FOR_ITEMS(i, items) {
        const int f = i % nfa, t = i / nfa;
        const int g = t % G, w = t / G;
        stage_item<N, S_IDX, DIR, DIT>(tile, ld, g * nft + f, w, tw);
    }
}

// natural in -> scrambled out; the caller has synchronised after filling the tile
template <int N, int DIR, int S_IDX = 0>
DDL_BODY void tile_fft_dif(cplx* tile, int ld, int nfa, int nft, int G, const cplx* __restrict__ tw) {
    stage_all<N, S_IDX, DIR, false>(tile, ld, nfa, nft, G, tw);
    DDL_SYNC();
    if constexpr (S_IDX + 1 < Fac<N>::S) tile_fft_dif<N, DIR, S_IDX + 1>(tile, ld, nfa, nft, G, tw);
}

// scrambled in -> natural out
template <int N, int DIR, int S_IDX = Fac<N>::S - 1>
DDL_BODY void tile_fft_dit(cplx* tile, int ld, int nfa, int nft, int G, const cplx* __restrict__ tw) {
    stage_all<N, S_IDX, DIR, true>(tile, ld, nfa, nft, G, tw);
    DDL_SYNC();
    if constexpr (S_IDX > 0) tile_fft_dit<N, DIR, S_IDX - 1>(tile, ld, nfa, nft, G, tw);
}

// ---- runtime-length variants (template length 0): same stage structure, radices from p.rt (fft_core.cuh)
template <int DIR, bool DIT>
DDL_BODY void stage_all_rt(cplx* tile, int ld, int nfa, int nft, int G, const cplx* __restrict__ tw, const RtFac& rt, int s) {
    int P = 1;
    for (int t = 0; t < s; ++t) P *= rt.radix[t];
    const int R = rt.radix[s], M = rt.n / P, Q = M / R;
    const int items = nfa * G * (rt.n / R);
    DDL_FOR_ITEMS(i, items) {
        const int f = i % nfa, t = i / nfa;
        const int g = t % G, w = t / G;
        stage_item_rt<DIR, DIT>(tile, ld, g * nft + f, w, rt.n, R, M, Q, P, tw);
    }
}

template <int N, int DIR>
DDL_BODY void fft_dif(const RtFac& rt, cplx* tile, int ld, int nfa, int nft, int G, const cplx* __restrict__ tw) {
    if constexpr (N > 0) {
        tile_fft_dif<N, DIR>(tile, ld, nfa, nft, G, tw);
    } else {
        for (int s = 0; s < rt.S; ++s) { stage_all_rt<DIR, false>(tile, ld, nfa, nft, G, tw, rt, s); DDL_SYNC(); }
    }
}
template <int N, int DIR>
DDL_BODY void fft_dit(const RtFac& rt, cplx* tile, int ld, int nfa, int nft, int G, const cplx* __restrict__ tw) {
    if constexpr (N > 0) {
        tile_fft_dit<N, DIR>(tile, ld, nfa, nft, G, tw);
    } else {
        for (int s = rt.S - 1; s >= 0; --s) { stage_all_rt<DIR, true>(tile, ld, nfa, nft, G, tw, rt, s); DDL_SYNC(); }
    }
}
template <int N> DDL_HD int tpos(const RtFac& rt, int k) {
    if constexpr (N > 0) return pos_of_index<N>(k); else return pos_of_index_rt(rt, k);
}
template <int N> DDL_HD int tidx(const RtFac& rt, int pos) {
    if constexpr (N > 0) return index_of_pos<N>(pos); else return index_of_pos_rt(rt, pos);
}

// running maximum of non-negative values; NaN wins (numpy's max propagates it)
DDL_HD double tmax_nn(double m, double a) { return (a > m || a != a) ? a : m; }

DDL_HD long long outer_off(const TileSide& s, int o) {
    return (long long)(s.outer_tab ? s.outer_tab[o] : o) * s.s_outer;
}

// exp(sign * i * ((S kx) y) t), the argument rounded in the reference's order (representations.py:607-611,674,692)
DDL_HD cplx shear_phase(const TileParams& p, int k, int line, double sign) {
    const double arg = ((p.sh_S * p.sh_kx[k]) * ((double)line * p.sh_dy)) * p.sh_t;
    double s, c;
#if DDL_DEVICE_BUILD
    sincos(arg, &s, &c);
#else
    s = std::sin(arg); c = std::cos(arg);
#endif
    return mk(c, sign * s);
}

template <int N, int MODE, int DIR, class PHYS>
DDL_BODY void tile_block(const TileParams& p, cplx* tile, int bx, int by, int bz) {
    const int ld = p.ld;
    const int n_len = (N > 0) ? N : p.rt.n;      // compile-time length, or the runtime one of the N == 0 instantiation
    const cplx* __restrict__ tw = p.tw;

    if constexpr (MODE == TM_C2C) {
        const cplx* __restrict__ in = (const cplx*)p.in[bz];
        cplx* __restrict__ out = (cplx*)p.out[bz];
        const int i0 = bx * p.G;
        const int npc = (p.inner_len - i0) < p.G ? (p.inner_len - i0) : p.G;
        const long long ib = outer_off(p.si, by) + (long long)i0 * p.si.s_inner;
        const long long ob = outer_off(p.so, by) + (long long)i0 * p.so.s_inner;
        const bool in_nfast = (p.si.s_n == 1);
        DDL_FOR_ITEMS(i, n_len * npc) {
            int n, c;
            if (in_nfast) { n = i % n_len; c = i / n_len; } else { c = i % npc; n = i / npc; }
            const int pn = p.si.n_tab ? p.si.n_tab[n] : n;
            cplx v = mk(0.0, 0.0);
            if (pn >= 0) v = in[ib + row_off(p.si, pn) + (long long)c * p.si.s_inner];
            tile[n * ld + c] = v;
        }
        DDL_SYNC();
        fft_dif<N, DIR>(p.rt, tile, ld, 1, 1, npc, tw);
        const bool out_nfast = (p.so.s_n == 1);
        const double sc = p.scale;
        DDL_FOR_ITEMS(i, n_len * npc) {
            int k, c, pos;
            if (out_nfast) { k = i % n_len; c = i / n_len; pos = tpos<N>(p.rt, k); }
            else { c = i % npc; pos = i / npc; k = tidx<N>(p.rt, pos); }
            const int pk = p.so.n_tab ? p.so.n_tab[k] : k;
            if (pk >= 0) out[ob + row_off(p.so, pk) + (long long)c * p.so.s_inner] = scal(tile[pos * ld + c], sc);
        }
    } else {
        constexpr int NI = PHYS::NI, NO = PHYS::NO;
        const int nft = p.nft;
        const int line0 = bx * 2 * p.G;
        int ng = (p.inner_len - line0 + 1) / 2;
        if (ng > p.G) ng = p.G;
        const long long ib = outer_off(p.si, by), ob = outer_off(p.so, by);

        if constexpr (MODE == TM_C2R || MODE == TM_FUSED) {
            // Hermitian pack of line pairs: Z[k] = A[k] + i B[k], Z[N-k] = conj(A[k]) + i conj(B[k])
            const int H = n_len / 2 + 1;
            const bool nfast = (p.si.s_n == 1);
            DDL_FOR_ITEMS(i, H * ng * NI) {
                int k, g, f;
                if (nfast) { k = i % H; int r = i / H; f = r % NI; g = r / NI; }
                else { g = i % ng; int r = i / ng; k = r % H; f = r / H; }
                const int l0 = line0 + 2 * g, l1 = l0 + 1;
                cplx A = mk(0.0, 0.0), B = mk(0.0, 0.0);
                if (k < p.kn) {
                    const int pk = p.si.n_tab ? p.si.n_tab[k] : k;
                    if (pk >= 0) {
                        const cplx* __restrict__ src = (const cplx*)p.in[f];
                        const long long a = ib + (long long)pk * p.si.s_n;
                        A = src[a + (long long)l0 * p.si.s_inner];
                        if (l1 < p.inner_len) B = src[a + (long long)l1 * p.si.s_inner];
                        if constexpr (N == 0) {
                            if (p.sh_on) { A = cmul(A, shear_phase(p, k, l0, -1.0)); B = cmul(B, shear_phase(p, k, l1, -1.0)); }
                        }
                    }
                }
                const int c = g * nft + f;
                if (k == 0 || 2 * k == n_len) {
                    tile[k * ld + c] = mk(A.x, B.x);
                } else {
                    tile[k * ld + c] = mk(A.x - B.y, A.y + B.x);
                    tile[(n_len - k) * ld + c] = mk(A.x + B.y, B.x - A.y);
                }
            }
            DDL_SYNC();
            fft_dif<N, +1>(p.rt, tile, ld, NI, nft, ng, tw);
        }

        if constexpr (MODE == TM_R2C) {
            const bool nfast = (p.si.s_n == 1);
            DDL_FOR_ITEMS(i, n_len * ng * NI) {
                int n, g, f;
                if (nfast) { n = i % n_len; int r = i / n_len; f = r % NI; g = r / NI; }
                else { g = i % ng; int r = i / ng; n = r % n_len; f = r / n_len; }
                const int l0 = line0 + 2 * g, l1 = l0 + 1;
                const double* __restrict__ src = (const double*)p.in[f];
                const long long a = ib + (long long)n * p.si.s_n;
                const double va = src[a + (long long)l0 * p.si.s_inner];
                const double vb = (l1 < p.inner_len) ? src[a + (long long)l1 * p.si.s_inner] : 0.0;
                tile[tpos<N>(p.rt, n) * ld + g * nft + f] = mk(va, vb);
            }
            DDL_SYNC();
        }

        if constexpr (MODE == TM_FUSED) {
            // real-space products at every grid point of both lines of the pair
            double m0 = 0.0, m1 = 0.0;
            DDL_FOR_ITEMS(i, n_len * ng) {
                const int g = i % ng, pos = i / ng;
                cplx* row = tile + pos * ld + g * nft;
                double ax[NI], ay[NI], ox[NO], oy[NO];
#pragma unroll
                for (int f = 0; f < NI; ++f) { cplx v = row[f]; ax[f] = v.x; ay[f] = v.y; }
                if (p.cfl) {
                    double a0 = 0.0, a1 = 0.0, b0 = 0.0, b1 = 0.0;
#pragma unroll
                    for (int f = 0; f < NI; ++f) {
                        if (f < PHYS::NDIM) { a0 = tmax_nn(a0, ax[f] * ax[f]); a1 = tmax_nn(a1, ay[f] * ay[f]); }
                        else if (f < PHYS::NDIM + PHYS::NG1) { b0 = tmax_nn(b0, ax[f] * ax[f]); b1 = tmax_nn(b1, ay[f] * ay[f]); }
                    }
                    m0 = tmax_nn(tmax_nn(m0, a0), a1);
                    m1 = tmax_nn(tmax_nn(m1, b0), b1);
                }
                PHYS::apply(ax, ox, p.pc);
                PHYS::apply(ay, oy, p.pc);
#pragma unroll
                for (int f = 0; f < NO; ++f) row[f] = mk(ox[f], oy[f]);
            }
            if (p.cfl) {      // uniform over the CTA
#if DDL_DEVICE_BUILD
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) {
                    m0 = tmax_nn(m0, __shfl_xor_sync(0xffffffffu, m0, off));
                    m1 = tmax_nn(m1, __shfl_xor_sync(0xffffffffu, m1, off));
                }
                if ((threadIdx.x & 31) == 0) {     // non-negative doubles order like their bit patterns
                    atomicMax(reinterpret_cast<unsigned long long*>(p.cfl), (unsigned long long)__double_as_longlong(m0));
                    atomicMax(reinterpret_cast<unsigned long long*>(p.cfl) + 1, (unsigned long long)__double_as_longlong(m1));
                }
#else
                p.cfl[0] = tmax_nn(p.cfl[0], m0);
                p.cfl[1] = tmax_nn(p.cfl[1], m1);
#endif
            }
            DDL_SYNC();
        }

        if constexpr (MODE == TM_R2C || MODE == TM_FUSED) {
            fft_dit<N, -1>(p.rt, tile, ld, NO, nft, ng, tw);
            const bool nfast = (p.so.s_n == 1);
            const double h = 0.5 * p.scale;
            const int kn = p.kn;
            DDL_FOR_ITEMS(i, kn * ng * NO) {
                int k, g, f;
                if (nfast) { k = i % kn; int r = i / kn; f = r % NO; g = r / NO; }
                else { g = i % ng; int r = i / ng; k = r % kn; f = r / kn; }
                const int pk = p.so.n_tab ? p.so.n_tab[k] : k;
                if (pk < 0) continue;
                const int c = g * nft + f;
                const cplx Zk = tile[k * ld + c], Zm = tile[((n_len - k) % n_len) * ld + c];
                const int l0 = line0 + 2 * g, l1 = l0 + 1;
                cplx* __restrict__ dst = (cplx*)p.out[f];
                const long long a = ob + (long long)pk * p.so.s_n;
                cplx Ak = mk((Zk.x + Zm.x) * h, (Zk.y - Zm.y) * h), Bk = mk((Zk.y + Zm.y) * h, (Zm.x - Zk.x) * h);
                if constexpr (N == 0) {
                    if (p.sh_on) { Ak = cmul(Ak, shear_phase(p, k, l0, +1.0)); Bk = cmul(Bk, shear_phase(p, k, l1, +1.0)); }
                }
                dst[a + (long long)l0 * p.so.s_inner] = Ak;
                if (l1 < p.inner_len) dst[a + (long long)l1 * p.so.s_inner] = Bk;
            }
        }

        if constexpr (MODE == TM_C2R) {
            const bool nfast = (p.so.s_n == 1);
            const double sc = p.scale;
            DDL_FOR_ITEMS(i, n_len * ng * NI) {
                int n, g, f;
                if (nfast) { n = i % n_len; int r = i / n_len; f = r % NI; g = r / NI; }
                else { g = i % ng; int r = i / ng; n = r % n_len; f = r / n_len; }
                const int l0 = line0 + 2 * g, l1 = l0 + 1;
                const cplx z = tile[tpos<N>(p.rt, n) * ld + g * nft + f];
                double* __restrict__ dst = (double*)p.out[f];
                const long long a = ob + (long long)n * p.so.s_n;
                dst[a + (long long)l0 * p.so.s_inner] = z.x * sc;
                if (l1 < p.inner_len) dst[a + (long long)l1 * p.so.s_inner] = z.y * sc;
            }
        }
    }
}

#define DDL_TILE_MAX_THREADS 768
#if DDL_DEVICE_BUILD
// round32() in api.cu launches at most 768 threads: the bound keeps every instantiation within the 64 K registers of an SM
// at that size (the advective-form MHD policy would otherwise take 88 registers x 768 threads and fail to launch)
template <int N, int MODE, int DIR, class PHYS>
__global__ void __launch_bounds__(DDL_TILE_MAX_THREADS) tile_kernel(const __grid_constant__ TileParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    tile_block<N, MODE, DIR, PHYS>(p, reinterpret_cast<cplx*>(smem_raw), blockIdx.x, blockIdx.y, blockIdx.z);
}
#endif

// Launch geometry + dispatch.  Returns 0 or a negative error code.
template <int N, int MODE, int DIR, class PHYS>
int launch_tile(const TileParams& p, int nthreads, ddl_stream_t stream) {
    const int per_tile = (MODE == TM_C2C) ? p.G : 2 * p.G;
    const int gx = (p.inner_len + per_tile - 1) / per_tile;
    const int gz = (MODE == TM_C2C) ? p.nf_in : 1;
    const int np = (MODE == TM_C2C) ? p.G : p.G * p.nft;
    if (p.ld < np) { set_error("launch_tile: ld %d < pencils %d", p.ld, np); return -1; }
    const size_t smem = (size_t)(N > 0 ? N : p.rt.n) * p.ld * sizeof(cplx);
#if DDL_DEVICE_BUILD
    auto kern = tile_kernel<N, MODE, DIR, PHYS>;
    if (smem > 48 * 1024) DDL_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid(gx, p.n_outer, gz);
    prof_begin(p.name, stream);
    kern<<<grid, nthreads, smem, stream>>>(p);
    prof_end(stream);
    DDL_CUDA_CHECK(cudaGetLastError());
#else
    (void)nthreads;
    prof_begin(p.name, stream);
    cplx* tile = (cplx*)malloc(smem);
    for (int bz = 0; bz < gz; ++bz)
        for (int by = 0; by < p.n_outer; ++by)
            for (int bx = 0; bx < gx; ++bx) tile_block<N, MODE, DIR, PHYS>(p, tile, bx, by, bz);
    free(tile);
#endif
    return 0;
}

}  // namespace ddl
