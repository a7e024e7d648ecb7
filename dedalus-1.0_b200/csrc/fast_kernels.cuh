// Specialised sm_100a kernels for the hot 3-D passes (device build only).  Same mathematics as
// the generic tile kernel (tile_kernel.cuh), which stays as the reference implementation for
// the host emulation and for the shapes these kernels do not cover; the GPU tests compare both
// against the oracle.
//
//  strided_fast : complex pass along a strided axis (y / z passes).  One thread = one first-
//                 stage butterfly of one pencil.  First-stage inputs go global -> registers,
//                 last-stage outputs registers -> global (pruned rows predicated, no index
//                 tables), so shared memory only carries the inter-stage exchanges.  All index
//                 arithmetic is compile-time; CX*16 B contiguous per row keeps every warp access
//                 on whole 128-B lines (workspace pitch is a multiple of 8 complex).
#pragma once
#include "ddl_common.cuh"
#include "tile_kernel.cuh"

#if DDL_DEVICE_BUILD
namespace ddl {

struct FastSide {
    long long s_n, s_outer;   // element strides of the transform axis and of the outer index
    const int* outer_tab;     // outer index -> stored outer index (NULL = identity)
    int m;                    // retained rows |index| <= m, or -1 = every row present
    int compact;              // retained rows stored contiguously (workspace) instead of in place (state)
    const int* row_tab;       // if set: stored row of logical row r is row_tab[r] (-1 = pruned)
    // slab exchange buffers: stored row r sits at (r >> split_shift) * s_blk + (r & split_mask) * s_n
    // (split_shift = 31, split_mask = 0x7fffffff: plain rows)
    int split_shift, split_mask;
    long long s_blk;
    // output side only, slab exchange fused into the pass: block b of field f starts at
    // peer_tab[f * nblk + b] (a pointer into rank b's arena, mapped through CUDA IPC) instead of
    // out[f] + b * s_blk; the block of a row is own_tab[row] if given, else row >> split_shift
    cplx* const* peer_tab;
    const unsigned char* own_tab;
    int nblk;
    long long peer_off;       // added to every peer address (sub-range of outer planes)
};

// tile t -> (inner chunk bx, outer by, field bz) of a gx x gy x gz tile grid without integer division: t / d == umulhi(t, m) with
// m = ceil(2^32 / d) for t < 2^32 / d (checked by the launchers); d == 1 has no such 32-bit m (2^32) and is the identity
struct TileDecode {
    int gx, gxy;
    unsigned mx, mxy;
};
__device__ __forceinline__ void tile_decode(const TileDecode& d, int t, int& bx, int& by, int& bz) {
    bz = d.gxy == 1 ? t : (int)__umulhi((unsigned)t, d.mxy);
    const int rem = t - bz * d.gxy;
    by = d.gx == 1 ? rem : (int)__umulhi((unsigned)rem, d.mx);
    bx = rem - by * d.gx;
}
inline unsigned tile_magic(long long d) { return d <= 1 ? 0u : (unsigned)(((1ULL << 32) + d - 1) / d); }

__device__ __forceinline__ long long fast_row_off(const FastSide& s, int r) {
    return (long long)(r >> s.split_shift) * s.s_blk + (long long)(r & s.split_mask) * s.s_n;
}

struct FastParams {
    const cplx* in[DDL_MAXF];
    cplx* out[DDL_MAXF];
    FastSide si, so;
    int inner_len;
    double scale;
    const cplx* tw;
    int gx, gy, gz;           // EXT kernels walk the (inner chunk, outer, field) tiles themselves: tile t = (t % gx, t / gx % gy, t / (gx gy))
    int stagger_ns, nsm;      // strided_staged: the k-th CTA of an SM starts k * stagger_ns late (see xfused_rot.cuh)
    unsigned mx, mxy;         // ceil(2^32 / gx), ceil(2^32 / (gx gy)): t / d == umulhi(t, m) for t < 2^32 / d (strided_staged)
};

// Slab passes only (EXT): number of CTAs a peer-store pass may occupy, 0 = one CTA per tile (ddl_set_option("peer_pass_ctas", n)).
// A pass whose stores cross NVLink is bound by the links, not by the SMs: a few resident CTAs per SM pair saturate them, and a
// grid limited to that many CTAs leaves the rest of every SM to the HBM / FP64-bound pass running on the other stream.
extern int g_peer_pass_ctas;
extern int g_persist_stagger_ns;

// stored row of logical row r, or -1 if the row is pruned
template <int N>
__device__ __forceinline__ int fast_row(int r, int m, int compact) {
    if (m < 0) return r;
    if (r <= m) return r;
    if (r >= N - m) return compact ? r - (N - 2 * m - 1) : r;
    return -1;
}

template <int R, int DIR>
__device__ __forceinline__ void twiddles_ld(cplx (&v)[R], int step, const cplx* __restrict__ tw) {
#pragma unroll
    for (int r = 1; r < R; ++r) {
        cplx w = __ldg(&tw[r * step]);
        if (DIR > 0) w.y = -w.y;
        v[r] = cmul(v[r], w);
    }
}

// middle stage s (neither first nor last): smem -> smem
template <int N, int S_IDX, int DIR, int CX, int T>
__device__ __forceinline__ void fast_mid_stage(cplx* tile, int c, int a, const cplx* __restrict__ tw) {
    constexpr int R = Fac<N>::radix(S_IDX);
    constexpr int P = Fac<N>::P(S_IDX);
    constexpr int M = N / P;
    constexpr int Q = M / R;
    constexpr int ITEMS = N / R;
#pragma unroll
    for (int w = 0; w < ITEMS; w += T) {
        const int wi = w + a;
        if (ITEMS % T != 0 && wi >= ITEMS) break;
        const int q = wi / Q, b = wi % Q;
        cplx* base = tile + (q * M + b) * CX + c;
        cplx v[R];
#pragma unroll
        for (int j = 0; j < R; ++j) v[j] = base[j * Q * CX];
        dftR<R, DIR>(v);
        twiddles_ld<R, DIR>(v, b * P, tw);
#pragma unroll
        for (int j = 0; j < R; ++j) base[j * Q * CX] = v[j];
    }
}

template <int N, int DIR, int CX, int S_IDX>
struct FastMid {
    static constexpr int T = N / Fac<N>::radix(0);
    __device__ __forceinline__ static void run(cplx* tile, int c, int a, const cplx* __restrict__ tw) {
        if constexpr (S_IDX < Fac<N>::S - 1) {
            fast_mid_stage<N, S_IDX, DIR, CX, T>(tile, c, a, tw);
            __syncthreads();
            FastMid<N, DIR, CX, S_IDX + 1>::run(tile, c, a, tw);
        }
    }
};

// EXT = false: plain rows on both sides (every single-GPU pass) - the lean path;
// EXT = true : slab extensions (peer-blocked rows, row-position tables, peer-store output)
template <int N, int DIR, int CX, bool EXT>
__global__ void __launch_bounds__(CX * (N / Fac<N>::radix(0)), (CX * (N / Fac<N>::radix(0)) <= 512) ? 2 : 1)
strided_fast(const __grid_constant__ FastParams p) {
    static_assert(Fac<N>::S >= 2, "strided_fast needs at least two stages");
    constexpr int R0 = Fac<N>::radix(0);
    constexpr int T = N / R0;              // threads per pencil
    constexpr int Q0 = N / R0;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cplx* tile = reinterpret_cast<cplx*>(smem_raw);
    const cplx* __restrict__ tw = p.tw;

    const int c = threadIdx.x % CX, a = threadIdx.x / CX;
    // EXT: a (possibly limited) 1-D grid of CTAs walks the tiles; plain kernels: one CTA per tile of the 3-D grid
    const int ntiles = EXT ? p.gx * p.gy * p.gz : 1;
#pragma unroll 1
    for (int t = EXT ? (int)blockIdx.x : 0; t < ntiles; t += EXT ? (int)gridDim.x : 1) {
    // EXT: tile -> (inner chunk, outer, field) by multiply-high (p.mx, p.mxy: exact below 2^32 / divisor, checked by the launcher);
    // three integer divisions per tile and thread would cost a measurable part of a butterfly stage
    int bx = (int)blockIdx.x, by = (int)blockIdx.y, bz = (int)blockIdx.z;
    if constexpr (EXT) tile_decode(TileDecode{p.gx, p.gx * p.gy, p.mx, p.mxy}, t, bx, by, bz);
    if (EXT && t != (int)blockIdx.x) __syncthreads();        // the previous tile's last stage has read the shared tile
    const int inner = bx * CX + c;
    const bool live = inner < p.inner_len;
    const cplx* __restrict__ in = p.in[bz];
    cplx* __restrict__ out = p.out[bz];
    const long long ib = (long long)(p.si.outer_tab ? p.si.outer_tab[by] : by) * p.si.s_outer + inner;
    const long long ob = (long long)(p.so.outer_tab ? p.so.outer_tab[by] : by) * p.so.s_outer + inner;

    // ---- stage 0: global -> registers -> smem
    {
        cplx v[R0];
#pragma unroll
        for (int j = 0; j < R0; ++j) {
            if constexpr (EXT) {
                const int row = p.si.row_tab ? p.si.row_tab[a + j * Q0] : fast_row<N>(a + j * Q0, p.si.m, p.si.compact);
                v[j] = (live && row >= 0) ? in[ib + fast_row_off(p.si, row)] : mk(0.0, 0.0);
            } else {
                const int row = fast_row<N>(a + j * Q0, p.si.m, p.si.compact);
                v[j] = (live && row >= 0) ? in[ib + (long long)row * p.si.s_n] : mk(0.0, 0.0);
            }
        }
        dftR<R0, DIR>(v);
        twiddles_ld<R0, DIR>(v, a, tw);
#pragma unroll
        for (int r = 0; r < R0; ++r) tile[(r * Q0 + a) * CX + c] = v[r];
    }
    __syncthreads();
    // ---- middle stages: smem -> smem
    FastMid<N, DIR, CX, 1>::run(tile, c, a, tw);
    // ---- last stage: smem -> registers -> global
    {
        constexpr int SL = Fac<N>::S - 1;
        constexpr int R = Fac<N>::radix(SL);
        constexpr int ITEMS = N / R;
        const double sc = p.scale;
#pragma unroll
        for (int w = 0; w < ITEMS; w += T) {
            const int q = w + a;
            if (ITEMS % T != 0 && q >= ITEMS) break;
            cplx v[R];
            const cplx* base = tile + (q * R) * CX + c;
#pragma unroll
            for (int j = 0; j < R; ++j) v[j] = base[j * CX];
            dftR<R, DIR>(v);
            const int k0 = index_of_pos<N>(q * R);
#pragma unroll
            for (int r = 0; r < R; ++r) {
                if constexpr (!EXT) {
                    const int row = fast_row<N>(k0 + r * (N / R), p.so.m, p.so.compact);
                    if (live && row >= 0) out[ob + (long long)row * p.so.s_n] = scal(v[r], sc);
                    continue;
                }
                const int row = p.so.row_tab ? p.so.row_tab[k0 + r * (N / R)] : fast_row<N>(k0 + r * (N / R), p.so.m, p.so.compact);
                if (live && row >= 0) {
                    if (p.so.peer_tab) {
                        const int blk = p.so.own_tab ? p.so.own_tab[row] : (row >> p.so.split_shift);
                        cplx* __restrict__ dst = p.so.peer_tab[bz * p.so.nblk + blk];
                        dst[ob + p.so.peer_off + (long long)(row & p.so.split_mask) * p.so.s_n] = scal(v[r], sc);
                    } else {
                        out[ob + fast_row_off(p.so, row)] = scal(v[r], sc);
                    }
                }
            }
        }
    }
    }   // tile loop
}

// pencils per tile for the strided pass of length N
template <int N> struct FastCX {
    static constexpr int T = N / Fac<N>::radix(0);
    // N = 512: 128-byte row segments (CX = 8, two 512-thread CTAs per SM); 64-byte segments cost 25 % there.
    // N = 1024: measured the other way round (y passes 5 % faster with two 512-thread CTAs of CX = 4 than with
    // one 1024-thread CTA of CX = 8), profiles/ncu_r1.md
    static constexpr int value = (T >= 128) ? 4 : (T >= 64 ? 8 : (T >= 32 ? 16 : 32));
};

template <int N, int DIR, bool EXT>
int launch_strided_fast_v(const FastParams& p, int nf, int n_outer, const char* name, ddl_stream_t stream);

template <int N, int DIR> int launch_strided_staged(const FastParams& p, int nf, int n_outer, const char* name, ddl_stream_t stream);   // fast_staged.cuh
extern int g_strided_staged;

template <int N, int DIR>
int launch_strided_fast(const FastParams& p, int nf, int n_outer, const char* name, ddl_stream_t stream) {
    const bool ext = p.si.row_tab || p.so.row_tab || p.so.peer_tab || p.si.split_shift != 31 || p.so.split_shift != 31;
    if (!ext && g_strided_staged) {
        // persistent CTAs with the next tile's first-stage inputs fetched by cp.async (fast_staged.cuh); 1 = not applicable
        const int rc = launch_strided_staged<N, DIR>(p, nf, n_outer, name, stream);
        if (rc != 1) return rc;
    }
    return ext ? launch_strided_fast_v<N, DIR, true>(p, nf, n_outer, name, stream)
               : launch_strided_fast_v<N, DIR, false>(p, nf, n_outer, name, stream);
}

template <int N, int DIR, bool EXT>
int launch_strided_fast_v(const FastParams& p, int nf, int n_outer, const char* name, ddl_stream_t stream) {
    constexpr int CX = FastCX<N>::value;
    constexpr int T = N / Fac<N>::radix(0);
    auto kern = strided_fast<N, DIR, CX, EXT>;
    const size_t smem = (size_t)N * CX * sizeof(cplx);
    static DeviceOnce once;
    if (once.get([&]() -> int {
            DDL_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            return 1;
        }) < 0) return -2;
    dim3 grid((p.inner_len + CX - 1) / CX, n_outer, nf);
    if constexpr (EXT) {
        FastParams q = p;
        q.gx = (int)grid.x; q.gy = (int)grid.y; q.gz = (int)grid.z;
        const long long ntiles = (long long)q.gx * q.gy * q.gz;
        const long long gxy = (long long)q.gx * q.gy;
        if (ntiles >= (1LL << 32) / gxy) { set_error("strided_fast: %lld tiles exceed the range of the tile decode", ntiles); return -1; }
        q.mx = tile_magic(q.gx);
        q.mxy = tile_magic(gxy);
        const bool peer = p.so.peer_tab != nullptr;
        const long long lim = (peer && g_peer_pass_ctas > 0 && g_peer_pass_ctas < ntiles) ? g_peer_pass_ctas : ntiles;
        prof_begin(name, stream);
        kern<<<(unsigned)lim, CX * T, smem, stream>>>(q);
        prof_end(stream);
        DDL_CUDA_CHECK(cudaGetLastError());
        return 0;
    }
    prof_begin(name, stream);
    kern<<<grid, CX * T, smem, stream>>>(p);
    prof_end(stream);
    DDL_CUDA_CHECK(cudaGetLastError());
    return 0;
}

}  // namespace ddl
#endif  // DDL_DEVICE_BUILD
