// Persistent, software-pipelined variant of the strided pencil pass (y and z passes of the one-rank pipeline; device build
// only; ddl_set_option("strided_staged", 1)).
//
// strided_fast (fast_kernels.cuh) loads the first butterfly stage straight from global memory into registers: a CTA's life is
// load (DRAM latency, nothing else to do) -> three compute stages -> store, and with two 512-thread CTAs per SM (64 registers,
// 64 KB tile each) the loads of one CTA overlap the compute of only one other.  ncu: 31-36 % of the warp samples are
// long-scoreboard stalls on exactly those loads, at 73-86 % of the HBM roof, with every execution pipe below 50 %.
// Here the same CTA shape stays on its SM and walks over tiles, and the first-stage inputs of the NEXT tile are fetched with
// cp.async (global -> shared, no registers held) into a staging buffer while the current tile is in its middle and last
// stages.  Each thread stages exactly the elements its own first-stage butterfly will read, so completion is a per-thread
// cp.async.wait_all and needs no barrier.  Staging holds the rows the pass really reads: the 2m+1 retained rows of a pruned
// input side (inverse passes), or the first 5 of the 8 first-stage rows of a full input side (forward passes; the other three are
// loaded directly as before) -- 43 KB next to the 64 KB tile, so two CTAs still share an SM.
// Same butterflies, same pruning by formula, same stores as strided_fast; results are bit-identical.
#pragma once
#include "fast_kernels.cuh"

#if DDL_DEVICE_BUILD
namespace ddl {

template <int N, int CX> struct StagedCfg {
    static constexpr int R0 = Fac<N>::radix(0);
    static constexpr int Q0 = N / R0;
    static constexpr int T = Q0;
    static constexpr int NT = CX * T;
    static constexpr int SROWS = (2 * (N / 3) + 1 + 7) / 8 * 8;                 // rows of staging: the retained rows of the 2/3 rule
    static constexpr int JS_FULL = SROWS / Q0;                                  // full input side: first-stage inputs j < JS_FULL are staged
    static constexpr size_t TILE = (size_t)N * CX * sizeof(cplx);
    static constexpr size_t SMEM = TILE + (size_t)SROWS * CX * sizeof(cplx);
    static constexpr bool ok = Fac<N>::S >= 2 && NT <= 512 && SMEM * 2 <= 227 * 1024 && JS_FULL >= 1;
};

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// PRUNED: the input side stores the retained rows only or in place (si.m >= 0: inverse passes); else every row (forward passes)
template <int N, int DIR, int CX, bool PRUNED>
__global__ void __launch_bounds__((StagedCfg<N, CX>::NT), 2)
strided_staged(const __grid_constant__ FastParams p) {
    using Cfg = StagedCfg<N, CX>;
    constexpr int R0 = Cfg::R0, Q0 = Cfg::Q0, T = Cfg::T;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cplx* tile = reinterpret_cast<cplx*>(smem_raw);
    cplx* stg = reinterpret_cast<cplx*>(smem_raw + Cfg::TILE);
    const cplx* __restrict__ tw = p.tw;
    const int c = threadIdx.x % CX, a = threadIdx.x / CX;
    const int ntiles = p.gx * p.gy * p.gz;
    const int m = p.si.m;                      // PRUNED: every retained row |index| <= m is staged; else first-stage inputs j < JS_FULL

    // staging slot of first-stage input j of this thread, or -1 (pruned row / loaded directly)
    auto slot = [&](int j) -> int {
        const int r = a + j * Q0;
        if constexpr (!PRUNED) return j < Cfg::JS_FULL ? r : -1;
        if (r <= m) return r;
        if (r >= N - m) return r - (N - 2 * m - 1);
        return -1;
    };
    // tile t -> (inner chunk, outer, field) without integer divisions (three per tile and thread would cost as much as a butterfly)
    auto decode = [&](int t, int& bx, int& by, int& bz) { tile_decode(TileDecode{p.gx, p.gx * p.gy, p.mx, p.mxy}, t, bx, by, bz); };
    auto prefetch = [&](int t) {
        int bx, by, bz;
        decode(t, bx, by, bz);
        const int inner = bx * CX + c;
        if (inner >= p.inner_len) return;
        const cplx* __restrict__ in = p.in[bz];
        const long long ib = (long long)(p.si.outer_tab ? p.si.outer_tab[by] : by) * p.si.s_outer + inner;
#pragma unroll
        for (int j = 0; j < R0; ++j) {
            const int s = slot(j);
            if (s >= 0) {
                const int row = fast_row<N>(a + j * Q0, m, p.si.compact);
                cp_async16(&stg[s * CX + c], &in[ib + (long long)row * p.si.s_n]);
            }
        }
    };

    if (p.stagger_ns > 0) {
        for (int k = blockIdx.x / p.nsm; k > 0; --k) __nanosleep(p.stagger_ns);      // break the lock step of an SM's CTAs
    }
    int t = blockIdx.x;
    if (t < ntiles) prefetch(t);
#pragma unroll 1
    for (; t < ntiles; t += gridDim.x) {
        int bx, by, bz;
        decode(t, bx, by, bz);
        const bool live = bx * CX + c < p.inner_len;

        // ---- stage 0: staged (or directly loaded) inputs -> registers -> tile
        {
            cplx v[R0];
            if constexpr (!PRUNED) {
                // the rows that are not staged first, so that their latency overlaps the wait for the staged ones
                const cplx* __restrict__ in = p.in[bz] + ((long long)(p.si.outer_tab ? p.si.outer_tab[by] : by) * p.si.s_outer + bx * CX + c);
#pragma unroll
                for (int j = Cfg::JS_FULL; j < R0; ++j) v[j] = live ? in[(long long)(a + j * Q0) * p.si.s_n] : mk(0.0, 0.0);
            }
            cp_async_wait_all();
#pragma unroll
            for (int j = 0; j < R0; ++j) {
                const int s = slot(j);
                if (s >= 0) v[j] = live ? stg[s * CX + c] : mk(0.0, 0.0);
                else if constexpr (PRUNED) v[j] = mk(0.0, 0.0);
            }
            dftR<R0, DIR>(v);
            twiddles_ld<R0, DIR>(v, a, tw);
#pragma unroll
            for (int r = 0; r < R0; ++r) tile[(r * Q0 + a) * CX + c] = v[r];
        }
        __syncthreads();
        // every thread has read its staged inputs: the buffer takes the next tile while this one is in its later stages
        if (t + (int)gridDim.x < ntiles) prefetch(t + gridDim.x);
        // ---- middle stages: smem -> smem
        FastMid<N, DIR, CX, 1>::run(tile, c, a, tw);
        // ---- last stage: smem -> registers -> global
        {
            constexpr int SL = Fac<N>::S - 1;
            constexpr int R = Fac<N>::radix(SL);
            constexpr int ITEMS = N / R;
            const double sc = p.scale;
            cplx* __restrict__ out = p.out[bz] + ((long long)(p.so.outer_tab ? p.so.outer_tab[by] : by) * p.so.s_outer + bx * CX + c);
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
                    const int row = fast_row<N>(k0 + r * (N / R), p.so.m, p.so.compact);
                    if (live && row >= 0) out[(long long)row * p.so.s_n] = scal(v[r], sc);
                }
            }
        }
        __syncthreads();          // the tile is free for the next iteration's first stage
    }
}

// returns 0 on success, 1 if this pass has no staged kernel (the caller launches strided_fast)
template <int N, int DIR>
int launch_strided_staged(const FastParams& p, int nf, int n_outer, const char* name, ddl_stream_t stream) {
    constexpr int CX = FastCX<N>::value;
    using Cfg = StagedCfg<N, CX>;
    if constexpr (!Cfg::ok) {
        return 1;
    } else {
        if (p.si.m >= 0 && 2 * p.si.m + 1 > Cfg::SROWS) return 1;
        auto kern = p.si.m >= 0 ? strided_staged<N, DIR, CX, true> : strided_staged<N, DIR, CX, false>;
        static DeviceOnce once[2];
        const int ctas = once[p.si.m >= 0 ? 1 : 0].get([&]() -> int {
            DDL_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM));
            DDL_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
            int dev = 0, sms = 0, per_sm = 0;
            DDL_CUDA_CHECK(cudaGetDevice(&dev));
            DDL_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
            DDL_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, Cfg::NT, Cfg::SMEM));
            if (per_sm < 1) { set_error("strided_staged: kernel does not fit an SM"); return -2; }
            return sms * per_sm;
        });
        if (ctas < 0) return -2;
        FastParams q = p;
        q.gx = (p.inner_len + CX - 1) / CX; q.gy = n_outer; q.gz = nf;
        const long long ntiles = (long long)q.gx * q.gy * q.gz;
        if (ntiles < 1) return 0;
        const long long gxy = (long long)q.gx * q.gy;
        if (ntiles >= (1LL << 32) / gxy) return 1;             // the multiply-high decode is exact below 2^32 / divisor
        q.mx = tile_magic(q.gx);
        q.mxy = tile_magic(gxy);
        q.stagger_ns = g_persist_stagger_ns;
        q.nsm = ctas >= 2 ? ctas / 2 : 1;
        const int grid = (int)(ntiles < ctas ? ntiles : ctas);
        prof_begin(name, stream);
        kern<<<grid, Cfg::NT, Cfg::SMEM, stream>>>(q);
        prof_end(stream);
        DDL_CUDA_CHECK(cudaGetLastError());
        return 0;
    }
}

}  // namespace ddl
#endif  // DDL_DEVICE_BUILD
