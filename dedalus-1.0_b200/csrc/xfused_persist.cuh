// Persistent variant of the fused x pass (launch variant 4; device build only).
//
// Same mathematics, same shared-memory tile and the same butterflies as xfused_kernel.cuh.  What changes is how the
// half-spectra reach the SM.  In the one-shot kernel every CTA starts with the Hermitian pack straight from global memory
// into registers: all of its warps sit on the DRAM latency at the same moment (ncu: a quarter of all warp samples are
// long-scoreboard stalls on those loads, at 43 % of the HBM roof).  Here
//   * one CTA per resident slot stays on its SM and walks over the line pairs;
//   * the 2 x NI retained half-lines of the NEXT pair (NI x 2 x kn x 16 B, 33 KB for 3-D MHD at 512) are brought into a
//     staging buffer by the bulk-copy engine (cp.async.bulk -> mbarrier complete_tx) while the current pair is in its
//     butterflies: a warp that would otherwise idle through the inverse phase waits for the staging buffer to be drained
//     (second mbarrier, one arrival per inverse warp) and issues the 2 x NI copies;
//   * the first butterfly stage reads the staged lines from shared memory (conflict-free 16-byte accesses), so no warp
//     ever waits on DRAM inside the loop; the loop-invariant stage-0 twiddle stays in registers;
//   * one warp per pencil slot (NS warps): the forward half is a single round, the inverse half uses NI of them;
//   * 2 CTAs x 9 warps per SM for 3-D MHD leave 112 registers per thread: no spills.
// Shared memory per CTA: NS x N x 16 B tile + NI x 2 x SL x 16 B staging (72 + 33.8 KB for MHD at 512), 2 CTAs per SM.
#pragma once
#include "xfused_kernel.cuh"

#if DDL_DEVICE_BUILD
namespace ddl {

__device__ __forceinline__ unsigned xp_smem(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void xp_mbar_init(unsigned long long* b, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(xp_smem(b)), "r"(count) : "memory");
}
__device__ __forceinline__ void xp_mbar_expect_tx(unsigned long long* b, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(xp_smem(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void xp_mbar_arrive(unsigned long long* b) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(xp_smem(b)) : "memory");
}
__device__ __forceinline__ void xp_mbar_wait(unsigned long long* b, unsigned parity) {
    unsigned ok;
    do {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                     : "=r"(ok) : "r"(xp_smem(b)), "r"(parity) : "memory");
    } while (!ok);
}
// global -> shared bulk copy (bytes: multiple of 16, both addresses 16-byte aligned), completion counted on the mbarrier
__device__ __forceinline__ void xp_bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* b) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(xp_smem(dst)), "l"(src), "r"(bytes), "r"(xp_smem(b)) : "memory");
}

template <int N, class PHYS> struct XPersistCfg {
    static constexpr int NI = PHYS::NI, NO = PHYS::NO;
    static constexpr int NS = NI > NO ? NI : NO;
    static constexpr int R0 = XFac<N>::radix(0);
    static constexpr int TP = N / R0;
    static constexpr bool ok = XFac<N>::ok && TP == 32 && XFac<N>::S == 2;      // one warp per pencil, two outer stages: N = 512
    static constexpr int NW = NS;
    static constexpr int NT = 32 * NW;
    static constexpr int SL = ((N / 3 + 1 + 7) / 8) * 8;                        // staged modes per line (2/3 rule, whole 128-B lines)
    static constexpr size_t TILE = (size_t)NS * N * sizeof(cplx);
    static constexpr size_t STAGE = (size_t)NI * 2 * SL * sizeof(cplx);
    static constexpr size_t SMEM = TILE + STAGE + 64;
    static constexpr int BY_SMEM = (int)((227 * 1024) / SMEM);
    static constexpr int BY_THREADS = 2048 / NT;
    static constexpr int MINB = BY_SMEM < BY_THREADS ? (BY_SMEM < 1 ? 1 : BY_SMEM) : BY_THREADS;
};

template <int N, class PHYS, bool CFL>
__global__ void __launch_bounds__((XPersistCfg<N, PHYS>::NT), (XPersistCfg<N, PHYS>::MINB))
xfused_persist_kernel(const __grid_constant__ XFusedParams p, int n_outer, int pairs_per_plane) {
    using Cfg = XPersistCfg<N, PHYS>;
    constexpr int NI = Cfg::NI, NO = Cfg::NO, NS = Cfg::NS, R0 = Cfg::R0, Q0 = N / R0, NT = Cfg::NT, SL = Cfg::SL;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    cplx* tile = reinterpret_cast<cplx*>(smem_raw);
    cplx* stage = reinterpret_cast<cplx*>(smem_raw + Cfg::TILE);
    unsigned long long* full = reinterpret_cast<unsigned long long*>(smem_raw + Cfg::TILE + Cfg::STAGE);
    unsigned long long* empty = full + 1;

    const int w = threadIdx.x >> 5, a = threadIdx.x & 31;
    const cplx* __restrict__ tw = p.tw;
    const int kn = p.kn;
    const unsigned line_bytes = (unsigned)kn * (unsigned)sizeof(cplx);
    const long long total = (long long)n_outer * pairs_per_plane;
    // the warp that issues the copies: the last one, which has no pencil in the inverse half when NO > NI
    constexpr int PRODW = Cfg::NW - 1;

    if (threadIdx.x == 0) {
        xp_mbar_init(full, 1);
        xp_mbar_init(empty, NI);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    auto issue = [&](long long q) {
        const int by = (int)(q / pairs_per_plane), bx = (int)(q % pairs_per_plane);
        const long long off = (long long)by * p.s_outer + (long long)(2 * bx) * p.pitch;
        xp_mbar_expect_tx(full, 2u * NI * line_bytes);
#pragma unroll
        for (int f = 0; f < NI; ++f) {
            xp_bulk_g2s(stage + (f * 2 + 0) * SL, p.in[f] + off, line_bytes, full);
            xp_bulk_g2s(stage + (f * 2 + 1) * SL, p.in[f] + off + p.pitch, line_bytes, full);
        }
    };
    long long q = blockIdx.x;
    if (threadIdx.x == PRODW * 32 && q < total) issue(q);

    // loop invariants: the stage-0 twiddle of this lane, the scale of the unpack
    const cplx w0 = (a != 0) ? DDL_LDG(&tw[a]) : mk(1.0, 0.0);
    const double h = 0.5 * p.scale;
    double m0 = 0.0, m1 = 0.0;
    unsigned phase = 0;

    for (; q < total; q += gridDim.x, phase ^= 1) {
        const int by = (int)(q / pairs_per_plane), bx = (int)(q % pairs_per_plane);
        const long long off = (long long)by * p.s_outer + (long long)(2 * bx) * p.pitch;

        // ================= inverse: warp f owns pencil f.  Stage 0: staged half-lines -> Hermitian pack -> registers -> tile
        if (w < NI) {
            xp_mbar_wait(full, phase);
            const cplx* __restrict__ A = stage + (w * 2) * SL;
            const cplx* __restrict__ B = A + SL;
            cplx v[R0];
#pragma unroll
            for (int j = 0; j < R0; ++j) {
                const int e = a + j * Q0;
                cplx z = mk(0.0, 0.0);
                if (e < kn) {
                    const cplx za = A[e], zb = B[e];
                    z = (e == 0) ? mk(za.x, zb.x) : mk(za.x - zb.y, za.y + zb.x);
                } else if (N - e < kn) {
                    const cplx za = A[N - e], zb = B[N - e];
                    z = mk(za.x + zb.y, zb.x - za.y);
                }
                v[j] = z;
            }
            // the staged lines of this pencil are in registers: one arrival per inverse warp frees the buffer.  The reads above went
            // through the generic proxy, the refill will be written by the async proxy: fence.proxy.async orders the two (the
            // mbarrier release / acquire pair orders the threads).  compute-sanitizer's racecheck does not follow this hand-over
            // and reports the refill against these reads (profiles/r2/sanitizer_racecheck_xfused512.log).
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (a == 0) xp_mbar_arrive(empty);
            xdft<R0, +1>(v);
            if (a != 0) xtwiddle<R0, true>(v, conj(w0));
            cplx* T = tile + w * N;
            const int sb = xsw<N>(a);
#pragma unroll
            for (int r = 0; r < R0; ++r) T[sb ^ xsw<N>(r * Q0)] = v[xreg<R0>(r)];
            __syncwarp();
            xstage<N, 1, +1, false, 32>(T, a, tw);
        }
        if (threadIdx.x == PRODW * 32) {
            const long long qn = q + gridDim.x;
            if (qn < total) {
                xp_mbar_wait(empty, phase);      // every inverse warp has read its staged lines of pair q
                issue(qn);
            }
        }
        __syncthreads();

        // ================= innermost radix-2 (inverse) -> products -> innermost radix-2 (forward), as in xfused_block
        {
            double* td = reinterpret_cast<double*>(tile);
#pragma unroll 1
            for (int i = threadIdx.x; i < N; i += NT) {
                const int c = i & 1, wp = i >> 1;
                const int s0 = 2 * xsw<N>(2 * wp) + c, s1 = s0 ^ 2;
                double u0[NI], u1[NI], o0[NO], o1[NO];
#pragma unroll
                for (int f = 0; f < NI; ++f) {
                    const double x0 = td[f * 2 * N + s0], x1 = td[f * 2 * N + s1];
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
                    td[f * 2 * N + s0] = o0[f] + o1[f];
                    td[f * 2 * N + s1] = o0[f] - o1[f];
                }
            }
        }
        __syncthreads();

        // ================= forward: warp f owns pencil f: stage 1 (shared), stage 0 -> registers, shuffle unpack, store
        if (w < NO) {
            cplx* T = tile + w * N;
            xstage<N, 1, -1, true, 32>(T, a, tw);
            __syncwarp();
            const int sb = xsw<N>(a);
            cplx v[R0];
#pragma unroll
            for (int j = 0; j < R0; ++j) v[j] = T[sb ^ xsw<N>(j * Q0)];
            if (a != 0) xtwiddle<R0, false>(v, w0);
            xdft<R0, -1>(v);
            cplx* __restrict__ dst = p.out[w] + off;
            const int src = (32 - a) & 31;
            const int rmax = (kn - 1) / Q0;
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
                    if (k < kn) {
                        dst[k] = mk((Zk.x + Zm.x) * h, (Zk.y - Zm.y) * h);
                        dst[p.pitch + k] = mk((Zk.y + Zm.y) * h, (Zm.x - Zk.x) * h);
                    }
                }
            }
        }
        // no barrier here: in the next iteration warp f overwrites slot f, which only warp f has just read; slots >= NI are
        // written next in the product phase, behind the barrier that follows the inverse half
    }
    if constexpr (CFL) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            m0 = xmax_nn(m0, __shfl_xor_sync(0xffffffffu, m0, o));
            m1 = xmax_nn(m1, __shfl_xor_sync(0xffffffffu, m1, o));
        }
        if (a == 0) {
            atomicMax(reinterpret_cast<unsigned long long*>(p.cfl), (unsigned long long)__double_as_longlong(m0));
            atomicMax(reinterpret_cast<unsigned long long*>(p.cfl) + 1, (unsigned long long)__double_as_longlong(m1));
        }
    }
}

// returns 0 on success, 1 if this (N, physics, mask) has no persistent kernel (the caller uses the one-shot kernel)
template <int N, class PHYS, bool CFL>
int launch_xfused_persist(const XFusedParams& p, int n_outer, ddl_stream_t stream) {
    using Cfg = XPersistCfg<N, PHYS>;
    if constexpr (!Cfg::ok) {
        return 1;
    } else {
        if (p.kn > Cfg::SL || (p.n_lines & 1) || (p.pitch & 7)) return 1;
        for (int f = 0; f < Cfg::NI; ++f)
            if (reinterpret_cast<uintptr_t>(p.in[f]) & 15) return 1;
        auto kern = xfused_persist_kernel<N, PHYS, CFL>;
        static DeviceOnce once;                                // resident CTAs of this instantiation on the current device
        const int ctas = once.get([&]() -> int {
            DDL_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM));
            DDL_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
            int dev = 0, sms = 0, per_sm = 0;
            DDL_CUDA_CHECK(cudaGetDevice(&dev));
            DDL_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
            DDL_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, Cfg::NT, Cfg::SMEM));
            if (per_sm < 1) { set_error("xfused_persist: kernel does not fit an SM"); return -2; }
            return sms * per_sm;
        });
        if (ctas < 0) return -2;
        const int pairs = p.n_lines / 2;
        const long long total = (long long)pairs * n_outer;
        const int grid = (int)(total < ctas ? total : ctas);
        if (grid < 1) return 0;
        prof_begin("x_fused", stream);
        kern<<<grid, Cfg::NT, Cfg::SMEM, stream>>>(p, n_outer, pairs);
        prof_end(stream);
        DDL_CUDA_CHECK(cudaGetLastError());
        return 0;
    }
}

}  // namespace ddl
#endif  // DDL_DEVICE_BUILD
