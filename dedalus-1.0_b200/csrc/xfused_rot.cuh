// Persistent variant of the fused x pass WITHOUT a staging buffer (launch variant 5; device build only).
//
// What the measurements of variants 0-4 say: the kernel is bound by how well FP64 pipe, issue slots and the
// shared-memory pipe overlap at 18 warps per SM; three independent 6-warp CTAs per SM beat two 9-warp ones, and hiding the
// load latency behind a staging buffer costs the third CTA and loses.  This variant keeps the winning shape -- 6 warps, one
// 72 KB tile, 3 CTAs per SM, first-stage inputs straight from global memory into registers -- and removes the two things that
// shape still wastes:
//   * three of the six warps idle through the second forward round (9 product pencils on 6 warps): here they start the
//     inverse transforms of the NEXT line pair into the tile slots the first forward round has already drained (slot f is read
//     last by warp f in that round and written next by warp f, so program order is the only synchronisation needed);
//   * 131 072 CTA launches per pass with their set-up (parameters, twiddle loads, address arithmetic): the CTAs stay resident
//     and walk the line pairs; the stage-0 twiddle of a lane is loop-invariant and stays in registers.
// Same butterflies, same order of operations per pencil as xfused_kernel: results equal to round-off.
// MEASURED SLOWER than the one-shot kernel (5.07 vs 3.99 ms): opt-in, kept with its test as the record.
#pragma once
#include "xfused_kernel.cuh"

#if DDL_DEVICE_BUILD
namespace ddl {

extern int g_persist_stagger_ns;      // ddl_set_option("persist_stagger_ns", ns)

template <int N, class PHYS> struct XRotCfg {
    static constexpr int NI = PHYS::NI, NO = PHYS::NO;
    static constexpr int NS = NI > NO ? NI : NO;
    static constexpr int R0 = XFac<N>::radix(0);
    static constexpr int TP = N / R0;
    static constexpr int NW = 6;
    static constexpr int NT = 32 * NW;
    // one warp per pencil, two outer stages (N = 512), at most two forward rounds, every inverse pencil on a warp of its own
    static constexpr bool ok = XFac<N>::ok && TP == 32 && XFac<N>::S == 2 && NI <= NW && NO <= 2 * NW;
    static constexpr size_t SMEM = (size_t)NS * N * sizeof(cplx);
    static constexpr int BY_SMEM = (int)((225 * 1024) / (SMEM + 1024));
    static constexpr int MINB = BY_SMEM < 3 ? (BY_SMEM < 1 ? 1 : BY_SMEM) : 3;
    static constexpr int WPS = (NW * MINB + 3) / 4;
    static constexpr int MAXREG = (16384 / (WPS * 32) / 8 * 8) > 255 ? 255 : (16384 / (WPS * 32) / 8 * 8);
};

template <int N, class PHYS, bool CFL>
__global__ void __launch_bounds__((XRotCfg<N, PHYS>::NT), (XRotCfg<N, PHYS>::MINB)) __maxnreg__((XRotCfg<N, PHYS>::MAXREG))
xfused_rot_kernel(const __grid_constant__ XFusedParams p, int n_outer, int pairs_per_plane, int stagger_ns, int nsm) {
    using Cfg = XRotCfg<N, PHYS>;
    constexpr int NI = Cfg::NI, NO = Cfg::NO, NW = Cfg::NW, R0 = Cfg::R0, Q0 = N / R0, NT = Cfg::NT;
    constexpr int NR2 = NO > NW ? NO - NW : 0;          // pencils of the second forward round (warps 0 .. NR2-1)
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cplx* tile = reinterpret_cast<cplx*>(smem_raw);
    const int w = threadIdx.x >> 5, a = threadIdx.x & 31;
    const cplx* __restrict__ tw = p.tw;
    const int kn = p.kn;
    const int total = n_outer * pairs_per_plane;          // line pairs of this launch (checked < 2^31 by the launcher)
    const cplx w0 = (a != 0) ? DDL_LDG(&tw[a]) : mk(1.0, 0.0);
    double m0 = 0.0, m1 = 0.0;

    auto offset = [&](int q) -> long long {
        const int by = q / pairs_per_plane, bx = q - by * pairs_per_plane;
        return (long long)by * p.s_outer + (long long)(2 * bx) * p.pitch;
    };
    // inverse transform of pencil f of the pair at `off`: Hermitian pack from global -> stage 0 -> slot f -> stage 1
    auto inverse = [&](int f, long long off) {
        const cplx* __restrict__ A = p.in[f] + off;
        const cplx* __restrict__ B = A + p.pitch;
        cplx v[R0];
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
        xdft<R0, +1>(v);
        if (a != 0) xtwiddle<R0, true>(v, conj(w0));
        cplx* T = tile + f * N;
        const int sb = xsw<N>(a);
#pragma unroll
        for (int r = 0; r < R0; ++r) T[sb ^ xsw<N>(r * Q0)] = v[xreg<R0>(r)];
        __syncwarp();
        xstage<N, 1, +1, false, 32>(T, a, tw);
    };
    // forward transform of pencil f: stage 1 (shared) -> stage 0 -> registers -> shuffle unpack -> retained modes to global
    auto forward = [&](int f, long long off) {
        cplx* T = tile + f * N;
        xstage<N, 1, -1, true, 32>(T, a, tw);
        __syncwarp();
        cplx v[R0];
        const int sb = xsw<N>(a);
#pragma unroll
        for (int j = 0; j < R0; ++j) v[j] = T[sb ^ xsw<N>(j * Q0)];
        if (a != 0) xtwiddle<R0, false>(v, w0);
        xdft<R0, -1>(v);
        cplx* __restrict__ dst = p.out[f] + off;
        const int src = (32 - a) & 31;
        const int rmax = (kn - 1) / Q0;
        const double h = 0.5 * p.scale;
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
    };

    // Experiment knob (ddl_set_option("persist_stagger_ns")): resident CTAs that start together and do identical work per pair
    // could stay in lock step (all in the shared-memory-bound product phase, then all in the FP64-bound butterflies); the k-th
    // CTA of an SM (blockIdx / #SMs) can be started k * stagger_ns late.  Measured: no effect.
    if (stagger_ns > 0) {
        for (int k = blockIdx.x / nsm; k > 0; --k) __nanosleep(stagger_ns);
    }
    int q = blockIdx.x;
    if (q < total && w < NI) inverse(w, offset(q));          // prologue: every inverse pencil of the first pair
#pragma unroll 1
    for (; q < total; q += gridDim.x) {
        __syncthreads();                                     // all inverse pencils of pair q are in the tile
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
        // forward: pencil w in slot w (round 1), then -- warps 0 .. NR2-1 only -- pencil 6 + w (round 2): one call site, two trips
        {
            const long long off = offset(q);
#pragma unroll 1
            for (int f = w; f < NO; f += NW) forward(f, off);
        }
        // slot w belongs to warp w again: start the next pair now.  The warps without a second-round pencil get here while the
        // others are still in round 2, which is the point.
        if (q + (int)gridDim.x < total && w < NI) inverse(w, offset(q + gridDim.x));
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

// returns 0 on success, 1 if this (N, physics, mask) has no such kernel (the caller uses the one-shot kernel)
template <int N, class PHYS, bool CFL>
int launch_xfused_rot(const XFusedParams& p, int n_outer, ddl_stream_t stream) {
    using Cfg = XRotCfg<N, PHYS>;
    if constexpr (!Cfg::ok) {
        return 1;
    } else {
        if (p.n_lines & 1) return 1;
        auto kern = xfused_rot_kernel<N, PHYS, CFL>;
        static DeviceOnce once;
        const int ctas = once.get([&]() -> int {
            DDL_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM));
            DDL_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
            int dev = 0, sms = 0, per_sm = 0;
            DDL_CUDA_CHECK(cudaGetDevice(&dev));
            DDL_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
            DDL_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, Cfg::NT, Cfg::SMEM));
            if (per_sm < 1) { set_error("xfused_rot: kernel does not fit an SM"); return -2; }
            return sms * per_sm;
        });
        if (ctas < 0) return -2;
        const int pairs = p.n_lines / 2;
        const long long total = (long long)pairs * n_outer;
        if (total >= (1LL << 31) - ctas) return 1;
        const int grid = (int)(total < ctas ? total : ctas);
        if (grid < 1) return 0;
        prof_begin("x_fused", stream);
        kern<<<grid, Cfg::NT, Cfg::SMEM, stream>>>(p, n_outer, pairs, g_persist_stagger_ns, ctas / Cfg::MINB > 0 ? ctas / Cfg::MINB : 1);
        prof_end(stream);
        DDL_CUDA_CHECK(cudaGetLastError());
        return 0;
    }
}

}  // namespace ddl
#endif  // DDL_DEVICE_BUILD
