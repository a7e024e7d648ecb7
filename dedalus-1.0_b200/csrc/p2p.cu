// Peer-to-peer slab exchange over NVLink: CUDA-IPC mapped arenas, copy-engine pushes, flag
// signalling (include/ddl.h: ddl_p2p_*).
//
// FFTW-MPI's global transpose (dedalus/utils/fftw/_fftw.pyx:272-304, one blocking MPI all-to-all
// inside every fftw_execute) becomes, per field: the z-/y-pass kernel writes peer-blocked
// pencils into this rank's arena; a communication stream waits for that kernel, pushes block s
// straight into rank s's arena with cudaMemcpyAsync (copy engines: no SM is taken from the
// passes that run meanwhile) and then raises this rank's arrival flag in every peer's arena; the
// consumer pass is preceded by a one-warp kernel that spins until all peers' flags have reached
// the exchange's sequence number.  No host synchronisation anywhere.
//
// Why arrival flags alone are enough (no "buffer free" credits): inside the RHS pipeline the
// region of a peer's buffer this rank overwrites in exchange e is exactly the block that peer
// sent to THIS rank in the previous exchange that used the buffer, and this rank has waited for
// that block's arrival flag earlier in its own program order (multi-GPU section).
#include <cstdint>
#include <cstring>
#include <vector>

#include "../../include/ddl.h"
#include "ddl_common.cuh"

namespace ddl {
// ddl_set_option("p2p_timeout_s", s): how long a consumer pass waits for a peer's arrival flag before it traps;
// 0 = wait for ever, like the blocking MPI all-to-all it replaces (_fftw.pyx:272-304).  A rank may legitimately be late by
// minutes (rank-0-only analysis or I/O, a snapshot write between two RHS evaluations, a debugger, first-call lazy set-up).
int g_p2p_timeout_s = 600;
}

#if DDL_DEVICE_BUILD
namespace ddl {

__device__ __forceinline__ unsigned long long p2p_now_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

__global__ void p2p_signal_kernel(unsigned* const* flags, int n, int slot, unsigned value) {
    const int i = threadIdx.x;
    if (i < n) {
        __threadfence_system();
        volatile unsigned* f = flags[i] + slot;
        *f = value;
        __threadfence_system();
    }
}

// spin until every watched flag has reached `value` (monotone sequence numbers; the
// comparison is wrap-safe); after timeout_ns of wall time (globaltimer: independent of the SM clock; 0 = never) trap, so that
// a lost peer fails loudly instead of hanging
__global__ void p2p_wait_kernel(const unsigned* flags, int n, int skip, unsigned value, unsigned long long timeout_ns) {
    const int i = threadIdx.x;
    if (i < n && i != skip) {
        const volatile unsigned* f = flags + i;
        const unsigned long long t0 = p2p_now_ns();
        while ((int)(*f - value) < 0) {
            if (timeout_ns && p2p_now_ns() - t0 > timeout_ns) {
                printf("ddl p2p: rank flag %d stuck at %u waiting for %u\n", i, *f, value);
                __trap();
            }
            __nanosleep(200);
        }
    }
    __threadfence_system();
}

// ---- all-to-all blocks pushed by a small SM kernel ---------------------------------------------------------------------
// Entry e: nrows[e] rows of row_bytes[e] bytes (a multiple of 16) from src[e] to dst[e] (a peer's arena, mapped through CUDA
// IPC), both with the same row pitch.  A few CTAs with several 16-byte loads in flight per thread saturate the NVLink ports of
// this GPU (posted stores, no acknowledgement to wait for), and their footprint -- no shared memory, <= 32 registers -- lets
// them share every SM with whatever HBM / FP64-bound pass runs on the other stream.  That is the point: the copy engines lose
// two thirds of their rate while such a pass runs (profiles/p2p_contention.py), and a pass that stores to the peers itself
// (ddl_slab_*_peer) turns into an NVLink-bound kernel that occupies the whole GPU.
// Footprint, measured (profiles/r2/slab_sweep.md): the strided passes fill the register file exactly (2 x 512 threads x 64
// registers), so ANY co-resident CTA evicts one of their two CTAs from its SM.  Many small push CTAs therefore halve the
// passes' occupancy everywhere; a few LARGE ones (1024 threads, half a register file each) cost one CTA slot on a few SMs
// only -- the NCCL recipe of giving a handful of SMs to communication -- and still keep 4 x 16 B per thread in flight.
#define DDL_PUSH_MAX 64
#define DDL_PUSH_THREADS 1024
#define DDL_PUSH_UNIT (DDL_PUSH_THREADS * 64)      // bytes of one work unit (one CTA iteration): 4 vectors of 16 B per thread
struct PushTab {
    const char* src[DDL_PUSH_MAX];
    char* dst[DDL_PUSH_MAX];
    long long row_bytes[DDL_PUSH_MAX], pitch[DDL_PUSH_MAX];
    int unit0[DDL_PUSH_MAX + 1];       // first work unit of entry e (prefix sums); unit0[n] = total
    int upr[DDL_PUSH_MAX];             // units per row
    int n;
};

__global__ void __launch_bounds__(DDL_PUSH_THREADS) p2p_push_kernel(const __grid_constant__ PushTab t) {
    const int total = t.unit0[t.n];
    for (int u = blockIdx.x; u < total; u += gridDim.x) {
        int e = 0;
        while (u >= t.unit0[e + 1]) ++e;                 // <= 64 entries: a short scan per 16 KB moved
        const int v = u - t.unit0[e];
        const int row = v / t.upr[e], part = v % t.upr[e];
        const long long off = (long long)row * t.pitch[e] + (long long)part * DDL_PUSH_UNIT;
        long long left = t.row_bytes[e] - (long long)part * DDL_PUSH_UNIT;
        const int nvec = (int)((left < DDL_PUSH_UNIT ? left : DDL_PUSH_UNIT) >> 4);
        const uint4* __restrict__ s4 = reinterpret_cast<const uint4*>(t.src[e] + off);
        uint4* __restrict__ d4 = reinterpret_cast<uint4*>(t.dst[e] + off);
        // four independent loads in flight per thread, then four stores
        uint4 r[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int i = threadIdx.x + k * DDL_PUSH_THREADS;
            if (i < nvec) r[k] = __ldg(s4 + i);
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int i = threadIdx.x + k * DDL_PUSH_THREADS;
            if (i < nvec) d4[i] = r[k];
        }
    }
}

// ---- the same table driven by the bulk-copy engine (TMA): ONE thread per CTA ---------------------------------------------
// global -> shared (cp.async.bulk, mbarrier complete_tx) -> peer global (cp.async.bulk.global.shared::cta, bulk groups), NB
// buffers of UB bytes in flight per CTA.  No registers to speak of, 32 threads: what such a CTA takes from an SM is 64 KB of
// shared memory, so it co-resides with two strided-pass CTAs (2 x 64 KB) and costs the fused x pass one of its three CTAs on
// the few SMs that host one.  The SM's load/store units and register file stay with the compute pass.
#define DDL_TMA_UB 16384
#define DDL_TMA_NB 4
struct TmaPushTab {                    // PushTab with units of DDL_TMA_UB bytes
    const char* src[DDL_PUSH_MAX];
    char* dst[DDL_PUSH_MAX];
    long long row_bytes[DDL_PUSH_MAX], pitch[DDL_PUSH_MAX];
    int unit0[DDL_PUSH_MAX + 1];
    int upr[DDL_PUSH_MAX];
    int n;
};

__device__ __forceinline__ unsigned tp_smem(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(32) p2p_tma_push_kernel(const __grid_constant__ TmaPushTab t) {
    extern __shared__ __align__(128) unsigned char tp_buf[];
    __shared__ __align__(8) unsigned long long full[DDL_TMA_NB];
    if (threadIdx.x != 0) return;
    for (int b = 0; b < DDL_TMA_NB; ++b)
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(tp_smem(&full[b])) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    const int total = t.unit0[t.n];
    // unit u -> (src, dst, bytes)
    auto locate = [&](int u, const char*& s, char*& d, unsigned& bytes) {
        int e = 0;
        while (u >= t.unit0[e + 1]) ++e;
        const int v = u - t.unit0[e];
        const int row = v / t.upr[e], part = v % t.upr[e];
        const long long off = (long long)row * t.pitch[e] + (long long)part * DDL_TMA_UB;
        const long long left = t.row_bytes[e] - (long long)part * DDL_TMA_UB;
        bytes = (unsigned)(left < DDL_TMA_UB ? left : DDL_TMA_UB);
        s = t.src[e] + off;
        d = t.dst[e] + off;
    };
    auto load = [&](int k, int u) {        // k-th unit of this CTA into buffer k % NB
        const char* s; char* d; unsigned bytes;
        locate(u, s, d, bytes);
        const int b = k % DDL_TMA_NB;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(tp_smem(&full[b])), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(tp_smem(tp_buf + (size_t)b * DDL_TMA_UB)), "l"(s), "r"(bytes), "r"(tp_smem(&full[b])) : "memory");
    };
    int nk = 0;                            // units of this CTA
    for (int u = blockIdx.x; u < total; u += gridDim.x) ++nk;
    for (int k = 0; k < nk && k < DDL_TMA_NB; ++k) load(k, blockIdx.x + k * gridDim.x);
    for (int k = 0; k < nk; ++k) {
        const int b = k % DDL_TMA_NB;
        const unsigned parity = (unsigned)(k / DDL_TMA_NB) & 1u;
        unsigned ok;
        do {
            asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                         : "=r"(ok) : "r"(tp_smem(&full[b])), "r"(parity) : "memory");
        } while (!ok);
        const char* s; char* d; unsigned bytes;
        locate(blockIdx.x + k * gridDim.x, s, d, bytes);
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                     ::"l"(d), "r"(tp_smem(tp_buf + (size_t)b * DDL_TMA_UB)), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        // the buffer of the PREVIOUS store may be refilled once that store has read it: all but the latest group done reading
        if (k >= 1 && k - 1 + DDL_TMA_NB < nk) {
            asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
            load(k - 1 + DDL_TMA_NB, blockIdx.x + (k - 1 + DDL_TMA_NB) * gridDim.x);
        }
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");      // every store complete before the kernel (and the flag behind it) ends
}

int g_push_tma = 1;                    // ddl_set_option("push_tma", 0 / 1): bulk-copy engine (default) or the thread-copy kernel above

}  // namespace ddl
#endif

#define DDL_P2P_RING 64
#define DDL_P2P_HDR 1024

struct ddl_p2p {
    int nranks = 1, rank = 0;
    void* base = nullptr;             // [flags: DDL_P2P_HDR bytes][data]; the flags sit at the same
    size_t bytes = 0;                 // offset in every rank's arena (the data sizes differ per rank)
    std::vector<void*> peer;          // peer[r]: rank r's arena in this process's address space
    unsigned** d_flagptrs = nullptr;  // device array: flag base of every rank
    unsigned seq = 0;
    bool copied[64] = {};             // ring: exchange used copy-engine pushes (its local block has an event)
#if DDL_DEVICE_BUILD
    cudaStream_t comm = nullptr;
    cudaEvent_t ready[DDL_P2P_RING], self[DDL_P2P_RING];
#endif
};

#if DDL_DEVICE_BUILD
using namespace ddl;

extern "C" int ddl_p2p_create(ddl_p2p** out, int nranks, int rank, size_t data_bytes, char* handle_out64) {
    ddl_p2p* c = new ddl_p2p();
    c->nranks = nranks; c->rank = rank;
    c->bytes = DDL_P2P_HDR + data_bytes;
    DDL_CUDA_CHECK(cudaMalloc(&c->base, c->bytes));
    DDL_CUDA_CHECK(cudaMemset(c->base, 0, c->bytes));
    cudaIpcMemHandle_t h;
    DDL_CUDA_CHECK(cudaIpcGetMemHandle(&h, c->base));
    static_assert(sizeof(h) == 64, "IPC handle size");
    memcpy(handle_out64, &h, 64);
    DDL_CUDA_CHECK(cudaStreamCreateWithFlags(&c->comm, cudaStreamNonBlocking));
    for (int i = 0; i < DDL_P2P_RING; ++i) {
        DDL_CUDA_CHECK(cudaEventCreateWithFlags(&c->ready[i], cudaEventDisableTiming));
        DDL_CUDA_CHECK(cudaEventCreateWithFlags(&c->self[i], cudaEventDisableTiming));
    }
    c->peer.assign(nranks, nullptr);
    c->peer[rank] = c->base;
    DDL_CUDA_CHECK(cudaDeviceSynchronize());
    *out = c;
    return 0;
}

// handles: nranks x 64 bytes, rank-major (all-gathered by the host layer)
extern "C" int ddl_p2p_connect(ddl_p2p* c, const char* handles) {
    for (int r = 0; r < c->nranks; ++r) {
        if (r == c->rank) continue;
        cudaIpcMemHandle_t h;
        memcpy(&h, handles + 64 * r, 64);
        DDL_CUDA_CHECK(cudaIpcOpenMemHandle(&c->peer[r], h, cudaIpcMemLazyEnablePeerAccess));
    }
    std::vector<unsigned*> fp(c->nranks);
    for (int r = 0; r < c->nranks; ++r) fp[r] = (unsigned*)c->peer[r];
    DDL_CUDA_CHECK(cudaMalloc(&c->d_flagptrs, sizeof(unsigned*) * c->nranks));
    DDL_CUDA_CHECK(cudaMemcpy(c->d_flagptrs, fp.data(), sizeof(unsigned*) * c->nranks, cudaMemcpyHostToDevice));
    return 0;
}

// base of the data part; the offsets given to ddl_p2p_exchange are relative to it
extern "C" void* ddl_p2p_base(ddl_p2p* c) { return (char*)c->base + DDL_P2P_HDR; }

// One exchange: after everything enqueued so far on `stream`, copy n blocks
// (this arena + src_off[i]) -> (rank dst_rank[i]'s arena + dst_off[i]), nbytes[i] each, on the
// communication stream, then publish the exchange's sequence number to every peer.
// Returns the sequence number (> 0) to hand to ddl_p2p_wait, or a negative error code.
extern "C" long long ddl_p2p_exchange(ddl_p2p* c, int n, const int* dst_rank, const int64_t* src_off, const int64_t* dst_off,
                                      const int64_t* nbytes, void* stream) {
    const unsigned seq = ++c->seq;
    const int slot = seq % DDL_P2P_RING;
    c->copied[slot] = true;
    cudaStream_t st = (cudaStream_t)stream;
    DDL_CUDA_CHECK(cudaEventRecord(c->ready[slot], st));
    DDL_CUDA_CHECK(cudaStreamWaitEvent(c->comm, c->ready[slot], 0));
    // the local block first (its consumer waits for `self`), then the peers
    for (int pass = 0; pass < 2; ++pass) {
        for (int i = 0; i < n; ++i) {
            const bool local = dst_rank[i] == c->rank;
            if (local != (pass == 0) || nbytes[i] <= 0) continue;
            DDL_CUDA_CHECK(cudaMemcpyAsync((char*)c->peer[dst_rank[i]] + DDL_P2P_HDR + dst_off[i],
                                           (const char*)c->base + DDL_P2P_HDR + src_off[i],
                                           (size_t)nbytes[i], cudaMemcpyDeviceToDevice, c->comm));
        }
        if (pass == 0) DDL_CUDA_CHECK(cudaEventRecord(c->self[slot], c->comm));
    }
    p2p_signal_kernel<<<1, 32, 0, c->comm>>>(c->d_flagptrs, c->nranks, c->rank, seq);
    DDL_CUDA_CHECK(cudaGetLastError());
    return (long long)seq;
}

// Data base of rank r's arena as mapped in THIS process (for the peer-store passes' tables).
extern "C" void* ddl_p2p_peer_base(ddl_p2p* c, int r) { return (char*)c->peer[r] + DDL_P2P_HDR; }

// Exchange whose data movement was done by the producing kernel itself (peer stores fused into
// the pass, ddl_slab_zinv_peer / ddl_slab_yfwd_peer): behind the work on `stream`, publish the
// next sequence number to every peer.  Returns the sequence number for ddl_p2p_wait.
extern "C" long long ddl_p2p_signal(ddl_p2p* c, void* stream) {
    const unsigned seq = ++c->seq;
    c->copied[seq % DDL_P2P_RING] = false;
    p2p_signal_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(c->d_flagptrs, c->nranks, c->rank, seq);
    DDL_CUDA_CHECK(cudaGetLastError());
    return (long long)seq;
}

// Blocks pushed by the SM kernel above instead of the copy engines: after everything enqueued so far on `stream`, entry i copies
// nrows[i] rows of row_bytes[i] bytes (multiples of 16) with pitch[i] bytes between rows from (this arena + src_off[i]) to
// (rank dst_rank[i]'s arena + dst_off[i]); own-rank entries are plain local copies.  `ctas` CTAs (0: 32).  With `publish`
// the exchange's sequence number is raised in every peer behind the copies and returned (> 0) for ddl_p2p_wait; without, 0
// is returned and a later call publishes (chunked pushes: one flag per exchange, after its last chunk).
extern "C" long long ddl_p2p_push(ddl_p2p* c, int n, const int* dst_rank, const int64_t* src_off, const int64_t* dst_off,
                                  const int64_t* row_bytes, const int64_t* nrows, const int64_t* pitch, int ctas, int publish,
                                  void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    for (int i0 = 0; i0 < n; i0 += DDL_PUSH_MAX) {
        PushTab t;
        t.n = 0;
        int units = 0;
        for (int i = i0; i < n && t.n < DDL_PUSH_MAX; ++i) {
            if (row_bytes[i] <= 0 || nrows[i] <= 0) continue;
            if ((row_bytes[i] | pitch[i] | src_off[i] | dst_off[i]) & 15) { set_error("ddl_p2p_push: sizes and offsets must be multiples of 16"); return -1; }
            const int e = t.n++;
            t.src[e] = (const char*)c->base + DDL_P2P_HDR + src_off[i];
            t.dst[e] = (char*)c->peer[dst_rank[i]] + DDL_P2P_HDR + dst_off[i];
            t.row_bytes[e] = row_bytes[i];
            t.pitch[e] = pitch[i];
            t.upr[e] = (int)((row_bytes[i] + DDL_PUSH_UNIT - 1) / DDL_PUSH_UNIT);
            t.unit0[e] = units;
            units += t.upr[e] * (int)nrows[i];
        }
        t.unit0[t.n] = units;
        if (units == 0) continue;
        const int want = ctas > 0 ? ctas : 32;
        if (g_push_tma) {
            TmaPushTab q;
            q.n = t.n;
            int tu = 0;
            for (int e = 0; e < t.n; ++e) {
                q.src[e] = t.src[e]; q.dst[e] = t.dst[e]; q.row_bytes[e] = t.row_bytes[e]; q.pitch[e] = t.pitch[e];
                q.upr[e] = (int)((t.row_bytes[e] + DDL_TMA_UB - 1) / DDL_TMA_UB);
                q.unit0[e] = tu;
                tu += q.upr[e] * ((t.unit0[e + 1] - t.unit0[e]) / t.upr[e]);
            }
            q.unit0[q.n] = tu;
            static DeviceOnce once;
            if (once.get([&]() -> int {
                    DDL_CUDA_CHECK(cudaFuncSetAttribute(p2p_tma_push_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, DDL_TMA_UB * DDL_TMA_NB));
                    return 1;
                }) < 0) return -2;
            prof_begin("push", st);
            p2p_tma_push_kernel<<<tu < want ? tu : want, 32, DDL_TMA_UB * DDL_TMA_NB, st>>>(q);
            prof_end(st);
        } else {
            prof_begin("push", st);
            p2p_push_kernel<<<units < want ? units : want, DDL_PUSH_THREADS, 0, st>>>(t);
            prof_end(st);
        }
        DDL_CUDA_CHECK(cudaGetLastError());
    }
    if (!publish) return 0;
    const unsigned seq = ++c->seq;
    c->copied[seq % DDL_P2P_RING] = false;
    p2p_signal_kernel<<<1, 32, 0, st>>>(c->d_flagptrs, c->nranks, c->rank, seq);
    DDL_CUDA_CHECK(cudaGetLastError());
    return (long long)seq;
}

// Make `stream` wait until exchange `seq` has fully arrived in this rank's arena.
extern "C" int ddl_p2p_wait(ddl_p2p* c, long long seq, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    const int slot = (unsigned)seq % DDL_P2P_RING;
    if (c->copied[slot]) DDL_CUDA_CHECK(cudaStreamWaitEvent(st, c->self[slot], 0));
    const unsigned* flags = (const unsigned*)c->base;
    p2p_wait_kernel<<<1, 32, 0, st>>>(flags, c->nranks, c->rank, (unsigned)seq,
                                      (unsigned long long)(ddl::g_p2p_timeout_s > 0 ? ddl::g_p2p_timeout_s : 0) * 1000000000ULL);
    DDL_CUDA_CHECK(cudaGetLastError());
    return 0;
}

extern "C" int ddl_p2p_destroy(ddl_p2p* c) {
    if (!c) return 0;
    cudaDeviceSynchronize();
    for (int r = 0; r < c->nranks; ++r)
        if (r != c->rank && c->peer[r]) cudaIpcCloseMemHandle(c->peer[r]);
    if (c->d_flagptrs) cudaFree(c->d_flagptrs);
    if (c->comm) cudaStreamDestroy(c->comm);
    for (int i = 0; i < DDL_P2P_RING; ++i) { cudaEventDestroy(c->ready[i]); cudaEventDestroy(c->self[i]); }
    if (c->base) cudaFree(c->base);
    delete c;
    return 0;
}

#else   // host emulation: there is no peer memory; the tests exchange through torch.distributed

extern "C" int ddl_p2p_create(ddl_p2p**, int, int, size_t, char*) { ddl::set_error("ddl_p2p needs the CUDA build"); return -1; }
extern "C" int ddl_p2p_connect(ddl_p2p*, const char*) { return -1; }
extern "C" void* ddl_p2p_base(ddl_p2p*) { return nullptr; }
extern "C" long long ddl_p2p_exchange(ddl_p2p*, int, const int*, const int64_t*, const int64_t*, const int64_t*, void*) { return -1; }
extern "C" int ddl_p2p_wait(ddl_p2p*, long long, void*) { return -1; }
extern "C" void* ddl_p2p_peer_base(ddl_p2p*, int) { return nullptr; }
extern "C" long long ddl_p2p_signal(ddl_p2p*, void*) { return -1; }
extern "C" long long ddl_p2p_push(ddl_p2p*, int, const int*, const int64_t*, const int64_t*, const int64_t*, const int64_t*, const int64_t*, int, int,
                                  void*) { return -1; }
extern "C" int ddl_p2p_destroy(ddl_p2p*) { return 0; }

#endif
