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
extern "C" int ddl_p2p_destroy(ddl_p2p*) { return 0; }

#endif
