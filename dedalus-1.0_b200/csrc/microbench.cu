// Measured roofs for the FP64 vector pipe (include/ddl.h: ddl_measure_fp64).  bench.py reports the dominant kernel against
// max(bytes / HBM bandwidth, flops / FP64 peak); MEASURED_PEAKS.json carries no FP64 figure, so the peak is measured here,
// live, on the device the benchmark runs on: independent DFMA chains (2 flops each) and, because an FFT butterfly is
// mostly additions, the same with DADD (1 flop each, same issue slot).
#include "ddl_common.cuh"
#include "../../include/ddl.h"

namespace ddl {
#if DDL_DEVICE_BUILD
template <bool FMA>
__global__ void __launch_bounds__(256) fp64_rate_kernel(double* out, int iters, double seed) {
    double a[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = seed + threadIdx.x * 1e-3 + i;
    const double m = 1.0 - 1e-9, c = 1e-9;
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int i = 0; i < 8; ++i) a[i] = FMA ? fma(a[i], m, c) : a[i] + c;
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += a[i];
    if (s == 123.456) out[0] = s;      // never true: keeps the chains alive
}
#endif
}  // namespace ddl

extern "C" int ddl_measure_fp64(double* out2, void* stream) {
#if DDL_DEVICE_BUILD
    using namespace ddl;
    cudaStream_t st = (cudaStream_t)stream;
    int dev = 0, sms = 0;
    DDL_CUDA_CHECK(cudaGetDevice(&dev));
    DDL_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    double* sink = (double*)dev_alloc(16);
    if (!sink) { set_error("ddl_measure_fp64: allocation failed"); return -2; }
    const int blocks = sms * 8, iters = 4096;
    cudaEvent_t e0, e1;
    DDL_CUDA_CHECK(cudaEventCreate(&e0));
    DDL_CUDA_CHECK(cudaEventCreate(&e1));
    for (int which = 0; which < 2; ++which) {
        float best = 1e30f;
        for (int rep = 0; rep < 4; ++rep) {
            DDL_CUDA_CHECK(cudaEventRecord(e0, st));
            if (which == 0) fp64_rate_kernel<true><<<blocks, 256, 0, st>>>(sink, iters, 1.0);
            else fp64_rate_kernel<false><<<blocks, 256, 0, st>>>(sink, iters, 1.0);
            DDL_CUDA_CHECK(cudaEventRecord(e1, st));
            DDL_CUDA_CHECK(cudaEventSynchronize(e1));
            float ms = 0.f;
            DDL_CUDA_CHECK(cudaEventElapsedTime(&ms, e0, e1));
            if (rep > 0 && ms < best) best = ms;
        }
        const double instr = (double)blocks * 256.0 * iters * 32.0;      // per-thread FP64 instructions
        out2[which] = instr * (which == 0 ? 2.0 : 1.0) / (best * 1e-3) / 1e12;   // TFLOP/s
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    dev_free(sink);
    DDL_CUDA_CHECK(cudaGetLastError());
    return 0;
#else
    (void)out2; (void)stream;
    ddl::set_error("ddl_measure_fp64: device build only");
    return -1;
#endif
}
