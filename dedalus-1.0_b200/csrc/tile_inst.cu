// Instantiates the generic tile kernel for ONE transform length (compile with -DDDL_N=<N>);
// one translation unit per length so the build parallelises.  -DDDL_N=0 is the runtime-length
// instantiation (mixed radix, fft_core.cuh RtFac) that serves every length without a unit of its own.
#include "tile_kernel.cuh"
#include "fast_kernels.cuh"
#include "fast_staged.cuh"
#include "xfused_kernel.cuh"
#include "xfused_persist.cuh"
#include "xfused_rot.cuh"
#include "fast_two.cuh"

#ifndef DDL_N
#error "compile with -DDDL_N=<transform length>"
#endif
#define DDL_CAT2(a, b) a##b
#define DDL_CAT(a, b) DDL_CAT2(a, b)

namespace ddl {

int DDL_CAT(run_tile_, DDL_N)(int mode, int dir, int phys, const TileParams& p, int nthreads, ddl_stream_t s) {
    constexpr int N = DDL_N;
    switch (mode) {
        case TM_C2C:
            return dir < 0 ? launch_tile<N, TM_C2C, -1, PhysNone>(p, nthreads, s)
                           : launch_tile<N, TM_C2C, +1, PhysNone>(p, nthreads, s);
        case TM_C2R: return launch_tile<N, TM_C2R, +1, PhysNone>(p, nthreads, s);
        case TM_R2C: return launch_tile<N, TM_R2C, -1, PhysNone>(p, nthreads, s);
        case TM_FUSED:
            switch (phys) {
                case 0: return launch_tile<N, TM_FUSED, 0, Hydro2C>(p, nthreads, s);
                case 1: return launch_tile<N, TM_FUSED, 0, Bouss2C>(p, nthreads, s);
                case 2: return launch_tile<N, TM_FUSED, 0, MHD2C>(p, nthreads, s);
                case 3: return launch_tile<N, TM_FUSED, 0, Hydro3C>(p, nthreads, s);
                case 4: return launch_tile<N, TM_FUSED, 0, Bouss3C>(p, nthreads, s);
                case 5: return launch_tile<N, TM_FUSED, 0, MHD3C>(p, nthreads, s);
                // advective-form policies for non-solenoidal states (physics_ops.cuh AdvOf)
                case 6: return launch_tile<N, TM_FUSED, 0, Hydro2A>(p, nthreads, s);
                case 7: return launch_tile<N, TM_FUSED, 0, Bouss2A>(p, nthreads, s);
                case 8: return launch_tile<N, TM_FUSED, 0, MHD2A>(p, nthreads, s);
                case 9: return launch_tile<N, TM_FUSED, 0, Hydro3A>(p, nthreads, s);
                case 10: return launch_tile<N, TM_FUSED, 0, Bouss3A>(p, nthreads, s);
                case 11: return launch_tile<N, TM_FUSED, 0, MHD3A>(p, nthreads, s);
                // traceless-flux policies of the one-rank 3-D RHS (physics_ops.cuh)
                case 12: return launch_tile<N, TM_FUSED, 0, Hydro3T>(p, nthreads, s);
                case 13: return launch_tile<N, TM_FUSED, 0, Bouss3T>(p, nthreads, s);
                case 14: return launch_tile<N, TM_FUSED, 0, MHD3T>(p, nthreads, s);
            }
    }
    set_error("run_tile: bad mode/physics %d/%d", mode, phys);
    return -1;
}

// specialised fused x-pass of the 3-D RHS; returns 1 if this length / physics has none
int DDL_CAT(run_xfused_, DDL_N)(int phys, const XFusedParams& p, int n_outer, int variant, ddl_stream_t s) {
    constexpr int N = DDL_N;
    if constexpr (XFac<N>::ok) {
        switch (phys) {
            case 3: return launch_xfused<N, Hydro3C>(p, n_outer, variant, s);
            case 4: return launch_xfused<N, Bouss3C>(p, n_outer, variant, s);
            case 5: return launch_xfused<N, MHD3C>(p, n_outer, variant, s);
            // traceless-flux policies (what the one-rank ddl_rhs runs by default): the default CTA shape and its CFL-capture twin
            // only -- the opt-in launch variants were built and measured on the six-product policies and stay with them
            // (ddl_set_option("traceless_flux", 0) + "xfused_variant")
            case 12: return launch_xfused_basic<N, Hydro3T>(p, n_outer, s);
            case 13: return launch_xfused_basic<N, Bouss3T>(p, n_outer, s);
            case 14: return launch_xfused_basic<N, MHD3T>(p, n_outer, s);
            // advective-form policies (states that are not solenoidal, e.g. the reference's own 3-D turb_new fields): the two
            // whose field counts fit the kernel's pointer tables; Boussinesq (13 products) stays on the generic tile kernel
            case 9: return launch_xfused_basic<N, Hydro3A>(p, n_outer, s);
            case 11: return launch_xfused_basic<N, MHD3A>(p, n_outer, s);
        }
    }
    return 1;
}

// two-stage strided pass for plain rows (fast_two.cuh); returns 1 if this length has none
template <int N, bool OK = TwoFac<N>::ok> struct TwoStrided {
    static int run(int, int, const TwoParams&, int, int, const char*, ddl_stream_t) { return 1; }
};
template <int N> struct TwoStrided<N, true> {
    static int run(int dir, int variant, const TwoParams& p, int nf, int n_outer, const char* name, ddl_stream_t s) {
        if (variant == 2)      // twiddles from the table
            return dir < 0 ? launch_strided_two<N, -1, true>(p, nf, n_outer, name, s) : launch_strided_two<N, +1, true>(p, nf, n_outer, name, s);
        return dir < 0 ? launch_strided_two<N, -1, false>(p, nf, n_outer, name, s) : launch_strided_two<N, +1, false>(p, nf, n_outer, name, s);
    }
};
int DDL_CAT(run_two_strided_, DDL_N)(int dir, int variant, const TwoParams& p, int nf, int n_outer, const char* name, ddl_stream_t s) {
    return TwoStrided<DDL_N>::run(dir, variant, p, nf, n_outer, name, s);
}

#if DDL_DEVICE_BUILD
// specialised strided pass; returns 1 if this length has no fast kernel (caller falls back)
template <int N> struct FastStrided {      // a template, so that the discarded branch is never instantiated (Fac<0>)
    static int run(int dir, const FastParams& p, int nf, int n_outer, const char* name, ddl_stream_t s) {
        if constexpr (Fac<N>::S >= 2) {
            return dir < 0 ? launch_strided_fast<N, -1>(p, nf, n_outer, name, s) : launch_strided_fast<N, +1>(p, nf, n_outer, name, s);
        } else {
            return 1;
        }
    }
};
template <> struct FastStrided<0> {
    static int run(int, const FastParams&, int, int, const char*, ddl_stream_t) { return 1; }
};
int DDL_CAT(run_fast_strided_, DDL_N)(int dir, const FastParams& p, int nf, int n_outer, const char* name, ddl_stream_t s) {
    return FastStrided<DDL_N>::run(dir, p, nf, n_outer, name, s);
}
#endif

}  // namespace ddl
