/* ddl.h -- C ABI of the B200-native pseudospectral RHS + timestep library (libddl_b200.so).
 *
 * Drop-in boundary for the hot path of jsoishi/dedalus-1.0 (paths below are relative to the
 * reference tree).  Every entry point takes raw DEVICE pointers (storage is owned by the
 * caller -- torch tensors in the Python host layer), enqueues work on the given cudaStream_t
 * and returns immediately: 0 on success, a negative code on error with the text available
 * from ddl_last_error().  Nothing allocates per call; nothing synchronises (ddl_sync only).
 *
 * Layouts (dedalus/data_objects/representations.py:84-93,174-176):
 *   3-D  x-space double [nz][ny][nx]            k-space complex128 [ny][nz][nx/2+1]
 *   2-D  x-space double [ny][nx]                k-space complex128 [nx/2+1][ny]
 * Forward transforms are normalised by 1/N_total, inverse unnormalised (:321,:329,:333).
 */
#ifndef DDL_H
#define DDL_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct ddl_plan ddl_plan;

/* physics ids (dedalus/physics/physics.py:419,612,724) */
enum { DDL_HYDRO = 0, DDL_BOUSSINESQ = 1, DDL_MHD = 2,
       /* The same right-hand sides for a state that is NOT solenoidal.  The fused pipeline evaluates the
        * nonlinear terms in conservative form (d_j(u_i u_j), ...), which equals the reference's advective
        * form (physics.py:197-228 XgradY) only where div u = div B = 0; the reference evaluates u.grad u
        * whatever the state.  The *_ADV policies add the missing u_i div u (- B_i div B / 4 pi rho0,
        * T div u) products: every `state` pointer list then carries, AFTER the state components, one
        * scratch k-space array (two for MHD) that receives the divergence spectra.  ddl_rhs,
        * ddl_reduce_max_square and the ddl_slab_* phases accept them (ddl_slab_theta fills the scratch
        * arrays); ddl_rhs_stage / ddl_slab_assemble_stage do not. */
       DDL_HYDRO_ADV = 3, DDL_BOUSSINESQ_ADV = 4, DDL_MHD_ADV = 5 };

/* stage-update kinds (dedalus/time_stepping/forward_step_cy_3d.pyx:17,34,63,95) */
enum { DDL_EULER = 0, DDL_ETD1 = 1, DDL_ETD2RK1 = 2, DDL_ETD2RK2 = 3 };

/* ddl_rhs flags */
enum {
    DDL_RHS_ZERO_FILL = 1,      /* also write zeros to the masked-out modes of deriv */
    DDL_RHS_DEALIAS_STATE = 2   /* zero the masked-out modes of the input state in place
                                   (what IncompressibleMHD does to it, physics.py:797-815 via
                                   representations.py:353) */
};

/* stage-update flags */
enum {
    DDL_STAGE_RETAINED_ONLY = 1 /* every operand is known to vanish outside the dealias mask:
                                   sweep the retained modes only (the update maps 0 -> 0 there) */
};

typedef struct ddl_phys_params {
    double rho0;        /* physics.py:753 */
    double g, alpha_t, beta;  /* physics.py:643-645 */
    int boussinesq_dir; /* component index of boussinesq_direction: x=0, y=1, z=2 (physics.py:646) */
    int reserved;
} ddl_phys_params;

/* Plan = wavenumbers, dealias mask, twiddles, index tables.  Replaces
 * fftw.create_data / fftw.rPlan (dedalus/utils/fftw/_fftw.pyx:81-185,246-309),
 * FourierRepresentation._setup_k and set_dealiasing (representations.py:204-233,359-382).
 *   shape_x  : x-space shape, (nz,ny,nx) or (ny,nx); every length in 2..2048 with prime factors <= 64 (FFTW, the
 *              reference's backend, takes any length; powers of two >= 8 run the specialised kernels)
 *   kx,ky,kz : the wavenumber VALUES along each axis exactly as the host computed them
 *              (kx: nx/2+1 entries, ky: ny, kz: nz; kz NULL in 2-D)
 *   keepx..  : 1 where the mode survives the dealias mask on that axis, 0 where it is zeroed
 *              (must be of the form |index| <= m on every axis)                              */
int ddl_plan_create(ddl_plan** out, int ndim, const int64_t* shape_x,
                    const double* kx, const double* ky, const double* kz,
                    const uint8_t* keepx, const uint8_t* keepy, const uint8_t* keepz);
int ddl_plan_destroy(ddl_plan* plan);

/* Same, for rank `rank` of an `nranks`-way slab decomposition (3-D only; nranks a power of two
 * dividing nz and ny).  Partition of the reference / FFTW-MPI (_fftw.pyx:114-148,
 * representations.py:180-186,231-233): x-space is split along z (nz/nranks planes per rank),
 * k-space along ky (ny/nranks rows per rank).  kx,ky,kz,keep* are the GLOBAL axis arrays; every
 * k-space pointer later passed with this plan is the rank's LOCAL slab [ny/nranks][nz][nx/2+1],
 * every x-space pointer its local [nz/nranks][ny][nx].  ddl_dealias / ddl_deriv / ddl_stage /
 * ddl_rk4_stage / ddl_cn_step work on the local slab unchanged; transforms and the RHS go
 * through the ddl_slab_* phases below with the exchange done by the caller between them.
 * ky_layout 0 = block slabs (the reference's: rank r owns ky rows [r*ny/P, (r+1)*ny/P));
 * ky_layout 1 = cyclic (rank r owns rows r, r+P, ...): same local shape, balanced under 2/3
 * dealiasing, where block slabs leave the middle ranks without retained modes. */
int ddl_plan_create_slab(ddl_plan** out, int ndim, const int64_t* shape_x,
                         const double* kx, const double* ky, const double* kz,
                         const uint8_t* keepx, const uint8_t* keepy, const uint8_t* keepz,
                         int nranks, int rank, int ky_layout);

/* scratch requirement of ddl_rhs / ddl_forward / ddl_backward for n_in inverse and n_out
 * forward transforms in flight (the transforms use n_in = n_out = 1) */
size_t ddl_workspace_bytes(const ddl_plan* plan, int n_in, int n_out);
size_t ddl_rhs_workspace_bytes(const ddl_plan* plan, int physics);

/* representations.py:335-345 forward(): x -> k, normalised, transposed out, dealiased */
int ddl_forward(ddl_plan* plan, const double* x, void* k, void* work, size_t work_bytes, void* stream);
/* representations.py:347-357 backward(): dealias k IN PLACE, then k -> x (unnormalised) */
int ddl_backward(ddl_plan* plan, void* k, double* x, void* work, size_t work_bytes, void* stream);
/* dealias_cy_{2,3}d.pyx dealias_23 / representations.py:442-455 zero_nyquist, in place */
int ddl_dealias(ddl_plan* plan, void* k, void* stream);
/* dealias_23 with the reference's own signature (dealias_cy_2d.pyx:13-41, dealias_cy_3d.pyx:13-46), no plan: the k-space
 * array (shape kshape: (ky,kz,kx) or (kx,ky)), the wavenumber VALUES of every axis, the Nyquist values in k-space axis order, and
 * ky either per row (ky_dense = 0) or dense over (ky, kx) / (kx, ky) (ky_dense = 1: the kernels' second branch, taken by the
 * shearing box, whose ky drifts with kx and time).  Zero where any k >= 2/3 k_nyquist or k <= -2/3 k_nyquist. */
int ddl_dealias_array(int ndim, const int64_t* kshape, void* data, const double* kx, const double* ky, const double* kz,
                      int ky_dense, const double* knyquist, void* stream);
/* representations.py:144-166 __setitem__ copies the caller's array into the fixed buffer; here the caller's array is in HOST memory
 * and the buffer on the device.  For a spectrum whose modes outside the dealias mask are known to be zero only the retained
 * box has to cross PCIe: copy the nbox boxes (six int64 each: lo0, hi0, lo1, hi1, lo2, hi2, half-open, in elements) of a 3-D
 * array of extents shape[3] (2-D spectra: shape[0] = 1) and elem_bytes per element between dst and src, both in the same full
 * layout; to_device = 1 host -> device, 0 device -> host; asynchronous on `stream` (pinned host memory), nothing outside the
 * boxes is read or written. */
int ddl_copy_boxes(void* dst, const void* src, const int64_t* shape, int nbox, const int64_t* boxes, int elem_bytes,
                   int to_device, void* stream);
/* representations.py:419-425 deriv(): out = i * k_axis * in   (axis: 0=x, 1=y, 2=z) */
int ddl_deriv(ddl_plan* plan, const void* k_in, void* k_out, int axis, void* stream);

/* Shearing box (FourierShearRepresentation, representations.py:558-740): between the x pass and the y pass the
 * half-transformed line at y = j * dy is multiplied by exp(+i ((S kx) y) t) on the way to k-space and by its conjugate
 * on the way back (fwd_np / rev_np :700-740; FFTW route :667-698).  While enabled, ddl_forward / ddl_backward of this
 * plan apply that factor inside their x pass (no extra pass over memory).  The plan of a shearing box is created with a
 * ky mask that keeps EVERY row (all ones, Nyquist included): the sheared wavenumber ky - S kx t, its wrap and its 2/3 mask
 * depend on kx and time (:627-642) and belong to the caller (ddl_dealias_array); kx and kz are masked as usual.  Slab-decomposed
 * plans too: ddl_slab_xr2c / ddl_slab_xc2r apply the factor, the other phases take the generic kernels for such a plan. */
int ddl_set_shear(ddl_plan* plan, int enable, double shear_rate, double time, double dy);

/* physics.py:527-599 / 664-712 / 770-819: deriv = RHS(state), all pointers k-space arrays in
 * StateData insertion order (u_x,u_y[,u_z] then T or B_x,B_y[,B_z]) */
int ddl_rhs(ddl_plan* plan, int physics, const ddl_phys_params* params,
            void* const* state, void* const* deriv, void* work, size_t work_bytes,
            int flags, void* stream);

/* ---- 3-D pipeline phases (one rank: ddl_rhs / ddl_forward / ddl_backward chain them) --------
 * A transform is  z pass -> [exchange] -> y pass -> x pass  (inverse) and the mirror (forward);
 * FFTW-MPI hides the exchange inside fftw_execute (_fftw.pyx:232-234,272-304), here it is the
 * caller's: between the z and y passes block s of every "k-side" array goes to rank s and lands
 * as block r of the peer's "x-side" array (an all-to-all of contiguous blocks; sizes below).
 *   k-side array, per field: [peer s][cyl][nzl][CX]      x-side array: [cy][nzl][CX]
 *   b / c arrays (x-pass input / output): [nzl][ny][CX]  e array: [cyl][cz][CX]
 * cyl = retained ky rows owned by this rank, cy = all retained ky rows, cz = retained kz rows,
 * nzl = local z planes, CX = pitch of the retained kx axis.
 * ddl_slab_info fills out[17] = {nranks, rank, nzl, nyl, cyl, cy0, cy, cz, CX, nkx,
 *   k-side elements, x-side elements, b elements, e elements (per field), z0, ky0, ky_layout};
 * ddl_slab_rows fills cyl of every rank: rank r's block in an x-side array is rows
 * [sum_{q<r} cyl_q, +cyl_r), i.e. cyl_r*nzl*CX elements, and this rank sends cyl*nzl*CX
 * elements to every peer. */
int ddl_slab_info(const ddl_plan* plan, int64_t* out16);
int ddl_slab_rows(const ddl_plan* plan, int64_t* cyl_of_rank);
/* *_ADV physics only: theta_u = i k.u (theta_B = i k.B) into the scratch arrays that follow the state */
int ddl_slab_theta(ddl_plan* plan, int physics, void* const* state, void* stream);
/* inverse z pass of nf fields: local k slabs (retained modes only are read) -> k-side arrays */
int ddl_slab_zinv(ddl_plan* plan, int nf, void* const* k_in, void* const* kside_out, void* stream);
/* inverse y pass: x-side arrays -> b arrays */
int ddl_slab_yinv(ddl_plan* plan, int nf, void* const* xside_in, void* const* b_out, void* stream);
/* x pass with the real-space products of `physics` fused in (physics.py:197-228,309-354):
 * b arrays of the state components -> c arrays of the product fields, normalised by 1/N_total */
int ddl_slab_xfused(ddl_plan* plan, int physics, const ddl_phys_params* params, void* const* b_in,
                    void* const* c_out, void* stream);
/* plain x passes of the transform API: b -> x-space (unnormalised), x-space -> c (1/N_total) */
int ddl_slab_xc2r(ddl_plan* plan, const void* b_in, double* x_out, void* stream);
int ddl_slab_xr2c(ddl_plan* plan, const double* x_in, void* c_out, void* stream);
/* forward y pass: c arrays -> x-side arrays;  forward z pass: k-side arrays -> e arrays, or
 * (full_out) straight into local k slabs (retained modes only are written) */
int ddl_slab_yfwd(ddl_plan* plan, int nf, void* const* c_in, void* const* xside_out, void* stream);
int ddl_slab_zfwd(ddl_plan* plan, int nf, void* const* kside_in, void* const* out, int full_out, void* stream);
/* spectral assembly: derivatives, curl, solenoidal projection (physics.py:180-195,374-416,588-599) */
int ddl_slab_assemble(ddl_plan* plan, int physics, const ddl_phys_params* params, void* const* e_in,
                      void* const* state, void* const* deriv, void* stream);

/* ---- peer-to-peer exchange over NVLink (one process per GPU; csrc/p2p.cu) -------------------
 * Replaces the MPI all-to-all FFTW-MPI performs inside every transform (_fftw.pyx:272-304).
 * Each rank owns an "arena" (cudaMalloc + CUDA IPC handle) holding its k-side / x-side arrays;
 * ddl_p2p_connect maps every peer's arena.  ddl_p2p_exchange enqueues, behind the work already
 * on `stream`, copy-engine pushes of n blocks into the peers' arenas on an internal stream and
 * then raises this rank's arrival flag (the returned sequence number) in every peer's arena;
 * ddl_p2p_wait makes `stream` wait (device-side spin on the flags) until all peers' pushes of
 * that sequence number have landed here.  Every rank must issue the same sequence of exchanges. */
typedef struct ddl_p2p ddl_p2p;
int ddl_p2p_create(ddl_p2p** out, int nranks, int rank, size_t data_bytes, char* ipc_handle_out64);
int ddl_p2p_connect(ddl_p2p* ctx, const char* ipc_handles /* nranks x 64 bytes */);
void* ddl_p2p_base(ddl_p2p* ctx);
long long ddl_p2p_exchange(ddl_p2p* ctx, int n, const int* dst_rank, const int64_t* src_off,
                           const int64_t* dst_off, const int64_t* nbytes, void* stream);
int ddl_p2p_wait(ddl_p2p* ctx, long long seq, void* stream);
/* Exchange fused into the producing pass: ddl_slab_zinv_peer / ddl_slab_yfwd_peer store every
 * output row straight into the owning rank's arena over NVLink (no send buffer, no copy);
 * peer_tab is a DEVICE array of nf x nranks pointers, entry [f][s] =
 *   zinv: rank s's x-side field f + cy0_me*nzl*CX      yfwd: rank s's k-side field f + (me*cyl_s - cy0_s)*nzl*CX
 * built from ddl_p2p_peer_base.  ddl_p2p_signal then publishes the arrival flag behind the pass. */
void* ddl_p2p_peer_base(ddl_p2p* ctx, int rank);
long long ddl_p2p_signal(ddl_p2p* ctx, void* stream);
/* The same blocks as ddl_p2p_exchange, moved by a small SM kernel on `stream` instead of the copy engines (which lose most of
 * their rate while an HBM-bound pass runs): entry i = nrows[i] rows of row_bytes[i] bytes, pitch[i] bytes apart on both sides
 * (all multiples of 16), from this arena + src_off[i] to rank dst_rank[i]'s arena + dst_off[i].  A few CTAs (`ctas`, 0 = 32)
 * without shared memory saturate NVLink and share the SMs with the pass running on another stream -- the all-to-all FFTW-MPI
 * performs inside every transform (_fftw.pyx:272-304), overlapped with the 1-D transforms.  publish = 1: raise this
 * exchange's arrival flag in every peer behind the copies and return its sequence number; 0: return 0 (more chunks follow). */
long long ddl_p2p_push(ddl_p2p* ctx, int n, const int* dst_rank, const int64_t* src_off, const int64_t* dst_off,
                       const int64_t* row_bytes, const int64_t* nrows, const int64_t* pitch, int ctas, int publish, void* stream);
/* forward y pass of local planes [z0, z0 + nzc) only, full-size arrays on both sides (chunked so that the push of chunk c
 * overlaps the x pass of chunk c + 1): c_in[f] = C[nzl][y][CX]  ->  xs_out[f] = x-side pencils [cy][nzl][CX] */
int ddl_slab_yfwd_planes(ddl_plan* plan, int nf, void* const* c_in, void* const* xs_out, int z0, int nzc, void* stream);
int ddl_slab_zinv_peer(ddl_plan* plan, int nf, void* const* k_in, void* const* peer_tab, void* stream);
/* yfwd_peer and xfused_planes work on local planes [z0, z0 + nzc) only, so that the forward y
 * pass of one chunk (NVLink-bound) overlaps the x pass of the next (compute-bound) */
int ddl_slab_yfwd_peer(ddl_plan* plan, int nf, void* const* c_in, void* const* peer_tab, int z0, int nzc, void* stream);
int ddl_slab_xfused_planes(ddl_plan* plan, int physics, const ddl_phys_params* params, void* const* b_in,
                           void* const* c_out, int z0, int nzc, void* stream);
int ddl_p2p_destroy(ddl_p2p* ctx);

/* forward_step_cy_{2d,3d}.pyx euler/etd1/etd2rk1/etd2rk2 for ncomp components at once.
 * The integrating factor is not an array: Z = -coeff[c] * (k^2)^visc_order * dt is formed
 * from the plan's wavenumbers (coeff = nu / kappa / eta, 0 => the reference's IF None =>
 * Euler branch).  deriv2 may be NULL for EULER / ETD1.  out may alias start.            */
int ddl_stage(ddl_plan* plan, int kind, int ncomp, void* const* start, void* const* out,
              void* const* deriv1, void* const* deriv2, const double* coeff, int visc_order,
              double dt, int flags, void* stream);

/* The same four kernels with the reference's own signature (forward_step_cy_3d.pyx:17-21,34-38,63-68,95-100; _2d: f0 from
 * the series in the small-|Z| branch, :55): complex arrays of `count` elements and the integrating factor as a real ARRAY
 * (Z = intfactor[i] * dt; callers pass -IF, time_step.py:289,301).  intfactor NULL = the Euler forms the reference takes
 * where integrating_factor is None (time_step.py:285-304,372-386).  No plan: nothing but the arrays.  For factors that are
 * not c (k^2)^n of static wavenumbers (shearing box), and for binding the Cython kernels one to one. */
int ddl_step_array(int kind, int ndim, long long count, const void* start, void* out, const void* deriv1, const void* deriv2,
                   const double* intfactor, double dt, void* stream);

/* One stage of the restated RK4 (time_step.py:395-483 + forward_step :187-221):
 *   total = (first ? 0 : total) + k / wdiv ;  if (!last) out = S(y, k, dt_step)
 *   else out = S(y, total, dt_step),  S = euler / etd1(-IF)                              */
int ddl_rk4_stage(ddl_plan* plan, int ncomp, void* const* y, void* const* k, void* const* total,
                  void* const* out, const double* coeff, int visc_order, double wdiv,
                  double dt_step, int first, int last, int flags, void* stream);

/* The RHS with the spectral assembly fused into a stage update: the derivative is formed in
 * registers and consumed by the update (same arithmetic as ddl_stage / ddl_rk4_stage / ddl_cn_step),
 * written only if k_out is given.  kind: DDL_EULER / DDL_ETD1 (the derivative of this RHS is THE
 * derivative), DDL_ETD2RK1 / DDL_ETD2RK2 (it is the SECOND one; deriv1 holds the first),
 * DDL_FUSE_RK4 (ddl_rk4_stage semantics with total / wdiv / first / last), DDL_FUSE_CN.
 * The sweep visits the retained modes only: total, deriv1 and k_out must vanish outside the dealias mask (what
 * DDL_STAGE_RETAINED_ONLY asserts); state / y may carry content out there, which ddl_stage_outside then updates.
 * ddl_slab_assemble_stage is the same tail for the slab phases. */
enum { DDL_FUSE_RK4 = 4, DDL_FUSE_CN = 5 };
typedef struct ddl_stage_fuse {
    void* const* y;       /* start state of the update (ncomp k-space arrays) */
    void* const* total;   /* RK4: running (k1 + 2k2 + 2k3 + k4)/6, else NULL */
    void* const* out;     /* state written (may alias y) */
    const double* coeff;  /* nu / kappa / eta per component (0: no integrating factor) */
    int visc_order;
    int first, last;      /* RK4 */
    double wdiv, dt_step; /* RK4 weight divisor; step of this update */
    int kind;
    int reserved;
    void* const* deriv1;  /* ETD2RK1/2: first derivative (read), else NULL */
    void* const* k_out;   /* optional: where to store the derivative formed here, else NULL */
} ddl_stage_fuse;
int ddl_rhs_stage(ddl_plan* plan, int physics, const ddl_phys_params* params, void* const* state, void* work,
                  size_t work_bytes, int flags, const ddl_stage_fuse* fuse, void* stream);
int ddl_slab_assemble_stage(ddl_plan* plan, int physics, const ddl_phys_params* params, void* const* e_in,
                            void* const* state, const ddl_stage_fuse* fuse, void* stream);

/* Companion of ddl_rhs_stage for a state that carries content OUTSIDE the dealias mask (hydro-type physics never dealias their
 * state; the reference's 2-D Taylor-Green field has such entries): there every derivative of the fused pipeline
 * vanishes and the update is the integrating factor alone, out = S(start, 0).  Same arithmetic as ddl_stage / ddl_rk4_stage /
 * ddl_cn_step with zero derivatives (bit-identical to their full sweeps), over the masked-out entries only.  kind as in
 * ddl_stage_fuse.  With it such states keep the fused stage kernel for their retained modes. */
int ddl_stage_outside(ddl_plan* plan, int kind, int ncomp, void* const* start, void* const* out, const double* coeff,
                      int visc_order, double dt, void* stream);

/* restated CrankNicholsonVisc (time_step.py:486-506): y = (top/bottom) y + k / bottom */
int ddl_cn_step(ddl_plan* plan, int ncomp, void* const* y, void* const* k, const double* coeff,
                int visc_order, double dt, int flags, void* stream);

/* ---- reductions: invariants and the CFL limit -------------------------------------------------
 * Replace the volume-average tasks of dedalus/analysis/volume_average.py:71-331 and
 * VectorFieldBase.max_square / compute_dt / set_dtlist (dedalus/data_objects/fields.py:153-157,
 * dedalus/physics/physics.py:151-158,601-610,714-721,821-836).  Results land in DEVICE memory
 * (caller-owned); a rank reduces its own slab, the caller combines ranks (sum / max).
 *
 * ddl_reduce_invariants: ONE sweep over the state (u, then T or B; StateData order) fills
 * out[DDL_NINV].  w = 1 on the kx = 0 plane, 2 elsewhere (volume_average.py:84-97); W = k x u,
 * J = k x B.  flags: DDL_STAGE_RETAINED_ONLY when the state is known to vanish outside the mask. */
enum {
    DDL_INV_EKIN = 0,        /* sum w |u|^2 / 2                        ekin      :108-118 */
    DDL_INV_E2 = 1,          /* sum w |B|^2 / 2 (emag :179-188) or sum w |T|^2 / 2 (temp2 / 2 :172-177) */
    DDL_INV_DIV_SUM = 2,     /* sum |i k.u|, unweighted                divergence_sum :287-295 */
    DDL_INV_MAG_DIV_SUM = 3, /* sum |i k.B|                            mag_div_sum    :315-322 */
    DDL_INV_ENSTROPHY = 4,   /* sum w |W|^2 / 2   (2-D: enstrophy :190-199; 3-D: energy_dissipation / (2 nu) :250-260) */
    DDL_INV_CURRENT2 = 5,    /* sum w |J|^2 / 2 */
    DDL_INV_HEL_KIN = 6,     /* sum w Re(u . conj(i W)),  3-D only */
    DDL_INV_HEL_CROSS = 7,   /* sum w Re(u . conj(B)) */
    DDL_INV_DIV_RE = 8, DDL_INV_DIV_IM = 9,         /* sum w (i k.u)  divergence :273-280 */
    DDL_INV_MAG_DIV_RE = 10, DDL_INV_MAG_DIV_IM = 11, /* mag_div :305-312 */
    DDL_INV_CENK_NUM = 12,   /* sum (k^2)^1.5 |u|^2/2 and */
    DDL_INV_CENK_DEN = 13,   /* sum k^2 |u|^2/2, k^2(0) := 1, en[0,0] := 0   vort_cenk :201-212 */
    DDL_INV_MSQ = 14,        /* +c: sum w |component c|^2, c < 6       ux2 .. bz2, temp2 :123-177 */
    DDL_INV_HEL_MAG = 20,    /* sum w Re(A . conj(B)), A = i J / k^2, 3-D MHD only */
    DDL_INV_GRAD2_T = 21,    /* sum w k^2 |T|^2                        thermal_energy_dissipation / kappa :262-271 */
    DDL_INV_DIV2 = 22,       /* sum w |k.u|^2: with 2 * ENSTROPHY it splits sum w |k|^2 |u|^2 into compressive + solenoidal */
    DDL_INV_MAG_DIV2 = 23,   /* sum w |k.B|^2 (with 2 * CURRENT2); the host layer picks the *_ADV policies from these */
    DDL_NINV = 24
};
int ddl_reduce_invariants(ddl_plan* plan, int physics, void* const* state, int flags, double* out, void* stream);

/* out[i] (device, 8 doubles) = number of non-zero entries of k-space array i (n <= 8) OUTSIDE the plan's dealias mask.  The
 * reference never knows whether a spectrum is dealiased and masks on every transform (representations.py:344,353); here a buffer
 * that IS zero out there takes the retained-modes-only sweeps and the fused stage kernel, and this one read-only pass over the
 * masked-out entries re-establishes that knowledge for buffers the caller has written. */
int ddl_reduce_outside_mask(ddl_plan* plan, int n, void* const* arrays, double* out, void* stream);

/* CFL capture: the only place where u(x), B(x) exist is inside the x pass of the RHS, so the
 * maxima the time-step limit needs are reduced THERE.  While out2 (device, 2 doubles, zeroed by
 * the caller) is set, every x pass of ddl_rhs / ddl_rhs_stage / ddl_slab_xfused* on this plan does
 *   out2[0] = max(out2[0], max_{x,i} u_i(x)^2),   out2[1] = max(out2[1], max_{x,i} B_i(x)^2)
 * (Boussinesq: T(x)^2) -- the maximum over space AND components of the squared component, which is what
 * VectorFieldBase.max_square returns (fields.py:153-157) -- of the state the RHS was evaluated at.
 * NULL switches the capture off (the default; the x pass then is the plain kernel). */
int ddl_rhs_capture_max(ddl_plan* plan, double* out2);
/* max_square on its own (compute_dt outside a step): the inverse half of the RHS pipeline with
 * the capture on, nothing else written but workspace.  One-rank plans; a slab-decomposed caller
 * runs ddl_slab_zinv / yinv / xfused with ddl_rhs_capture_max set.  flags: DDL_RHS_DEALIAS_STATE
 * (the reference's max_square transforms the state itself, which masks it in place). */
int ddl_reduce_max_square(ddl_plan* plan, int physics, const ddl_phys_params* params, void* const* state,
                          void* work, size_t work_bytes, int flags, double* out2, void* stream);

/* launch accounting: cumulative number of kernels this library has launched in the process,
 * and optional per-launch CUDA-event timing aggregated by kernel label (bench.py roofline) */
long long ddl_launch_count(void);
int ddl_profile_enable(int on);
int ddl_profile_report(char* json_out, size_t nbytes);

/* "fast_kernels" = 0 routes every pass through the generic tile kernel (tests compare both);
 * "xfused_variant" = 0..6, 8 picks the variant of the fused x pass (csrc/xfused_kernel.cuh: CTA shapes, retained count at compile
 *   time, persistent / staged kernels, branch-free input packs; 0 = the measured best; unknown values run 0);
 * "traceless_flux" = 0/1 (default 1): the one-rank 3-D ddl_rhs / ddl_rhs_stage form 5 momentum products (T_ij - delta_ij T_zz) instead
 *   of 6 -- the difference is a pressure, removed by the solenoidal projection mode by mode -- i.e. one forward transform fewer per
 *   evaluation; same derivative to the rounding of the projection (1e-16); the slab-decomposed phases always form 6;
 * "strided_two" = 0/1/2: the y / z passes of lengths 256 and 512 on plain rows run as two register butterflies (16 x 16, 16 x 32)
 *   around one trip through shared memory (csrc/fast_two.cuh; 1 = stage-0 twiddles generated in registers, the default; 2 = loaded
 *   from the table) or as the three radix-8 stages of strided_fast (0); the slab-decomposed passes always take the latter;
 * "assemble_variant" = 0..4 picks the launch variant of the spectral assembly fused with the stage update (ddl_rhs_stage): 3 (the
 *   default) loads every operand of a mode before any arithmetic and runs two 256-thread CTAs per SM at 128 registers, 0 is the
 *   round-1 kernel (loads between the per-component updates, four CTAs per SM); all variants are the same arithmetic, bit for bit;
 * "rhs_plane_chunk" = n > 0 runs y_inv -> x -> y_fwd of the one-rank 3-D RHS over chunks of n z-planes with chunk-sized,
 *   reused half-transformed arrays (an L2-residency experiment, default 0 = off; measured slower on B200);
 * "peer_pass_ctas" = n > 0 limits the slab passes that store to the peers over NVLink (ddl_slab_zinv_peer / ddl_slab_yfwd_peer) to n
 *   CTAs that walk the tiles themselves (NVLink-bound: a few CTAs per SM pair saturate the links and the rest of the SMs stays
 *   free for the pass running on the other stream); 0 = one CTA per tile;
 * "p2p_timeout_s" = seconds a consumer pass of the peer exchange waits for a peer's arrival flag before it traps the context
 *   (default 600; 0 = wait for ever, like the blocking MPI all-to-all of _fftw.pyx:272-304 it replaces) */
int ddl_set_option(const char* name, int value);

/* Measured roofs of the FP64 vector pipe on the current device (bench.py roofline; nothing in the reference corresponds):
 * out2[0] = TFLOP/s of independent DFMA chains (2 flops each), out2[1] = TFLOP/s of DADD chains (1 flop each: the
 * instruction-issue roof an FFT butterfly, which is mostly additions, actually meets).  Host pointers; synchronises. */
int ddl_measure_fp64(double* out2, void* stream);

int ddl_sync(void* stream);
const char* ddl_last_error(void);
const char* ddl_version(void);

#ifdef __cplusplus
}
#endif
#endif /* DDL_H */
