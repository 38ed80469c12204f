/* ies_b200 -- C-ABI of the B200-native time-stepping engine for IES.
 *
 * This is the drop-in boundary for the reference's hot path
 * (space.Basic3D.updateH / updateE and the per-step work fused around it).
 * The reference has no FFI of its own (it is pure Python on numpy/cupy); each
 * entry point below names the reference code it replaces (file:line in
 * steve1029/IES).  The Python host layer (ies_b200/space.py, source.py,
 * collector.py) binds these through ctypes and mirrors the reference's classes.
 *
 * Conventions: every function returns 0 on success, non-zero on error;
 * ies_last_error() gives the message of the last failure on the calling
 * thread.  All arrays are C-order (z fastest), complex values interleaved
 * (re, im).  Field components are numbered Ex=0 Ey=1 Ez=2 Hx=3 Hy=4 Hz=5.
 * There is NO CPU fallback: every compute entry point launches sm_100a kernels.
 */
#ifndef IES_B200_H
#define IES_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct ies_ctx ies_ctx;           /* one x-slab of one Basic3D space on one GPU */
typedef struct ies_dft ies_dft;           /* one Sx/Sy/Sz running-DFT collector        */
typedef struct ies_probe ies_probe;       /* one FieldAtPoint time-series recorder      */

enum { IES_F32 = 0, IES_F64 = 1, IES_C64 = 2, IES_C128 = 3 };      /* field_dtype, space.py:117-122 */
enum { IES_FDTD = 0, IES_SHPF = 1, IES_PSTD = 2 };                 /* method,      space.py:79-89   */
enum { IES_EX = 0, IES_EY, IES_EZ, IES_HX, IES_HY, IES_HZ };
enum { IES_HALF_H = 0, IES_HALF_E = 1 };
/* derivative slots of one half-step; F is the field being differentiated
 * (E in updateH, H in updateE): space.py:193-205 */
enum { IES_D_YFZ = 0, IES_D_ZFY, IES_D_ZFX, IES_D_XFZ, IES_D_XFY, IES_D_YFX };

typedef struct {
    int32_t nx, ny, nz;        /* LOCAL slab (myNx, Ny, Nz), space.py:114-115 */
    int32_t dtype;             /* IES_F32..IES_C128 */
    int32_t method;            /* IES_FDTD / IES_SHPF / IES_PSTD */
    int32_t rank, nranks;      /* x-slab index, space.py:46-48, 108-137 */
    int32_t device;            /* CUDA device ordinal */
    double dx, dy, dz, dt;
} ies_config;

/* One CPML correction term = one "Update <comp> at <face>" block of
 * space.py:1110-1712:  psi = b*psi + a*d ;  F += sign * C2 * (kf*d + psi),
 * kf = 1/kappa - 1, on the box [lo,hi) of local cell indices.  The profile
 * tables are already gathered per cell along `axis` (length hi[axis]-lo[axis]),
 * so the odd/even half-cell sampling of the reference stays on the host. */
typedef struct {
    int32_t half;              /* IES_HALF_H / IES_HALF_E */
    int32_t comp;              /* 0,1,2 = x,y,z component of the updated field */
    int32_t diff;              /* IES_D_* slot feeding this term */
    int32_t axis;              /* 0,1,2 */
    int32_t lo[3], hi[3];      /* field box */
    int32_t psi_off;           /* psi index along axis = (idx - lo[axis]) + psi_off */
    int32_t psi_thick;         /* extent of the psi array along axis (npml) */
    double sign;               /* +1 / -1 */
    const double* b;           /* host pointers, copied by the call */
    const double* a;
    const double* kf;
} ies_pml_term;

/* ---- lifetime ------------------------------------------------------------ */
const char* ies_last_error(void);
int ies_device_count(int* n);
/* Basic3D.__init__ + malloc (space.py:9-237): allocates the six fields and the
 * per-method scratch on cfg->device; fields start at zero. */
int ies_create(const ies_config* cfg, ies_ctx** out);
int ies_destroy(ies_ctx* ctx);
/* Run the context's work on an externally owned stream (use_own_stream = 0; a handle of 0 is then
 * the caller's legacy default stream, not a special value) or go back to the context's own
 * non-blocking stream (use_own_stream = 1, cuda_stream ignored). */
int ies_set_stream(ies_ctx* ctx, void* cuda_stream, int use_own_stream);
int ies_sync(ies_ctx* ctx);
/* Engine tuning knobs (no reference counterpart; defaults need no call; every combination computes
 * bit-identical fp64 fields).  Names:
 *  "fused"      SHPF, real dtypes, ny == nz in {64,128,256,512}: 1 = the half-step as ONE launch, z-line and
 *               y-line tiles as two roles of one grid [shpf_fused.cuh]; 0 = z-line derivative kernel +
 *               y-line update kernel; -1 [default] = fused where it is faster (fp64, lines <= 256 points);
 *  "fused_lead" planes the z role runs ahead (6); "fused_ring" scratch ring size in planes (0 = full-size
 *               scratch); "fused_zb" z tiles per z-role CTA (1, 2 [default] or 4); "fused_prefetch" (0); "fused_discard" 1 [default] =
 *               a y tile drops the z-derivative scratch lines it has consumed from L2 (discard.global.L2), so
 *               the dirty scratch is never written back to HBM;
 *  "pml_split"  CPML corrections: 1 = a separate pass over the absorber cells after the update kernels
 *               [k_pml_terms], 0 = inside the update kernels, -1 [default] = separate when a y/z face carries
 *               terms or the slab has >= 2^23 cells;
 *  "ctile"      1 [default] = tiles whose cells share one coefficient skip the coefficient array;
 *  "palette"    1 = palette-compressed coefficient arrays when they hold <= 32 distinct values (0);
 *  "fdtd_vec"   1 [default] = 16-byte vectorised FDTD kernel when nz allows;
 *  "reset_psi"  zero the CPML auxiliary arrays;  "fused_prof" development cycle counters. */
int ies_set_option(ies_ctx* ctx, const char* name, int64_t value);

/* ---- setup --------------------------------------------------------------- */
/* init_update_constants (space.py:445-553) with zero conductivity: one f64
 * array per half-step, CH2 = -2dt/(2mu), CE2 = 2dt/(2eps), full local grid. */
int ies_set_coeff(ies_ctx* ctx, int half, const double* host, int64_t n);
/* Main-update sub-volume of component comp (0..5) -- the slices of
 * space.py:801-825, 1017-1037 -- as a box of local indices. */
int ies_set_update_box(ies_ctx* ctx, int comp, const int32_t lo[3], const int32_t hi[3]);
/* Spectral multiplier of one axis for one half-step: `n` complex128 values on
 * the FULL spectrum (fftfreq order), i.e. ik*exp(+-ik d/2) (- i k_Bloch ...) of
 * space.py:164-181, 709-755, 1893-1950 with the rfft Hermitian extension done
 * by the host. */
int ies_set_multiplier(ies_ctx* ctx, int half, int axis, const double* re_im, int32_t n);
/* Drop all CPML terms / add one (space.py:239-361, 1054-1712). */
int ies_clear_pml(ies_ctx* ctx);
int ies_add_pml_term(ies_ctx* ctx, const ies_pml_term* term);
/* FDTD ghost-plane copies of _exchange_BBC{x,y,z} (space.py:1714-1858,
 * 1981-2033): F[-1] = F[1]*pp ; F[0] = F[-2]*pm.  enabled=0 switches an axis off. */
int ies_set_ghost(ies_ctx* ctx, int axis, int enabled,
                  double pp_re, double pp_im, double pm_re, double pm_im);
/* x-slab neighbours present?  (rank != 0, rank != size-1; space.py:645, 662) */
int ies_set_neighbours(ies_ctx* ctx, int has_prev, int has_next);

/* ---- hot path ------------------------------------------------------------ */
/* Basic3D.updateH / updateE (space.py:639-840, 842-1052) WITHOUT the MPI
 * exchange: the halo planes must already be in the buffers of ies_halo_recv_ptr. */
int ies_update_h(ies_ctx* ctx, int64_t tstep);
int ies_update_e(ies_ctx* ctx, int64_t tstep);
/* The same half-step in two parts, so that a slab's halo exchange (space.py:645-670, 863-887)
 * can overlap the part that needs no neighbour plane: phase 0 = z-line / x-line derivative
 * passes (nothing for FDTD), phase 1 = the rest.  phase 0 then phase 1 == ies_update_h/e. */
int ies_update_phase(ies_ctx* ctx, int half, int phase);
/* Device pointers for the halo exchange (space.py:645-670, 863-887).
 * half = IES_HALF_H: send = Ey[0], Ez[0] (to rank-1), recv = planes of rank+1;
 * half = IES_HALF_E: send = Hy[-1], Hz[-1] (to rank+1), recv = planes of rank-1.
 * which = 0 (y component) or 1 (z component).  *bytes = plane size. */
int ies_halo_send_ptr(ies_ctx* ctx, int half, int which, void** dev, int64_t* bytes);
int ies_halo_recv_ptr(ies_ctx* ctx, int half, int which, void** dev, int64_t* bytes);
/* Same-process neighbour exchange (one process driving several GPUs):
 * copies src's send planes into dst's recv planes with cudaMemcpyPeerAsync,
 * ordered after src's stream and before dst's next update. */
int ies_halo_copy(ies_ctx* dst, ies_ctx* src, int half);
/* One process per GPU (the reference's `mpirun -n R`): the blocking pickled mpi4py send/recv of
 * space.py:645-670, 863-887 becomes copy-engine transfers over NVLink into peer-mapped memory.
 *   ies_halo_ipc_export : 64-byte CUDA IPC handle of this slab's receive planes + arrival flags;
 *   ies_halo_ipc_connect: map a neighbour's handle (nbr 0 = rank-1, 1 = rank+1);
 *   ies_halo_push(half) : enqueue the copy of my two send planes (half H: Ey[0],Ez[0] -> rank-1,
 *                         half E: Hy[-1],Hz[-1] -> rank+1) into that neighbour's receive planes and,
 *                         in stream order after them, the write of the push count to its flag;
 *   ies_halo_wait(half) : enqueue a stream memory wait (flag >= my wait count) in front of the
 *                         update that reads the planes.
 * Every updateH / updateE of a slab with a neighbour does one push and one wait; nothing blocks
 * the host and no kernel runs for the exchange. */
int ies_halo_ipc_export(ies_ctx* ctx, void* handle64);
int ies_halo_ipc_connect(ies_ctx* ctx, int nbr, const void* handle64);
int ies_halo_push(ies_ctx* ctx, int half);
int ies_halo_wait(ies_ctx* ctx, int half);

/* Setter.put_src (source.py:167-253): F[lo:hi] (+)= pulse * px[i]*py[j]*pz[k].
 * px/py/pz are complex128 (re,im) tables of the box extents or NULL (=1). */
int ies_put_src(ies_ctx* ctx, int comp, const int32_t lo[3], const int32_t hi[3],
                double re, double im, int hard,
                const double* px, const double* py, const double* pz);

/* ---- field access (gather / snapshots / tests) ---------------------------- */
int ies_get_field(ies_ctx* ctx, int comp, const int32_t lo[3], const int32_t hi[3], void* host);
int ies_set_field(ies_ctx* ctx, int comp, const int32_t lo[3], const int32_t hi[3], const void* host);
int ies_field_ptr(ies_ctx* ctx, int comp, void** dev);

/* ---- collectors ----------------------------------------------------------- */
/* Sx/Sy/Sz (collector.py:266-801): running DFT of four components on the box
 * [lo,hi) (one axis has extent 1).  comps[4] are field ids; the DFT arrays are
 * complex128 of shape (nf, n1, n2). */
int ies_dft_create(ies_ctx* ctx, const int32_t lo[3], const int32_t hi[3],
                   const int32_t comps[4], const double* freqs, int32_t nf, ies_dft** out);
/* do_RFT(tstep): DFT += (A - B)[box] * exp(2 pi i f tstep dt) * dt.  b may be NULL
 * (plain space) or the incident-field context (Empty3D.get_SF, space.py:2157-2179,
 * evaluated lazily on the collector plane only).  The call samples the plane into a ring
 * of 16 time slots; the accumulators are updated every 16 calls (and on ies_dft_read) in
 * time order, i.e. with the same arithmetic and summation order as a per-step update. */
int ies_dft_accumulate(ies_dft* d, ies_ctx* a, ies_ctx* b, int64_t tstep);
int ies_dft_read(ies_dft* d, int which, void* host_c128);
int ies_dft_destroy(ies_dft* d);
/* FieldAtPoint (collector.py:123-201): six components at one cell per step. */
int ies_probe_create(ies_ctx* ctx, int32_t i, int32_t j, int32_t k, int64_t tsteps, ies_probe** out);
int ies_probe_record(ies_probe* p, ies_ctx* a, ies_ctx* b, int64_t tstep);
int ies_probe_read(ies_probe* p, int comp, void* host);
int ies_probe_destroy(ies_probe* p);

/* ---- measurement ---------------------------------------------------------- */
/* Kernel launches issued by this library since process start (bench.py's gpu_launches). */
int64_t ies_launch_count(void);
/* CUDA events on the context's stream: ies_timer_start ... ies_timer_stop(&ms). */
int ies_timer_start(ies_ctx* ctx);
int ies_timer_stop(ies_ctx* ctx, double* ms);
/* Per-kernel CUDA-event timing.  ies_profile(ctx, 1) starts bracketing every hot-path
 * launch with events (0 stops and clears); ies_profile_read sums the durations of slot
 * 0 = z-line derivative, 1 = y-line derivative + fused update, 2 = x-line derivative
 * (PSTD), 3 = FDTD update. */
int ies_profile(ies_ctx* ctx, int on);
int ies_profile_read(ies_ctx* ctx, int slot, double* ms_total, int64_t* launches);
/* Development: SM-cycle totals per phase of the fused SHPF kernel since ies_set_option("fused_prof", 1):
 * [0] z FFT, [1] z ring wait, [2] z store+signal, [3] y FFT, [4] y wait for z, [5] y update,
 * [8] z tiles, [9] y tiles. */
int ies_fused_prof_read(ies_ctx* ctx, uint64_t* out16);

#ifdef __cplusplus
}
#endif
#endif
