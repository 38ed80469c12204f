// Spectral-derivative kernels for SHPF / PSTD (space.py:709-755, 903-961) and the
// fused y-line derivative + field update + CPML kernel (space.py:801-811, 1017-1025,
// 1110-1712).  Included by the per-dtype translation units spectral_<dtype>.cu.
//
// Pass structure of one half-step (F = field being differentiated, G = updated field):
//   k_zline : z lines (contiguous), pair (F_y, F_x) packed as one complex line
//             -> scratch dz[0] = d/dz F_y, dz[1] = d/dz F_x
//   k_xline : PSTD only, x lines (stride ny*nz), pair (F_z, F_y) -> scratch dxs[0..1]
//   k_yline_update : y lines (stride nz) of the pair (F_z, F_x) transformed in
//             registers, then the whole cell update (x difference, curl, CPML, store).
// For real fields two real lines share one complex FFT: the derivative operator is a
// real circulant, so D(a + i b) = Da + i Db and no spectrum splitting is needed; the
// multiplier table is the Hermitian extension of the reference's rfft multiplier.
#pragma once
#include <type_traits>
#include "engine.h"
#include "fft_dev.cuh"
#include "update_dev.cuh"

namespace ies {

template <typename T, bool CPLX> struct Fld;
template <typename T> struct Fld<T, false> {
    using C = typename Cx<T>::type;
    static constexpr int NF = 1;          // FFTs per line pair
    // pack element idx of real arrays A,B into one complex value
    static __device__ __forceinline__ C ld(const void* A, const void* B, size_t i, int) {
        C v; v.x = ((const T*)A)[i]; v.y = ((const T*)B)[i]; return v;
    }
    static __device__ __forceinline__ void st(void* dA, void* dB, size_t i, C v, int) {
        ((T*)dA)[i] = v.x; ((T*)dB)[i] = v.y;
    }
    static __device__ __forceinline__ void st_stream(void* dA, void* dB, size_t i, C v, int) {
        __stcs((T*)dA + i, v.x); __stcs((T*)dB + i, v.y);
    }
};
template <typename T> struct Fld<T, true> {
    using C = typename Cx<T>::type;
    static constexpr int NF = 2;
    static __device__ __forceinline__ C ld(const void* A, const void* B, size_t i, int f) {
        return ((const C*)(f == 0 ? A : B))[i];
    }
    static __device__ __forceinline__ void st(void* dA, void* dB, size_t i, C v, int f) {
        ((C*)(f == 0 ? dA : dB))[i] = v;
    }
    static __device__ __forceinline__ void st_stream(void* dA, void* dB, size_t i, C v, int f) {
        __stcs((C*)(f == 0 ? dA : dB) + i, v);
    }
};

template <typename C>
__device__ __forceinline__ void load_tables(C* tw, C* ml, const C* twg, const C* mlg, int n) {
    for (int i = threadIdx.x; i < n; i += blockDim.x) { tw[i] = twg[i]; ml[i] = mlg[i]; }
    __syncthreads();
}

// ---------------------------------------------------------------- z lines -----
template <int N, int VPT> struct ZCfg {
    static constexpr int TT = N / VPT;
    static constexpr int LPB = (256 / TT) > 0 ? (256 / TT) : 1;   // lines per block
    static constexpr int THREADS = LPB * TT;
};
template <typename C, int N, int VPT> struct ZXchg { using type = XchgContig<C, N>; };
template <typename C, int N> struct ZXchg<C, N, 8> { using type = XchgContig8<C, N>; };
// 512-point z lines cross threads with the split real / imaginary exchange: 64 instead of 98 KB of shared memory
// per CTA (two CTAs per SM by registers either way: 100 KB more L1 for the streaming loads / stores); Mie slab 4.30 -> 4.19 ms/step
// (same-box A/B of two builds, tools/gpu_call_r2i.sh).  -DIES_Z512_NOSPLIT restores the padded complex exchange.
#ifndef IES_Z512_NOSPLIT
template <typename C> struct ZXchg<C, 512, 16> { using type = XchgContigSplit<C, 512>; };
#endif

// One CTA = LPB adjacent z lines.  Input lines start at line0 (+ blockIdx.x*LPB), the
// derivative lines go to oline0 (+ blockIdx.x*LPB) of the scratch arrays.  VPT = register
// values per thread (16: fewer, fatter threads; 8: twice the warps at half the registers).
template <typename T, bool CPLX, int N, int VPT>
__global__ void __launch_bounds__(ZCfg<N, VPT>::THREADS, (VPT == 8 ? 3 : 2))
k_zline(const void* __restrict__ A, const void* __restrict__ B, void* __restrict__ dA,
        void* __restrict__ dB, long line0, long oline0, long nlines,
        const typename Cx<T>::type* __restrict__ twg, const typename Cx<T>::type* __restrict__ mlg) {
    using C = typename Cx<T>::type;
    using F = Fld<T, CPLX>;
    using X = typename ZXchg<C, N, VPT>::type;
    // Shared-memory tables: the twiddles of the second forward and the last inverse stage
    // transposed to [m][jj] (the threads of a line read consecutive entries; the master
    // table's stride-m access costs 2.1x the wavefronts at N = 256 and more at N = 512), the
    // master table only when a third stage needs it, and the spectral multiplier.
    using TT_ = TwTables<N, VPT>;
    using P = PlanV<N, VPT>;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    C* tw = reinterpret_cast<C*>(smem_raw);                       // master (TT_::NEED_MASTER ? N : 0)
    C* twf = tw + (TT_::NEED_MASTER ? N : 0);
    C* twi = TT_::SHARED ? twf : twf + TT_::FWD;
    C* ml = twi + TT_::INV;
    C* xbuf = ml + N;
    for (int q = threadIdx.x; q < N; q += blockDim.x) {
        ml[q] = mlg[q];
        if (TT_::NEED_MASTER) tw[q] = twg[q];
        if (N > VPT) {
            constexpr int NSI = N / VPT;
            twi[q] = twg[((q / NSI) * (q % NSI)) & (N - 1)];
            if (!TT_::SHARED && q < TT_::FWD) twf[q] = twg[((q / VPT) * (q % VPT) * TT_::FSTEP) & (N - 1)];
        }
    }
    __syncthreads();
    constexpr bool TWT = N > VPT;
    constexpr int TT = ZCfg<N, VPT>::TT, LPB = ZCfg<N, VPT>::LPB;
    const int t = threadIdx.x % TT, l = threadIdx.x / TT;
    const long line = (long)blockIdx.x * LPB + l;
    const bool ok = line < nlines;
    const size_t ibase = (size_t)(line0 + line) * N;
    const size_t obase = (size_t)(oline0 + line) * N;
    X xb{reinterpret_cast<decltype(X::base)>(xbuf + (size_t)l * X::LS)};
#pragma unroll 1
    for (int f = 0; f < F::NF; ++f) {
        C v[VPT];
#pragma unroll
        for (int q = 0; q < VPT; ++q) {
            if (ok) v[q] = F::ld(A, B, ibase + line_index_v<N, VPT>(t, q), f);
            else { v[q].x = 0; v[q].y = 0; }
        }
        fft_forward_v<N, VPT, TWT>(v, t, tw, xb, twf);
#pragma unroll
        for (int q = 0; q < VPT; ++q) v[q] = cmul(v[q], ml[spec_index_v<N, VPT>(t, q)]);
        fft_inverse_v<N, VPT, TWT>(v, t, tw, xb, twi);
        if (ok) {
#pragma unroll
            // st.global.cs: the scratch is read once, a whole grid sweep later
            for (int q = 0; q < VPT; ++q) F::st_stream(dA, dB, obase + line_index_v<N, VPT>(t, q), v[q], f);
        }
    }
}

// ------------------------------------------------------- strided lines (x) -----
template <typename T, bool CPLX, int N> struct SCfg {
    static constexpr int TT = N / 16;
    static constexpr int ES = (int)sizeof(T) * (CPLX ? 2 : 1);
    static constexpr int THREADS = 256;
    static constexpr int W = THREADS / TT;                        // adjacent lines (columns) per CTA
    static_assert(W * ES >= 32, "row segment below one sector");
};

// Tile of the fused y-line kernel: W = THREADS/TT adjacent columns x all N rows.  Complex
// dtypes keep two stash buffers (one per transformed field), so their CTAs are half as wide
// (128 threads, 64 KB of shared memory for c128) to keep several CTAs per SM.
template <typename T, bool CPLX, int N> struct YCfg {
    static constexpr int TT = N / 16;
    static constexpr int ES = (int)sizeof(T) * (CPLX ? 2 : 1);
#ifndef IES_Y512_THREADS
#define IES_Y512_THREADS 256
#endif
    static constexpr int THREADS = (CPLX && TT <= 64) ? 128 : ((!CPLX && N == 512) ? IES_Y512_THREADS : 256);
    static constexpr int W = THREADS / TT;
    static constexpr int MINB = CPLX ? 3 : (THREADS > 256 ? 1 : 2);   // (real dtypes: 128-thread CTAs 1.14 -> 1.32 ms at N = 256; 512-thread CTAs at N = 512 +-3 %)   // (128-thread CTAs for real dtypes: 64-byte row segments, 1.14 -> 1.32 ms)                     // CTAs per SM the kernel is compiled for
    static_assert(W >= 1 && W * ES >= 32, "row segment below one sector");
};

template <typename T, bool CPLX, int N>
__global__ void __launch_bounds__(SCfg<T, CPLX, N>::THREADS)
k_xline(const void* __restrict__ A, const void* __restrict__ B, void* __restrict__ dA,
        void* __restrict__ dB, long ncols, long stride, long batch_stride,
        const typename Cx<T>::type* __restrict__ twg, const typename Cx<T>::type* __restrict__ mlg) {
    using C = typename Cx<T>::type;
    using F = Fld<T, CPLX>;
    constexpr int W = SCfg<T, CPLX, N>::W;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    C* tw = reinterpret_cast<C*>(smem_raw);
    C* ml = tw + N;
    C* xbuf = ml + N;
    load_tables(tw, ml, twg, mlg, N);
    const int c = threadIdx.x % W, t = threadIdx.x / W;
    const long col = (long)blockIdx.x * W + c;
    const bool ok = col < ncols;
    const size_t boff = (size_t)blockIdx.y * (size_t)batch_stride;   // batch = one x-plane of y lines
    XchgStrided<C, W> xb{xbuf + c};
#pragma unroll 1
    for (int f = 0; f < F::NF; ++f) {
        C v[16];
#pragma unroll
        for (int q = 0; q < 16; ++q) {
            if (ok) v[q] = F::ld(A, B, boff + (size_t)line_index<N>(t, q) * stride + col, f);
            else { v[q].x = 0; v[q].y = 0; }
        }
        fft_forward<N>(v, t, tw, xb);
#pragma unroll
        for (int q = 0; q < 16; ++q) v[q] = cmul(v[q], ml[spec_index<N>(t, q)]);
        fft_inverse<N>(v, t, tw, xb);
        if (ok) {
#pragma unroll
            for (int q = 0; q < 16; ++q) F::st_stream(dA, dB, boff + (size_t)line_index<N>(t, q) * stride + col, v[q], f);
        }
    }
}

// ------------------------------------------ y lines + fused field update -----
// Column -> slot of a y tile's shared-memory rows (exchange buffer and derivative stash).  fp64 real tiles
// are 16+ columns of 16-byte pairs; phase A has lane = column (a quarter-warp = columns 0..7 or 8..15), phase B
// has lane = column PAIR (a quarter-warp = the even or the odd columns of one row), and a 16-byte access is
// conflict free when its quarter-warp covers the eight 16-byte bank groups.  Swapping neighbours in every second
// group of eight columns, c ^ ((c >> 3) & 1), satisfies both (the identity gave phase B a 2-way conflict on every stash read:
// 16.8 M of 284 M shared wavefronts per launch in the ncu capture of the headline).
template <typename T, bool CPLX, int W>
__device__ __forceinline__ int stash_col(const int c) {
    if constexpr (!CPLX && sizeof(T) == 8 && W >= 16) return c ^ ((c >> 3) & 1);
    else return c;
}

// Phase B of k_yline_update.  FAST: no CPML term touches the tile and every update box either
// contains or misses it (upd = component mask, CTA-uniform): straight-line interior code.
// CM: where the coefficient comes from -- 0 the f64 array, 1 the palette form, 2 one value for
// the whole tile (cuni; materials are piecewise constant, so most tiles are uniform and skip the
// coefficient array altogether: one array pass less for the HBM-bound kernel).
// dz_off: element offset of the z-derivative scratch relative to the field index (0 for the
// full-size scratch of the two-kernel path, the ring slot offset in the fused kernel).
// DISC: compiled with the scratch-line discard of the fused kernel (run-time switch p.dz_discard).
template <typename T, bool CPLX, int N, int CM, bool FAST, bool DISC = false>
__device__ __forceinline__ void yline_phase_b(const UpdParams& p, const int i, const int k0, const unsigned mask,
                                              const int upd, const typename Cx<T>::type* xbuf, const double cuni,
                                              const long long dz_off) {
    using C = typename Cx<T>::type;
    using A = typename AccT<CPLX>::type;
    using VV = Vec<T, CPLX>;
    using S = YCfg<T, CPLX, N>;
    constexpr int W = S::W;
    constexpr int V = VV::V;
    const size_t plane = (size_t)p.ny * p.nz;
    const int in = i + p.dir;                       // x neighbour plane
    const bool nb_inside = (in >= 0 && in < p.nx);
    const bool nb_any = nb_inside || p.halo[0] != nullptr;
    const void* nFy = nb_inside ? p.F[1] : p.halo[0];
    const void* nFz = nb_inside ? p.F[2] : p.halo[1];
    const size_t nbase = nb_inside ? (size_t)in * plane : 0;
    constexpr int CG = W / V;                    // column groups
    constexpr int RP = S::THREADS / CG;          // rows per pass
    constexpr int NPASS = N / RP;
    // passes whose loads are batched: two when one pass takes 40 registers (f64: 2 cells per
    // vector, c128: 1), one when it takes 80 (f32: 4 cells, c64: 2 complex cells) -- two such
    // passes spilled (f32 SHPF 24.6 -> 32.8 Gcell/s with one)
    constexpr int PB = (NPASS % 2 == 0 && V * (CPLX ? 2 : 1) <= 2) ? 2 : 1;
    const int cg = threadIdx.x % CG, tr = threadIdx.x / CG;
    const int k = k0 + cg * V;
    if (k >= p.nz) return;
    const double sx = p.dir > 0 ? p.rdx : -p.rdx;
#pragma unroll 1
    for (int pass0 = 0; pass0 < NPASS; pass0 += PB) {
        A dz0[PB][V], dz1[PB][V], a3[PB][V], a4[PB][V], b3[PB][V], b4[PB][V], g[PB][3][V];
        double cf[PB][V];
#pragma unroll
        for (int u = 0; u < PB; ++u) {
            const int j = tr + (pass0 + u) * RP;
            const size_t idx = (size_t)i * plane + (size_t)j * p.nz + k;
            VV::ld(p.dz[0], (size_t)((long long)idx + dz_off), dz0[u]);
            VV::ld(p.dz[1], (size_t)((long long)idx + dz_off), dz1[u]);
            if (p.pstd) {
                VV::ld(p.dxs[0], idx, a3[u]);
                VV::ld(p.dxs[1], idx, a4[u]);
            } else if (nb_any) {
                const size_t nidx = nbase + (size_t)j * p.nz + k;
                VV::ld(nFz, nidx, a3[u]);
                VV::ld(nFy, nidx, a4[u]);
                VV::ld(p.F[2], idx, b3[u]);
                VV::ld(p.F[1], idx, b4[u]);
            }
#pragma unroll
            for (int c = 0; c < 3; ++c) VV::ld(p.G[c], idx, g[u][c]);
            if constexpr (CM == 2) {
#pragma unroll
                for (int v = 0; v < V; ++v) cf[u][v] = cuni;
            } else {
                ld_coeff<V, CM == 1>(p, idx, cf[u]);
            }
        }
        // A tile spans all rows, so with CPML on the y faces every tile carries CPML terms; the
        // rows of this batch may still be interior: then they take the straight-line path too
        // (batch-uniform branch around two separately compiled bodies).
        auto compute = [&](auto fastrow, const unsigned bmask) {
            constexpr bool FR = decltype(fastrow)::value;
#pragma unroll
            for (int u = 0; u < PB; ++u) {
                const int j = tr + (pass0 + u) * RP;
#pragma unroll
                for (int v = 0; v < V; ++v) {
                    A d[6];
                    const C r0 = xbuf[(size_t)j * W + stash_col<T, CPLX, W>(cg * V + v)];
                    if constexpr (CPLX) {
                        const C r1 = xbuf[(size_t)N * W + (size_t)j * W + stash_col<T, CPLX, W>(cg * V + v)];
                        d[0] = make_double2((double)r0.x, (double)r0.y);
                        d[5] = make_double2((double)r1.x, (double)r1.y);
                    } else {
                        d[0] = (double)r0.x; d[5] = (double)r0.y;
                    }
                    d[1] = dz0[u][v];
                    d[2] = dz1[u][v];
                    if (p.pstd) { d[3] = a3[u][v]; d[4] = a4[u][v]; }
                    else if (nb_any) {
                        d[3] = a_scale(sx, a_sub(a3[u][v], b3[u][v]));
                        d[4] = a_scale(sx, a_sub(a4[u][v], b4[u][v]));
                    } else { d[3] = a_zero(A()); d[4] = a_zero(A()); }
                    A gg[3] = {g[u][0][v], g[u][1][v], g[u][2][v]};
                    if constexpr (FR) cell_update_fast<CPLX>(upd, cf[u][v], d, gg);
                    else cell_update_regs<T, CPLX>(p, bmask, i, j, k + v, cf[u][v], d, gg);
                    g[u][0][v] = gg[0]; g[u][1][v] = gg[1]; g[u][2][v] = gg[2];
                }
            }
        };
        if constexpr (FAST) {
            compute(std::true_type{}, 0u);
        } else {
            const unsigned bm = term_mask(p, i, i + 1, pass0 * RP, (pass0 + PB) * RP, k0, k0 + W);
            if (upd >= 0 && bm == 0u) compute(std::true_type{}, 0u); else compute(std::false_type{}, bm);
        }
#pragma unroll
        for (int u = 0; u < PB; ++u) {
            const int j = tr + (pass0 + u) * RP;
            const size_t idx = (size_t)i * plane + (size_t)j * p.nz + k;
            // streaming stores (st.global.cs, evict-first): the updated field is not touched again
            // in this half-step; measured 1.29 -> 1.14 ms per launch (loads stay allocating:
            // ld.cs / L1::no_allocate / ld.cg on the streaming operands cost 8-13 %)
#pragma unroll
            for (int c = 0; c < 3; ++c) vst_stream<VV>(p.G[c], idx, g[u][c], 0);
            if constexpr (DISC && !CPLX && sizeof(T) == 8 && W % 16 == 0) {      // (W = 8 at N = 512: a line spans two tiles)
                // fused kernel: this tile was the only reader of its z-derivative scratch lines (a row segment of
                // the tile is whole 128-byte lines, read by lanes of this warp in the load above): drop them from
                // L2 so the dirty lines are never written back to HBM (lines the CPML pass still reads are kept)
                if (p.dz_discard && (cg * V) % 16 == 0 && k >= p.dz_keep_lo && k + 16 <= p.dz_keep_hi) {
                    const size_t e = (size_t)((long long)idx + dz_off);
                    asm volatile("discard.global.L2 [%0], 128;" :: "l"((const double*)p.dz[0] + e) : "memory");
                    asm volatile("discard.global.L2 [%0], 128;" :: "l"((const double*)p.dz[1] + e) : "memory");
                }
            }
        }
    }
}

// Phase A: W adjacent y lines (one per column, lane = column) are transformed in
//          registers; the derivative pair lands in shared memory in tile layout
//          stash[row * W + col].
// Phase B: the CTA re-maps to 16-byte vectors along z (V cells per thread) and streams
//          the cell update: PB row groups of loads are issued before any arithmetic so
//          enough bytes are in flight to cover HBM latency.
// One CTA = the tile (plane i, columns k0 .. k0+W-1, all rows).  PAL: coefficients come
// from the palette form (update_dev.cuh ld_coeff).
// The twiddle / multiplier tables are read straight from global memory (L1-resident,
// 8 KB): that keeps the CTA at N*W*16 B of shared memory, so two CTAs fit the 132 KB
// carve-out and 96 KB of L1 remain for loads in flight (measured: the kernel's
// bandwidth follows the L1 size left by the carve-out).
template <typename T, bool CPLX, int N>
__device__ __forceinline__ void yline_phase_a(const UpdParams& p, const int i, const int k0,
                                              typename Cx<T>::type* xbuf,
                                              const typename Cx<T>::type* __restrict__ tw,
                                              const typename Cx<T>::type* __restrict__ ml) {
    using C = typename Cx<T>::type;
    using F = Fld<T, CPLX>;
    using S = YCfg<T, CPLX, N>;
    constexpr int W = S::W;
    constexpr int NF = F::NF;
    const size_t plane = (size_t)p.ny * p.nz;
    const int c = threadIdx.x % W, t = threadIdx.x / W;
    const int k = k0 + c;
    const bool ok = k < p.nz;
    const size_t pbase = (size_t)i * plane + k;
    // pair (F_z, F_x): Re -> d/dy F_z (slot 0), Im -> d/dy F_x (slot 5)
#pragma unroll
    for (int f = 0; f < NF; ++f) {
        XchgStrided<C, W> xb{xbuf + (size_t)f * N * W + stash_col<T, CPLX, W>(c)};
        C v[16];
#pragma unroll
        for (int q = 0; q < 16; ++q) {
            if (ok) v[q] = F::ld(p.F[2], p.F[0], pbase + (size_t)line_index<N>(t, q) * p.nz, f);
            else { v[q].x = 0; v[q].y = 0; }
        }
        fft_forward<N>(v, t, tw, xb);
#pragma unroll
        for (int q = 0; q < 16; ++q) v[q] = cmul(v[q], ml[spec_index<N>(t, q)]);
        fft_inverse<N>(v, t, tw, xb);
        __syncthreads();                 // everyone finished reading the exchange buffer
#pragma unroll
        for (int q = 0; q < 16; ++q) xb.st(line_index<N>(t, q), v[q]);
        if (p.dy_side && ok) {
            // separate CPML pass (engine.cu k_pml_terms): the y derivatives of the absorber rows go to the
            // side buffer straight from the registers (lane = column: 16-byte stores, contiguous per row).
            // Doing this inside the update phase's load loop cost that phase 12 %.
            C* side = (C*)p.dy_side + (size_t)f * p.nx * p.ys_rows * p.nz + (size_t)i * p.ys_rows * p.nz + k;
#pragma unroll
            for (int q = 0; q < 16; ++q) {
                const int row = line_index<N>(t, q);
                const int jj = row < p.ys_lo_n ? row : (row >= p.ys_hi_0 ? row - p.ys_hi_0 + p.ys_lo_n : -1);
                if (jj >= 0) side[(size_t)jj * p.nz] = v[q];
            }
        }
    }
}

// phase B for the tile (plane i, column block kb of ntk): picks the straight-line / uniform-
// coefficient variants per tile.
template <typename T, bool CPLX, int N, bool PAL, bool DISC = false>
__device__ __forceinline__ void yline_phase_b_dispatch(const UpdParams& p, const int i, const int kb, const int ntk,
                                                       const typename Cx<T>::type* xbuf, const long long dz_off) {
    constexpr int W = YCfg<T, CPLX, N>::W;
    const int k0 = kb * W;
    const unsigned mask = term_mask(p, i, i + 1, 0, p.ny, k0, k0 + W);
    const int upd = tile_update_class(p, i, i + 1, 0, p.ny, k0, min(k0 + W, p.nz));
    // per-tile uniform coefficient (engine.cu: k_tile_uniform), NaN when the tile is not uniform
    const double cuni = p.Ctile ? p.Ctile[(size_t)i * ntk + kb] : __longlong_as_double(0x7ff8000000000000LL);
    const bool fast = mask == 0u && upd >= 0;
    if (fast && cuni == cuni) yline_phase_b<T, CPLX, N, 2, true, DISC>(p, i, k0, mask, upd, xbuf, cuni, dz_off);
    else if (fast) yline_phase_b<T, CPLX, N, (PAL ? 1 : 0), true, DISC>(p, i, k0, mask, upd, xbuf, 0.0, dz_off);
    else if (cuni == cuni) yline_phase_b<T, CPLX, N, 2, false, DISC>(p, i, k0, mask, upd, xbuf, cuni, dz_off);   // CPML tile, one coefficient
    else yline_phase_b<T, CPLX, N, (PAL ? 1 : 0), false, DISC>(p, i, k0, mask, upd, xbuf, 0.0, dz_off);
}

template <typename T, bool CPLX, int N, bool PAL>
__global__ void __launch_bounds__(YCfg<T, CPLX, N>::THREADS, YCfg<T, CPLX, N>::MINB)
k_yline_update(const UpdParams p, const typename Cx<T>::type* __restrict__ tw,
               const typename Cx<T>::type* __restrict__ ml) {
    using C = typename Cx<T>::type;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    C* xbuf = reinterpret_cast<C*>(smem_raw);      // NF buffers of N*W: exchange, then derivative stash
    const int i = p.i0 + (int)blockIdx.y;
    if (p.nterms) prefetch_tile_psi<T, CPLX>(p, i, (int)blockIdx.x * YCfg<T, CPLX, N>::W,
                                             min(((int)blockIdx.x + 1) * YCfg<T, CPLX, N>::W, p.nz));
    yline_phase_a<T, CPLX, N>(p, i, (int)blockIdx.x * YCfg<T, CPLX, N>::W, xbuf, tw, ml);
    __syncthreads();
    yline_phase_b_dispatch<T, CPLX, N, PAL>(p, i, (int)blockIdx.x, (int)gridDim.x, xbuf, p.dz_off);
}

// ------------------------------------------------------------- launchers -----
template <typename K>
static int set_smem(K kernel, size_t bytes) {
    if (bytes > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
        if (e != cudaSuccess) { set_error(std::string("cudaFuncSetAttribute: ") + cudaGetErrorString(e)); return 1; }
    }
    return 0;
}

#define IES_FOR_N(n, MACRO)                                             \
    switch (n) {                                                        \
        case 16: MACRO(16); break;   case 32: MACRO(32); break;         \
        case 64: MACRO(64); break;   case 128: MACRO(128); break;       \
        case 256: MACRO(256); break; case 512: MACRO(512); break;       \
        default: set_error("unsupported FFT length " + std::to_string(n) + \
                           " (SHPF/PSTD axes must be a power of two in 16..512)"); return 1; \
    }

template <typename T, bool CPLX>
int launch_zline(Ctx* c, const void* A, const void* B, void* dA, void* dB, int half, int i0, int i1, int out_i0,
                 cudaStream_t st) {
    using C = typename Cx<T>::type;
    if (!st) st = c->stream;
    const int n = c->cfg.nz;
    const long nlines = (long)(i1 - i0) * c->cfg.ny;
    const long line0 = (long)i0 * c->cfg.ny;
    const long oline0 = (long)out_i0 * c->cfg.ny;
    if (nlines <= 0) return 0;
#define Z_LAUNCH(NN, VV) {                                                                  \
        constexpr int LPB = ZCfg<NN, VV>::LPB;                                              \
        using TW = TwTables<NN, VV>;                                                        \
        size_t sm = sizeof(C) * ((TW::NEED_MASTER ? NN : 0) + (TW::SHARED ? 0 : TW::FWD) + TW::INV + NN + \
                                 (size_t)LPB * ZXchg<C, NN, VV>::type::LS);                  \
        auto kern = k_zline<T, CPLX, NN, VV>;                                               \
        if (set_smem(kern, sm)) return 1;                                                   \
        unsigned grid = (unsigned)((nlines + LPB - 1) / LPB);                               \
        kern<<<grid, ZCfg<NN, VV>::THREADS, sm, st>>>(A, B, dA, dB, line0, oline0, nlines,   \
            (const C*)c->tw[2], (const C*)c->mult[half][2]);                                \
    }
#define Z_CASE(NN) Z_LAUNCH(NN, 16)       // 8 values per thread measured 54 % slower (more exchanges)
    if (st == c->stream) prof_mark(c, PROF_ZLINE, 0);
    IES_FOR_N(n, Z_CASE)
    if (st == c->stream) prof_mark(c, PROF_ZLINE, 1);
#undef Z_CASE
#undef Z_LAUNCH
    count_launch();
    IES_CUDA(cudaGetLastError());
    return 0;
}

// Strided-line derivative of the pair (A, B): axis 0 = x lines of the whole slab (PSTD),
// axis 1 = y lines of the x-planes [i0, i1).
template <typename T, bool CPLX>
int launch_sline(Ctx* c, const void* A, const void* B, void* dA, void* dB, int half, int axis, int i0, int i1) {
    using C = typename Cx<T>::type;
    const size_t es = sizeof(T) * (CPLX ? 2 : 1);
    const long plane = (long)c->cfg.ny * c->cfg.nz;
    const int n = axis == 0 ? c->cfg.nx : c->cfg.ny;
    const long ncols = axis == 0 ? plane : (long)c->cfg.nz;
    const long stride = ncols;
    const long batch_stride = axis == 0 ? 0 : plane;
    const unsigned batches = axis == 0 ? 1u : (unsigned)(i1 - i0);
    if (axis == 1) {
        if (i1 <= i0) return 0;
        const size_t off = (size_t)i0 * plane * es;
        A = (const char*)A + off; B = (const char*)B + off; dA = (char*)dA + off; dB = (char*)dB + off;
    }
#define X_CASE(NN) {                                                                        \
        using S = SCfg<T, CPLX, NN>;                                                        \
        size_t sm = sizeof(C) * (2 * NN + (size_t)NN * S::W);                               \
        auto kern = k_xline<T, CPLX, NN>;                                                   \
        if (set_smem(kern, sm)) return 1;                                                   \
        dim3 grid((unsigned)((ncols + S::W - 1) / S::W), batches);                          \
        kern<<<grid, S::THREADS, sm, c->stream>>>(A, B, dA, dB, ncols, stride, batch_stride, \
            (const C*)c->tw[axis], (const C*)c->mult[half][axis]);                          \
    }
    prof_mark(c, PROF_XLINE, 0);
    IES_FOR_N(n, X_CASE)
    prof_mark(c, PROF_XLINE, 1);
#undef X_CASE
    count_launch();
    IES_CUDA(cudaGetLastError());
    return 0;
}

template <typename T, bool CPLX>
int launch_yline_update(Ctx* c, const UpdParams& p, int half) {
    using C = typename Cx<T>::type;
    const int n = c->cfg.ny;
    if (p.i1 <= p.i0) return 0;
    const bool pal = p.Cidx != nullptr;
#define Y_CASE(NN) {                                                                        \
        using S = YCfg<T, CPLX, NN>;                                                        \
        size_t sm = sizeof(C) * ((size_t)NN * S::W * Fld<T, CPLX>::NF);                     \
        auto kern = pal ? k_yline_update<T, CPLX, NN, true> : k_yline_update<T, CPLX, NN, false>; \
        if (set_smem(kern, sm)) return 1;                                                   \
        dim3 grid((unsigned)((c->cfg.nz + S::W - 1) / S::W), (unsigned)(p.i1 - p.i0));      \
        kern<<<grid, S::THREADS, sm, c->stream>>>(p, (const C*)c->tw[1],                     \
            (const C*)c->mult[half][1]);                                                    \
    }
    prof_mark(c, PROF_YLINE_UPDATE, 0);
    IES_FOR_N(n, Y_CASE)
    prof_mark(c, PROF_YLINE_UPDATE, 1);
#undef Y_CASE
    count_launch();
    IES_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace ies
