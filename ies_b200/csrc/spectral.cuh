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
    template <bool HINT> static __device__ __forceinline__ C ldp(const void* A, const void* B, size_t i, int, uint64_t pol) {
        C v; v.x = ld_pol<HINT>((const T*)A + i, pol); v.y = ld_pol<HINT>((const T*)B + i, pol); return v;
    }
    template <bool HINT> static __device__ __forceinline__ void stp(void* dA, void* dB, size_t i, C v, int, uint64_t pol) {
        st_pol<HINT>((T*)dA + i, v.x, pol); st_pol<HINT>((T*)dB + i, v.y, pol);
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
    template <bool HINT> static __device__ __forceinline__ C ldp(const void* A, const void* B, size_t i, int f, uint64_t pol) {
        return ld_pol<HINT>((const C*)(f == 0 ? A : B) + i, pol);
    }
    template <bool HINT> static __device__ __forceinline__ void stp(void* dA, void* dB, size_t i, C v, int f, uint64_t pol) {
        st_pol<HINT>((C*)(f == 0 ? dA : dB) + i, v, pol);
    }
};

template <typename C>
__device__ __forceinline__ void load_tables(C* tw, C* ml, const C* twg, const C* mlg, int n) {
    for (int i = threadIdx.x; i < n; i += blockDim.x) { tw[i] = twg[i]; ml[i] = mlg[i]; }
    __syncthreads();
}

// ---------------------------------------------------------------- z lines -----
template <int N> struct ZCfg {
    static constexpr int TT = N / 16;
    static constexpr int LPB = (256 / TT) > 0 ? (256 / TT) : 1;   // lines per block
    static constexpr int THREADS = LPB * TT;
};

__device__ __forceinline__ int ld_acquire(const int* p) {
    int v; asm volatile("ld.global.acquire.gpu.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v;
}
__device__ __forceinline__ void red_release(int* p) {
    asm volatile("red.global.release.gpu.add.s32 [%0], 1;" :: "l"(p) : "memory");
}
// thread 0 spins until *ctr >= need (no-op when ctr is null); callers follow with a CTA barrier
__device__ __forceinline__ void wait_counter(const int* ctr, int need) {
    if (need > 0 && threadIdx.x == 0) {
        while (ld_acquire(ctr) < need) __nanosleep(32);
    }
}

// One work item = LPB adjacent z lines.  Input lines start at in_line0 (+ item*LPB), the
// derivative lines go to out_line0 (+ item*LPB) of the scratch arrays.
template <typename T, bool CPLX, int N, bool HINT = false>
__device__ __forceinline__ void zline_item(const void* __restrict__ A, const void* __restrict__ B,
                                           void* __restrict__ dA, void* __restrict__ dB,
                                           long in_line0, long out_line0, long nlines, long item,
                                           const typename Cx<T>::type* tw, const typename Cx<T>::type* ml,
                                           typename Cx<T>::type* xbuf,
                                           const int* dep_ctr = nullptr, int dep_need = 0,
                                           uint64_t pin = 0, uint64_t pout = 0) {
    using C = typename Cx<T>::type;
    using F = Fld<T, CPLX>;
    constexpr int TT = ZCfg<N>::TT, LPB = ZCfg<N>::LPB;
    const int t = threadIdx.x % TT, l = threadIdx.x / TT;
    const long line = item * LPB + l;
    const bool ok = line < nlines;
    const size_t ibase = (size_t)(in_line0 + line) * N;
    const size_t obase = (size_t)(out_line0 + line) * N;
    XchgContig<C, N> xb{xbuf + (size_t)l * XchgContig<C, N>::LS};
#pragma unroll 1
    for (int f = 0; f < F::NF; ++f) {
        C v[16];
#pragma unroll
        for (int q = 0; q < 16; ++q) {
            if (ok) v[q] = F::template ldp<HINT>(A, B, ibase + line_index<N>(t, q), f, pin);
            else { v[q].x = 0; v[q].y = 0; }
        }
        fft_forward<N>(v, t, tw, xb);
#pragma unroll
        for (int q = 0; q < 16; ++q) v[q] = cmul(v[q], ml[spec_index<N>(t, q)]);
        fft_inverse<N>(v, t, tw, xb);
        if (dep_need > 0 && f == 0) {            // output slot free? (fused kernel only; CTA-uniform)
            wait_counter(dep_ctr, dep_need);
            __syncthreads();
        }
        if (ok) {
#pragma unroll
            for (int q = 0; q < 16; ++q) F::template stp<HINT>(dA, dB, obase + line_index<N>(t, q), v[q], f, pout);
        }
    }
}

template <typename T, bool CPLX, int N, bool HINT>
__global__ void __launch_bounds__(ZCfg<N>::THREADS)
k_zline(const void* __restrict__ A, const void* __restrict__ B, void* __restrict__ dA,
        void* __restrict__ dB, long line0, long oline0, long nlines,
        const typename Cx<T>::type* __restrict__ twg, const typename Cx<T>::type* __restrict__ mlg,
        const int pol_in, const int pol_out) {
    using C = typename Cx<T>::type;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    C* tw = reinterpret_cast<C*>(smem_raw);
    C* ml = tw + N;
    C* xbuf = ml + N;
    load_tables(tw, ml, twg, mlg, N);
    uint64_t pin = 0, pout = 0;
    if constexpr (HINT) { pin = make_policy(pol_in); pout = make_policy(pol_out); }
    zline_item<T, CPLX, N, HINT>(A, B, dA, dB, line0, oline0, nlines, (long)blockIdx.x, tw, ml, xbuf,
                                 nullptr, 0, pin, pout);
}

// ------------------------------------------------------- strided lines (x) -----
template <typename T, bool CPLX, int N> struct SCfg {
    static constexpr int TT = N / 16;
    static constexpr int ES = (int)sizeof(T) * (CPLX ? 2 : 1);
    static constexpr int THREADS = 256;
    static constexpr int W = THREADS / TT;                        // adjacent lines (columns) per CTA
    static_assert(W * ES >= 32, "row segment below one sector");
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
            for (int q = 0; q < 16; ++q) F::st(dA, dB, boff + (size_t)line_index<N>(t, q) * stride + col, v[q], f);
        }
    }
}

// ------------------------------------------ y lines + fused field update -----
// Phase A: W adjacent y lines (one per column, lane = column) are transformed in
//          registers; the derivative pair lands in shared memory in tile layout
//          stash[row * W + col].
// Phase B: the CTA re-maps to 16-byte vectors along z (V cells per thread) and streams
//          the cell update: PB row groups of loads are issued before any arithmetic so
//          enough bytes are in flight to cover HBM latency.
// One work item = the tile (plane i, columns kb*W .. kb*W+W-1, all rows).  dz_off is the
// element offset added to a cell index when reading the z-derivative scratch (0 for
// full-size scratch, ring-slot offset in the fused kernel); DZCG selects ld.global.cg
// for those reads (data produced by other CTAs of the same launch).
template <typename T, bool CPLX, int N, bool DZCG, bool HINT = false>
__device__ __forceinline__ void yline_item(const UpdParams& p, const int i, const int kb, const long long dz_off,
                                           const typename Cx<T>::type* tw, const typename Cx<T>::type* ml,
                                           typename Cx<T>::type* xbuf,
                                           const int* dep_ctr = nullptr, int dep_need = 0) {
    using C = typename Cx<T>::type;
    using F = Fld<T, CPLX>;
    using A = typename AccT<CPLX>::type;
    using VV = Vec<T, CPLX>;
    using S = SCfg<T, CPLX, N>;
    constexpr int W = S::W;
    constexpr int V = VV::V;
    constexpr int NF = F::NF;
    const int k0 = kb * W;
    const size_t plane = (size_t)p.ny * p.nz;
    {
        const int c = threadIdx.x % W, t = threadIdx.x / W;
        const int k = k0 + c;
        const bool ok = k < p.nz;
        const size_t pbase = (size_t)i * plane + k;
        // pair (F_z, F_x): Re -> d/dy F_z (slot 0), Im -> d/dy F_x (slot 5)
#pragma unroll
        for (int f = 0; f < NF; ++f) {
            XchgStrided<C, W> xb{xbuf + (size_t)f * N * W + c};
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
        }
        wait_counter(dep_ctr, dep_need);     // z derivatives of this chunk produced? (fused kernel only)
        __syncthreads();
    }
    // ---------------- phase B: vectorised streaming update ----------------
    constexpr int CG = W / V;                    // column groups
    constexpr int RP = S::THREADS / CG;          // rows per pass
    constexpr int NPASS = N / RP;
    constexpr int PB = (NPASS % 2 == 0) ? 2 : 1; // passes whose loads are batched
    const int cg = threadIdx.x % CG, tr = threadIdx.x / CG;
    const int k = k0 + cg * V;
    if (k >= p.nz) return;
    const unsigned mask = term_mask(p, i, i + 1, 0, p.ny, k0, k0 + W);
    const int in = i + p.dir;                       // x neighbour plane
    const bool nb_inside = (in >= 0 && in < p.nx);
    const bool nb_any = nb_inside || p.halo[0] != nullptr;
    const void* nFy = nb_inside ? p.F[1] : p.halo[0];
    const void* nFz = nb_inside ? p.F[2] : p.halo[1];
    const size_t nbase = nb_inside ? (size_t)in * plane : 0;
    const double sx = p.dir > 0 ? p.rdx : -p.rdx;
    uint64_t pdz = 0, pg = 0;
    if constexpr (HINT) { pdz = make_policy(p.pol_dz); pg = make_policy(p.pol_g); }
#pragma unroll 1
    for (int pass0 = 0; pass0 < NPASS; pass0 += PB) {
        A dz0[PB][V], dz1[PB][V], a3[PB][V], a4[PB][V], b3[PB][V], b4[PB][V], g[PB][3][V];
        double cf[PB][V];
#pragma unroll
        for (int u = 0; u < PB; ++u) {
            const int j = tr + (pass0 + u) * RP;
            const size_t idx = (size_t)i * plane + (size_t)j * p.nz + k;
            if constexpr (DZCG) {
                VV::template ldx<true>(p.dz[0], (size_t)((long long)idx + dz_off), dz0[u]);
                VV::template ldx<true>(p.dz[1], (size_t)((long long)idx + dz_off), dz1[u]);
            } else {
                VV::template ldp<HINT>(p.dz[0], (size_t)((long long)idx + dz_off), dz0[u], pdz);
                VV::template ldp<HINT>(p.dz[1], (size_t)((long long)idx + dz_off), dz1[u], pdz);
            }
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
            for (int c = 0; c < 3; ++c) VV::template ldp<HINT>(p.G[c], idx, g[u][c], pg);
            ld_coeff<V>(p, idx, cf[u]);
        }
#pragma unroll
        for (int u = 0; u < PB; ++u) {
            const int j = tr + (pass0 + u) * RP;
            const size_t idx = (size_t)i * plane + (size_t)j * p.nz + k;
#pragma unroll
            for (int v = 0; v < V; ++v) {
                A d[6];
                const C r0 = xbuf[(size_t)j * W + cg * V + v];
                if constexpr (CPLX) {
                    const C r1 = xbuf[(size_t)N * W + (size_t)j * W + cg * V + v];
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
                cell_update_regs<T, CPLX>(p, mask, i, j, k + v, cf[u][v], d, gg);
                g[u][0][v] = gg[0]; g[u][1][v] = gg[1]; g[u][2][v] = gg[2];
            }
#pragma unroll
            for (int c = 0; c < 3; ++c) VV::template stp<HINT>(p.G[c], idx, g[u][c], pg);
        }
    }
}

template <typename T, bool CPLX, int N, bool HINT>
__global__ void __launch_bounds__(256, (CPLX ? 1 : 2))
k_yline_update(const UpdParams p, const typename Cx<T>::type* __restrict__ twg,
               const typename Cx<T>::type* __restrict__ mlg, const int tables_in_smem) {
    using C = typename Cx<T>::type;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    // The twiddle / multiplier tables are read straight from global memory (L1-resident,
    // 8 KB) by default: that keeps the CTA at N*W*16 B of shared memory, so two CTAs fit the
    // 132 KB carve-out and 96 KB of L1 remain for loads in flight (measured: the kernel's
    // bandwidth follows the L1 size left by the carve-out).
    const C* tw = twg;
    const C* ml = mlg;
    C* xbuf = reinterpret_cast<C*>(smem_raw);      // NF buffers of N*W: exchange, then derivative stash
    if (tables_in_smem) {
        C* stw = reinterpret_cast<C*>(smem_raw);
        C* sml = stw + N;
        xbuf = sml + N;
        load_tables(stw, sml, twg, mlg, N);
        tw = stw; ml = sml;
    }
    yline_item<T, CPLX, N, false, HINT>(p, p.i0 + (int)blockIdx.y, (int)blockIdx.x, p.dz_off, tw, ml, xbuf);
}

// ------------------------------------------------ fused persistent half-step -----
// SHPF half-step as ONE persistent launch: z-line items and y-line/update items are
// pulled from a global queue ordered so that a chunk's z derivatives are produced two
// groups before they are consumed; the scratch is a 3-chunk ring that stays resident in
// the 126 MB L2, so the z derivatives never travel to HBM.  Dependencies are per-chunk
// completion counters (release/acquire at gpu scope); an item is only claimed by a
// running CTA and never waits on an item claimed later, so the scheme cannot deadlock.
struct FusedArgs {
    int total;
    int* ctr;            // [0] queue head, [1 .. nc] z items done, [1+nc .. 2nc] y items done
    int nc, cx, la;      // chunks, planes per chunk, look-ahead of the z items (chunks)
    int slots;           // ring slots of the z-derivative scratch
    int kblocks;         // y-line items per plane
    int zitems_full;     // z items of a full chunk (cx*ny/LPB)
    int debug_skip;      // development: 1 = skip y work, 2 = skip z work
};


template <typename T, bool CPLX, int N>
__global__ void __launch_bounds__(256, (CPLX ? 1 : 2))
k_shpf_fused(const UpdParams p, const FusedArgs fa,
             const typename Cx<T>::type* __restrict__ twz, const typename Cx<T>::type* __restrict__ mlz,
             const typename Cx<T>::type* __restrict__ twy, const typename Cx<T>::type* __restrict__ mly) {
    using C = typename Cx<T>::type;
    static_assert(ZCfg<N>::THREADS == 256 && SCfg<T, CPLX, N>::THREADS == 256, "fused kernel needs 256-thread items");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    C* tz = reinterpret_cast<C*>(smem_raw);
    C* mz = tz + N;
    C* ty = mz + N;
    C* my = ty + N;
    C* xbuf = my + N;
    for (int q = threadIdx.x; q < N; q += blockDim.x) { tz[q] = twz[q]; mz[q] = mlz[q]; ty[q] = twy[q]; my[q] = mly[q]; }
    __shared__ int s_item[2];
    const size_t plane = (size_t)p.ny * p.nz;
    // queue layout (all chunks are full: cx divides nx):
    //   [z(0) .. z(la-1)] then for c = 0 .. nc-1: y(c), z(c+la) (z only while c+la < nc)
    const int zf = fa.zitems_full, yf = fa.cx * fa.kblocks;
    const int head = min(fa.la, fa.nc) * zf;
    const int npair = max(fa.nc - fa.la, 0);            // chunks c that are followed by z(c+la)
    if (threadIdx.x == 0) s_item[0] = atomicAdd(&fa.ctr[0], 1);
    int par = 0;
    while (true) {
        __syncthreads();                       // s_item[par] visible; previous item done with smem
        const int item = s_item[par];
        if (item >= fa.total) break;
        if (threadIdx.x == 0) s_item[par ^ 1] = atomicAdd(&fa.ctr[0], 1);   // prefetch the next claim
        par ^= 1;
        int type, c, local;
        if (item < head) { type = 0; c = item / zf; local = item - c * zf; }
        else {
            const int r = item - head;
            const int pairs_items = npair * (yf + zf);
            if (r < pairs_items) {
                const int pr = r / (yf + zf), w = r - pr * (yf + zf);
                if (w < yf) { type = 1; c = pr; local = w; }
                else { type = 0; c = pr + fa.la; local = w - yf; }
            } else {
                const int r2 = r - pairs_items;
                type = 1; c = npair + r2 / yf; local = r2 - (r2 / yf) * yf;
            }
        }
        const int pl0 = c * fa.cx;                                   // first plane of the chunk
        const int slot = c % fa.slots;
        if (type == 0) {
            // the ring slot must have been consumed by the y items of chunk c-slots
            if (fa.debug_skip == 2) { if (threadIdx.x == 0) red_release(&fa.ctr[1 + c]); continue; }
            const bool wrap = c >= fa.slots;
            const int* dep = &fa.ctr[wrap ? 1 + fa.nc + (c - fa.slots) : 0];
            zline_item<T, CPLX, N>(p.F[1], p.F[0], const_cast<void*>(p.dz[0]), const_cast<void*>(p.dz[1]),
                                   (long)pl0 * p.ny, (long)slot * fa.cx * p.ny, (long)fa.cx * p.ny, (long)local,
                                   tz, mz, xbuf, dep, wrap ? yf : 0);
            __syncthreads();
            if (threadIdx.x == 0) { __threadfence(); red_release(&fa.ctr[1 + c]); }
        } else {
            if (fa.debug_skip == 1) { if (threadIdx.x == 0) red_release(&fa.ctr[1 + fa.nc + c]); continue; }
            const int i = pl0 + local / fa.kblocks;
            const int kb = local - (local / fa.kblocks) * fa.kblocks;
            const long long dz_off = ((long long)slot * fa.cx - pl0) * (long long)plane;
            yline_item<T, CPLX, N, true>(p, i, kb, dz_off, ty, my, xbuf, &fa.ctr[1 + c], zf);
            __syncthreads();
            if (threadIdx.x == 0) { __threadfence(); red_release(&fa.ctr[1 + fa.nc + c]); }
        }
    }
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
    const bool hint = (c->pol_zin | c->pol_zout) != 0;
    const int n = c->cfg.nz;
    const long nlines = (long)(i1 - i0) * c->cfg.ny;
    const long line0 = (long)i0 * c->cfg.ny;
    const long oline0 = (long)out_i0 * c->cfg.ny;
    if (nlines <= 0) return 0;
#define Z_CASE(NN) {                                                                        \
        constexpr int LPB = ZCfg<NN>::LPB;                                                  \
        size_t sm = sizeof(C) * (2 * NN + (size_t)LPB * XchgContig<C, NN>::LS);             \
        auto kern = hint ? k_zline<T, CPLX, NN, true> : k_zline<T, CPLX, NN, false>;        \
        if (set_smem(kern, sm)) return 1;                                                   \
        unsigned grid = (unsigned)((nlines + LPB - 1) / LPB);                               \
        kern<<<grid, ZCfg<NN>::THREADS, sm, st>>>(A, B, dA, dB, line0, oline0, nlines,       \
            (const C*)c->tw[2], (const C*)c->mult[half][2], c->pol_zin, c->pol_zout);       \
    }
    if (st == c->stream) prof_mark(c, PROF_ZLINE, 0);
    IES_FOR_N(n, Z_CASE)
    if (st == c->stream) prof_mark(c, PROF_ZLINE, 1);
#undef Z_CASE
    count_launch();
    IES_CUDA(cudaGetLastError());
    return 0;
}

// Strided-line derivative of the pair (A, B): axis 0 = x lines of the whole slab (PSTD),
// axis 1 = y lines of the x-planes [i0, i1) (refresh of the alternating SHPF path's scratch).
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
    const bool hint = (p.pol_dz | p.pol_g) != 0;
    int tsm = 0;
    if (const char* e = getenv("IES_B200_TABLES_SMEM")) tsm = atoi(e);
#define Y_CASE(NN) {                                                                        \
        using S = SCfg<T, CPLX, NN>;                                                        \
        size_t sm = sizeof(C) * ((tsm ? 2 * NN : 0) + (size_t)NN * S::W * Fld<T, CPLX>::NF); \
        auto kern = hint ? k_yline_update<T, CPLX, NN, true> : k_yline_update<T, CPLX, NN, false>; \
        if (set_smem(kern, sm)) return 1;                                                   \
        dim3 grid((unsigned)((c->cfg.nz + S::W - 1) / S::W), (unsigned)(p.i1 - p.i0));      \
        kern<<<grid, S::THREADS, sm, c->stream>>>(p, (const C*)c->tw[1],                     \
            (const C*)c->mult[half][1], tsm);                                               \
    }
    prof_mark(c, PROF_YLINE_UPDATE, 0);
    IES_FOR_N(n, Y_CASE)
    prof_mark(c, PROF_YLINE_UPDATE, 1);
#undef Y_CASE
    count_launch();
    IES_CUDA(cudaGetLastError());
    return 0;
}

// Returns 0 = launched, 1 = error, 2 = configuration not covered by the fused kernel.
template <typename T, bool CPLX>
int launch_shpf_fused(Ctx* c, const UpdParams& p, int half) {
    using C = typename Cx<T>::type;
    const int n = c->cfg.ny;
    if (c->cfg.nz != n || n < 64 || n > 512) return 2;
    FusedPlan& fp = c->fused;
    if (!fp.ready) return 2;
    FusedArgs fa;
    fa.total = fp.total;
    fa.ctr = fp.ctr; fa.nc = fp.nc; fa.cx = fp.cx; fa.la = fp.la; fa.slots = fp.slots; fa.kblocks = fp.kblocks; fa.zitems_full = fp.zitems_full;
    fa.debug_skip = 0;
    if (const char* e = getenv("IES_B200_FUSED_SKIP")) fa.debug_skip = atoi(e);
    UpdParams q = p;
    q.dz[0] = fp.ring[0]; q.dz[1] = fp.ring[1];
    IES_CUDA(cudaMemsetAsync(fp.ctr, 0, sizeof(int) * (size_t)(1 + 2 * fp.nc), c->stream));
    prof_mark(c, PROF_YLINE_UPDATE, 0);
#define F_CASE(NN) {                                                                        \
        using S = SCfg<T, CPLX, NN>;                                                        \
        size_t xz = (size_t)ZCfg<NN>::LPB * XchgContig<C, NN>::LS;                          \
        size_t xy = (size_t)NN * S::W * Fld<T, CPLX>::NF;                                   \
        size_t sm = sizeof(C) * (4 * NN + (xz > xy ? xz : xy));                             \
        auto kern = k_shpf_fused<T, CPLX, NN>;                                              \
        if (set_smem(kern, sm)) return 1;                                                   \
        if (fp.grid[half] == 0) {                                                           \
            int per = 0, nsm = 0;                                                           \
            IES_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per, kern, 256, sm));   \
            IES_CUDA(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, c->cfg.device)); \
            if (per < 1) { set_error("fused kernel does not fit on an SM"); return 1; }     \
            fp.grid[half] = per * nsm;                                                      \
        }                                                                                   \
        kern<<<fp.grid[half], 256, sm, c->stream>>>(q, fa, (const C*)c->tw[2],               \
            (const C*)c->mult[half][2], (const C*)c->tw[1], (const C*)c->mult[half][1]);    \
    }
    switch (n) {
        case 64: F_CASE(64); break;   case 128: F_CASE(128); break;
        case 256: F_CASE(256); break; case 512: F_CASE(512); break;
        default: return 2;
    }
#undef F_CASE
    prof_mark(c, PROF_YLINE_UPDATE, 1);
    count_launch();
    IES_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace ies
