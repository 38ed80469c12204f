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

template <typename T, bool CPLX, int N>
__global__ void __launch_bounds__(ZCfg<N>::THREADS)
k_zline(const void* __restrict__ A, const void* __restrict__ B, void* __restrict__ dA,
        void* __restrict__ dB, long line0, long nlines,
        const typename Cx<T>::type* __restrict__ twg, const typename Cx<T>::type* __restrict__ mlg) {
    using C = typename Cx<T>::type;
    using F = Fld<T, CPLX>;
    constexpr int TT = ZCfg<N>::TT, LPB = ZCfg<N>::LPB;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    C* tw = reinterpret_cast<C*>(smem_raw);
    C* ml = tw + N;
    C* xbuf = ml + N;
    load_tables(tw, ml, twg, mlg, N);
    const int t = threadIdx.x % TT, l = threadIdx.x / TT;
    const long line = (long)blockIdx.x * LPB + l;
    const bool ok = line < nlines;
    const size_t base = (size_t)(line0 + line) * N;
    XchgContig<C, N> xb{xbuf + (size_t)l * XchgContig<C, N>::LS};
#pragma unroll 1
    for (int f = 0; f < F::NF; ++f) {
        C v[16];
#pragma unroll
        for (int q = 0; q < 16; ++q) {
            if (ok) v[q] = F::ld(A, B, base + line_index<N>(t, q), f);
            else { v[q].x = 0; v[q].y = 0; }
        }
        fft_forward<N>(v, t, tw, xb);
#pragma unroll
        for (int q = 0; q < 16; ++q) v[q] = cmul(v[q], ml[spec_index<N>(t, q)]);
        fft_inverse<N>(v, t, tw, xb);
        if (ok) {
#pragma unroll
            for (int q = 0; q < 16; ++q) F::st(dA, dB, base + line_index<N>(t, q), v[q], f);
        }
    }
}

// ------------------------------------------------------- strided lines (x) -----
template <typename T, bool CPLX, int N> struct SCfg {
    static constexpr int TT = N / 16;
    static constexpr int ES = (int)sizeof(T) * (CPLX ? 2 : 1);
    static constexpr int WMIN = 128 / ES;                         // one 128-byte segment per row
    static constexpr int W = (128 / TT) > WMIN ? (128 / TT) : WMIN;
    static constexpr int THREADS = W * TT;
};

template <typename T, bool CPLX, int N>
__global__ void __launch_bounds__(SCfg<T, CPLX, N>::THREADS)
k_xline(const void* __restrict__ A, const void* __restrict__ B, void* __restrict__ dA,
        void* __restrict__ dB, long ncols, long stride,
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
    XchgStrided<C, W> xb{xbuf + c};
#pragma unroll 1
    for (int f = 0; f < F::NF; ++f) {
        C v[16];
#pragma unroll
        for (int q = 0; q < 16; ++q) {
            if (ok) v[q] = F::ld(A, B, (size_t)line_index<N>(t, q) * stride + col, f);
            else { v[q].x = 0; v[q].y = 0; }
        }
        fft_forward<N>(v, t, tw, xb);
#pragma unroll
        for (int q = 0; q < 16; ++q) v[q] = cmul(v[q], ml[spec_index<N>(t, q)]);
        fft_inverse<N>(v, t, tw, xb);
        if (ok) {
#pragma unroll
            for (int q = 0; q < 16; ++q) F::st(dA, dB, (size_t)line_index<N>(t, q) * stride + col, v[q], f);
        }
    }
}

// ------------------------------------------ y lines + fused field update -----
template <bool CPLX> struct Conv;
template <> struct Conv<false> {       // real fields: one packed FFT gave both derivatives
    template <typename C>
    static __device__ __forceinline__ void set(double& d0, double& d5, const C (&res)[1][16], int q) {
        d0 = (double)res[0][q].x; d5 = (double)res[0][q].y;
    }
};
template <> struct Conv<true> {
    template <typename C>
    static __device__ __forceinline__ void set(double2& d0, double2& d5, const C (&res)[2][16], int q) {
        d0 = make_double2((double)res[0][q].x, (double)res[0][q].y);
        d5 = make_double2((double)res[1][q].x, (double)res[1][q].y);
    }
};

template <typename T, bool CPLX, int N>
__global__ void __launch_bounds__(SCfg<T, CPLX, N>::THREADS)
k_yline_update(const UpdParams p, const typename Cx<T>::type* __restrict__ twg,
               const typename Cx<T>::type* __restrict__ mlg) {
    using C = typename Cx<T>::type;
    using F = Fld<T, CPLX>;
    using A = typename AccT<CPLX>::type;
    using E = Elem<T, CPLX>;
    constexpr int W = SCfg<T, CPLX, N>::W;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    C* tw = reinterpret_cast<C*>(smem_raw);
    C* ml = tw + N;
    C* xbuf = ml + N;
    load_tables(tw, ml, twg, mlg, N);
    const int c = threadIdx.x % W, t = threadIdx.x / W;
    const int k0 = blockIdx.x * W;
    const int k = k0 + c;
    const int i = p.i0 + blockIdx.y;
    const bool ok = k < p.nz;
    const size_t plane = (size_t)p.ny * p.nz;
    const size_t pbase = (size_t)i * plane + k;
    XchgStrided<C, W> xb{xbuf + c};
    // pair (F_z, F_x): Re -> d/dy F_z (slot 0), Im -> d/dy F_x (slot 5)
    C res[F::NF][16];
#pragma unroll
    for (int f = 0; f < F::NF; ++f) {
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
#pragma unroll
        for (int q = 0; q < 16; ++q) res[f][q] = v[q];
    }
    if (!ok) return;
    const unsigned mask = term_mask(p, i, i + 1, 0, p.ny, k0, k0 + W);
    const int in = i + p.dir;                       // x neighbour plane
    const bool nb_inside = (in >= 0 && in < p.nx);
    const bool nb_halo = !nb_inside && p.halo[0] != nullptr;
#pragma unroll
    for (int q = 0; q < 16; ++q) {
        const int j = line_index<N>(t, q);
        const size_t idx = pbase + (size_t)j * p.nz;
        A d[6];
        Conv<CPLX>::set(d[0], d[5], res, q);
        d[1] = E::ld(p.dz[0], idx);
        d[2] = E::ld(p.dz[1], idx);
        if (p.pstd) {
            d[3] = E::ld(p.dxs[0], idx);
            d[4] = E::ld(p.dxs[1], idx);
        } else if (nb_inside || nb_halo) {
            const size_t nidx = nb_inside ? (size_t)in * plane + (size_t)j * p.nz + k
                                          : (size_t)j * p.nz + k;
            const void* Fy = nb_inside ? p.F[1] : p.halo[0];
            const void* Fz = nb_inside ? p.F[2] : p.halo[1];
            const double s = p.dir > 0 ? p.rdx : -p.rdx;
            d[3] = a_scale(s, a_sub(E::ld(Fz, nidx), E::ld(p.F[2], idx)));
            d[4] = a_scale(s, a_sub(E::ld(Fy, nidx), E::ld(p.F[1], idx)));
        } else {
            d[3] = a_zero(A());
            d[4] = a_zero(A());
        }
        cell_update<T, CPLX>(p, mask, i, j, k, d);
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
int launch_zline(Ctx* c, const void* A, const void* B, void* dA, void* dB, int half, int i0, int i1) {
    using C = typename Cx<T>::type;
    const int n = c->cfg.nz;
    const long nlines = (long)(i1 - i0) * c->cfg.ny;
    const long line0 = (long)i0 * c->cfg.ny;
    if (nlines <= 0) return 0;
#define Z_CASE(NN) {                                                                        \
        constexpr int LPB = ZCfg<NN>::LPB;                                                  \
        size_t sm = sizeof(C) * (2 * NN + (size_t)LPB * XchgContig<C, NN>::LS);             \
        auto kern = k_zline<T, CPLX, NN>;                                                   \
        if (set_smem(kern, sm)) return 1;                                                   \
        unsigned grid = (unsigned)((nlines + LPB - 1) / LPB);                               \
        kern<<<grid, ZCfg<NN>::THREADS, sm, c->stream>>>(A, B, dA, dB, line0, nlines,        \
            (const C*)c->tw[2], (const C*)c->mult[half][2]);                                \
    }
    prof_mark(c, PROF_ZLINE, 0);
    IES_FOR_N(n, Z_CASE)
    prof_mark(c, PROF_ZLINE, 1);
#undef Z_CASE
    count_launch();
    IES_CUDA(cudaGetLastError());
    return 0;
}

template <typename T, bool CPLX>
int launch_xline(Ctx* c, const void* A, const void* B, void* dA, void* dB, int half) {
    using C = typename Cx<T>::type;
    const int n = c->cfg.nx;
    const long ncols = (long)c->cfg.ny * c->cfg.nz;
#define X_CASE(NN) {                                                                        \
        using S = SCfg<T, CPLX, NN>;                                                        \
        size_t sm = sizeof(C) * (2 * NN + (size_t)NN * S::W);                               \
        auto kern = k_xline<T, CPLX, NN>;                                                   \
        if (set_smem(kern, sm)) return 1;                                                   \
        unsigned grid = (unsigned)((ncols + S::W - 1) / S::W);                              \
        kern<<<grid, S::THREADS, sm, c->stream>>>(A, B, dA, dB, ncols, ncols,                \
            (const C*)c->tw[0], (const C*)c->mult[half][0]);                                \
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
#define Y_CASE(NN) {                                                                        \
        using S = SCfg<T, CPLX, NN>;                                                        \
        size_t sm = sizeof(C) * (2 * NN + (size_t)NN * S::W);                               \
        auto kern = k_yline_update<T, CPLX, NN>;                                            \
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
