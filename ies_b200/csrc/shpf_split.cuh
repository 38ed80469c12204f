// SHPF half-step as two kernels with the cell update SPLIT between them
// (space.py:709-727 + 801-811 for updateH, 953-970 + 1017-1025 for updateE, CPML 1110-1712).
//
// The three curl components need different derivatives:
//     G_x += C (d/dy F_z - d/dz F_y)      G_y += C (d/dz F_x - d/dx F_z)      G_z += C (d/dx F_y - d/dy F_x)
// so G_y needs only the z-line derivative (plus the x difference) and G_z only the y-line
// derivative; G_x is the one component that needs both.  Therefore
//
//   k_zline_update (tiles = NL whole z lines of one x-plane)
//      A: d/dz (F_y, F_x) of the tile by FFT                     -> shared-memory stash
//      B: stream: G_y += C (d/dz F_x - d/dx F_z) (+ its CPML terms), d/dz F_y -> scratch dz[0]
//   k_yline_update<SPLIT> (spectral.cuh; tiles = W whole y lines of one x-plane)
//      A: d/dy (F_z, F_x) -> stash;   B: G_x and G_z from the stash, dz[0] and d/dx F_y
//
// Per half-step and cell the pair moves 16 array passes (z: read F_x F_y F_z G_y C, write G_y
// dz0; y: read F_x F_y F_z dz0 G_x G_z C, write G_x G_z) like the z-line + full y-line pair, but
// the z-line kernel -- bound by the FFT arithmetic, HBM half idle -- now carries 7 of them and the
// HBM-bound y-line kernel 9 instead of 12.  Every component sees the same expression with the
// same operands as in the unsplit update, so the fields are bit-identical.
#pragma once
#include "spectral.cuh"

namespace ies {

// Phase B of k_zline_update.  FAST: interior tile (no CPML term, update box resolved per tile).
template <typename T, bool CPLX, int N, bool PAL, bool FAST>
__device__ __forceinline__ void zline_phase_b(const UpdParams& p, const int i, const int j0,
                                              const unsigned mask, const int upd, const typename Cx<T>::type* stash) {
    using C = typename Cx<T>::type;
    using A = typename AccT<CPLX>::type;
    using S = typename Elem<T, CPLX>::S;
    using VV = Vec<T, CPLX>;
    constexpr int TT = N / 16;
    constexpr int NL = 256 / TT;                 // lines (rows) per tile
    constexpr int V = VV::V;
    constexpr int TILE = N * NL;
    constexpr int CGN = N / V;                   // vectors per tile row
    constexpr int NIT = TILE / (V * 256);        // vectors per thread (16 / V)
    constexpr int PB = (NIT % 2 == 0) ? 2 : 1;   // iterations whose loads are batched
    static_assert(N % V == 0, "vector width must divide the line");
    const int tid = threadIdx.x;
    const size_t plane = (size_t)p.ny * p.nz;
    const int in = i + p.dir;                    // x neighbour plane
    const bool nb_inside = (in >= 0 && in < p.nx);
    const bool nb_any = nb_inside || p.halo[0] != nullptr;
    const void* nFz = nb_inside ? p.F[2] : p.halo[1];
    const size_t nbase = nb_inside ? (size_t)in * plane : 0;
    const double sx = p.dir > 0 ? p.rdx : -p.rdx;
    void* SA = const_cast<void*>(p.dz[0]);
#pragma unroll 1
    for (int it0 = 0; it0 < NIT; it0 += PB) {
        A a3[PB][V], b3[PB][V], g[PB][V];
        double cf[PB][V];
        bool ok[PB];
#pragma unroll
        for (int u = 0; u < PB; ++u) {
            const int e = tid + (it0 + u) * 256;
            const int r = e / CGN, cc = (e % CGN) * V;
            const int j = j0 + r;
            ok[u] = j < p.ny;
            if (!ok[u]) continue;
            const size_t idx = (size_t)i * plane + (size_t)j * p.nz + cc;
            if (nb_any) {
                VV::ld(nFz, nbase + (size_t)j * p.nz + cc, a3[u]);
                VV::ld(p.F[2], idx, b3[u]);
            }
            VV::ld(p.G[1], idx, g[u]);
            ld_coeff<V, PAL>(p, idx, cf[u]);
        }
#pragma unroll
        for (int u = 0; u < PB; ++u) {
            if (!ok[u]) continue;
            const int e = tid + (it0 + u) * 256;
            const int r = e / CGN, cc = (e % CGN) * V;
            const int j = j0 + r;
            const size_t idx = (size_t)i * plane + (size_t)j * p.nz + cc;
            A dzy[V];
#pragma unroll
            for (int v = 0; v < V; ++v) {
                const int slot = r * N + XchgContigSw<C, N>::phys(cc + v);
                A d[6];
                const C r0 = stash[slot];
                if constexpr (CPLX) {
                    const C r1 = stash[TILE + slot];
                    dzy[v] = make_double2((double)r0.x, (double)r0.y);
                    d[2] = make_double2((double)r1.x, (double)r1.y);
                } else {
                    dzy[v] = (double)r0.x; d[2] = (double)r0.y;
                }
                d[0] = a_zero(A()); d[1] = a_zero(A()); d[4] = a_zero(A()); d[5] = a_zero(A());
                d[3] = nb_any ? a_scale(sx, a_sub(a3[u][v], b3[u][v])) : a_zero(A());
                A gg[3] = {a_zero(A()), g[u][v], a_zero(A())};
                if constexpr (FAST) cell_update_fast<CPLX, 2>(upd, cf[u][v], d, gg);
                else cell_update_regs<T, CPLX, 2>(p, mask, i, j, cc + v, cf[u][v], d, gg);
                g[u][v] = gg[1];
            }
            VV::st(SA, idx, dzy);                // d/dz F_y for the y-line kernel (exact: stash precision)
            VV::st(p.G[1], idx, g[u]);
        }
    }
    (void)sizeof(S);
}

template <typename T, bool CPLX, int N, bool PAL>
__global__ void __launch_bounds__(256, (CPLX ? 1 : 2))
k_zline_update(const UpdParams p, const typename Cx<T>::type* __restrict__ tw,
               const typename Cx<T>::type* __restrict__ ml) {
    using C = typename Cx<T>::type;
    using F = Fld<T, CPLX>;
    constexpr int TT = N / 16;                   // threads per line
    constexpr int NL = 256 / TT;                 // lines per tile
    constexpr int NF = F::NF;
    constexpr int TILE = N * NL;                 // 4096 cells
    extern __shared__ __align__(16) unsigned char smem_raw[];
    C* stash = reinterpret_cast<C*>(smem_raw);   // NF buffers of TILE elements: exchange, then stash
    const int i = p.i0 + (int)blockIdx.y;
    const int j0 = (int)blockIdx.x * NL;
    const size_t plane = (size_t)p.ny * p.nz;
    const int tid = threadIdx.x;
    const int t = tid % TT, l = tid / TT;
    const bool line_ok = j0 + l < p.ny;
    const size_t lbase = (size_t)i * plane + (size_t)(j0 + l) * p.nz;
    // ---------------- phase A: d/dz of the pair (F_y, F_x) ----------------
#pragma unroll
    for (int f = 0; f < NF; ++f) {
        C v[16];
#pragma unroll
        for (int q = 0; q < 16; ++q) {
            if (line_ok) v[q] = F::ld(p.F[1], p.F[0], lbase + (size_t)line_index<N>(t, q), f);
            else { v[q].x = 0; v[q].y = 0; }
        }
        XchgContigSw<C, N> xb{stash + (size_t)f * TILE + (size_t)l * N};
        fft_forward<N>(v, t, tw, xb);
#pragma unroll
        for (int q = 0; q < 16; ++q) v[q] = cmul(v[q], ml[spec_index<N>(t, q)]);
        fft_inverse<N>(v, t, tw, xb);
        xb.sync();
#pragma unroll
        for (int q = 0; q < 16; ++q) xb.st(line_index<N>(t, q), v[q]);
    }
    __syncthreads();
    // ---------------- phase B: G_y update + d/dz F_y to the scratch ----------------
    const unsigned mask = term_mask(p, i, i + 1, j0, j0 + NL, 0, p.nz);
    const int upd = tile_update_class(p, i, i + 1, j0, min(j0 + NL, p.ny), 0, p.nz);
    if (mask == 0u && upd >= 0) zline_phase_b<T, CPLX, N, PAL, true>(p, i, j0, mask, upd, stash);
    else zline_phase_b<T, CPLX, N, PAL, false>(p, i, j0, mask, upd, stash);
}

template <typename T, bool CPLX>
int launch_zline_update(Ctx* c, const UpdParams& p, int half) {
    using C = typename Cx<T>::type;
    if (p.i1 <= p.i0) return 0;
    const int n = c->cfg.nz;
    const C* tw = (const C*)c->tw[2];
    const C* ml = (const C*)c->mult[half][2];
    const size_t sm = sizeof(C) * 4096 * Fld<T, CPLX>::NF;
    const bool pal = p.Cidx != nullptr;
#define ZU_CASE(NN) {                                                                       \
        constexpr int NL = 256 / (NN / 16);                                                 \
        dim3 grid((unsigned)((c->cfg.ny + NL - 1) / NL), (unsigned)(p.i1 - p.i0));          \
        auto kern = pal ? k_zline_update<T, CPLX, NN, true> : k_zline_update<T, CPLX, NN, false>; \
        if (set_smem(kern, sm)) return 1;                                                   \
        kern<<<grid, 256, sm, c->stream>>>(p, tw, ml);                                      \
    }
    prof_mark(c, PROF_ZLINE, 0);
    IES_FOR_N(n, ZU_CASE)
    prof_mark(c, PROF_ZLINE, 1);
#undef ZU_CASE
    count_launch();
    IES_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace ies
