// SHPF half-step as ONE kernel with alternating tile orientation
// (space.py:709-727 + 801-811 for updateH, 953-970 + 1017-1025 for updateE, CPML 1110-1712).
//
// The two spectral axes need whole lines along y and along z, so a single tile can only
// supply one of the two derivative pairs of its own cells.  The other pair is produced one
// half-step EARLIER by the kernel that updated the field being differentiated:
//
//   updateH  (tiles = LPB whole z lines of one x-plane, "ORI_Z")
//      A: d/dz (E_y, E_x) of the tile by FFT                -> shared-memory stash
//      B: stream the cell update: d/dy (E_z, E_x) from the scratch, x differences from the
//         neighbour plane, curl, H += C*curl, CPML; new (H_y, H_x) replace the stash
//      C: d/dz (H_y, H_x) of the tile by FFT (updateE's multiplier) -> scratch, in place
//   updateE  (tiles = W whole y lines of one x-plane, "ORI_Y")
//      A: d/dy (H_z, H_x) -> stash;  B: update with d/dz (H_y, H_x) from the scratch; new
//         (E_z, E_x) replace the stash;  C: d/dy (E_z, E_x) (updateH's multiplier) -> scratch
//
// Per half-step every cell moves 14 array passes through HBM (read F x3, G x3, C, scratch x2;
// write G x3, scratch x2) instead of the 16 of the z-line + y-line kernel pair, in one launch.
// The scratch always describes the fields as they were when the producing kernel ran: the
// host layer tracks writes that happen in between (put_src, set_field) and refreshes the
// affected planes with the stand-alone derivative kernels before the next update.
#pragma once
#include "spectral.cuh"

namespace ies {

enum { ORI_Y = 0, ORI_Z = 1 };
#ifndef IES_HALF_NT
#define IES_HALF_NT 256
#endif
constexpr int HALF_NT = IES_HALF_NT;          // threads per CTA of k_shpf_half (tile = 16 * HALF_NT cells)

// Phase B of k_shpf_half: streaming cell update of the tile; the tile's own derivative pair
// comes from the stash, the other pair from the scratch; the new G pair replaces the stash
// entry.  FAST: interior tile (no CPML term, update boxes resolved per tile).
template <typename T, bool CPLX, int N, int ORI, bool PAL, bool FAST>
__device__ __forceinline__ void half_phase_b(const UpdParams& p, const int i, const int j0, const int k0,
                                             const unsigned mask, const int upd, typename Cx<T>::type* stash) {
    using C = typename Cx<T>::type;
    using A = typename AccT<CPLX>::type;
    using E = Elem<T, CPLX>;
    using VV = Vec<T, CPLX>;
    constexpr int TT = N / 16;
    constexpr int NL = HALF_NT / TT;
    constexpr int V = VV::V;
    constexpr int TILE = N * NL;
    constexpr int COLS = ORI == ORI_Y ? NL : N;
    constexpr int CGN = COLS / V;                // vectors per tile row
    constexpr int NIT = TILE / (V * HALF_NT);    // vectors per thread (16 / V)
    constexpr int PB = (NIT % 2 == 0) ? 2 : 1;   // iterations whose loads are batched
    const int tid = threadIdx.x;
    const size_t plane = (size_t)p.ny * p.nz;
    auto cell_slot = [](int r, int cc) -> int {
        if (ORI == ORI_Y) return r * NL + cc;
        return r * N + XchgContigSw<C, N>::phys(cc);
    };
    const int in = i + p.dir;                    // x neighbour plane
    const bool nb_inside = (in >= 0 && in < p.nx);
    const bool nb_any = nb_inside || p.halo[0] != nullptr;
    const void* nFy = nb_inside ? p.F[1] : p.halo[0];
    const void* nFz = nb_inside ? p.F[2] : p.halo[1];
    const size_t nbase = nb_inside ? (size_t)in * plane : 0;
    const double sx = p.dir > 0 ? p.rdx : -p.rdx;
#pragma unroll 1
    for (int it0 = 0; it0 < NIT; it0 += PB) {
        A s0[PB][V], s1[PB][V], a3[PB][V], a4[PB][V], b3[PB][V], b4[PB][V], g[PB][3][V];
        double cf[PB][V];
        bool ok[PB];
#pragma unroll
        for (int u = 0; u < PB; ++u) {
            const int e = tid + (it0 + u) * HALF_NT;
            const int r = e / CGN, cc = (e % CGN) * V;
            const int j = j0 + r, k = k0 + cc;
            ok[u] = (j < p.ny) && (k < p.nz);
            if (!ok[u]) continue;
            const size_t idx = (size_t)i * plane + (size_t)j * p.nz + k;
            VV::ld(p.dz[0], idx, s0[u]);
            VV::ld(p.dz[1], idx, s1[u]);
            if (nb_any) {
                const size_t nidx = nbase + (size_t)j * p.nz + k;
                VV::ld(nFz, nidx, a3[u]);
                VV::ld(nFy, nidx, a4[u]);
                VV::ld(p.F[2], idx, b3[u]);
                VV::ld(p.F[1], idx, b4[u]);
            }
#pragma unroll
            for (int c = 0; c < 3; ++c) VV::ld(p.G[c], idx, g[u][c]);
            ld_coeff<V, PAL>(p, idx, cf[u]);
        }
#pragma unroll
        for (int u = 0; u < PB; ++u) {
            if (!ok[u]) continue;
            const int e = tid + (it0 + u) * HALF_NT;
            const int r = e / CGN, cc = (e % CGN) * V;
            const int j = j0 + r, k = k0 + cc;
            const size_t idx = (size_t)i * plane + (size_t)j * p.nz + k;
#pragma unroll
            for (int v = 0; v < V; ++v) {
                const int slot = cell_slot(r, cc + v);
                A d[6];
                A t0, t1;                         // the tile's own derivative pair
                const C r0 = stash[slot];
                if constexpr (CPLX) {
                    const C r1 = stash[TILE + slot];
                    t0 = make_double2((double)r0.x, (double)r0.y);
                    t1 = make_double2((double)r1.x, (double)r1.y);
                } else {
                    t0 = (double)r0.x; t1 = (double)r0.y;
                }
                if (ORI == ORI_Y) { d[0] = t0; d[5] = t1; d[1] = s0[u][v]; d[2] = s1[u][v]; }
                else              { d[1] = t0; d[2] = t1; d[0] = s0[u][v]; d[5] = s1[u][v]; }
                if (nb_any) {
                    d[3] = a_scale(sx, a_sub(a3[u][v], b3[u][v]));
                    d[4] = a_scale(sx, a_sub(a4[u][v], b4[u][v]));
                } else { d[3] = a_zero(A()); d[4] = a_zero(A()); }
                A gg[3] = {g[u][0][v], g[u][1][v], g[u][2][v]};
                if constexpr (FAST) cell_update_fast<CPLX>(upd, cf[u][v], d, gg);
                else cell_update_regs<T, CPLX>(p, mask, i, j, k + v, cf[u][v], d, gg);
                g[u][0][v] = gg[0]; g[u][1][v] = gg[1]; g[u][2][v] = gg[2];
                // the updated pair, as it is stored (field precision), replaces the consumed
                // derivative in the stash: ORI_Y (G_z, G_x), ORI_Z (G_y, G_x)
                const A na = E::rnd(ORI == ORI_Y ? gg[2] : gg[1]);
                const A nb = E::rnd(gg[0]);
                if constexpr (CPLX) {
                    C w0, w1;
                    w0.x = (T)na.x; w0.y = (T)na.y; w1.x = (T)nb.x; w1.y = (T)nb.y;
                    stash[slot] = w0; stash[TILE + slot] = w1;
                } else {
                    C w; w.x = (T)na; w.y = (T)nb;
                    stash[slot] = w;
                }
            }
#pragma unroll
            for (int c = 0; c < 3; ++c) VV::st(p.G[c], idx, g[u][c]);
        }
    }
}

template <typename T, bool CPLX, int N, int ORI, bool PAL>
__global__ void __launch_bounds__(HALF_NT, (CPLX ? 1 : 2) * (256 / HALF_NT))
k_shpf_half(const UpdParams p, const typename Cx<T>::type* __restrict__ tw,
            const typename Cx<T>::type* __restrict__ ml_in, const typename Cx<T>::type* __restrict__ ml_out) {
    using C = typename Cx<T>::type;
    using F = Fld<T, CPLX>;
    using A = typename AccT<CPLX>::type;
    using E = Elem<T, CPLX>;
    using VV = Vec<T, CPLX>;
    constexpr int TT = N / 16;                   // threads per line
    constexpr int NL = HALF_NT / TT;             // lines per tile (W columns for ORI_Y, LPB rows for ORI_Z)
    constexpr int V = VV::V;
    constexpr int NF = F::NF;
    constexpr int TILE = N * NL;                 // 4096 cells
    constexpr int ROWS = ORI == ORI_Y ? N : NL;
    constexpr int COLS = ORI == ORI_Y ? NL : N;
    static_assert(COLS % V == 0, "vector width must divide the tile row");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    C* stash = reinterpret_cast<C*>(smem_raw);   // NF buffers of TILE elements: exchange, then stash

    const int i = p.i0 + (int)blockIdx.y;
    const int blk = (int)blockIdx.x;
    const int j0 = ORI == ORI_Y ? 0 : blk * NL;  // first row of the tile
    const int k0 = ORI == ORI_Y ? blk * NL : 0;  // first column of the tile
    const size_t plane = (size_t)p.ny * p.nz;
    const int tid = threadIdx.x;

    // line geometry of this thread for phases A and C
    int t, l;                                    // FFT slot within the line, line within the tile
    if (ORI == ORI_Y) { l = tid % NL; t = tid / NL; } else { t = tid % TT; l = tid / TT; }
    const bool line_ok = ORI == ORI_Y ? (k0 + l < p.nz) : (j0 + l < p.ny);
    const size_t lbase = ORI == ORI_Y ? (size_t)i * plane + (size_t)(k0 + l)
                                      : (size_t)i * plane + (size_t)(j0 + l) * p.nz;
    const size_t lstride = ORI == ORI_Y ? (size_t)p.nz : (size_t)1;
    // stash address of cell (row r, column cc) of the tile
    auto cell_slot = [](int r, int cc) -> int {
        if (ORI == ORI_Y) return r * NL + cc;
        return r * N + XchgContigSw<C, N>::phys(cc);
    };

    // ---------------- phase A: derivative of the F pair along the tile's line axis ----------------
    // ORI_Y: pair (F_z, F_x) -> d/dy F_z (slot 0), d/dy F_x (slot 5)
    // ORI_Z: pair (F_y, F_x) -> d/dz F_y (slot 1), d/dz F_x (slot 2)
    const void* PA = ORI == ORI_Y ? p.F[2] : p.F[1];
    const void* PB_ = p.F[0];
#pragma unroll
    for (int f = 0; f < NF; ++f) {
        C v[16];
#pragma unroll
        for (int q = 0; q < 16; ++q) {
            if (line_ok) v[q] = F::ld(PA, PB_, lbase + (size_t)line_index<N>(t, q) * lstride, f);
            else { v[q].x = 0; v[q].y = 0; }
        }
        if constexpr (ORI == ORI_Y) {
            XchgStrided<C, NL> xb{stash + (size_t)f * TILE + l};
            fft_forward<N>(v, t, tw, xb);
#pragma unroll
            for (int q = 0; q < 16; ++q) v[q] = cmul(v[q], ml_in[spec_index<N>(t, q)]);
            fft_inverse<N>(v, t, tw, xb);
            xb.sync();
#pragma unroll
            for (int q = 0; q < 16; ++q) xb.st(line_index<N>(t, q), v[q]);
        } else {
            XchgContigSw<C, N> xb{stash + (size_t)f * TILE + (size_t)l * N};
            fft_forward<N>(v, t, tw, xb);
#pragma unroll
            for (int q = 0; q < 16; ++q) v[q] = cmul(v[q], ml_in[spec_index<N>(t, q)]);
            fft_inverse<N>(v, t, tw, xb);
            xb.sync();
#pragma unroll
            for (int q = 0; q < 16; ++q) xb.st(line_index<N>(t, q), v[q]);
        }
    }
    __syncthreads();

    // ---------------- phase B: vectorised streaming update ----------------
    {
        const unsigned mask = ORI == ORI_Y ? term_mask(p, i, i + 1, 0, p.ny, k0, k0 + NL)
                                           : term_mask(p, i, i + 1, j0, j0 + NL, 0, p.nz);
        const int upd = ORI == ORI_Y ? tile_update_class(p, i, i + 1, 0, p.ny, k0, min(k0 + NL, p.nz))
                                     : tile_update_class(p, i, i + 1, j0, min(j0 + NL, p.ny), 0, p.nz);
        if (mask == 0u && upd >= 0) half_phase_b<T, CPLX, N, ORI, PAL, true>(p, i, j0, k0, mask, upd, stash);
        else half_phase_b<T, CPLX, N, ORI, PAL, false>(p, i, j0, k0, mask, upd, stash);
    }
    __syncthreads();

    // ---------------- phase C: derivative of the NEW G pair -> scratch (next half-step) ----------------
    void* SA = const_cast<void*>(p.dz[0]);
    void* SB = const_cast<void*>(p.dz[1]);
#pragma unroll
    for (int f = 0; f < NF; ++f) {
        C v[16];
        if constexpr (ORI == ORI_Y) {
            XchgStrided<C, NL> xb{stash + (size_t)f * TILE + l};
#pragma unroll
            for (int q = 0; q < 16; ++q) v[q] = xb.ld(line_index<N>(t, q));
            fft_forward<N>(v, t, tw, xb);
#pragma unroll
            for (int q = 0; q < 16; ++q) v[q] = cmul(v[q], ml_out[spec_index<N>(t, q)]);
            fft_inverse<N>(v, t, tw, xb);
        } else {
            XchgContigSw<C, N> xb{stash + (size_t)f * TILE + (size_t)l * N};
#pragma unroll
            for (int q = 0; q < 16; ++q) v[q] = xb.ld(line_index<N>(t, q));
            fft_forward<N>(v, t, tw, xb);
#pragma unroll
            for (int q = 0; q < 16; ++q) v[q] = cmul(v[q], ml_out[spec_index<N>(t, q)]);
            fft_inverse<N>(v, t, tw, xb);
        }
        if (line_ok) {
#pragma unroll
            for (int q = 0; q < 16; ++q) F::st(SA, SB, lbase + (size_t)line_index<N>(t, q) * lstride, v[q], f);
        }
    }
}

// H update: ORI_Z (line axis z, N = nz); E update: ORI_Y (line axis y, N = ny).
template <typename T, bool CPLX>
int launch_shpf_half(Ctx* c, const UpdParams& p, int half) {
    using C = typename Cx<T>::type;
    if (p.i1 <= p.i0) return 0;
    const int ori = half == IES_HALF_H ? ORI_Z : ORI_Y;
    const int axis = ori == ORI_Y ? 1 : 2;
    const int n = ori == ORI_Y ? c->cfg.ny : c->cfg.nz;
    const int other = ori == ORI_Y ? c->cfg.nz : c->cfg.ny;
    const C* tw = (const C*)c->tw[axis];
    const C* ml_in = (const C*)c->mult[half][axis];
    const C* ml_out = (const C*)c->mult[half ^ 1][axis];
    const size_t sm = sizeof(C) * 16 * HALF_NT * Fld<T, CPLX>::NF;
    const bool pal = p.Cidx != nullptr;
#define H_LAUNCH(NN, OO, PP) {                                                              \
        auto kern = k_shpf_half<T, CPLX, NN, OO, PP>;                                       \
        if (set_smem(kern, sm)) return 1;                                                   \
        kern<<<grid, HALF_NT, sm, c->stream>>>(p, tw, ml_in, ml_out);                           \
    }
#define H_CASE(NN) {                                                                        \
        constexpr int NL = HALF_NT / (NN / 16);                                             \
        dim3 grid((unsigned)((other + NL - 1) / NL), (unsigned)(p.i1 - p.i0));              \
        if (ori == ORI_Y) { if (pal) H_LAUNCH(NN, ORI_Y, true) else H_LAUNCH(NN, ORI_Y, false) } \
        else              { if (pal) H_LAUNCH(NN, ORI_Z, true) else H_LAUNCH(NN, ORI_Z, false) } \
    }
    prof_mark(c, PROF_YLINE_UPDATE, 0);
    IES_FOR_N(n, H_CASE)
    prof_mark(c, PROF_YLINE_UPDATE, 1);
#undef H_CASE
#undef H_LAUNCH
    count_launch();
    IES_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace ies
