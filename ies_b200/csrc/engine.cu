// libies_b200.so -- C-ABI host layer + the non-spectral kernels (FDTD update,
// ghost-plane copies, source injection, collectors, field pack/unpack).
// See include/ies_b200.h for the reference lines each entry point replaces.
#include <cmath>
#include <cstdio>
#include <cstring>
#include <atomic>
#include <algorithm>
#include <cstdlib>
#include <type_traits>

#include "engine.h"
#include "fft_dev.cuh"
#include "update_dev.cuh"

namespace ies {

static thread_local std::string g_err;
static std::atomic<long long> g_launches{0};
void set_error(const std::string& s) { g_err = s; }
void count_launch(int n) { g_launches += n; }
void prof_mark(Ctx* c, int slot, int end) {
    if (!c->profiling) return;
    cudaEvent_t e;
    if (cudaEventCreate(&e) != cudaSuccess) return;
    cudaEventRecord(e, c->stream);
    c->prof_ev[slot][end].push_back(e);
}
bool fft_len_supported(int n) { return n >= 16 && n <= 512 && (n & (n - 1)) == 0; }

// ------------------------------------------------------------------ FDTD -----
// Yee curl with two-point differences (space.py:760-779, 975-994) + the fused
// update/CPML of update_dev.cuh.  One thread per cell, z fastest (coalesced).
template <typename T, bool CPLX, bool PAL>
__global__ void __launch_bounds__(256) k_fdtd(const UpdParams p) {
    using A = typename AccT<CPLX>::type;
    using E = Elem<T, CPLX>;
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    const int j = blockIdx.y * blockDim.y + threadIdx.y;
    const int i = p.i0 + blockIdx.z;
    if (k >= p.nz || j >= p.ny) return;
    const size_t plane = (size_t)p.ny * p.nz;
    const size_t idx = (size_t)i * plane + (size_t)j * p.nz + k;
    const int dir = p.dir;
    A d[6];
    const A fx = E::ld(p.F[0], idx), fy = E::ld(p.F[1], idx), fz = E::ld(p.F[2], idx);
    // y neighbour
    const int jn = j + dir, kn = k + dir, in = i + dir;
    const double sy = dir > 0 ? p.rdy : -p.rdy, sz = dir > 0 ? p.rdz : -p.rdz, sx = dir > 0 ? p.rdx : -p.rdx;
    if (jn >= 0 && jn < p.ny) {
        const size_t n = idx + (ptrdiff_t)dir * p.nz;
        d[0] = a_scale(sy, a_sub(E::ld(p.F[2], n), fz));
        d[5] = a_scale(sy, a_sub(E::ld(p.F[0], n), fx));
    } else { d[0] = a_zero(A()); d[5] = a_zero(A()); }
    if (kn >= 0 && kn < p.nz) {
        const size_t n = idx + (ptrdiff_t)dir;
        d[1] = a_scale(sz, a_sub(E::ld(p.F[1], n), fy));
        d[2] = a_scale(sz, a_sub(E::ld(p.F[0], n), fx));
    } else { d[1] = a_zero(A()); d[2] = a_zero(A()); }
    if (in >= 0 && in < p.nx) {
        const size_t n = idx + (ptrdiff_t)dir * plane;
        d[3] = a_scale(sx, a_sub(E::ld(p.F[2], n), fz));
        d[4] = a_scale(sx, a_sub(E::ld(p.F[1], n), fy));
    } else if (p.halo[0] != nullptr) {
        const size_t n = (size_t)j * p.nz + k;
        d[3] = a_scale(sx, a_sub(E::ld(p.halo[1], n), fz));
        d[4] = a_scale(sx, a_sub(E::ld(p.halo[0], n), fy));
    } else { d[3] = a_zero(A()); d[4] = a_zero(A()); }
    const unsigned mask = p.nterms ? ((1u << p.nterms) - 1u) : 0u;
    cell_update<T, CPLX, PAL>(p, mask, i, j, k, d);
}

// Vectorised variant: 16-byte accesses, V cells along z per thread, block = 32 z-vectors x 8
// rows of one x-plane; the z neighbour of a vector is its own shifted lanes plus one scalar.
// Blocks whose tile lies inside all three update boxes and touches no CPML term (FAST) run
// straight-line code.  Needs nz % V == 0 (16-byte aligned rows); do_update falls back to k_fdtd.
template <typename T, bool CPLX, bool PAL, bool FAST>
__device__ __forceinline__ void fdtd_vec_body(const UpdParams& p, const unsigned mask, const int upd) {
    using A = typename AccT<CPLX>::type;
    using E = Elem<T, CPLX>;
    using VV = Vec<T, CPLX>;
    constexpr int V = VV::V;
    const int k = (blockIdx.x * 32 + threadIdx.x) * V;
    const int j = blockIdx.y * 8 + threadIdx.y;
    const int i = p.i0 + blockIdx.z;
    if (k >= p.nz || j >= p.ny) return;
    const size_t plane = (size_t)p.ny * p.nz;
    const size_t idx = (size_t)i * plane + (size_t)j * p.nz + k;
    const int dir = p.dir;
    const double sy = dir > 0 ? p.rdy : -p.rdy, sz = dir > 0 ? p.rdz : -p.rdz, sx = dir > 0 ? p.rdx : -p.rdx;
    A fx[V], fy[V], fz[V], g[3][V], ny_z[V], ny_x[V], nx_z[V], nx_y[V];
    double cf[V];
    VV::ld(p.F[0], idx, fx); VV::ld(p.F[1], idx, fy); VV::ld(p.F[2], idx, fz);
    const int jn = j + dir, in = i + dir;
    const bool has_y = jn >= 0 && jn < p.ny;
    const bool x_in = in >= 0 && in < p.nx;
    const bool has_x = x_in || p.halo[0] != nullptr;
    if (has_y) {
        const size_t n = idx + (ptrdiff_t)dir * p.nz;
        VV::ld(p.F[2], n, ny_z); VV::ld(p.F[0], n, ny_x);
    }
    if (has_x) {
        const size_t n = x_in ? idx + (ptrdiff_t)dir * plane : (size_t)j * p.nz + k;
        VV::ld(x_in ? p.F[2] : p.halo[1], n, nx_z);
        VV::ld(x_in ? p.F[1] : p.halo[0], n, nx_y);
    }
    // the one z neighbour outside the vector: element k+V (dir > 0) or k-1 (dir < 0)
    const int ke = dir > 0 ? k + V : k - 1;
    const bool has_ze = ke >= 0 && ke < p.nz;
    A ez_y = a_zero(A()), ez_x = a_zero(A());
    if (has_ze) {
        const size_t n = (size_t)i * plane + (size_t)j * p.nz + ke;
        ez_y = E::ld(p.F[1], n); ez_x = E::ld(p.F[0], n);
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) VV::ld(p.G[c], idx, g[c]);
    ld_coeff<V, PAL>(p, idx, cf);
#pragma unroll
    for (int v = 0; v < V; ++v) {
        A d[6];
        if (has_y) { d[0] = a_scale(sy, a_sub(ny_z[v], fz[v])); d[5] = a_scale(sy, a_sub(ny_x[v], fx[v])); }
        else { d[0] = a_zero(A()); d[5] = a_zero(A()); }
        // z neighbour: a lane of the same vector, or the extra element at its end
        const A upy = (v + 1 < V) ? fy[(v + 1 < V) ? v + 1 : v] : ez_y, upx = (v + 1 < V) ? fx[(v + 1 < V) ? v + 1 : v] : ez_x;
        const A dny = (v >= 1) ? fy[(v >= 1) ? v - 1 : v] : ez_y, dnx = (v >= 1) ? fx[(v >= 1) ? v - 1 : v] : ez_x;
        const bool zok = dir > 0 ? ((v + 1 < V) || has_ze) : ((v >= 1) || has_ze);
        if (zok) {
            const A zy = dir > 0 ? upy : dny, zx = dir > 0 ? upx : dnx;
            d[1] = a_scale(sz, a_sub(zy, fy[v])); d[2] = a_scale(sz, a_sub(zx, fx[v]));
        } else { d[1] = a_zero(A()); d[2] = a_zero(A()); }
        if (has_x) { d[3] = a_scale(sx, a_sub(nx_z[v], fz[v])); d[4] = a_scale(sx, a_sub(nx_y[v], fy[v])); }
        else { d[3] = a_zero(A()); d[4] = a_zero(A()); }
        A gg[3] = {g[0][v], g[1][v], g[2][v]};
        if constexpr (FAST) cell_update_fast<CPLX>(upd, cf[v], d, gg);
        else cell_update_regs<T, CPLX>(p, mask, i, j, k + v, cf[v], d, gg);
        g[0][v] = gg[0]; g[1][v] = gg[1]; g[2][v] = gg[2];
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) vst_stream<VV>(p.G[c], idx, g[c], 0);     // st.global.cs (not re-read in this half-step)
}

template <typename T, bool CPLX, bool PAL>
__global__ void __launch_bounds__(256) k_fdtd_vec(const UpdParams p) {
    constexpr int V = Vec<T, CPLX>::V;
    const int i = p.i0 + blockIdx.z;
    const int j0 = blockIdx.y * 8, k0 = blockIdx.x * 32 * V;
    const int j1 = min(j0 + 8, p.ny), k1 = min(k0 + 32 * V, p.nz);
    const unsigned mask = term_mask(p, i, i + 1, j0, j1, k0, k1);
    const int upd = tile_update_class(p, i, i + 1, j0, j1, k0, k1);
    if (mask == 0u && upd >= 0) fdtd_vec_body<T, CPLX, PAL, true>(p, mask, upd);
    else fdtd_vec_body<T, CPLX, PAL, false>(p, mask, upd);
}

// Ghost-plane copies F[-1] = F[1]*pp ; F[0] = F[-2]*pm along `axis` for the three
// components, and for the two halo planes along y/z (space.py:1714-1796).
template <typename T, bool CPLX>
__global__ void k_ghost(void* f0, void* f1, void* f2, void* h0, void* h1, int nx, int ny, int nz,
                        int axis, double2 pp, double2 pm) {
    using A = typename AccT<CPLX>::type;
    using E = Elem<T, CPLX>;
    const int na = axis == 0 ? ny : nx, nb = axis == 2 ? ny : nz;   // the two other extents
    const long tid = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const long per = (long)na * nb;
    auto mulp = [](A v, double2 ph) -> A {
        if constexpr (CPLX) return make_double2(v.x * ph.x - v.y * ph.y, v.x * ph.y + v.y * ph.x);
        else return v * ph.x;
    };
    if (tid < per * 3) {
        const int comp = (int)(tid / per);
        const long r = tid % per;
        const int a = (int)(r / nb), b = (int)(r % nb);
        void* f = comp == 0 ? f0 : comp == 1 ? f1 : f2;
        auto at = [&](int s) -> size_t {
            int i = axis == 0 ? s : a;
            int j = axis == 1 ? s : (axis == 0 ? a : b);
            int k = axis == 2 ? s : b;
            return ((size_t)i * ny + j) * nz + k;
        };
        const int N = axis == 0 ? nx : axis == 1 ? ny : nz;
        E::st(f, at(N - 1), mulp(E::ld(f, at(1)), pp));
        E::st(f, at(0), mulp(E::ld(f, at(N - 2)), pm));
    } else if (h0 != nullptr && axis > 0) {
        // halo planes are (ny, nz); patch along y (axis 1) or z (axis 2)
        const long r = tid - per * 3;
        const int other = axis == 1 ? nz : ny;
        if (r < 2L * other) {
            void* h = r < other ? h0 : h1;
            const int o = (int)(r % other);
            auto at = [&](int s) -> size_t { return axis == 1 ? (size_t)s * nz + o : (size_t)o * nz + s; };
            const int N = axis == 1 ? ny : nz;
            E::st(h, at(N - 1), mulp(E::ld(h, at(1)), pp));
            E::st(h, at(0), mulp(E::ld(h, at(N - 2)), pm));
        }
    }
}

// Setter.put_src (source.py:167-253)
template <typename T, bool CPLX>
__global__ void k_put_src(void* f, int ny, int nz, Box bx, double2 pulse, int hard,
                          const double2* px, const double2* py, const double2* pz) {
    using A = typename AccT<CPLX>::type;
    using E = Elem<T, CPLX>;
    const int ex = bx.hi[0] - bx.lo[0], ey = bx.hi[1] - bx.lo[1], ez = bx.hi[2] - bx.lo[2];
    const long tid = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= (long)ex * ey * ez) return;
    const int c = (int)(tid % ez), b = (int)((tid / ez) % ey), a = (int)(tid / ((long)ez * ey));
    double vr = pulse.x, vi = pulse.y;
    if (px) {
        // pulse * ((px*py)*pz), the product order of source.py:231-232
        const double2 X = px[a], Y = py[b], Z = pz[c];
        const double tr = X.x * Y.x - X.y * Y.y, ti = X.x * Y.y + X.y * Y.x;
        const double ur = tr * Z.x - ti * Z.y, ui = tr * Z.y + ti * Z.x;
        const double wr = vr * ur - vi * ui, wi = vr * ui + vi * ur;
        vr = wr; vi = wi;
    }
    const size_t idx = ((size_t)(bx.lo[0] + a) * ny + (bx.lo[1] + b)) * nz + (bx.lo[2] + c);
    A add;
    if constexpr (CPLX) add = make_double2(vr, vi); else add = vr;
    if (hard) E::st(f, idx, add);
    else E::st(f, idx, a_add(E::ld(f, idx), add));
}

// ---- spectral axes of ANY length (the reference takes any N: space.py:145-162) -----------------
// The FFT kernels cover powers of two in 16..512.  For every other length the derivative
// ifft(M fft(x)) is applied as what it is, a circulant matrix: y_j = sum_k c[(j - k) mod N] x_k with
// c = ifft(M) (host, ies_set_multiplier).  O(N) per cell instead of O(log N) -- the shipped 50^3
// SHPF set-up costs 50 multiply-adds per derivative and cell -- but on the GPU, in the field
// precision rules of the FFT path, with the same update / CPML code behind it (k_update_generic).
template <typename T, bool CPLX>
__global__ void __launch_bounds__(256)
k_circ(const void* __restrict__ src, void* __restrict__ dst, const typename Cx<T>::type* __restrict__ c,
       int nx, int ny, int nz, int axis) {
    using A = typename AccT<CPLX>::type;
    using E = Elem<T, CPLX>;
    const size_t ncell = (size_t)nx * ny * nz;
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= ncell) return;
    const int k = (int)(idx % nz), j = (int)((idx / nz) % ny), i = (int)(idx / ((size_t)nz * ny));
    const int n = axis == 0 ? nx : axis == 1 ? ny : nz;
    const int me = axis == 0 ? i : axis == 1 ? j : k;
    const size_t stride = axis == 0 ? (size_t)ny * nz : axis == 1 ? (size_t)nz : 1;
    const size_t base = idx - (size_t)me * stride;
    // accumulate in the transform's precision (T), like the FFT path
    T ar = 0, ai = 0;
    int m = me;                                          // (me - q) mod n, walked downwards
    for (int q = 0; q < n; ++q) {
        const typename Cx<T>::type cc = c[m];
        if constexpr (CPLX) {
            const A x = E::ld(src, base + (size_t)q * stride);
            ar = r_add(ar, r_sub(r_mul(cc.x, (T)x.x), r_mul(cc.y, (T)x.y)));
            ai = r_add(ai, r_add(r_mul(cc.x, (T)x.y), r_mul(cc.y, (T)x.x)));
        } else {
            ar = r_fma(cc.x, (T)E::ld(src, base + (size_t)q * stride), ar);
        }
        m = m == 0 ? n - 1 : m - 1;
    }
    if constexpr (CPLX) E::st(dst, idx, make_double2((double)ar, (double)ai));
    else E::st(dst, idx, (double)ar);
}

// cell update with all spectral derivatives read from scratch (direct-circulant path)
template <typename T, bool CPLX, bool PAL>
__global__ void __launch_bounds__(256) k_update_generic(const UpdParams p) {
    using A = typename AccT<CPLX>::type;
    using E = Elem<T, CPLX>;
    const size_t ncell = (size_t)p.nx * p.ny * p.nz;
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= ncell) return;
    const int k = (int)(idx % p.nz), j = (int)((idx / p.nz) % p.ny), i = (int)(idx / ((size_t)p.nz * p.ny));
    const size_t plane = (size_t)p.ny * p.nz;
    A d[6];
    d[0] = E::ld(p.dys[0], idx); d[5] = E::ld(p.dys[1], idx);
    d[1] = E::ld(p.dz[0], idx);  d[2] = E::ld(p.dz[1], idx);
    if (p.pstd) { d[3] = E::ld(p.dxs[0], idx); d[4] = E::ld(p.dxs[1], idx); }
    else {
        const int in = i + p.dir;
        const double sx = p.dir > 0 ? p.rdx : -p.rdx;
        if (in >= 0 && in < p.nx) {
            const size_t nb = idx + (ptrdiff_t)p.dir * plane;
            d[3] = a_scale(sx, a_sub(E::ld(p.F[2], nb), E::ld(p.F[2], idx)));
            d[4] = a_scale(sx, a_sub(E::ld(p.F[1], nb), E::ld(p.F[1], idx)));
        } else if (p.halo[0] != nullptr) {
            const size_t nb = (size_t)j * p.nz + k;
            d[3] = a_scale(sx, a_sub(E::ld(p.halo[1], nb), E::ld(p.F[2], idx)));
            d[4] = a_scale(sx, a_sub(E::ld(p.halo[0], nb), E::ld(p.F[1], idx)));
        } else { d[3] = a_zero(A()); d[4] = a_zero(A()); }
    }
    const unsigned mask = p.nterms ? ((1u << p.nterms) - 1u) : 0u;
    cell_update<T, CPLX, PAL>(p, mask, i, j, k, d);
}

// CPML corrections as a pass of their own over the absorber cells (space.py:1110-1712), after the
// update kernels ran WITHOUT terms.  Inside the update kernels a term costs every warp that meets an
// absorber row one dependent HBM round trip per term (~2.8 us under load) in front of its stores, and
// terms on the y / z faces sit in every y-line tile: the all-face-CPML step ran 35-40 % slower than
// the x-only one.  Here one thread = one (term, cell): the derivative the term consumes is fetched
// again -- x differences and FDTD differences recomputed from F (unchanged during the half-step), z
// derivatives from the spectral scratch, y derivatives from the side buffer the update phase saved
// for the absorber rows -- then psi = b psi + a d ; G += sign C2 (kf d + psi), the reference's
// statements in the reference's order (main update first, then faces y, z, x: one launch per axis,
// the terms of one launch touch disjoint (cell, component) pairs).
// VW = cells per thread along z: Vec<T,CPLX>::V (16-byte accesses) when the term's box and the
// arrays allow it (x and y faces on even grids), else 1.
template <typename T, bool CPLX, int VW>
__global__ void __launch_bounds__(256)
k_pml_terms(const UpdParams p, const int t0) {
    using A = typename AccT<CPLX>::type;
    using E = Elem<T, CPLX>;
    using VV = Vec<T, CPLX>;
    const PmlTermDev& q = p.terms[t0 + blockIdx.y];
    const int ax = q.axis, df = q.diff, dir = p.dir;
    const unsigned ey = (unsigned)(q.hi[1] - q.lo[1]), ezv = (unsigned)(q.hi[2] - q.lo[2]) / VW;
    const unsigned ngrp = (unsigned)(q.hi[0] - q.lo[0]) * ey * ezv;          // < 2^32 (checked by the launcher)
    const size_t plane = (size_t)p.ny * p.nz;
    // where the term's derivative comes from
    const int fc = (df == 3 || df == 0) ? 2 : (df == 4 || df == 1) ? 1 : 0;  // F component differentiated
    for (unsigned g = blockIdx.x * blockDim.x + threadIdx.x; g < ngrp; g += gridDim.x * blockDim.x) {
        const int k = q.lo[2] + (int)(g % ezv) * VW, j = q.lo[1] + (int)((g / ezv) % ey), i = q.lo[0] + (int)(g / (ezv * ey));
        const size_t idx = (size_t)i * plane + (size_t)j * p.nz + k;
        A d[VW];
#pragma unroll
        for (int v = 0; v < VW; ++v) d[v] = a_zero(A());
        auto ldv = [&](const void* arr, size_t at, A (&o)[VW]) {
            if constexpr (VW == 1) o[0] = E::ld(arr, at); else VV::ld(arr, at, o);
        };
        if (df == 3 || df == 4) {                       // d/dx F_z (3), d/dx F_y (4)
            if (p.pstd) ldv(p.dxs[df == 3 ? 0 : 1], idx, d);
            else {
                const int in = i + dir;
                const double sx = dir > 0 ? p.rdx : -p.rdx;
                const bool inside = in >= 0 && in < p.nx;
                if (inside || p.halo[0] != nullptr) {
                    A a[VW], b[VW];
                    if (inside) ldv(p.F[fc], idx + (ptrdiff_t)dir * plane, a);
                    else ldv(p.halo[df == 3 ? 1 : 0], (size_t)j * p.nz + k, a);
                    ldv(p.F[fc], idx, b);
#pragma unroll
                    for (int v = 0; v < VW; ++v) d[v] = a_scale(sx, a_sub(a[v], b[v]));
                }
            }
        } else if (df == 1 || df == 2) {                // d/dz F_y (1), d/dz F_x (2)
            if (p.fdtd) {
                const double sz = dir > 0 ? p.rdz : -p.rdz;
#pragma unroll
                for (int v = 0; v < VW; ++v) {
                    const int kn = k + v + dir;
                    if (kn >= 0 && kn < p.nz) d[v] = a_scale(sz, a_sub(E::ld(p.F[fc], idx + v + (ptrdiff_t)dir), E::ld(p.F[fc], idx + v)));
                }
            } else ldv(p.dz[df == 1 ? 0 : 1], (size_t)((long long)idx + p.dz_off), d);
        } else {                                        // d/dy F_z (0), d/dy F_x (5)
            if (p.fdtd) {
                const int jn = j + dir;
                const double sy = dir > 0 ? p.rdy : -p.rdy;
                if (jn >= 0 && jn < p.ny) {
                    A a[VW], b[VW];
                    ldv(p.F[fc], idx + (ptrdiff_t)dir * p.nz, a);
                    ldv(p.F[fc], idx, b);
#pragma unroll
                    for (int v = 0; v < VW; ++v) d[v] = a_scale(sy, a_sub(a[v], b[v]));
                }
            } else {
                const int jj = j < p.ys_lo_n ? j : j - p.ys_hi_0 + p.ys_lo_n;
                const size_t sidx = ((size_t)i * p.ys_rows + jj) * p.nz + k;
#pragma unroll
                for (int v = 0; v < VW; ++v) {
                    if constexpr (CPLX) {
                        d[v] = E::ld(p.dy_side, (df == 0 ? 0 : (size_t)p.nx * p.ys_rows * p.nz) + sidx + v);
                    } else {
                        using C2 = typename std::conditional<std::is_same<T, float>::value, float2, double2>::type;
                        const C2 w = ((const C2*)p.dy_side)[sidx + v];
                        d[v] = df == 0 ? (double)w.x : (double)w.y;
                    }
                }
            }
        }
        // psi index of the first cell; the VW cells are adjacent in psi for x / y faces (VW > 1 only there)
        const int n0 = (ax == 0 ? i - q.lo[0] : ax == 1 ? j - q.lo[1] : k - q.lo[2]);
        const int pn = n0 + q.psi_off;
        const int p0 = ax == 0 ? pn : i, p1 = ax == 1 ? pn : j, p2 = ax == 2 ? pn : k;
        const size_t pidx = ((size_t)p0 * q.pdim[1] + p1) * q.pdim[2] + p2;
        double cf[VW];
        if (p.Cidx) ld_coeff<VW, true>(p, idx, cf); else ld_coeff<VW, false>(p, idx, cf);
        A psi[VW], gg[VW];
        ldv(q.psi, pidx, psi);
        ldv(p.G[q.comp], idx, gg);
        const double tb = q.b[n0], ta = q.a[n0], tk = q.kf[n0];           // VW > 1: same n for the VW cells
#pragma unroll
        for (int v = 0; v < VW; ++v) {
            psi[v] = a_add(a_scale(tb, psi[v]), a_scale(ta, d[v]));
            const A pr = E::rnd(psi[v]);
            gg[v] = a_add(gg[v], a_scale(q.sign, a_scale(cf[v], a_add(a_scale(tk, d[v]), pr))));
        }
        if constexpr (VW == 1) { E::st(q.psi, pidx, psi[0]); E::st(p.G[q.comp], idx, gg[0]); }
        else { VV::st(q.psi, pidx, psi); VV::st(p.G[q.comp], idx, gg); }
    }
}

template <typename T, bool CP>
static int launch_pml_terms(ies_ctx* c, const UpdParams& p) {
    // one launch per face axis in the reference's order y, z, x (terms are stored in that order)
    constexpr int V = Vec<T, CP>::V;
    int t = 0;
    while (t < p.nterms) {
        int t1 = t;
        long maxc = 0;
        bool vec = V > 1 && p.terms[t].axis != 2 && p.nz % V == 0;
        while (t1 < p.nterms && p.terms[t1].axis == p.terms[t].axis) {
            const PmlTermDev& q = p.terms[t1];
            const long cells = (long)(q.hi[0] - q.lo[0]) * (q.hi[1] - q.lo[1]) * (q.hi[2] - q.lo[2]);
            if (cells >= (1L << 32)) { set_error("CPML box too large"); return 1; }
            maxc = std::max(maxc, cells);
            // 16-byte accesses need the z range to start and end on a vector boundary (psi of x / y faces has nz columns)
            vec = vec && q.lo[2] % V == 0 && q.hi[2] % V == 0 && q.pdim[2] % V == 0;
            ++t1;
        }
        if (maxc > 0) {
            const long groups = vec ? maxc / V : maxc;
            const unsigned gx = (unsigned)std::min<long>((groups + 255) / 256, 148L * 32);
            const dim3 grid(gx, (unsigned)(t1 - t));
            if (vec) k_pml_terms<T, CP, V><<<grid, 256, 0, c->stream>>>(p, t);
            else k_pml_terms<T, CP, 1><<<grid, 256, 0, c->stream>>>(p, t);
            count_launch();
            IES_CUDA(cudaGetLastError());
        }
        t = t1;
    }
    return 0;
}

// Lossless palette compression of a coefficient array: distinct bit patterns are appended to
// `pal` (<= 256 slots, EMPTY-initialised) with atomicCAS; *overflow is set when they do not fit.
#define IES_PAL_EMPTY 0xFFFFFFFFFFFFFFFFull
__global__ void k_palette_build(const unsigned long long* __restrict__ cbits, size_t n,
                                unsigned long long* pal, int* overflow) {
    unsigned long long last = IES_PAL_EMPTY;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const unsigned long long v = cbits[i];
        if (v == last) continue;
        last = v;
        int k = 0;
        for (; k < 256; ++k) {
            unsigned long long cur = *((volatile unsigned long long*)pal + k);
            if (cur == v) break;
            if (cur == IES_PAL_EMPTY) {
                cur = atomicCAS(pal + k, IES_PAL_EMPTY, v);
                if (cur == IES_PAL_EMPTY || cur == v) break;
            }
        }
        if (k == 256) *overflow = 1;
    }
}
__global__ void k_palette_index(const unsigned long long* __restrict__ cbits, size_t n,
                                const unsigned long long* __restrict__ pal, uint8_t* __restrict__ idx) {
    __shared__ unsigned long long sp[256];
    sp[threadIdx.x] = pal[threadIdx.x];           // blockDim.x == 256
    __syncthreads();
    int last = 0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const unsigned long long v = cbits[i];
        int k = last;
        if (sp[k] != v) { k = 0; while (k < 255 && sp[k] != v) ++k; }
        last = k;
        idx[i] = (uint8_t)k;
    }
}

// One CTA per y-line tile (plane i, column block kb of w columns, all rows): writes the tile's
// coefficient if every cell holds the same bit pattern, else a NaN.
__global__ void k_tile_uniform(const unsigned long long* __restrict__ cbits, int ny, int nz, int w, double* __restrict__ out) {
    const int i = blockIdx.y, kb = blockIdx.x;
    const int k0 = kb * w, cols = min(w, nz - k0);
    const size_t base = (size_t)i * ny * nz + k0;
    const unsigned long long first = cbits[base];
    int diff = 0;
    for (int e = threadIdx.x; e < ny * cols; e += blockDim.x) {
        const int j = e / cols, c = e - j * cols;
        diff |= cbits[base + (size_t)j * nz + c] != first;
    }
    diff = __syncthreads_or(diff);
    if (threadIdx.x == 0)
        out[(size_t)i * gridDim.x + kb] = diff ? __longlong_as_double(0x7ff8000000000000LL) : __longlong_as_double((long long)first);
}

// pack / unpack a box for ies_get_field / ies_set_field
template <typename S>
__global__ void k_pack(const S* f, S* out, int ny, int nz, Box bx, int unpack) {
    const int ex = bx.hi[0] - bx.lo[0], ey = bx.hi[1] - bx.lo[1], ez = bx.hi[2] - bx.lo[2];
    const long n = (long)ex * ey * ez;
    for (long tid = (long)blockIdx.x * blockDim.x + threadIdx.x; tid < n; tid += (long)gridDim.x * blockDim.x) {
        const int c = (int)(tid % ez), b = (int)((tid / ez) % ey), a = (int)(tid / ((long)ez * ey));
        const size_t idx = ((size_t)(bx.lo[0] + a) * ny + (bx.lo[1] + b)) * nz + (bx.lo[2] + c);
        if (unpack) const_cast<S*>(f)[idx] = out[tid]; else out[tid] = f[idx];
    }
}

// ------------------------------------------------------------- collectors -----
struct DftDev {
    double2* acc[4];
    int nf;
    long ncell;
};

// Running DFT of the collectors, time-blocked: do_RFT only SAMPLES the four tangential
// components on the plane into a ring of DFT_TB time slots (2 array passes over the plane);
// every DFT_TB steps (or when a result is read) k_dft_acc_block adds the slots to the
// accumulators in time order -- the (Nf, plane) complex accumulators, the dominant traffic of
// collector.py:323-338 (4 x Nf x plane x 32 B per step), are read and written once per block.
// The per-step arithmetic (operand order of collector.py:334) and the summation order are
// unchanged, so the spectra are bit-identical to a step-by-step accumulation.
constexpr int DFT_TB = 16;
struct DftTimes { double t[DFT_TB]; };

// phase[s][f] = exp(2 pi i f t_s dt)
__global__ void k_dft_phase(const double* freqs, int nf, DftTimes ts, int nslots, double dt, double2* phase) {
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    const int s = blockIdx.y;
    if (f >= nf || s >= nslots) return;
    const double arg = ((2.0 * 3.141592653589793) * freqs[f]) * ts.t[s] * dt;
    double sn, cs;
    sincos(arg, &sn, &cs);
    phase[(size_t)s * nf + f] = make_double2(cs, sn);
}

// ring[(slot*4 + q) * ncell + cell] = component q on the collector plane (SF = TF - IF lazily)
template <typename T, bool CPLX>
__global__ void __launch_bounds__(256)
k_dft_sample(double2* __restrict__ ring, long ncell, int slot, const void* a0, const void* a1, const void* a2,
             const void* a3, const void* b0, const void* b1, const void* b2, const void* b3, int ny, int nz, Box bx) {
    using E = Elem<T, CPLX>;
    const int ey = bx.hi[1] - bx.lo[1], ez = bx.hi[2] - bx.lo[2];
    const long tid = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= ncell) return;
    const int c = (int)(tid % ez), b = (int)((tid / ez) % ey), a = (int)(tid / ((long)ez * ey));
    const size_t idx = ((size_t)(bx.lo[0] + a) * ny + (bx.lo[1] + b)) * nz + (bx.lo[2] + c);
    const void* fa[4] = {a0, a1, a2, a3};
    const void* fb[4] = {b0, b1, b2, b3};
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        auto x = E::ld(fa[q], idx);
        if (fb[q]) {
            // Empty3D.get_SF: SF = TF - IF, evaluated in field precision (space.py:2173-2179)
            x = E::rnd(a_sub(x, E::ld(fb[q], idx)));
        }
        double2 v;
        if constexpr (CPLX) v = x; else v = make_double2(x, 0.0);
        ring[((size_t)slot * 4 + q) * ncell + tid] = v;
    }
}

// One thread = one (component, cell): its nslots samples stay in registers while it walks the
// frequencies; phase table [slot][f] in shared memory.
template <bool CPLX>
__global__ void __launch_bounds__(256)
k_dft_acc_block(DftDev d, const double2* __restrict__ ring, const double2* __restrict__ phase, int nslots, double dt) {
    extern __shared__ double2 sph[];
    for (int e = threadIdx.x; e < nslots * d.nf; e += blockDim.x) sph[e] = phase[e];
    __syncthreads();
    const long tid = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const int q = blockIdx.y;
    if (tid >= d.ncell) return;
    double2 v[DFT_TB];
#pragma unroll
    for (int s = 0; s < DFT_TB; ++s)
        v[s] = s < nslots ? ring[((size_t)s * 4 + q) * d.ncell + tid] : make_double2(0.0, 0.0);
    for (int f = 0; f < d.nf; ++f) {
        double2* acc = d.acc[q] + (size_t)f * d.ncell + tid;
        double2 sum = *acc;
#pragma unroll
        for (int s = 0; s < DFT_TB; ++s) {
            if (s < nslots) {
                const double2 ph = sph[s * d.nf + f];
                // (F * exp(..)) * dt
                double re, im;
                if constexpr (CPLX) { re = v[s].x * ph.x - v[s].y * ph.y; im = v[s].x * ph.y + v[s].y * ph.x; }
                else { re = v[s].x * ph.x; im = v[s].x * ph.y; }
                sum.x += re * dt; sum.y += im * dt;
            }
        }
        *acc = sum;
    }
}

template <typename T, bool CPLX>
__global__ void k_probe(void* out, long tsteps, long tstep, const void* a0, const void* a1, const void* a2,
                        const void* a3, const void* a4, const void* a5, const void* b0, const void* b1,
                        const void* b2, const void* b3, const void* b4, const void* b5, size_t idx) {
    using E = Elem<T, CPLX>;
    const int q = threadIdx.x;
    if (q >= 6) return;
    const void* fa[6] = {a0, a1, a2, a3, a4, a5};
    const void* fb[6] = {b0, b1, b2, b3, b4, b5};
    auto x = E::ld(fa[q], idx);
    if (fb[q]) x = a_sub(x, E::ld(fb[q], idx));
    E::st(out, (size_t)q * tsteps + tstep, x);
}

}  // namespace ies

// =============================================================== C ABI =======
using namespace ies;

struct ies_ctx : public Ctx {};
struct ies_dft {
    ies_ctx* ctx;
    Box box;
    int comps[4];
    int nf;
    long ncell;
    double2* acc[4];
    double* freqs;
    double2* phase;             // [DFT_TB][nf]
    double2* ring;              // [DFT_TB][4][ncell] sampled plane values not yet accumulated
    int npend;                  // filled slots
    double pend_t[DFT_TB];      // their time steps
    double dt;
    bool cplx;
    cudaStream_t last_stream;
};
struct ies_probe {
    ies_ctx* ctx;
    size_t idx;
    long tsteps;
    void* buf;
};

#define DISPATCH(ctx, ...)                                             \
    switch ((ctx)->cfg.dtype) {                                         \
        case IES_F32:  { using T = float;  constexpr bool CP = false; __VA_ARGS__; } break;  \
        case IES_F64:  { using T = double; constexpr bool CP = false; __VA_ARGS__; } break;  \
        case IES_C64:  { using T = float;  constexpr bool CP = true;  __VA_ARGS__; } break;  \
        case IES_C128: { using T = double; constexpr bool CP = true;  __VA_ARGS__; } break;  \
        default: set_error("bad dtype"); return 1;                      \
    }

namespace ies {
int dev_alloc(Ctx* c, void** p, size_t bytes, bool zero) {
    IES_CUDA(cudaMalloc(p, bytes ? bytes : 1));
    c->owned.push_back(*p);
    if (zero) IES_CUDA(cudaMemsetAsync(*p, 0, bytes, c->stream));
    return 0;
}
}  // namespace ies

static int fill_params(ies_ctx* c, int half, UpdParams& p) {
    const int fo = half == IES_HALF_H ? 0 : 3, go = half == IES_HALF_H ? 3 : 0;
    for (int q = 0; q < 3; ++q) { p.F[q] = c->F[fo + q]; p.G[q] = c->F[go + q]; p.box[q] = c->ubox[go + q]; }
    p.C = c->C[half];
    const bool pal = c->use_palette && c->Cnpal[half] > 0;
    p.Cidx = pal ? c->Cidx[half] : nullptr;
    p.Ctile = (c->use_ctile && c->cfg.method != IES_FDTD) ? c->Ctile[half] : nullptr;
    for (int q = 0; q < MAX_PAL; ++q) p.cpal[q] = pal ? c->Cpal_host[half][q] : 0.0;
    if (!p.C) { set_error("init_update_constants() has not been called (no coefficients uploaded)"); return 1; }
    const bool nb = half == IES_HALF_H ? c->has_next : c->has_prev;
    p.halo[0] = nb ? c->halo_recv[half][0] : nullptr;
    p.halo[1] = nb ? c->halo_recv[half][1] : nullptr;
    p.dz[0] = c->scratch[0]; p.dz[1] = c->scratch[1];
    p.dxs[0] = c->scratch[2]; p.dxs[1] = c->scratch[3];
    p.dys[0] = p.dys[1] = nullptr;
    p.nx = c->cfg.nx; p.ny = c->cfg.ny; p.nz = c->cfg.nz;
    p.dir = half == IES_HALF_H ? +1 : -1;
    p.i0 = 0; p.i1 = c->cfg.nx;
    p.pstd = c->cfg.method == IES_PSTD;
    p.dz_off = 0; p.dz_discard = 0; p.dz_keep_lo = 0; p.dz_keep_hi = (short)std::min(p.nz, 32767);
    p.rdx = 1.0 / c->cfg.dx; p.rdy = 1.0 / c->cfg.dy; p.rdz = 1.0 / c->cfg.dz;
    p.nterms = (int)c->terms[half].size();
    for (int t = 0; t < p.nterms; ++t) p.terms[t] = c->terms[half][t];
    p.dy_side = nullptr; p.ys_lo_n = 0; p.ys_hi_0 = p.ny; p.ys_rows = 0;
    p.fdtd = c->cfg.method == IES_FDTD;
    return 0;
}

static int ensure_scratch(ies_ctx* c, int first, int last) {
    const size_t fbytes = (size_t)c->cfg.nx * c->cfg.ny * c->cfg.nz * c->esize;
    for (int q = first; q <= last; ++q)
        if (!c->scratch[q]) if (dev_alloc(c, &c->scratch[q], fbytes)) return 1;
    return 0;
}

// SHPF with a Bloch / periodic x axis: ghost-plane copies of the DIFFERENTIATED field after
// the update (_updateH_BBC_SHPF / _updateE_BBC_SHPF, space.py:1898-1912, 2073-2085).
template <typename T, bool CP>
static int post_ghost_x(ies_ctx* c, const UpdParams& p) {
    if (c->cfg.method != IES_SHPF || !c->ghost_on[0]) return 0;
    if (c->cfg.nx < 4) { set_error("periodic axis needs >= 4 cells"); return 1; }
    const long tot = (long)c->cfg.ny * c->cfg.nz * 3;
    k_ghost<T, CP><<<(unsigned)((tot + 255) / 256), 256, 0, c->stream>>>(
        const_cast<void*>(p.F[0]), const_cast<void*>(p.F[1]), const_cast<void*>(p.F[2]), nullptr, nullptr,
        c->cfg.nx, c->cfg.ny, c->cfg.nz, 0, make_double2(c->ghost_pp[0][0], c->ghost_pp[0][1]),
        make_double2(c->ghost_pm[0][0], c->ghost_pm[0][1]));
    count_launch();
    IES_CUDA(cudaGetLastError());
    return 0;
}

// phase: -1 = whole half-step; 0 = the part that needs no neighbour plane (the z-line and
// x-line derivative passes), so that a slab's halo exchange can overlap it; 1 = the rest.
template <typename T, bool CP>
static int do_update(ies_ctx* c, int half, int phase) {
    UpdParams p;
    if (fill_params(c, half, p)) return 1;
    const int nx = c->cfg.nx, ny = c->cfg.ny, nz = c->cfg.nz;
    // use_fused: 1 = wherever instantiated, 0 = never, -1 (default) = where it was measured faster: lines
    // up to 256 points (at 512 the z role's stage tables and padded exchange take 98 KB per CTA and
    // leave 32 KB of L1 to the streaming phase: 4.58 vs 3.94 ms/step on 256x512x512)
    const bool fused_ok = c->cfg.method == IES_SHPF && !CP && c->cfg.ny == c->cfg.nz &&
                          (c->cfg.ny == 64 || c->cfg.ny == 128 || c->cfg.ny == 256 || c->cfg.ny == 512);
    // ... and no CPML on the y / z faces (their terms sit in every y-line tile; with the psi lines
    // prefetched into L2 the two-kernel path runs them faster: 3.87 vs 4.17 ms/step on 1024x256x256)
    bool yz_pml = false;
    for (int t = 0; t < p.nterms; ++t) yz_pml |= p.terms[t].axis != 0;
    // ... and double precision (fp32: 2.32 vs 2.09 ms/step on 1024x256x256 -- half the bytes per tile, the
    // roles' fixed costs weigh twice as much)
    // CPML corrections in a pass of their own (k_pml_terms): the update kernels then see no term at all
    // (default: whenever a y or z face carries terms, and for x-only absorbers on slabs large enough that the
    //  extra launch is noise -- headline 3.17 -> 3.13 ms/step; a 256x64x64 FDTD step is launch-bound)
    const bool split = p.nterms > 0 && !c->generic && (c->use_pml_split > 0 ||
                                        (c->use_pml_split < 0 && (yz_pml || (size_t)nx * ny * nz >= ((size_t)1 << 23))));
    const bool fused = fused_ok && (c->use_fused > 0 || (c->use_fused < 0 && c->cfg.ny <= 256 && (!yz_pml || split) && c->dbl));
    const bool overlap_ok = c->cfg.method != IES_FDTD && !fused;
    if (!overlap_ok) { if (phase == 0) return 0; phase = -1; }       // everything in phase 1
    if (split && c->cfg.method != IES_FDTD) {
        // absorber rows of the y faces: the update phase saves their y derivatives for the correction pass
        int lo_n = 0, hi_0 = ny;
        for (int t = 0; t < p.nterms; ++t) {
            const PmlTermDev& q = p.terms[t];
            if (q.axis != 1 || q.hi[1] <= q.lo[1]) continue;
            if (q.lo[1] < ny / 2) lo_n = std::max(lo_n, q.hi[1]); else hi_0 = std::min(hi_0, q.lo[1]);
        }
        if (lo_n > hi_0) { lo_n = ny; hi_0 = ny; }          // faces meet: every row is saved
        const int rows = lo_n + (ny - hi_0);
        if (rows > 0) {
            const size_t need = (size_t)(CP ? 2 : 1) * nx * rows * nz * (c->dbl ? 16 : 8);
            if (c->dy_side_bytes < need) {
                if (dev_alloc(c, &c->dy_side, need, false)) return 1;       // (an outgrown buffer stays owned until destroy)
                c->dy_side_bytes = need;
            }
            p.dy_side = c->dy_side; p.ys_lo_n = lo_n; p.ys_hi_0 = hi_0; p.ys_rows = rows;
        }
    }
    UpdParams pm = p;                       // what the update kernels see
    if (split) pm.nterms = 0;
    if (c->cfg.method == IES_FDTD) {
        // ghost copies on the differentiated field, axis order x, y, z (space.py:1798-1858)
        for (int a = 0; a < 3; ++a) {
            if (!c->ghost_on[a]) continue;
            const int dims[3] = {nx, ny, nz};
            if (dims[a] < 4) { set_error("periodic axis needs >= 4 cells"); return 1; }
            const long per = (long)(a == 0 ? ny : nx) * (a == 2 ? ny : nz);
            const long tot = per * 3 + 2L * (ny > nz ? ny : nz);
            void* h0 = (a > 0 && p.halo[0]) ? const_cast<void*>(p.halo[0]) : nullptr;
            void* h1 = (a > 0 && p.halo[1]) ? const_cast<void*>(p.halo[1]) : nullptr;
            k_ghost<T, CP><<<(unsigned)((tot + 255) / 256), 256, 0, c->stream>>>(
                const_cast<void*>(p.F[0]), const_cast<void*>(p.F[1]), const_cast<void*>(p.F[2]), h0, h1,
                nx, ny, nz, a, make_double2(c->ghost_pp[a][0], c->ghost_pp[a][1]),
                make_double2(c->ghost_pm[a][0], c->ghost_pm[a][1]));
            count_launch();
        }
        constexpr int V = Vec<T, CP>::V;
        prof_mark(c, PROF_FDTD, 0);
        if (c->fdtd_vec && nz % V == 0) {
            dim3 blk(32, 8, 1);
            dim3 grid((nz + 32 * V - 1) / (32 * V), (ny + 7) / 8, nx);
            if (p.Cidx) k_fdtd_vec<T, CP, true><<<grid, blk, 0, c->stream>>>(pm);
            else k_fdtd_vec<T, CP, false><<<grid, blk, 0, c->stream>>>(pm);
        } else {
            dim3 blk(nz >= 64 ? 64 : 32, nz >= 64 ? 4 : 8, 1);
            dim3 grid((nz + blk.x - 1) / blk.x, (ny + blk.y - 1) / blk.y, nx);
            if (p.Cidx) k_fdtd<T, CP, true><<<grid, blk, 0, c->stream>>>(pm);
            else k_fdtd<T, CP, false><<<grid, blk, 0, c->stream>>>(pm);
        }
        prof_mark(c, PROF_FDTD, 1);
        count_launch();
        IES_CUDA(cudaGetLastError());
        if (split) return launch_pml_terms<T, CP>(c, p);
        return 0;
    }
    for (int a = 1; a < 3; ++a)
        if (!c->mult[half][a]) { set_error("spectral multiplier not set (malloc()/init_update_constants() missing)"); return 1; }
    if (c->generic) {
        if (phase == 0) return 0;
        using C = typename Cx<T>::type;
        const size_t fbytes = (size_t)nx * ny * nz * c->esize;
        for (int q = 0; q < 6; ++q) {
            if ((q == 2 || q == 3) && c->cfg.method != IES_PSTD) continue;
            if (!c->scratch[q]) if (dev_alloc(c, &c->scratch[q], fbytes)) return 1;
        }
        const size_t ncell = (size_t)nx * ny * nz;
        const unsigned grid = (unsigned)((ncell + 255) / 256);
        auto circ = [&](const void* src, void* dst, int axis) {
            k_circ<T, CP><<<grid, 256, 0, c->stream>>>(src, dst, (const C*)c->circ[half][axis], nx, ny, nz, axis);
            count_launch();
        };
        circ(p.F[1], c->scratch[0], 2); circ(p.F[0], c->scratch[1], 2);         // d/dz F_y, d/dz F_x
        circ(p.F[2], c->scratch[4], 1); circ(p.F[0], c->scratch[5], 1);         // d/dy F_z, d/dy F_x
        if (c->cfg.method == IES_PSTD) {
            if (!c->circ[half][0]) { set_error("x multiplier not set"); return 1; }
            circ(p.F[2], c->scratch[2], 0); circ(p.F[1], c->scratch[3], 0);     // d/dx F_z, d/dx F_y
        }
        UpdParams pg = p;                   // CPML inside the update (terms walk of cell_update)
        pg.dz[0] = c->scratch[0]; pg.dz[1] = c->scratch[1];
        pg.dxs[0] = c->scratch[2]; pg.dxs[1] = c->scratch[3];
        pg.dys[0] = c->scratch[4]; pg.dys[1] = c->scratch[5];
        pg.dy_side = nullptr;
        prof_mark(c, PROF_YLINE_UPDATE, 0);
        if (pg.Cidx) k_update_generic<T, CP, true><<<grid, 256, 0, c->stream>>>(pg);
        else k_update_generic<T, CP, false><<<grid, 256, 0, c->stream>>>(pg);
        prof_mark(c, PROF_YLINE_UPDATE, 1);
        count_launch();
        IES_CUDA(cudaGetLastError());
        return post_ghost_x<T, CP>(c, p);
    }
    {
        const bool ring_only = fused && !split && c->fused_ring_planes > 0 && c->fused_ring_planes < c->cfg.nx;
        if (!ring_only && ensure_scratch(c, 0, c->cfg.method == IES_PSTD ? 3 : 1)) return 1;
    }
    p.dz[0] = pm.dz[0] = c->scratch[0]; p.dz[1] = pm.dz[1] = c->scratch[1];
    if (fused) {
        // one launch: z-line tiles run LEAD planes ahead of the y-line update tiles (shpf_fused.cuh)
        int ring = c->fused_ring_planes;
        if (split) ring = 0;                // the correction pass reads the z derivatives of the whole slab afterwards
        if (ring <= 0 || ring >= c->cfg.nx) ring = c->cfg.nx;
        if (ring < c->cfg.nx) {
            if (ring < c->fused_lead + 2) ring = c->fused_lead + 2;
            if (c->fused_ring_alloc < ring) {
                const size_t rb = (size_t)ring * ny * nz * c->esize;
                for (int q = 0; q < 2; ++q) if (dev_alloc(c, &c->fused_ring[q], rb)) return 1;
                c->fused_ring_alloc = ring;
            }
            pm.dz[0] = c->fused_ring[0]; pm.dz[1] = c->fused_ring[1];
        }
        // the scratch is dead once the y role has read it, except the columns of the z-face absorbers when the
        // correction pass runs afterwards (it differentiates along z there)
        pm.dz_discard = c->fused_discard ? 1 : 0;
        pm.dz_keep_lo = 0; pm.dz_keep_hi = (short)nz;        // (fused: nz <= 512)
        if (split)
            for (int t = 0; t < p.nterms; ++t) {
                const PmlTermDev& q = p.terms[t];
                if (q.axis != 2 || q.hi[2] <= q.lo[2]) continue;
                if (q.lo[2] < nz / 2) pm.dz_keep_lo = (short)std::max((int)pm.dz_keep_lo, q.hi[2]); else pm.dz_keep_hi = (short)std::min((int)pm.dz_keep_hi, q.lo[2]);
            }
        const int keep = c->fused_ring_planes;
        c->fused_ring_planes = ring;
        const int rc = launch_shpf_fused<T, CP>(c, pm, half);
        c->fused_ring_planes = keep;
        if (rc == 1) return 1;
        if (rc == 0) {
            if (split && launch_pml_terms<T, CP>(c, p)) return 1;
            return post_ghost_x<T, CP>(c, p);
        }
        pm.dz[0] = c->scratch[0]; pm.dz[1] = c->scratch[1];       // no instantiation: two-kernel path
        pm.dz_discard = 0;
    }
    p.dxs[0] = pm.dxs[0] = c->scratch[2]; p.dxs[1] = pm.dxs[1] = c->scratch[3];
    if (phase != 1) {
        if (c->cfg.method == IES_PSTD) {
            if (!c->mult[half][0]) { set_error("x multiplier not set"); return 1; }
            if (launch_sline<T, CP>(c, p.F[2], p.F[1], c->scratch[2], c->scratch[3], half, 0, 0, nx)) return 1;
        }
        if (launch_zline<T, CP>(c, p.F[1], p.F[0], c->scratch[0], c->scratch[1], half, 0, nx, 0)) return 1;
    }
    if (phase != 0) {
        if (launch_yline_update<T, CP>(c, pm, half)) return 1;
        if (split && launch_pml_terms<T, CP>(c, p)) return 1;
        if (post_ghost_x<T, CP>(c, p)) return 1;
    }
    return 0;
}

extern "C" {

const char* ies_last_error(void) { return g_err.c_str(); }
int64_t ies_launch_count(void) { return g_launches.load(); }

int ies_device_count(int* n) { IES_CUDA(cudaGetDeviceCount(n)); return 0; }

static int create_impl(const ies_config* cfg, ies_ctx* c);

int ies_create(const ies_config* cfg, ies_ctx** out) {
    if (!cfg || !out) { set_error("null argument"); return 1; }
    if (cfg->nx < 1 || cfg->ny < 2 || cfg->nz < 2) { set_error("bad grid"); return 1; }
    if (cfg->dtype < 0 || cfg->dtype > 3 || cfg->method < 0 || cfg->method > 2) { set_error("bad dtype/method"); return 1; }
    if (cfg->method != IES_FDTD && (cfg->ny < 2 || cfg->nz < 2)) { set_error("bad grid"); return 1; }
    int ndev = 0;
    IES_CUDA(cudaGetDeviceCount(&ndev));
    if (cfg->device < 0 || cfg->device >= ndev) { set_error("no such CUDA device"); return 1; }
    IES_CUDA(cudaSetDevice(cfg->device));
    ies_ctx* c = new ies_ctx();
    c->own_stream = nullptr; c->ev_halo = nullptr; c->ev_t0 = c->ev_t1 = nullptr;
    c->stage = nullptr; c->stage_bytes = 0; c->seq_ring = nullptr; c->src_tab = nullptr; c->src_tab_bytes = 0;
    c->peer_block[0] = c->peer_block[1] = nullptr;
    if (create_impl(cfg, c)) {              // the error text survives the clean-up
        const std::string keep = ies_last_error();
        if (c->own_stream) { c->stream = c->own_stream; ies_destroy(c); }
        else { for (void* p : c->owned) cudaFree(p); delete c; }
        set_error(keep);
        return 1;
    }
    *out = c;
    return 0;
}

static int create_impl(const ies_config* cfg, ies_ctx* c) {
    c->cfg = *cfg;
    // any spectral axis the FFT kernels do not cover switches the context to direct circulant sums
    c->generic = cfg->method != IES_FDTD &&
                 (!fft_len_supported(cfg->ny) || !fft_len_supported(cfg->nz) ||
                  (cfg->method == IES_PSTD && !fft_len_supported(cfg->nx)));
    for (int h = 0; h < 2; ++h) for (int a = 0; a < 3; ++a) c->circ[h][a] = nullptr;
    c->cplx = cfg->dtype >= 2;
    c->dbl = (cfg->dtype & 1) != 0;
    c->esize = (c->dbl ? 8 : 4) * (c->cplx ? 2 : 1);
    IES_CUDA(cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking));
    c->stream = c->own_stream;
    IES_CUDA(cudaEventCreateWithFlags(&c->ev_halo, cudaEventDisableTiming));
    IES_CUDA(cudaEventCreate(&c->ev_t0));
    IES_CUDA(cudaEventCreate(&c->ev_t1));
    c->profiling = 0;
    const size_t ncell = (size_t)cfg->nx * cfg->ny * cfg->nz;
    const size_t fbytes = ncell * c->esize;
    for (int q = 0; q < 6; ++q) if (dev_alloc(c, &c->F[q], fbytes)) return 1;
    for (int h = 0; h < 2; ++h) { c->C[h] = nullptr; c->Cidx[h] = nullptr; c->Cpal[h] = nullptr; c->Cnpal[h] = 0; c->Ctile[h] = nullptr; }
    c->use_ctile = 1;
    c->fdtd_vec = 1;
    if (const char* e = getenv("IES_B200_FDTD_VEC")) c->fdtd_vec = atoi(e);
    if (const char* e = getenv("IES_B200_CTILE")) c->use_ctile = atoi(e);
    c->use_palette = 0;     // measured slower than the f64 array (index -> value dependent loads), kept as an option
    if (const char* e = getenv("IES_B200_PALETTE")) c->use_palette = atoi(e);
    for (int q = 0; q < 6; ++q) c->scratch[q] = nullptr;
    // spectral scratch is allocated on first use
    c->use_pml_split = -1; c->dy_side = nullptr; c->dy_side_bytes = 0;
    if (const char* e = getenv("IES_B200_PML_SPLIT")) c->use_pml_split = atoi(e);
    c->use_fused = -1; c->fused_zb = 2; c->fused_prefetch = 0; c->fused_discard = 1; c->fused_lead = 6; c->fused_ring_planes = 0; c->fused_ring_alloc = 0;
    c->fused_ring[0] = c->fused_ring[1] = nullptr; c->fused_sync = nullptr; c->twz_t = nullptr; c->fused_prof = nullptr; c->fused_prof_mem = nullptr;
    if (const char* e = getenv("IES_B200_FUSED")) c->use_fused = atoi(e);
    if (const char* e = getenv("IES_B200_FUSED_LEAD")) c->fused_lead = std::max(1, atoi(e));
    if (const char* e = getenv("IES_B200_FUSED_RING")) c->fused_ring_planes = atoi(e);
    {
        // halo block: [H y | H z | E y | E z | flags], planes 256-byte aligned
        const size_t pbytes = (((size_t)cfg->ny * cfg->nz * c->esize) + 255) / 256 * 256;
        c->halo_block_bytes = 4 * pbytes + 256;
        if (dev_alloc(c, &c->halo_block, c->halo_block_bytes)) return 1;
        for (int h = 0; h < 2; ++h) {
            for (int w = 0; w < 2; ++w) c->halo_recv[h][w] = (char*)c->halo_block + (size_t)(2 * h + w) * pbytes;
            c->halo_flag[h] = (unsigned*)((char*)c->halo_block + 4 * pbytes) + 16 * h;
            c->push_seq[h] = c->wait_seq[h] = 0;
        }
        for (int n = 0; n < 2; ++n) {
            c->peer_block[n] = nullptr;
            for (int h = 0; h < 2; ++h) { c->peer_flag[n][h] = nullptr; c->peer_recv[n][h][0] = c->peer_recv[n][h][1] = nullptr; }
        }
        c->seq_ring = nullptr; c->flag_write_mode = 0;
    }
    for (int h = 0; h < 2; ++h) for (int a = 0; a < 3; ++a) c->mult[h][a] = nullptr;
    // master twiddles W_N[k] = exp(-2 pi i k / N) per axis, in FFT precision
    const int dims[3] = {cfg->nx, cfg->ny, cfg->nz};
    for (int a = 0; a < 3; ++a) {
        c->tw[a] = nullptr;
        if (cfg->method == IES_FDTD || (a == 0 && cfg->method != IES_PSTD) || c->generic) continue;
        const int n = dims[a];
        std::vector<double> hd(2 * n);
        std::vector<float> hf(2 * n);
        for (int k = 0; k < n; ++k) {
            const long double ang = -2.0L * 3.14159265358979323846264338327950288L * (long double)k / (long double)n;
            hd[2 * k] = (double)cosl(ang); hd[2 * k + 1] = (double)sinl(ang);
            hf[2 * k] = (float)hd[2 * k]; hf[2 * k + 1] = (float)hd[2 * k + 1];
        }
        const size_t b = (size_t)2 * n * (c->dbl ? 8 : 4);
        if (dev_alloc(c, &c->tw[a], b, false)) return 1;
        IES_CUDA(cudaMemcpyAsync(c->tw[a], c->dbl ? (void*)hd.data() : (void*)hf.data(), b, cudaMemcpyHostToDevice, c->stream));
        IES_CUDA(cudaStreamSynchronize(c->stream));
    }
    if (cfg->method == IES_SHPF) {
        // fused half-step: ticket + per-plane counters
        void* q; if (dev_alloc(c, &q, sizeof(unsigned) * (size_t)(1 + 2 * cfg->nx + 32 * cfg->nx))) return 1; c->fused_sync = (unsigned*)q;
    }
    // stage tables of the y and z axes transposed to [m][jj] (fft_dev.cuh TwTables: forward stage NS = 16,
    // inverse stage NS = N/16; one table at N = 256): lines handled by adjacent lanes read them coalesced
    for (int a = 1; a < 3; ++a) {
        c->tw_t[a] = nullptr;
        if (cfg->method == IES_FDTD || c->generic) continue;
        const int n = dims[a];
        if (n <= 16) continue;
        const int r1 = (n / 16 >= 16) ? 16 : n / 16, fwd = r1 * 16, nsi = n / 16, fstep = n / (16 * r1);
        const bool shared = n == 256;
        const int tot = (shared ? 0 : fwd) + n;
        std::vector<double> hd(2 * (size_t)tot);
        std::vector<float> hf(2 * (size_t)tot);
        auto W = [&](int k, int at) {
            const long double ang = -2.0L * 3.14159265358979323846264338327950288L * (long double)(k & (n - 1)) / (long double)n;
            hd[2 * at] = (double)cosl(ang); hd[2 * at + 1] = (double)sinl(ang);
            hf[2 * at] = (float)hd[2 * at]; hf[2 * at + 1] = (float)hd[2 * at + 1];
        };
        int at = 0;
        if (!shared) for (int q = 0; q < fwd; ++q) W((q / 16) * (q % 16) * fstep, at++);
        for (int q = 0; q < n; ++q) W((q / nsi) * (q % nsi), at++);
        const size_t b = (size_t)2 * tot * (c->dbl ? 8 : 4);
        if (dev_alloc(c, &c->tw_t[a], b, false)) return 1;
        IES_CUDA(cudaMemcpyAsync(c->tw_t[a], c->dbl ? (void*)hd.data() : (void*)hf.data(), b, cudaMemcpyHostToDevice, c->stream));
        IES_CUDA(cudaStreamSynchronize(c->stream));
    }
    c->tw_t[0] = nullptr;
    c->twz_t = c->tw_t[2];
    for (int q = 0; q < 6; ++q) {
        c->ubox[q].lo[0] = c->ubox[q].lo[1] = c->ubox[q].lo[2] = 0;
        c->ubox[q].hi[0] = cfg->nx; c->ubox[q].hi[1] = cfg->ny; c->ubox[q].hi[2] = cfg->nz;
    }
    for (int a = 0; a < 3; ++a) c->ghost_on[a] = 0;
    c->has_prev = cfg->rank > 0;
    c->has_next = cfg->rank < cfg->nranks - 1;
    IES_CUDA(cudaStreamSynchronize(c->stream));
    return 0;
}

int ies_destroy(ies_ctx* c) {
    if (!c) return 0;
    cudaSetDevice(c->cfg.device);
    cudaStreamSynchronize(c->stream);
    for (void* p : c->owned) cudaFree(p);
    if (c->stage) cudaFree(c->stage);
    if (c->src_tab) cudaFree(c->src_tab);
    for (int n = 0; n < 2; ++n) if (c->peer_block[n]) cudaIpcCloseMemHandle(c->peer_block[n]);
    if (c->seq_ring) cudaFreeHost(c->seq_ring);
    if (c->ev_halo) cudaEventDestroy(c->ev_halo);
    if (c->ev_t0) cudaEventDestroy(c->ev_t0);
    if (c->ev_t1) cudaEventDestroy(c->ev_t1);
    cudaStreamDestroy(c->own_stream);
    delete c;
    return 0;
}

int ies_set_option(ies_ctx* c, const char* name, int64_t value) {
    if (!c || !name) { set_error("null argument"); return 1; }
    const std::string n(name);
    const int v = (int)value;
    IES_CUDA(cudaSetDevice(c->cfg.device));
    IES_CUDA(cudaStreamSynchronize(c->stream));
    if (n == "palette") c->use_palette = v;
    else if (n == "ctile") c->use_ctile = v;
    else if (n == "fdtd_vec") c->fdtd_vec = v;
    else if (n == "pml_split") c->use_pml_split = v;
    else if (n == "fused") c->use_fused = v;
    else if (n == "fused_lead") c->fused_lead = v < 1 ? 1 : v;
    else if (n == "fused_zb") c->fused_zb = v >= 4 ? 4 : v >= 2 ? 2 : 1;
    else if (n == "fused_prefetch") c->fused_prefetch = v;
    else if (n == "fused_discard") c->fused_discard = v;
    else if (n == "fused_ring") c->fused_ring_planes = v;
    else if (n == "fused_prof") {               // 1: start (zeroed) per-phase cycle counters, 0: stop
        if (v && !c->fused_prof_mem) { void* q; if (dev_alloc(c, &q, 16 * 8)) return 1; c->fused_prof_mem = (unsigned long long*)q; }
        if (v) IES_CUDA(cudaMemsetAsync(c->fused_prof_mem, 0, 16 * 8, c->stream));
        c->fused_prof = v ? c->fused_prof_mem : nullptr;
    }
    else if (n == "reset_psi") {               // zero the CPML auxiliary state (restart a run on new fields)
        for (int h = 0; h < 2; ++h)
            for (const PmlTermDev& t : c->terms[h])
                IES_CUDA(cudaMemsetAsync(t.psi, 0, (size_t)t.pdim[0] * t.pdim[1] * t.pdim[2] * c->esize, c->stream));
        IES_CUDA(cudaStreamSynchronize(c->stream));
    }
    else { set_error("unknown option " + n); return 1; }
    return 0;
}

int ies_set_stream(ies_ctx* c, void* s, int use_own_stream) {
    // work already queued on the old stream (set-up, a source injection) must be visible to
    // whatever is enqueued on the new one
    if (!c) { set_error("null context"); return 1; }
    IES_CUDA(cudaSetDevice(c->cfg.device));
    IES_CUDA(cudaStreamSynchronize(c->stream));
    // an external handle of 0 is the legacy default stream of the caller -- a legitimate choice
    // (e.g. torch's default stream), not a request for the context's own stream
    c->stream = use_own_stream ? c->own_stream : (cudaStream_t)s;
    return 0;
}
int ies_timer_start(ies_ctx* c) {
    IES_CUDA(cudaSetDevice(c->cfg.device));
    IES_CUDA(cudaEventRecord(c->ev_t0, c->stream));
    return 0;
}
int ies_timer_stop(ies_ctx* c, double* ms) {
    IES_CUDA(cudaSetDevice(c->cfg.device));
    IES_CUDA(cudaEventRecord(c->ev_t1, c->stream));
    IES_CUDA(cudaEventSynchronize(c->ev_t1));
    float f = 0.f;
    IES_CUDA(cudaEventElapsedTime(&f, c->ev_t0, c->ev_t1));
    *ms = (double)f;
    return 0;
}
int ies_profile(ies_ctx* c, int on) {
    for (int s = 0; s < 4; ++s) for (int e = 0; e < 2; ++e) {
        for (cudaEvent_t ev : c->prof_ev[s][e]) cudaEventDestroy(ev);
        c->prof_ev[s][e].clear();
    }
    c->profiling = on;
    return 0;
}
int ies_profile_read(ies_ctx* c, int slot, double* ms_total, int64_t* launches) {
    if (slot < 0 || slot > 3) { set_error("bad profile slot"); return 1; }
    IES_CUDA(cudaSetDevice(c->cfg.device));
    IES_CUDA(cudaStreamSynchronize(c->stream));
    double tot = 0.0;
    const size_t n = std::min(c->prof_ev[slot][0].size(), c->prof_ev[slot][1].size());
    for (size_t i = 0; i < n; ++i) {
        float f = 0.f;
        IES_CUDA(cudaEventElapsedTime(&f, c->prof_ev[slot][0][i], c->prof_ev[slot][1][i]));
        tot += f;
    }
    *ms_total = tot; *launches = (int64_t)n;
    return 0;
}
int ies_fused_prof_read(ies_ctx* c, uint64_t* out16) {
    IES_CUDA(cudaSetDevice(c->cfg.device));
    IES_CUDA(cudaStreamSynchronize(c->stream));
    if (!c->fused_prof_mem) { for (int q = 0; q < 16; ++q) out16[q] = 0; return 0; }
    IES_CUDA(cudaMemcpy(out16, c->fused_prof_mem, 16 * 8, cudaMemcpyDeviceToHost));
    return 0;
}
int ies_sync(ies_ctx* c) { IES_CUDA(cudaSetDevice(c->cfg.device)); IES_CUDA(cudaStreamSynchronize(c->stream)); return 0; }

int ies_set_coeff(ies_ctx* c, int half, const double* host, int64_t n) {
    const size_t ncell = (size_t)c->cfg.nx * c->cfg.ny * c->cfg.nz;
    if (half < 0 || half > 1 || (size_t)n != ncell) { set_error("ies_set_coeff: bad size"); return 1; }
    IES_CUDA(cudaSetDevice(c->cfg.device));
    if (!c->C[half]) { void* p; if (dev_alloc(c, &p, ncell * 8, false)) return 1; c->C[half] = (double*)p; }
    IES_CUDA(cudaMemcpyAsync(c->C[half], host, ncell * 8, cudaMemcpyHostToDevice, c->stream));
    // palette form: materials are piecewise constant, so the array usually holds a handful of
    // distinct values; the update kernels then read one index byte per cell instead of 8 bytes
    if (c->cfg.method != IES_FDTD && !c->generic) {
        // per-tile uniform coefficients for the y-line kernel (tile = 4096/ny columns of one plane)
        // = YCfg::W of spectral.cuh: 256 threads (128 for complex dtypes, lines up to 1024) / (ny/16)
#ifndef IES_Y512_THREADS
#define IES_Y512_THREADS 256
#endif
        const int thr = (c->cplx && c->cfg.ny / 16 <= 64) ? 128 : ((!c->cplx && c->cfg.ny == 512) ? IES_Y512_THREADS : 256);
        const int w = thr * 16 / c->cfg.ny > 0 ? thr * 16 / c->cfg.ny : 1;
        const int kt = (c->cfg.nz + w - 1) / w;
        if (!c->Ctile[half]) { void* p; if (dev_alloc(c, &p, (size_t)c->cfg.nx * kt * 8, false)) return 1; c->Ctile[half] = (double*)p; }
        k_tile_uniform<<<dim3((unsigned)kt, (unsigned)c->cfg.nx), 256, 0, c->stream>>>(
            (const unsigned long long*)c->C[half], c->cfg.ny, c->cfg.nz, w, c->Ctile[half]);
        count_launch();
        IES_CUDA(cudaGetLastError());
    }
    c->Cnpal[half] = 0;
    if (!c->use_palette) return 0;
    if (!c->Cidx[half]) {
        void* p; if (dev_alloc(c, &p, ncell, false)) return 1; c->Cidx[half] = (uint8_t*)p;
        if (dev_alloc(c, &p, 257 * 8, false)) return 1; c->Cpal[half] = (double*)p;
    }
    int* d_over = (int*)(c->Cpal[half] + 256);
    IES_CUDA(cudaMemsetAsync(c->Cpal[half], 0xFF, 256 * 8, c->stream));
    IES_CUDA(cudaMemsetAsync(d_over, 0, 8, c->stream));
    k_palette_build<<<1184, 256, 0, c->stream>>>((const unsigned long long*)c->C[half], ncell,
                                                 (unsigned long long*)c->Cpal[half], d_over);
    count_launch();
    unsigned long long hp[257];
    IES_CUDA(cudaMemcpyAsync(hp, c->Cpal[half], 257 * 8, cudaMemcpyDeviceToHost, c->stream));
    IES_CUDA(cudaStreamSynchronize(c->stream));
    if ((int)(hp[256] & 0xffffffffu) == 0) {
        int np = 0;
        while (np < 256 && hp[np] != IES_PAL_EMPTY) ++np;
        if (np > MAX_PAL) return 0;            // too many distinct values: keep the f64 array
        memcpy(c->Cpal_host[half], hp, sizeof(double) * np);
        k_palette_index<<<1184, 256, 0, c->stream>>>((const unsigned long long*)c->C[half], ncell,
                                                     (const unsigned long long*)c->Cpal[half], c->Cidx[half]);
        count_launch();
        IES_CUDA(cudaGetLastError());
        IES_CUDA(cudaStreamSynchronize(c->stream));
        c->Cnpal[half] = np;
    }
    return 0;
}

int ies_set_update_box(ies_ctx* c, int comp, const int32_t lo[3], const int32_t hi[3]) {
    if (comp < 0 || comp > 5) { set_error("bad component"); return 1; }
    for (int a = 0; a < 3; ++a) { c->ubox[comp].lo[a] = lo[a]; c->ubox[comp].hi[a] = hi[a]; }
    return 0;
}

int ies_set_multiplier(ies_ctx* c, int half, int axis, const double* re_im, int32_t n) {
    const int dims[3] = {c->cfg.nx, c->cfg.ny, c->cfg.nz};
    if (half < 0 || half > 1 || axis < 0 || axis > 2 || n != dims[axis]) { set_error("ies_set_multiplier: bad args"); return 1; }
    IES_CUDA(cudaSetDevice(c->cfg.device));
    const size_t b = (size_t)2 * n * (c->dbl ? 8 : 4);
    if (c->generic) {
        // direct-circulant path: first column c = ifft(M), c[m] = (1/n) sum_k M[k] exp(+2 pi i k m / n)
        std::vector<double> hd(2 * (size_t)n);
        std::vector<float> hf(2 * (size_t)n);
        const long double w = 2.0L * 3.14159265358979323846264338327950288L / (long double)n;
        for (int m = 0; m < n; ++m) {
            long double sr = 0, si = 0;
            for (int k = 0; k < n; ++k) {
                const long double ang = w * (long double)(((long long)k * m) % n);
                const long double cr = cosl(ang), ci = sinl(ang);
                sr += (long double)re_im[2 * k] * cr - (long double)re_im[2 * k + 1] * ci;
                si += (long double)re_im[2 * k] * ci + (long double)re_im[2 * k + 1] * cr;
            }
            hd[2 * m] = (double)(sr / n); hd[2 * m + 1] = (double)(si / n);
            hf[2 * m] = (float)hd[2 * m]; hf[2 * m + 1] = (float)hd[2 * m + 1];
        }
        if (!c->circ[half][axis]) if (dev_alloc(c, &c->circ[half][axis], b, false)) return 1;
        IES_CUDA(cudaMemcpyAsync(c->circ[half][axis], c->dbl ? (void*)hd.data() : (void*)hf.data(), b, cudaMemcpyHostToDevice, c->stream));
        IES_CUDA(cudaStreamSynchronize(c->stream));
        c->mult[half][axis] = c->circ[half][axis];       // "multiplier is set" for the checks of do_update
        return 0;
    }
    if (!c->mult[half][axis]) if (dev_alloc(c, &c->mult[half][axis], b, false)) return 1;
    // fold the 1/N of the inverse transform into the table (exact: N is a power of two)
    std::vector<double> hd(2 * n);
    std::vector<float> hf(2 * n);
    for (int k = 0; k < 2 * n; ++k) { hd[k] = re_im[k] / (double)n; hf[k] = (float)hd[k]; }
    IES_CUDA(cudaMemcpyAsync(c->mult[half][axis], c->dbl ? (void*)hd.data() : (void*)hf.data(), b, cudaMemcpyHostToDevice, c->stream));
    IES_CUDA(cudaStreamSynchronize(c->stream));
    return 0;
}

int ies_clear_pml(ies_ctx* c) {
    // the psi arrays and profile tables of the dropped terms go back to the allocator
    IES_CUDA(cudaSetDevice(c->cfg.device));
    IES_CUDA(cudaStreamSynchronize(c->stream));
    for (int h = 0; h < 2; ++h) {
        for (const PmlTermDev& t : c->terms[h]) {
            void* ptrs[2] = {t.psi, const_cast<double*>(t.b)};       // b, a, kf share one allocation
            for (void* q : ptrs) {
                auto it = std::find(c->owned.begin(), c->owned.end(), q);
                if (it != c->owned.end()) { cudaFree(q); c->owned.erase(it); }
            }
        }
        c->terms[h].clear();
    }
    return 0;
}

int ies_add_pml_term(ies_ctx* c, const ies_pml_term* t) {
    if (!t || t->half < 0 || t->half > 1 || t->comp < 0 || t->comp > 2 || t->diff < 0 || t->diff > 5 || t->axis < 0 || t->axis > 2) {
        set_error("ies_add_pml_term: bad term"); return 1;
    }
    if (t->diff / 2 != t->comp) { set_error("ies_add_pml_term: derivative slot does not belong to the component's curl"); return 1; }
    if ((int)c->terms[t->half].size() >= MAX_TERMS) { set_error("too many CPML terms"); return 1; }
    IES_CUDA(cudaSetDevice(c->cfg.device));
    PmlTermDev d;
    d.comp = t->comp; d.diff = t->diff; d.axis = t->axis; d.psi_off = t->psi_off; d.sign = t->sign;
    const int dims[3] = {c->cfg.nx, c->cfg.ny, c->cfg.nz};
    for (int a = 0; a < 3; ++a) {
        d.lo[a] = t->lo[a]; d.hi[a] = t->hi[a];
        if (t->lo[a] < 0 || t->hi[a] > dims[a]) { set_error("CPML box out of range"); return 1; }
        d.pdim[a] = (a == t->axis) ? t->psi_thick : dims[a];
    }
    const int len = t->hi[t->axis] - t->lo[t->axis];
    if (len < 0 || len + t->psi_off > t->psi_thick) { set_error("CPML psi range"); return 1; }
    const size_t pb = (size_t)d.pdim[0] * d.pdim[1] * d.pdim[2] * c->esize;
    if (dev_alloc(c, &d.psi, pb)) return 1;
    void* tb;
    const int L = len > 0 ? len : 1;
    if (dev_alloc(c, &tb, (size_t)3 * L * 8, false)) return 1;
    double* tbd = (double*)tb;
    if (len > 0) {
        IES_CUDA(cudaMemcpyAsync(tbd, t->b, (size_t)len * 8, cudaMemcpyHostToDevice, c->stream));
        IES_CUDA(cudaMemcpyAsync(tbd + L, t->a, (size_t)len * 8, cudaMemcpyHostToDevice, c->stream));
        IES_CUDA(cudaMemcpyAsync(tbd + 2 * L, t->kf, (size_t)len * 8, cudaMemcpyHostToDevice, c->stream));
        IES_CUDA(cudaStreamSynchronize(c->stream));
    }
    d.b = tbd; d.a = tbd + L; d.kf = tbd + 2 * L;
    c->terms[t->half].push_back(d);
    return 0;
}

int ies_set_ghost(ies_ctx* c, int axis, int enabled, double ppr, double ppi, double pmr, double pmi) {
    if (axis < 0 || axis > 2) { set_error("bad axis"); return 1; }
    c->ghost_on[axis] = enabled;
    c->ghost_pp[axis][0] = ppr; c->ghost_pp[axis][1] = ppi;
    c->ghost_pm[axis][0] = pmr; c->ghost_pm[axis][1] = pmi;
    return 0;
}

int ies_set_neighbours(ies_ctx* c, int has_prev, int has_next) { c->has_prev = has_prev; c->has_next = has_next; return 0; }

int ies_update_h(ies_ctx* c, int64_t) {
    IES_CUDA(cudaSetDevice(c->cfg.device));
    DISPATCH(c, return (do_update<T, CP>(c, IES_HALF_H, -1)));
    return 0;
}
int ies_update_e(ies_ctx* c, int64_t) {
    IES_CUDA(cudaSetDevice(c->cfg.device));
    DISPATCH(c, return (do_update<T, CP>(c, IES_HALF_E, -1)));
    return 0;
}
int ies_update_phase(ies_ctx* c, int half, int phase) {
    if (half < 0 || half > 1 || phase < 0 || phase > 1) { set_error("ies_update_phase: bad half/phase"); return 1; }
    IES_CUDA(cudaSetDevice(c->cfg.device));
    DISPATCH(c, return (do_update<T, CP>(c, half, phase)));
    return 0;
}

int ies_halo_send_ptr(ies_ctx* c, int half, int which, void** dev, int64_t* bytes) {
    const size_t pb = (size_t)c->cfg.ny * c->cfg.nz * c->esize;
    if (half == IES_HALF_H) *dev = c->F[which ? IES_EZ : IES_EY];                                   // plane 0
    else *dev = (char*)c->F[which ? IES_HZ : IES_HY] + (size_t)(c->cfg.nx - 1) * pb;                // plane -1
    *bytes = (int64_t)pb;
    return 0;
}
int ies_halo_recv_ptr(ies_ctx* c, int half, int which, void** dev, int64_t* bytes) {
    *dev = c->halo_recv[half][which];
    *bytes = (int64_t)c->cfg.ny * c->cfg.nz * c->esize;
    return 0;
}
// ---- inter-process halo over CUDA IPC (one process per GPU) -----------------------------------
// The receive planes and two arrival flags of a context live in ONE allocation whose IPC handle
// the neighbours map.  A push = copy-engine transfers of my two send planes into the neighbour's
// receive planes followed, in stream order, by a 4-byte write of the push count to its flag; the
// receiver puts a stream memory wait (cuStreamWaitValue32, >=) in front of the update that reads
// the planes.  No kernel, no host synchronisation, no event hand-shake between the processes.
typedef int (*wait32_fn)(cudaStream_t, unsigned long long, unsigned, unsigned);    // CUresult (CUstream, CUdeviceptr, cuuint32_t, flags)
static wait32_fn drv_entry(const char* name) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint(name, &fn, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) return nullptr;
    return (wait32_fn)fn;
}

int ies_halo_ipc_export(ies_ctx* c, void* handle64) {
    if (!c || !handle64) { set_error("null argument"); return 1; }
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    IES_CUDA(cudaSetDevice(c->cfg.device));
    IES_CUDA(cudaStreamSynchronize(c->stream));
    cudaIpcMemHandle_t h;
    IES_CUDA(cudaIpcGetMemHandle(&h, c->halo_block));
    memcpy(handle64, &h, 64);
    return 0;
}

int ies_halo_ipc_connect(ies_ctx* c, int nbr, const void* handle64) {
    if (!c || !handle64 || nbr < 0 || nbr > 1) { set_error("ies_halo_ipc_connect: bad argument"); return 1; }
    IES_CUDA(cudaSetDevice(c->cfg.device));
    if (c->peer_block[nbr]) { cudaIpcCloseMemHandle(c->peer_block[nbr]); c->peer_block[nbr] = nullptr; }
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    void* base = nullptr;
    IES_CUDA(cudaIpcOpenMemHandle(&base, h, cudaIpcMemLazyEnablePeerAccess));
    c->peer_block[nbr] = base;
    const size_t pbytes = (c->halo_block_bytes - 256) / 4;          // same grid on every rank
    for (int hf = 0; hf < 2; ++hf) {
        for (int w = 0; w < 2; ++w) c->peer_recv[nbr][hf][w] = (char*)base + (size_t)(2 * hf + w) * pbytes;
        c->peer_flag[nbr][hf] = (unsigned*)((char*)base + 4 * pbytes) + 16 * hf;
    }
    if (!c->seq_ring) {
        IES_CUDA(cudaHostAlloc((void**)&c->seq_ring, 4096 * sizeof(unsigned), cudaHostAllocDefault));
        if (const char* e = getenv("IES_B200_FLAG_MEMCPY")) c->flag_write_mode = atoi(e);
        if (!drv_entry("cuStreamWriteValue32") || !drv_entry("cuStreamWaitValue32")) c->flag_write_mode = 1;
    }
    return 0;
}

int ies_halo_push(ies_ctx* c, int half) {
    if (half < 0 || half > 1) { set_error("bad half"); return 1; }
    // updateH: my first planes go to rank-1 (neighbour 0); updateE: my last planes to rank+1 (neighbour 1)
    const int nbr = half == IES_HALF_H ? 0 : 1;
    if (!(nbr == 0 ? c->has_prev : c->has_next)) return 0;
    if (!c->peer_block[nbr]) { set_error("ies_halo_push: neighbour not connected (ies_halo_ipc_connect)"); return 1; }
    IES_CUDA(cudaSetDevice(c->cfg.device));
    for (int w = 0; w < 2; ++w) {
        void* sp; int64_t b;
        ies_halo_send_ptr(c, half, w, &sp, &b);
        IES_CUDA(cudaMemcpyAsync(c->peer_recv[nbr][half][w], sp, (size_t)b, cudaMemcpyDeviceToDevice, c->stream));
    }
    const unsigned seq = ++c->push_seq[half];
    if (c->flag_write_mode == 0) {
        static wait32_fn wr = drv_entry("cuStreamWriteValue32");
        const int rc = wr(c->stream, (unsigned long long)(uintptr_t)c->peer_flag[nbr][half], seq, 0u);
        if (rc != 0) { set_error("cuStreamWriteValue32 failed (" + std::to_string(rc) + "); set IES_B200_FLAG_MEMCPY=1"); return 1; }
    } else {
        // the ring slot is rewritten 4096 pushes later; the launch queue is far shorter than that
        unsigned* slot = c->seq_ring + (seq & 4095u);
        *slot = seq;
        IES_CUDA(cudaMemcpyAsync(c->peer_flag[nbr][half], slot, sizeof(unsigned), cudaMemcpyHostToDevice, c->stream));
    }
    return 0;
}

int ies_halo_wait(ies_ctx* c, int half) {
    if (half < 0 || half > 1) { set_error("bad half"); return 1; }
    // updateH reads the planes of rank+1, updateE those of rank-1
    if (!(half == IES_HALF_H ? c->has_next : c->has_prev)) return 0;
    IES_CUDA(cudaSetDevice(c->cfg.device));
    const unsigned seq = ++c->wait_seq[half];
    static wait32_fn wt = drv_entry("cuStreamWaitValue32");
    if (!wt) { set_error("cuStreamWaitValue32 is not available in this driver"); return 1; }
    const int rc = wt(c->stream, (unsigned long long)(uintptr_t)c->halo_flag[half], seq, 1u /* CU_STREAM_WAIT_VALUE_GEQ */);
    if (rc != 0) { set_error("cuStreamWaitValue32 failed (" + std::to_string(rc) + ")"); return 1; }
    return 0;
}

int ies_halo_copy(ies_ctx* dst, ies_ctx* src, int half) {
    // order after everything queued on src's stream, run the copies on dst's stream
    IES_CUDA(cudaSetDevice(src->cfg.device));
    IES_CUDA(cudaEventRecord(src->ev_halo, src->stream));
    IES_CUDA(cudaSetDevice(dst->cfg.device));
    IES_CUDA(cudaStreamWaitEvent(dst->stream, src->ev_halo, 0));
    for (int w = 0; w < 2; ++w) {
        void* s; int64_t b;
        ies_halo_send_ptr(src, half, w, &s, &b);
        IES_CUDA(cudaMemcpyPeerAsync(dst->halo_recv[half][w], dst->cfg.device, s, src->cfg.device, (size_t)b, dst->stream));
    }
    // src must not overwrite the sent planes before the copy ran
    IES_CUDA(cudaEventRecord(dst->ev_halo, dst->stream));
    IES_CUDA(cudaSetDevice(src->cfg.device));
    IES_CUDA(cudaStreamWaitEvent(src->stream, dst->ev_halo, 0));
    return 0;
}

static int check_box(ies_ctx* c, const int32_t lo[3], const int32_t hi[3], Box& bx) {
    const int dims[3] = {c->cfg.nx, c->cfg.ny, c->cfg.nz};
    for (int a = 0; a < 3; ++a) {
        if (lo[a] < 0 || hi[a] > dims[a] || hi[a] < lo[a]) { set_error("box out of range"); return 1; }
        bx.lo[a] = lo[a]; bx.hi[a] = hi[a];
    }
    return 0;
}

int ies_put_src(ies_ctx* c, int comp, const int32_t lo[3], const int32_t hi[3], double re, double im, int hard,
                const double* px, const double* py, const double* pz) {
    if (comp < 0 || comp > 5) { set_error("bad component"); return 1; }
    Box bx; if (check_box(c, lo, hi, bx)) return 1;
    const long n = (long)(hi[0] - lo[0]) * (hi[1] - lo[1]) * (hi[2] - lo[2]);
    if (n <= 0) return 0;
    IES_CUDA(cudaSetDevice(c->cfg.device));
    double2 *dpx = nullptr, *dpy = nullptr, *dpz = nullptr;
    if (px && py && pz) {
        if (!c->cplx) { set_error("Bloch phase tables need a complex field dtype"); return 1; }
        // The phase tables of a Setter never change between calls: they live in a per-context device
        // buffer with a host shadow, and are uploaded only when their content differs (a per-call
        // cudaMallocAsync + three pageable H2D copies put a stream synchronisation and milliseconds of
        // host time into every step of the Bloch-boundary runs).
        const int ex = hi[0] - lo[0], ey = hi[1] - lo[1], ez = hi[2] - lo[2];
        const size_t cnt = (size_t)(ex + ey + ez) * 2;
        std::vector<double> want(cnt);
        memcpy(want.data(), px, (size_t)ex * 16);
        memcpy(want.data() + 2 * ex, py, (size_t)ey * 16);
        memcpy(want.data() + 2 * (ex + ey), pz, (size_t)ez * 16);
        if (c->src_tab_bytes < cnt * 8) {
            if (c->src_tab) { IES_CUDA(cudaStreamSynchronize(c->stream)); cudaFree(c->src_tab); c->src_tab = nullptr; }
            IES_CUDA(cudaMalloc(&c->src_tab, cnt * 8));
            c->src_tab_bytes = cnt * 8;
            c->src_tab_host.clear();
        }
        if (c->src_tab_host != want) {
            IES_CUDA(cudaStreamSynchronize(c->stream));      // a queued injection may still read the old tables
            IES_CUDA(cudaMemcpy(c->src_tab, want.data(), cnt * 8, cudaMemcpyHostToDevice));
            c->src_tab_host.swap(want);
        }
        dpx = (double2*)c->src_tab; dpy = dpx + ex; dpz = dpy + ey;
    }
    DISPATCH(c, k_put_src<T, CP><<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(
        c->F[comp], c->cfg.ny, c->cfg.nz, bx, make_double2(re, im), hard, dpx, dpy, dpz));
    count_launch();
    IES_CUDA(cudaGetLastError());
    return 0;
}

static int ensure_stage(ies_ctx* c, size_t bytes) {
    if (c->stage_bytes >= bytes) return 0;
    if (c->stage) cudaFree(c->stage);
    c->stage = nullptr; c->stage_bytes = 0;
    IES_CUDA(cudaMalloc(&c->stage, bytes));
    c->stage_bytes = bytes;
    return 0;
}

static int pack_launch(ies_ctx* c, void* f, Box bx, long n, int unpack) {
    const unsigned grid = (unsigned)std::min<long>((n + 255) / 256, 148L * 16);
    switch (c->esize) {
        case 4:  k_pack<float><<<grid, 256, 0, c->stream>>>((const float*)f, (float*)c->stage, c->cfg.ny, c->cfg.nz, bx, unpack); break;
        case 8:  k_pack<double><<<grid, 256, 0, c->stream>>>((const double*)f, (double*)c->stage, c->cfg.ny, c->cfg.nz, bx, unpack); break;
        default: k_pack<double2><<<grid, 256, 0, c->stream>>>((const double2*)f, (double2*)c->stage, c->cfg.ny, c->cfg.nz, bx, unpack); break;
    }
    count_launch();
    IES_CUDA(cudaGetLastError());
    return 0;
}

int ies_get_field(ies_ctx* c, int comp, const int32_t lo[3], const int32_t hi[3], void* host) {
    if (comp < 0 || comp > 5) { set_error("bad component"); return 1; }
    Box bx; if (check_box(c, lo, hi, bx)) return 1;
    const long n = (long)(hi[0] - lo[0]) * (hi[1] - lo[1]) * (hi[2] - lo[2]);
    if (n <= 0) return 0;
    IES_CUDA(cudaSetDevice(c->cfg.device));
    const bool whole_planes = lo[1] == 0 && lo[2] == 0 && hi[1] == c->cfg.ny && hi[2] == c->cfg.nz;
    if (whole_planes) {
        const size_t pb = (size_t)c->cfg.ny * c->cfg.nz * c->esize;
        IES_CUDA(cudaMemcpyAsync(host, (char*)c->F[comp] + lo[0] * pb, (size_t)(hi[0] - lo[0]) * pb, cudaMemcpyDeviceToHost, c->stream));
    } else {
        if (ensure_stage(c, (size_t)n * c->esize)) return 1;
        if (pack_launch(c, c->F[comp], bx, n, 0)) return 1;
        IES_CUDA(cudaMemcpyAsync(host, c->stage, (size_t)n * c->esize, cudaMemcpyDeviceToHost, c->stream));
    }
    IES_CUDA(cudaStreamSynchronize(c->stream));
    return 0;
}

int ies_set_field(ies_ctx* c, int comp, const int32_t lo[3], const int32_t hi[3], const void* host) {
    if (comp < 0 || comp > 5) { set_error("bad component"); return 1; }
    Box bx; if (check_box(c, lo, hi, bx)) return 1;
    const long n = (long)(hi[0] - lo[0]) * (hi[1] - lo[1]) * (hi[2] - lo[2]);
    if (n <= 0) return 0;
    IES_CUDA(cudaSetDevice(c->cfg.device));
    const bool whole_planes = lo[1] == 0 && lo[2] == 0 && hi[1] == c->cfg.ny && hi[2] == c->cfg.nz;
    if (whole_planes) {
        const size_t pb = (size_t)c->cfg.ny * c->cfg.nz * c->esize;
        IES_CUDA(cudaMemcpyAsync((char*)c->F[comp] + lo[0] * pb, host, (size_t)(hi[0] - lo[0]) * pb, cudaMemcpyHostToDevice, c->stream));
    } else {
        if (ensure_stage(c, (size_t)n * c->esize)) return 1;
        IES_CUDA(cudaMemcpyAsync(c->stage, host, (size_t)n * c->esize, cudaMemcpyHostToDevice, c->stream));
        if (pack_launch(c, c->F[comp], bx, n, 1)) return 1;
    }
    IES_CUDA(cudaStreamSynchronize(c->stream));
    return 0;
}

int ies_field_ptr(ies_ctx* c, int comp, void** dev) {
    if (comp < 0 || comp > 5) { set_error("bad component"); return 1; }
    *dev = c->F[comp];
    return 0;
}

// ---------------------------------------------------------------- collectors
int ies_dft_create(ies_ctx* c, const int32_t lo[3], const int32_t hi[3], const int32_t comps[4],
                   const double* freqs, int32_t nf, ies_dft** out) {
    Box bx; if (check_box(c, lo, hi, bx)) return 1;
    if (nf < 1) { set_error("nf < 1"); return 1; }
    IES_CUDA(cudaSetDevice(c->cfg.device));
    ies_dft* d = new ies_dft();
    d->ctx = c; d->box = bx; d->nf = nf;
    d->ncell = (long)(hi[0] - lo[0]) * (hi[1] - lo[1]) * (hi[2] - lo[2]);
    for (int q = 0; q < 4; ++q) {
        if (comps[q] < 0 || comps[q] > 5) { set_error("bad component"); delete d; return 1; }
        d->comps[q] = comps[q];
        const size_t b = (size_t)nf * (d->ncell > 0 ? d->ncell : 1) * 16;
        IES_CUDA(cudaMalloc((void**)&d->acc[q], b));
        IES_CUDA(cudaMemsetAsync(d->acc[q], 0, b, c->stream));
    }
    IES_CUDA(cudaMalloc((void**)&d->freqs, (size_t)nf * 8));
    IES_CUDA(cudaMalloc((void**)&d->phase, (size_t)DFT_TB * nf * 16));
    IES_CUDA(cudaMalloc((void**)&d->ring, (size_t)DFT_TB * 4 * (d->ncell > 0 ? d->ncell : 1) * 16));
    d->npend = 0; d->dt = c->cfg.dt; d->cplx = c->cplx; d->last_stream = c->stream;
    IES_CUDA(cudaMemcpyAsync(d->freqs, freqs, (size_t)nf * 8, cudaMemcpyHostToDevice, c->stream));
    IES_CUDA(cudaStreamSynchronize(c->stream));
    *out = d;
    return 0;
}

// adds the pending time slots to the accumulators (stream order: after their sampling kernels)
static int dft_flush(ies_dft* d) {
    if (d->npend == 0 || d->ncell <= 0) { d->npend = 0; return 0; }
    cudaStream_t st = d->last_stream;
    DftTimes ts;
    for (int s = 0; s < DFT_TB; ++s) ts.t[s] = s < d->npend ? d->pend_t[s] : 0.0;
    k_dft_phase<<<dim3((d->nf + 127) / 128, d->npend), 128, 0, st>>>(d->freqs, d->nf, ts, d->npend, d->dt, d->phase);
    DftDev dd; dd.nf = d->nf; dd.ncell = d->ncell;
    for (int q = 0; q < 4; ++q) dd.acc[q] = d->acc[q];
    const size_t sm = (size_t)d->npend * d->nf * 16;
    const dim3 grid((unsigned)((d->ncell + 255) / 256), 4);
    if (d->cplx) {
        if (sm > 48 * 1024) IES_CUDA(cudaFuncSetAttribute(k_dft_acc_block<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
        k_dft_acc_block<true><<<grid, 256, sm, st>>>(dd, d->ring, d->phase, d->npend, d->dt);
    } else {
        if (sm > 48 * 1024) IES_CUDA(cudaFuncSetAttribute(k_dft_acc_block<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
        k_dft_acc_block<false><<<grid, 256, sm, st>>>(dd, d->ring, d->phase, d->npend, d->dt);
    }
    count_launch(2);
    IES_CUDA(cudaGetLastError());
    d->npend = 0;
    return 0;
}

int ies_dft_accumulate(ies_dft* d, ies_ctx* a, ies_ctx* b, int64_t tstep) {
    if (d->ncell <= 0) return 0;
    if (a->cfg.device != d->ctx->cfg.device || (b && b->cfg.device != a->cfg.device)) { set_error("collector spaces must share a device"); return 1; }
    if (b && (b->cfg.dtype != a->cfg.dtype || b->cfg.nx != a->cfg.nx || b->cfg.ny != a->cfg.ny || b->cfg.nz != a->cfg.nz)) {
        set_error("get_SF: TF and IF differ in shape/dtype"); return 1;
    }
    if ((size_t)DFT_TB * d->nf * 16 > 200 * 1024) { set_error("collector: too many frequencies for the blocked DFT (nf <= 800)"); return 1; }
    IES_CUDA(cudaSetDevice(a->cfg.device));
    cudaStream_t st = a->stream;
    if (st != d->last_stream) {
        // pending slots were sampled on another stream: accumulate them there first
        if (dft_flush(d)) return 1;
        IES_CUDA(cudaStreamSynchronize(d->last_stream));
        d->last_stream = st;
    }
    if (b && b->stream != a->stream) {
        IES_CUDA(cudaEventRecord(b->ev_halo, b->stream));
        IES_CUDA(cudaStreamWaitEvent(st, b->ev_halo, 0));
    }
    const void* fa[4]; const void* fb[4];
    for (int q = 0; q < 4; ++q) { fa[q] = a->F[d->comps[q]]; fb[q] = b ? b->F[d->comps[q]] : nullptr; }
    d->cplx = a->cplx; d->dt = a->cfg.dt;
    DISPATCH(a, k_dft_sample<T, CP><<<(unsigned)((d->ncell + 255) / 256), 256, 0, st>>>(
        d->ring, d->ncell, d->npend, fa[0], fa[1], fa[2], fa[3], fb[0], fb[1], fb[2], fb[3], a->cfg.ny, a->cfg.nz, d->box));
    count_launch();
    IES_CUDA(cudaGetLastError());
    d->pend_t[d->npend++] = (double)tstep;
    if (b && b->stream != a->stream) {
        IES_CUDA(cudaEventRecord(a->ev_halo, st));
        IES_CUDA(cudaStreamWaitEvent(b->stream, a->ev_halo, 0));
    }
    if (d->npend == DFT_TB) return dft_flush(d);
    return 0;
}

int ies_dft_read(ies_dft* d, int which, void* host) {
    if (which < 0 || which > 3) { set_error("bad index"); return 1; }
    IES_CUDA(cudaSetDevice(d->ctx->cfg.device));
    if (dft_flush(d)) return 1;
    IES_CUDA(cudaStreamSynchronize(d->last_stream));
    IES_CUDA(cudaStreamSynchronize(d->ctx->stream));
    IES_CUDA(cudaDeviceSynchronize());
    IES_CUDA(cudaMemcpy(host, d->acc[which], (size_t)d->nf * d->ncell * 16, cudaMemcpyDeviceToHost));
    return 0;
}

int ies_dft_destroy(ies_dft* d) {
    if (!d) return 0;
    cudaSetDevice(d->ctx->cfg.device);
    cudaDeviceSynchronize();
    for (int q = 0; q < 4; ++q) cudaFree(d->acc[q]);
    cudaFree(d->freqs); cudaFree(d->phase); cudaFree(d->ring);
    delete d;
    return 0;
}

int ies_probe_create(ies_ctx* c, int32_t i, int32_t j, int32_t k, int64_t tsteps, ies_probe** out) {
    if (i < 0 || i >= c->cfg.nx || j < 0 || j >= c->cfg.ny || k < 0 || k >= c->cfg.nz || tsteps < 1) { set_error("probe out of range"); return 1; }
    IES_CUDA(cudaSetDevice(c->cfg.device));
    ies_probe* p = new ies_probe();
    p->ctx = c; p->tsteps = (long)tsteps;
    p->idx = ((size_t)i * c->cfg.ny + j) * c->cfg.nz + k;
    IES_CUDA(cudaMalloc(&p->buf, (size_t)6 * tsteps * c->esize));
    IES_CUDA(cudaMemset(p->buf, 0, (size_t)6 * tsteps * c->esize));
    *out = p;
    return 0;
}

int ies_probe_record(ies_probe* p, ies_ctx* a, ies_ctx* b, int64_t tstep) {
    if (tstep < 0 || tstep >= p->tsteps) { set_error("probe: tstep out of range"); return 1; }
    IES_CUDA(cudaSetDevice(a->cfg.device));
    if (b && b->stream != a->stream) {
        IES_CUDA(cudaEventRecord(b->ev_halo, b->stream));
        IES_CUDA(cudaStreamWaitEvent(a->stream, b->ev_halo, 0));
    }
    const void* fb[6];
    for (int q = 0; q < 6; ++q) fb[q] = b ? b->F[q] : nullptr;
    DISPATCH(a, k_probe<T, CP><<<1, 32, 0, a->stream>>>(p->buf, p->tsteps, (long)tstep, a->F[0], a->F[1], a->F[2],
        a->F[3], a->F[4], a->F[5], fb[0], fb[1], fb[2], fb[3], fb[4], fb[5], p->idx));
    count_launch();
    IES_CUDA(cudaGetLastError());
    if (b && b->stream != a->stream) {
        IES_CUDA(cudaEventRecord(a->ev_halo, a->stream));
        IES_CUDA(cudaStreamWaitEvent(b->stream, a->ev_halo, 0));
    }
    return 0;
}

int ies_probe_read(ies_probe* p, int comp, void* host) {
    if (comp < 0 || comp > 5) { set_error("bad component"); return 1; }
    IES_CUDA(cudaSetDevice(p->ctx->cfg.device));
    IES_CUDA(cudaDeviceSynchronize());
    const size_t b = (size_t)p->tsteps * p->ctx->esize;
    IES_CUDA(cudaMemcpy(host, (char*)p->buf + comp * b, b, cudaMemcpyDeviceToHost));
    return 0;
}

int ies_probe_destroy(ies_probe* p) {
    if (!p) return 0;
    cudaSetDevice(p->ctx->cfg.device);
    cudaDeviceSynchronize();
    cudaFree(p->buf);
    delete p;
    return 0;
}

}  // extern "C"
