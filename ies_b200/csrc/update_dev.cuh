// Per-cell leap-frog update with the CPML corrections fused in
// (space.py:801-825, 1017-1037 main update; 1110-1712 CPML faces).
#pragma once
#include "engine.h"

namespace ies {

// compute type of the update arithmetic: always double precision, like the
// reference (f64 coefficient arrays promote the expression)
template <bool CPLX> struct AccT;
template <> struct AccT<false> { using type = double; };
template <> struct AccT<true>  { using type = double2; };

__device__ __forceinline__ double  a_add(double a, double b) { return a + b; }
__device__ __forceinline__ double  a_sub(double a, double b) { return a - b; }
__device__ __forceinline__ double  a_scale(double s, double a) { return s * a; }
__device__ __forceinline__ double  a_zero(double) { return 0.0; }
__device__ __forceinline__ double2 a_add(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ double2 a_sub(double2 a, double2 b) { return make_double2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ double2 a_scale(double s, double2 a) { return make_double2(s * a.x, s * a.y); }
__device__ __forceinline__ double2 a_zero(double2) { return make_double2(0.0, 0.0); }

template <typename T, bool CPLX> struct Elem;
template <typename T> struct Elem<T, false> {
    using S = T;
    static __device__ __forceinline__ double ld(const void* p, size_t i) { return (double)((const T*)p)[i]; }
    static __device__ __forceinline__ void st(void* p, size_t i, double v) { ((T*)p)[i] = (T)v; }
    static __device__ __forceinline__ double rnd(double v) { return (double)(T)v; }
};
template <> struct Elem<float, true> {
    using S = float2;
    static __device__ __forceinline__ double2 ld(const void* p, size_t i) {
        float2 v = ((const float2*)p)[i]; return make_double2(v.x, v.y);
    }
    static __device__ __forceinline__ void st(void* p, size_t i, double2 v) {
        ((float2*)p)[i] = make_float2((float)v.x, (float)v.y);
    }
    static __device__ __forceinline__ double2 rnd(double2 v) { return make_double2((double)(float)v.x, (double)(float)v.y); }
};
template <> struct Elem<double, true> {
    using S = double2;
    static __device__ __forceinline__ double2 ld(const void* p, size_t i) { return ((const double2*)p)[i]; }
    static __device__ __forceinline__ void st(void* p, size_t i, double2 v) { ((double2*)p)[i] = v; }
    static __device__ __forceinline__ double2 rnd(double2 v) { return v; }
};

__device__ __forceinline__ bool in_box(const int (&lo)[3], const int (&hi)[3], int i, int j, int k) {
    return i >= lo[0] && i < hi[0] && j >= lo[1] && j < hi[1] && k >= lo[2] && k < hi[2];
}

// Bit mask of the CPML terms whose box intersects [lo,hi) (CTA-level culling).
__device__ __forceinline__ unsigned term_mask(const UpdParams& p, int i0, int i1, int j0, int j1, int k0, int k1) {
    unsigned m = 0;
    for (int t = 0; t < p.nterms; ++t) {
        const PmlTermDev& q = p.terms[t];
        if (q.lo[0] < i1 && q.hi[0] > i0 && q.lo[1] < j1 && q.hi[1] > j0 && q.lo[2] < k1 && q.hi[2] > k0)
            m |= 1u << t;
    }
    return m;
}

// d[] = the six derivatives of this half-step at the cell, slot order IES_D_*:
//   comp x: d[0]-d[1]   comp y: d[2]-d[3]   comp z: d[4]-d[5]
template <typename T, bool CPLX>
__device__ __forceinline__ void cell_update(const UpdParams& p, unsigned mask, int i, int j, int k,
                                            const typename AccT<CPLX>::type (&d)[6]) {
    using A = typename AccT<CPLX>::type;
    using E = Elem<T, CPLX>;
    const size_t idx = ((size_t)i * p.ny + j) * p.nz + k;
    const double C = p.C[idx];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        bool touched = false;
        A g = a_zero(A());
        if (in_box(p.box[c].lo, p.box[c].hi, i, j, k)) {
            g = E::ld(p.G[c], idx);
            g = a_add(g, a_scale(C, a_sub(d[2 * c], d[2 * c + 1])));
            touched = true;
        }
        unsigned m = mask;
        while (m) {
            const int t = __ffs(m) - 1;
            m &= m - 1;
            const PmlTermDev& q = p.terms[t];
            if (q.comp != c || !in_box(q.lo, q.hi, i, j, k)) continue;
            if (!touched) { g = E::ld(p.G[c], idx); touched = true; }
            const int ax = q.axis;
            const int n = (ax == 0 ? i - q.lo[0] : ax == 1 ? j - q.lo[1] : k - q.lo[2]);
            const int pn = n + q.psi_off;
            const int p0 = ax == 0 ? pn : i, p1 = ax == 1 ? pn : j, p2 = ax == 2 ? pn : k;
            const size_t pidx = ((size_t)p0 * q.pdim[1] + p1) * q.pdim[2] + p2;
            A dd = d[0];
#pragma unroll
            for (int s = 1; s < 6; ++s) if (q.diff == s) dd = d[s];
            A psi = E::ld(q.psi, pidx);
            psi = a_add(a_scale(q.b[n], psi), a_scale(q.a[n], dd));
            // psi is stored in field precision and the rounded value is what the
            // field correction uses (space.py:1154-1156)
            E::st(q.psi, pidx, psi);
            psi = E::rnd(psi);
            const A corr = a_scale(C, a_add(a_scale(q.kf[n], dd), psi));
            g = a_add(g, a_scale(q.sign, corr));
        }
        if (touched) E::st(p.G[c], idx, g);
    }
}

}  // namespace ies
