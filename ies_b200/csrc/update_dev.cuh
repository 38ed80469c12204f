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

// Explicitly rounded (never contracted into FMAs): the update is then the reference's own
// NumPy expression -- one rounding per multiply / add / subtract, space.py:801-803 -- and gives
// the same bits in every kernel that inlines it.
__device__ __forceinline__ double  a_add(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double  a_sub(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ double  a_scale(double s, double a) { return __dmul_rn(s, a); }
__device__ __forceinline__ double  a_zero(double) { return 0.0; }
__device__ __forceinline__ double2 a_add(double2 a, double2 b) { return make_double2(__dadd_rn(a.x, b.x), __dadd_rn(a.y, b.y)); }
__device__ __forceinline__ double2 a_sub(double2 a, double2 b) { return make_double2(__dsub_rn(a.x, b.x), __dsub_rn(a.y, b.y)); }
__device__ __forceinline__ double2 a_scale(double s, double2 a) { return make_double2(__dmul_rn(s, a.x), __dmul_rn(s, a.y)); }
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

// Vector access: V = 16 bytes / element consecutive cells along z per thread.
template <typename T, bool CPLX> struct Vec;
template <> struct Vec<double, false> {
    static constexpr int V = 2;
    static __device__ __forceinline__ void ld(const void* p, size_t i, double (&o)[2]) {
        const double2 v = *reinterpret_cast<const double2*>((const double*)p + i); o[0] = v.x; o[1] = v.y;
    }
    // streaming variants (ld/st.global.cs: evict-first): for data touched once per half-step
    static __device__ __forceinline__ void ld_stream(const void* p, size_t i, double (&o)[2]) {
        const double2 v = __ldcs(reinterpret_cast<const double2*>((const double*)p + i)); o[0] = v.x; o[1] = v.y;
    }
    static __device__ __forceinline__ void st_stream(void* p, size_t i, const double (&o)[2]) {
        __stcs(reinterpret_cast<double2*>((double*)p + i), make_double2(o[0], o[1]));
    }
    static __device__ __forceinline__ void st(void* p, size_t i, const double (&o)[2]) {
        *reinterpret_cast<double2*>((double*)p + i) = make_double2(o[0], o[1]);
    }
};
template <> struct Vec<float, false> {
    static constexpr int V = 4;
    static __device__ __forceinline__ void ld(const void* p, size_t i, double (&o)[4]) {
        const float4 v = *reinterpret_cast<const float4*>((const float*)p + i);
        o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w;
    }
    static __device__ __forceinline__ void st(void* p, size_t i, const double (&o)[4]) {
        *reinterpret_cast<float4*>((float*)p + i) = make_float4((float)o[0], (float)o[1], (float)o[2], (float)o[3]);
    }
    static __device__ __forceinline__ void st_stream(void* p, size_t i, const double (&o)[4]) {
        __stcs(reinterpret_cast<float4*>((float*)p + i), make_float4((float)o[0], (float)o[1], (float)o[2], (float)o[3]));
    }
};
template <> struct Vec<float, true> {
    static constexpr int V = 2;
    static __device__ __forceinline__ void ld(const void* p, size_t i, double2 (&o)[2]) {
        const float4 v = *reinterpret_cast<const float4*>((const float2*)p + i);
        o[0] = make_double2(v.x, v.y); o[1] = make_double2(v.z, v.w);
    }
    static __device__ __forceinline__ void st(void* p, size_t i, const double2 (&o)[2]) {
        *reinterpret_cast<float4*>((float2*)p + i) =
            make_float4((float)o[0].x, (float)o[0].y, (float)o[1].x, (float)o[1].y);
    }
    static __device__ __forceinline__ void st_stream(void* p, size_t i, const double2 (&o)[2]) {
        __stcs(reinterpret_cast<float4*>((float2*)p + i),
               make_float4((float)o[0].x, (float)o[0].y, (float)o[1].x, (float)o[1].y));
    }
};
template <> struct Vec<double, true> {
    static constexpr int V = 1;
    static __device__ __forceinline__ void ld(const void* p, size_t i, double2 (&o)[1]) { o[0] = ((const double2*)p)[i]; }
    static __device__ __forceinline__ void st(void* p, size_t i, const double2 (&o)[1]) { ((double2*)p)[i] = o[0]; }
    static __device__ __forceinline__ void st_stream(void* p, size_t i, const double2 (&o)[1]) { __stcs((double2*)p + i, o[0]); }
};
// Streaming accessors: Vec<T,CPLX>::ld_stream / st_stream when the specialisation has them
// (fp64 real), else the plain ones.
template <typename VV, typename X, int V> __device__ __forceinline__ auto vld_stream(const void* p, size_t i, X (&o)[V], int)
    -> decltype(VV::ld_stream(p, i, o)) { VV::ld_stream(p, i, o); }
template <typename VV, typename X, int V> __device__ __forceinline__ void vld_stream(const void* p, size_t i, X (&o)[V], long) { VV::ld(p, i, o); }
template <typename VV, typename X, int V> __device__ __forceinline__ auto vst_stream(void* p, size_t i, const X (&o)[V], int)
    -> decltype(VV::st_stream(p, i, o)) { VV::st_stream(p, i, o); }
template <typename VV, typename X, int V> __device__ __forceinline__ void vst_stream(void* p, size_t i, const X (&o)[V], long) { VV::st(p, i, o); }

// Coefficients of V consecutive cells: from the palette form (one index byte per cell, the
// palette in the kernel-parameter constant bank) when PAL, else from the f64 array.
template <int V, bool PAL> __device__ __forceinline__ void ld_coeff(const UpdParams& q, size_t i, double (&o)[V]) {
    if constexpr (PAL) {
        if constexpr (V == 1) { o[0] = q.cpal[q.Cidx[i]]; }
        else if constexpr (V == 2) {
            const unsigned short w = *reinterpret_cast<const unsigned short*>(q.Cidx + i);
            o[0] = q.cpal[w & 0xff]; o[1] = q.cpal[w >> 8];
        } else {
            static_assert(V == 1 || V == 2 || V == 4, "vector width");
            const unsigned w = *reinterpret_cast<const unsigned*>(q.Cidx + i);
#pragma unroll
            for (int v = 0; v < V; ++v) o[v] = q.cpal[(w >> (8 * v)) & 0xff];
        }
    } else {
        const double* p = q.C;
        if constexpr (V == 1) { o[0] = p[i]; }
        else {
#pragma unroll
            for (int v = 0; v < V; v += 2) {
                const double2 t = *reinterpret_cast<const double2*>(p + i + v); o[v] = t.x; o[v + 1] = t.y;
            }
        }
    }
}

// Update of the three components of one cell held in registers.
//   d[] = the six derivatives of this half-step, slot order IES_D_*:
//   comp x: d[0]-d[1]   comp y: d[2]-d[3]   comp z: d[4]-d[5]
// A CPML term on component c always consumes one of that component's own two curl
// derivatives (space.py:1153-1162 etc.), selected by the parity of its slot.
// COMPS: compile-time mask of the components this kernel updates (the split SHPF half-step
// updates G_y in the z-line kernel and G_x, G_z in the y-line kernel); d[] / g[] entries of
// the other components are not touched.
template <typename T, bool CPLX, int COMPS = 7>
__device__ __forceinline__ void cell_update_regs(const UpdParams& p, unsigned mask, int i, int j, int k,
                                                 const double C, const typename AccT<CPLX>::type (&d)[6],
                                                 typename AccT<CPLX>::type (&g)[3]) {
    using A = typename AccT<CPLX>::type;
    using E = Elem<T, CPLX>;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        if (!(COMPS & (1 << c))) continue;
        if (in_box(p.box[c].lo, p.box[c].hi, i, j, k))
            g[c] = a_add(g[c], a_scale(C, a_sub(d[2 * c], d[2 * c + 1])));
    }
    // CPML terms, one walk over the set bits (terms are stored in the order the reference
    // applies them, so each component still receives its corrections in that order)
    unsigned m = mask;
    while (m) {
        const int t = __ffs(m) - 1;
        m &= m - 1;
        const PmlTermDev& q = p.terms[t];
        const int c = q.comp;
        if (!((COMPS >> c) & 1) || !in_box(q.lo, q.hi, i, j, k)) continue;
        const int ax = q.axis;
        const int n = (ax == 0 ? i - q.lo[0] : ax == 1 ? j - q.lo[1] : k - q.lo[2]);
        const int pn = n + q.psi_off;
        const int p0 = ax == 0 ? pn : i, p1 = ax == 1 ? pn : j, p2 = ax == 2 ? pn : k;
        const size_t pidx = ((size_t)p0 * q.pdim[1] + p1) * q.pdim[2] + p2;
        const int df = q.diff;                          // derivative slot 0..5 (register select, no local array)
        const A dd = df == 0 ? d[0] : df == 1 ? d[1] : df == 2 ? d[2] : df == 3 ? d[3] : df == 4 ? d[4] : d[5];
        A psi = E::ld(q.psi, pidx);
        psi = a_add(a_scale(q.b[n], psi), a_scale(q.a[n], dd));
        // psi is stored in field precision and the rounded value is what the
        // field correction uses (space.py:1154-1156)
        E::st(q.psi, pidx, psi);
        psi = E::rnd(psi);
        const A corr = a_scale(q.sign, a_scale(C, a_add(a_scale(q.kf[n], dd), psi)));
        if (c == 0) g[0] = a_add(g[0], corr);
        else if (c == 1) g[1] = a_add(g[1], corr);
        else g[2] = a_add(g[2], corr);
    }
}

// L2 prefetch of the CPML auxiliary fields a y-line tile (plane i, columns [k0, k1), all rows) is
// going to touch.  The term walk of cell_update_regs loads psi inside a data-dependent loop, one
// dependent round trip per term; under load a round trip to HBM costs ~2.8 us (15 MB in flight at
// 5.6 TB/s), an L2 hit a fraction of it.  Issued at the start of the tile's work (before its FFT
// phase, ~10 us ahead of the use), costs no registers and nothing to wait for; the psi lines of a
// tile are a few KB (unlike the tile's field operands, whose prefetch thrashes L2).
template <typename T, bool CPLX>
__device__ __forceinline__ void prefetch_tile_psi(const UpdParams& p, int i, int k0, int k1) {
    constexpr int ES = (int)sizeof(T) * (CPLX ? 2 : 1);
    for (int t = 0; t < p.nterms; ++t) {
        const PmlTermDev& q = p.terms[t];
        if (i < q.lo[0] || i >= q.hi[0]) continue;
        const int ka = max(k0, q.lo[2]), kb = min(k1, q.hi[2]) - 1;
        if (ka > kb) continue;
        for (int j = q.lo[1] + (int)threadIdx.x; j < q.hi[1]; j += (int)blockDim.x) {
            const int ax = q.axis;
            const int na = (ax == 0 ? i - q.lo[0] : ax == 1 ? j - q.lo[1] : ka - q.lo[2]) + q.psi_off;
            const int nb = (ax == 0 ? i - q.lo[0] : ax == 1 ? j - q.lo[1] : kb - q.lo[2]) + q.psi_off;
            const size_t ia = ((size_t)(ax == 0 ? na : i) * q.pdim[1] + (ax == 1 ? na : j)) * q.pdim[2] + (ax == 2 ? na : ka);
            const size_t ib = ((size_t)(ax == 0 ? nb : i) * q.pdim[1] + (ax == 1 ? nb : j)) * q.pdim[2] + (ax == 2 ? nb : kb);
            const char* pa = (const char*)q.psi + ia * ES;
            const char* pb = (const char*)q.psi + ib * ES;
            asm volatile("prefetch.global.L2 [%0];" ::"l"(pa));
            if (((size_t)pa >> 7) != ((size_t)pb >> 7)) asm volatile("prefetch.global.L2 [%0];" ::"l"(pb));
        }
    }
}

// Tile-level classification against the three update boxes: returns the bit mask of the
// components whose box contains the whole tile [i0,i1) x [j0,j1) x [k0,k1) when every box
// either contains the tile or misses it completely, else -1 (per-cell tests needed).
__device__ __forceinline__ int tile_update_class(const UpdParams& p, int i0, int i1, int j0, int j1, int k0, int k1) {
    int upd = 0;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const Box& b = p.box[c];
        const bool inside = b.lo[0] <= i0 && i1 <= b.hi[0] && b.lo[1] <= j0 && j1 <= b.hi[1] && b.lo[2] <= k0 && k1 <= b.hi[2];
        const bool apart = b.hi[0] <= i0 || i1 <= b.lo[0] || b.hi[1] <= j0 || j1 <= b.lo[1] || b.hi[2] <= k0 || k1 <= b.lo[2];
        if (inside) upd |= 1 << c;
        else if (!apart) return -1;
    }
    return upd;
}

// Interior fast path of cell_update_regs: no CPML term touches the tile and the update boxes
// were resolved per tile (upd = tile_update_class(...) >= 0, CTA-uniform).  Same expression
// as the general path, so both give identical bits.
template <bool CPLX, int COMPS = 7>
__device__ __forceinline__ void cell_update_fast(const int upd, const double C,
                                                 const typename AccT<CPLX>::type (&d)[6],
                                                 typename AccT<CPLX>::type (&g)[3]) {
    if ((upd & COMPS) == COMPS) {
#pragma unroll
        for (int c = 0; c < 3; ++c)
            if (COMPS & (1 << c)) g[c] = a_add(g[c], a_scale(C, a_sub(d[2 * c], d[2 * c + 1])));
    } else {
#pragma unroll
        for (int c = 0; c < 3; ++c)
            if ((COMPS & (1 << c)) && (upd & (1 << c))) g[c] = a_add(g[c], a_scale(C, a_sub(d[2 * c], d[2 * c + 1])));
    }
}

// Scalar per-cell variant (loads and stores the field itself).
template <typename T, bool CPLX, bool PAL>
__device__ __forceinline__ void cell_update(const UpdParams& p, unsigned mask, int i, int j, int k,
                                            const typename AccT<CPLX>::type (&d)[6]) {
    using A = typename AccT<CPLX>::type;
    using E = Elem<T, CPLX>;
    const size_t idx = ((size_t)i * p.ny + j) * p.nz + k;
    A g[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) g[c] = E::ld(p.G[c], idx);
    double cf[1];
    ld_coeff<1, PAL>(p, idx, cf);
    cell_update_regs<T, CPLX>(p, mask, i, j, k, cf[0], d, g);
#pragma unroll
    for (int c = 0; c < 3; ++c) E::st(p.G[c], idx, g[c]);
}

}  // namespace ies
