// Register/shared-memory Stockham FFT building block for the spectral derivatives
// (replaces xp.fft.{r,}fftn / i{r,}fftn of space.py:145-162, call sites 709-755, 903-961).
//
// One FFT line of length N (power of two, 16..1024) is handled by T = N/16 threads,
// each holding 16 complex values in registers.  Stages are radix 16/16/(N/256);
// values cross threads through a shared-memory exchange buffer whose addressing is
// supplied by the caller (contiguous-line or strided-line policy).  The forward
// transform ends, and the inverse begins, with the same register<->index map, so the
// spectral multiplier ik*exp(+-ik d/2) is applied in registers with no exchange.
#pragma once
#include <cuda_runtime.h>

namespace ies {

template <typename T> struct Cx;
template <> struct Cx<float>  { using type = float2; };
template <> struct Cx<double> { using type = double2; };

// Every floating-point operation of the transforms is spelled with an explicit rounding
// intrinsic: the compiler may then neither contract a*b+c differently in two kernels that
// inline the same code nor re-associate, so every kernel variant (two-kernel path, fused
// single-launch path, any dtype instantiation) produces the same bits for the same input.
__device__ __forceinline__ float  r_add(float a, float b)   { return __fadd_rn(a, b); }
__device__ __forceinline__ double r_add(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ float  r_sub(float a, float b)   { return __fsub_rn(a, b); }
__device__ __forceinline__ double r_sub(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ float  r_mul(float a, float b)   { return __fmul_rn(a, b); }
__device__ __forceinline__ double r_mul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ float  r_fma(float a, float b, float c)    { return __fmaf_rn(a, b, c); }
__device__ __forceinline__ double r_fma(double a, double b, double c) { return __fma_rn(a, b, c); }

template <typename C> __device__ __forceinline__ C cadd(C a, C b) { C r; r.x = r_add(a.x, b.x); r.y = r_add(a.y, b.y); return r; }
template <typename C> __device__ __forceinline__ C csub(C a, C b) { C r; r.x = r_sub(a.x, b.x); r.y = r_sub(a.y, b.y); return r; }
template <typename C> __device__ __forceinline__ C cmul(C a, C b) {
    C r; r.x = r_fma(a.x, b.x, -r_mul(a.y, b.y)); r.y = r_fma(a.x, b.y, r_mul(a.y, b.x)); return r;
}
template <typename C> __device__ __forceinline__ C cmulc(C a, C b) {   // a * conj(b)
    C r; r.x = r_fma(a.x, b.x, r_mul(a.y, b.y)); r.y = r_fma(a.y, b.x, -r_mul(a.x, b.y)); return r;
}

// cos/sin(2*pi*k/16), k = 0..7
__device__ __forceinline__ constexpr double tw16_cos(int k) {
    return k == 0 ? 1.0 : k == 1 ? 0.92387953251128673848 : k == 2 ? 0.70710678118654752440
         : k == 3 ? 0.38268343236508977173 : k == 4 ? 0.0 : k == 5 ? -0.38268343236508977173
         : k == 6 ? -0.70710678118654752440 : -0.92387953251128673848;
}
__device__ __forceinline__ constexpr double tw16_sin(int k) {
    return k == 0 ? 0.0 : k == 1 ? 0.38268343236508977173 : k == 2 ? 0.70710678118654752440
         : k == 3 ? 0.92387953251128673848 : k == 4 ? 1.0 : k == 5 ? 0.92387953251128673848
         : k == 6 ? 0.70710678118654752440 : 0.38268343236508977173;
}

// In-register R-point DFT, decimation in time, natural order in and out.
// INV=false: exp(-2 pi i nk/R); INV=true: exp(+2 pi i nk/R), unnormalised.
template <int R, bool INV, typename C> struct Dft;

template <bool INV, typename C> struct Dft<1, INV, C> {
    static __device__ __forceinline__ void run(const C (&in)[1], C (&out)[1]) { out[0] = in[0]; }
};
template <bool INV, typename C> struct Dft<2, INV, C> {
    static __device__ __forceinline__ void run(const C (&in)[2], C (&out)[2]) {
        out[0] = cadd(in[0], in[1]);
        out[1] = csub(in[0], in[1]);
    }
};
template <int R, bool INV, typename C> struct Dft {
    static __device__ __forceinline__ void run(const C (&in)[R], C (&out)[R]) {
        C e[R / 2], o[R / 2], E[R / 2], O[R / 2];
#pragma unroll
        for (int i = 0; i < R / 2; ++i) { e[i] = in[2 * i]; o[i] = in[2 * i + 1]; }
        Dft<R / 2, INV, C>::run(e, E);
        Dft<R / 2, INV, C>::run(o, O);
#pragma unroll
        for (int k = 0; k < R / 2; ++k) {
            C t;
            constexpr int S = 16 / R;                  // index step into the 16th-root table
            if (k == 0) {
                t = O[k];
            } else if (4 * k == R) {                   // -i (forward) / +i (inverse)
                if (INV) { t.x = -O[k].y; t.y = O[k].x; } else { t.x = O[k].y; t.y = -O[k].x; }
            } else {
                using Tr = decltype(t.x);
                const Tr c = (Tr)tw16_cos(k * S), s = (Tr)tw16_sin(k * S);
                if (INV) { t.x = r_fma(O[k].x, c, -r_mul(O[k].y, s)); t.y = r_fma(O[k].x, s, r_mul(O[k].y, c)); }
                else     { t.x = r_fma(O[k].x, c, r_mul(O[k].y, s)); t.y = r_fma(O[k].y, c, -r_mul(O[k].x, s)); }
            }
            out[k]         = cadd(E[k], t);
            out[k + R / 2] = csub(E[k], t);
        }
    }
};

// Radix plan for VPT register values per thread (VPT = 16: radix 16/16/(N/256), one line =
// N/16 threads; VPT = 8: radix 8/8/(N/64), one line = N/8 threads, half the registers).
template <int N, int VPT> struct PlanV {
    static_assert(N >= 16 && N <= 1024 && (N & (N - 1)) == 0, "FFT length must be 16..1024, power of two");
    static_assert(VPT == 8 || VPT == 16, "values per thread");
    static constexpr int T  = N / VPT;                           // threads per line
    static constexpr int R1 = (N / VPT >= VPT) ? VPT : N / VPT;  // second radix (1 if N == VPT)
    static constexpr int R2 = N / (VPT * R1);                    // third radix
    static_assert(R2 <= VPT, "FFT length too large for this plan");
    static constexpr int RL = (R2 > 1) ? R2 : (R1 > 1 ? R1 : VPT);   // radix adjacent to the spectrum
};
template <int N> using Plan = PlanV<N, 16>;

// One Stockham stage on the VPT register values of a thread.
//   v[b*R + m] <-> input element  j + m*N/R,  j = t + T*b   (b < VPT/R butterflies per thread)
// LOAD : read inputs from the exchange buffer;  STORE: write outputs to it at
//   (j/NS)*NS*R + (j%NS) + m*NS.   Without STORE the outputs stay in v[b*R+m].
// tw: master twiddle table W_N[k] = exp(-2 pi i k/N) (shared or global memory).
// X::SPLIT (XchgContigSplit): the buffer holds one real per element; a STORE stage writes the real
// parts and leaves the imaginary parts in v[].y, the LOAD of the next stage (which is told that
// stage's radix PR and stride PNS) reads the real parts, then passes the imaginary parts through
// the same buffer.  Same values, same arithmetic -- only the route through shared memory differs.
template <int N, int VPT, int R, int NS, bool INV, bool LOAD, bool STORE, bool TWT = false, int PR = 1, int PNS = 1,
          typename C, typename X>
__device__ __forceinline__ void fft_stage_v(C (&v)[VPT], const int t, const C* __restrict__ tw, X& xb) {
    constexpr int T = N / VPT;
    constexpr int NB = VPT / R;
    if constexpr (LOAD && X::SPLIT) {
        decltype(v[0].x) re[VPT];
        xb.sync();
#pragma unroll
        for (int b = 0; b < NB; ++b) {
            const int j = t + T * b;
#pragma unroll
            for (int m = 0; m < R; ++m) re[b * R + m] = xb.ldx(j + m * (N / R));
        }
        xb.sync();
#pragma unroll
        for (int b = 0; b < VPT / PR; ++b) {
            const int j = t + T * b;
            const int base = (j / PNS) * PNS * PR + (j & (PNS - 1));
#pragma unroll
            for (int m = 0; m < PR; ++m) xb.stx(base + m * PNS, v[b * PR + m].y);
        }
        xb.sync();
#pragma unroll
        for (int b = 0; b < NB; ++b) {
            const int j = t + T * b;
#pragma unroll
            for (int m = 0; m < R; ++m) { v[b * R + m].x = re[b * R + m]; v[b * R + m].y = xb.ldx(j + m * (N / R)); }
        }
    } else if (LOAD) {
        xb.sync();
#pragma unroll
        for (int b = 0; b < NB; ++b) {
            const int j = t + T * b;
#pragma unroll
            for (int m = 0; m < R; ++m) v[b * R + m] = xb.ld(j + m * (N / R));
        }
    }
    if (STORE) xb.sync();
#pragma unroll
    for (int b = 0; b < NB; ++b) {
        const int j = t + T * b;
        C in[R], out[R];
#pragma unroll
        for (int m = 0; m < R; ++m) in[m] = v[b * R + m];
        if (NS > 1) {
            const int jj = j & (NS - 1);
            constexpr int step = N / (NS * R);
#pragma unroll
            for (int m = 1; m < R; ++m) {
                // TWT: tw is THIS stage's table transposed to [m][jj], row length NS (consecutive
                // threads read consecutive entries: no shared-memory bank conflicts); else the
                // master table W_N[k]
                const C w = TWT ? tw[m * NS + jj] : tw[(jj * m * step) & (N - 1)];
                in[m] = INV ? cmulc(in[m], w) : cmul(in[m], w);
            }
        }
        Dft<R, INV, C>::run(in, out);
        if constexpr (STORE && X::SPLIT) {
            const int base = (j / NS) * NS * R + (j & (NS - 1));
#pragma unroll
            for (int m = 0; m < R; ++m) { xb.stx(base + m * NS, out[m].x); v[b * R + m].y = out[m].y; }
        } else if (STORE) {
            const int base = (j / NS) * NS * R + (j & (NS - 1));
#pragma unroll
            for (int m = 0; m < R; ++m) xb.st(base + m * NS, out[m]);
        } else {
#pragma unroll
            for (int m = 0; m < R; ++m) v[b * R + m] = out[m];
        }
    }
}

// Spectrum index held in v[q] after fft_forward (and expected by fft_inverse).
template <int N, int VPT> __device__ __forceinline__ int spec_index_v(int t, int q) {
    constexpr int RL = PlanV<N, VPT>::RL;
    constexpr int T = N / VPT;
    const int b = q / RL, m = q % RL;
    return t + T * b + m * (N / RL);
}
// Line index held in v[q] before fft_forward and after fft_inverse.
template <int N, int VPT> __device__ __forceinline__ int line_index_v(int t, int q) { return t + q * (N / VPT); }

// Transposed stage tables (TWT): twf = table of the second forward stage (radix R1, NS = VPT):
// twf[m*VPT + jj] = W_N^(jj*m*N/(VPT*R1)); twi = table of the last inverse stage (radix VPT,
// NS = N/VPT): twi[m*(N/VPT) + jj] = W_N^(jj*m).  For N = VPT*VPT both are the same table.
// The remaining twiddled stages (only when R2 > 1) read the master table tw.
template <int N, int VPT> struct TwTables {
    using P = PlanV<N, VPT>;
    static constexpr int FWD = (N > VPT) ? P::R1 * VPT : 0;          // entries of twf
    static constexpr int INV = (N > VPT) ? N : 0;                    // entries of twi
    static constexpr bool SHARED = (N == VPT * VPT);                 // twf == twi
    static constexpr bool NEED_MASTER = P::R2 > 1;
    static constexpr int FSTEP = (N > VPT) ? N / (VPT * P::R1) : 1;
};

template <int N, int VPT, bool TWT = false, typename C, typename X>
__device__ __forceinline__ void fft_forward_v(C (&v)[VPT], int t, const C* tw, X& xb, const C* twf = nullptr) {
    using P = PlanV<N, VPT>;
    if (N == VPT) {
        fft_stage_v<N, VPT, VPT, 1, false, false, false>(v, t, tw, xb);
    } else {
        fft_stage_v<N, VPT, VPT, 1, false, false, true>(v, t, tw, xb);
        if (P::R2 == 1) {
            fft_stage_v<N, VPT, P::R1, VPT, false, true, false, TWT, VPT, 1>(v, t, TWT ? twf : tw, xb);
        } else {
            fft_stage_v<N, VPT, P::R1, VPT, false, true, true, TWT, VPT, 1>(v, t, TWT ? twf : tw, xb);
            fft_stage_v<N, VPT, (P::R2 > 1 ? P::R2 : 2), VPT * P::R1, false, true, false, false, P::R1, VPT>(v, t, tw, xb);
        }
    }
}

template <int N, int VPT, bool TWT = false, typename C, typename X>
__device__ __forceinline__ void fft_inverse_v(C (&v)[VPT], int t, const C* tw, X& xb, const C* twi = nullptr) {
    using P = PlanV<N, VPT>;
    if (N == VPT) {
        fft_stage_v<N, VPT, VPT, 1, true, false, false>(v, t, tw, xb);
    } else if (P::R2 == 1) {
        fft_stage_v<N, VPT, P::R1, 1, true, false, true>(v, t, tw, xb);
        fft_stage_v<N, VPT, VPT, P::R1, true, true, false, TWT, P::R1, 1>(v, t, TWT ? twi : tw, xb);
    } else {
        constexpr int R2 = (P::R2 > 1 ? P::R2 : 2);
        fft_stage_v<N, VPT, R2, 1, true, false, true>(v, t, tw, xb);
        fft_stage_v<N, VPT, P::R1, R2, true, true, true, false, R2, 1>(v, t, tw, xb);
        fft_stage_v<N, VPT, VPT, R2 * P::R1, true, true, false, TWT, P::R1, R2>(v, t, TWT ? twi : tw, xb);
    }
}

// 16 values per thread (the default plan)
template <int N> __device__ __forceinline__ int spec_index(int t, int q) { return spec_index_v<N, 16>(t, q); }
template <int N> __device__ __forceinline__ int line_index(int t, int q) { return line_index_v<N, 16>(t, q); }
template <int N, typename C, typename X>
__device__ __forceinline__ void fft_forward(C (&v)[16], int t, const C* tw, X& xb) { fft_forward_v<N, 16>(v, t, tw, xb); }
template <int N, typename C, typename X>
__device__ __forceinline__ void fft_inverse(C (&v)[16], int t, const C* tw, X& xb) { fft_inverse_v<N, 16>(v, t, tw, xb); }

// Exchange policies ---------------------------------------------------------------
// Contiguous lines (z axis): threads of a line are adjacent lanes; one line per
// T lanes; padded by one element per 16 to keep the radix-16 scatter conflict free.
template <typename C, int N> struct XchgContig {
    static constexpr bool SPLIT = false;
    C* base;            // this line's slice of the exchange buffer
    static constexpr int LS = N + N / 16;
    __device__ __forceinline__ C ld(int i) const { return base[i + (i >> 4)]; }
    __device__ __forceinline__ void st(int i, C v) const { base[i + (i >> 4)] = v; }
    __device__ __forceinline__ void sync() const {
        if (N / 16 <= 32) __syncwarp(); else __syncthreads();
    }
};
// Same for the radix-8 plan: one pad element per 8 keeps the radix-8 scatter (8t + m) conflict free.
template <typename C, int N> struct XchgContig8 {
    static constexpr bool SPLIT = false;
    C* base;
    static constexpr int LS = N + N / 8;
    __device__ __forceinline__ C ld(int i) const { return base[i + (i >> 3)]; }
    __device__ __forceinline__ void st(int i, C v) const { base[i + (i >> 3)] = v; }
    __device__ __forceinline__ void sync() const {
        if (N / 8 <= 32) __syncwarp(); else __syncthreads();
    }
};
// Contiguous lines without padding: element i lives at i ^ ((i >> 4) & 15), which keeps both
// the radix-16 scatter (16t + m) and the strided gather (t + 16m) conflict free while a line
// occupies exactly N elements (k_zline_update keeps its tile at 64 KB).
template <typename C, int N> struct XchgContigSw {
    static constexpr bool SPLIT = false;
    C* base;
    static constexpr int LS = N;
    static __device__ __forceinline__ int phys(int i) { return i ^ ((i >> 4) & 15); }
    __device__ __forceinline__ C ld(int i) const { return base[phys(i)]; }
    __device__ __forceinline__ void st(int i, C v) const { base[phys(i)] = v; }
    __device__ __forceinline__ void sync() const {
        if (N / 16 <= 32) __syncwarp(); else __syncthreads();
    }
};
// Contiguous lines, real and imaginary parts crossing one after the other (fft_stage_v, SPLIT): the
// line's slice holds N padded REALS -- half the shared memory of XchgContig (35 instead of 70 KB for
// the 8 lines of a 512-point tile) for twice the 8-byte accesses.  LS is in units of C.
template <typename C, int N> struct XchgContigSplit {
    static constexpr bool SPLIT = true;
    using R = decltype(C::x);
    R* base;
    static constexpr int LS = (N + N / 16) / 2;
    __device__ __forceinline__ R ldx(int i) const { return base[i + (i >> 4)]; }
    __device__ __forceinline__ void stx(int i, R v) const { base[i + (i >> 4)] = v; }
    __device__ __forceinline__ C ld(int i) const { return C(); }          // (never taken: SPLIT stages use ldx / stx)
    __device__ __forceinline__ void st(int, C) const {}
    __device__ __forceinline__ void sync() const {
        if (N / 16 <= 32) __syncwarp(); else __syncthreads();
    }
};
// Strided lines (y or x axis): W adjacent lines per CTA, lane index = column.
template <typename C, int W> struct XchgStrided {
    static constexpr bool SPLIT = false;
    C* base;            // buffer + column
    __device__ __forceinline__ C ld(int i) const { return base[i * W]; }
    __device__ __forceinline__ void st(int i, C v) const { base[i * W] = v; }
    __device__ __forceinline__ void sync() const { __syncthreads(); }
};

}  // namespace ies
