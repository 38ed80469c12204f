// SHPF half-step as ONE launch (space.py:709-727 + 801-811 for updateH, 953-970 + 1017-1025
// for updateE, CPML 1110-1712): the z-line derivative tiles and the y-line update tiles of
// spectral.cuh run as two ROLES of the same grid, ordered by a ticket so that the z tiles of
// plane p are handed out LEAD planes ahead of the y tiles that consume them.
//
//   ticket t -> group g = t / (ZT + YT), slot r = t % (ZT + YT)
//      r <  ZT : z role, tile r of plane g            (if g < nx)
//      r >= ZT : y role, tile r - ZT of plane g - LEAD (if g >= LEAD)
//
// Why: the two-kernel path writes the z-derivative scratch (2 arrays) to HBM in one sweep and
// reads it back a whole sweep later, and reads F_x / F_y once per sweep: 6 of its 15-16 array
// passes.  Here a scratch plane is consumed a few planes after it was produced, from a RING of
// `ring` planes that stays in the 126 MB L2, and the second reads of F_x / F_y hit L2 as well.
//
// Ordering: zdone[p] / ydone[p] count the finished tiles of plane p (the launcher zeroes the
// counters and the ticket in stream order before every launch).  A y tile waits for
// zdone[plane] after its own FFT phase; a z tile waits for ydone[plane - ring] before it
// overwrites a ring slot.  Both waits are on tickets handed out EARLIER (ring > LEAD), and a
// ticket is drawn by a CTA that is already resident, so every wait is on a running or finished
// CTA: no deadlock whatever the hardware's block scheduling order is.
//
// Arithmetic is the same expression on the same operands as in k_zline + k_yline_update, so the
// fields are bit-identical to the two-kernel path.
#pragma once
#include <algorithm>
#include "spectral.cuh"

namespace ies {

struct FusedParams {
    unsigned* ticket;           // one counter
    unsigned* zdone;            // [nx]
    unsigned* ydone;            // [nx]
    unsigned* pair;             // [nx][yt/2]: finished y tiles per tile pair (tiles narrower than a 128-byte line share the discard)
    int lead;                   // planes the z role runs ahead of the y role
    int ring;                   // scratch planes (ring > lead; ring >= nx: no wrap)
    int zt, yt;                 // tiles per plane of each role
    int prefetch;               // z role prefetches the y role's streaming operands of its rows into L2
    unsigned long long* prof;   // null, or 16 counters of SM cycles per role phase (development: option fused_prof)
};

// development timers: thread 0 of a CTA adds the cycles since *t0 to prof[slot] and restarts the clock
__device__ __forceinline__ void fprof(const FusedParams& fp, int slot, long long* t0) {
    if (fp.prof && threadIdx.x == 0) {
        const long long t = clock64();
        atomicAdd(fp.prof + slot, (unsigned long long)(t - *t0));
        *t0 = t;
    }
}

// FENCE: acquire fence after the poll.  Needed when the data behind the counter may sit stale in
// this SM's L1 (a re-used ring slot); with a full-size scratch every line is read by exactly one
// CTA per launch and L1 starts every launch empty, so the poll + barrier order the loads and the
// fence (an L1 invalidation for the whole SM, also for the CTA next door) is skipped.
__device__ __forceinline__ void wait_counter(const unsigned* ctr, unsigned target, bool fence) {
    if (threadIdx.x == 0) {
        while (*(volatile const unsigned*)ctr < target) __nanosleep(20);
        if (fence) __threadfence();
    }
    __syncthreads();
}
__device__ __forceinline__ void signal_counter(unsigned* ctr) {
    __syncthreads();            // every thread's stores / loads of the tile are issued
    if (threadIdx.x == 0) {
        __threadfence();        // release (cumulative over the CTA through the barrier)
        atomicAdd(ctr, 1u);
    }
}

// z role: LPB adjacent z lines of plane pz -> ring slot.  Same transform as k_zline; the
// exchange buffer is the unpadded swizzled one (64 KB like the y role's stash) and the stage
// tables come transposed from global memory (twt = [forward stage | inverse stage]).
// ZM = 0: tables from global memory (L1), unpadded swizzled exchange (64 KB of shared memory);
// ZM = 1: k_zline's layout -- stage tables + multiplier copied to shared memory, padded exchange.
// ZM = 2: as 1 with the split real / imaginary exchange (XchgContigSplit): 64 instead of 98 KB at N = 512,
//         below the y role's 80 KB, so the CTA pair leaves 90 KB of L1 to the streaming phase.
template <int N, int ZM> struct FusedZSmem {
    using TT_ = TwTables<N, 16>;
    static constexpr int TABLES = (ZM >= 1) ? ((TT_::NEED_MASTER ? N : 0) + (TT_::SHARED ? 0 : TT_::FWD) + TT_::INV + N) : 0;
    static constexpr int LS = (ZM == 2) ? (N + N / 16) / 2 : ((ZM == 1) ? (N + N / 16) : N);
    static constexpr int ELEMS = TABLES + ZCfg<N, 16>::LPB * LS;      // complex elements
};

// ZB: z tiles per CTA, done one after the other with the tables staged once and ONE signal at the
// end -- the release fence in front of the signal holds the CTA's slot until its stores are
// acknowledged (2.9 k cycles per tile when every tile signals for itself).
template <typename T, bool CPLX, int N, int ZM, int ZB>
__device__ __forceinline__ void fused_z_role(const UpdParams& p, const FusedParams& fp, const int pz, const int tile0,
                                             typename Cx<T>::type* smem,
                                             const typename Cx<T>::type* __restrict__ tw,
                                             const typename Cx<T>::type* __restrict__ twt,
                                             const typename Cx<T>::type* __restrict__ ml) {
    using C = typename Cx<T>::type;
    using F = Fld<T, CPLX>;
    using X = typename std::conditional<ZM == 2, XchgContigSplit<C, N>,
                  typename std::conditional<ZM == 1, XchgContig<C, N>, XchgContigSw<C, N>>::type>::type;
    using TT_ = TwTables<N, 16>;
    constexpr bool TWT = N > 16;
    constexpr int TT = ZCfg<N, 16>::TT, LPB = ZCfg<N, 16>::LPB;
    const C* twf = twt;
    const C* twi = TT_::SHARED ? twt : twt + TT_::FWD;
    C* xbuf = smem;
    const int t = threadIdx.x % TT, l = threadIdx.x / TT;
    // The first tile's field loads are issued BEFORE the tables are staged, so that the two global round
    // trips (tables, fields) overlap instead of following each other across the staging barrier.
    C v[F::NF][16];
    auto load_tile = [&](const int zb) {
        const int row = (tile0 + zb) * LPB + l;
        const size_t ibase = ((size_t)pz * p.ny + row) * N;
#pragma unroll
        for (int f = 0; f < F::NF; ++f) {
#pragma unroll
            for (int q = 0; q < 16; ++q) {
                if (row < p.ny) v[f][q] = F::ld(p.F[1], p.F[0], ibase + line_index_v<N, 16>(t, q), f);
                else { v[f][q].x = 0; v[f][q].y = 0; }
            }
        }
    };
    constexpr bool PRELOAD = F::NF == 1;
    if constexpr (PRELOAD) load_tile(0);
    if constexpr (ZM >= 1) {
        C* s_tw = smem;
        C* s_twf = s_tw + (TT_::NEED_MASTER ? N : 0);
        C* s_twi = TT_::SHARED ? s_twf : s_twf + TT_::FWD;
        C* s_ml = s_twi + TT_::INV;
        xbuf = s_ml + N;
        for (int q = threadIdx.x; q < N; q += blockDim.x) {
            s_ml[q] = ml[q];
            if (TT_::NEED_MASTER) s_tw[q] = tw[q];
            if (N > 16) {
                s_twi[q] = twi[q];
                if (!TT_::SHARED && q < TT_::FWD) s_twf[q] = twf[q];
            }
        }
        __syncthreads();
        tw = s_tw; twf = s_twf; twi = s_twi; ml = s_ml;
    }
    X xb{reinterpret_cast<decltype(X::base)>(xbuf + (size_t)l * X::LS)};
    long long t0 = fp.prof ? clock64() : 0;
    // Experiment (option fused_prefetch, off): the z role asks L2 for the rows of F_z and the three G
    // arrays its plane's y tiles will stream `lead` planes later.  Measured: y update phase 17.7 k ->
    // 18.8 k cycles, 3.07 -> 3.22 ms/step -- the update phase is not waiting for HBM latency (it is
    // paced by the L1/LSU wavefront rate of its 16-byte row-segment accesses), so L2 hits do not help.
    if (fp.prefetch) {
        constexpr int ES = (int)sizeof(T) * (CPLX ? 2 : 1);
        const size_t rows0 = ((size_t)pz * p.ny + (size_t)tile0 * LPB) * N * ES;
        const int bytes = min(ZB * LPB, p.ny - tile0 * LPB) * N * ES;
        const char* arr[4] = {(const char*)p.F[2], (const char*)p.G[0], (const char*)p.G[1], (const char*)p.G[2]};
#pragma unroll
        for (int a = 0; a < 4; ++a)
            for (int off = (int)threadIdx.x * 128; off < bytes; off += (int)blockDim.x * 128)
                asm volatile("prefetch.global.L2 [%0];" ::"l"(arr[a] + rows0 + off));
    }
#pragma unroll 1
    for (int zb = 0; zb < ZB; ++zb) {
    const int row = (tile0 + zb) * LPB + l;
    const bool ok = row < p.ny;
    const size_t obase = ((size_t)(pz % fp.ring) * p.ny + row) * N;
    if (!PRELOAD || zb > 0) load_tile(zb);
#pragma unroll
    for (int f = 0; f < F::NF; ++f) {
        fft_forward_v<N, 16, TWT>(v[f], t, tw, xb, twf);
#pragma unroll
        for (int q = 0; q < 16; ++q) v[f][q] = cmul(v[f][q], ml[spec_index_v<N, 16>(t, q)]);
        fft_inverse_v<N, 16, TWT>(v[f], t, tw, xb, twi);
    }
    fprof(fp, 0, &t0);
    // the slot's previous tenant (plane pz - ring) must have been consumed
    if (pz - fp.ring >= p.i0) wait_counter(fp.ydone + (pz - fp.ring), (unsigned)fp.yt, true);
    fprof(fp, 1, &t0);
    if (ok) {
        void* dA = const_cast<void*>(p.dz[0]);
        void* dB = const_cast<void*>(p.dz[1]);
#pragma unroll
        for (int f = 0; f < F::NF; ++f) {
#pragma unroll
            for (int q = 0; q < 16; ++q) F::st(dA, dB, obase + line_index_v<N, 16>(t, q), v[f][q], f);
        }
    }
    if (zb + 1 < ZB) { xb.sync(); continue; }       // the exchange buffer is re-used by the next tile
    signal_counter(fp.zdone + pz);
    fprof(fp, 2, &t0);
    if (fp.prof && threadIdx.x == 0) atomicAdd(fp.prof + 8, (unsigned long long)ZB);
    }
}

// YM = 0: the y role reads its twiddle / multiplier tables from global memory (L1), like
// k_yline_update; YM = 1: it copies them behind the stash in shared memory first (+2N entries).
template <typename T, bool CPLX, int NY, int NZ, int ZM, int YM, int ZB>
__global__ void __launch_bounds__(256, 2)
k_shpf_fused(const UpdParams p, const FusedParams fp,
             const typename Cx<T>::type* __restrict__ twy, const typename Cx<T>::type* __restrict__ mly,
             const typename Cx<T>::type* __restrict__ twz, const typename Cx<T>::type* __restrict__ twzt,
             const typename Cx<T>::type* __restrict__ mlz) {
    using C = typename Cx<T>::type;
    static_assert(!CPLX, "fused SHPF kernel: real field dtypes (complex tiles are 128 threads wide)");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    C* xbuf = reinterpret_cast<C*>(smem_raw);
    __shared__ unsigned s_ticket;
    // Measured alternatives: persistent CTAs looping over tickets (the loop keeps
    // both roles' state live: 740 B of spills, 4.3 ms/step); one CTA doing a z tile and then a y
    // tile (3.05 ms/step: all CTAs run the same phase sequence and the phases overlap less);
    // tables through L1 instead of shared memory (z FFT 14.6 k cycles per tile instead of 8.5 k,
    // y FFT 16.2 k instead of 12.6 k).
    if (threadIdx.x == 0) s_ticket = atomicAdd(fp.ticket, 1u);
    __syncthreads();
    const int zc = fp.zt / ZB;                      // z CTAs per plane (the launcher picks ZB | zt)
    const int per = zc + fp.yt;
    const int g = (int)(s_ticket / (unsigned)per), r = (int)(s_ticket % (unsigned)per);
    const int nplanes = p.i1 - p.i0;
    if (r < zc) {
        if (g >= nplanes) return;
        fused_z_role<T, CPLX, NZ, ZM, ZB>(p, fp, p.i0 + g, r * ZB, xbuf, twz, twzt, mlz);
    } else {
        const int py = g - fp.lead;
        if (py < 0) return;
        const int i = p.i0 + py;
        const int kb = r - zc;
        long long t0 = fp.prof ? clock64() : 0;
        if (p.nterms) prefetch_tile_psi<T, CPLX>(p, i, kb * YCfg<T, CPLX, NY>::W, min((kb + 1) * YCfg<T, CPLX, NY>::W, p.nz));
        // y tables (master W_N + multiplier) behind the stash; no barrier of their own: they are first
        // read after the two barriers of the first exchange
        C* s_tw = xbuf + (size_t)NY * YCfg<T, CPLX, NY>::W;
        C* s_ml = s_tw + NY;
        for (int q = threadIdx.x; q < NY; q += blockDim.x) { s_tw[q] = twy[q]; s_ml[q] = mly[q]; }
        yline_phase_a<T, CPLX, NY>(p, i, kb * YCfg<T, CPLX, NY>::W, xbuf, s_tw, s_ml);
        fprof(fp, 3, &t0);
        wait_counter(fp.zdone + i, (unsigned)zc, fp.ring < p.i1 - p.i0);     // also the barrier that publishes the stash
        fprof(fp, 4, &t0);
        const long long plane = (long long)p.ny * p.nz;
        const long long dz_off = ((long long)(i % fp.ring) - (long long)i) * plane;
        yline_phase_b_dispatch<T, CPLX, NY, false, true>(p, i, kb, fp.yt, xbuf, dz_off);
        if constexpr (YCfg<T, CPLX, NY>::W * sizeof(T) == 64) {
            // 8-column tiles (N = 512): a 128-byte scratch line belongs to two adjacent tiles; the one that
            // finishes second drops the pair's lines from L2 (see yline_phase_b for the wide-tile case)
            if (p.dz_discard) {
                __syncthreads();                    // this tile's loads are consumed
                if (threadIdx.x == 0) s_ticket = atomicAdd(fp.pair + (size_t)(i - p.i0) * (fp.yt / 2) + kb / 2, 1u);
                __syncthreads();
                const int k = (kb & ~1) * YCfg<T, CPLX, NY>::W;
                if (s_ticket == 1u && k >= p.dz_keep_lo && k + 16 <= p.dz_keep_hi) {
                    for (int j = (int)threadIdx.x; j < p.ny; j += (int)blockDim.x) {
                        const size_t e = (size_t)((long long)i * plane + (long long)j * p.nz + k + dz_off);
                        asm volatile("discard.global.L2 [%0], 128;" :: "l"((const T*)p.dz[0] + e) : "memory");
                        asm volatile("discard.global.L2 [%0], 128;" :: "l"((const T*)p.dz[1] + e) : "memory");
                    }
                }
            }
        }
        // only a z tile that re-uses this plane's ring slot waits for it
        if (i + fp.ring < p.i1) signal_counter(fp.ydone + i);
        fprof(fp, 5, &t0);
        if (fp.prof && threadIdx.x == 0) atomicAdd(fp.prof + 9, 1ull);
    }
}

// Host side: launch the fused half-step on planes [p.i0, p.i1).  Returns 0 / 1 like the other
// launchers, 2 when this (dtype, ny, nz) has no fused instantiation (caller takes the two-kernel path).
template <typename T, bool CPLX>
int launch_shpf_fused(Ctx* c, const UpdParams& p, int half) {
    if constexpr (CPLX) { return 2; }
    else {
        using C = typename Cx<T>::type;
        const int ny = c->cfg.ny, nz = c->cfg.nz;
        if (ny != nz) return 2;
        if (p.i1 <= p.i0) return 0;
        if (p.Cidx != nullptr) return 2;            // palette-form coefficients: two-kernel path
        FusedParams fp;
        fp.ticket = c->fused_sync; fp.zdone = c->fused_sync + 1; fp.ydone = c->fused_sync + 1 + c->cfg.nx;
        fp.lead = c->fused_lead;
        fp.ring = c->fused_ring_planes;
        fp.prof = c->fused_prof;
        fp.prefetch = c->fused_prefetch;
        // (measured alternative: counters that only grow, target = tiles x launch number, no memset --
        //  0.8% SLOWER on the headline in a same-box A/B of the two builds, 2.923 vs 2.895 ms; kept the memset)
        fp.pair = c->fused_sync + 1 + 2 * c->cfg.nx;
        const bool pairs = p.dz_discard && ny == 512;
        IES_CUDA(cudaMemsetAsync(c->fused_sync, 0, sizeof(unsigned) * (size_t)(1 + 2 * c->cfg.nx + (pairs ? 32 * c->cfg.nx : 0)), c->stream));
        const int nplanes = p.i1 - p.i0;
#define F_CASE(NN) {                                                                        \
            fp.zt = (ny + ZCfg<NN, 16>::LPB - 1) / ZCfg<NN, 16>::LPB;                       \
            fp.yt = (nz + YCfg<T, CPLX, NN>::W - 1) / YCfg<T, CPLX, NN>::W;                  \
            const size_t sm0 = sizeof(C) * (size_t)NN * YCfg<T, CPLX, NN>::W;                \
            constexpr int ZM_ = (NN == 512) ? 2 : 1;                                        \
            const size_t sm = std::max(sizeof(C) * (size_t)FusedZSmem<NN, ZM_>::ELEMS, sm0 + sizeof(C) * 2 * NN); \
            const int zb = (c->fused_zb == 4 && fp.zt % 4 == 0) ? 4 : (c->fused_zb >= 2 && fp.zt % 2 == 0) ? 2 : 1; \
            auto kern = zb == 4 ? k_shpf_fused<T, CPLX, NN, NN, ZM_, 1, 4> : zb == 2 ? k_shpf_fused<T, CPLX, NN, NN, ZM_, 1, 2> \
                                                                             : k_shpf_fused<T, CPLX, NN, NN, ZM_, 1, 1>; \
            if (set_smem(kern, sm)) return 1;                                               \
            const unsigned grid = (unsigned)((nplanes + fp.lead) * (fp.zt / zb + fp.yt));          \
            kern<<<grid, 256, sm, c->stream>>>(p, fp, (const C*)c->tw[1], (const C*)c->mult[half][1], \
                (const C*)c->tw[2], (const C*)c->twz_t, (const C*)c->mult[half][2]);        \
        }
        prof_mark(c, PROF_YLINE_UPDATE, 0);
        switch (ny) {
            case 64: F_CASE(64); break;
            case 128: F_CASE(128); break;
            case 256: F_CASE(256); break;
            case 512: F_CASE(512); break;
            default: return 2;
        }
        prof_mark(c, PROF_YLINE_UPDATE, 1);
#undef F_CASE
        count_launch();
        IES_CUDA(cudaGetLastError());
        return 0;
    }
}

}  // namespace ies
