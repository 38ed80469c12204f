// Global-memory accessors carrying an L2 eviction policy (createpolicy + .L2::cache_hint).
// The chunked SHPF half-step keeps the z-derivative scratch of a few x-planes resident in
// the 126 MB L2 between the kernel that writes it and the kernel that reads it: those
// accesses are tagged evict_last, the once-through field traffic evict_first.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace ies {

enum { POL_NONE = 0, POL_LAST = 1, POL_FIRST = 2 };

__device__ __forceinline__ uint64_t make_policy(int kind) {
    uint64_t p;
    if (kind == POL_LAST)       asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    else if (kind == POL_FIRST) asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    else                        asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(p));
    return p;
}

// S is a 4-, 8- or 16-byte POD.  HINT = false -> plain access (pol ignored).
template <bool HINT, typename S>
__device__ __forceinline__ S ld_pol(const S* p, uint64_t pol) {
    if constexpr (!HINT) return *p;
    S out;
    if constexpr (sizeof(S) == 4) {
        uint32_t r;
        asm("ld.global.L2::cache_hint.b32 %0, [%1], %2;" : "=r"(r) : "l"(p), "l"(pol));
        __builtin_memcpy(&out, &r, 4);
    } else if constexpr (sizeof(S) == 8) {
        uint64_t r;
        asm("ld.global.L2::cache_hint.b64 %0, [%1], %2;" : "=l"(r) : "l"(p), "l"(pol));
        __builtin_memcpy(&out, &r, 8);
    } else {
        static_assert(sizeof(S) == 16, "ld_pol: 4, 8 or 16 bytes");
        uint64_t r[2];
        asm("ld.global.L2::cache_hint.v2.b64 {%0,%1}, [%2], %3;" : "=l"(r[0]), "=l"(r[1]) : "l"(p), "l"(pol));
        __builtin_memcpy(&out, r, 16);
    }
    return out;
}

template <bool HINT, typename S>
__device__ __forceinline__ void st_pol(S* p, const S& v, uint64_t pol) {
    if constexpr (!HINT) { *p = v; return; }
    if constexpr (sizeof(S) == 4) {
        uint32_t r; __builtin_memcpy(&r, &v, 4);
        asm volatile("st.global.L2::cache_hint.b32 [%0], %1, %2;" :: "l"(p), "r"(r), "l"(pol) : "memory");
    } else if constexpr (sizeof(S) == 8) {
        uint64_t r; __builtin_memcpy(&r, &v, 8);
        asm volatile("st.global.L2::cache_hint.b64 [%0], %1, %2;" :: "l"(p), "l"(r), "l"(pol) : "memory");
    } else {
        static_assert(sizeof(S) == 16, "st_pol: 4, 8 or 16 bytes");
        uint64_t r[2]; __builtin_memcpy(r, &v, 16);
        asm volatile("st.global.L2::cache_hint.v2.b64 [%0], {%1,%2}, %3;" :: "l"(p), "l"(r[0]), "l"(r[1]), "l"(pol) : "memory");
    }
}

}  // namespace ies
