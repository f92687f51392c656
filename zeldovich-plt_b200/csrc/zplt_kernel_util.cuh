// Small device helpers shared by the FFT / emission translation units: streaming global accesses, L2 residency
// hints, 256-bit record stores, mbarrier + TMA wrappers, warp reductions for the emission statistics.
#pragma once
#include <cuda.h>  // CUtensorMap (types only)

#include "zplt_device.cuh"
#include "zplt_internal.h"

namespace zplt {

// streaming (read-once / write-once) global accesses: do not keep the lines in L1
__device__ __forceinline__ cplx ld_stream(const cplx *p) {
    cplx r;
    asm volatile("ld.global.L1::no_allocate.v2.f64 {%0, %1}, [%2];" : "=d"(r.x), "=d"(r.y) : "l"(p));
    return r;
}
__device__ __forceinline__ void st_stream(cplx *p, cplx v) { __stcs(p, v); }

// resident CTAs per SM the register budget is tuned for: 512 threads of 128 registers fill an SM
constexpr int min_ctas(int threads) { return threads >= 512 ? 1 : (512 / threads > 8 ? 8 : 512 / threads); }

__device__ __forceinline__ unsigned smid() {
    unsigned r;
    asm("mov.u32 %0, %%smid;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// L2 residency of the emission kernel's parking areas: they are rewritten for every tile (128 KB per SM, 19 MB in all)
// while ~160 KB per tile stream past them; without a hint ~9 GB of them per pass were written back to HBM (ncu:
// 43.3 GB written against 34.4 GB of records).  Parked values are stored and loaded with an evict_last policy.
__device__ __forceinline__ uint64_t l2_evict_last() {
    uint64_t p;
    asm("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
// One 32-byte RVZel record as a single 256-bit store (STG.E.ENL2.256 on sm_100): every lane fills a whole sector with one
// instruction instead of two 16-byte halves at a 32-byte stride.  Needs 32-byte aligned records (EmitParams::wide_records).
__device__ __forceinline__ void st_record32(unsigned char *rec, float4 lo, float4 hi) {
    asm volatile("st.global.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(rec), "f"(lo.x), "f"(lo.y), "f"(lo.z), "f"(lo.w),
                 "f"(hi.x), "f"(hi.y), "f"(hi.z), "f"(hi.w)
                 : "memory");
}
template <bool G>
__device__ __forceinline__ void park_st(float *p, float v, uint64_t pol) {
    if constexpr (G)
        asm volatile("st.global.L2::cache_hint.f32 [%0], %1, %2;" ::"l"(p), "f"(v), "l"(pol) : "memory");
    else
        *p = v;
}
template <bool G>
__device__ __forceinline__ float park_ld(const float *p, uint64_t pol) {
    if constexpr (G) {
        float v;
        asm volatile("ld.global.L2::cache_hint.f32 %0, [%1], %2;" : "=f"(v) : "l"(p), "l"(pol) : "memory");
        return v;
    } else {
        return *p;
    }
}
__device__ __forceinline__ void park_std(double *p, double v, uint64_t pol) {
    asm volatile("st.global.L2::cache_hint.f64 [%0], %1, %2;" ::"l"(p), "d"(v), "l"(pol) : "memory");
}
__device__ __forceinline__ double park_ldd(const double *p, uint64_t pol) {
    double v;
    asm volatile("ld.global.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(v) : "l"(p), "l"(pol) : "memory");
    return v;
}
__device__ __forceinline__ void park_st2(float2 *p, float2 v, uint64_t pol) {
    asm volatile("st.global.L2::cache_hint.v2.f32 [%0], {%1, %2}, %3;" ::"l"(p), "f"(v.x), "f"(v.y), "l"(pol) : "memory");
}
__device__ __forceinline__ float2 park_ld2(const float2 *p, uint64_t pol) {
    float2 v;
    asm volatile("ld.global.L2::cache_hint.v2.f32 {%0, %1}, [%2], %3;" : "=f"(v.x), "=f"(v.y) : "l"(p), "l"(pol) : "memory");
    return v;
}


// ------------------------------------------------------------------ mbarrier + TMA
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t) __cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, unsigned parity) {
    asm volatile(
       "{\n"
       ".reg .pred P1;\n"
       "ZPLT_WAIT:\n"
       "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
       "@P1 bra ZPLT_DONE;\n"
       "bra ZPLT_WAIT;\n"
       "ZPLT_DONE:\n"
       "}" ::"r"(smem_u32(bar)),
       "r"(parity)
       : "memory");
}
// one slice of a tile: box (2T doubles along x, 1 row, M points along the transform axis, 1 array)
__device__ __forceinline__ void tma_load_4d(void *dst, const CUtensorMap *map, int c0, int c1, int c2, int c3, uint64_t *bar) {
    asm volatile(
       "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];" ::"r"(
          smem_u32(dst)),
       "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(smem_u32(bar))
       : "memory");
}


// ------------------------------------------------------------------ waiting for another kernel's progress
// The z pass + exchange kernel of a slab rank is resident for the whole of stage 1 and consumes row groups as the generation
// kernels (another stream) complete them: after each generation kernel the host enqueues a stream-ordered memset of one flag.
__device__ __forceinline__ unsigned int ld_volatile_u32(const unsigned int *p) {
    unsigned int v;
    asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// one thread: true when flags[grp] is set, false after ~15 s of polling
__device__ __forceinline__ bool wait_group(const unsigned int *flags, int grp) {
    if (ld_volatile_u32(flags + grp) != 0u) return true;
    const long long t0 = clock64();
    while (ld_volatile_u32(flags + grp) == 0u) {
        __nanosleep(256);
        if (clock64() - t0 > 30000000000ll) return false;
    }
    __threadfence();
    return true;
}

// ------------------------------------------------------------------ statistics
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// Reduce three per-thread partials over the warp and leave them in s_red[warp][slot]
// (slot 0 is a sum, every other slot a maximum; slot 7 is scratch).
// ACC: the CTA walks several tiles (persistent kernel) and accumulates into s_red, which it zeroed at the start.
template <int NT, bool ACC = false>
__device__ __forceinline__ void fold_stats(double (*s_red)[8], int tid, double q0, int s0, double q1, int s1, double q2, int s2) {
    if constexpr (NT >= 32) {
        q0 = (s0 == 0) ? warp_sum(q0) : warp_max(q0);
        q1 = warp_max(q1);
        q2 = warp_max(q2);
        if ((tid & 31) == 0) {
            double *r = s_red[tid >> 5];
            if constexpr (ACC) {
                r[s0] = (s0 == 0) ? r[s0] + q0 : fmax(r[s0], q0);
                r[s1] = fmax(r[s1], q1), r[s2] = fmax(r[s2], q2);
            } else {
                r[s0] = q0, r[s1] = q1, r[s2] = q2;
            }
        }
    } else {
        s_red[tid][s0] = q0, s_red[tid][s1] = q1, s_red[tid][s2] = q2;
    }
}


}  // namespace zplt
