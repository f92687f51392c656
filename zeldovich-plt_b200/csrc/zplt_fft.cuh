// FP64 complex 1-D backward FFT of power-of-two length N (16..4096), register-resident.
//
// A pencil of N points is owned by M = N/16 threads; each thread keeps 16 complex
// values in registers.  The transform is 1..3 Stockham-style passes of radix
// (16, R2, R3) with R2 = min(16, N/16), R3 = N/(16*R2); radix-2/4/8 passes run 8/4/2
// butterflies per thread.  Between passes the pencil goes through shared memory once
// (write, barrier, read); inter-pass twiddles come from a W_N table in global memory
// (L1-resident).  Input and output use the same ownership: register e of slot b holds
// element b + M*e, so the caller's global loads and stores are identical for every N.
//
// Replaces fftw_execute_dft on plan1d/plan2d (reference src/zeldovich.cpp:83-114,
// sign +1, unnormalised).  Bandwidth-bound FP64: no tensor cores.
#pragma once
#include "zplt_device.cuh"

namespace zplt {

__device__ __forceinline__ cplx cadd(cplx a, cplx b) { return make_double2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ cplx csub(cplx a, cplx b) { return make_double2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ cplx cmul(cplx a, cplx b) { return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
__device__ __forceinline__ cplx cmuli(cplx a) { return make_double2(-a.y, a.x); }  // * (+i)

// ---- small backward DFTs on registers, natural order in and out -------------------
__device__ __forceinline__ void dft2(cplx &a, cplx &b) {
    cplx t = csub(a, b);
    a      = cadd(a, b);
    b      = t;
}
__device__ __forceinline__ void dft4(cplx &v0, cplx &v1, cplx &v2, cplx &v3) {
    cplx t0 = cadd(v0, v2), t1 = csub(v0, v2), t2 = cadd(v1, v3), t3 = cmuli(csub(v1, v3));
    v0 = cadd(t0, t2);
    v1 = cadd(t1, t3);
    v2 = csub(t0, t2);
    v3 = csub(t1, t3);
}
#define ZPLT_SQRT1_2 0.70710678118654752440
#define ZPLT_COS_PI_8 0.92387953251128675613
#define ZPLT_SIN_PI_8 0.38268343236508977173
__device__ __forceinline__ void dft8(cplx &v0, cplx &v1, cplx &v2, cplx &v3, cplx &v4, cplx &v5, cplx &v6, cplx &v7) {
    // decimation in time: evens and odds
    dft4(v0, v2, v4, v6);
    dft4(v1, v3, v5, v7);
    // odd outputs times W8^k, W8 = exp(+i pi/4)
    cplx o1 = make_double2(ZPLT_SQRT1_2 * (v3.x - v3.y), ZPLT_SQRT1_2 * (v3.x + v3.y));
    cplx o2 = cmuli(v5);
    cplx o3 = make_double2(-ZPLT_SQRT1_2 * (v7.x + v7.y), ZPLT_SQRT1_2 * (v7.x - v7.y));
    cplx e0 = v0, e1 = v2, e2 = v4, e3 = v6, o0 = v1;
    v0 = cadd(e0, o0);
    v4 = csub(e0, o0);
    v1 = cadd(e1, o1);
    v5 = csub(e1, o1);
    v2 = cadd(e2, o2);
    v6 = csub(e2, o2);
    v3 = cadd(e3, o3);
    v7 = csub(e3, o3);
}
__device__ __forceinline__ void dft16(cplx (&v)[16]) {
    dft8(v[0], v[2], v[4], v[6], v[8], v[10], v[12], v[14]);
    dft8(v[1], v[3], v[5], v[7], v[9], v[11], v[13], v[15]);
    // E[k] = v[2k], O[k] = v[2k+1]; X[k] = E[k] + W16^k O[k], X[k+8] = E[k] - W16^k O[k]
    const double c1 = ZPLT_COS_PI_8, s1 = ZPLT_SIN_PI_8, r = ZPLT_SQRT1_2;
    cplx o[8];
    o[0] = v[1];
    o[1] = cmul(v[3], make_double2(c1, s1));
    o[2] = make_double2(r * (v[5].x - v[5].y), r * (v[5].x + v[5].y));
    o[3] = cmul(v[7], make_double2(s1, c1));
    o[4] = cmuli(v[9]);
    o[5] = cmul(v[11], make_double2(-s1, c1));
    o[6] = make_double2(-r * (v[13].x + v[13].y), r * (v[13].x - v[13].y));
    o[7] = cmul(v[15], make_double2(-c1, s1));
    cplx e[8];
#pragma unroll
    for (int k = 0; k < 8; k++) e[k] = v[2 * k];
#pragma unroll
    for (int k = 0; k < 8; k++) {
        v[k]     = cadd(e[k], o[k]);
        v[k + 8] = csub(e[k], o[k]);
    }
}

// w^1 .. w^(R-1) from w by a shallow multiplication tree (depth <= 4 products, error ~ a few ulp)
// instead of R-1 table loads: the FP64 pipe has head-room, the load/store unit does not.
template <int R>
__device__ __forceinline__ void twiddle_powers(cplx w, cplx (&pw)[16]) {
    pw[1] = w;
    if (R > 2) {
        pw[2] = cmul(w, w);
        pw[3] = cmul(pw[2], w);
    }
    if (R > 4) {
        pw[4] = cmul(pw[2], pw[2]);
        pw[5] = cmul(pw[4], w);
        pw[6] = cmul(pw[4], pw[2]);
        pw[7] = cmul(pw[4], pw[3]);
    }
    if (R > 8) {
        pw[8] = cmul(pw[4], pw[4]);
#pragma unroll
        for (int k = 9; k < 16; k++) pw[k] = cmul(pw[8], pw[k - 8]);
    }
}

// R-point DFTs on consecutive groups of a 16-register array: group j uses v[j*R .. j*R+R-1]
template <int R>
__device__ __forceinline__ void dft_groups(cplx (&v)[16]);
template <>
__device__ __forceinline__ void dft_groups<16>(cplx (&v)[16]) {
    dft16(v);
}
template <>
__device__ __forceinline__ void dft_groups<8>(cplx (&v)[16]) {
    dft8(v[0], v[1], v[2], v[3], v[4], v[5], v[6], v[7]);
    dft8(v[8], v[9], v[10], v[11], v[12], v[13], v[14], v[15]);
}
template <>
__device__ __forceinline__ void dft_groups<4>(cplx (&v)[16]) {
#pragma unroll
    for (int j = 0; j < 4; j++) dft4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
}
template <>
__device__ __forceinline__ void dft_groups<2>(cplx (&v)[16]) {
#pragma unroll
    for (int j = 0; j < 8; j++) dft2(v[2 * j], v[2 * j + 1]);
}

template <int N>
struct FftPlan {
    static_assert(N >= 16 && N <= 4096 && (N & (N - 1)) == 0, "power-of-two length 16..4096");
    static constexpr int M      = N / 16;                  // threads (slots) per pencil
    static constexpr int R1     = 16;
    static constexpr int R2     = (M >= 16) ? 16 : M;      // 1 when N == 16
    static constexpr int R3     = N / (16 * R2);           // 1 when N <= 256
    static constexpr int PASSES = (R2 == 1) ? 1 : ((R3 == 1) ? 2 : 3);
    static constexpr int RLAST  = (PASSES == 1) ? 16 : ((PASSES == 2) ? R2 : R3);
};

constexpr int ilog2(int v) { return v <= 1 ? 0 : 1 + ilog2(v / 2); }
template <int N>
struct FftLog2 {
    static constexpr int value = ilog2(N);
};

// Shared-memory image of the T pencils a CTA transforms together.  Thread index = slot*T + pencil,
// 128-bit accesses are served per quarter-warp (8 lanes), conflict-free iff the 8 sixteen-byte
// slots differ mod 8.  T >= 8: the 8 lanes are 8 pencils of one slot, an odd pencil stride is
// enough.  T = 4 or 2: a quarter-warp mixes 2 or 4 slots; pencil stride N+2 (N+4) plus an XOR of
// the low three index bits with bits [log2(last radix) ..] makes every exchange pattern of every
// pass conflict-free (searched exhaustively by tools/bank_sim.py).
template <int N, int T>
struct FftSmem {
    static constexpr int PSTRIDE = N + (T >= 8 ? 1 : (T == 4 ? 2 : 4));
    static constexpr int SHIFT   = (T >= 8) ? -1 : ilog2(FftPlan<N>::RLAST);
    __device__ static __forceinline__ int at(int a) { return SHIFT < 0 ? a : (a ^ ((a >> (SHIFT < 0 ? 0 : SHIFT)) & 7)); }
};

// One middle/last pass: radix R, K = product of earlier radices, L = N/(K*R).
// Reads its inputs from the pencil's shared-memory image S (written by the previous pass),
// leaves the DFT outputs (twiddled unless LAST) in v[j*R + k] for butterfly j = 0..16/R-1.
// W^(base*k), k = 1..R-1: either the multiplication tree or, for kernels that are FP64-bound and have
// load/store head-room (the contiguous-row generation kernel), R-1 loads from the W_N table.
template <int R, bool TWLOAD>
__device__ __forceinline__ void twiddle_set(const cplx *__restrict__ tw, int base, cplx (&pw)[16]) {
    if constexpr (TWLOAD) {
#pragma unroll
        for (int k = 1; k < R; k++) pw[k] = __ldg(&tw[base * k]);
    } else {
        twiddle_powers<R>(__ldg(&tw[base]), pw);
    }
}

template <int N, int T, int R, int K, bool LAST, bool TWLOAD = false>
__device__ __forceinline__ void fft_pass(cplx (&v)[16], cplx *S, int b, const cplx *__restrict__ tw) {
    constexpr int M  = N / 16;
    constexpr int L  = N / (K * R);
    constexpr int NB = 16 / R;  // butterflies per thread
#pragma unroll
    for (int j = 0; j < NB; j++) {
        const int q  = b + j * M;
        const int kk = q / L, l = q % L;
#pragma unroll
        for (int n = 0; n < R; n++) v[j * R + n] = S[FftSmem<N, T>::at(kk * (R * L) + n * L + l)];
    }
    dft_groups<R>(v);
    if constexpr (!LAST) {
        __syncthreads();  // everyone has read S before it is overwritten
#pragma unroll
        for (int j = 0; j < NB; j++) {
            const int q  = b + j * M;
            const int kk = q / L, l = q % L;
            cplx pw[16];
            twiddle_set<R, TWLOAD>(tw, K * l, pw);
#pragma unroll
            for (int k = 0; k < R; k++) {
                cplx x = v[j * R + k];
                if (k > 0) x = cmul(x, pw[k]);
                S[FftSmem<N, T>::at((kk + K * k) * L + l)] = x;
            }
        }
        __syncthreads();
    }
}

// In-register FFT of one pencil.  v[e] holds x[b + M*e] on entry and X[bo + M*e] on exit, where
// bo is the returned output slot (bo == b except for 3-pass lengths, where it is a permutation of
// the slots — callers use it for every index derived from the transformed data).
// Every thread of the CTA must call this (it contains __syncthreads()).  S = this pencil's
// shared-memory image (FftSmem<N,T>::PSTRIDE elements), tw = W_N^j table.
//
// 3-pass lengths (N >= 512, radices 16, 16, R3): only the first exchange is a transpose across
// the whole pencil (CTA barrier).  Pass 2 reads 16 locations and, after its DFT and twiddle, writes
// its results back to the very same locations; pass 3 then only needs what the R3 neighbouring
// slots of the same warp wrote, so a __syncwarp() replaces the two CTA barriers of the second
// exchange and the warps of a CTA drift apart (FP64 and shared-memory phases overlap).
// WL = false keeps the natural output order (bo == b) at the price of two more CTA barriers: the
// contiguous-row kernel needs it, because the slot permutation would break its coalesced row stores.
template <int N, int T, bool WL = true, bool TWLOAD = false>
__device__ __forceinline__ int fft_pencil(cplx (&v)[16], cplx *S, int b, const cplx *__restrict__ tw) {
    typedef FftPlan<N> P;
    typedef FftSmem<N, T> SM;
    int bo = b;
    // pass 1: radix 16 over stride M, K = 1, L = M
    dft16(v);
    if constexpr (P::PASSES > 1) {
        {
            cplx pw[16];
            twiddle_set<16, TWLOAD>(tw, b, pw);
#pragma unroll
            for (int k = 0; k < 16; k++) {
                cplx x = v[k];
                if (k > 0) x = cmul(x, pw[k]);
                S[SM::at(k * P::M + b)] = x;
            }
        }
        __syncthreads();
        if constexpr (P::PASSES == 2) {
            fft_pass<N, T, P::R2, 16, true>(v, S, b, tw);
        } else if constexpr (!WL) {
            fft_pass<N, T, P::R2, 16, false, TWLOAD>(v, S, b, tw);
            fft_pass<N, T, P::R3, 16 * P::R2, true>(v, S, b, tw);
        } else {
            constexpr int R3 = P::R3;
            static_assert(P::R2 == 16 && R3 * T <= 32 && 32 % (R3 * T) == 0, "pass-3 partners must share a warp");
            const int k1 = b / R3, i = b % R3;
            const int base = k1 * (16 * R3);
            // pass 2: radix 16 over stride R3, in place on this thread's own 16 locations
#pragma unroll
            for (int n = 0; n < 16; n++) v[n] = S[SM::at(base + n * R3 + i)];
            dft16(v);
            {
                cplx pw[16];
                twiddle_powers<16>(__ldg(&tw[16 * i]), pw);
#pragma unroll
                for (int k = 0; k < 16; k++) {
                    cplx x = v[k];
                    if (k > 0) x = cmul(x, pw[k]);
                    S[SM::at(base + k * R3 + i)] = x;
                }
            }
            __syncwarp();
            // pass 3: radix R3 butterflies k2 = i + R3*j of this slot group
            constexpr int NB3 = 16 / R3;
#pragma unroll
            for (int j = 0; j < NB3; j++)
#pragma unroll
                for (int n = 0; n < R3; n++) v[j * R3 + n] = S[SM::at(base + (i + R3 * j) * R3 + n)];
            dft_groups<R3>(v);
            bo = k1 + 16 * i;
        }
        // last pass, radix R, butterfly j: v[j*R + k] = X[bo + M*(j + (16/R)*k)]  -> reorder to slot order
        constexpr int RL = (P::PASSES == 2) ? P::R2 : P::R3;
        constexpr int NB = 16 / RL;
        if constexpr (NB > 1) {
            cplx t[16];
#pragma unroll
            for (int j = 0; j < NB; j++)
#pragma unroll
                for (int k = 0; k < RL; k++) t[j + NB * k] = v[j * RL + k];
#pragma unroll
            for (int e = 0; e < 16; e++) v[e] = t[e];
        }
    }  // PASSES > 1
    return bo;
}

}  // namespace zplt
