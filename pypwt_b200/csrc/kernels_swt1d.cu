// Stationary (a trous) transform, batched 1D (every row an independent signal), compile-time filter length.
//
//   forward  A[g], D[g] = sum_j (L, H)[F-1-j] * x[(g + (j - (F/2-1)) * s) mod N]        separable.cu:409-446
//   inverse  x[g] = sum_j IL[F-1-j]/2 * a[(g + (j - F/2) * s) mod N] + sum_j IH[F-1-j]/2 * d[...]   separable.cu:553-590
//   s = 2^(level-1)
//
// The generic fallback reads every tap of every output from global memory (one scalar load per tap through L1:
// 0.7-2 TB/s algorithmic).  Here a CTA owns a 1024-column tile and walks down the rows: each row segment (+ the
// dilated filter reach, periodic wrap resolved per 16-byte group) is staged once with cp.async into a
// double-buffered shared row, a thread produces 4 adjacent outputs, and every tap is one aligned 128-bit
// shared load (s % 4 == 0) or comes from the few aligned vectors that cover the window (s = 1, 2).
// Arithmetic order is the reference's (j ascending, the two synthesis sums kept apart and added last).
#include <stdlib.h>

#include "pwt_internal.h"

namespace {

constexpr int NT = 256;
constexpr int TW = 4 * NT;        // columns per tile
constexpr int RPS = 2;            // rows per stage
constexpr int MAXSLOT = 3;        // 16-byte staging groups per thread and row: reach of the filter <= 2048 columns

__device__ __forceinline__ void cp_async16(float* dst, const float* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ float2 fma2s(float x, float2 t, float2 acc) { return __ffma2_rn(make_float2(x, x), t, acc); }

struct TapsDup {                  // synthesis taps, halved and duplicated: l[j] = (IL[F-1-j]/2, same), h likewise
    float2 l[PWT_MAX_TAPS];
    float2 h[PWT_MAX_TAPS];
};

__host__ __device__ inline int halo_l(int c, int s) { return (c * s + 3) & ~3; }

// SMODE 1, 2: dilation s = SMODE (window streamed from contiguous vectors); SMODE 0: s % 4 == 0 (aligned tap loads)
template <int F, int SMODE>
__global__ void __launch_bounds__(NT, 3)
k_swt1d_fwd(const float* __restrict__ in, float* __restrict__ A, float* __restrict__ D, int rows, int Nc, int s, int QS,
            const __grid_constant__ PwtTapsFwd f) {
    constexpr int C = F / 2 - 1;
    extern __shared__ __align__(16) float sm[];
    const int tid = threadIdx.x;
    const int HL = halo_l(C, s), HR = halo_l(F / 2, s);
    const int NG = (TW + HL + HR) >> 2;                 // 16-byte groups of a staged row
    const int pitch = 4 * NG;
    const int x0 = blockIdx.x * TW;
    const int q0 = blockIdx.y * QS, q1 = min(q0 + QS, rows);
    if (q0 >= q1) return;
    int gcol[MAXSLOT];
#pragma unroll
    for (int k = 0; k < MAXSLOT; k++) {
        int g = (x0 - HL + 4 * (tid + k * NT)) % Nc;
        gcol[k] = g < 0 ? g + Nc : g;
    }
    auto stage = [&](int row, int buf) {
#pragma unroll
        for (int rr = 0; rr < RPS; rr++) {
            const float* src = in + (long long)min(row + rr, rows - 1) * Nc;
            float* dst = sm + (buf * RPS + rr) * pitch;
#pragma unroll
            for (int k = 0; k < MAXSLOT; k++)
                if (tid + k * NT < NG) cp_async16(dst + 4 * (tid + k * NT), src + gcol[k]);
        }
        cp_async_commit();
    };
    const int col = x0 + 4 * tid;
    const float2 zero2 = make_float2(0.f, 0.f);
    pwt_pdl_trigger();                                  // programmatic dependent launch: see pwt_internal.h
    pwt_pdl_wait();
    stage(q0, 0);
    int buf = 0;
    for (int row = q0; row < q1; row += RPS, buf ^= 1) {
        cp_async_wait_all();
        __syncthreads();
        if (row + RPS < q1) stage(row + RPS, buf ^ 1);
#pragma unroll
        for (int rr = 0; rr < RPS; rr++) {
            const float* base = sm + (buf * RPS + rr) * pitch + 4 * tid;       // staged column x0 - HL + 4*tid
            float2 acc[4] = {zero2, zero2, zero2, zero2};
            if (SMODE == 0) {
                const float* p = base + HL - C * s;
#pragma unroll
                for (int j = 0; j < F; j++) {
                    const float4 v = *reinterpret_cast<const float4*>(p + j * s);
                    acc[0] = fma2s(v.x, f.t[j], acc[0]);
                    acc[1] = fma2s(v.y, f.t[j], acc[1]);
                    acc[2] = fma2s(v.z, f.t[j], acc[2]);
                    acc[3] = fma2s(v.w, f.t[j], acc[3]);
                }
            } else {
                constexpr int S = SMODE ? SMODE : 1;
                constexpr int HLs = (C * S + 3) & ~3, DX = HLs - C * S, NV = (DX + 4 + (F - 1) * S + 3) / 4;
#pragma unroll
                for (int k = 0; k < NV; k++) {
                    const float4 v = *reinterpret_cast<const float4*>(base + 4 * k);
                    const float xv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                    for (int ee = 0; ee < 4; ee++)
#pragma unroll
                        for (int e = 0; e < 4; e++) {
                            const int d = 4 * k + ee - DX - e;
                            if (d >= 0 && d % S == 0 && d / S < F) acc[e] = fma2s(xv[ee], f.t[d / S], acc[e]);
                        }
                }
            }
            if (row + rr < q1 && col < Nc) {
                const long long o = (long long)(row + rr) * Nc + col;
                *reinterpret_cast<float4*>(A + o) = make_float4(acc[0].x, acc[1].x, acc[2].x, acc[3].x);
                *reinterpret_cast<float4*>(D + o) = make_float4(acc[0].y, acc[1].y, acc[2].y, acc[3].y);
            }
        }
    }
}

template <int F, int SMODE>
__global__ void __launch_bounds__(NT, 3)
k_swt1d_inv(const float* __restrict__ A, const float* __restrict__ D, float* __restrict__ out, int rows, int Nc, int s,
            int QS, const __grid_constant__ TapsDup f) {
    constexpr int C = F / 2;
    extern __shared__ __align__(16) float sm[];
    const int tid = threadIdx.x;
    const int HL = halo_l(C, s), HR = halo_l(F / 2 - 1, s);
    const int NG = (TW + HL + HR) >> 2;
    const int pitch = 4 * NG;
    const int x0 = blockIdx.x * TW;
    const int q0 = blockIdx.y * QS, q1 = min(q0 + QS, rows);
    if (q0 >= q1) return;
    int gcol[MAXSLOT];
#pragma unroll
    for (int k = 0; k < MAXSLOT; k++) {
        int g = (x0 - HL + 4 * (tid + k * NT)) % Nc;
        gcol[k] = g < 0 ? g + Nc : g;
    }
    auto stage = [&](int row, int buf) {                // buffer layout: [buf][row in stage][plane a | plane d]
#pragma unroll
        for (int rr = 0; rr < RPS; rr++) {
            const long long ro = (long long)min(row + rr, rows - 1) * Nc;
            float* dst = sm + ((buf * RPS + rr) * 2) * pitch;
#pragma unroll
            for (int k = 0; k < MAXSLOT; k++)
                if (tid + k * NT < NG) {
                    cp_async16(dst + 4 * (tid + k * NT), A + ro + gcol[k]);
                    cp_async16(dst + pitch + 4 * (tid + k * NT), D + ro + gcol[k]);
                }
        }
        cp_async_commit();
    };
    const int col = x0 + 4 * tid;
    const float2 zero2 = make_float2(0.f, 0.f);
    pwt_pdl_trigger();                                  // programmatic dependent launch: see pwt_internal.h
    pwt_pdl_wait();
    stage(q0, 0);
    int buf = 0;
    for (int row = q0; row < q1; row += RPS, buf ^= 1) {
        cp_async_wait_all();
        __syncthreads();
        if (row + RPS < q1) stage(row + RPS, buf ^ 1);
#pragma unroll
        for (int rr = 0; rr < RPS; rr++) {
            const float* ba = sm + ((buf * RPS + rr) * 2) * pitch + 4 * tid;
            const float* bd = ba + pitch;
            float4 r;
            if (SMODE == 0) {
                // pairs of adjacent outputs: (r1[0], r1[1]), (r1[2], r1[3]) from a; (r2..) from d
                float2 a01 = zero2, a23 = zero2, d01 = zero2, d23 = zero2;
                const int o = HL - C * s;
#pragma unroll
                for (int j = 0; j < F; j++) {
                    const float4 va = *reinterpret_cast<const float4*>(ba + o + j * s);
                    const float4 vd = *reinterpret_cast<const float4*>(bd + o + j * s);
                    a01 = __ffma2_rn(make_float2(va.x, va.y), f.l[j], a01);
                    a23 = __ffma2_rn(make_float2(va.z, va.w), f.l[j], a23);
                    d01 = __ffma2_rn(make_float2(vd.x, vd.y), f.h[j], d01);
                    d23 = __ffma2_rn(make_float2(vd.z, vd.w), f.h[j], d23);
                }
                r = make_float4(a01.x + d01.x, a01.y + d01.y, a23.x + d23.x, a23.y + d23.y);
            } else {
                constexpr int S = SMODE ? SMODE : 1;
                constexpr int HLs = (C * S + 3) & ~3, DX = HLs - C * S, NV = (DX + 4 + (F - 1) * S + 3) / 4;
                float r1[4] = {0.f, 0.f, 0.f, 0.f}, r2[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                for (int k = 0; k < NV; k++) {
                    const float4 va = *reinterpret_cast<const float4*>(ba + 4 * k);
                    const float4 vd = *reinterpret_cast<const float4*>(bd + 4 * k);
                    const float xa[4] = {va.x, va.y, va.z, va.w}, xd[4] = {vd.x, vd.y, vd.z, vd.w};
#pragma unroll
                    for (int ee = 0; ee < 4; ee++)
#pragma unroll
                        for (int e = 0; e < 4; e++) {
                            const int d = 4 * k + ee - DX - e;
                            if (d >= 0 && d % S == 0 && d / S < F) {
                                r1[e] = fmaf(xa[ee], f.l[d / S].x, r1[e]);
                                r2[e] = fmaf(xd[ee], f.h[d / S].x, r2[e]);
                            }
                        }
                }
                r = make_float4(r1[0] + r2[0], r1[1] + r2[1], r1[2] + r2[2], r1[3] + r2[3]);
            }
            if (row + rr < q1 && col < Nc) *reinterpret_cast<float4*>(out + (long long)(row + rr) * Nc + col) = r;
        }
    }
}

inline int sms() { return pwt_sm_count(); }
int pick_qs(int ntiles, int rows) {
    const int want = (12 * sms() + ntiles - 1) / ntiles;          // ~4 waves of 3 CTAs per SM
    int qs = (rows + want - 1) / want;
    qs = ((qs + RPS - 1) / RPS) * RPS;
    if (qs < 2 * RPS) qs = 2 * RPS;
    while ((rows + qs - 1) / qs > 65535) qs += RPS;
    return qs;
}

template <int F, int SMODE>
int launch_fwd(const float* in, float* A, float* D, int rows, int Nc, int s, const PwtFilters& f, cudaStream_t st) {
    const int NG = (TW + halo_l(F / 2 - 1, s) + halo_l(F / 2, s)) >> 2;
    if (NG > MAXSLOT * NT) return 0;
    const size_t smem = sizeof(float) * 2 * RPS * 4 * (size_t)NG;
    static PwtKernelOnce once;                         // largest staged row this instantiation accepts (NG <= MAXSLOT * NT)
    const size_t smem_max = sizeof(float) * 2 * RPS * 4 * (size_t)(MAXSLOT * NT);
    if (!pwt_kernel_once(once, k_swt1d_fwd<F, SMODE>, NT, smem_max, smem_max)) return 0;
    const int ntiles = (Nc + TW - 1) / TW, QS = pick_qs(ntiles, rows);
    dim3 grid(ntiles, (rows + QS - 1) / QS);
    pwt_launch_pdl(k_swt1d_fwd<F, SMODE>, grid, NT, smem, st, in, A, D, rows, Nc, s, QS, pwt_pack_taps_fwd(f, F));
    return 1;
}
template <int F, int SMODE>
int launch_inv(const float* A, const float* D, float* out, int rows, int Nc, int s, const PwtFilters& f, cudaStream_t st) {
    const int NG = (TW + halo_l(F / 2, s) + halo_l(F / 2 - 1, s)) >> 2;
    if (NG > MAXSLOT * NT) return 0;
    const size_t smem = sizeof(float) * 2 * RPS * 2 * 4 * (size_t)NG;
    static PwtKernelOnce once;                         // largest staged row this instantiation accepts (NG <= MAXSLOT * NT)
    const size_t smem_max = sizeof(float) * 2 * RPS * 2 * 4 * (size_t)(MAXSLOT * NT);
    if (!pwt_kernel_once(once, k_swt1d_inv<F, SMODE>, NT, smem_max, smem_max)) return 0;
    TapsDup t;
    for (int j = 0; j < PWT_MAX_TAPS; j++) {
        const float l = j < F ? 0.5f * f.IL[F - 1 - j] : 0.f, h = j < F ? 0.5f * f.IH[F - 1 - j] : 0.f;
        t.l[j] = make_float2(l, l);
        t.h[j] = make_float2(h, h);
    }
    const int ntiles = (Nc + TW - 1) / TW, QS = pick_qs(ntiles, rows);
    dim3 grid(ntiles, (rows + QS - 1) / QS);
    pwt_launch_pdl(k_swt1d_inv<F, SMODE>, grid, NT, smem, st, A, D, out, rows, Nc, s, QS, t);
    return 1;
}
template <int F>
int fwd_s(const float* in, float* A, float* D, int rows, int Nc, int s, const PwtFilters& f, cudaStream_t st) {
    if (s == 1) return launch_fwd<F, 1>(in, A, D, rows, Nc, s, f, st);
    if (s == 2) return launch_fwd<F, 2>(in, A, D, rows, Nc, s, f, st);
    return launch_fwd<F, 0>(in, A, D, rows, Nc, s, f, st);
}
template <int F>
int inv_s(const float* A, const float* D, float* out, int rows, int Nc, int s, const PwtFilters& f, cudaStream_t st) {
    if (s == 1) return launch_inv<F, 1>(A, D, out, rows, Nc, s, f, st);
    if (s == 2) return launch_inv<F, 2>(A, D, out, rows, Nc, s, f, st);
    return launch_inv<F, 0>(A, D, out, rows, Nc, s, f, st);
}

}  // namespace

#define PWT_SWT1D_CASES(X) X(2) X(4) X(6) X(8) X(10) X(12) X(14) X(16) X(18) X(20) X(22) X(24) X(26) X(28) X(30) X(32) X(34) X(36) X(38) X(40)

// Return 0 when the configuration is not covered (width not a multiple of 4, unaligned planes, filter reach
// beyond the staged row): the caller falls back to the generic kernels.
int pwt_fast_swt_fwd1d(const float* in, float* A, float* D, int rows, int Nc, int level, const PwtFilters& f,
                       cudaStream_t st) {
    if (level < 1 || level > 20 || (Nc & 3) || rows < 1 || ((((uintptr_t)in) | ((uintptr_t)A) | ((uintptr_t)D)) & 15)) return 0;
    const int s = 1 << (level - 1);
    switch (f.hlen) {
#define X(FF) case FF: return fwd_s<FF>(in, A, D, rows, Nc, s, f, st);
        PWT_SWT1D_CASES(X)
#undef X
        default: return 0;
    }
}
int pwt_fast_swt_inv1d(const float* A, const float* D, float* out, int rows, int Nc, int level, const PwtFilters& f,
                       cudaStream_t st) {
    if (level < 1 || level > 20 || (Nc & 3) || rows < 1 || ((((uintptr_t)out) | ((uintptr_t)A) | ((uintptr_t)D)) & 15)) return 0;
    const int s = 1 << (level - 1);
    switch (f.hlen) {
#define X(FF) case FF: return inv_s<FF>(A, D, out, rows, Nc, s, f, st);
        PWT_SWT1D_CASES(X)
#undef X
        default: return 0;
    }
}
