// Specialised separable-DWT kernels for the headline path (compile-time filter length).
//
// One launch per level, fused 2-pass tile kernel, HBM-bound by design:
//   forward : column (y) analysis in REGISTERS while walking down the image (each thread owns 4
//             adjacent columns = one 128-bit load per row, a sliding window of F rows, no horizontal
//             neighbour needed), results to a small double-buffered shared-memory ring, then the row
//             (x) analysis from shared memory (conflict-free 128-bit reads) and 64-bit coalesced stores.
//   inverse : column synthesis in registers (polyphase, F/2 taps per output, window of F/2(+1) band
//             rows for the four bands), ring in shared memory, row synthesis from shared memory,
//             128-bit stores.
// Shared-memory traffic is ~10 B per pixel instead of ~18 B for a "stage the input tile" design and
// the input is touched exactly once by 128-bit loads.  Periodic wrap / odd sizes are handled by the
// same kernels: threads whose 4 columns are not interior-and-aligned gather them one by one.
// Intermediate approximations (read by the next level's launch) are stored with an L2 evict_last
// policy, coefficient bands and the final image with evict_first, so the ping-pong plane of the
// next level is served by the 126 MB L2.
//
// The analysis runs columns-then-rows (the reference does rows-then-columns, separable.cu:196-197);
// the two orders differ only in fp32 rounding (~1e-7 relative, tolerance is 1e-5).  For Haar the
// butterfly order is exactly the reference's (haar.cu:27-35).
#include <stdlib.h>

#include <type_traits>

#include "pwt_internal.h"

namespace {

__device__ __forceinline__ int wrap_dwt(int i, int N) {
    const int Ne = N + (N & 1);
    i %= Ne;
    if (i < 0) i += Ne;
    return i >= N ? N - 1 : i;
}
__device__ __forceinline__ int wrap_per(int i, int N) {
    i %= N;
    return i < 0 ? i + N : i;
}

// single-step versions, valid for -N <= i < 2N (the dispatcher only uses these kernels for N >= 64)
__device__ __forceinline__ int wrap1_dwt(int i, int N) {
    const int Ne = N + (N & 1);
    if (i < 0) i += Ne;
    if (i >= Ne) i -= Ne;
    return i >= N ? N - 1 : i;
}
__device__ __forceinline__ int wrap1_per(int i, int N) {
    if (i < 0) i += N;
    if (i >= N) i -= N;
    return i;
}

// ---- cache-policy helpers ---------------------------------------------------------------------
__device__ __forceinline__ unsigned long long policy_evict_last() {
    unsigned long long p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ unsigned long long policy_evict_first() {
    unsigned long long p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ float4 ldg4(const float* p, unsigned long long pol) {
    float4 v;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.f32 {%0,%1,%2,%3}, [%4], %5;"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                 : "l"(p), "l"(pol));
    return v;
}
__device__ __forceinline__ void stg2(float* p, float a, float b, unsigned long long pol) {
    asm volatile("st.global.L2::cache_hint.v2.f32 [%0], {%1,%2}, %3;" ::"l"(p), "f"(a), "f"(b), "l"(pol) : "memory");
}
__device__ __forceinline__ void stg4(float* p, float4 v, unsigned long long pol) {
    asm volatile("st.global.L2::cache_hint.v4.f32 [%0], {%1,%2,%3,%4}, %5;" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z),
                 "f"(v.w), "l"(pol)
                 : "memory");
}
__device__ __forceinline__ void stg1(float* p, float a, unsigned long long pol) {
    asm volatile("st.global.L2::cache_hint.f32 [%0], %1, %2;" ::"l"(p), "f"(a), "l"(pol) : "memory");
}

__device__ __forceinline__ float2 fma2s(float x, float2 t, float2 acc) { return __ffma2_rn(make_float2(x, x), t, acc); }
__device__ __forceinline__ void fma4(float4& acc, const float4& v, float t) {
    acc.x = fmaf(v.x, t, acc.x);
    acc.y = fmaf(v.y, t, acc.y);
    acc.z = fmaf(v.z, t, acc.z);
    acc.w = fmaf(v.w, t, acc.w);
}

constexpr int FLAG_VEC_IN = 1;    // input rows are 16-byte aligned (Nc % 4 == 0)
constexpr int FLAG_VEC_OUT = 2;   // output rows allow vector stores
constexpr int FLAG_A_KEEP = 4;    // the approximation output feeds another launch: keep it in L2
constexpr int FLAG_IN_LAST = 8;   // the input was produced by the previous launch (L2 resident): last use

// =========================================================================================
// forward
// =========================================================================================
// Border handling stays out of the hot path: a thread whose 4 columns are interior and 16-byte aligned
// (`vec`) and a chunk whose rows do not wrap (`rows_plain`, CTA-uniform) load with plain 128-bit loads
// at `row * Nc + xcol`; only the few border threads / chunks take the gather path.
template <int F, bool HAAR, int NT, int R, int MINB>
__global__ void __launch_bounds__(NT, MINB)
k_fwd(const float* __restrict__ in, float* __restrict__ A, float* __restrict__ Hb, float* __restrict__ V,
      float* __restrict__ D, int Nr, int Nc, int TX, int TYT, long long in_bs, long long out_bs, int flags,
      const __grid_constant__ PwtFilters f, const __grid_constant__ PwtTapsFwd tp) {
    constexpr int C = F / 2 - 1;                  // window start offset (separable.cu:104)
    constexpr int CL = (C + 3) & ~3;              // left halo rounded to the vector width
    constexpr int DELTA = CL - C;
    constexpr int SW = 4 * NT + 4;                // ring row pitch (floats)
    constexpr int NV = (DELTA + F + 2 + 3) / 4;   // float4 reads per item in the row pass
    extern __shared__ __align__(16) float sm[];   // [2 bufs][2 planes][R][SW]

    const int tid = threadIdx.x;
    const int Nr2 = (Nr + 1) >> 1, Nc2 = (Nc + 1) >> 1;
    const int kx0 = blockIdx.x * TX, ky0 = blockIdx.y * TYT;
    in += blockIdx.z * in_bs;
    A += blockIdx.z * out_bs;
    Hb += blockIdx.z * out_bs;
    V += blockIdx.z * out_bs;
    D += blockIdx.z * out_bs;
    const unsigned long long pol_in = policy_evict_first();
    const unsigned long long pol_det = policy_evict_first();
    const unsigned long long pol_a = (flags & FLAG_A_KEEP) ? policy_evict_last() : policy_evict_first();

    const int xcol = 2 * kx0 - CL + 4 * tid;                       // this thread's 4 columns
    const int txe = min(TX, Nc2 - kx0);                            // valid output columns of the tile
    const bool col_active = xcol <= 2 * kx0 + 2 * txe + F / 2 - 2;
    const bool vec = (flags & FLAG_VEC_IN) && xcol >= 0 && xcol + 3 < Nc;
    const float* in_t = in + xcol;
    const int ky_end = min(ky0 + TYT, Nr2);
    int cx[4];
#pragma unroll
    for (int i = 0; i < 4; i++) cx[i] = wrap_dwt(xcol + i, Nc);

    // any row, any thread (border path)
    auto load_any = [&](int grow) -> float4 {
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (col_active) {
            const float* row = in + (long long)wrap1_dwt(grow, Nr) * Nc;
            if (vec) {
                v = ldg4(row + xcol, pol_in);
            } else {
                v.x = __ldg(row + cx[0]);
                v.y = __ldg(row + cx[1]);
                v.z = __ldg(row + cx[2]);
                v.w = __ldg(row + cx[3]);
            }
        }
        return v;
    };

    float4 w[F];        // sliding window: w[j] = input row 2*ky - C + j
    float4 q[2 * R];    // prefetch queue: the 2R new rows of the NEXT chunk are in flight during this one
    {
        const int r0 = 2 * ky0 - C;
        if (vec && r0 >= 0 && r0 + F - 2 + 2 * R <= Nr) {
            const float* p = in_t + (long long)r0 * Nc;
#pragma unroll
            for (int j = 0; j < F - 2; j++) w[j] = ldg4(p + (long long)j * Nc, pol_in);
#pragma unroll
            for (int i = 0; i < 2 * R; i++) q[i] = ldg4(p + (long long)(F - 2 + i) * Nc, pol_in);
        } else {
#pragma unroll
            for (int j = 0; j < F - 2; j++) w[j] = load_any(r0 + j);
#pragma unroll
            for (int i = 0; i < 2 * R; i++) q[i] = load_any(r0 + F - 2 + i);
        }
    }

    // pass 1 of one chunk: consume 2 queued rows per output row, refill the queue, column analysis
    auto pass1 = [&](auto mode, float* s_lo, float* s_hi, int rn0) {
        constexpr int MODE = decltype(mode)::value;     // 0: no next chunk, 1: plain loads, 2: border loads
        const float* pn = in_t + (long long)rn0 * Nc;
#pragma unroll
        for (int i = 0; i < R; i++) {
            w[F - 2] = q[2 * i];
            w[F - 1] = q[2 * i + 1];
            if (MODE == 1) {
                q[2 * i] = ldg4(pn + (long long)(2 * i) * Nc, pol_in);
                q[2 * i + 1] = ldg4(pn + (long long)(2 * i + 1) * Nc, pol_in);
            } else if (MODE == 2) {
                q[2 * i] = load_any(rn0 + 2 * i);
                q[2 * i + 1] = load_any(rn0 + 2 * i + 1);
            }
            float4 lo, hi;
            if (HAAR) {
                lo = make_float4(w[0].x + w[1].x, w[0].y + w[1].y, w[0].z + w[1].z, w[0].w + w[1].w);
                hi = make_float4(w[0].x - w[1].x, w[0].y - w[1].y, w[0].z - w[1].z, w[0].w - w[1].w);
            } else {
                lo = make_float4(0.f, 0.f, 0.f, 0.f);
                hi = lo;
#pragma unroll
                for (int j = 0; j < F; j++) {
                    fma4(lo, w[j], f.L[F - 1 - j]);
                    fma4(hi, w[j], f.H[F - 1 - j]);
                }
            }
            *reinterpret_cast<float4*>(s_lo + i * SW + 4 * tid) = lo;
            *reinterpret_cast<float4*>(s_hi + i * SW + 4 * tid) = hi;
#pragma unroll
            for (int j = 0; j < F - 2; j++) w[j] = w[j + 2];
        }
    };

    const int npairs = (txe + 1) >> 1;
    const bool vec_out = (flags & FLAG_VEC_OUT) != 0;
    int buf = 0;
    for (int kyc = ky0; kyc < ky_end; kyc += R, buf ^= 1) {
        float* s_lo = sm + buf * (2 * R * SW);
        float* s_hi = s_lo + R * SW;
        const int rn0 = 2 * (kyc + R) - C + F - 2;                 // first new row of the next chunk
        if (kyc + R >= ky_end)
            pass1(std::integral_constant<int, 0>{}, s_lo, s_hi, rn0);
        else if (vec && rn0 >= 0 && rn0 + 2 * R <= Nr)
            pass1(std::integral_constant<int, 1>{}, s_lo, s_hi, rn0);
        else
            pass1(std::integral_constant<int, 2>{}, s_lo, s_hi, rn0);
        __syncthreads();
        // ---- pass 2: row analysis from the ring, two output columns per thread ----
        if (tid < npairs) {
            int o = kyc * Nc2 + kx0 + 2 * tid;       // < 2^31: one image holds < 2^31 samples
            const bool pair_ok = vec_out && (kx0 + 2 * tid + 1 < Nc2);
            const bool second = kx0 + 2 * tid + 1 < Nc2;
#pragma unroll
            for (int i = 0; i < R; i++, o += Nc2) {
                if (kyc + i >= Nr2) break;
                float vl[4 * NV], vh[4 * NV];
#pragma unroll
                for (int k = 0; k < NV; k++) {
                    const float4 a = *reinterpret_cast<const float4*>(s_lo + i * SW + 4 * tid + 4 * k);
                    const float4 b = *reinterpret_cast<const float4*>(s_hi + i * SW + 4 * tid + 4 * k);
                    vl[4 * k] = a.x; vl[4 * k + 1] = a.y; vl[4 * k + 2] = a.z; vl[4 * k + 3] = a.w;
                    vh[4 * k] = b.x; vh[4 * k + 1] = b.y; vh[4 * k + 2] = b.z; vh[4 * k + 3] = b.w;
                }
                float a0, a1, v0, v1, h0, h1, d0, d1;
                if (HAAR) {
                    a0 = 0.5f * (vl[DELTA] + vl[DELTA + 1]);     a1 = 0.5f * (vl[DELTA + 2] + vl[DELTA + 3]);
                    v0 = 0.5f * (vl[DELTA] - vl[DELTA + 1]);     v1 = 0.5f * (vl[DELTA + 2] - vl[DELTA + 3]);
                    h0 = 0.5f * (vh[DELTA] + vh[DELTA + 1]);     h1 = 0.5f * (vh[DELTA + 2] + vh[DELTA + 3]);
                    d0 = 0.5f * (vh[DELTA] - vh[DELTA + 1]);     d1 = 0.5f * (vh[DELTA + 2] - vh[DELTA + 3]);
                } else {
                    if (F >= 10) {      // long rows of FMAs: issue them 2-wide (sample x (low-pass, high-pass) tap pair)
                        float2 av0 = make_float2(0.f, 0.f), av1 = av0, hd0 = av0, hd1 = av0;
#pragma unroll
                        for (int j = 0; j < F; j++) {
                            av0 = fma2s(vl[DELTA + j], tp.t[j], av0);     av1 = fma2s(vl[DELTA + 2 + j], tp.t[j], av1);
                            hd0 = fma2s(vh[DELTA + j], tp.t[j], hd0);     hd1 = fma2s(vh[DELTA + 2 + j], tp.t[j], hd1);
                        }
                        a0 = av0.x; v0 = av0.y; a1 = av1.x; v1 = av1.y;
                        h0 = hd0.x; d0 = hd0.y; h1 = hd1.x; d1 = hd1.y;
                    } else {
                    a0 = a1 = v0 = v1 = h0 = h1 = d0 = d1 = 0.f;
#pragma unroll
                    for (int j = 0; j < F; j++) {
                        const float tl = f.L[F - 1 - j], th = f.H[F - 1 - j];
                        a0 = fmaf(vl[DELTA + j], tl, a0);     a1 = fmaf(vl[DELTA + 2 + j], tl, a1);
                        v0 = fmaf(vl[DELTA + j], th, v0);     v1 = fmaf(vl[DELTA + 2 + j], th, v1);
                        h0 = fmaf(vh[DELTA + j], tl, h0);     h1 = fmaf(vh[DELTA + 2 + j], tl, h1);
                        d0 = fmaf(vh[DELTA + j], th, d0);     d1 = fmaf(vh[DELTA + 2 + j], th, d1);
                    }
                    }
                }
                if (pair_ok) {
                    stg2(A + o, a0, a1, pol_a);
                    stg2(Hb + o, h0, h1, pol_det);
                    stg2(V + o, v0, v1, pol_det);
                    stg2(D + o, d0, d1, pol_det);
                } else {
                    stg1(A + o, a0, pol_a);
                    stg1(Hb + o, h0, pol_det);
                    stg1(V + o, v0, pol_det);
                    stg1(D + o, d0, pol_det);
                    if (second) {
                        stg1(A + o + 1, a1, pol_a);
                        stg1(Hb + o + 1, h1, pol_det);
                        stg1(V + o + 1, v1, pol_det);
                        stg1(D + o + 1, d1, pol_det);
                    }
                }
            }
        }
    }
}

// =========================================================================================
// inverse
// =========================================================================================
// NT threads = two roles of NT/2 threads: role 0 synthesises t1 = syn_y(A, H), role 1 synthesises
// t2 = syn_y(V, D) (halves the register footprint of the sliding windows); both roles share the
// row-synthesis pass.
template <int F, bool HAAR, int NT, int R, int MINB>
__global__ void __launch_bounds__(NT, MINB)
k_inv(const float* __restrict__ A, const float* __restrict__ Hb, const float* __restrict__ V,
      const float* __restrict__ D, float* __restrict__ out, int nr, int nc, int Nr_out, int Nc_out, int TXH,
      int TYH, long long in_bs, long long out_bs, int flags, const __grid_constant__ PwtFilters f,
      const __grid_constant__ PwtTapsInv tp) {
    constexpr int NC = NT / 2;                           // column groups (threads per role)
    constexpr int P = F / 2 - 1, HALF = F / 2;
    constexpr int S0 = P >> 1, E0 = P & 1;               // output parity 0: row shift / first tap
    constexpr int S1 = (P + 1) >> 1, E1 = (P + 1) & 1;   // output parity 1
    constexpr int WIN = HALF + (S1 - S0);                // band rows alive per output pair
    constexpr int HL = S1;                               // horizontal reach on both sides
    constexpr int HLr = (HL + 3) & ~3;
    constexpr int SW = 4 * NC + 4;
    constexpr int NV = (4 + 2 * HLr) / 4;
    extern __shared__ __align__(16) float sm[];          // [2 bufs][2 planes][2R][SW]

    const int tid = threadIdx.x;
    const int role = tid / NC, t = tid - role * NC;
    const int x0h = blockIdx.x * TXH, y0h = blockIdx.y * TYH;       // tile origin in band coordinates
    const long long ib = blockIdx.z * in_bs;
    const float* __restrict__ ba = (role ? V : A) + ib;              // low-pass partner of this role
    const float* __restrict__ bd = (role ? D : Hb) + ib;             // high-pass partner
    out += blockIdx.z * out_bs;
    const unsigned long long pol_in = policy_evict_first();
    const unsigned long long pol_out = (flags & FLAG_A_KEEP) ? policy_evict_last() : policy_evict_first();

    const int txe = min(TXH, nc - x0h);
    const int kcol = x0h - HLr + 4 * t;
    const bool col_active = kcol <= x0h + txe - 1 + HL;
    const bool vec = (flags & FLAG_VEC_IN) && kcol >= 0 && kcol + 3 < nc;
    int cx[4];
#pragma unroll
    for (int i = 0; i < 4; i++) cx[i] = wrap_per(kcol + i, nc);

    auto load_row = [&](const float* band, int grow) -> float4 {
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (col_active) {
            const float* row = band + (long long)wrap1_per(grow, nr) * nc;
            if (vec) {
                v = ldg4(row + kcol, pol_in);
            } else {
                v.x = __ldg(row + cx[0]);
                v.y = __ldg(row + cx[1]);
                v.z = __ldg(row + cx[2]);
                v.w = __ldg(row + cx[3]);
            }
        }
        return v;
    };

    // window row j <-> band row q + S0 - (HALF-1) + j for the current q
    float4 wa[WIN], wd[WIN];
    float4 qa[R], qd[R];      // prefetch queue: the R new band rows of the NEXT chunk
    const float* ba_t = ba + kcol;
    const float* bd_t = bd + kcol;
    {
        const int r0 = y0h + S0 - (HALF - 1);
        if (vec && r0 >= 0 && r0 + WIN - 1 + R <= nr) {
#pragma unroll
            for (int j = 0; j < WIN - 1; j++) {
                wa[j] = ldg4(ba_t + (long long)(r0 + j) * nc, pol_in);
                wd[j] = ldg4(bd_t + (long long)(r0 + j) * nc, pol_in);
            }
#pragma unroll
            for (int i = 0; i < R; i++) {
                qa[i] = ldg4(ba_t + (long long)(r0 + WIN - 1 + i) * nc, pol_in);
                qd[i] = ldg4(bd_t + (long long)(r0 + WIN - 1 + i) * nc, pol_in);
            }
        } else {
#pragma unroll
            for (int j = 0; j < WIN - 1; j++) {
                wa[j] = load_row(ba, r0 + j);
                wd[j] = load_row(bd, r0 + j);
            }
#pragma unroll
            for (int i = 0; i < R; i++) {
                qa[i] = load_row(ba, r0 + WIN - 1 + i);
                qd[i] = load_row(bd, r0 + WIN - 1 + i);
            }
        }
    }

    const int ngroups = (txe + 3) >> 2;
    const bool vec_out = (flags & FLAG_VEC_OUT) != 0;
    const int q_end = min(y0h + TYH, nr);

    auto pass1 = [&](auto mode, float* s_mine, int rn0) {
        constexpr int MODE = decltype(mode)::value;     // 0: no next chunk, 1: plain loads, 2: border loads
        const float* pa = ba_t + (long long)rn0 * nc;
        const float* pd = bd_t + (long long)rn0 * nc;
#pragma unroll
        for (int i = 0; i < R; i++) {
            wa[WIN - 1] = qa[i];
            wd[WIN - 1] = qd[i];
            if (MODE == 1) {
                qa[i] = ldg4(pa + (long long)i * nc, pol_in);
                qd[i] = ldg4(pd + (long long)i * nc, pol_in);
            } else if (MODE == 2) {
                qa[i] = load_row(ba, rn0 + i);
                qd[i] = load_row(bd, rn0 + i);
            }
            float4 te, to;
            if (HAAR) {
                te = make_float4(wa[0].x + wd[0].x, wa[0].y + wd[0].y, wa[0].z + wd[0].z, wa[0].w + wd[0].w);
                to = make_float4(wa[0].x - wd[0].x, wa[0].y - wd[0].y, wa[0].z - wd[0].z, wa[0].w - wd[0].w);
            } else {
                te = make_float4(0.f, 0.f, 0.f, 0.f);
                to = te;
#pragma unroll
                for (int jj = 0; jj < HALF; jj++) {
                    const int je = HALF - 1 - jj, jo = HALF - 1 - jj + (S1 - S0);
                    fma4(te, wa[je], f.IL[2 * jj + E0]);
                    fma4(te, wd[je], f.IH[2 * jj + E0]);
                    fma4(to, wa[jo], f.IL[2 * jj + E1]);
                    fma4(to, wd[jo], f.IH[2 * jj + E1]);
                }
            }
            *reinterpret_cast<float4*>(s_mine + (2 * i) * SW + 4 * t) = te;
            *reinterpret_cast<float4*>(s_mine + (2 * i + 1) * SW + 4 * t) = to;
#pragma unroll
            for (int j = 0; j < WIN - 1; j++) {
                wa[j] = wa[j + 1];
                wd[j] = wd[j + 1];
            }
        }
    };

    int buf = 0;
    for (int qc = y0h; qc < q_end; qc += R, buf ^= 1) {
        float* s_mine = sm + buf * (4 * R * SW) + role * (2 * R * SW);
        const int rn0 = qc + R + S1;                                // first new band row of the next chunk
        if (qc + R >= q_end)
            pass1(std::integral_constant<int, 0>{}, s_mine, rn0);
        else if (vec && rn0 + R <= nr)
            pass1(std::integral_constant<int, 1>{}, s_mine, rn0);
        else
            pass1(std::integral_constant<int, 2>{}, s_mine, rn0);
        __syncthreads();
        // ---- row synthesis: 4 band columns -> 8 output columns per item; items = 2R rows x ngroups ----
        const float* s_1 = sm + buf * (4 * R * SW);
        const float* s_2 = s_1 + 2 * R * SW;
        const int rows_here = min(2 * R, min(Nr_out - 2 * qc, 2 * (nr - qc)));
        for (int it = tid; it < rows_here * ngroups; it += NT) {
            const int i = it / ngroups, u = it - i * ngroups;
            float v1[4 * NV], v2[4 * NV];
#pragma unroll
            for (int k = 0; k < NV; k++) {
                const float4 a = *reinterpret_cast<const float4*>(s_1 + i * SW + 4 * u + 4 * k);
                const float4 b = *reinterpret_cast<const float4*>(s_2 + i * SW + 4 * u + 4 * k);
                v1[4 * k] = a.x; v1[4 * k + 1] = a.y; v1[4 * k + 2] = a.z; v1[4 * k + 3] = a.w;
                v2[4 * k] = b.x; v2[4 * k + 1] = b.y; v2[4 * k + 2] = b.z; v2[4 * k + 3] = b.w;
            }
            float o[8];
#pragma unroll
            for (int c = 0; c < 4; c++) {
                if (HAAR) {
                    o[2 * c] = 0.5f * (v1[HLr + c] + v2[HLr + c]);
                    o[2 * c + 1] = 0.5f * (v1[HLr + c] - v2[HLr + c]);
                } else {
                    if (F >= 10) {
                        // 2-wide: the sample at band offset w - S1 times the (even-phase, odd-phase) tap pair;
                        // walking w downwards keeps the reference's summation order (jj ascending)
                        float2 eo = make_float2(0.f, 0.f);
#pragma unroll
                        for (int w = 2 * HL; w >= 0; w--) {
                            eo = fma2s(v1[HLr + c + w - S1], tp.l[w], eo);
                            eo = fma2s(v2[HLr + c + w - S1], tp.h[w], eo);
                        }
                        o[2 * c] = eo.x;
                        o[2 * c + 1] = eo.y;
                    } else {
                    float e = 0.f, od = 0.f;
#pragma unroll
                    for (int jj = 0; jj < HALF; jj++) {
                        e = fmaf(v1[HLr + c + S0 - jj], f.IL[2 * jj + E0], e);
                        e = fmaf(v2[HLr + c + S0 - jj], f.IH[2 * jj + E0], e);
                        od = fmaf(v1[HLr + c + S1 - jj], f.IL[2 * jj + E1], od);
                        od = fmaf(v2[HLr + c + S1 - jj], f.IH[2 * jj + E1], od);
                    }
                    o[2 * c] = e;
                    o[2 * c + 1] = od;
                    }
                }
            }
            const int gy = 2 * qc + i;
            const int gx = 2 * (x0h + 4 * u);
            float* dst = out + (long long)gy * Nc_out + gx;
            if (vec_out && gx + 7 < Nc_out) {
                stg4(dst, make_float4(o[0], o[1], o[2], o[3]), pol_out);
                stg4(dst + 4, make_float4(o[4], o[5], o[6], o[7]), pol_out);
            } else {
#pragma unroll
                for (int c = 0; c < 8; c++)
                    if (gx + c < Nc_out) stg1(dst + c, o[c], pol_out);
            }
        }
    }
}

inline int cdiv(int a, int b) { return (a + b - 1) / b; }

inline int num_sms() { return pwt_sm_count(); }

// Pick the tile height (multiple of `quantum` rows) so that the grid is a whole number of waves when
// possible: total CTAs close below a multiple of (#SM x resident CTAs).
int pick_tile_rows(int rows, int nx, int batch, int quantum, int resident) {
    if (const int v = pwt_tuning().fast_tile_rows) {         // tuning override (multiple of the chunk size)
        if (v > 0) return ((v + quantum - 1) / quantum) * quantum;
    }
    const int slots = num_sms() * resident;
    int best = quantum * 8, best_waste = 1 << 30;
    for (int t = 4; t <= 64; t++) {                 // tile heights 4..64 quanta
        const int th = t * quantum;
        const long long ctas = (long long)cdiv(rows, th) * nx * batch;
        const long long waves = (ctas + slots - 1) / slots;
        // wasted work: idle slots in the last wave + halo rows, in units of quantum-rows
        const long long idle = (waves * slots - ctas) * t;
        const long long waste = idle + ctas * 1;    // ~1 quantum of halo/prologue cost per CTA
        if (waste < best_waste) {
            best_waste = (int)waste;
            best = th;
        }
    }
    return best;
}

template <int F, bool HAAR, int NT, int R, int MINB>
int launch_fwd(const float* in, float* A, float* Hb, float* V, float* D, int batch, int Nr, int Nc,
               long long in_bs, long long out_bs, int flags, const PwtFilters& f, cudaStream_t st) {
    constexpr int C = F / 2 - 1, CL = (C + 3) & ~3;
    constexpr int SW = 4 * NT + 4;
    const int Nr2 = (Nr + 1) / 2, Nc2 = (Nc + 1) / 2;
    int txmax = ((4 * NT - CL - F / 2 + 1) / 2) & ~3;
    const int nx = cdiv(Nc2, txmax);
    const int TX = min(txmax, (cdiv(Nc2, nx) + 3) & ~3);
    const size_t smem = sizeof(float) * 2 * 2 * R * SW;
    static PwtKernelOnce once;
    const int resident = pwt_kernel_once(once, k_fwd<F, HAAR, NT, R, MINB>, NT, smem, smem);
    if (!resident) return 0;
    const int TYT = pick_tile_rows(Nr2, cdiv(Nc2, TX), batch, R, resident);
    dim3 grid(cdiv(Nc2, TX), cdiv(Nr2, TYT), batch);
    k_fwd<F, HAAR, NT, R, MINB><<<grid, NT, smem, st>>>(in, A, Hb, V, D, Nr, Nc, TX, TYT, in_bs, out_bs, flags, f,
                                                  pwt_pack_taps_fwd(f, F));
    return 1;
}

template <int F, bool HAAR, int NT, int R, int MINB>
int launch_inv(const float* A, const float* Hb, const float* V, const float* D, float* out, int batch,
               int nr, int nc, int Nr_out, int Nc_out, long long in_bs, long long out_bs, int flags,
               const PwtFilters& f, cudaStream_t st) {
    constexpr int P = F / 2 - 1, HL = (P + 1) >> 1, HLr = (HL + 3) & ~3;
    constexpr int NC = NT / 2;
    constexpr int SW = 4 * NC + 4;
    int txmax = (4 * NC - HL - HLr) & ~3;
    const int nx = cdiv(nc, txmax);
    const int TXH = min(txmax, (cdiv(nc, nx) + 3) & ~3);
    const size_t smem = sizeof(float) * 2 * 2 * 2 * R * SW;
    static PwtKernelOnce once;
    const int resident = pwt_kernel_once(once, k_inv<F, HAAR, NT, R, MINB>, NT, smem, smem);
    if (!resident) return 0;
    const int TYH = pick_tile_rows(nr, cdiv(nc, TXH), batch, R, resident);
    dim3 grid(cdiv(nc, TXH), cdiv(nr, TYH), batch);
    k_inv<F, HAAR, NT, R, MINB><<<grid, NT, smem, st>>>(A, Hb, V, D, out, nr, nc, Nr_out, Nc_out, TXH, TYH, in_bs,
                                                  out_bs, flags, f, pwt_pack_taps_inv(f, F));
    return 1;
}

}  // namespace

int pwt_fast_dwt_fwd2d(const float* in, float* A, float* Hb, float* V, float* D, int batch, int Nr,
                       int Nc, long long in_bs, long long out_bs, const PwtFilters& f, bool haar,
                       int hint_flags, cudaStream_t st) {
    const int F = haar ? 2 : f.hlen;
    if (Nr < 64 || Nc < 64) return 0;                // tiny levels: the generic kernel handles them
    const int Nc2 = (Nc + 1) / 2;
    int flags = hint_flags & (FLAG_A_KEEP | FLAG_IN_LAST);
    if (Nc % 4 == 0 && in_bs % 4 == 0 && ((uintptr_t)in & 15) == 0) flags |= FLAG_VEC_IN;
    if (Nc2 % 2 == 0 && out_bs % 2 == 0 && (((uintptr_t)A | (uintptr_t)Hb | (uintptr_t)V | (uintptr_t)D) & 7) == 0)
        flags |= FLAG_VEC_OUT;
#define FWD(FF, HH, RR, MB) return launch_fwd<FF, HH, 128, RR, MB>(in, A, Hb, V, D, batch, Nr, Nc, in_bs, out_bs, flags, f, st)
    if (haar) FWD(2, true, 4, 6);
    switch (F) {
        case 2: FWD(2, false, 4, 6);
        case 4:
            FWD(4, false, 4, 5);
        case 6: FWD(6, false, 4, 4);
        case 8: FWD(8, false, 4, 4);
        case 10: FWD(10, false, 2, 4);
        case 12: FWD(12, false, 2, 4);
        case 14: FWD(14, false, 2, 3);
        case 16: FWD(16, false, 2, 3);
        case 18: FWD(18, false, 2, 3);
        case 20: FWD(20, false, 2, 3);
        default: return 0;
    }
#undef FWD
}

int pwt_fast_dwt_inv2d(const float* A, const float* Hb, const float* V, const float* D, float* out,
                       int batch, int nr, int nc, int Nr_out, int Nc_out, long long in_bs,
                       long long out_bs, const PwtFilters& f, bool haar, int hint_flags, cudaStream_t st) {
    const int F = haar ? 2 : f.hlen;
    if (nr < 64 || nc < 64) return 0;
    int flags = hint_flags & (FLAG_A_KEEP | FLAG_IN_LAST);
    if (nc % 4 == 0 && in_bs % 4 == 0 && (((uintptr_t)A | (uintptr_t)Hb | (uintptr_t)V | (uintptr_t)D) & 15) == 0)
        flags |= FLAG_VEC_IN;
    if (Nc_out % 4 == 0 && out_bs % 4 == 0 && ((uintptr_t)out & 15) == 0) flags |= FLAG_VEC_OUT;
#define INV(FF, HH, RR, MB) return launch_inv<FF, HH, 256, RR, MB>(A, Hb, V, D, out, batch, nr, nc, Nr_out, Nc_out, in_bs, out_bs, flags, f, st)
    if (haar) INV(2, true, 4, 3);
    switch (F) {
        case 2: INV(2, false, 4, 3);
        case 4: INV(4, false, 4, 3);
        case 6: INV(6, false, 4, 2);
        case 8: INV(8, false, 4, 2);
        case 10: INV(10, false, 2, 2);
        case 12: INV(12, false, 2, 2);
        case 14: INV(14, false, 2, 1);
        case 16: INV(16, false, 2, 1);
        case 18: INV(18, false, 2, 1);
        case 20: INV(20, false, 2, 1);
        default: return 0;
    }
#undef INV
}
