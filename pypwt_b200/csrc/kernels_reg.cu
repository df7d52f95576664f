// Register-resident separable DWT kernels for short filters (F <= 10) on 128-column-aligned images:
// the headline path (haar/db2 ... on 8192^2, 4096^2, 2048^2 ...).
//
// Profiling the shared-memory version (profiles/r01_*) showed the L1TEX/LSU data pipe as the most
// loaded unit (65 % of peak: ~24 B of shared-memory + global traffic per pixel through a 128 B/clk
// pipe) with DRAM at ~70 % of the measured copy bandwidth.  These kernels remove shared memory
// altogether:
//   * one WARP owns a strip of 128 input columns (one 128-bit load per lane per row) and walks down a
//     band of rows; warps are completely independent (no __syncthreads, no shared memory);
//   * the horizontal pass runs on the just-loaded row: the F-2 halo samples come from the two
//     neighbouring lanes through warp shuffles, lanes 0 / 31 fetch theirs with predicated scalar loads
//     (periodic wrap happens there);
//   * the vertical pass is a sliding window held in registers (each lane filters exactly the columns it
//     produced horizontally), 64/128-bit coalesced stores.
// L1 traffic drops to ~10 B/px (load 4 + shuffles 2 + store 4) and the instruction count to ~11/px.
// forward : rows then columns (the reference's order, separable.cu:196-197)
// inverse : rows then columns as well (the reference runs columns first, separable.cu:351-361); the
//           result differs only by fp32 rounding.  Haar uses the reference's exact butterfly order.
#include <stdlib.h>

#include <type_traits>

#include "pwt_internal.h"

namespace {

__device__ __forceinline__ int wrap1_dwt(int i, int N) {   // -N <= i < 2N
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
    if (pol == 0) return __ldg(reinterpret_cast<const float4*>(p));
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.f32 {%0,%1,%2,%3}, [%4], %5;"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                 : "l"(p), "l"(pol));
    return v;
}
__device__ __forceinline__ void stg2(float* p, float a, float b, unsigned long long pol) {
    if (pol == 0) { *reinterpret_cast<float2*>(p) = make_float2(a, b); return; }
    asm volatile("st.global.L2::cache_hint.v2.f32 [%0], {%1,%2}, %3;" ::"l"(p), "f"(a), "f"(b), "l"(pol) : "memory");
}
__device__ __forceinline__ void stg4(float* p, float a, float b, float c, float d, unsigned long long pol) {
    if (pol == 0) { *reinterpret_cast<float4*>(p) = make_float4(a, b, c, d); return; }
    asm volatile("st.global.L2::cache_hint.v4.f32 [%0], {%1,%2,%3,%4}, %5;" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d),
                 "l"(pol)
                 : "memory");
}

constexpr unsigned FULL = 0xffffffffu;
constexpr int kWarps = 4;                 // independent warps per CTA
constexpr int FLAG_OUT_KEEP = PWT_HINT_OUT_FEEDS_NEXT;

__device__ __forceinline__ float comp(const float4& v, int c) {
    return c == 0 ? v.x : (c == 1 ? v.y : (c == 2 ? v.z : v.w));
}

// Horizontal neighbourhood of a lane's 4 samples: ext[HW + c] = sample (4*lane + c), c in [-HW, 4+HW).
// Interior lanes get the halo from lanes +-1 by shuffle, lane 0 / 31 from eL / eR (own scalar loads).
// e[i]: on lane 0 the i-th sample LEFT of the strip, on lane 31 the i-th sample RIGHT of it (other lanes:
// unused).  It is fetched by one branch-free scalar load per row (see the ecol[] set-up in the kernels).
template <int HW>
__device__ __forceinline__ void build_ext(const float4& v, const float* e, int lane, float* ext) {
    const bool first = lane == 0, last = lane == 31;
#pragma unroll
    for (int c = 0; c < 4; c++) ext[HW + c] = comp(v, c);
#pragma unroll
    for (int i = 0; i < HW; i++) {
        const float l = __shfl_up_sync(FULL, comp(v, 4 - HW + i), 1);      // lane-1's last HW samples
        const float r = __shfl_down_sync(FULL, comp(v, i), 1);             // lane+1's first HW samples
        ext[i] = first ? e[i] : l;
        ext[HW + 4 + i] = last ? e[i] : r;
    }
}


// N consecutive input rows starting at row rb: one 128-bit load + HW halo scalars per lane and row.
// The wrap test is uniform and hoisted out of the row loop.
template <int N, int HW>
__device__ __forceinline__ void load_rows_fwd(const float* __restrict__ in, int rb, int Nr, int Nc, int px0,
                                              const int* ecol, unsigned long long pol, float4* v,
                                              float (*e)[HW > 0 ? HW : 1]) {
    if (rb >= 0 && rb + N <= Nr) {
        const float* p = in + (long long)rb * Nc;
#pragma unroll
        for (int i = 0; i < N; i++, p += Nc) {
            v[i] = ldg4(p + px0, pol);
#pragma unroll
            for (int c = 0; c < HW; c++) e[i][c] = __ldg(p + ecol[c]);
        }
    } else {
#pragma unroll
        for (int i = 0; i < N; i++) {
            const float* p = in + (long long)wrap1_dwt(rb + i, Nr) * Nc;
            v[i] = ldg4(p + px0, pol);
#pragma unroll
            for (int c = 0; c < HW; c++) e[i][c] = __ldg(p + ecol[c]);
        }
    }
}

// N consecutive band rows of the four bands
template <int N, int HW>
__device__ __forceinline__ void load_rows_inv(const float* __restrict__ A, const float* __restrict__ Hb,
                                              const float* __restrict__ V, const float* __restrict__ D, int rb,
                                              int nr, int nc, int x0, const int* ecol, unsigned long long pol,
                                              float4 (*b)[4], float (*e)[4][HW > 0 ? HW : 1]) {
    const bool plain = rb >= 0 && rb + N <= nr;
#pragma unroll
    for (int i = 0; i < N; i++) {
        const long long ro = (long long)(plain ? rb + i : wrap1_per(rb + i, nr)) * nc;
        b[i][0] = ldg4(A + ro + x0, pol);
        b[i][1] = ldg4(Hb + ro + x0, pol);
        b[i][2] = ldg4(V + ro + x0, pol);
        b[i][3] = ldg4(D + ro + x0, pol);
#pragma unroll
        for (int c = 0; c < HW; c++) {
            e[i][0][c] = __ldg(A + ro + ecol[c]);
            e[i][1][c] = __ldg(Hb + ro + ecol[c]);
            e[i][2][c] = __ldg(V + ro + ecol[c]);
            e[i][3][c] = __ldg(D + ro + ecol[c]);
        }
    }
}

// =========================================================================================
// forward: 128 input columns per warp -> 64 output columns, TYW output rows per task
// =========================================================================================
template <int F, bool HAAR, int U, int MINB>
__global__ void __launch_bounds__(32 * kWarps, MINB)
k_fwd_reg(const float* __restrict__ in, float* __restrict__ A, float* __restrict__ Hb, float* __restrict__ V,
          float* __restrict__ D, int Nr, int Nc, int TYW, long long in_bs, long long out_bs, int flags,
          const __grid_constant__ PwtFilters f) {
    constexpr int C = F / 2 - 1;          // window start offset = halo on each side
    constexpr int HW = C;
    pwt_pdl_trigger();                    // programmatic dependent launch: see pwt_internal.h
    pwt_pdl_wait();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int strips = Nc >> 7;
    const int task = blockIdx.x * kWarps + warp;
    const int strip = task % strips, band = task / strips;
    const int Nr2 = (Nr + 1) >> 1, Nc2 = Nc >> 1;
    const int ky0 = band * TYW;
    if (ky0 >= Nr2) return;
    const int ky1 = min(ky0 + TYW, Nr2);
    in += blockIdx.y * in_bs;
    const long long ob = blockIdx.y * out_bs;
    const bool nohint = flags & 1024;
    const unsigned long long pol_in = nohint ? 0ull : policy_evict_first();
    const unsigned long long pol_det = nohint ? 0ull : policy_evict_first();
    const unsigned long long pol_a = nohint ? 0ull : ((flags & FLAG_OUT_KEEP) ? policy_evict_last() : policy_evict_first());

    const int px0 = (strip << 7) + 4 * lane;
    const float* in_t = in + px0;
    // halo samples: lane 0 fetches the HW samples left of the strip, lane 31 the HW samples right of it;
    // the other lanes re-read their own first sample so that the extra load needs no branch
    int ecol[HW > 0 ? HW : 1];
#pragma unroll
    for (int i = 0; i < HW; i++)
        ecol[i] = lane == 0 ? wrap1_per((strip << 7) - HW + i, Nc)
                            : (lane == 31 ? wrap1_per((strip << 7) + 128 + i, Nc) : px0);
    int o = ky0 * Nc2 + (strip << 6) + 2 * lane;      // output offset (elements, < 2^31)
    A += ob; Hb += ob; V += ob; D += ob;

    if (HAAR) {
        // haar.cu:27-35: A = .5((a+c)+(b+d)), V = .5((a+c)-(b+d)), H = .5((a-c)+(b-d)), D = .5((a-c)-(b-d))
        for (int ky = ky0; ky < ky1; ky += U) {
            float4 r0[U], r1[U];
#pragma unroll
            for (int u = 0; u < U; u++) {
                const int y0 = 2 * (ky + u), y1 = wrap1_dwt(y0 + 1, Nr);
                const bool ok = ky + u < ky1;
                r0[u] = ok ? ldg4(in_t + (long long)y0 * Nc, pol_in) : make_float4(0.f, 0.f, 0.f, 0.f);
                r1[u] = ok ? ldg4(in_t + (long long)y1 * Nc, pol_in) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int u = 0; u < U; u++, o += Nc2) {
                if (ky + u >= ky1) break;
                const float sx = r0[u].x + r1[u].x, sy = r0[u].y + r1[u].y, sz = r0[u].z + r1[u].z, sw = r0[u].w + r1[u].w;
                const float dx = r0[u].x - r1[u].x, dy = r0[u].y - r1[u].y, dz = r0[u].z - r1[u].z, dw = r0[u].w - r1[u].w;
                stg2(A + o, 0.5f * (sx + sy), 0.5f * (sz + sw), pol_a);
                stg2(V + o, 0.5f * (sx - sy), 0.5f * (sz - sw), pol_det);
                stg2(Hb + o, 0.5f * (dx + dy), 0.5f * (dz + dw), pol_det);
                stg2(D + o, 0.5f * (dx - dy), 0.5f * (dz - dw), pol_det);
            }
        }
        return;
    }

    // horizontal analysis of one input row -> (lo0, lo1, hi0, hi1) of this lane's two output columns
    auto hpass = [&](const float4& v, const float* e) -> float4 {
        float ext[F + 2];
        build_ext<HW>(v, e, lane, ext);
        float lo0 = 0.f, lo1 = 0.f, hi0 = 0.f, hi1 = 0.f;
#pragma unroll
        for (int j = 0; j < F; j++) {
            const float tl = f.L[F - 1 - j], th = f.H[F - 1 - j];
            lo0 = fmaf(ext[j], tl, lo0);
            lo1 = fmaf(ext[j + 2], tl, lo1);
            hi0 = fmaf(ext[j], th, hi0);
            hi1 = fmaf(ext[j + 2], th, hi1);
        }
        return make_float4(lo0, lo1, hi0, hi1);
    };
    // vertical sliding window: hw[j] = horizontally filtered row (2*ky - C + j)
    float4 hw[F];
    {
        constexpr int NPRE = F > 2 ? F - 2 : 1;
        float4 v[NPRE];
        float e[NPRE][HW > 0 ? HW : 1];
        load_rows_fwd<F - 2, HW>(in, 2 * ky0 - C, Nr, Nc, px0, ecol, pol_in, v, e);
#pragma unroll
        for (int j = 0; j < F - 2; j++) hw[j] = hpass(v[j], e[j]);
    }
    for (int ky = ky0; ky < ky1; ky += U) {
        float4 v[2 * U];
        float e[2 * U][HW > 0 ? HW : 1];
        load_rows_fwd<2 * U, HW>(in, 2 * ky - C + F - 2, Nr, Nc, px0, ecol, pol_in, v, e);
#pragma unroll
        for (int u = 0; u < U; u++, o += Nc2) {
            hw[F - 2] = hpass(v[2 * u], e[2 * u]);
            hw[F - 1] = hpass(v[2 * u + 1], e[2 * u + 1]);
            float a0 = 0.f, a1 = 0.f, h0 = 0.f, h1 = 0.f, v0 = 0.f, v1 = 0.f, d0 = 0.f, d1 = 0.f;
#pragma unroll
            for (int j = 0; j < F; j++) {
                const float tl = f.L[F - 1 - j], th = f.H[F - 1 - j];
                a0 = fmaf(hw[j].x, tl, a0);
                a1 = fmaf(hw[j].y, tl, a1);
                h0 = fmaf(hw[j].x, th, h0);     // (Lx, Hy)
                h1 = fmaf(hw[j].y, th, h1);
                v0 = fmaf(hw[j].z, tl, v0);     // (Hx, Ly)
                v1 = fmaf(hw[j].w, tl, v1);
                d0 = fmaf(hw[j].z, th, d0);
                d1 = fmaf(hw[j].w, th, d1);
            }
            if (ky + u < ky1) {
                stg2(A + o, a0, a1, pol_a);
                stg2(Hb + o, h0, h1, pol_det);
                stg2(V + o, v0, v1, pol_det);
                stg2(D + o, d0, d1, pol_det);
            }
#pragma unroll
            for (int j = 0; j < F - 2; j++) hw[j] = hw[j + 2];
        }
    }
}

// =========================================================================================
// inverse: 128 band columns per warp -> 256 output columns, TYW band rows per task
// =========================================================================================
template <int F, bool HAAR, int U, int MINB>
__global__ void __launch_bounds__(32 * kWarps, MINB)
k_inv_reg(const float* __restrict__ A, const float* __restrict__ Hb, const float* __restrict__ V,
          const float* __restrict__ D, float* __restrict__ out, int nr, int nc, int Nr_out, int Nc_out, int TYW,
          long long in_bs, long long out_bs, int flags, const __grid_constant__ PwtFilters f) {
    constexpr int P = F / 2 - 1, HALF = F / 2;
    constexpr int S0 = P >> 1, E0 = P & 1;               // output parity 0: shift / first tap
    constexpr int S1 = (P + 1) >> 1, E1 = (P + 1) & 1;   // output parity 1
    constexpr int WIN = HALF + (S1 - S0);                // band rows alive per output row pair
    constexpr int HW = S1;                               // horizontal halo (band samples) on each side
    pwt_pdl_trigger();
    pwt_pdl_wait();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int strips = nc >> 7;
    const int task = blockIdx.x * kWarps + warp;
    const int strip = task % strips, band = task / strips;
    const int q0 = band * TYW;
    if (q0 >= nr) return;
    const int q1 = min(q0 + TYW, nr);
    const long long ib = blockIdx.y * in_bs;
    A += ib; Hb += ib; V += ib; D += ib;
    out += blockIdx.y * out_bs;
    const bool nohint = flags & 1024;
    const unsigned long long pol_in = nohint ? 0ull : policy_evict_first();
    const unsigned long long pol_out = nohint ? 0ull : ((flags & FLAG_OUT_KEEP) ? policy_evict_last() : policy_evict_first());

    const int x0 = (strip << 7) + 4 * lane;
    int ecol[HW > 0 ? HW : 1];
#pragma unroll
    for (int i = 0; i < HW; i++)
        ecol[i] = lane == 0 ? wrap1_per((strip << 7) - HW + i, nc)
                            : (lane == 31 ? wrap1_per((strip << 7) + 128 + i, nc) : x0);
    float* out_t = out + (strip << 8) + 8 * lane;

    if (HAAR) {
        // haar.cu:41-58 with a=A, b=V, c=H, d=D: (0,0)=.5((a+c)+(b+d)) (0,1)=.5((a+c)-(b+d)) (1,0)=.5((a-c)+(b-d)) (1,1)=.5((a-c)-(b-d))
        for (int q = q0; q < q1; q += U) {
            float4 a[U], h[U], v[U], d[U];
#pragma unroll
            for (int u = 0; u < U; u++) {
                const long long ro = (long long)min(q + u, nr - 1) * nc + x0;
                a[u] = ldg4(A + ro, pol_in);
                h[u] = ldg4(Hb + ro, pol_in);
                v[u] = ldg4(V + ro, pol_in);
                d[u] = ldg4(D + ro, pol_in);
            }
#pragma unroll
            for (int u = 0; u < U; u++) {
                if (q + u >= q1) break;
                float e[8], o[8];
#pragma unroll
                for (int c = 0; c < 4; c++) {
                    const float ac = comp(a[u], c) + comp(h[u], c), bd = comp(v[u], c) + comp(d[u], c);
                    const float am = comp(a[u], c) - comp(h[u], c), bm = comp(v[u], c) - comp(d[u], c);
                    e[2 * c] = 0.5f * (ac + bd);
                    e[2 * c + 1] = 0.5f * (ac - bd);
                    o[2 * c] = 0.5f * (am + bm);
                    o[2 * c + 1] = 0.5f * (am - bm);
                }
                const int gy = 2 * (q + u);
                float* p = out_t + (long long)gy * Nc_out;
                stg4(p, e[0], e[1], e[2], e[3], pol_out);
                stg4(p + 4, e[4], e[5], e[6], e[7], pol_out);
                if (gy + 1 < Nr_out) {
                    stg4(p + Nc_out, o[0], o[1], o[2], o[3], pol_out);
                    stg4(p + Nc_out + 4, o[4], o[5], o[6], o[7], pol_out);
                }
            }
        }
        return;
    }

    struct Row8 { float u1[8], u2[8]; };     // horizontally synthesised band row: u1 = syn_x(A,V), u2 = syn_x(H,D)
    auto hpass = [&](const float4* b, float (*e)[HW > 0 ? HW : 1]) -> Row8 {
        float xa[4 + 2 * HW], xh[4 + 2 * HW], xv[4 + 2 * HW], xd[4 + 2 * HW];
        build_ext<HW>(b[0], e[0], lane, xa);
        build_ext<HW>(b[1], e[1], lane, xh);
        build_ext<HW>(b[2], e[2], lane, xv);
        build_ext<HW>(b[3], e[3], lane, xd);
        Row8 r;
#pragma unroll
        for (int c = 0; c < 4; c++) {
            float e1 = 0.f, o1 = 0.f, e2 = 0.f, o2 = 0.f;
#pragma unroll
            for (int jj = 0; jj < HALF; jj++) {
                const float le = f.IL[2 * jj + E0], he = f.IH[2 * jj + E0];
                const float lo = f.IL[2 * jj + E1], ho = f.IH[2 * jj + E1];
                e1 = fmaf(xa[HW + c + S0 - jj], le, e1);
                e1 = fmaf(xv[HW + c + S0 - jj], he, e1);
                o1 = fmaf(xa[HW + c + S1 - jj], lo, o1);
                o1 = fmaf(xv[HW + c + S1 - jj], ho, o1);
                e2 = fmaf(xh[HW + c + S0 - jj], le, e2);
                e2 = fmaf(xd[HW + c + S0 - jj], he, e2);
                o2 = fmaf(xh[HW + c + S1 - jj], lo, o2);
                o2 = fmaf(xd[HW + c + S1 - jj], ho, o2);
            }
            r.u1[2 * c] = e1;
            r.u1[2 * c + 1] = o1;
            r.u2[2 * c] = e2;
            r.u2[2 * c + 1] = o2;
        }
        return r;
    };

    // vertical window: w[j] <-> band row q + S0 - (HALF-1) + j
    Row8 w[WIN];
    {
        constexpr int NPRE = WIN > 1 ? WIN - 1 : 1;
        float4 b[NPRE][4];
        float e[NPRE][4][HW > 0 ? HW : 1];
        load_rows_inv<WIN - 1, HW>(A, Hb, V, D, q0 + S0 - (HALF - 1), nr, nc, x0, ecol, pol_in, b, e);
#pragma unroll
        for (int j = 0; j < WIN - 1; j++) w[j] = hpass(b[j], e[j]);
    }
    for (int q = q0; q < q1; q += U) {
        float4 b[U][4];
        float e[U][4][HW > 0 ? HW : 1];
        load_rows_inv<U, HW>(A, Hb, V, D, q + S1, nr, nc, x0, ecol, pol_in, b, e);
#pragma unroll
        for (int u = 0; u < U; u++) {
            w[WIN - 1] = hpass(b[u], e[u]);
            float ev[8], od[8];
#pragma unroll
            for (int c = 0; c < 8; c++) {
                float s0 = 0.f, s1 = 0.f;
#pragma unroll
                for (int jj = 0; jj < HALF; jj++) {
                    const int je = HALF - 1 - jj, jo = HALF - 1 - jj + (S1 - S0);
                    s0 = fmaf(w[je].u1[c], f.IL[2 * jj + E0], s0);
                    s0 = fmaf(w[je].u2[c], f.IH[2 * jj + E0], s0);
                    s1 = fmaf(w[jo].u1[c], f.IL[2 * jj + E1], s1);
                    s1 = fmaf(w[jo].u2[c], f.IH[2 * jj + E1], s1);
                }
                ev[c] = s0;
                od[c] = s1;
            }
            const int gy = 2 * (q + u);
            if (q + u < q1) {
                float* p = out_t + (long long)gy * Nc_out;
                stg4(p, ev[0], ev[1], ev[2], ev[3], pol_out);
                stg4(p + 4, ev[4], ev[5], ev[6], ev[7], pol_out);
                if (gy + 1 < Nr_out) {
                    stg4(p + Nc_out, od[0], od[1], od[2], od[3], pol_out);
                    stg4(p + Nc_out + 4, od[4], od[5], od[6], od[7], pol_out);
                }
            }
#pragma unroll
            for (int j = 0; j < WIN - 1; j++) w[j] = w[j + 1];
        }
    }
}

inline int cdiv(int a, int b) { return (a + b - 1) / b; }

template <int F, bool HAAR, int U, int MINB>
int launch_fwd(const float* in, float* A, float* Hb, float* V, float* D, int batch, int Nr, int Nc,
               long long in_bs, long long out_bs, int flags, const PwtFilters& f, cudaStream_t st) {
    const int Nr2 = (Nr + 1) / 2;
    int TYW = pwt_tuning().reg_tile_rows;
    TYW = ((TYW + U - 1) / U) * U;
    const int tasks = (Nc / 128) * cdiv(Nr2, TYW);
    dim3 grid(cdiv(tasks, kWarps), batch);
    pwt_launch_pdl(k_fwd_reg<F, HAAR, U, MINB>, dim3(grid), 32 * kWarps, 0, st, in, A, Hb, V, D, Nr, Nc, TYW, in_bs, out_bs, flags, f);
    return 1;
}

template <int F, bool HAAR, int U, int MINB>
int launch_inv(const float* A, const float* Hb, const float* V, const float* D, float* out, int batch,
               int nr, int nc, int Nr_out, int Nc_out, long long in_bs, long long out_bs, int flags,
               const PwtFilters& f, cudaStream_t st) {
    int TYW = pwt_tuning().reg_tile_rows / 2;
    if (TYW < U) TYW = U;
    TYW = ((TYW + U - 1) / U) * U;
    const int tasks = (nc / 128) * cdiv(nr, TYW);
    dim3 grid(cdiv(tasks, kWarps), batch);
    pwt_launch_pdl(k_inv_reg<F, HAAR, U, MINB>, grid, 32 * kWarps, 0, st, A, Hb, V, D, out, nr, nc, Nr_out, Nc_out, TYW, in_bs,
                   out_bs, flags, f);
    return 1;
}

}  // namespace

// Covered: even filter lengths 2..10, input width a multiple of 128 and 16-byte aligned planes.
int pwt_reg_dwt_fwd2d(const float* in, float* A, float* Hb, float* V, float* D, int batch, int Nr,
                      int Nc, long long in_bs, long long out_bs, const PwtFilters& f, bool haar,
                      int hint_flags, cudaStream_t st) {
    const int F = haar ? 2 : f.hlen;
    if (F > 10 || (F & 1) || Nc % 128 != 0 || Nr < 32 || batch > 65535) return 0;
    if (in_bs % 4 != 0 || out_bs % 2 != 0 || ((uintptr_t)in & 15) != 0 ||
        (((uintptr_t)A | (uintptr_t)Hb | (uintptr_t)V | (uintptr_t)D) & 7) != 0)
        return 0;
    const int flags = hint_flags | (pwt_tuning().use_hints ? 0 : 1024);
#define FWD(FF, HH, UU, MB) return launch_fwd<FF, HH, UU, MB>(in, A, Hb, V, D, batch, Nr, Nc, in_bs, out_bs, flags, f, st)
    if (haar) FWD(2, true, 4, 8);
    switch (F) {
        case 4:
            switch (pwt_tuning().reg_fwd_variant) {
                case 1: FWD(4, false, 2, 8);
                case 3: FWD(4, false, 4, 5);
                default: FWD(4, false, 2, 6);
            }
        case 6: FWD(6, false, 2, 6);
        case 8: FWD(8, false, 2, 5);
        case 10: FWD(10, false, 2, 4);
        default: return 0;
    }
#undef FWD
}

int pwt_reg_dwt_inv2d(const float* A, const float* Hb, const float* V, const float* D, float* out,
                      int batch, int nr, int nc, int Nr_out, int Nc_out, long long in_bs,
                      long long out_bs, const PwtFilters& f, bool haar, int hint_flags, cudaStream_t st) {
    const int F = haar ? 2 : f.hlen;
    if (F > 10 || (F & 1) || nc % 128 != 0 || nr < 16 || Nc_out != 2 * nc || batch > 65535) return 0;
    if (in_bs % 4 != 0 || out_bs % 4 != 0 || ((uintptr_t)out & 15) != 0 ||
        (((uintptr_t)A | (uintptr_t)Hb | (uintptr_t)V | (uintptr_t)D) & 15) != 0)
        return 0;
    const int flags = hint_flags | (pwt_tuning().use_hints ? 0 : 1024);
#define INV(FF, HH, UU, MB) return launch_inv<FF, HH, UU, MB>(A, Hb, V, D, out, batch, nr, nc, Nr_out, Nc_out, in_bs, out_bs, flags, f, st)
    if (haar) INV(2, true, 2, 8);
    switch (F) {
        case 4:
            INV(4, false, 2, 4);
        case 6: INV(6, false, 2, 4);
        case 8: INV(8, false, 1, 3);
        case 10: INV(10, false, 1, 3);
        default: return 0;
    }
#undef INV
}
