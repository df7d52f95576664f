// Streaming strip kernels for the mid-length and long filters (F = 6 .. 40, compile-time), any image size.
//
// ncu on the shared-memory tile kernels (profiles/r01_ncu_long_filters.csv) showed, for sym8 (F = 16) at 8192^2:
// L1/shared data pipe at 71 % (52 B of shared-memory traffic per pixel: every sample is re-read ~5x by the row
// pass and ~4.5x by the column pass), 45 % of all executed instructions in the staging loop (a modulo per tile
// row), a 22 % halo recompute in the row pass, FMA pipe only 48 % busy.  This design removes all three:
//   * a CTA owns a strip of 256 image columns and walks DOWN a segment of rows in chunks of R rows
//     (cp.async staging of chunk c+1 overlaps the arithmetic of chunk c; wrap resolved per 16-byte group);
//   * first pass along the rows from shared memory, one warp per row, 8 pixels per lane: the F+6 sample window
//     is read once per 8 pixels (swizzled layout, conflict-free 128-bit reads);
//   * second pass down the columns in TRANSPOSED form: each thread owns one column of one intermediate plane,
//     reads each intermediate sample exactly once and scatters it into the F/2 output rows it contributes to,
//     held as rotating register accumulators (static indices: the chunk height is the rotation period);
//     finished rows go straight to global memory, coalesced.  No vertical halo is ever recomputed inside a
//     segment (only F-2 rows where a segment starts).
//   Shared-memory traffic drops to ~25 B/px and the instruction stream is ~80 % FFMA2.
// Every multiply-add is issued 2-wide (FFMA2): one sample times a packed pair of taps -- (low-pass, high-pass)
// in the analysis, (even phase, odd phase) in the synthesis.
// forward : rows then columns (the reference's order, separable.cu:196-197), taps in the reference's order.
// inverse : rows then columns (the reference runs columns first, separable.cu:351-361; same sums, the result
//           differs only by fp32 rounding, like kernels_reg.cu).
#include <stdlib.h>

#include "pwt_internal.h"

namespace {

constexpr int NT = 256;       // threads per CTA
constexpr int NWARP = NT / 32;
constexpr int SW = 256;       // image columns per strip (forward: input columns, inverse: output columns)
constexpr int HC = SW / 2;    // half-resolution columns per strip

__device__ __forceinline__ int mod_pos(int i, int n) {
    i %= n;
    return i < 0 ? i + n : i;
}
// reference extension of the analysis (separable.cu:98-131): periodic over the size rounded up to even, the
// extra sample of an odd size repeats the last one
__device__ __forceinline__ int wrap_dwt(int i, int N) {
    const int Ne = N + (N & 1);
    if (i < 0) i += Ne;
    else if (i >= Ne) i -= Ne;
    if ((unsigned)i >= (unsigned)Ne) i = mod_pos(i, Ne);      // tiny images only
    return i >= N ? N - 1 : i;
}
__device__ __forceinline__ int wrap_per(int i, int N) {
    if (i < 0) i += N;
    else if (i >= N) i -= N;
    if ((unsigned)i >= (unsigned)N) i = mod_pos(i, N);
    return i;
}

__device__ __forceinline__ void cp_async16(float* dst, const float* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async4(float* dst, const float* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((unsigned)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ float2 fma2s(float x, float2 t, float2 acc) { return __ffma2_rn(make_float2(x, x), t, acc); }
__device__ __forceinline__ float2 mul2s(float x, float2 t) { return make_float2(x * t.x, x * t.y); }

// 128-bit groups of a staged row are stored at q ^ ((q >> 3) & 1): lanes that read windows 8 floats apart
// (float4 index 2*lane + k) then hit 8 different 16-byte banks per quarter warp
__device__ __forceinline__ int swz2(int q) { return q ^ ((q >> 3) & 1); }
// rows written as 4 consecutive float4 per lane (float4 index 4*lane + k) use q ^ ((q >> 3) & 3)
__device__ __forceinline__ int swz4(int q) { return q ^ ((q >> 3) & 3); }

// keep a precomputed value in its register (stops the compiler from re-deriving it in every chunk)
__device__ __forceinline__ void pin(int& v) { asm volatile("" : "+r"(v)); }
// predicated scalar store at a 32-bit element offset: no branch, one IMAD.WIDE + one STG
__device__ __forceinline__ void stg_if(float* base, unsigned off, float v, unsigned u, unsigned n) {
    asm volatile(
        "{ .reg .pred p; .reg .u64 a;\n"
        "  setp.lt.u32 p, %3, %4;\n"
        "  mad.wide.u32 a, %1, 4, %0;\n"
        "  @p st.global.f32 [a], %2; }\n" ::"l"(base), "r"(off), "f"(v), "r"(u), "r"(n));
}

using TapsFwd = PwtTapsFwd;
using TapsInv = PwtTapsInv;

// chunk height = K rotation periods, K chosen so that the rows of a chunk fill the 8 warps of the row pass
__host__ __device__ constexpr int pick_k(int period, int max_rows) {
    int best = 1, best_eff = 0;
    for (int k = 1; k * period <= max_rows || k == 1; k++) {
        const int r = k * period, eff = 1000 * r / (NWARP * ((r + NWARP - 1) / NWARP));
        if (eff > best_eff + 40) { best_eff = eff; best = k; }
        if (k * period > max_rows) break;
    }
    return best;
}

// ---- forward ------------------------------------------------------------------------------------
template <int F>
struct FwdGeo {
    static constexpr int C = F / 2 - 1;                    // output k reads inputs 2k - C .. 2k - C + F - 1
    static constexpr int CL = (C + 3) & ~3;                // the staged row starts CL (aligned) columns left of 2*kx0
    static constexpr int DX = CL - C;
    static constexpr int NV = (DX + F + 6 + 3) / 4;        // float4 per lane window (8 pixels -> 4 outputs)
    static constexpr int IW4 = 2 * 31 + NV;                // float4 groups of a staged row
    static constexpr int P = 4 * ((IW4 + 1) & ~1);         // raw row pitch (floats)
    static constexpr int K = pick_k(F, 24);
    static constexpr int R = K * F;                        // rows per chunk (F = rotation period of the column pass)
    static constexpr int NBUF = R <= 24 ? 2 : 1;           // staging buffers
    static constexpr int NS = (R * IW4 + NT - 1) / NT;     // 16-byte staging slots per thread and chunk
    static constexpr size_t smem = sizeof(float) * ((size_t)NBUF * R * P + (size_t)R * SW) + sizeof(int) * (size_t)(4 * IW4);
};

// NRM: the norm reduction fused into the pass -- every thread sums |c| and c^2 of the coefficients it stores
// (fp32 within a chunk, fp64 across chunks), the CTA writes one pair of doubles to partials[linear CTA index]
// (plain stores, no atomics, no memset; wt.cu:368-416 computes the same sums with cuBLAS over the stored planes).
// count_a: include the approximation plane (last level only).
template <int F, int MB, bool NRM>
__global__ void __launch_bounds__(NT, MB)
k_strip_fwd(const float* __restrict__ in, float* __restrict__ A, float* __restrict__ Hb, float* __restrict__ V,
            float* __restrict__ D, int Nr, int Nc, long long in_bs, long long out_bs, int QS,
            const __grid_constant__ TapsFwd f, double* __restrict__ partials, int count_a) {
    using G = FwdGeo<F>;
    constexpr int C = G::C, CL = G::CL, DX = G::DX, NV = G::NV, IW4 = G::IW4, P = G::P, R = G::R, NBUF = G::NBUF, NS = G::NS;
    constexpr int HF = F / 2;
    extern __shared__ __align__(16) float sm[];
    float* raw = sm;                                  // [NBUF][R][P]
    float* rp = sm + NBUF * R * P;                    // [R][SW]: low-pass half | high-pass half of every row
    int* colidx = reinterpret_cast<int*>(rp + R * SW);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int Nr2 = (Nr + 1) >> 1, Nc2 = (Nc + 1) >> 1;
    const int kx0 = blockIdx.x * HC;
    const int q0 = blockIdx.y * QS, q1 = min(q0 + QS, Nr2);
    const size_t cta = ((size_t)blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
    if (q0 >= q1) {
        if (NRM && tid == 0) partials[2 * cta] = partials[2 * cta + 1] = 0.0;
        return;
    }
    in += blockIdx.z * in_bs;
    const long long ob = blockIdx.z * out_bs;
    const int i0 = 2 * q0 - C;                        // image row of stream row 0
    const int nrows = 2 * (q1 - q0) + F - 2;          // stream rows this segment consumes
    const int nchunks = (nrows + R - 1) / R;
    const int xs = 2 * kx0 - CL;                      // image column of staged column 0 (multiple of 4)
    const bool vec = (Nc & 3) == 0 && Nc >= 4 * IW4 && (((uintptr_t)in) & 15) == 0 && (in_bs & 3) == 0;

    // staging slots of this thread: fixed (row in chunk, 16-byte group) -> shared offset, image offset (one image
    // holds < 2^31 samples, so offsets inside an image are 32-bit)
    int s_off[NS], s_img[NS];
#pragma unroll
    for (int s = 0; s < NS; s++) {
        const int idx = tid + s * NT, r = idx / IW4, q = idx - r * IW4;
        s_off[s] = r * P + 4 * swz2(q);
        int gc = xs + 4 * q;
        if (gc < 0) gc += Nc;
        else if (gc >= Nc) gc -= Nc;
        s_img[s] = r * Nc + gc;
        if (NS <= 6) { pin(s_off[s]); pin(s_img[s]); }
    }
    if (!vec) {
        for (int j = tid; j < 4 * IW4; j += NT) colidx[j] = wrap_dwt(xs + j, Nc);
        __syncthreads();
    }
    auto stage = [&](int c) {
        float* dst = raw + (NBUF == 2 ? (c & 1) * R * P : 0);
        const int ibase = i0 + c * R;
        if (vec) {
            if (ibase >= 0 && ibase + R <= Nr) {       // interior chunk: no wrap
                const int b = ibase * Nc;
#pragma unroll
                for (int s = 0; s < NS; s++)
                    if (s < NS - 1 || tid + s * NT < R * IW4) cp_async16(dst + s_off[s], in + (unsigned)(b + s_img[s]));
            } else {
#pragma unroll
                for (int s = 0; s < NS; s++) {
                    const int idx = tid + s * NT, r = idx / IW4;
                    if (s < NS - 1 || idx < R * IW4)
                        cp_async16(dst + s_off[s], in + (unsigned)(wrap_dwt(ibase + r, Nr) * Nc + (s_img[s] - r * Nc)));
                }
            }
        } else {
            for (int e = tid; e < R * 4 * IW4; e += NT) {
                const int r = e / (4 * IW4), j = e - r * (4 * IW4);
                const float* row = in + (long long)wrap_dwt(ibase + r, Nr) * Nc;
                cp_async4(dst + r * P + 4 * swz2(j >> 2) + (j & 3), row + colidx[j]);
            }
        }
        cp_async_commit();
    };

    // window offsets of this lane in a staged row (float4 index 2*lane + k, swizzled)
    int w_off[NV];
#pragma unroll
    for (int k = 0; k < NV; k++) { w_off[k] = warp * P + 4 * swz2(2 * lane + k); pin(w_off[k]); }

    // column pass ownership: thread -> (plane, half-resolution column)
    const int pl = tid >> 7, col = tid & (HC - 1);
    const int kx = kx0 + col;
    const unsigned nvalid = kx < Nc2 ? (unsigned)(q1 - q0) : 0u;      // outputs u in [0, nvalid) are stored
    float* o0 = (pl ? V : A) + ob + kx;                // low-pass down the column
    float* o1 = (pl ? D : Hb) + ob + kx;               // high-pass down the column
    int ubase = -HF;                                   // chunk c completes outputs u = ubase + 1 .. ubase + R/2
    int obase = (q0 - HF) * Nc2;
    const float2 zero2 = make_float2(0.f, 0.f);
    float2 acc[HF];
#pragma unroll
    for (int a = 0; a < HF; a++) acc[a] = zero2;
    float s1 = 0.f, s2 = 0.f;                          // norm partial sums of the current chunk
    double d1 = 0.0, d2 = 0.0;
    const bool cnt_x = pl || count_a;                  // plane 0 stores A in .x: counted on the last level only

    pwt_pdl_wait();                                        // everything above is independent of the previous launch
    stage(0);
    for (int c = 0; c < nchunks; c++) {
        if (c == nchunks - 1) pwt_pdl_trigger();           // last chunk: the next launch may start being scheduled
        if (NBUF == 2) {
            if (c + 1 < nchunks) { stage(c + 1); cp_async_wait<1>(); }
            else cp_async_wait<0>();
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();                               // chunk c staged; everybody is done with rp
        const float* rb = raw + (NBUF == 2 ? (c & 1) * R * P : 0);
        // ---- row pass: warp per row, lane -> outputs 4*lane .. 4*lane+3 (both filters) ----
#pragma unroll
        for (int rr = 0; rr < (R + NWARP - 1) / NWARP; rr++) {
            const int r = warp + rr * NWARP;
            if (R % NWARP != 0 && r >= R) break;
            // streamed window: every 128-bit group is consumed as soon as it is loaded (sample i feeds output o with
            // tap j = i - DX - 2o), so the live state is the 4 accumulator pairs; each sum still runs over j ascending
            float2 p[4];
#pragma unroll
            for (int o = 0; o < 4; o++) p[o] = zero2;
#pragma unroll
            for (int k = 0; k < NV; k++) {
                const float4 v = *reinterpret_cast<const float4*>(rb + rr * NWARP * P + w_off[k]);
                const float xv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                for (int e = 0; e < 4; e++)
#pragma unroll
                    for (int o = 0; o < 4; o++) {
                        const int j = 4 * k + e - DX - 2 * o;
                        if (j >= 0 && j < F) p[o] = fma2s(xv[e], f.t[j], p[o]);
                    }
            }
            *reinterpret_cast<float4*>(rp + r * SW + 4 * lane) = make_float4(p[0].x, p[1].x, p[2].x, p[3].x);
            *reinterpret_cast<float4*>(rp + r * SW + HC + 4 * lane) = make_float4(p[0].y, p[1].y, p[2].y, p[3].y);
        }
        __syncthreads();                               // rp complete, raw buffer free
        if (NBUF == 1 && c + 1 < nchunks) stage(c + 1);
        // ---- column pass, transposed form: stream row n = c*R + j feeds outputs u = n/2 - d with tap (n&1) + 2d ----
#pragma unroll
        for (int j = 0; j < R; j++) {
            const float x = rp[j * SW + tid];
#pragma unroll
            for (int d = 0; d < HF; d++) {
                const int a = (((j >> 1) - d) % HF + HF) % HF, m = (j & 1) + 2 * d;
                acc[a] = fma2s(x, f.t[m], m == 0 ? zero2 : acc[a]);
            }
            if (j & 1) {
                // completes output u = ubase + (j+1)/2, accumulator ((j+1)/2) mod F/2
                const int a = ((j + 1) >> 1) % HF, k = (j + 1) >> 1;
                stg_if(o0, (unsigned)(obase + k * Nc2), acc[a].x, (unsigned)(ubase + k), nvalid);
                stg_if(o1, (unsigned)(obase + k * Nc2), acc[a].y, (unsigned)(ubase + k), nvalid);
                if (NRM) {
                    const bool ok = (unsigned)(ubase + k) < nvalid;
                    const float vx = ok && cnt_x ? acc[a].x : 0.f, vy = ok ? acc[a].y : 0.f;
                    s1 += fabsf(vx) + fabsf(vy);
                    s2 = fmaf(vx, vx, s2);
                    s2 = fmaf(vy, vy, s2);
                }
            }
        }
        if (NRM) { d1 += (double)s1; d2 += (double)s2; s1 = s2 = 0.f; }
        ubase += R / 2;
        obase += (R / 2) * Nc2;
    }
    if (NRM) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            d1 += __shfl_xor_sync(0xffffffffu, d1, o);
            d2 += __shfl_xor_sync(0xffffffffu, d2, o);
        }
        __syncthreads();                               // everybody is done with rp: reuse it for the 8 warp sums
        double* red = reinterpret_cast<double*>(rp);
        if (lane == 0) { red[2 * warp] = d1; red[2 * warp + 1] = d2; }
        __syncthreads();
        if (tid == 0) {
            double t1 = 0.0, t2 = 0.0;
#pragma unroll
            for (int w = 0; w < NWARP; w++) { t1 += red[2 * w]; t2 += red[2 * w + 1]; }
            partials[2 * cta] = t1;
            partials[2 * cta + 1] = t2;
        }
    }
}

// ---- inverse ------------------------------------------------------------------------------------
// Synthesis taps of the column pass, packed DENSE: the even output row 2v and the odd output row 2(v-SH)+1 read
// the same band rows v - S1 + w, w = 0 .. F/2-1, so every FFMA2 carries two useful products
// (l[w] = (IL even-phase tap, IL odd-phase tap), h[w] likewise from IH).
struct TapsInvDense {
    float2 l[PWT_MAX_TAPS / 2];
    float2 h[PWT_MAX_TAPS / 2];
};

template <int F>
struct InvGeo {
    static constexpr int Pp = F / 2 - 1, HALF = F / 2;
    static constexpr int S0 = Pp >> 1, E0 = Pp & 1, S1 = (Pp + 1) >> 1, E1 = (Pp + 1) & 1;
    static constexpr int NW = 2 * S1 + 1;                  // window of the row pass: band offsets -S1 .. +S1
    static constexpr int SH = Pp & 1;                      // odd rows lag the even rows by SH row pairs in the column pass
    static constexpr int HLr = (S1 + 3) & ~3;              // aligned column halo of the staged band rows
    static constexpr int DX = HLr - S1;
    static constexpr int NV = (DX + 8 + 2 * S1 + 3) / 4;   // float4 per lane window (8 band columns -> 16 outputs)
    static constexpr int BW4 = 2 * 15 + NV;                // float4 groups of a staged band row
    static constexpr int P = 4 * ((BW4 + 1) & ~1);
    static constexpr int K = pick_k(HALF, 24);
    static constexpr int R = K * HALF;                     // band rows per chunk (HALF = rotation period)
    static constexpr int NBUF = R <= 12 ? 2 : 1;
    static constexpr int NS = (4 * R * BW4 + NT - 1) / NT;
    static constexpr size_t smem = sizeof(float) * ((size_t)NBUF * 4 * R * P + (size_t)R * 2 * SW) + sizeof(int) * (size_t)(4 * BW4);
    // which output phases use window position w (same rule as pwt_pack_taps_inv)
    __host__ __device__ static constexpr bool use_e(int w) { return S0 + S1 - w >= 0 && S0 + S1 - w < HALF; }
    __host__ __device__ static constexpr bool use_o(int w) { return 2 * S1 - w >= 0 && 2 * S1 - w < HALF; }
};

// A soft / hard threshold recorded by the plan and not yet applied to memory (THR = 1 / 2): every thread applies
// it to the 16-byte groups it staged itself, right after its cp.async copies land (common.cu:19 / :63).
struct StripThr {
    float beta;       // detail bands of this level
    float beta_app;   // approximation band (coarsest level of a call with do_threshold_appcoeffs)
    int app;
};
template <int THR>
__device__ __forceinline__ float thr1(float v, float beta) {
    if (THR == 1) return copysignf(fmaxf(fabsf(v) - beta, 0.0f), v);
    return (fabsf(v) - beta > 0.0f) ? v : 0.0f * v;
}

template <int F, int MB, int THR>
__global__ void __launch_bounds__(NT, MB)
k_strip_inv(const float* __restrict__ A, const float* __restrict__ Hb, const float* __restrict__ V,
            const float* __restrict__ D, float* __restrict__ out, int nr, int nc, int Nr_out, int Nc_out,
            long long in_bs, long long out_bs, int QS, const __grid_constant__ TapsInv f,
            const __grid_constant__ TapsInvDense fd, const StripThr thr) {
    using G = InvGeo<F>;
    constexpr int S1 = G::S1, NW = G::NW, SH = G::SH, HALF = G::HALF, HLr = G::HLr, DX = G::DX, NV = G::NV, BW4 = G::BW4,
                  P = G::P, R = G::R, NBUF = G::NBUF, NS = G::NS;
    extern __shared__ __align__(16) float sm[];
    float* raw = sm;                                  // [NBUF][4 bands][R][P]
    float* us = sm + NBUF * 4 * R * P;                // [R][2 planes][SW] row-synthesised planes (swizzled float4 groups)
    int* colidx = reinterpret_cast<int*>(us + R * 2 * SW);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int x0 = blockIdx.x * HC;                   // first band column of the strip
    const int q0 = blockIdx.y * QS, q1 = min(q0 + QS, nr);
    if (q0 >= q1) return;
    const long long ib = blockIdx.z * in_bs;
    out += blockIdx.z * out_bs;
    A += ib; Hb += ib; V += ib; D += ib;
    const int r0 = q0 - S1;                           // band row of stream row 0
    const int nrows = (q1 - q0) + SH + HALF - 1;      // stream rows this segment consumes
    const int nchunks = (nrows + R - 1) / R;
    const int xs = x0 - HLr;
    const bool vec = (nc & 3) == 0 && nc >= 4 * BW4 && (in_bs & 3) == 0 &&
                     ((((uintptr_t)A) | ((uintptr_t)Hb) | ((uintptr_t)V) | ((uintptr_t)D)) & 15) == 0;

    int s_off[NS], s_img[NS];
#pragma unroll
    for (int s = 0; s < NS; s++) {
        const int idx = tid + s * NT, br = idx / BW4, q = idx - br * BW4;     // br = band * R + row
        s_off[s] = br * P + 4 * swz2(q);
        int gc = xs + 4 * q;
        if (gc < 0) gc += nc;
        else if (gc >= nc) gc -= nc;
        s_img[s] = (br % R) * nc + gc;
        if (NS <= 6) { pin(s_off[s]); pin(s_img[s]); }
    }
    if (!vec) {
        for (int j = tid; j < 4 * BW4; j += NT) colidx[j] = wrap_per(xs + j, nc);
        __syncthreads();
    }
    auto stage = [&](int c) {
        float* dst = raw + (NBUF == 2 ? (c & 1) * 4 * R * P : 0);
        const int rbase = r0 + c * R;
        if (vec) {
            const bool interior = rbase >= 0 && rbase + R <= nr;
            const int bofs = rbase * nc;
#pragma unroll
            for (int s = 0; s < NS; s++) {
                const int idx = tid + s * NT, br = idx / BW4;
                if (s < NS - 1 || idx < 4 * R * BW4) {
                    const int b = br / R, r = br - b * R;
                    const float* band = b == 0 ? A : b == 1 ? Hb : b == 2 ? V : D;
                    const int o = interior ? bofs + s_img[s] : wrap_per(rbase + r, nr) * nc + (s_img[s] - r * nc);
                    cp_async16(dst + s_off[s], band + (unsigned)o);
                }
            }
        } else {
            for (int e = tid; e < 4 * R * 4 * BW4; e += NT) {
                const int br = e / (4 * BW4), j = e - br * (4 * BW4);
                const int b = br / R, r = br - b * R;
                const float* row = (b == 0 ? A : b == 1 ? Hb : b == 2 ? V : D) + (long long)wrap_per(rbase + r, nr) * nc;
                cp_async4(dst + br * P + 4 * swz2(j >> 2) + (j & 3), row + colidx[j]);
            }
        }
        cp_async_commit();
    };

    auto apply_thr = [&](int c) {                       // this thread's own staged groups of chunk c
        float* dst = raw + (NBUF == 2 ? (c & 1) * 4 * R * P : 0);
        if (vec) {
#pragma unroll
            for (int s = 0; s < NS; s++) {
                const int idx = tid + s * NT, br = idx / BW4;
                if (s < NS - 1 || idx < 4 * R * BW4) {
                    const int b = br / R;
                    if (b == 0 && !thr.app) continue;
                    const float beta = b == 0 ? thr.beta_app : thr.beta;
                    float4 v = *reinterpret_cast<float4*>(dst + s_off[s]);
                    v.x = thr1<THR ? THR : 1>(v.x, beta); v.y = thr1<THR ? THR : 1>(v.y, beta);
                    v.z = thr1<THR ? THR : 1>(v.z, beta); v.w = thr1<THR ? THR : 1>(v.w, beta);
                    *reinterpret_cast<float4*>(dst + s_off[s]) = v;
                }
            }
        } else {
            for (int e = tid; e < 4 * R * 4 * BW4; e += NT) {
                const int br = e / (4 * BW4), j = e - br * (4 * BW4), b = br / R;
                if (b == 0 && !thr.app) continue;
                float* q = dst + br * P + 4 * swz2(j >> 2) + (j & 3);
                *q = thr1<THR ? THR : 1>(*q, b == 0 ? thr.beta_app : thr.beta);
            }
        }
    };
    // row pass ownership: lanes 0-15 -> plane 0 (A with V), lanes 16-31 -> plane 1 (H with D); 8 band columns each
    const int rpl = lane >> 4, g = lane & 15;
    int w_off[NV];
#pragma unroll
    for (int k = 0; k < NV; k++) { w_off[k] = (rpl * R + warp) * P + 4 * swz2(2 * g + k); if (F <= 24) pin(w_off[k]); }
    int u_off[4];                                     // where this lane writes its 16 outputs in a us row
#pragma unroll
    for (int k = 0; k < 4; k++) { u_off[k] = warp * 2 * SW + rpl * SW + 4 * swz4(4 * g + k); if (F <= 24) pin(u_off[k]); }

    // column pass ownership: thread -> output column.  Stream row n feeds the row pairs u = n - w (relative to q0);
    // pair u = (output row 2(q0+u), output row 2(q0+u-SH)+1).
    const int X = 2 * x0 + tid;
    int cu_off = 4 * swz4(tid >> 2) + (tid & 3);
    pin(cu_off);
    const unsigned nv_e = X < Nc_out ? (unsigned)(q1 - q0) : 0u;
    const unsigned nv_o = X < Nc_out ? (unsigned)max(min(q1, Nr_out >> 1) - q0, 0) : 0u;
    float* op = out + X;
    int ubase = -(HALF - 1);                          // chunk c completes pairs u = ubase .. ubase + R - 1
    int obase = 2 * (q0 - (HALF - 1)) * Nc_out;
    const float2 zero2 = make_float2(0.f, 0.f);
    float2 acc[HALF];
#pragma unroll
    for (int a = 0; a < HALF; a++) acc[a] = zero2;

    pwt_pdl_wait();                                        // everything above is independent of the previous launch
    stage(0);
    for (int c = 0; c < nchunks; c++) {
        if (c == nchunks - 1) pwt_pdl_trigger();           // last chunk: the next launch may start being scheduled
        if (NBUF == 2) {
            if (c + 1 < nchunks) { stage(c + 1); cp_async_wait<1>(); }
            else cp_async_wait<0>();
        } else {
            cp_async_wait<0>();
        }
        if (THR) apply_thr(c);
        __syncthreads();
        const float* rb = raw + (NBUF == 2 ? (c & 1) * 4 * R * P : 0);
        // ---- row synthesis: band row -> two planes of 2*HC samples (low-pass-column plane from A,V; high from H,D) ----
#pragma unroll
        for (int rr = 0; rr < (R + NWARP - 1) / NWARP; rr++) {
            const int r = warp + rr * NWARP;
            if (R % NWARP != 0 && r >= R) break;
            const float* px = rb + rr * NWARP * P;               // + w_off: A or H row r (row low-pass source)
            const float* py = px + 2 * R * P;                    //          V or D row r (row high-pass source)
            // streamed window, walked downwards (the reference's order: jj ascending = window position descending)
            float2 eo[8];
#pragma unroll
            for (int cidx = 0; cidx < 8; cidx++) eo[cidx] = zero2;
#pragma unroll
            for (int k = NV - 1; k >= 0; k--) {
                const float4 a = *reinterpret_cast<const float4*>(px + w_off[k]);
                const float4 b = *reinterpret_cast<const float4*>(py + w_off[k]);
                const float xa[4] = {a.x, a.y, a.z, a.w}, xb[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
                for (int e = 3; e >= 0; e--)
#pragma unroll
                    for (int cidx = 0; cidx < 8; cidx++) {
                        const int w = 4 * k + e - DX - cidx;
                        if (w < 0 || w >= NW) continue;
                        if (G::use_e(w) && G::use_o(w)) {
                            eo[cidx] = fma2s(xa[e], f.l[w], eo[cidx]);
                            eo[cidx] = fma2s(xb[e], f.h[w], eo[cidx]);
                        } else if (G::use_e(w)) {
                            eo[cidx].x = fmaf(xa[e], f.l[w].x, eo[cidx].x);
                            eo[cidx].x = fmaf(xb[e], f.h[w].x, eo[cidx].x);
                        } else if (G::use_o(w)) {
                            eo[cidx].y = fmaf(xa[e], f.l[w].y, eo[cidx].y);
                            eo[cidx].y = fmaf(xb[e], f.h[w].y, eo[cidx].y);
                        }
                    }
            }
            float o[16];
#pragma unroll
            for (int cidx = 0; cidx < 8; cidx++) { o[2 * cidx] = eo[cidx].x; o[2 * cidx + 1] = eo[cidx].y; }
            float* ur = us + rr * NWARP * 2 * SW;
#pragma unroll
            for (int k = 0; k < 4; k++)
                *reinterpret_cast<float4*>(ur + u_off[k]) = make_float4(o[4 * k], o[4 * k + 1], o[4 * k + 2], o[4 * k + 3]);
        }
        __syncthreads();
        if (NBUF == 1 && c + 1 < nchunks) stage(c + 1);
        // ---- column synthesis, transposed form ----
#pragma unroll
        for (int j = 0; j < R; j++) {
            const float ul = us[j * 2 * SW + cu_off];
            const float uh = us[j * 2 * SW + SW + cu_off];
#pragma unroll
            for (int w = 0; w < HALF; w++) {
                const int a = ((j - w) % HALF + HALF) % HALF;
                acc[a] = fma2s(ul, fd.l[w], w == 0 ? zero2 : acc[a]);
                acc[a] = fma2s(uh, fd.h[w], acc[a]);
            }
            {
                // completes pair u = ubase + j, accumulator (j+1) mod HALF
                const int a = (j + 1) % HALF;
                const int u = ubase + j;
                stg_if(op, (unsigned)(obase + 2 * j * Nc_out), acc[a].x, (unsigned)u, nv_e);
                stg_if(op, (unsigned)(obase + (2 * (j - SH) + 1) * Nc_out), acc[a].y, (unsigned)(u - SH), nv_o);
            }
        }
        ubase += R;
        obase += 2 * R * Nc_out;
    }
}

// ---- batched 1D (ndim = 1: every row of the input is an independent signal) ------------------------------
// The row passes of the 2D kernels on their own: a CTA walks down a strip of 256 columns in chunks of 16 rows
// (cp.async double buffer), one warp per row, results go straight to global memory (the synthesis regroups its
// 16 outputs per lane through a warp-private shared row so that the stores are contiguous 512-byte runs).
// Reference: separable.cu:91-131 (w_kern_forward_pass1 on its own), 210-255 (inverse).
constexpr int R1 = 16;

template <int F>
struct Fwd1Geo {
    static constexpr int C = FwdGeo<F>::C, CL = FwdGeo<F>::CL, DX = FwdGeo<F>::DX, NV = FwdGeo<F>::NV, IW4 = FwdGeo<F>::IW4,
                         P = FwdGeo<F>::P;
    static constexpr int NS = (R1 * IW4 + NT - 1) / NT;
    static constexpr size_t smem = sizeof(float) * ((size_t)2 * R1 * P) + sizeof(int) * (size_t)(4 * IW4);
};

template <int F>
__global__ void __launch_bounds__(NT, 3)
k_strip_fwd1d(const float* __restrict__ in, float* __restrict__ A, float* __restrict__ D, int rows, int Nc, int QS,
              const __grid_constant__ TapsFwd f) {
    using G = Fwd1Geo<F>;
    constexpr int CL = G::CL, DX = G::DX, NV = G::NV, IW4 = G::IW4, P = G::P, NS = G::NS, R = R1;
    extern __shared__ __align__(16) float sm[];
    float* raw = sm;                                  // [2][R][P]
    int* colidx = reinterpret_cast<int*>(sm + 2 * R * P);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int Nc2 = (Nc + 1) >> 1;
    const int kx0 = blockIdx.x * HC;
    const int q0 = blockIdx.y * QS, q1 = min(q0 + QS, rows);
    if (q0 >= q1) return;
    const int nchunks = (q1 - q0 + R - 1) / R;
    const int xs = 2 * kx0 - CL;
    const bool vec = (Nc & 3) == 0 && Nc >= 4 * IW4 && (((uintptr_t)in) & 15) == 0;
    int s_off[NS], s_col[NS];
#pragma unroll
    for (int s = 0; s < NS; s++) {
        const int idx = tid + s * NT, r = idx / IW4, q = idx - r * IW4;
        s_off[s] = r * P + 4 * swz2(q);
        int gc = xs + 4 * q;
        if (gc < 0) gc += Nc;
        else if (gc >= Nc) gc -= Nc;
        s_col[s] = gc;
        if (NS <= 6) { pin(s_off[s]); pin(s_col[s]); }
    }
    if (!vec) {
        for (int j = tid; j < 4 * IW4; j += NT) colidx[j] = wrap_dwt(xs + j, Nc);
        __syncthreads();
    }
    auto stage = [&](int c) {
        float* dst = raw + (c & 1) * R * P;
        const int ibase = q0 + c * R;
        const int rmax = rows - 1 - ibase;            // rows past the end re-read the last one (never stored)
        const float* src = in + (long long)ibase * Nc;
        if (vec) {
#pragma unroll
            for (int s = 0; s < NS; s++) {
                const int idx = tid + s * NT, r = idx / IW4;
                if (s < NS - 1 || idx < R * IW4) cp_async16(dst + s_off[s], src + (unsigned)(min(r, rmax) * Nc + s_col[s]));
            }
        } else {
            for (int e = tid; e < R * 4 * IW4; e += NT) {
                const int r = e / (4 * IW4), j = e - r * (4 * IW4);
                cp_async4(dst + r * P + 4 * swz2(j >> 2) + (j & 3), src + (unsigned)(min(r, rmax) * Nc + colidx[j]));
            }
        }
        cp_async_commit();
    };
    int w_off[NV];
#pragma unroll
    for (int k = 0; k < NV; k++) { w_off[k] = warp * P + 4 * swz2(2 * lane + k); pin(w_off[k]); }
    const int kx = kx0 + 4 * lane;
    const bool vst = (Nc2 & 3) == 0 && kx + 3 < Nc2 && ((((uintptr_t)A) | ((uintptr_t)D)) & 15) == 0;
    const float2 zero2 = make_float2(0.f, 0.f);

    pwt_pdl_wait();
    stage(0);
    for (int c = 0; c < nchunks; c++) {
        if (c == nchunks - 1) pwt_pdl_trigger();
        cp_async_wait<0>();
        __syncthreads();                               // chunk c staged; the other buffer is free
        if (c + 1 < nchunks) stage(c + 1);
        const float* rb = raw + (c & 1) * R * P;
#pragma unroll
        for (int rr = 0; rr < R / NWARP; rr++) {
            const int row = q0 + c * R + warp + rr * NWARP;
            float2 p[4];
#pragma unroll
            for (int o = 0; o < 4; o++) p[o] = zero2;
#pragma unroll
            for (int k = 0; k < NV; k++) {
                const float4 v = *reinterpret_cast<const float4*>(rb + rr * NWARP * P + w_off[k]);
                const float xv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                for (int e = 0; e < 4; e++)
#pragma unroll
                    for (int o = 0; o < 4; o++) {
                        const int j = 4 * k + e - DX - 2 * o;
                        if (j >= 0 && j < F) p[o] = fma2s(xv[e], f.t[j], p[o]);
                    }
            }
            if (row < q1) {
                const long long ofs = (long long)row * Nc2 + kx;
                if (vst) {
                    *reinterpret_cast<float4*>(A + ofs) = make_float4(p[0].x, p[1].x, p[2].x, p[3].x);
                    *reinterpret_cast<float4*>(D + ofs) = make_float4(p[0].y, p[1].y, p[2].y, p[3].y);
                } else {
#pragma unroll
                    for (int o = 0; o < 4; o++)
                        if (kx + o < Nc2) { A[ofs + o] = p[o].x; D[ofs + o] = p[o].y; }
                }
            }
        }
    }
}

template <int F>
struct Inv1Geo {
    using G2 = InvGeo<F>;
    static constexpr int NS = (2 * R1 * G2::BW4 + NT - 1) / NT;
    static constexpr size_t smem = sizeof(float) * ((size_t)2 * 2 * R1 * G2::P + (size_t)NWARP * 2 * SW) + sizeof(int) * (size_t)(4 * G2::BW4);
};

template <int F>
__global__ void __launch_bounds__(NT, 3)
k_strip_inv1d(const float* __restrict__ A, const float* __restrict__ D, float* __restrict__ out, int rows, int nc,
              int Nc_out, int QS, const __grid_constant__ TapsInv f) {
    using G = InvGeo<F>;
    constexpr int NW = G::NW, HLr = G::HLr, DX = G::DX, NV = G::NV, BW4 = G::BW4, P = G::P, R = R1, NS = Inv1Geo<F>::NS;
    extern __shared__ __align__(16) float sm[];
    float* raw = sm;                                  // [2][2 bands][R][P]
    float* us = sm + 2 * 2 * R * P;                   // [NWARP][2 rows][SW] warp-private output rows
    int* colidx = reinterpret_cast<int*>(us + NWARP * 2 * SW);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int x0 = blockIdx.x * HC;
    const int q0 = blockIdx.y * QS, q1 = min(q0 + QS, rows);
    if (q0 >= q1) return;
    const int nchunks = (q1 - q0 + R - 1) / R;
    const int xs = x0 - HLr;
    const bool vec = (nc & 3) == 0 && nc >= 4 * BW4 && ((((uintptr_t)A) | ((uintptr_t)D)) & 15) == 0;
    int s_off[NS], s_col[NS];
#pragma unroll
    for (int s = 0; s < NS; s++) {
        const int idx = tid + s * NT, br = idx / BW4, q = idx - br * BW4;     // br = band * R + row
        s_off[s] = br * P + 4 * swz2(q);
        int gc = xs + 4 * q;
        if (gc < 0) gc += nc;
        else if (gc >= nc) gc -= nc;
        s_col[s] = gc;
        if (NS <= 6) { pin(s_off[s]); pin(s_col[s]); }
    }
    if (!vec) {
        for (int j = tid; j < 4 * BW4; j += NT) colidx[j] = wrap_per(xs + j, nc);
        __syncthreads();
    }
    auto stage = [&](int c) {
        float* dst = raw + (c & 1) * 2 * R * P;
        const int ibase = q0 + c * R;
        const int rmax = rows - 1 - ibase;
        const long long rb = (long long)ibase * nc;
        if (vec) {
#pragma unroll
            for (int s = 0; s < NS; s++) {
                const int idx = tid + s * NT, br = idx / BW4;
                if (s < NS - 1 || idx < 2 * R * BW4) {
                    const int b = br / R, r = br - b * R;
                    cp_async16(dst + s_off[s], (b ? D : A) + rb + (unsigned)(min(r, rmax) * nc + s_col[s]));
                }
            }
        } else {
            for (int e = tid; e < 2 * R * 4 * BW4; e += NT) {
                const int br = e / (4 * BW4), j = e - br * (4 * BW4);
                const int b = br / R, r = br - b * R;
                cp_async4(dst + br * P + 4 * swz2(j >> 2) + (j & 3), (b ? D : A) + rb + (unsigned)(min(r, rmax) * nc + colidx[j]));
            }
        }
        cp_async_commit();
    };
    // lanes 0-15 -> row 2*warp, lanes 16-31 -> row 2*warp+1 of the chunk; 8 band columns (16 outputs) per lane
    const int rpl = lane >> 4, g = lane & 15;
    int w_off[NV];
#pragma unroll
    for (int k = 0; k < NV; k++) { w_off[k] = (2 * warp + rpl) * P + 4 * swz2(2 * g + k); if (F <= 24) pin(w_off[k]); }
    float* usw = us + warp * 2 * SW;
    const bool vst = (Nc_out & 3) == 0 && (((uintptr_t)out) & 15) == 0;
    const float2 zero2 = make_float2(0.f, 0.f);

    pwt_pdl_wait();
    stage(0);
    for (int c = 0; c < nchunks; c++) {
        if (c == nchunks - 1) pwt_pdl_trigger();
        cp_async_wait<0>();
        __syncthreads();
        if (c + 1 < nchunks) stage(c + 1);
        const float* px = raw + (c & 1) * 2 * R * P;   // + w_off: approximation row
        const float* py = px + R * P;                  //          detail row
        float2 eo[8];
#pragma unroll
        for (int cidx = 0; cidx < 8; cidx++) eo[cidx] = zero2;
#pragma unroll
        for (int k = NV - 1; k >= 0; k--) {
            const float4 a = *reinterpret_cast<const float4*>(px + w_off[k]);
            const float4 b = *reinterpret_cast<const float4*>(py + w_off[k]);
            const float xa[4] = {a.x, a.y, a.z, a.w}, xb[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int e = 3; e >= 0; e--)
#pragma unroll
                for (int cidx = 0; cidx < 8; cidx++) {
                    const int w = 4 * k + e - DX - cidx;
                    if (w < 0 || w >= NW) continue;
                    if (G::use_e(w) && G::use_o(w)) {
                        eo[cidx] = fma2s(xa[e], f.l[w], eo[cidx]);
                        eo[cidx] = fma2s(xb[e], f.h[w], eo[cidx]);
                    } else if (G::use_e(w)) {
                        eo[cidx].x = fmaf(xa[e], f.l[w].x, eo[cidx].x);
                        eo[cidx].x = fmaf(xb[e], f.h[w].x, eo[cidx].x);
                    } else if (G::use_o(w)) {
                        eo[cidx].y = fmaf(xa[e], f.l[w].y, eo[cidx].y);
                        eo[cidx].y = fmaf(xb[e], f.h[w].y, eo[cidx].y);
                    }
                }
        }
#pragma unroll
        for (int k = 0; k < 4; k++)
            *reinterpret_cast<float4*>(usw + rpl * SW + 4 * swz4(4 * g + k)) =
                make_float4(eo[2 * k].x, eo[2 * k].y, eo[2 * k + 1].x, eo[2 * k + 1].y);
        __syncwarp();
#pragma unroll
        for (int h = 0; h < 4; h++) {
            const int idx = h * 32 + lane, rw = idx >> 6, q = idx & 63;
            const float4 v = *reinterpret_cast<const float4*>(usw + rw * SW + 4 * swz4(q));
            const int row = q0 + c * R + 2 * warp + rw, col = 2 * x0 + 4 * q;
            if (row < q1) {
                float* dst = out + (long long)row * Nc_out + col;
                if (vst && col + 3 < Nc_out) {
                    *reinterpret_cast<float4*>(dst) = v;
                } else {
                    const float ev[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                    for (int e = 0; e < 4; e++)
                        if (col + e < Nc_out) dst[e] = ev[e];
                }
            }
        }
        __syncwarp();
    }
}

inline int cdiv(int a, int b) { return (a + b - 1) / b; }


// Segment height: the grid is (strips, segments, images); pick the segment count whose last wave is fullest,
// discounting the rows a segment start costs (halo rows + rounding of the stream to whole chunks).
int pick_segments(int nstrips, int batch, int rows_out, int rows_per_out, int halo, int R, int slots) {
    const long long base = (long long)nstrips * batch;
    int best = 1;
    double best_eff = -1.0;
    const int max_seg = rows_out < 512 ? (rows_out + 7) / 8 : 64;
    for (int ns = 1; ns <= max_seg; ns++) {
        const int qs = cdiv(rows_out, ns);
        const int nseg = cdiv(rows_out, qs);
        const double ctas = (double)base * nseg;
        const double waves = ctas / slots;
        const double wave_eff = waves / (double)((long long)((ctas + slots - 1) / slots));
        const int stream = qs * rows_per_out + halo;
        const double chunk_eff = (double)(qs * rows_per_out) / (double)(cdiv(stream, R) * R);
        const double eff = wave_eff * chunk_eff;
        if (eff > best_eff + 1e-9) { best_eff = eff; best = nseg; }
    }
    return best;
}

inline int sm_count() { return pwt_sm_count(); }

struct NormSink {             // optional fused norm reduction of a forward launch
    double* partials;         // receives one (sum |c|, sum c^2) pair per CTA, or null
    int cap;                  // pairs available
    int count_a;              // include the approximation plane
    int written;              // out: pairs written (0: the launch ran without the reduction)
};
template <int F, int MB, bool NRM>
int launch_fwd_mb(const float* in, float* A, float* Hb, float* V, float* D, int batch, int Nr, int Nc,
                  long long in_bs, long long out_bs, const PwtFilters& f, NormSink* ns, cudaStream_t st) {
    using G = FwdGeo<F>;
    static PwtKernelOnce once;
    const int per_sm = pwt_kernel_once(once, k_strip_fwd<F, MB, NRM>, NT, G::smem, G::smem);
    if (!per_sm) return 0;
    const int Nr2 = (Nr + 1) / 2, Nc2 = (Nc + 1) / 2;
    const int nstrips = cdiv(Nc2, HC);
    const int force = pwt_tuning().strip_segs;
    const int nseg = force > 0 ? force : pick_segments(nstrips, batch, Nr2, 2, F - 2, G::R, per_sm * sm_count());
    const int QS = cdiv(Nr2, nseg);
    dim3 grid(nstrips, cdiv(Nr2, QS), batch);
    const TapsFwd t = pwt_pack_taps_fwd(f, F);
    if (NRM) {
        const long long ctas = (long long)grid.x * grid.y * grid.z;
        if (!ns || !ns->partials || ctas > ns->cap) return -1;      // caller retries without the reduction
        pwt_launch_pdl(k_strip_fwd<F, MB, NRM>, grid, NT, G::smem, st, in, A, Hb, V, D, Nr, Nc, in_bs, out_bs, QS, t, ns->partials, ns->count_a);
        ns->written = (int)ctas;
    } else {
        pwt_launch_pdl(k_strip_fwd<F, MB, NRM>, grid, NT, G::smem, st, in, A, Hb, V, D, Nr, Nc, in_bs, out_bs, QS, t, (double*)nullptr, 0);
    }
    return 1;
}
// resident CTAs per SM the kernels are compiled for (register cap 80 / 128)
int occ_fwd(int F) {
    const int o = pwt_tuning().strip_occ_fwd;
    return o ? o : (F >= 14 && F <= 16 ? 3 : 2);
}
int occ_inv(int F) {
    const int o = pwt_tuning().strip_occ_inv;
    return o ? o : (F >= 14 && F <= 16 ? 3 : 2);
}
template <int F>
int launch_fwd(const float* in, float* A, float* Hb, float* V, float* D, int batch, int Nr, int Nc,
               long long in_bs, long long out_bs, const PwtFilters& f, NormSink* ns, cudaStream_t st) {
    if (ns && ns->partials) {                      // 2 CTAs/SM variant only: the sums cost 6 registers
        ns->written = 0;
        if (launch_fwd_mb<F, 2, true>(in, A, Hb, V, D, batch, Nr, Nc, in_bs, out_bs, f, ns, st) == 1) return 1;
    }
    if (F <= 16 && occ_fwd(F) == 3)
        return launch_fwd_mb<F, (F <= 16 ? 3 : 2), false>(in, A, Hb, V, D, batch, Nr, Nc, in_bs, out_bs, f, nullptr, st);
    return launch_fwd_mb<F, 2, false>(in, A, Hb, V, D, batch, Nr, Nc, in_bs, out_bs, f, nullptr, st);
}
template <int F, int MB, int THR>
int launch_inv_mb(const float* A, const float* Hb, const float* V, const float* D, float* out, int batch, int nr,
                  int nc, int Nr_out, int Nc_out, long long in_bs, long long out_bs, const PwtFilters& f,
                  const StripThr& thr, cudaStream_t st) {
    using G = InvGeo<F>;
    static PwtKernelOnce once;
    const int per_sm = pwt_kernel_once(once, k_strip_inv<F, MB, THR>, NT, G::smem, G::smem);
    if (!per_sm) return 0;
    const int nstrips = cdiv(nc, HC);
    const int force = pwt_tuning().strip_segs;
    const int nseg = force > 0 ? force : pick_segments(nstrips, batch, nr, 1, G::SH + G::HALF - 1, G::R, per_sm * sm_count());
    const int QS = cdiv(nr, nseg);
    dim3 grid(nstrips, cdiv(nr, QS), batch);
    const TapsInv t = pwt_pack_taps_inv(f, F);
    TapsInvDense td;
    for (int w = 0; w < PWT_MAX_TAPS / 2; w++) {
        td.l[w] = make_float2(t.l[w].x, t.l[w + G::SH].y);
        td.h[w] = make_float2(t.h[w].x, t.h[w + G::SH].y);
    }
    pwt_launch_pdl(k_strip_inv<F, MB, THR>, grid, NT, G::smem, st, A, Hb, V, D, out, nr, nc, Nr_out, Nc_out, in_bs, out_bs, QS, t, td, thr);
    return 1;
}
template <int F>
int launch_inv(const float* A, const float* Hb, const float* V, const float* D, float* out, int batch, int nr,
               int nc, int Nr_out, int Nc_out, long long in_bs, long long out_bs, const PwtFilters& f, int thr_op,
               const StripThr& thr, cudaStream_t st) {
    // F = 14: 3 CTAs per SM like the plain inverse (64 x 2048^2 db7 fwd+soft+inv 1.184 -> 1.150 ms); F = 16 spills at the
    // 80-register cap with the threshold code and loses (1.23 -> 1.28 ms), so it stays at 2
    if (F == 14 && occ_inv(F) == 3 && pwt_tuning().strip_thr_occ3) {
        if (thr_op == PWT_OP_SOFT) return launch_inv_mb<F, (F == 14 ? 3 : 2), 1>(A, Hb, V, D, out, batch, nr, nc, Nr_out, Nc_out, in_bs, out_bs, f, thr, st);
        if (thr_op == PWT_OP_HARD) return launch_inv_mb<F, (F == 14 ? 3 : 2), 2>(A, Hb, V, D, out, batch, nr, nc, Nr_out, Nc_out, in_bs, out_bs, f, thr, st);
    }
    if (thr_op == PWT_OP_SOFT) return launch_inv_mb<F, 2, 1>(A, Hb, V, D, out, batch, nr, nc, Nr_out, Nc_out, in_bs, out_bs, f, thr, st);
    if (thr_op == PWT_OP_HARD) return launch_inv_mb<F, 2, 2>(A, Hb, V, D, out, batch, nr, nc, Nr_out, Nc_out, in_bs, out_bs, f, thr, st);
    if (F <= 16 && occ_inv(F) == 3)
        return launch_inv_mb<F, (F <= 16 ? 3 : 2), 0>(A, Hb, V, D, out, batch, nr, nc, Nr_out, Nc_out, in_bs, out_bs, f, thr, st);
    return launch_inv_mb<F, 2, 0>(A, Hb, V, D, out, batch, nr, nc, Nr_out, Nc_out, in_bs, out_bs, f, thr, st);
}

}  // namespace

#define PWT_STRIP_CASES(X) X(4) X(6) X(8) X(10) X(12) X(14) X(16) X(18) X(20) X(22) X(24) X(26) X(28) X(30) X(32) X(34) X(36) X(38) X(40)

// partials (may be null): receives one (sum |c|, sum c^2) pair per CTA of the launch for the detail planes (and
// the approximation when count_a); *written = number of pairs (0 when the launch ran without the reduction).
int pwt_strip_dwt_fwd2d_norms(const float* in, float* A, float* Hb, float* V, float* D, int batch, int Nr, int Nc,
                              long long in_bs, long long out_bs, const PwtFilters& f, double* partials, int cap,
                              int count_a, int* written, cudaStream_t st) {
    if (written) *written = 0;
    if (batch > 65535 || Nr < 2 || Nc < 2 || (long long)(Nr + 64) * Nc >= (1LL << 31)) return 0;   // 32-bit offsets inside an image
    NormSink ns;
    ns.partials = partials;
    ns.cap = cap;
    ns.count_a = count_a;
    ns.written = 0;
    int rc = 0;
    switch (f.hlen) {
#define X(FF) case FF: rc = launch_fwd<FF>(in, A, Hb, V, D, batch, Nr, Nc, in_bs, out_bs, f, partials ? &ns : nullptr, st); break;
        PWT_STRIP_CASES(X)
#undef X
        default: return 0;
    }
    if (written) *written = ns.written;
    return rc;
}
int pwt_strip_dwt_fwd2d(const float* in, float* A, float* Hb, float* V, float* D, int batch, int Nr, int Nc,
                        long long in_bs, long long out_bs, const PwtFilters& f, cudaStream_t st) {
    return pwt_strip_dwt_fwd2d_norms(in, A, Hb, V, D, batch, Nr, Nc, in_bs, out_bs, f, nullptr, 0, 0, nullptr, st);
}

// thr_op < 0: nothing pending; PWT_OP_SOFT / PWT_OP_HARD: the recorded threshold is applied to the staged
// coefficients (beta: detail bands of this level; beta_app for the approximation band when app != 0).
int pwt_strip_dwt_inv2d_thr(const float* A, const float* Hb, const float* V, const float* D, float* out, int batch,
                            int nr, int nc, int Nr_out, int Nc_out, long long in_bs, long long out_bs,
                            const PwtFilters& f, int thr_op, float beta, int app, float beta_app, cudaStream_t st) {
    if (batch > 65535 || nr < 1 || nc < 1 || (long long)(Nr_out + 64) * Nc_out >= (1LL << 31)) return 0;
    StripThr thr;
    thr.beta = beta;
    thr.beta_app = beta_app;
    thr.app = app;
    switch (f.hlen) {
#define X(FF) case FF: return launch_inv<FF>(A, Hb, V, D, out, batch, nr, nc, Nr_out, Nc_out, in_bs, out_bs, f, thr_op, thr, st);
        PWT_STRIP_CASES(X)
#undef X
        default: return 0;
    }
}
int pwt_strip_dwt_inv2d(const float* A, const float* Hb, const float* V, const float* D, float* out, int batch,
                        int nr, int nc, int Nr_out, int Nc_out, long long in_bs, long long out_bs,
                        const PwtFilters& f, cudaStream_t st) {
    return pwt_strip_dwt_inv2d_thr(A, Hb, V, D, out, batch, nr, nc, Nr_out, Nc_out, in_bs, out_bs, f, -1, 0.f, 0, 0.f, st);
}

// Haar, batched 1D, widths that are a multiple of 8: no halo and dense rows make the whole stack one flat stream
// (8 samples in -> 4 + 4 out per item).  Same butterfly as the generic kernel / reference (haar.cu:10-42).
namespace {
__global__ void __launch_bounds__(256)
k_haar_fwd1d_flat(const float4* __restrict__ in, float4* __restrict__ A, float4* __restrict__ D, long long items) {
    const float c = 0.70710678118654746f;
    pwt_pdl_trigger();
    pwt_pdl_wait();
    for (long long i = blockIdx.x * 256LL + threadIdx.x; i < items; i += gridDim.x * 256LL) {
        const float4 u = __ldg(in + 2 * i), v = __ldg(in + 2 * i + 1);
        A[i] = make_float4(c * (u.x + u.y), c * (u.z + u.w), c * (v.x + v.y), c * (v.z + v.w));
        D[i] = make_float4(c * (u.x - u.y), c * (u.z - u.w), c * (v.x - v.y), c * (v.z - v.w));
    }
}
__global__ void __launch_bounds__(256)
k_haar_inv1d_flat(const float4* __restrict__ A, const float4* __restrict__ D, float4* __restrict__ out, long long items) {
    const float c = 0.70710678118654746f;
    pwt_pdl_trigger();
    pwt_pdl_wait();
    for (long long i = blockIdx.x * 256LL + threadIdx.x; i < items; i += gridDim.x * 256LL) {
        const float4 a = __ldg(A + i), d = __ldg(D + i);
        out[2 * i] = make_float4(c * (a.x + d.x), c * (a.x - d.x), c * (a.y + d.y), c * (a.y - d.y));
        out[2 * i + 1] = make_float4(c * (a.z + d.z), c * (a.z - d.z), c * (a.w + d.w), c * (a.w - d.w));
    }
}
}  // namespace

int pwt_haar_fwd1d_flat(const float* in, float* A, float* D, int rows, int Nc, cudaStream_t st) {
    if ((Nc & 7) || ((((uintptr_t)in) | ((uintptr_t)A) | ((uintptr_t)D)) & 15)) return 0;
    const long long items = (long long)rows * (Nc / 8);
    const long long want = (items + 255) / 256, cap = (long long)sm_count() * 16;
    pwt_launch_pdl(k_haar_fwd1d_flat, dim3((unsigned)(want < cap ? want : cap)), 256, 0, st, reinterpret_cast<const float4*>(in),
                   reinterpret_cast<float4*>(A), reinterpret_cast<float4*>(D), items);
    return 1;
}
int pwt_haar_inv1d_flat(const float* A, const float* D, float* out, int rows, int nc, int Nc_out, cudaStream_t st) {
    if ((nc & 3) || Nc_out != 2 * nc || ((((uintptr_t)out) | ((uintptr_t)A) | ((uintptr_t)D)) & 15)) return 0;
    const long long items = (long long)rows * (nc / 4);
    const long long want = (items + 255) / 256, cap = (long long)sm_count() * 16;
    pwt_launch_pdl(k_haar_inv1d_flat, dim3((unsigned)(want < cap ? want : cap)), 256, 0, st, reinterpret_cast<const float4*>(A),
                   reinterpret_cast<const float4*>(D), reinterpret_cast<float4*>(out), items);
    return 1;
}

#define PWT_STRIP1D_CASES(X) X(4) X(6) X(8) X(10) X(12) X(14) X(16) X(18) X(20) X(22) X(24) X(26) X(28) X(30) X(32) X(34) X(36) X(38) X(40)

namespace {
int pick_segments_1d(int nstrips, int rows, int slots) {
    // enough CTAs for ~4 waves, segments a multiple of the chunk height
    const int want = (4 * slots + nstrips - 1) / nstrips;
    int qs = (rows + want - 1) / want;
    qs = ((qs + R1 - 1) / R1) * R1;
    if (qs < R1) qs = R1;
    while ((rows + qs - 1) / qs > 65535) qs += R1;
    return qs;
}
template <int F>
int launch_fwd1d(const float* in, float* A, float* D, int rows, int Nc, const PwtFilters& f, cudaStream_t st) {
    using G = Fwd1Geo<F>;
    static PwtKernelOnce once;
    const int per_sm = pwt_kernel_once(once, k_strip_fwd1d<F>, NT, G::smem, G::smem);
    if (!per_sm) return 0;
    const int nstrips = cdiv((Nc + 1) / 2, HC);
    const int QS = pick_segments_1d(nstrips, rows, per_sm * sm_count());
    dim3 grid(nstrips, cdiv(rows, QS), 1);
    pwt_launch_pdl(k_strip_fwd1d<F>, grid, NT, G::smem, st, in, A, D, rows, Nc, QS, pwt_pack_taps_fwd(f, F));
    return 1;
}
template <int F>
int launch_inv1d(const float* A, const float* D, float* out, int rows, int nc, int Nc_out, const PwtFilters& f,
                 cudaStream_t st) {
    using G = Inv1Geo<F>;
    static PwtKernelOnce once;
    const int per_sm = pwt_kernel_once(once, k_strip_inv1d<F>, NT, G::smem, G::smem);
    if (!per_sm) return 0;
    const int nstrips = cdiv(nc, HC);
    const int QS = pick_segments_1d(nstrips, rows, per_sm * sm_count());
    dim3 grid(nstrips, cdiv(rows, QS), 1);
    pwt_launch_pdl(k_strip_inv1d<F>, grid, NT, G::smem, st, A, D, out, rows, nc, Nc_out, QS, pwt_pack_taps_inv(f, F));
    return 1;
}
}  // namespace

int pwt_strip_dwt_fwd1d(const float* in, float* A, float* D, int rows, int Nc, const PwtFilters& f, cudaStream_t st) {
    if (rows < 1 || Nc < 2 || Nc >= (1 << 26)) return 0;
    switch (f.hlen) {
#define X(FF) case FF: return launch_fwd1d<FF>(in, A, D, rows, Nc, f, st);
        PWT_STRIP1D_CASES(X)
#undef X
        default: return 0;
    }
}
int pwt_strip_dwt_inv1d(const float* A, const float* D, float* out, int rows, int nc, int Nc_out, const PwtFilters& f,
                        cudaStream_t st) {
    if (rows < 1 || nc < 1 || nc >= (1 << 26)) return 0;
    switch (f.hlen) {
#define X(FF) case FF: return launch_inv1d<FF>(A, D, out, rows, nc, Nc_out, f, st);
        PWT_STRIP1D_CASES(X)
#undef X
        default: return 0;
    }
}
