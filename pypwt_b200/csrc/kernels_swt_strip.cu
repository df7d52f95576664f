// Stationary (a trous) transform, 2D separable FORWARD level, streaming strip design (F = 2 .. 16).
//
//   out[y][x] = sum_jy sum_jx fy[F-1-jy] * fx[F-1-jx] * in[(y + (jy-c)*s) mod Nr][(x + (jx-c)*s) mod Nc],
//   c = F/2-1, s = 2^(level-1); (fx, fy) = (L,L) -> A, (L,H) -> H, (H,L) -> V, (H,H) -> D      separable.cu:409-493
//
// ncu on the register kernel of kernels_swt.cu (k_swt_fwd, one barrier and one row of prefetch per input row):
// 315 us per 8192^2 level for db4 = 51 % of the DRAM peak, long-scoreboard stalls 3.7 per issue (latency-bound),
// 86 instructions per pixel (scalar FFMA).  This kernel keeps the lattice idea (a CTA walks one residue class of
// rows y = r + q*s down a 256-column strip, so the dilated column filter is an ordinary one) and borrows the
// strip kernels' structure (kernels_strip.cu):
//   * chunks of R lattice rows are staged with cp.async one chunk ahead (double buffer, periodic wrap per 16 bytes);
//   * row pass from shared memory: 64 threads per row, 4 adjacent outputs each, every dilated tap an aligned
//     128-bit load (s % 4 == 0) or taken from the contiguous window (s = 1, 2); FFMA2 with (L, H) tap pairs;
//   * column pass in TRANSPOSED form: a thread owns two adjacent columns of one row-filtered plane, reads each
//     sample once and adds it to the F pending output rows (rotating register accumulators, static indices);
//     every stream row completes one output row of two bands: 64-bit coalesced stores.
// Same summation order as the reference in both directions (taps ascending).
#include <stdlib.h>

#include "pwt_internal.h"

namespace {

constexpr int NT = 256;
constexpr int SWC = 256;          // output columns per strip

__device__ __forceinline__ void cp_async16(float* dst, const float* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ float2 fma2s(float x, float2 t, float2 acc) { return __ffma2_rn(make_float2(x, x), t, acc); }
__device__ __forceinline__ int wrap1(int i, int N) {   // -N <= i < 2N
    if (i < 0) i += N;
    else if (i >= N) i -= N;
    return i;
}
// predicated 64-bit store at a 32-bit element offset
__device__ __forceinline__ void stg2_if(float* base, unsigned off, float a, float b, unsigned u, unsigned n) {
    asm volatile(
        "{ .reg .pred p; .reg .u64 q;\n"
        "  setp.lt.u32 p, %4, %5;\n"
        "  mad.wide.u32 q, %1, 4, %0;\n"
        "  @p st.global.v2.f32 [q], {%2, %3}; }\n" ::"l"(base), "r"(off), "f"(a), "f"(b), "r"(u), "r"(n));
}

__host__ __device__ constexpr int chunk_rows(int F) {      // multiple of F (rotation period) and of 4 (row-pass rounds), >= 8
    int r = F;
    while (r % 4 != 0 || r < 8) r += F;
    return r;
}

template <int F, int SMODE>
__global__ void __launch_bounds__(NT, F <= 8 ? 3 : 2)
k_swt_strip_fwd(const float* __restrict__ in, float* __restrict__ A, float* __restrict__ Hb, float* __restrict__ V,
                float* __restrict__ D, int Nr, int Nc, int s, int TQ, int nseg, long long plane,
                const __grid_constant__ PwtTapsFwd f) {
    constexpr int C = F / 2 - 1, R = chunk_rows(F);
    extern __shared__ __align__(16) float sm[];
    const int tid = threadIdx.x;
    const int HL = (C * s + 3) & ~3, HR = ((F - 1 - C) * s + 3) & ~3;
    const int NG = (SWC + HL + HR) >> 2, pitch = 4 * NG;
    float* raw = sm;                                   // [2][R][pitch]
    float* rp = sm + 2 * R * pitch;                    // [R][2 planes][SWC]
    const int x0 = blockIdx.x * SWC;
    const int r = blockIdx.y / nseg, seg = blockIdx.y - r * nseg;      // residue class of rows, segment of the lattice
    const int nq = (Nr - r + s - 1) / s;
    const int q0 = seg * TQ, q1 = min(q0 + TQ, nq);
    if (q0 >= q1) return;
    in += blockIdx.z * plane;
    const long long ob = blockIdx.z * plane;
    const int m0 = q0 - C, m_last = q1 - 1 + (F - 1 - C);             // lattice indices of the stream
    const int nchunks = (m_last - m0 + R) / R;
    // staging: 4 rows x 64 groups per pass; this thread's (wrapped) image columns
    const int srow = tid >> 6, sg = tid & 63;
    int gcol[2];
#pragma unroll
    for (int k = 0; k < 2; k++) {
        int g = (x0 - HL + 4 * (sg + 64 * k)) % Nc;
        gcol[k] = g < 0 ? g + Nc : g;
    }
    auto stage = [&](int c) {
        float* dst = raw + (c & 1) * R * pitch;
#pragma unroll
        for (int i = 0; i < R / 4; i++) {
            const int rr = srow + 4 * i;
            const int m = min(m0 + c * R + rr, m_last);
            const float* src = in + (unsigned)(wrap1(r + m * s, Nr) * Nc);
#pragma unroll
            for (int k = 0; k < 2; k++)
                if (sg + 64 * k < NG) cp_async16(dst + rr * pitch + 4 * (sg + 64 * k), src + gcol[k]);
        }
        cp_async_commit();
    };
    // column pass ownership: two adjacent columns of one row-filtered plane
    const int pl = tid >> 7, cp = tid & 127;
    const int X = x0 + 2 * cp;
    const unsigned nvalid = X < Nc ? (unsigned)(q1 - q0) : 0u;
    float* o0 = (pl ? V : A) + ob + X;                 // column low-pass
    float* o1 = (pl ? D : Hb) + ob + X;                // column high-pass
    const float2 zero2 = make_float2(0.f, 0.f);
    float2 acc[F][2];
#pragma unroll
    for (int a = 0; a < F; a++) acc[a][0] = acc[a][1] = zero2;
    int ubase = -(F - 1);                              // chunk c completes outputs u = ubase + j, j = 0 .. R-1

    pwt_pdl_wait();                                    // programmatic dependent launch: see pwt_internal.h
    stage(0);
    for (int c = 0; c < nchunks; c++) {
        if (c == nchunks - 1) pwt_pdl_trigger();
        if (c + 1 < nchunks) { stage(c + 1); cp_async_wait<1>(); }
        else cp_async_wait<0>();
        __syncthreads();                               // chunk c staged; everybody is done with rp
        const float* rb = raw + (c & 1) * R * pitch;
        // ---- row pass: 64 threads per row, 4 adjacent outputs (both filters) ----
#pragma unroll
        for (int i = 0; i < R / 4; i++) {
            const int rr = srow + 4 * i;
            const float* base = rb + rr * pitch + 4 * sg;            // staged column x0 - HL + 4*sg
            float2 p[4] = {zero2, zero2, zero2, zero2};
            if (SMODE == 0) {
                const float* q = base + HL - C * s;
#pragma unroll
                for (int j = 0; j < F; j++) {
                    const float4 v = *reinterpret_cast<const float4*>(q + j * s);
                    p[0] = fma2s(v.x, f.t[j], p[0]);
                    p[1] = fma2s(v.y, f.t[j], p[1]);
                    p[2] = fma2s(v.z, f.t[j], p[2]);
                    p[3] = fma2s(v.w, f.t[j], p[3]);
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
                            if (d >= 0 && d % S == 0 && d / S < F) p[e] = fma2s(xv[ee], f.t[d / S], p[e]);
                        }
                }
            }
            *reinterpret_cast<float4*>(rp + rr * 2 * SWC + 4 * sg) = make_float4(p[0].x, p[1].x, p[2].x, p[3].x);
            *reinterpret_cast<float4*>(rp + rr * 2 * SWC + SWC + 4 * sg) = make_float4(p[0].y, p[1].y, p[2].y, p[3].y);
        }
        __syncthreads();                               // rp complete, raw buffer free
        // ---- column pass, transposed form: stream row n = c*R + j feeds outputs u = n - jj with tap jj ----
#pragma unroll
        for (int j = 0; j < R; j++) {
            const float2 x = *reinterpret_cast<const float2*>(rp + j * 2 * SWC + pl * SWC + 2 * cp);
#pragma unroll
            for (int jj = 0; jj < F; jj++) {
                const int a = ((j - jj) % F + F) % F;
                acc[a][0] = fma2s(x.x, f.t[jj], jj == 0 ? zero2 : acc[a][0]);
                acc[a][1] = fma2s(x.y, f.t[jj], jj == 0 ? zero2 : acc[a][1]);
            }
            {
                const int a = (j + 1) % F;             // completed by tap F-1: u = n - F + 1
                const int u = ubase + j;
                const unsigned off = (unsigned)((r + (q0 + u) * s) * Nc);
                stg2_if(o0, off, acc[a][0].x, acc[a][1].x, (unsigned)u, nvalid);
                stg2_if(o1, off, acc[a][0].y, acc[a][1].y, (unsigned)u, nvalid);
            }
        }
        ubase += R;
    }
}

inline int cdiv(int a, int b) { return (a + b - 1) / b; }
inline int sms() { return pwt_sm_count(); }
// segments per residue class: fullest last wave, discounting the F-1 halo rows and the rounding to whole chunks
int pick_segments(long long base, int rows_out, int halo, int R, int slots) {
    int best = 1;
    double best_eff = -1.0;
    const int max_seg = rows_out < 512 ? (rows_out + 7) / 8 : 64;
    for (int ns = 1; ns <= max_seg; ns++) {
        const int qs = cdiv(rows_out, ns), nseg = cdiv(rows_out, qs);
        const double ctas = (double)base * nseg;
        const double waves = ctas / slots;
        const double wave_eff = waves / (double)((long long)((ctas + slots - 1) / slots));
        const double chunk_eff = (double)qs / (double)(cdiv(qs + halo, R) * R);
        if (wave_eff * chunk_eff > best_eff + 1e-9) { best_eff = wave_eff * chunk_eff; best = nseg; }
    }
    return best;
}

template <int F, int SMODE>
int launch(const float* in, float* A, float* Hb, float* V, float* D, int batch, int Nr, int Nc, int s,
           const PwtFilters& f, cudaStream_t st) {
    constexpr int C = F / 2 - 1, R = chunk_rows(F);
    const int HL = (C * s + 3) & ~3, HR = ((F - 1 - C) * s + 3) & ~3;
    const int NG = (SWC + HL + HR) >> 2;
    if (NG > 128) return 0;                            // filter reach beyond the two staging groups per thread
    const size_t smem = sizeof(float) * ((size_t)2 * R * 4 * NG + (size_t)R * 2 * SWC);
    // the staged row pitch depends on the dilation: opt in (once per device) to the largest size this instantiation can
    // ask for; the occupancy depends on the actual size, so it is queried per launch
    static PwtKernelOnce once;
    const size_t smem_max = sizeof(float) * ((size_t)2 * R * 4 * 128 + (size_t)R * 2 * SWC);
    if (!pwt_kernel_once(once, k_swt_strip_fwd<F, SMODE>, NT, smem_max, smem_max)) return 0;
    int per_sm = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_swt_strip_fwd<F, SMODE>, NT, smem);
    if (per_sm <= 0) per_sm = 1;
    const int strips = cdiv(Nc, SWC);
    const int nq = cdiv(Nr, s);                        // longest residue class
    const int nseg = pick_segments((long long)strips * s * batch, nq, F - 1, R, per_sm * sms());
    const int TQ = cdiv(nq, nseg);
    if ((long long)s * cdiv(nq, TQ) > 65535) return 0;
    dim3 grid(strips, s * cdiv(nq, TQ), batch);
    pwt_launch_pdl(k_swt_strip_fwd<F, SMODE>, grid, NT, smem, st, in, A, Hb, V, D, Nr, Nc, s, TQ, cdiv(nq, TQ), (long long)Nr * Nc,
                   pwt_pack_taps_fwd(f, F));
    return 1;
}
template <int F>
int launch_s(const float* in, float* A, float* Hb, float* V, float* D, int batch, int Nr, int Nc, int s,
             const PwtFilters& f, cudaStream_t st) {
    if (s == 1) return launch<F, 1>(in, A, Hb, V, D, batch, Nr, Nc, s, f, st);
    if (s == 2) return launch<F, 2>(in, A, Hb, V, D, batch, Nr, Nc, s, f, st);
    return launch<F, 0>(in, A, Hb, V, D, batch, Nr, Nc, s, f, st);
}

}  // namespace

// ---- inverse level ----------------------------------------------------------------------------------------
//   x = syn_x(syn_y(A, H), syn_y(V, D)),  syn(a, d)[g] = sum_j IL[F-1-j]/2 * a[g + (j-F/2)*s] + IH[F-1-j]/2 * d[...]
//   (columns first, like the reference: separable.cu:553-626)
// Same lattice walk.  The staged tile is 256 columns INCLUDING the reach of the row filter; the column synthesis
// runs in transposed form on all of them (a thread owns two adjacent columns of t1 or t2: 2-wide FMAs over the
// column pair with duplicated taps), finished t rows go to shared memory, the row synthesis produces the
// 256 - (F-1)*s columns the strip owns (aligned 128-bit tap loads) and stores them with 128-bit stores.
// A deferred soft / hard threshold (THR) is applied to the coefficients as the column pass reads them.
namespace {

struct TapsDup {                  // synthesis taps, halved and duplicated: l[j] = (IL[F-1-j]/2, same), h likewise
    float2 l[PWT_MAX_TAPS];
    float2 h[PWT_MAX_TAPS];
    // shifted pairs for two outputs one dilation step apart that read the same sample: lp[j] = (tap j, tap j-1)
    // (tap -1 = tap F = 0), so that one FFMA2 serves both outputs in the contiguous-window row pass (s = 1, 2)
    float2 lp[PWT_MAX_TAPS + 1];
    float2 hp[PWT_MAX_TAPS + 1];
};
struct Thr {
    float beta, beta_app;
    int app;
};
template <int THR>
__device__ __forceinline__ float thr1(float v, float beta) {
    // common.cu:19 (soft) / common.cu:63 (hard, strict >)
    if (THR == 1) return copysignf(fmaxf(fabsf(v) - beta, 0.0f), v);
    return (fabsf(v) - beta > 0.0f) ? v : 0.0f * v;
}

template <int F, int SMODE, int THR, int NBUF>
__global__ void __launch_bounds__(NT, F <= 8 ? (NBUF == 1 ? 4 : 3) : 2)
k_swt_strip_inv(const float* __restrict__ A, const float* __restrict__ Hb, const float* __restrict__ V,
                const float* __restrict__ D, float* __restrict__ out, int Nr, int Nc, int s, int TQ, int nseg,
                long long plane, const Thr thr, const __grid_constant__ TapsDup f) {
    constexpr int C = F / 2, R = chunk_rows(F);
    constexpr int TWD = 256;                           // staged columns = 64 groups
    extern __shared__ __align__(16) float sm[];
    float* raw = sm;                                   // [NBUF][R][4 bands][TWD]
    float* tb = sm + NBUF * R * 4 * TWD;               // [R][2][TWD] finished t1 | t2 rows
    const int tid = threadIdx.x;
    const int HL = (C * s + 3) & ~3, HR = ((F - 1 - C) * s + 3) & ~3;
    const int OWN = TWD - HL - HR;                     // columns this strip stores
    const int x0 = blockIdx.x * OWN;
    const int r = blockIdx.y / nseg, seg = blockIdx.y - r * nseg;
    const int nq = (Nr - r + s - 1) / s;
    const int q0 = seg * TQ, q1 = min(q0 + TQ, nq);
    if (q0 >= q1) return;
    const long long ib = blockIdx.z * plane;
    out += ib;
    const int m0 = q0 - C, m_last = q1 - 1 + (F - 1 - C);
    const int nchunks = (m_last - m0 + R) / R;
    // staging: thread -> (band, 16-byte group), all R rows
    const int sb = tid >> 6, sg = tid & 63;
    int gcol = (x0 - HL + 4 * sg) % Nc;
    if (gcol < 0) gcol += Nc;
    const float* band = (sb == 0 ? A : sb == 1 ? Hb : sb == 2 ? V : D) + ib + gcol;
    auto stage = [&](int c) {
        float* dst = raw + (NBUF == 2 ? (c & 1) * R * 4 * TWD : 0) + sb * TWD + 4 * sg;
#pragma unroll
        for (int rr = 0; rr < R; rr++) {
            const int m = min(m0 + c * R + rr, m_last);
            cp_async16(dst + rr * 4 * TWD, band + (unsigned)(wrap1(r + m * s, Nr) * Nc));
        }
        cp_async_commit();
    };
    // column pass ownership: two adjacent staged columns of t1 (from A, H) or t2 (from V, D)
    const int pl = tid >> 7, cp = tid & 127;
    const float2 zero2 = make_float2(0.f, 0.f);
    float2 acc[F];
#pragma unroll
    for (int a = 0; a < F; a++) acc[a] = zero2;
    // row pass ownership.  s = 1, 2: 64 threads per row, 4 adjacent output columns each (contiguous window).
    // s % 4 == 0: 32 threads per row, two 16-byte groups one tap step apart (q and q + s/4): the F+1 vectors
    // q + jj*s/4 feed both, which halves the shared-memory reads that bound these levels.  The t rows are stored
    // with their 16-byte groups XOR-swizzled (bit 3 of the group index into bit log2(s/4)) so that the
    // quarter-warps of those strided reads hit 8 different banks.
    const int sig = s >> 2;
    const int swm = (SMODE == 0 && sig < 8) ? sig : 0;
    const int tw_off = pl * TWD + 4 * ((cp >> 1) ^ (((cp >> 4) & 1) * swm)) + 2 * (cp & 1);
    const int prow = SMODE == 0 ? tid >> 5 : tid >> 6;
    const int pg = SMODE == 0 ? ((tid & 31) / max(sig, 1)) * 2 * sig + ((tid & 31) & (sig - 1)) : tid & 63;
    const int X = x0 + 4 * pg;
    const bool pvalid = 4 * pg < OWN && X < Nc;
    const bool pvalid2 = 4 * (pg + sig) < OWN && X + 4 * sig < Nc;
    int ubase = -(F - 1);

    pwt_pdl_wait();
    stage(0);
    for (int c = 0; c < nchunks; c++) {
        if (c == nchunks - 1) pwt_pdl_trigger();
        if (NBUF == 2 && c + 1 < nchunks) { stage(c + 1); cp_async_wait<1>(); }
        else cp_async_wait<0>();
        __syncthreads();                               // chunk c staged; everybody is done with tb
        const float* rb = raw + (NBUF == 2 ? (c & 1) * R * 4 * TWD : 0) + 2 * pl * TWD + 2 * cp;
        // ---- column synthesis, transposed form: stream row n = c*R + j feeds rows u = n - jj with tap jj ----
#pragma unroll
        for (int j = 0; j < R; j++) {
            float2 xa = *reinterpret_cast<const float2*>(rb + j * 4 * TWD);          // A or V
            float2 xd = *reinterpret_cast<const float2*>(rb + j * 4 * TWD + TWD);    // H or D
            if (THR) {
                xd.x = thr1<THR>(xd.x, thr.beta); xd.y = thr1<THR>(xd.y, thr.beta);
                if (pl) { xa.x = thr1<THR>(xa.x, thr.beta); xa.y = thr1<THR>(xa.y, thr.beta); }
                else if (thr.app) { xa.x = thr1<THR>(xa.x, thr.beta_app); xa.y = thr1<THR>(xa.y, thr.beta_app); }
            }
#pragma unroll
            for (int jj = 0; jj < F; jj++) {
                const int a = ((j - jj) % F + F) % F;
                acc[a] = __ffma2_rn(xa, f.l[jj], jj == 0 ? zero2 : acc[a]);
                acc[a] = __ffma2_rn(xd, f.h[jj], acc[a]);
            }
            *reinterpret_cast<float2*>(tb + j * 2 * TWD + tw_off) = acc[(j + 1) % F];               // row u = ubase + j
        }
        __syncthreads();                               // t rows complete, raw buffer free
        if (NBUF == 1 && c + 1 < nchunks) stage(c + 1);               // overlaps the row synthesis
        // ---- row synthesis of the R finished rows ----
        if (SMODE == 0) {
#pragma unroll
            for (int i = 0; i < (R + 7) / 8; i++) {
                const int j = prow + 8 * i, u = ubase + j;
                if (R % 8 != 0 && j >= R) break;
                const float* t1 = tb + j * 2 * TWD;
                float2 a01 = zero2, a23 = zero2, b01 = zero2, b23 = zero2;        // group pg, group pg + sig
#pragma unroll
                for (int jj = 0; jj <= F; jj++) {                                 // HL == C*s: tap jj of group pg sits at pg + jj*sig
                    const int g = min(pg + jj * sig, 63);
                    const int o = 4 * (g ^ (((g >> 3) & 1) * swm));
                    const float4 v1 = *reinterpret_cast<const float4*>(t1 + o);
                    const float4 v2 = *reinterpret_cast<const float4*>(t1 + TWD + o);
                    if (jj < F) {
                        a01 = __ffma2_rn(make_float2(v1.x, v1.y), f.l[jj], a01);
                        a23 = __ffma2_rn(make_float2(v1.z, v1.w), f.l[jj], a23);
                        a01 = __ffma2_rn(make_float2(v2.x, v2.y), f.h[jj], a01);
                        a23 = __ffma2_rn(make_float2(v2.z, v2.w), f.h[jj], a23);
                    }
                    if (jj >= 1) {
                        b01 = __ffma2_rn(make_float2(v1.x, v1.y), f.l[jj - 1], b01);
                        b23 = __ffma2_rn(make_float2(v1.z, v1.w), f.l[jj - 1], b23);
                        b01 = __ffma2_rn(make_float2(v2.x, v2.y), f.h[jj - 1], b01);
                        b23 = __ffma2_rn(make_float2(v2.z, v2.w), f.h[jj - 1], b23);
                    }
                }
                if (u >= 0 && u < q1 - q0) {
                    float* dst = out + (unsigned)((r + (q0 + u) * s) * Nc + X);
                    if (pvalid) *reinterpret_cast<float4*>(dst) = make_float4(a01.x, a01.y, a23.x, a23.y);
                    if (pvalid2) *reinterpret_cast<float4*>(dst + 4 * sig) = make_float4(b01.x, b01.y, b23.x, b23.y);
                }
            }
        } else {
#pragma unroll
            for (int i = 0; i < R / 4; i++) {
                const int j = prow + 4 * i, u = ubase + j;
                const float* t1 = tb + j * 2 * TWD + 4 * pg;                 // staged column x0 - HL + 4*pg
                const float* t2 = t1 + TWD;
                constexpr int S = SMODE ? SMODE : 1;
                constexpr int HLs = (C * S + 3) & ~3, DX = HLs - C * S, NV = (DX + 4 + (F - 1) * S + 3) / 4;
                // outputs e and e + S read the same sample with taps j and j - 1: pairs (0, S) and (1 or 2, ...)
                constexpr int EA0 = 0, EA1 = S == 1 ? 2 : 1;
                float2 pa = zero2, pb = zero2;                               // (out[EA0], out[EA0+S]), (out[EA1], out[EA1+S])
#pragma unroll
                for (int k = 0; k < NV; k++) {
                    const float4 v1 = *reinterpret_cast<const float4*>(t1 + 4 * k);
                    const float4 v2 = *reinterpret_cast<const float4*>(t2 + 4 * k);
                    const float x1[4] = {v1.x, v1.y, v1.z, v1.w}, x2[4] = {v2.x, v2.y, v2.z, v2.w};
#pragma unroll
                    for (int ee = 0; ee < 4; ee++) {
                        const int da = 4 * k + ee - DX - EA0, db = 4 * k + ee - DX - EA1;
                        if (da >= 0 && da % S == 0 && da / S <= F) {
                            pa = fma2s(x1[ee], f.lp[da / S], pa);
                            pa = fma2s(x2[ee], f.hp[da / S], pa);
                        }
                        if (db >= 0 && db % S == 0 && db / S <= F) {
                            pb = fma2s(x1[ee], f.lp[db / S], pb);
                            pb = fma2s(x2[ee], f.hp[db / S], pb);
                        }
                    }
                }
                if (pvalid && u >= 0 && u < q1 - q0)
                    *reinterpret_cast<float4*>(out + (unsigned)((r + (q0 + u) * s) * Nc + X)) =
                        S == 1 ? make_float4(pa.x, pa.y, pb.x, pb.y) : make_float4(pa.x, pb.x, pa.y, pb.y);
            }
        }
        ubase += R;
    }
}

template <int F, int SMODE, int THR, int NBUF>
int launch_inv_n(const float* A, const float* Hb, const float* V, const float* D, float* out, int batch, int Nr,
                 int Nc, int s, const PwtFilters& f, const Thr& thr, cudaStream_t st) {
    constexpr int C = F / 2, R = chunk_rows(F);
    const int HL = (C * s + 3) & ~3, HR = ((F - 1 - C) * s + 3) & ~3;
    const int OWN = 256 - HL - HR;
    if (OWN < 128) return 0;                           // dilation too large for overlapping 256-column tiles
    const size_t smem = sizeof(float) * ((size_t)NBUF * R * 4 * 256 + (size_t)R * 2 * 256);
    static PwtKernelOnce once;
    const int per_sm = pwt_kernel_once(once, k_swt_strip_inv<F, SMODE, THR, NBUF>, NT, smem, smem);
    if (!per_sm) return 0;
    const int strips = cdiv(Nc, OWN);
    const int nq = cdiv(Nr, s);
    const int nseg = pick_segments((long long)strips * s * batch, nq, F - 1, R, per_sm * sms());
    const int TQ = cdiv(nq, nseg);
    if ((long long)s * cdiv(nq, TQ) > 65535) return 0;
    TapsDup t;
    for (int j = 0; j < PWT_MAX_TAPS; j++) {
        const float l = j < F ? 0.5f * f.IL[F - 1 - j] : 0.f, h = j < F ? 0.5f * f.IH[F - 1 - j] : 0.f;
        t.l[j] = make_float2(l, l);
        t.h[j] = make_float2(h, h);
    }
    for (int j = 0; j <= PWT_MAX_TAPS; j++) {
        t.lp[j] = make_float2(j < F ? t.l[j].x : 0.f, j >= 1 && j <= F ? t.l[j - 1].x : 0.f);
        t.hp[j] = make_float2(j < F ? t.h[j].x : 0.f, j >= 1 && j <= F ? t.h[j - 1].x : 0.f);
    }
    dim3 grid(strips, s * cdiv(nq, TQ), batch);
    pwt_launch_pdl(k_swt_strip_inv<F, SMODE, THR, NBUF>, grid, NT, smem, st, A, Hb, V, D, out, Nr, Nc, s, TQ, cdiv(nq, TQ),
                   (long long)Nr * Nc, thr, t);
    return 1;
}
template <int F, int SMODE, int THR>
int launch_inv_t(const float* A, const float* Hb, const float* V, const float* D, float* out, int batch, int Nr,
                 int Nc, int s, const PwtFilters& f, const Thr& thr, cudaStream_t st) {
    const int nbuf = pwt_tuning().swt_nbuf;
    if (F <= 8 && nbuf == 2) return launch_inv_n<F, SMODE, THR, (F <= 8 ? 2 : 1)>(A, Hb, V, D, out, batch, Nr, Nc, s, f, thr, st);
    return launch_inv_n<F, SMODE, THR, 1>(A, Hb, V, D, out, batch, Nr, Nc, s, f, thr, st);
}
template <int F, int THR>
int launch_inv_s(const float* A, const float* Hb, const float* V, const float* D, float* out, int batch, int Nr,
                 int Nc, int s, const PwtFilters& f, const Thr& thr, cudaStream_t st) {
    if (s == 1) return launch_inv_t<F, 1, THR>(A, Hb, V, D, out, batch, Nr, Nc, s, f, thr, st);
    if (s == 2) return launch_inv_t<F, 2, THR>(A, Hb, V, D, out, batch, Nr, Nc, s, f, thr, st);
    return launch_inv_t<F, 0, THR>(A, Hb, V, D, out, batch, Nr, Nc, s, f, thr, st);
}
template <int F>
int launch_inv(const float* A, const float* Hb, const float* V, const float* D, float* out, int batch, int Nr,
               int Nc, int s, const PwtFilters& f, int thr_op, const Thr& thr, cudaStream_t st) {
    if (thr_op == PWT_OP_SOFT) return launch_inv_s<F, 1>(A, Hb, V, D, out, batch, Nr, Nc, s, f, thr, st);
    if (thr_op == PWT_OP_HARD) return launch_inv_s<F, 2>(A, Hb, V, D, out, batch, Nr, Nc, s, f, thr, st);
    return launch_inv_s<F, 0>(A, Hb, V, D, out, batch, Nr, Nc, s, f, thr, st);
}

}  // namespace

// Can the strip inverse run this level (and therefore apply a deferred threshold while loading)?
int pwt_strip_swt_inv2d_covers(int batch, int Nr, int Nc, int level, const PwtFilters& f, const void* A, const void* out) {
    const int F = f.hlen;
    if (level < 1 || level > 16) return 0;
    const int s = 1 << (level - 1);
    if ((F & 1) || F < 2 || F > 20 || (Nc & 3) || batch > 65535 || out == A) return 0;
    if ((F - 1) * s >= Nc || (F - 1) * s >= Nr || (long long)Nr * Nc >= (1LL << 31)) return 0;
    if ((((uintptr_t)out | (uintptr_t)A) & 15) != 0) return 0;
    if (pwt_tuning().no_strip_swt) return 0;
    const int C = F / 2, HL = (C * s + 3) & ~3, HR = ((F - 1 - C) * s + 3) & ~3;
    return 256 - HL - HR >= 128;
}

// thr_op < 0: no deferred operator; otherwise PWT_OP_SOFT / PWT_OP_HARD with beta (details of this level) and,
// when app != 0, beta_app for the approximation input.  Return 0 when not covered.
int pwt_strip_swt_inv2d(const float* A, const float* Hb, const float* V, const float* D, float* out, int batch,
                        int Nr, int Nc, int level, const PwtFilters& f, int thr_op, float beta, int app,
                        float beta_app, cudaStream_t st) {
    const int F = f.hlen;
    if (level < 1 || level > 16) return 0;
    const int s = 1 << (level - 1);
    if (!pwt_strip_swt_inv2d_covers(batch, Nr, Nc, level, f, A, out)) return 0;
    if ((((uintptr_t)Hb | (uintptr_t)V | (uintptr_t)D) & 15) != 0) return 0;
    Thr thr;
    thr.beta = beta;
    thr.beta_app = beta_app;
    thr.app = app;
    switch (F) {
#define X(FF) case FF: return launch_inv<FF>(A, Hb, V, D, out, batch, Nr, Nc, s, f, thr_op, thr, st);
        X(2) X(4) X(6) X(8) X(10) X(12) X(14) X(16) X(18) X(20)
#undef X
        default: return 0;
    }
}

// Return 0 when the configuration is not covered (the register kernels / generic kernels take over).
int pwt_strip_swt_fwd2d(const float* in, float* A, float* Hb, float* V, float* D, int batch, int Nr, int Nc,
                        int level, const PwtFilters& f, cudaStream_t st) {
    const int F = f.hlen;
    if (level < 1 || level > 16) return 0;
    const int s = 1 << (level - 1);
    if ((F & 1) || F < 2 || F > 20 || (Nc & 3) || batch > 65535 || in == A) return 0;
    if ((F - 1) * s >= Nc || (F - 1) * s >= Nr || (long long)Nr * Nc >= (1LL << 31)) return 0;
    if ((((uintptr_t)in | (uintptr_t)A | (uintptr_t)Hb | (uintptr_t)V | (uintptr_t)D) & 15) != 0) return 0;
    if (pwt_tuning().no_strip_swt) return 0;
    switch (F) {
#define X(FF) case FF: return launch_s<FF>(in, A, Hb, V, D, batch, Nr, Nc, s, f, st);
        X(2) X(4) X(6) X(8) X(10) X(12) X(14) X(16) X(18) X(20)
#undef X
        default: return 0;
    }
}
