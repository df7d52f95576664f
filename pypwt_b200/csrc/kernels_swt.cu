// Stationary (a trous) wavelet transform, 2D separable, one fused launch per level, in registers.
//
//   forward  out[g] = sum_j f[F-1-j] * x[(g + (j-c)*s) mod N], c = F/2-1, s = 2^(level-1)   separable.cu:409-493
//   inverse  x[g]   = sum_j (fl[F-1-j]/2) * a[(g + (j-F/2)*s) mod N] + (fh[F-1-j]/2) * d[...]  separable.cu:553-626
//
// Design (B200): no decimation means 4 output planes per input plane (20 B/px/level compulsory), so the
// kernels must not add intermediate planes (the reference, and our generic fallback, write and re-read
// two full-size row-pass planes: 36 B/px/level).  One thread owns 4 adjacent columns of a column strip and
// a CTA walks one residue class of rows (the lattice y = r + q*s), so the dilated column filter is an
// ordinary sliding window in registers.  The dilated ROW filter needs the neighbours' samples: each row is
// published once through a double-buffered shared-memory row (one barrier per row), every tap is an aligned
// 128-bit shared load at +-(j*s/4) vectors (s % 4 == 0) or comes from the few aligned vectors covering the
// window (s = 1, 2).  Each sample is read from HBM exactly once, with the next row's load already in flight.
//   forward : row filter (shared row) -> column sliding window (registers) -> 4 x 128-bit stores
//   inverse : column synthesis in transposed form (F pending output rows as accumulators) -> row synthesis
//             from the shared row; a deferred soft/hard threshold is applied to the bands as they are loaded.
#include <stdlib.h>

#include "pwt_internal.h"

namespace {

__device__ __forceinline__ int wrap1_per(int i, int N) {   // -N <= i < 2N
    if (i < 0) i += N;
    if (i >= N) i -= N;
    return i;
}
__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ void stg4(float* p, const float4& v) { *reinterpret_cast<float4*>(p) = v; }
__device__ __forceinline__ void fma4(float4& acc, const float4& v, float t) {
    acc.x = fmaf(v.x, t, acc.x);
    acc.y = fmaf(v.y, t, acc.y);
    acc.z = fmaf(v.z, t, acc.z);
    acc.w = fmaf(v.w, t, acc.w);
}

constexpr int kThreads = 128;

struct SwtGeom {
    int Nr, Nc, s, TQ;           // plane size, dilation, lattice steps per task
    int strips, chunks;          // column strips of 4*kThreads, chunks of TQ steps per residue class
    long long plane;             // elements per image
};

// ---- forward: in -> A, H, V, D -----------------------------------------------------------------
// One CTA walks a residue class of rows (lattice y = r + q*s) down a column strip.  Every input row is read
// from memory once (one 128-bit load per thread, the next row already in flight), published through a
// double-buffered shared-memory row (one barrier per row) and row-filtered from there with the dilated taps;
// the column pass is a sliding window of the F filtered rows in registers.  Column strips overlap by the reach
// of the row filter (threads in the overlap only load).
__host__ __device__ inline int swt_fwd_halo_l(int F, int s) { return ((F / 2 - 1) * s + 3) & ~3; }
__host__ __device__ inline int swt_fwd_halo_r(int F, int s) { return ((F / 2) * s + 3) & ~3; }

template <int F, int SMODE, int MINB>
__global__ void __launch_bounds__(kThreads, MINB)
k_swt_fwd(const float* __restrict__ in, float* __restrict__ A, float* __restrict__ Hb, float* __restrict__ V,
          float* __restrict__ D, const SwtGeom g, const __grid_constant__ PwtFilters f) {
    constexpr int C = F / 2 - 1;
    const int Nr = g.Nr, Nc = g.Nc, s = g.s;
    const int HLa = swt_fwd_halo_l(F, s), HRa = swt_fwd_halo_r(F, s);
    const int OWN = 4 * kThreads - HLa - HRa;
    const int strip = blockIdx.x % g.strips;
    const int rc = blockIdx.x / g.strips;
    const int r = rc % s, chunk = rc / s;                    // residue class of rows, chunk along the lattice
    const int nq = (Nr - r + s - 1) / s;                     // lattice length of this class
    const int q0 = chunk * g.TQ;
    if (q0 >= nq) return;                                    // CTA-uniform
    const int q1 = min(q0 + g.TQ, nq);
    const int tid = threadIdx.x;
    const int lc = 4 * tid;
    const int xg = strip * OWN - HLa + lc;
    int xw = xg % Nc;
    if (xw < 0) xw += Nc;
    const bool own = lc >= HLa && lc < HLa + OWN && xg < Nc;
    const long long ib = blockIdx.y * g.plane;
    in += ib + xw;
    const long long ob = ib + xw;

    __shared__ float4 srow[2][kThreads];
    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
    float4 wl[F], wh[F];                                     // row-filtered rows y + (j-C)*s of the current output row
#pragma unroll
    for (int j = 0; j < F; j++) wl[j] = wh[j] = z;
    int buf = 0;
    const int m_last = q1 - 1 + (F - 1 - C);
    float4 nx = ldg4(in + (long long)wrap1_per(r + (q0 - C) * s, Nr) * Nc);
    for (int m = q0 - C; m <= m_last; m++) {                 // input lattice index
        srow[buf][tid] = nx;
        if (m < m_last) nx = ldg4(in + (long long)wrap1_per(r + (m + 1) * s, Nr) * Nc);
        __syncthreads();
        if (own) {
            float4 lo = z, hi = z;
            if (SMODE == 0) {
                const int step = s >> 2;
#pragma unroll
                for (int j = 0; j < F; j++) {
                    const float4 v = srow[buf][tid + (j - C) * step];
                    fma4(lo, v, f.L[F - 1 - j]);
                    fma4(hi, v, f.H[F - 1 - j]);
                }
            } else {
                constexpr int S = SMODE;
                constexpr int BL = ((C * S + 3) / 4) * 4, BR = (((F - 1 - C) * S + 3) / 4) * 4;
                constexpr int NE = (BL + 4 + BR) / 4;
                float ext[4 * NE];
#pragma unroll
                for (int k = 0; k < NE; k++) {
                    const float4 v = srow[buf][tid - BL / 4 + k];
                    ext[4 * k] = v.x; ext[4 * k + 1] = v.y; ext[4 * k + 2] = v.z; ext[4 * k + 3] = v.w;
                }
#pragma unroll
                for (int j = 0; j < F; j++) {
                    const float ta = f.L[F - 1 - j], tb = f.H[F - 1 - j];
                    const int o = BL + (j - C) * S;
                    lo.x = fmaf(ext[o], ta, lo.x);     hi.x = fmaf(ext[o], tb, hi.x);
                    lo.y = fmaf(ext[o + 1], ta, lo.y); hi.y = fmaf(ext[o + 1], tb, hi.y);
                    lo.z = fmaf(ext[o + 2], ta, lo.z); hi.z = fmaf(ext[o + 2], tb, hi.z);
                    lo.w = fmaf(ext[o + 3], ta, lo.w); hi.w = fmaf(ext[o + 3], tb, hi.w);
                }
            }
            wl[F - 1] = lo;
            wh[F - 1] = hi;
            const int q = m - (F - 1) + C;                   // output row whose window is complete
            if (q >= q0) {
                float4 a = z, h = z, v = z, d = z;
#pragma unroll
                for (int j = 0; j < F; j++) {
                    const float tl = f.L[F - 1 - j], th = f.H[F - 1 - j];
                    fma4(a, wl[j], tl);
                    fma4(h, wl[j], th);      // (Lx, Hy)
                    fma4(v, wh[j], tl);      // (Hx, Ly)
                    fma4(d, wh[j], th);
                }
                const long long o = ob + (long long)(r + q * s) * Nc;
                stg4(A + o, a);
                stg4(Hb + o, h);
                stg4(V + o, v);
                stg4(D + o, d);
            }
#pragma unroll
            for (int j = 0; j < F - 1; j++) {
                wl[j] = wl[j + 1];
                wh[j] = wh[j + 1];
            }
        }
        buf ^= 1;
    }
}

// ---- inverse: A, H, V, D -> out ---------------------------------------------------------------
// Columns first, like the reference (separable.cu:553-626): t1 = syn_y(A, H), t2 = syn_y(V, D), then
// out = syn_x(t1, t2).  The column synthesis runs in TRANSPOSED form: a thread keeps the F pending output
// rows of t1 and t2 as accumulators and adds every band row it loads to all of them, so each coefficient
// is read from memory exactly once (4 x 128-bit loads per lattice row and thread; the direct form would
// need 4 sliding windows = 4F vector registers).  That single visit is also where a deferred soft / hard
// threshold is applied (THR), at no extra traffic.  Finished t1/t2 rows go through a double-buffered
// shared-memory row (one barrier per output row) for the dilated row synthesis; column strips overlap by
// the reach of the row filter (C*s on the left, (F/2-1)*s on the right), recomputed instead of exchanged.
template <int THR>
__device__ __forceinline__ float thr1(float v, float beta) {
    // common.cu:19 (soft) / common.cu:63 (hard, strict >)
    if (THR == 1) return copysignf(fmaxf(fabsf(v) - beta, 0.0f), v);
    return (fabsf(v) - beta > 0.0f) ? v : 0.0f * v;
}
template <int THR>
__device__ __forceinline__ float4 thr4(float4 v, float b) {
    return make_float4(thr1<THR>(v.x, b), thr1<THR>(v.y, b), thr1<THR>(v.z, b), thr1<THR>(v.w, b));
}

struct SwtThr {
    float beta;       // detail bands of this level
    float beta_app;   // approximation (only when app != 0: coarsest level of a thresholded-approximation call)
    int app;
};

template <int F, int SMODE>
struct SwtInvGeo {
    static constexpr int C = F / 2;                                          // separable.cu:565-568
    static constexpr int S = SMODE;                                          // 1, 2, or 0 (runtime s, s % 4 == 0)
};
__host__ __device__ inline int swt_halo_l(int F, int s) { return ((F / 2) * s + 3) & ~3; }
__host__ __device__ inline int swt_halo_r(int F, int s) { return ((F / 2 - 1) * s + 3) & ~3; }

template <int F, int SMODE, int MINB, int THR>
__global__ void __launch_bounds__(kThreads, MINB)
k_swt_inv(const float* __restrict__ A, const float* __restrict__ Hb, const float* __restrict__ V,
          const float* __restrict__ D, float* __restrict__ out, const SwtGeom g, const SwtThr thr,
          const __grid_constant__ PwtFilters f) {
    constexpr int C = F / 2;
    const int Nr = g.Nr, Nc = g.Nc, s = g.s;
    const int HLa = swt_halo_l(F, s), HRa = swt_halo_r(F, s);
    const int OWN = 4 * kThreads - HLa - HRa;                // columns a strip owns (stores)
    const int strip = blockIdx.x % g.strips;
    const int rc = blockIdx.x / g.strips;
    const int r = rc % s, chunk = rc / s;                    // residue class of rows, chunk along the lattice
    const int nq = (Nr - r + s - 1) / s;
    const int q0 = chunk * g.TQ;
    if (q0 >= nq) return;                                    // CTA-uniform
    const int q1 = min(q0 + g.TQ, nq);
    const int tid = threadIdx.x;
    const int lc = 4 * tid;                                  // local column of this thread's vector
    const int xg = strip * OWN - HLa + lc;                   // global column (may lie outside [0, Nc): wraps)
    int xw = xg % Nc;
    if (xw < 0) xw += Nc;                                    // Nc % 4 == 0: a vector never straddles the wrap
    const bool own = lc >= HLa && lc < HLa + OWN && xg < Nc;
    const long long ib = blockIdx.y * g.plane;
    A += ib + xw; Hb += ib + xw; V += ib + xw; D += ib + xw;
    out += ib + xw;

    __shared__ float4 sbuf[2][2][kThreads];                  // [buffer][t1 | t2][thread]
    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
    float4 acc1[F], acc2[F];                                 // pending rows of t1 / t2; slot 0 completes next
#pragma unroll
    for (int i = 0; i < F; i++) acc1[i] = acc2[i] = z;
    int buf = 0;
    const int m_last = q1 - 1 + (F - 1 - C);
    long long ro = (long long)wrap1_per(r + (q0 - C) * s, Nr) * Nc;
    float4 na = ldg4(A + ro), nh = ldg4(Hb + ro), nv = ldg4(V + ro), nd = ldg4(D + ro);
    for (int m = q0 - C; m <= m_last; m++) {                 // input lattice index
        float4 a = na, h = nh, v = nv, d = nd;
        if (m < m_last) {                                    // the next band rows are in flight during this one
            ro = (long long)wrap1_per(r + (m + 1) * s, Nr) * Nc;
            na = ldg4(A + ro); nh = ldg4(Hb + ro); nv = ldg4(V + ro); nd = ldg4(D + ro);
        }
        if (THR) {
            h = thr4<THR>(h, thr.beta);
            v = thr4<THR>(v, thr.beta);
            d = thr4<THR>(d, thr.beta);
            if (thr.app) a = thr4<THR>(a, thr.beta_app);
        }
        // input m feeds output q = m - j + C with tap IL[F-1-j] / 2; slot i = F-1-j  (separable.cu:621-622)
#pragma unroll
        for (int i = 0; i < F; i++) {
            const float cl = 0.5f * f.IL[i], ch = 0.5f * f.IH[i];
            fma4(acc1[i], a, cl);
            fma4(acc1[i], h, ch);
            fma4(acc2[i], v, cl);
            fma4(acc2[i], d, ch);
        }
        const int q = m - (F - 1) + C;                       // the output row that is complete now
        if (q >= q0) {                                       // CTA-uniform
            sbuf[buf][0][tid] = acc1[0];
            sbuf[buf][1][tid] = acc2[0];
            __syncthreads();
            if (own) {
                float4 o = z;
                if (SMODE == 0) {
                    const int step = s >> 2;                 // vectors per tap
#pragma unroll
                    for (int j = 0; j < F; j++) {
                        const int k = tid + (j - C) * step;
                        fma4(o, sbuf[buf][0][k], 0.5f * f.IL[F - 1 - j]);
                        fma4(o, sbuf[buf][1][k], 0.5f * f.IH[F - 1 - j]);
                    }
                } else {
                    constexpr int S = SMODE;
                    constexpr int BL = ((C * S + 3) / 4) * 4, BR = (((F - 1 - C) * S + 3) / 4) * 4;
                    constexpr int NE = (BL + 4 + BR) / 4;
                    float e1[4 * NE], e2[4 * NE];
#pragma unroll
                    for (int k = 0; k < NE; k++) {
                        const float4 p1 = sbuf[buf][0][tid - BL / 4 + k], p2 = sbuf[buf][1][tid - BL / 4 + k];
                        e1[4 * k] = p1.x; e1[4 * k + 1] = p1.y; e1[4 * k + 2] = p1.z; e1[4 * k + 3] = p1.w;
                        e2[4 * k] = p2.x; e2[4 * k + 1] = p2.y; e2[4 * k + 2] = p2.z; e2[4 * k + 3] = p2.w;
                    }
#pragma unroll
                    for (int j = 0; j < F; j++) {
                        const float tl = 0.5f * f.IL[F - 1 - j], th = 0.5f * f.IH[F - 1 - j];
                        const int o0 = BL + (j - C) * S;
                        o.x = fmaf(e1[o0], tl, o.x);     o.x = fmaf(e2[o0], th, o.x);
                        o.y = fmaf(e1[o0 + 1], tl, o.y); o.y = fmaf(e2[o0 + 1], th, o.y);
                        o.z = fmaf(e1[o0 + 2], tl, o.z); o.z = fmaf(e2[o0 + 2], th, o.z);
                        o.w = fmaf(e1[o0 + 3], tl, o.w); o.w = fmaf(e2[o0 + 3], th, o.w);
                    }
                }
                stg4(out + (long long)(r + q * s) * Nc, o);
            }
            buf ^= 1;
        }
#pragma unroll
        for (int i = 0; i < F - 1; i++) {
            acc1[i] = acc1[i + 1];
            acc2[i] = acc2[i + 1];
        }
        acc1[F - 1] = z;
        acc2[F - 1] = z;
    }
}

inline int cdiv(int a, int b) { return (a + b - 1) / b; }

SwtGeom make_geom(int Nr, int Nc, int level) {
    SwtGeom g;
    g.Nr = Nr;
    g.Nc = Nc;
    g.s = 1 << (level - 1);
    g.strips = cdiv(Nc, 4 * kThreads);
    const int nq = cdiv(Nr, g.s);
    int tq = 64;
    if (pwt_tuning().swt_tq > 0) tq = pwt_tuning().swt_tq;
    while (tq > 16 && (long long)g.strips * g.s * cdiv(nq, tq) < 2LL * pwt_sm_count() * 4) tq >>= 1;   // keep the GPU full
    g.TQ = tq;
    g.chunks = cdiv(nq, tq);
    g.plane = (long long)Nr * Nc;
    return g;
}

template <int F, int MINB>
int launch_fwd(const float* in, float* A, float* Hb, float* V, float* D, int batch, int Nr, int Nc, int level,
               const PwtFilters& f, cudaStream_t st) {
    SwtGeom g = make_geom(Nr, Nc, level);
    const int own = 4 * kThreads - swt_fwd_halo_l(F, g.s) - swt_fwd_halo_r(F, g.s);
    if (own < 2 * kThreads) return 0;                        // dilation too large for overlapping strips
    g.strips = cdiv(Nc, own);
    dim3 grid(g.strips * g.s * g.chunks, batch);
    if (g.s == 1) k_swt_fwd<F, 1, MINB><<<grid, kThreads, 0, st>>>(in, A, Hb, V, D, g, f);
    else if (g.s == 2) k_swt_fwd<F, 2, MINB><<<grid, kThreads, 0, st>>>(in, A, Hb, V, D, g, f);
    else k_swt_fwd<F, 0, MINB><<<grid, kThreads, 0, st>>>(in, A, Hb, V, D, g, f);
    return 1;
}
template <int F, int MINB, int THR>
int launch_inv_t(const float* A, const float* Hb, const float* V, const float* D, float* out, int batch, int Nr,
                 int Nc, int level, const PwtFilters& f, const SwtThr& thr, cudaStream_t st) {
    SwtGeom g = make_geom(Nr, Nc, level);
    const int own = 4 * kThreads - swt_halo_l(F, g.s) - swt_halo_r(F, g.s);
    if (own < 2 * kThreads) return 0;                        // dilation too large for overlapping strips
    g.strips = cdiv(Nc, own);
    dim3 grid(g.strips * g.s * g.chunks, batch);
    if (g.s == 1) k_swt_inv<F, 1, MINB, THR><<<grid, kThreads, 0, st>>>(A, Hb, V, D, out, g, thr, f);
    else if (g.s == 2) k_swt_inv<F, 2, MINB, THR><<<grid, kThreads, 0, st>>>(A, Hb, V, D, out, g, thr, f);
    else k_swt_inv<F, 0, MINB, THR><<<grid, kThreads, 0, st>>>(A, Hb, V, D, out, g, thr, f);
    return 1;
}
template <int F, int MINB>
int launch_inv(const float* A, const float* Hb, const float* V, const float* D, float* out, int batch, int Nr,
               int Nc, int level, const PwtFilters& f, int thr_op, const SwtThr& thr, cudaStream_t st) {
    if (thr_op == PWT_OP_SOFT) return launch_inv_t<F, MINB, 1>(A, Hb, V, D, out, batch, Nr, Nc, level, f, thr, st);
    if (thr_op == PWT_OP_HARD) return launch_inv_t<F, MINB, 2>(A, Hb, V, D, out, batch, Nr, Nc, level, f, thr, st);
    return launch_inv_t<F, MINB, 0>(A, Hb, V, D, out, batch, Nr, Nc, level, f, thr, st);
}

bool covered(int F, int batch, int Nr, int Nc, int level, const void* p0, const void* p1) {
    const int s = 1 << (level - 1);
    if ((F & 1) || F < 2 || F > 12 || Nc % 4 != 0 || batch > 65535) return false;
    if ((F - 1) * s >= Nc || (F - 1) * s >= Nr) return false;          // single-step wrap must suffice
    if ((((uintptr_t)p0 | (uintptr_t)p1) & 15) != 0) return false;
    if (pwt_tuning().no_fast_swt) return false;
    return true;
}

}  // namespace

// Return 0 when the configuration is not covered (the generic two-pass kernels take over).
int pwt_fast_swt_fwd2d(const float* in, float* A, float* Hb, float* V, float* D, int batch, int Nr, int Nc,
                       int level, const PwtFilters& f, cudaStream_t st) {
    const int F = f.hlen;
    if (!covered(F, batch, Nr, Nc, level, in, A) || in == A) return 0;
    if ((((uintptr_t)Hb | (uintptr_t)V | (uintptr_t)D) & 15) != 0) return 0;
    switch (F) {
        case 2: return launch_fwd<2, 4>(in, A, Hb, V, D, batch, Nr, Nc, level, f, st);
        case 4: return launch_fwd<4, 4>(in, A, Hb, V, D, batch, Nr, Nc, level, f, st);
        case 6: return launch_fwd<6, 3>(in, A, Hb, V, D, batch, Nr, Nc, level, f, st);
        case 8: return launch_fwd<8, 3>(in, A, Hb, V, D, batch, Nr, Nc, level, f, st);
        case 10: return launch_fwd<10, 2>(in, A, Hb, V, D, batch, Nr, Nc, level, f, st);
        case 12: return launch_fwd<12, 2>(in, A, Hb, V, D, batch, Nr, Nc, level, f, st);
        default: return 0;
    }
}

// Can the fused inverse run this level (and therefore apply a deferred threshold while loading)?
int pwt_fast_swt_inv2d_covers(int batch, int Nr, int Nc, int level, const PwtFilters& f, const void* A,
                              const void* out) {
    const int F = f.hlen, s = 1 << (level - 1);
    if (!covered(F, batch, Nr, Nc, level, A, out) || out == A) return 0;
    return 4 * kThreads - swt_halo_l(F, s) - swt_halo_r(F, s) >= 2 * kThreads;
}

// thr_op < 0: no deferred operator; otherwise PWT_OP_SOFT / PWT_OP_HARD with beta (details of this level) and,
// when app != 0, beta_app for the approximation input.
int pwt_fast_swt_inv2d(const float* A, const float* Hb, const float* V, const float* D, float* out, int batch,
                       int Nr, int Nc, int level, const PwtFilters& f, int thr_op, float beta, int app,
                       float beta_app, cudaStream_t st) {
    const int F = f.hlen;
    if (!pwt_fast_swt_inv2d_covers(batch, Nr, Nc, level, f, A, out)) return 0;
    if ((((uintptr_t)Hb | (uintptr_t)V | (uintptr_t)D) & 15) != 0) return 0;
    SwtThr thr;
    thr.beta = beta;
    thr.beta_app = beta_app;
    thr.app = app;
    switch (F) {
        case 2: return launch_inv<2, 4>(A, Hb, V, D, out, batch, Nr, Nc, level, f, thr_op, thr, st);
        case 4: return launch_inv<4, 4>(A, Hb, V, D, out, batch, Nr, Nc, level, f, thr_op, thr, st);
        case 6: return launch_inv<6, 3>(A, Hb, V, D, out, batch, Nr, Nc, level, f, thr_op, thr, st);
        case 8: return launch_inv<8, 3>(A, Hb, V, D, out, batch, Nr, Nc, level, f, thr_op, thr, st);
        case 10: return launch_inv<10, 2>(A, Hb, V, D, out, batch, Nr, Nc, level, f, thr_op, thr, st);
        case 12: return launch_inv<12, 2>(A, Hb, V, D, out, batch, Nr, Nc, level, f, thr_op, thr, st);
        default: return 0;
    }
}
