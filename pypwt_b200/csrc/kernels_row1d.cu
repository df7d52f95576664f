// Batched 1D DWT / IDWT, ALL levels in one launch (reference: w_forward_separable_1d separable.cu:214-236,
// w_inverse_separable_1d :368-395, kern_haar1d_fwd/inv haar.cu:132-160 -- one launch per level there, and here until
// round 2: 28 B/px moved for a 3-level forward + inverse against 16 compulsory).
//
// Rows are independent signals and a whole row fits in shared memory (8192 floats = 32 KB), so a CTA stages a group
// of rows ONCE, runs every level on chip -- the approximation ping-pongs between two shared buffers, the detail
// coefficients of each level go straight to global memory with 128-bit stores -- and only the final approximation
// leaves as well: 8 B/px per direction for any number of levels.  The periodic extension (period rounded up to even,
// the extra sample of an odd length repeating the last one, separable.cu:98-102) is materialised as a halo around the
// staged row, so the tap loops are plain window reads: each thread produces 4 low-pass + 4 high-pass outputs from one
// window of F + 6 samples read as aligned 128-bit groups, every multiply-add a 2-wide FFMA2 on (low, high) tap pairs.
// The inverse stages A_L and every detail band of the group up front (one wait for all of them), then synthesises
// level by level in the polyphase form (even, odd output pairs; no zero insertion), 8 outputs per thread.
// Lengths that are not multiples of 4 take scalar staging / stores on the levels concerned (same arithmetic).
#include "pwt_internal.h"

namespace {
constexpr int NT = 256;

__device__ __forceinline__ void cp_async16(float* dst, const float* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ float2 fma2s(float x, float2 t, float2 acc) { return __ffma2_rn(make_float2(x, x), t, acc); }
__device__ __forceinline__ void stg_cs(float4* p, float4 v) { __stcs(p, v); }

struct RowLevels {
    float* D[PWT_MAX_LEVELS];        // detail band of level l + 1, rows x n[l + 1]
    int n[PWT_MAX_LEVELS + 1];       // n[l]: signal length at level l (n[0] = Nc)
    int dofs[PWT_MAX_LEVELS];        // inverse: offset (floats) of level l + 1's staged details inside a row's region
    int L;
};

template <int F>
struct RowGeo {
    static constexpr int NP = 4;                           // output pairs per thread and step
    static constexpr int P = F / 2 - 1;                    // analysis: output k reads inputs 2k - P .. 2k - P + F - 1
    static constexpr int HL = (F / 2 + 6 + 3) & ~3;        // halo (floats) kept on both sides of a staged signal
    static constexpr int OFF = (4 - (P & 3)) & 3;          // the analysis window starts OFF floats after an aligned group
    static constexpr int NV = (OFF + F + 2 * (NP - 1) + 3) / 4;   // float4 groups of the analysis window (NP output pairs)
    static constexpr int S1 = (P + 1) >> 1;                // synthesis: output pair j reads bands j - S1 .. j - S1 + F/2
    static constexpr int OFFI = (4 - (S1 & 3)) & 3;
    static constexpr int NVI = (OFFI + F / 2 + NP + 3) / 4;       // float4 groups of the synthesis window (NP output pairs)
    static constexpr int W = F / 2 + 1;                    // synthesis window positions per output pair
};
// pitch of a staged signal of length n: halo + n rounded up to even (+ slack for the last chunk) + halo
__host__ __device__ constexpr int row_pitch(int n, int HL) { return HL + ((n + 1 + 3) & ~3) + HL + 8; }

// Measured and dropped (profiles/r02_notes.md): an XOR swizzle of the 128-bit groups (the window reads of neighbouring
// threads start 8 floats apart: 2-way bank conflicts), 8 output pairs per thread, and sector-complete paired stores --
// each slower than this plain version; the kernel is bound by instruction issue, not by shared-memory wavefronts.
__device__ __forceinline__ int sa(int e) { return e; }                                // element index -> float offset
__device__ __forceinline__ float4 lds4(const float* row, int q) { return *reinterpret_cast<const float4*>(row + 4 * q); }
__device__ __forceinline__ void sts4(float* row, int q, float4 v) { *reinterpret_cast<float4*>(row + 4 * q) = v; }

__device__ __forceinline__ int mod_pos(int i, int n) {
    i %= n;
    return i < 0 ? i + n : i;
}
// Work of a group is spread as (row, item) with a power-of-two number of threads per row, so that no index needs a
// division: threads per row T = 2^tb >= items per row (capped at the CTA), rows advance NT / T at a time.
__device__ __forceinline__ int log2_ceil_cap(int n) {            // smallest tb with 2^tb >= n, capped at log2(NT)
    int tb = 0;
    while ((1 << tb) < n && (1 << tb) < NT) tb++;
    return tb;
}
// periodic halo of the ANALYSIS input (period Ne = n rounded up to even, x~[n] = x[n-1] for odd n)
template <int HL>
__device__ __forceinline__ void fill_halo_fwd(float* buf, int pitch, int nrows, int n, int tid) {
    static_assert(2 * HL + 1 <= 64, "halo");
    const int Ne = n + (n & 1);
    const int j = tid & 63;
    if (j > 2 * HL) return;
    for (int r = tid >> 6; r < nrows; r += NT / 64) {
        float* b = buf + r * pitch;
        if (j < 2 * HL) {
            const int pos = j < HL ? -1 - j : Ne + j - HL;         // left, then right halo
            int s = pos < 0 ? pos + Ne : pos - Ne;
            if ((unsigned)s >= (unsigned)Ne) s = mod_pos(pos, Ne); // signals shorter than the halo
            b[sa(HL + pos)] = b[sa(HL + (s >= n ? n - 1 : s))];
        } else if (n & 1) {
            b[sa(HL + n)] = b[sa(HL + n - 1)];
        }
    }
}
// periodic halo of a SYNTHESIS input (period n)
template <int HL>
__device__ __forceinline__ void fill_halo_inv(float* buf, int pitch, int nrows, int n, int tid) {
    static_assert(2 * HL <= 64, "halo");
    const int j = tid & 63;
    if (j >= 2 * HL) return;
    for (int r = tid >> 6; r < nrows; r += NT / 64) {
        float* b = buf + r * pitch;
        const int pos = j < HL ? -1 - j : n + j - HL;
        int s = pos < 0 ? pos + n : pos - n;
        if ((unsigned)s >= (unsigned)n) s = mod_pos(pos, n);
        b[sa(HL + pos)] = b[sa(HL + s)];
    }
}
// stage `nrows` rows of length n (global, dense) into buf (+HL), 16-byte copies when every row start is aligned
__device__ __forceinline__ void stage_rows(float* buf, int pitch, int HL, const float* g, int nrows, int n, int tid) {
    if ((n & 3) == 0 && (((uintptr_t)g) & 15) == 0) {
        const int nv = n >> 2, tb = log2_ceil_cap(nv), T = 1 << tb;
        for (int r = tid >> tb; r < nrows; r += NT >> tb)
            for (int c = tid & (T - 1); c < nv; c += T) cp_async16(buf + r * pitch + HL + 4 * c, g + (size_t)r * n + 4 * c);
    } else {
        const int tb = log2_ceil_cap(n), T = 1 << tb;
        for (int r = tid >> tb; r < nrows; r += NT >> tb)
            for (int c = tid & (T - 1); c < n; c += T) buf[r * pitch + sa(HL + c)] = __ldg(g + (size_t)r * n + c);
    }
}
__device__ __forceinline__ void store4(float* g, int k0, int n, float4 v, bool vec) {
    if (vec) {
        stg_cs(reinterpret_cast<float4*>(g + k0), v);
    } else {
        if (k0 < n) g[k0] = v.x;
        if (k0 + 1 < n) g[k0 + 1] = v.y;
        if (k0 + 2 < n) g[k0 + 2] = v.z;
        if (k0 + 3 < n) g[k0 + 3] = v.w;
    }
}
// ---- forward: all levels --------------------------------------------------------------------------
template <int F>
__global__ void __launch_bounds__(NT)
k_row_fwd(const float* __restrict__ in, float* __restrict__ A, const __grid_constant__ RowLevels lv,
          const __grid_constant__ PwtTapsFwd tp, int rows, int R, int S0, int S1) {
    using G = RowGeo<F>;
    constexpr int NP = G::NP;
    extern __shared__ __align__(16) float sm[];
    float* buf0 = sm;
    float* buf1 = sm + (size_t)R * S0;
    const int tid = threadIdx.x;
    const int L = lv.L, N0 = lv.n[0];
    const int ngroups = (rows + R - 1) / R;
    pwt_pdl_wait();
    for (int g = blockIdx.x; g < ngroups; g += gridDim.x) {
        const int r0 = g * R, nr = rows - r0 < R ? rows - r0 : R;
        if (g + (int)gridDim.x >= ngroups) pwt_pdl_trigger();
        stage_rows(buf0, S0, G::HL, in + (size_t)r0 * N0, nr, N0, tid);
        cp_async_commit();
        cp_async_wait_all();
        __syncthreads();
        float* cur = buf0;
        float* nxt = buf1;
        int Sc = S0, Sn = S1;
        for (int l = 1; l <= L; l++) {
            const int nin = lv.n[l - 1], nout = lv.n[l];
            fill_halo_fwd<G::HL>(cur, Sc, nr, nin, tid);
            __syncthreads();
            const int nch = (nout + NP - 1) / NP;
            const bool last = l == L;
            float* Dg = lv.D[l - 1] + (size_t)r0 * nout;
            float* Ag = A + (size_t)r0 * nout;
            const bool vec = (nout & 3) == 0 && ((((uintptr_t)Dg) | ((uintptr_t)Ag)) & 15) == 0;
            const int tb = log2_ceil_cap(nch), T = 1 << tb;
            for (int r = tid >> tb; r < nr; r += NT >> tb) {
                const float* crow = cur + r * Sc;
                float* drow = Dg + (size_t)r * nout;
                float* arow = Ag + (size_t)r * nout;
                float* nrow = nxt + r * Sn;
                for (int k0 = NP * (tid & (T - 1)); k0 < nout; k0 += NP * T) {
                    float4 lo, hi;
                    if (F == 2) {                                                   // haar.cu:132-143
                        const float c = 0.70710678118654746f;
                        const float4 u = lds4(crow, (G::HL + 2 * k0) >> 2), v = lds4(crow, ((G::HL + 2 * k0) >> 2) + 1);
                        lo = make_float4(c * (u.x + u.y), c * (u.z + u.w), c * (v.x + v.y), c * (v.z + v.w));
                        hi = make_float4(c * (u.x - u.y), c * (u.z - u.w), c * (v.x - v.y), c * (v.z - v.w));
                    } else {
                        const int q0 = (G::HL + 2 * k0 - G::P - G::OFF) >> 2;
                        float w[4 * G::NV];
#pragma unroll
                        for (int q = 0; q < G::NV; q++) {
                            const float4 t = lds4(crow, q0 + q);
                            w[4 * q] = t.x; w[4 * q + 1] = t.y; w[4 * q + 2] = t.z; w[4 * q + 3] = t.w;
                        }
                        float2 p0 = make_float2(0.f, 0.f), p1 = p0, p2 = p0, p3 = p0; // (low, high) of outputs k0 .. k0 + 3
#pragma unroll
                        for (int j = 0; j < F; j++) {
                            p0 = fma2s(w[G::OFF + j], tp.t[j], p0);
                            p1 = fma2s(w[G::OFF + 2 + j], tp.t[j], p1);
                            p2 = fma2s(w[G::OFF + 4 + j], tp.t[j], p2);
                            p3 = fma2s(w[G::OFF + 6 + j], tp.t[j], p3);
                        }
                        lo = make_float4(p0.x, p1.x, p2.x, p3.x);
                        hi = make_float4(p0.y, p1.y, p2.y, p3.y);
                    }
                    store4(drow, k0, nout, hi, vec);
                    if (last) store4(arow, k0, nout, lo, vec);
                    else sts4(nrow, (G::HL + k0) >> 2, lo);
                }
            }
            __syncthreads();
            float* t = cur; cur = nxt; nxt = t;
            const int ts = Sc; Sc = Sn; Sn = ts;
        }
    }
}

// ---- inverse: all levels --------------------------------------------------------------------------
// shared layout per group: [R x SA0 : approximations of odd levels][R x SA1 : even levels][R x SD : details of every level]
template <int F>
__global__ void __launch_bounds__(NT)
k_row_inv(const float* __restrict__ A, float* __restrict__ out, const __grid_constant__ RowLevels lv,
          const __grid_constant__ PwtTapsInv tp, int rows, int R, int SA0, int SA1, int SD) {
    using G = RowGeo<F>;
    constexpr int NP = G::NP;
    extern __shared__ __align__(16) float sm[];
    float* const bufA0 = sm;                                // a_l of odd l
    float* const bufA1 = sm + (size_t)R * SA0;              // a_l of even l
    float* bufD = sm + (size_t)R * (SA0 + SA1);
    const int tid = threadIdx.x;
    const int L = lv.L;
    const int ngroups = (rows + R - 1) / R;
    pwt_pdl_wait();
    for (int g = blockIdx.x; g < ngroups; g += gridDim.x) {
        const int r0 = g * R, nr = rows - r0 < R ? rows - r0 : R;
        if (g + (int)gridDim.x >= ngroups) pwt_pdl_trigger();
        {
            stage_rows((L & 1) ? bufA0 : bufA1, (L & 1) ? SA0 : SA1, G::HL, A + (size_t)r0 * lv.n[L], nr, lv.n[L], tid);
            for (int l = 1; l <= L; l++)
                stage_rows(bufD + lv.dofs[l - 1], SD, G::HL, lv.D[l - 1] + (size_t)r0 * lv.n[l], nr, lv.n[l], tid);
            cp_async_commit();
            cp_async_wait_all();
            __syncthreads();
            for (int l = 1; l <= L; l++) fill_halo_inv<G::HL>(bufD + lv.dofs[l - 1], SD, nr, lv.n[l], tid);
        }
        for (int l = L; l >= 1; l--) {
            const bool odd = l & 1;
            const int n2 = lv.n[l], nout = lv.n[l - 1];
            float* ca = odd ? bufA0 : bufA1;
            const int Sa = odd ? SA0 : SA1;
            fill_halo_inv<G::HL>(ca, Sa, nr, n2, tid);
            __syncthreads();
            const float* cd = bufD + lv.dofs[l - 1];
            float* na = odd ? bufA1 : bufA0;
            const int Sn = odd ? SA1 : SA0;
            float* Og = out + (size_t)r0 * nout;
            const bool vec = (nout & 3) == 0 && (((uintptr_t)Og) & 15) == 0;
            const int nch = (n2 + NP - 1) / NP;
            const int tb = log2_ceil_cap(nch), T = 1 << tb;
            for (int r = tid >> tb; r < nr; r += NT >> tb) {
                const float* arow = ca + r * Sa;
                const float* drow = cd + r * SD;
                float* orow = Og + (size_t)r * nout;
                float* nrow = na + r * Sn;
                for (int j0 = NP * (tid & (T - 1)); j0 < n2; j0 += NP * T) {
                    float4 o0, o1;                                                  // outputs 2 j0 .. 2 j0 + 7
                    if (F == 2) {                                                   // haar.cu:146-160
                        const float c = 0.70710678118654746f;
                        const float4 a = lds4(arow, (G::HL + j0) >> 2), d = lds4(drow, (G::HL + j0) >> 2);
                        o0 = make_float4(c * (a.x + d.x), c * (a.x - d.x), c * (a.y + d.y), c * (a.y - d.y));
                        o1 = make_float4(c * (a.z + d.z), c * (a.z - d.z), c * (a.w + d.w), c * (a.w - d.w));
                    } else {
                        const int q0 = (G::HL + j0 - G::S1 - G::OFFI) >> 2;
                        float wa[4 * G::NVI], wd[4 * G::NVI];
#pragma unroll
                        for (int q = 0; q < G::NVI; q++) {
                            const float4 t = lds4(arow, q0 + q), u = lds4(drow, q0 + q);
                            wa[4 * q] = t.x; wa[4 * q + 1] = t.y; wa[4 * q + 2] = t.z; wa[4 * q + 3] = t.w;
                            wd[4 * q] = u.x; wd[4 * q + 1] = u.y; wd[4 * q + 2] = u.z; wd[4 * q + 3] = u.w;
                        }
                        float2 e0 = make_float2(0.f, 0.f), e1 = e0, e2 = e0, e3 = e0; // (even, odd) outputs of pairs j0 .. j0 + 3
#pragma unroll
                        for (int w = 0; w < G::W; w++) {
                            e0 = fma2s(wa[G::OFFI + w], tp.l[w], e0);     e0 = fma2s(wd[G::OFFI + w], tp.h[w], e0);
                            e1 = fma2s(wa[G::OFFI + 1 + w], tp.l[w], e1); e1 = fma2s(wd[G::OFFI + 1 + w], tp.h[w], e1);
                            e2 = fma2s(wa[G::OFFI + 2 + w], tp.l[w], e2); e2 = fma2s(wd[G::OFFI + 2 + w], tp.h[w], e2);
                            e3 = fma2s(wa[G::OFFI + 3 + w], tp.l[w], e3); e3 = fma2s(wd[G::OFFI + 3 + w], tp.h[w], e3);
                        }
                        o0 = make_float4(e0.x, e0.y, e1.x, e1.y);
                        o1 = make_float4(e2.x, e2.y, e3.x, e3.y);
                    }
                    if (l == 1) {
                        store4(orow, 2 * j0, nout, o0, vec);
                        store4(orow, 2 * j0 + 4, nout, o1, vec && 2 * j0 + 4 < nout);
                    } else {                               // (values past nout land in the halo / slack and are overwritten)
                        sts4(nrow, (G::HL + 2 * j0) >> 2, o0);
                        sts4(nrow, ((G::HL + 2 * j0) >> 2) + 1, o1);
                    }
                }
            }
            __syncthreads();
        }
    }
}

// =====================================================================================================
// Batched 1D stationary (a trous) transform, every level in one launch (reference: w_forward_swt_separable_1d
// separable.cu:519-537, w_inverse_swt_separable_1d :653-672: one launch per level, 12 B/px each; here 4 + 4 (L + 1) B/px
// per direction).  A row is staged once; level l filters it with dilation s = 2^(l-1) and period N (no odd
// extension): out[g] = sum_j f[F-1-j] * in[(g + (j - c) s) mod N], c = F/2 - 1 (analysis), c = F/2 and taps / 2
// (synthesis, separable.cu:553-626).  The approximation ping-pongs between two shared rows, detail rows go to / come
// from global memory.  4 adjacent outputs per thread: dilations that are multiples of 4 read one aligned 128-bit
// group per tap, s = 1 and s = 2 read the contiguous window once; wrap-around per 128-bit group (N % 4 == 0).
// =====================================================================================================
struct SwtLevels {
    float* D[PWT_MAX_LEVELS];
    int L;
};
struct TapsDup1d {                // synthesis taps halved and duplicated: l[j] = (IL[F-1-j] / 2, same), h likewise
    float2 l[PWT_MAX_TAPS];
    float2 h[PWT_MAX_TAPS];
};
__device__ __forceinline__ int wrapq(int q, int nq) {          // -nq <= q < 2 nq
    if (q < 0) q += nq;
    else if (q >= nq) q -= nq;
    return q;
}

template <int F, int SM>
__device__ __forceinline__ void swt_fwd_level(const float* __restrict__ cur, float* __restrict__ nxt, int pitch, float* __restrict__ Dg,
                                              float* __restrict__ Ag, bool last, int nr, int N, int s, const PwtTapsFwd& tp, int tid) {
    constexpr int C = F / 2 - 1;
    const int nq = N >> 2;
    const int tb = log2_ceil_cap(nq), T = 1 << tb;
    for (int r = tid >> tb; r < nr; r += NT >> tb) {
        const float* crow = cur + r * pitch;
        for (int g4 = tid & (T - 1); g4 < nq; g4 += T) {
            float2 p0 = make_float2(0.f, 0.f), p1 = p0, p2 = p0, p3 = p0;
            if (SM == 0) {
                const int sq = s >> 2;
                int q = mod_pos(g4 - C * sq, nq);
#pragma unroll
                for (int j = 0; j < F; j++) {
                    const float4 x = lds4(crow, q);
                    p0 = fma2s(x.x, tp.t[j], p0);
                    p1 = fma2s(x.y, tp.t[j], p1);
                    p2 = fma2s(x.z, tp.t[j], p2);
                    p3 = fma2s(x.w, tp.t[j], p3);
                    q += sq;
                    if (q >= nq) q -= nq;
                }
            } else {
                constexpr int S = SM;
                constexpr int OFF = (4 - (C * S) % 4) % 4;              // the window starts OFF floats into its first group
                constexpr int NV = (OFF + 4 + (F - 1) * S + 3) / 4;
                int q = mod_pos(g4 - (C * S + OFF) / 4, nq);
                float w[4 * NV];
#pragma unroll
                for (int k = 0; k < NV; k++) {
                    const float4 x = lds4(crow, q);
                    w[4 * k] = x.x; w[4 * k + 1] = x.y; w[4 * k + 2] = x.z; w[4 * k + 3] = x.w;
                    q = q + 1 == nq ? 0 : q + 1;
                }
#pragma unroll
                for (int j = 0; j < F; j++) {
                    p0 = fma2s(w[OFF + j * S], tp.t[j], p0);
                    p1 = fma2s(w[OFF + 1 + j * S], tp.t[j], p1);
                    p2 = fma2s(w[OFF + 2 + j * S], tp.t[j], p2);
                    p3 = fma2s(w[OFF + 3 + j * S], tp.t[j], p3);
                }
            }
            const size_t o = (size_t)r * N + 4 * g4;
            stg_cs(reinterpret_cast<float4*>(Dg + o), make_float4(p0.y, p1.y, p2.y, p3.y));
            const float4 lo = make_float4(p0.x, p1.x, p2.x, p3.x);
            if (last) stg_cs(reinterpret_cast<float4*>(Ag + o), lo);
            else sts4(nxt + r * pitch, g4, lo);
        }
    }
}

template <int F>
__global__ void __launch_bounds__(NT)
k_row_swt_fwd(const float* __restrict__ in, float* __restrict__ A, const __grid_constant__ SwtLevels lv,
              const __grid_constant__ PwtTapsFwd tp, int rows, int N, int R, int pitch) {
    extern __shared__ __align__(16) float sm[];
    float* buf0 = sm;
    float* buf1 = sm + (size_t)R * pitch;
    const int tid = threadIdx.x, L = lv.L;
    const int ngroups = (rows + R - 1) / R;
    pwt_pdl_wait();
    for (int g = blockIdx.x; g < ngroups; g += gridDim.x) {
        const int r0 = g * R, nr = rows - r0 < R ? rows - r0 : R;
        if (g + (int)gridDim.x >= ngroups) pwt_pdl_trigger();
        stage_rows(buf0, pitch, 0, in + (size_t)r0 * N, nr, N, tid);
        cp_async_commit();
        cp_async_wait_all();
        __syncthreads();
        float* cur = buf0;
        float* nxt = buf1;
        for (int l = 1; l <= L; l++) {
            float* Dg = lv.D[l - 1] + (size_t)r0 * N;
            float* Ag = A + (size_t)r0 * N;
            if (l == 1) swt_fwd_level<F, 1>(cur, nxt, pitch, Dg, Ag, l == L, nr, N, 1, tp, tid);
            else if (l == 2) swt_fwd_level<F, 2>(cur, nxt, pitch, Dg, Ag, l == L, nr, N, 2, tp, tid);
            else swt_fwd_level<F, 0>(cur, nxt, pitch, Dg, Ag, l == L, nr, N, 1 << (l - 1), tp, tid);
            __syncthreads();
            float* t = cur; cur = nxt; nxt = t;
        }
    }
}

template <int F, int SM>
__device__ __forceinline__ void swt_inv_level(const float* __restrict__ ca, const float* __restrict__ cd, float* __restrict__ na, int pitch,
                                              float* __restrict__ Og, bool final, int nr, int N, int s, const TapsDup1d& tp, int tid) {
    constexpr int C = F / 2;
    const int nq = N >> 2;
    const int tb = log2_ceil_cap(nq), T = 1 << tb;
    for (int r = tid >> tb; r < nr; r += NT >> tb) {
        const float* arow = ca + r * pitch;
        const float* drow = cd + r * pitch;
        for (int g4 = tid & (T - 1); g4 < nq; g4 += T) {
            float2 e01 = make_float2(0.f, 0.f), e23 = e01;             // outputs 4 g4 .. 4 g4 + 3
            if (SM == 0) {
                const int sq = s >> 2;
                int q = mod_pos(g4 - C * sq, nq);
#pragma unroll
                for (int j = 0; j < F; j++) {
                    const float4 x = lds4(arow, q), y = lds4(drow, q);
                    e01 = __ffma2_rn(make_float2(x.x, x.y), tp.l[j], e01);
                    e23 = __ffma2_rn(make_float2(x.z, x.w), tp.l[j], e23);
                    e01 = __ffma2_rn(make_float2(y.x, y.y), tp.h[j], e01);
                    e23 = __ffma2_rn(make_float2(y.z, y.w), tp.h[j], e23);
                    q += sq;
                    if (q >= nq) q -= nq;
                }
            } else {
                constexpr int S = SM;
                constexpr int OFF = (4 - (C * S) % 4) % 4;
                constexpr int NV = (OFF + 4 + (F - 1) * S + 3) / 4;
                int q = mod_pos(g4 - (C * S + OFF) / 4, nq);
                float wa[4 * NV], wd[4 * NV];
#pragma unroll
                for (int k = 0; k < NV; k++) {
                    const float4 x = lds4(arow, q), y = lds4(drow, q);
                    wa[4 * k] = x.x; wa[4 * k + 1] = x.y; wa[4 * k + 2] = x.z; wa[4 * k + 3] = x.w;
                    wd[4 * k] = y.x; wd[4 * k + 1] = y.y; wd[4 * k + 2] = y.z; wd[4 * k + 3] = y.w;
                    q = q + 1 == nq ? 0 : q + 1;
                }
#pragma unroll
                for (int j = 0; j < F; j++) {
                    e01 = __ffma2_rn(make_float2(wa[OFF + j * S], wa[OFF + 1 + j * S]), tp.l[j], e01);
                    e23 = __ffma2_rn(make_float2(wa[OFF + 2 + j * S], wa[OFF + 3 + j * S]), tp.l[j], e23);
                    e01 = __ffma2_rn(make_float2(wd[OFF + j * S], wd[OFF + 1 + j * S]), tp.h[j], e01);
                    e23 = __ffma2_rn(make_float2(wd[OFF + 2 + j * S], wd[OFF + 3 + j * S]), tp.h[j], e23);
                }
            }
            const float4 o = make_float4(e01.x, e01.y, e23.x, e23.y);
            if (final) stg_cs(reinterpret_cast<float4*>(Og + (size_t)r * N + 4 * g4), o);
            else sts4(na + r * pitch, g4, o);
        }
    }
}

// shared layout: [R x pitch : a][R x pitch : a'][R x pitch : details of the level being synthesised]
template <int F>
__global__ void __launch_bounds__(NT)
k_row_swt_inv(const float* __restrict__ A, float* __restrict__ out, const __grid_constant__ SwtLevels lv,
              const __grid_constant__ TapsDup1d tp, int rows, int N, int R, int pitch) {
    extern __shared__ __align__(16) float sm[];
    float* bufA = sm;
    float* bufB = sm + (size_t)R * pitch;
    float* bufD = sm + (size_t)2 * R * pitch;
    const int tid = threadIdx.x, L = lv.L;
    const int ngroups = (rows + R - 1) / R;
    pwt_pdl_wait();
    for (int g = blockIdx.x; g < ngroups; g += gridDim.x) {
        const int r0 = g * R, nr = rows - r0 < R ? rows - r0 : R;
        if (g + (int)gridDim.x >= ngroups) pwt_pdl_trigger();
        stage_rows(bufA, pitch, 0, A + (size_t)r0 * N, nr, N, tid);
        float* cur = bufA;
        float* nxt = bufB;
        for (int l = L; l >= 1; l--) {
            stage_rows(bufD, pitch, 0, lv.D[l - 1] + (size_t)r0 * N, nr, N, tid);
            cp_async_commit();
            cp_async_wait_all();
            __syncthreads();
            float* Og = out + (size_t)r0 * N;
            if (l == 1) swt_inv_level<F, 1>(cur, bufD, nxt, pitch, Og, true, nr, N, 1, tp, tid);
            else if (l == 2) swt_inv_level<F, 2>(cur, bufD, nxt, pitch, Og, false, nr, N, 2, tp, tid);
            else swt_inv_level<F, 0>(cur, bufD, nxt, pitch, Og, false, nr, N, 1 << (l - 1), tp, tid);
            __syncthreads();
            float* t = cur; cur = nxt; nxt = t;
        }
    }
}

// rows per CTA and launch geometry of the a-trous rows kernels (nbuf shared rows per staged row)
struct SwtPlan1d {
    int R, pitch, grid, ok;
    size_t smem;
};
inline SwtPlan1d make_swt_plan(int rows, int N, int L, int F, int nbuf) {
    SwtPlan1d pl = {};
    if ((N & 3) || N < 8 || L < 1 || L > 30) return pl;
    if ((long long)(F - 1) * (1LL << (L - 1)) >= N) return pl;         // one wrap must cover the reach (always true after level clipping)
    pl.pitch = N + 8;
    const size_t per_row = sizeof(float) * (size_t)nbuf * pl.pitch;
    const size_t budget = (nbuf == 2 ? 72 : 104) * 1024;              // 3 / 2 CTAs per SM
    int R = (int)(budget / per_row);
    if (R < 1) {
        if (per_row > 200 * 1024) return pl;
        R = 1;
    }
    const int want = (4096 + N - 1) / N;
    if (R > want) R = want;
    if (R > rows) R = rows;
    pl.R = R < 1 ? 1 : R;
    pl.smem = per_row * pl.R;
    const int ngroups = (rows + pl.R - 1) / pl.R;
    int per_sm = (int)((220 * 1024) / (pl.smem + 1024));
    per_sm = per_sm < 1 ? 1 : (per_sm > 8 ? 8 : per_sm);
    const int cap = pwt_sm_count() * per_sm;
    pl.grid = ngroups < cap ? ngroups : cap;
    pl.ok = 1;
    return pl;
}
template <int F>
int launch_row_swt_fwd(const float* in, float* A, float* const* D, int rows, int N, int L, const PwtFilters& f, cudaStream_t st) {
    const SwtPlan1d pl = make_swt_plan(rows, N, L, F, 2);
    if (!pl.ok || ((((uintptr_t)in) | ((uintptr_t)A)) & 15)) return 0;
    SwtLevels lv = {};
    lv.L = L;
    for (int l = 0; l < L; l++) {
        if (((uintptr_t)D[l]) & 15) return 0;
        lv.D[l] = D[l];
    }
    static PwtKernelOnce once;
    if (!pwt_kernel_once(once, k_row_swt_fwd<F>, NT, 200 * 1024, 72 * 1024)) return 0;
    const PwtTapsFwd t = pwt_pack_taps_fwd(f, F);
    pwt_launch_pdl(k_row_swt_fwd<F>, dim3((unsigned)pl.grid), NT, pl.smem, st, in, A, lv, t, rows, N, pl.R, pl.pitch);
    return 1;
}
template <int F>
int launch_row_swt_inv(const float* A, float* const* D, float* out, int rows, int N, int L, const PwtFilters& f, cudaStream_t st) {
    const SwtPlan1d pl = make_swt_plan(rows, N, L, F, 3);
    if (!pl.ok || ((((uintptr_t)out) | ((uintptr_t)A)) & 15)) return 0;
    SwtLevels lv = {};
    lv.L = L;
    for (int l = 0; l < L; l++) {
        if (((uintptr_t)D[l]) & 15) return 0;
        lv.D[l] = D[l];
    }
    static PwtKernelOnce once;
    if (!pwt_kernel_once(once, k_row_swt_inv<F>, NT, 200 * 1024, 104 * 1024)) return 0;
    TapsDup1d t;
    for (int j = 0; j < PWT_MAX_TAPS; j++) {
        const float l = j < F ? 0.5f * f.IL[F - 1 - j] : 0.f, h = j < F ? 0.5f * f.IH[F - 1 - j] : 0.f;
        t.l[j] = make_float2(l, l);
        t.h[j] = make_float2(h, h);
    }
    pwt_launch_pdl(k_row_swt_inv<F>, dim3((unsigned)pl.grid), NT, pl.smem, st, A, out, lv, t, rows, N, pl.R, pl.pitch);
    return 1;
}

struct Plan1d {
    RowLevels lv;
    int R, S0, S1, SD;
    size_t smem;
    int ok;
};
// rows per group: enough independent 4-wide chunks for the 256 threads at level 1, within the shared memory budget
template <int F>
Plan1d make_plan(int rows, int Nc, int L, bool inverse) {
    using G = RowGeo<F>;
    Plan1d pl = {};
    pl.lv.L = L;
    pl.lv.n[0] = Nc;
    for (int l = 1; l <= L; l++) pl.lv.n[l] = (pl.lv.n[l - 1] + 1) >> 1;
    for (int l = 1; l <= L; l++)
        if (pl.lv.n[l - 1] < 2) return pl;
    int sd = 0;
    for (int l = 1; l <= L; l++) {
        pl.lv.dofs[l - 1] = sd;
        sd += row_pitch(pl.lv.n[l], G::HL);
    }
    if (!inverse) {
        pl.S0 = row_pitch(Nc, G::HL);
        pl.S1 = L >= 2 ? row_pitch(pl.lv.n[1], G::HL) : 0;
        pl.SD = 0;
    } else {                                               // S0: a_l of odd l (largest n[1]), S1: even l (largest n[2])
        pl.S0 = row_pitch(pl.lv.n[1], G::HL);
        pl.S1 = L >= 2 ? row_pitch(pl.lv.n[2], G::HL) : 0;
        pl.SD = sd;
    }
    const size_t per_row = sizeof(float) * (size_t)(pl.S0 + pl.S1 + pl.SD);
    const size_t budget = 56 * 1024;                       // 4 CTAs per SM
    const size_t hard = 200 * 1024;
    int R = (int)(budget / per_row);
    if (R < 1) {
        if (per_row > hard) return pl;                     // row too long for one CTA: per-level kernels
        R = 1;
    }
    const int want = (4096 + Nc - 1) / Nc;                 // >= 4096 level-1 inputs per group keep every thread busy
    if (R > want) R = want;
    if (R > rows) R = rows;
    if (R < 1) R = 1;
    pl.R = R;
    pl.smem = per_row * R;
    pl.ok = 1;
    return pl;
}

template <int F>
int launch_row_fwd(const float* in, float* A, float* const* D, int rows, int Nc, int L, const PwtFilters& f, cudaStream_t st) {
    Plan1d pl = make_plan<F>(rows, Nc, L, false);
    if (!pl.ok) return 0;
    for (int l = 0; l < L; l++) pl.lv.D[l] = D[l];
    static PwtKernelOnce once;
    if (!pwt_kernel_once(once, k_row_fwd<F>, NT, 200 * 1024, 56 * 1024)) return 0;
    const int ngroups = (rows + pl.R - 1) / pl.R;
    int per_sm = (int)((220 * 1024) / (pl.smem + 1024));
    per_sm = per_sm < 1 ? 1 : (per_sm > 8 ? 8 : per_sm);
    const int cap = pwt_sm_count() * per_sm;
    const PwtTapsFwd t = pwt_pack_taps_fwd(f, F);
    pwt_launch_pdl(k_row_fwd<F>, dim3((unsigned)(ngroups < cap ? ngroups : cap)), NT, pl.smem, st, in, A, pl.lv, t, rows, pl.R, pl.S0, pl.S1);
    return 1;
}
template <int F>
int launch_row_inv(const float* A, float* const* D, float* out, int rows, int Nc, int L, const PwtFilters& f, cudaStream_t st) {
    Plan1d pl = make_plan<F>(rows, Nc, L, true);
    if (!pl.ok) return 0;
    for (int l = 0; l < L; l++) pl.lv.D[l] = D[l];
    static PwtKernelOnce once;
    if (!pwt_kernel_once(once, k_row_inv<F>, NT, 200 * 1024, 56 * 1024)) return 0;
    const int ngroups = (rows + pl.R - 1) / pl.R;
    int per_sm = (int)((220 * 1024) / (pl.smem + 1024));
    per_sm = per_sm < 1 ? 1 : (per_sm > 8 ? 8 : per_sm);
    const int cap = pwt_sm_count() * per_sm;
    const PwtTapsInv t = pwt_pack_taps_inv(f, F);
    pwt_launch_pdl(k_row_inv<F>, dim3((unsigned)(ngroups < cap ? ngroups : cap)), NT, pl.smem, st, A, out, pl.lv, t, rows, pl.R, pl.S0, pl.S1, pl.SD);
    return 1;
}
}  // namespace

#define PWT_ROWSWT_CASES(X) X(2) X(4) X(6) X(8) X(10) X(12) X(14) X(16) X(18) X(20)
#define PWT_ROW1D_CASES(X) X(2) X(4) X(6) X(8) X(10) X(12) X(14) X(16) X(18) X(20) X(22) X(24) X(26) X(28) X(30) X(32) X(34) X(36) X(38) X(40)

// D[l]: detail band of level l + 1 (rows x ceil(Nc / 2^(l+1))); A: rows x n[L].  Returns 0 when not covered.
int pwt_row_dwt_fwd1d_all(const float* in, float* A, float* const* D, int rows, int Nc, int L, const PwtFilters& f, cudaStream_t st) {
    if (rows < 1 || Nc < 2 || L < 1 || L > PWT_MAX_LEVELS || (long long)rows * Nc >= (1LL << 40)) return 0;
    switch (f.hlen) {
#define X(FF) case FF: return launch_row_fwd<FF>(in, A, D, rows, Nc, L, f, st);
        PWT_ROW1D_CASES(X)
#undef X
        default: return 0;
    }
}
int pwt_row_dwt_inv1d_all(const float* A, float* const* D, float* out, int rows, int Nc, int L, const PwtFilters& f, cudaStream_t st) {
    if (rows < 1 || Nc < 2 || L < 1 || L > PWT_MAX_LEVELS || (long long)rows * Nc >= (1LL << 40)) return 0;
    switch (f.hlen) {
#define X(FF) case FF: return launch_row_inv<FF>(A, D, out, rows, Nc, L, f, st);
        PWT_ROW1D_CASES(X)
#undef X
        default: return 0;
    }
}

// batched 1D a-trous transform, every level in one launch.  D[l]: detail band of level l + 1 (rows x Nc).  Return 0 when
// not covered (width not a multiple of 4, row too long for a CTA's shared memory, filter longer than 20 taps at s = 2).
int pwt_row_swt_fwd1d_all(const float* in, float* A, float* const* D, int rows, int Nc, int L, const PwtFilters& f, cudaStream_t st) {
    if (rows < 1 || L < 1 || L > PWT_MAX_LEVELS || (long long)rows * Nc >= (1LL << 40)) return 0;
    switch (f.hlen) {
#define X(FF) case FF: return launch_row_swt_fwd<FF>(in, A, D, rows, Nc, L, f, st);
        PWT_ROWSWT_CASES(X)
#undef X
        default: return 0;
    }
}
int pwt_row_swt_inv1d_all(const float* A, float* const* D, float* out, int rows, int Nc, int L, const PwtFilters& f, cudaStream_t st) {
    if (rows < 1 || L < 1 || L > PWT_MAX_LEVELS || (long long)rows * Nc >= (1LL << 40)) return 0;
    switch (f.hlen) {
#define X(FF) case FF: return launch_row_swt_inv<FF>(A, D, out, rows, Nc, L, f, st);
        PWT_ROWSWT_CASES(X)
#undef X
        default: return 0;
    }
}
