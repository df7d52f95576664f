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
