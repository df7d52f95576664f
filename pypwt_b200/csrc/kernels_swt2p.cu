// Stationary (a trous) 2D level as two streaming passes -- the fallback for what the fused SWT kernels do not cover
// (filters longer than 16 taps, widths that are not multiples of 4).  Reference: w_kern_forward_swt_pass1/2
// separable.cu:409-493, w_kern_inverse_swt_pass1/2 :553-626 (also two passes, one thread per output, every tap a global
// load).  The first fallback here worked the same way and fell off a cliff: 4096^2 db10 3 levels 5.6 ms against 0.58 ms
// for sym8 on the fused kernels.  These passes read every sample once:
//   rows    a CTA stages the inputs of TO consecutive outputs of a row (+ the (F-1) s reach, periodic wrap resolved while
//           staging) with coalesced loads; thread o reads sx[o + j s]: consecutive threads, consecutive addresses for any s;
//   columns a thread owns VEC adjacent columns and walks one residue class of rows y = r + q s, so the dilated filter is
//           an ordinary one over q: analysis keeps the F rows of the window in rotating registers (one new row per output),
//           synthesis scatters each new row of the two bands into F rotating accumulators (transposed form); rotations
//           unrolled over their period, every register index static.
// 36 B/px per level and direction (20 compulsory), at streaming speed.
#include <stdlib.h>

#include "pwt_internal.h"

namespace {
__device__ __forceinline__ float2 fma2s(float x, float2 t, float2 acc) { return __ffma2_rn(make_float2(x, x), t, acc); }
__device__ __forceinline__ int mod_pos(int i, int n) {
    i %= n;
    return i < 0 ? i + n : i;
}
__host__ __device__ inline int cdiv(int a, int b) { return (a + b - 1) / b; }

// ---- rows ------------------------------------------------------------------------------------------------------------
constexpr int kTO = 2048;                                  // outputs per tile
// analysis: out[g] = sum_j f[F-1-j] in[(g + (j - c) s) mod Nc], c = F/2 - 1  ->  lo, hi
template <int F>
__global__ void __launch_bounds__(256)
k_swt2p_rows_fwd(const float* __restrict__ in, float* __restrict__ lo, float* __restrict__ hi, long long rows, int Nc, int s,
                 const __grid_constant__ PwtTapsFwd tp) {
    extern __shared__ float sx[];
    const int c = F / 2 - 1, reach = (F - 1) * s, ntile = cdiv(Nc, kTO);
    pwt_pdl_wait();
    for (long long t = blockIdx.x; t < rows * ntile; t += gridDim.x) {
        const long long r = t / ntile;
        const int g0 = (int)(t - r * ntile) * kTO, x0 = g0 - c * s;
        const int nout = min(kTO, Nc - g0), nin = nout + reach;
        const float* row = in + r * Nc;
        if (x0 >= 0 && x0 + nin <= Nc) {
            for (int i = threadIdx.x; i < nin; i += 256) sx[i] = __ldg(row + x0 + i);
        } else {
            int xi = mod_pos(x0 + (int)threadIdx.x, Nc);
            const int step = 256 % Nc;
            for (int i = threadIdx.x; i < nin; i += 256) {
                sx[i] = __ldg(row + xi);
                xi += step;
                if (xi >= Nc) xi -= Nc;
            }
        }
        __syncthreads();
        for (int o = threadIdx.x; o < nout; o += 256) {
            float2 p = make_float2(0.f, 0.f);
#pragma unroll
            for (int j = 0; j < F; j++) p = fma2s(sx[o + j * s], tp.t[j], p);
            lo[r * Nc + g0 + o] = p.x;
            hi[r * Nc + g0 + o] = p.y;
        }
        __syncthreads();
    }
}
// synthesis: out[g] = sum_j (IL[F-1-j] / 2) t1[(g + (j - F/2) s) mod Nc] + (IH[F-1-j] / 2) t2[...]
struct TapsHalf {
    float l[PWT_MAX_TAPS], h[PWT_MAX_TAPS];
};
template <int F>
__global__ void __launch_bounds__(256)
k_swt2p_rows_inv(const float* __restrict__ t1, const float* __restrict__ t2, float* __restrict__ out, long long rows, int Nc, int s,
                 const __grid_constant__ TapsHalf tp) {
    extern __shared__ float sx[];
    const int c = F / 2, reach = (F - 1) * s, ntile = cdiv(Nc, kTO);
    float* sa = sx;
    float* sd = sx + kTO + reach;
    pwt_pdl_wait();
    for (long long t = blockIdx.x; t < rows * ntile; t += gridDim.x) {
        const long long r = t / ntile;
        const int g0 = (int)(t - r * ntile) * kTO, x0 = g0 - c * s;
        const int nout = min(kTO, Nc - g0), nin = nout + reach;
        const float* a = t1 + r * Nc;
        const float* d = t2 + r * Nc;
        int xi = mod_pos(x0 + (int)threadIdx.x, Nc);
        const int step = 256 % Nc;
        for (int i = threadIdx.x; i < nin; i += 256) {
            sa[i] = __ldg(a + xi);
            sd[i] = __ldg(d + xi);
            xi += step;
            if (xi >= Nc) xi -= Nc;
        }
        __syncthreads();
        for (int o = threadIdx.x; o < nout; o += 256) {
            float x = 0.f;
#pragma unroll
            for (int j = 0; j < F; j++) {
                x = fmaf(sa[o + j * s], tp.l[j], x);
                x = fmaf(sd[o + j * s], tp.h[j], x);
            }
            out[r * Nc + g0 + o] = x;
        }
        __syncthreads();
    }
}

// ---- columns -----------------------------------------------------------------------------------------------------------
template <int VEC>
__device__ __forceinline__ void ldv(float (&w)[VEC], const float* p) {
    if (VEC == 4) { const float4 v = __ldg(reinterpret_cast<const float4*>(p)); w[0] = v.x; w[VEC > 1 ? 1 : 0] = v.y; w[VEC > 2 ? 2 : 0] = v.z; w[VEC > 3 ? 3 : 0] = v.w; }
    else if (VEC == 2) { const float2 v = __ldg(reinterpret_cast<const float2*>(p)); w[0] = v.x; w[VEC > 1 ? 1 : 0] = v.y; }
    else w[0] = __ldg(p);
}
template <int VEC>
__device__ __forceinline__ void stv(float* p, const float (&w)[VEC]) {
    if (VEC == 4) __stcs(reinterpret_cast<float4*>(p), make_float4(w[0], w[VEC > 1 ? 1 : 0], w[VEC > 2 ? 2 : 0], w[VEC > 3 ? 3 : 0]));
    else if (VEC == 2) __stcs(reinterpret_cast<float2*>(p), make_float2(w[0], w[VEC > 1 ? 1 : 0]));
    else p[0] = w[0];
}
struct ColJobs {
    const float* a[2];     // analysis: lo / hi plane;  synthesis: A / V
    const float* b[2];     // synthesis: H / D
    float* o0[2];          // analysis: A / V;  synthesis: t1 / t2
    float* o1[2];          // analysis: H / D
};
// work item -> (residue class r, run of KS lattice positions, column group); blockIdx.y = job, blockIdx.z = image
template <int F, int VEC>
__global__ void __launch_bounds__(128)
k_swt2p_cols_fwd(const __grid_constant__ ColJobs jb, int Nr, int Nc, int s, int KS, long long plane, const __grid_constant__ PwtTapsFwd tp) {
    constexpr int C = F / 2 - 1;
    const float* __restrict__ in = jb.a[blockIdx.y] + blockIdx.z * plane;
    float* __restrict__ o0 = jb.o0[blockIdx.y] + blockIdx.z * plane;
    float* __restrict__ o1 = jb.o1[blockIdx.y] + blockIdx.z * plane;
    const int PV = Nc / VEC, nqmax = (Nr + s - 1) / s, nseg = (nqmax + KS - 1) / KS;
    pwt_pdl_wait();
    for (long long i = blockIdx.x * 128LL + threadIdx.x; i < (long long)PV * nseg * s; i += gridDim.x * 128LL) {
        const int p = (int)(i % PV) * VEC;
        const int rs = (int)(i / PV), r = rs % s, seg = rs / s;
        const int nq = (Nr - r + s - 1) / s;
        const int q0 = seg * KS, qend = min(q0 + KS, nq);
        if (q0 >= qend) continue;
        float w[F][VEC];
        int y = mod_pos(r + (q0 - C) * s, Nr);                 // row of window position 0
#pragma unroll
        for (int j = 0; j < F - 1; j++) {
            ldv<VEC>(w[j], in + (long long)y * Nc + p);
            y += s;
            if (y >= Nr) y -= Nr;
        }
        for (int qb = q0; qb < qend; qb += F) {
#pragma unroll
            for (int u = 0; u < F; u++) {
                const int q = qb + u;
                if (q < qend) {
                    ldv<VEC>(w[(F - 1 + u) % F], in + (long long)y * Nc + p);
                    y += s;
                    if (y >= Nr) y -= Nr;
                    float2 acc[VEC];
#pragma unroll
                    for (int v = 0; v < VEC; v++) acc[v] = make_float2(0.f, 0.f);
#pragma unroll
                    for (int j = 0; j < F; j++)
#pragma unroll
                        for (int v = 0; v < VEC; v++) acc[v] = fma2s(w[(j + u) % F][v], tp.t[j], acc[v]);
                    float a[VEC], d[VEC];
#pragma unroll
                    for (int v = 0; v < VEC; v++) { a[v] = acc[v].x; d[v] = acc[v].y; }
                    const long long o = (long long)(r + q * s) * Nc + p;
                    stv<VEC>(o0 + o, a);
                    stv<VEC>(o1 + o, d);
                }
            }
        }
    }
}
// synthesis, transposed form: stream row k (bands a and d at row (r + k s) mod Nr) adds tap j to output q = k - j + C;
// output q completes with j = F - 1.  slot(q) = (q - q0) mod F.
template <int F, int VEC>
__global__ void __launch_bounds__(128)
k_swt2p_cols_inv(const __grid_constant__ ColJobs jb, int Nr, int Nc, int s, int KS, long long plane, const __grid_constant__ TapsHalf tp) {
    constexpr int C = F / 2;
    const float* __restrict__ A = jb.a[blockIdx.y] + blockIdx.z * plane;
    const float* __restrict__ B = jb.b[blockIdx.y] + blockIdx.z * plane;
    float* __restrict__ out = jb.o0[blockIdx.y] + blockIdx.z * plane;
    const int PV = Nc / VEC, nqmax = (Nr + s - 1) / s, nseg = (nqmax + KS - 1) / KS;
    pwt_pdl_wait();
    for (long long i = blockIdx.x * 128LL + threadIdx.x; i < (long long)PV * nseg * s; i += gridDim.x * 128LL) {
        const int p = (int)(i % PV) * VEC;
        const int rs = (int)(i / PV), r = rs % s, seg = rs / s;
        const int nq = (Nr - r + s - 1) / s;
        const int q0 = seg * KS, qend = min(q0 + KS, nq);
        if (q0 >= qend) continue;
        float acc[F][VEC];
        int y = mod_pos(r + (q0 - C) * s, Nr);                 // stream step t = 0 is lattice position k = q0 - C
        const int nsteps = (qend - q0) + F - 1;
        for (int tb = 0; tb < nsteps; tb += F) {
#pragma unroll
            for (int u = 0; u < F; u++) {
                const int t = tb + u;
                if (t < nsteps) {
                    float xa[VEC], xd[VEC];
                    ldv<VEC>(xa, A + (long long)y * Nc + p);
                    ldv<VEC>(xd, B + (long long)y * Nc + p);
                    y += s;
                    if (y >= Nr) y -= Nr;
#pragma unroll
                    for (int j = 0; j < F; j++)                 // output q0 + t - j, slot (u - j) mod F; j = 0 opens the slot
#pragma unroll
                        for (int v = 0; v < VEC; v++) {
                            const float base = j == 0 ? 0.f : acc[((u - j) % F + F) % F][v];
                            acc[((u - j) % F + F) % F][v] = fmaf(xd[v], tp.h[j], fmaf(xa[v], tp.l[j], base));
                        }
                    const int q = q0 + t - (F - 1);             // completed by tap F - 1: slot (u + 1) mod F
                    if (q >= q0 && q < qend) stv<VEC>(out + (long long)(r + q * s) * Nc + p, acc[(u + 1) % F]);
                }
            }
        }
    }
}

inline int pick_ks(int nq, long long cols_total, int period) {
    const long long want = 3LL * 148 * 1024;
    long long nseg = (want + cols_total - 1) / cols_total;
    if (nseg < 1) nseg = 1;
    int ks = (int)((nq + nseg - 1) / nseg);
    if (ks < 4 * period) ks = 4 * period;
    return ((ks + period - 1) / period) * period;
}
inline unsigned grid_for(long long items, int threads) {
    long long g = (items + threads - 1) / threads;
    const long long cap = (long long)pwt_sm_count() * 64;
    return (unsigned)(g < 1 ? 1 : (g > cap ? cap : g));
}
template <int F, int VEC>
void launch_cols_fwd(const ColJobs& jb, int batch, int Nr, int Nc, int s, const PwtTapsFwd& t, cudaStream_t st) {
    const int nqmax = cdiv(Nr, s), PV = Nc / VEC;
    const int KS = pick_ks(nqmax, (long long)PV * s * 2 * batch, F);
    const long long items = (long long)PV * cdiv(nqmax, KS) * s;
    pwt_launch_pdl(k_swt2p_cols_fwd<F, VEC>, dim3(grid_for(items, 128), 2, batch), 128, 0, st, jb, Nr, Nc, s, KS, (long long)Nr * Nc, t);
}
template <int F, int VEC>
void launch_cols_inv(const ColJobs& jb, int batch, int Nr, int Nc, int s, const TapsHalf& t, cudaStream_t st) {
    const int nqmax = cdiv(Nr, s), PV = Nc / VEC;
    const int KS = pick_ks(nqmax, (long long)PV * s * 2 * batch, F);
    const long long items = (long long)PV * cdiv(nqmax, KS) * s;
    pwt_launch_pdl(k_swt2p_cols_inv<F, VEC>, dim3(grid_for(items, 128), 2, batch), 128, 0, st, jb, Nr, Nc, s, KS, (long long)Nr * Nc, t);
}
template <int F>
int launch_rows_fwd(const float* in, float* lo, float* hi, long long rows, int Nc, int s, size_t smem, const PwtTapsFwd& t, cudaStream_t st) {
    static PwtKernelOnce once;
    if (!pwt_kernel_once(once, k_swt2p_rows_fwd<F>, 256, 64 * 1024, smem)) return 0;
    pwt_launch_pdl(k_swt2p_rows_fwd<F>, dim3(grid_for(rows * cdiv(Nc, kTO) * 256, 256)), 256, smem, st, in, lo, hi, rows, Nc, s, t);
    return 1;
}
template <int F>
bool prep_rows_inv(size_t smem) {
    static PwtKernelOnce once;
    return pwt_kernel_once(once, k_swt2p_rows_inv<F>, 256, 64 * 1024, smem);
}
template <int F>
void launch_rows_inv(const float* t1, const float* t2, float* out, long long rows, int Nc, int s, size_t smem, const TapsHalf& t, cudaStream_t st) {
    pwt_launch_pdl(k_swt2p_rows_inv<F>, dim3(grid_for(rows * cdiv(Nc, kTO) * 256, 256)), 256, smem, st, t1, t2, out, rows, Nc, s, t);
}
inline int vec_cap() {                                    // PWT_SWT2P_VEC: cap of the columns per thread (A/B)
    static const int v = [] { const char* e = getenv("PWT_SWT2P_VEC"); return e && *e ? atoi(e) : 4; }();
    return v;
}
inline bool aligned16(const void* a, const void* b, const void* c, const void* d) {
    return ((((uintptr_t)a) | ((uintptr_t)b) | ((uintptr_t)c) | ((uintptr_t)d)) & 15) == 0;
}
}  // namespace

#define PWT_SWT2P_CASES(X) X(2) X(4) X(6) X(8) X(10) X(12) X(14) X(16) X(18) X(20) X(22) X(24) X(26) X(28) X(30) X(32) X(34) X(36) X(38) X(40)

// tmp: 2 * batch * Nr * Nc floats.  Returns the launches (2), or 0 when not covered.
int pwt_swt2p_fwd2d(const float* in, float* A, float* Hb, float* V, float* D, float* tmp, int batch, int Nr, int Nc, int level,
                    const PwtFilters& f, cudaStream_t st) {
    const int F = f.hlen, s = 1 << (level - 1);
    const long long reach = (long long)(F - 1) * s;
    if (F < 2 || F > PWT_MAX_TAPS || (F & 1) || s >= Nr || s >= Nc || reach > 8192 || level > 20) return 0;
    const long long rows = (long long)batch * Nr, planeN = rows * Nc;
    float* lo = tmp;
    float* hi = tmp + planeN;
    const PwtTapsFwd t = pwt_pack_taps_fwd(f, F);
    const size_t smem = sizeof(float) * (size_t)(kTO + reach);
    switch (F) {
#define X(FF) case FF: if (!launch_rows_fwd<FF>(in, lo, hi, rows, Nc, s, smem, t, st)) return 0; break;
        PWT_SWT2P_CASES(X)
#undef X
        default: return 0;
    }
    ColJobs jb = {};
    jb.a[0] = lo; jb.o0[0] = A; jb.o1[0] = Hb;
    jb.a[1] = hi; jb.o0[1] = V; jb.o1[1] = D;
    const bool al = aligned16(lo, A, Hb, V) && aligned16(hi, D, lo, lo) && (((long long)Nr * Nc) & 3) == 0;
    int vec = (F <= 20 && (Nc & 3) == 0 && al) ? 4 : ((Nc & 1) == 0 && al) ? 2 : 1;
    if (vec > vec_cap()) vec = vec_cap();
    switch (F) {
#define X(FF) case FF: if (vec == 4) launch_cols_fwd<FF, (FF <= 20 ? 4 : 2)>(jb, batch, Nr, Nc, s, t, st); \
                       else if (vec == 2) launch_cols_fwd<FF, 2>(jb, batch, Nr, Nc, s, t, st); \
                       else launch_cols_fwd<FF, 1>(jb, batch, Nr, Nc, s, t, st); break;
        PWT_SWT2P_CASES(X)
#undef X
        default: return 0;
    }
    return 2;
}
int pwt_swt2p_inv2d(const float* A, const float* Hb, const float* V, const float* D, float* out, float* tmp, int batch, int Nr, int Nc,
                    int level, const PwtFilters& f, cudaStream_t st) {
    const int F = f.hlen, s = 1 << (level - 1);
    const long long reach = (long long)(F - 1) * s;
    if (F < 2 || F > PWT_MAX_TAPS || (F & 1) || s >= Nr || s >= Nc || reach > 4096 || level > 20) return 0;
    const long long rows = (long long)batch * Nr, planeN = rows * Nc;
    float* t1 = tmp;
    float* t2 = tmp + planeN;
    TapsHalf t;
    for (int j = 0; j < PWT_MAX_TAPS; j++) {
        t.l[j] = j < F ? 0.5f * f.IL[F - 1 - j] : 0.f;
        t.h[j] = j < F ? 0.5f * f.IH[F - 1 - j] : 0.f;
    }
    const size_t smem = sizeof(float) * 2 * (size_t)(kTO + reach);
    switch (F) {
#define X(FF) case FF: if (!prep_rows_inv<FF>(smem)) return 0; break;
        PWT_SWT2P_CASES(X)
#undef X
        default: return 0;
    }
    ColJobs jb = {};
    jb.a[0] = A; jb.b[0] = Hb; jb.o0[0] = t1;
    jb.a[1] = V; jb.b[1] = D; jb.o0[1] = t2;
    const bool al = aligned16(A, Hb, V, D) && aligned16(t1, t2, t1, t1) && (((long long)Nr * Nc) & 3) == 0;
    int vec = (F <= 20 && (Nc & 3) == 0 && al) ? 4 : ((Nc & 1) == 0 && al) ? 2 : 1;
    if (vec > vec_cap()) vec = vec_cap();
    switch (F) {
#define X(FF) case FF: if (vec == 4) launch_cols_inv<FF, (FF <= 20 ? 4 : 2)>(jb, batch, Nr, Nc, s, t, st); \
                       else if (vec == 2) launch_cols_inv<FF, 2>(jb, batch, Nr, Nc, s, t, st); \
                       else launch_cols_inv<FF, 1>(jb, batch, Nr, Nc, s, t, st); break;
        PWT_SWT2P_CASES(X)
#undef X
        default: return 0;
    }
    switch (F) {
#define X(FF) case FF: launch_rows_inv<FF>(t1, t2, out, rows, Nc, s, smem, t, st); break;
        PWT_SWT2P_CASES(X)
#undef X
    }
    return 2;
}
