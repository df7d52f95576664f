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
// 36 B/px per level and direction (20 compulsory), at streaming speed.  Templated on the sample type: the double-precision plans
// run the same passes (two DFMA instead of one FFMA2 per tap pair, 2 columns per thread instead of 4).
#include <stdlib.h>

#include <type_traits>

#include "pwt_internal.h"

namespace {
// (low-pass, high-pass) pairs: float2 + FFMA2 in single precision, two DFMA in double precision
template <typename T> struct V2;
template <> struct V2<float> { using type = float2; };
template <> struct V2<double> { using type = double2; };
template <typename T> using v2_t = typename V2<T>::type;
__device__ __forceinline__ float2 fma2s(float x, float2 t, float2 acc) { return __ffma2_rn(make_float2(x, x), t, acc); }
__device__ __forceinline__ double2 fma2s(double x, double2 t, double2 acc) { return make_double2(fma(x, t.x, acc.x), fma(x, t.y, acc.y)); }
template <typename T> __device__ __forceinline__ v2_t<T> zero2();
template <> __device__ __forceinline__ float2 zero2<float>() { return make_float2(0.f, 0.f); }
template <> __device__ __forceinline__ double2 zero2<double>() { return make_double2(0.0, 0.0); }
template <typename T>
struct TapsPair {                                          // analysis, reversed: t[j] = (L[F-1-j], H[F-1-j])
    v2_t<T> t[PWT_MAX_TAPS];
};
__device__ __forceinline__ int mod_pos(int i, int n) {
    i %= n;
    return i < 0 ? i + n : i;
}
__host__ __device__ inline int cdiv(int a, int b) { return (a + b - 1) / b; }

// ---- rows ------------------------------------------------------------------------------------------------------------
constexpr int kTO = 2048;                                  // outputs per tile
// analysis: out[g] = sum_j f[F-1-j] in[(g + (j - c) s) mod Nc], c = F/2 - 1  ->  lo, hi
template <typename T, int F>
__global__ void __launch_bounds__(256)
k_swt2p_rows_fwd(const T* __restrict__ in, T* __restrict__ lo, T* __restrict__ hi, long long rows, int Nc, int s,
                 const __grid_constant__ TapsPair<T> tp) {
    extern __shared__ __align__(16) unsigned char smraw[];
    T* sx = reinterpret_cast<T*>(smraw);
    const int c = F / 2 - 1, reach = (F - 1) * s, ntile = cdiv(Nc, kTO);
    pwt_pdl_wait();
    for (long long t = blockIdx.x; t < rows * ntile; t += gridDim.x) {
        const long long r = t / ntile;
        const int g0 = (int)(t - r * ntile) * kTO, x0 = g0 - c * s;
        const int nout = min(kTO, Nc - g0), nin = nout + reach;
        const T* row = in + r * Nc;
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
            v2_t<T> p = zero2<T>();
#pragma unroll
            for (int j = 0; j < F; j++) p = fma2s(sx[o + j * s], tp.t[j], p);
            lo[r * Nc + g0 + o] = p.x;
            hi[r * Nc + g0 + o] = p.y;
        }
        __syncthreads();
    }
}
// synthesis: out[g] = sum_j (IL[F-1-j] / 2) t1[(g + (j - F/2) s) mod Nc] + (IH[F-1-j] / 2) t2[...]
template <typename T>
struct TapsHalf {
    T l[PWT_MAX_TAPS], h[PWT_MAX_TAPS];
};
template <typename T, int F>
__global__ void __launch_bounds__(256)
k_swt2p_rows_inv(const T* __restrict__ t1, const T* __restrict__ t2, T* __restrict__ out, long long rows, int Nc, int s,
                 const __grid_constant__ TapsHalf<T> tp) {
    extern __shared__ __align__(16) unsigned char smraw[];
    T* sx = reinterpret_cast<T*>(smraw);
    const int c = F / 2, reach = (F - 1) * s, ntile = cdiv(Nc, kTO);
    T* sa = sx;
    T* sd = sx + kTO + reach;
    pwt_pdl_wait();
    for (long long t = blockIdx.x; t < rows * ntile; t += gridDim.x) {
        const long long r = t / ntile;
        const int g0 = (int)(t - r * ntile) * kTO, x0 = g0 - c * s;
        const int nout = min(kTO, Nc - g0), nin = nout + reach;
        const T* a = t1 + r * Nc;
        const T* d = t2 + r * Nc;
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
            T x = 0;
#pragma unroll
            for (int j = 0; j < F; j++) {
                x = fma(sa[o + j * s], tp.l[j], x);
                x = fma(sd[o + j * s], tp.h[j], x);
            }
            out[r * Nc + g0 + o] = x;
        }
        __syncthreads();
    }
}

// ---- columns -----------------------------------------------------------------------------------------------------------
template <typename T, int VEC>
__device__ __forceinline__ void ldv(T (&w)[VEC], const T* p) {
    if constexpr (VEC == 1) w[0] = __ldg(p);
    else if constexpr (std::is_same<T, float>::value && VEC == 4) { const float4 v = __ldg(reinterpret_cast<const float4*>(p)); w[0] = v.x; w[1] = v.y; w[2] = v.z; w[3] = v.w; }
    else if constexpr (std::is_same<T, float>::value && VEC == 2) { const float2 v = __ldg(reinterpret_cast<const float2*>(p)); w[0] = v.x; w[1] = v.y; }
    else { static_assert(VEC == 2, "double: 1 or 2 columns per thread"); const double2 v = __ldg(reinterpret_cast<const double2*>(p)); w[0] = v.x; w[1] = v.y; }
}
template <typename T, int VEC>
__device__ __forceinline__ void stv(T* p, const T (&w)[VEC]) {
    if constexpr (VEC == 1) p[0] = w[0];
    else if constexpr (std::is_same<T, float>::value && VEC == 4) __stcs(reinterpret_cast<float4*>(p), make_float4(w[0], w[1], w[2], w[3]));
    else if constexpr (std::is_same<T, float>::value && VEC == 2) __stcs(reinterpret_cast<float2*>(p), make_float2(w[0], w[1]));
    else __stcs(reinterpret_cast<double2*>(p), make_double2(w[0], w[1]));
}
template <typename T>
struct ColJobs {
    const T* a[2];         // analysis: lo / hi plane;  synthesis: A / V
    const T* b[2];         // synthesis: H / D
    T* o0[2];              // analysis: A / V;  synthesis: t1 / t2
    T* o1[2];              // analysis: H / D
};
// work item -> (residue class r, run of KS lattice positions, column group); blockIdx.y = job, blockIdx.z = image
template <typename T, int F, int VEC>
__global__ void __launch_bounds__(128)
k_swt2p_cols_fwd(const __grid_constant__ ColJobs<T> jb, int Nr, int Nc, int s, int KS, long long plane, const __grid_constant__ TapsPair<T> tp) {
    constexpr int C = F / 2 - 1;
    const T* __restrict__ in = jb.a[blockIdx.y] + blockIdx.z * plane;
    T* __restrict__ o0 = jb.o0[blockIdx.y] + blockIdx.z * plane;
    T* __restrict__ o1 = jb.o1[blockIdx.y] + blockIdx.z * plane;
    const int PV = Nc / VEC, nqmax = (Nr + s - 1) / s, nseg = (nqmax + KS - 1) / KS;
    pwt_pdl_wait();
    for (long long i = blockIdx.x * 128LL + threadIdx.x; i < (long long)PV * nseg * s; i += gridDim.x * 128LL) {
        const int p = (int)(i % PV) * VEC;
        const int rs = (int)(i / PV), r = rs % s, seg = rs / s;
        const int nq = (Nr - r + s - 1) / s;
        const int q0 = seg * KS, qend = min(q0 + KS, nq);
        if (q0 >= qend) continue;
        T w[F][VEC];
        int y = mod_pos(r + (q0 - C) * s, Nr);                 // row of window position 0
#pragma unroll
        for (int j = 0; j < F - 1; j++) {
            ldv<T, VEC>(w[j], in + (long long)y * Nc + p);
            y += s;
            if (y >= Nr) y -= Nr;
        }
        for (int qb = q0; qb < qend; qb += F) {
#pragma unroll
            for (int u = 0; u < F; u++) {
                const int q = qb + u;
                if (q < qend) {
                    ldv<T, VEC>(w[(F - 1 + u) % F], in + (long long)y * Nc + p);
                    y += s;
                    if (y >= Nr) y -= Nr;
                    v2_t<T> acc[VEC];
#pragma unroll
                    for (int v = 0; v < VEC; v++) acc[v] = zero2<T>();
#pragma unroll
                    for (int j = 0; j < F; j++)
#pragma unroll
                        for (int v = 0; v < VEC; v++) acc[v] = fma2s(w[(j + u) % F][v], tp.t[j], acc[v]);
                    T a[VEC], d[VEC];
#pragma unroll
                    for (int v = 0; v < VEC; v++) { a[v] = acc[v].x; d[v] = acc[v].y; }
                    const long long o = (long long)(r + q * s) * Nc + p;
                    stv<T, VEC>(o0 + o, a);
                    stv<T, VEC>(o1 + o, d);
                }
            }
        }
    }
}
// synthesis, transposed form: stream row k (bands a and d at row (r + k s) mod Nr) adds tap j to output q = k - j + C;
// output q completes with j = F - 1.  slot(q) = (q - q0) mod F.
template <typename T, int F, int VEC>
__global__ void __launch_bounds__(128)
k_swt2p_cols_inv(const __grid_constant__ ColJobs<T> jb, int Nr, int Nc, int s, int KS, long long plane, const __grid_constant__ TapsHalf<T> tp) {
    constexpr int C = F / 2;
    const T* __restrict__ A = jb.a[blockIdx.y] + blockIdx.z * plane;
    const T* __restrict__ B = jb.b[blockIdx.y] + blockIdx.z * plane;
    T* __restrict__ out = jb.o0[blockIdx.y] + blockIdx.z * plane;
    const int PV = Nc / VEC, nqmax = (Nr + s - 1) / s, nseg = (nqmax + KS - 1) / KS;
    pwt_pdl_wait();
    for (long long i = blockIdx.x * 128LL + threadIdx.x; i < (long long)PV * nseg * s; i += gridDim.x * 128LL) {
        const int p = (int)(i % PV) * VEC;
        const int rs = (int)(i / PV), r = rs % s, seg = rs / s;
        const int nq = (Nr - r + s - 1) / s;
        const int q0 = seg * KS, qend = min(q0 + KS, nq);
        if (q0 >= qend) continue;
        T acc[F][VEC];
        int y = mod_pos(r + (q0 - C) * s, Nr);                 // stream step t = 0 is lattice position k = q0 - C
        const int nsteps = (qend - q0) + F - 1;
        for (int tb = 0; tb < nsteps; tb += F) {
#pragma unroll
            for (int u = 0; u < F; u++) {
                const int t = tb + u;
                if (t < nsteps) {
                    T xa[VEC], xd[VEC];
                    ldv<T, VEC>(xa, A + (long long)y * Nc + p);
                    ldv<T, VEC>(xd, B + (long long)y * Nc + p);
                    y += s;
                    if (y >= Nr) y -= Nr;
#pragma unroll
                    for (int j = 0; j < F; j++)                 // output q0 + t - j, slot (u - j) mod F; j = 0 opens the slot
#pragma unroll
                        for (int v = 0; v < VEC; v++) {
                            const T base = j == 0 ? T(0) : acc[((u - j) % F + F) % F][v];
                            acc[((u - j) % F + F) % F][v] = fma(xd[v], tp.h[j], fma(xa[v], tp.l[j], base));
                        }
                    const int q = q0 + t - (F - 1);             // completed by tap F - 1: slot (u + 1) mod F
                    if (q >= q0 && q < qend) stv<T, VEC>(out + (long long)(r + q * s) * Nc + p, acc[(u + 1) % F]);
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
template <typename T, int F, int VEC>
void launch_cols_fwd(const ColJobs<T>& jb, int batch, int Nr, int Nc, int s, const TapsPair<T>& t, cudaStream_t st) {
    const int nqmax = cdiv(Nr, s), PV = Nc / VEC;
    const int KS = pick_ks(nqmax, (long long)PV * s * 2 * batch, F);
    const long long items = (long long)PV * cdiv(nqmax, KS) * s;
    pwt_launch_pdl(k_swt2p_cols_fwd<T, F, VEC>, dim3(grid_for(items, 128), 2, batch), 128, 0, st, jb, Nr, Nc, s, KS, (long long)Nr * Nc, t);
}
template <typename T, int F, int VEC>
void launch_cols_inv(const ColJobs<T>& jb, int batch, int Nr, int Nc, int s, const TapsHalf<T>& t, cudaStream_t st) {
    const int nqmax = cdiv(Nr, s), PV = Nc / VEC;
    const int KS = pick_ks(nqmax, (long long)PV * s * 2 * batch, F);
    const long long items = (long long)PV * cdiv(nqmax, KS) * s;
    pwt_launch_pdl(k_swt2p_cols_inv<T, F, VEC>, dim3(grid_for(items, 128), 2, batch), 128, 0, st, jb, Nr, Nc, s, KS, (long long)Nr * Nc, t);
}
constexpr size_t kRowsSmemCap = 160 * 1024;
template <typename T, int F>
int launch_rows_fwd(const T* in, T* lo, T* hi, long long rows, int Nc, int s, size_t smem, const TapsPair<T>& t, cudaStream_t st) {
    static PwtKernelOnce once;
    if (!pwt_kernel_once(once, k_swt2p_rows_fwd<T, F>, 256, kRowsSmemCap, smem)) return 0;
    pwt_launch_pdl(k_swt2p_rows_fwd<T, F>, dim3(grid_for(rows * cdiv(Nc, kTO) * 256, 256)), 256, smem, st, in, lo, hi, rows, Nc, s, t);
    return 1;
}
template <typename T, int F>
bool prep_rows_inv(size_t smem) {
    static PwtKernelOnce once;
    return pwt_kernel_once(once, k_swt2p_rows_inv<T, F>, 256, kRowsSmemCap, smem);
}
template <typename T, int F>
void launch_rows_inv(const T* t1, const T* t2, T* out, long long rows, int Nc, int s, size_t smem, const TapsHalf<T>& t, cudaStream_t st) {
    pwt_launch_pdl(k_swt2p_rows_inv<T, F>, dim3(grid_for(rows * cdiv(Nc, kTO) * 256, 256)), 256, smem, st, t1, t2, out, rows, Nc, s, t);
}
inline int vec_cap() {                                    // PWT_SWT2P_VEC: cap of the columns per thread (A/B)
    static const int v = [] { const char* e = getenv("PWT_SWT2P_VEC"); return e && *e ? atoi(e) : 4; }();
    return v;
}
inline bool aligned16(const void* a, const void* b, const void* c, const void* d) {
    return ((((uintptr_t)a) | ((uintptr_t)b) | ((uintptr_t)c) | ((uintptr_t)d)) & 15) == 0;
}
// columns per thread: 16 bytes when the rows allow it (F <= 20: the window / the accumulators stay in registers)
template <typename T>
int pick_vec(int F, int Nr, int Nc, bool al) {
    constexpr int full = 16 / (int)sizeof(T);              // 4 floats, 2 doubles
    al = al && (((long long)Nr * Nc) % full) == 0;
    int vec = (F <= 20 && Nc % full == 0 && al) ? full : (sizeof(T) == 4 && (Nc & 1) == 0 && al) ? 2 : 1;
    return vec > vec_cap() ? vec_cap() : vec;
}

#define PWT_SWT2P_CASES(X) X(2) X(4) X(6) X(8) X(10) X(12) X(14) X(16) X(18) X(20) X(22) X(24) X(26) X(28) X(30) X(32) X(34) X(36) X(38) X(40)
// the vector widths a type has: float 4 (F <= 20) / 2 / 1, double 2 (F <= 20) / 1
#define PWT_SWT2P_COLS(T, FF, LAUNCH, ...)                                                             \
    if constexpr (sizeof(T) == 4) {                                                                    \
        if (vec == 4) LAUNCH<T, FF, (sizeof(T) == 4 && FF <= 20 ? 4 : 2)>(__VA_ARGS__);                \
        else if (vec == 2) LAUNCH<T, FF, 2>(__VA_ARGS__);                                              \
        else LAUNCH<T, FF, 1>(__VA_ARGS__);                                                            \
    } else {                                                                                           \
        if (vec == 2) LAUNCH<T, FF, (FF <= 20 ? 2 : 1)>(__VA_ARGS__);                                  \
        else LAUNCH<T, FF, 1>(__VA_ARGS__);                                                            \
    }

template <typename T>
int swt2p_fwd2d_t(const T* in, T* A, T* Hb, T* V, T* D, T* tmp, int batch, int Nr, int Nc, int level, const PwtFiltersT<T>& f,
                  cudaStream_t st) {
    const int F = f.hlen, s = 1 << (level - 1);
    const long long reach = (long long)(F - 1) * s;
    if (F < 2 || F > PWT_MAX_TAPS || (F & 1) || s >= Nr || s >= Nc || reach > 8192 || level > 20) return 0;
    const long long rows = (long long)batch * Nr, planeN = rows * Nc;
    T* lo = tmp;
    T* hi = tmp + planeN;
    TapsPair<T> t;
    for (int j = 0; j < PWT_MAX_TAPS; j++) {
        t.t[j].x = j < F ? f.L[F - 1 - j] : T(0);
        t.t[j].y = j < F ? f.H[F - 1 - j] : T(0);
    }
    const size_t smem = sizeof(T) * (size_t)(kTO + reach);
    switch (F) {
#define X(FF) case FF: if (!launch_rows_fwd<T, FF>(in, lo, hi, rows, Nc, s, smem, t, st)) return 0; break;
        PWT_SWT2P_CASES(X)
#undef X
        default: return 0;
    }
    ColJobs<T> jb = {};
    jb.a[0] = lo; jb.o0[0] = A; jb.o1[0] = Hb;
    jb.a[1] = hi; jb.o0[1] = V; jb.o1[1] = D;
    const int vec = pick_vec<T>(F, Nr, Nc, aligned16(lo, A, Hb, V) && aligned16(hi, D, lo, lo));
    switch (F) {
#define X(FF) case FF: PWT_SWT2P_COLS(T, FF, launch_cols_fwd, jb, batch, Nr, Nc, s, t, st) break;
        PWT_SWT2P_CASES(X)
#undef X
        default: return 0;
    }
    return 2;
}
template <typename T>
int swt2p_inv2d_t(const T* A, const T* Hb, const T* V, const T* D, T* out, T* tmp, int batch, int Nr, int Nc, int level,
                  const PwtFiltersT<T>& f, cudaStream_t st) {
    const int F = f.hlen, s = 1 << (level - 1);
    const long long reach = (long long)(F - 1) * s;
    if (F < 2 || F > PWT_MAX_TAPS || (F & 1) || s >= Nr || s >= Nc || reach > 4096 || level > 20) return 0;
    const long long rows = (long long)batch * Nr, planeN = rows * Nc;
    T* t1 = tmp;
    T* t2 = tmp + planeN;
    TapsHalf<T> t;
    for (int j = 0; j < PWT_MAX_TAPS; j++) {
        t.l[j] = j < F ? T(0.5) * f.IL[F - 1 - j] : T(0);
        t.h[j] = j < F ? T(0.5) * f.IH[F - 1 - j] : T(0);
    }
    const size_t smem = sizeof(T) * 2 * (size_t)(kTO + reach);
    switch (F) {
#define X(FF) case FF: if (!prep_rows_inv<T, FF>(smem)) return 0; break;
        PWT_SWT2P_CASES(X)
#undef X
        default: return 0;
    }
    ColJobs<T> jb = {};
    jb.a[0] = A; jb.b[0] = Hb; jb.o0[0] = t1;
    jb.a[1] = V; jb.b[1] = D; jb.o0[1] = t2;
    const int vec = pick_vec<T>(F, Nr, Nc, aligned16(A, Hb, V, D) && aligned16(t1, t2, t1, t1));
    switch (F) {
#define X(FF) case FF: PWT_SWT2P_COLS(T, FF, launch_cols_inv, jb, batch, Nr, Nc, s, t, st) break;
        PWT_SWT2P_CASES(X)
#undef X
        default: return 0;
    }
    switch (F) {
#define X(FF) case FF: launch_rows_inv<T, FF>(t1, t2, out, rows, Nc, s, smem, t, st); break;
        PWT_SWT2P_CASES(X)
#undef X
    }
    return 2;
}
}  // namespace

// tmp: 2 * batch * Nr * Nc samples.  Return the launches (2), or 0 when not covered.
int pwt_swt2p_fwd2d(const float* in, float* A, float* Hb, float* V, float* D, float* tmp, int batch, int Nr, int Nc, int level,
                    const PwtFilters& f, cudaStream_t st) {
    return swt2p_fwd2d_t<float>(in, A, Hb, V, D, tmp, batch, Nr, Nc, level, f, st);
}
int pwt_swt2p_inv2d(const float* A, const float* Hb, const float* V, const float* D, float* out, float* tmp, int batch, int Nr, int Nc,
                    int level, const PwtFilters& f, cudaStream_t st) {
    return swt2p_inv2d_t<float>(A, Hb, V, D, out, tmp, batch, Nr, Nc, level, f, st);
}
int pwt_swt2p_fwd2d_f64(const double* in, double* A, double* Hb, double* V, double* D, double* tmp, int batch, int Nr, int Nc, int level,
                        const PwtFilters64& f, cudaStream_t st) {
    return swt2p_fwd2d_t<double>(in, A, Hb, V, D, tmp, batch, Nr, Nc, level, f, st);
}
int pwt_swt2p_inv2d_f64(const double* A, const double* Hb, const double* V, const double* D, double* out, double* tmp, int batch, int Nr,
                        int Nc, int level, const PwtFilters64& f, cudaStream_t st) {
    return swt2p_inv2d_t<double>(A, Hb, V, D, out, tmp, batch, Nr, Nc, level, f, st);
}
