// Double-precision plans: the reference's DOUBLEPRECISION build (pdwt/src/filters.h:16-30 `#define DTYPE double`,
// pdwt/Makefile:36-39 libpdwtd.so), which its Python wrapper never reaches.  Same class surface as
// the fp32 plan -- wt.h:42-75 with DTYPE = double -- behind `pwt64_*` entry points (include/pwt_b200.h).
//
// Kernels: 2D DWT levels with F = 4 .. 20 run the two-pass streaming kernels below; everything else (Haar, longer filters,
// batched 1D, a-trous) the separable tile kernels of kernels_generic.cu instantiated for double.  Taps at the table's full precision.  The roofline doubles per pixel
// (16 B forward, 16 B inverse); the specialised fp32 families (register cascade, strip kernels) are not instantiated
// for double -- measured numbers in profiles/r02_notes.md.  Non-separable mode uses the rank-1 identity (the four
// F x F banks are outer products of the 1D bank): separable kernels, detail slots 1 and 2 swapped like the reference
// (nonseparable.cu:71-78).
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <strings.h>

#include "../../include/pwt_b200.h"
#include "pwt_internal.h"

int pwt_is_haar_alias(const char* wname);
int pwt_fill_filters64(const char* wname, PwtFilters64* out);
extern "C" const char* pwt_last_error(void);
// kernels_f64_fused.cu: the Haar butterfly of a 2D level, one thread per pair of band columns (0: switched off)
int pwt64_haar_fwd2d(const double* in, double* A, double* Hb, double* V, double* D, int batch, int Nr, int Nc, long long in_bs,
                     long long out_bs, cudaStream_t st);
int pwt64_haar_inv2d(const double* A, const double* Hb, const double* V, const double* D, double* out, int batch, int nr, int nc,
                     int Nro, int Nco, long long in_bs, long long out_bs, cudaStream_t st);
// kernels_swt2p.cu instantiated for double: a 2D a-trous level as two streaming passes (0: not covered)
int pwt_swt2p_fwd2d_f64(const double* in, double* A, double* Hb, double* V, double* D, double* tmp, int batch, int Nr, int Nc, int level,
                        const PwtFilters64& f, cudaStream_t st);
int pwt_swt2p_inv2d_f64(const double* A, const double* Hb, const double* V, const double* D, double* out, double* tmp, int batch, int Nr,
                        int Nc, int level, const PwtFilters64& f, cudaStream_t st);
int pwt_set_error(int code, const char* msg);      // pwt_plan.cu: stores the thread-local message, returns code

namespace {
int fail64(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    return pwt_set_error(code, buf);
}
#define CK64(call)                                                                                  \
    do {                                                                                            \
        cudaError_t e_ = (call);                                                                    \
        if (e_ != cudaSuccess) return fail64(PWT_ERR_CUDA, "%s failed: %s", #call, cudaGetErrorString(e_)); \
    } while (0)

inline int div2i(int n) { return (n + 1) >> 1; }
inline int ilog2i(int i) {
    int l = 0;
    while (i > 1) { i >>= 1; ++l; }
    return l;
}
inline size_t align32(size_t n) { return (n + 31) & ~(size_t)31; }

// ---- element-wise operators, norms, circshift (double) ------------------------------------------
enum { OP_SOFT = 0, OP_HARD = 1, OP_SCALE = 2 };
template <int OP>
__global__ void __launch_bounds__(256) k64_eltwise(double* __restrict__ p, long long n, double beta) {
    for (long long i = blockIdx.x * 256LL + threadIdx.x; i < n; i += gridDim.x * 256LL) {
        const double v = p[i];
        double r;
        if (OP == OP_SOFT) r = copysign(fmax(fabs(v) - beta, 0.0), v);          // common.cu:13-22
        else if (OP == OP_HARD) r = fabs(v) > beta ? v : 0.0;                    // common.cu:56-64
        else r = v * beta;                                                       // common.cu:347-371 (scal)
        p[i] = r;
    }
}
__global__ void __launch_bounds__(256) k64_norms(const double* __restrict__ p, long long n, double* __restrict__ acc) {
    double s1 = 0.0, s2 = 0.0;
    for (long long i = blockIdx.x * 256LL + threadIdx.x; i < n; i += gridDim.x * 256LL) {
        const double v = p[i];
        s1 += fabs(v);
        s2 = fma(v, v, s2);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        s1 += __shfl_xor_sync(0xffffffffu, s1, o);
        s2 += __shfl_xor_sync(0xffffffffu, s2, o);
    }
    __shared__ double w1[8], w2[8];
    if ((threadIdx.x & 31) == 0) { w1[threadIdx.x >> 5] = s1; w2[threadIdx.x >> 5] = s2; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int k = 1; k < 8; k++) { s1 += w1[k]; s2 += w2[k]; }
        atomicAdd(acc, s1);
        atomicAdd(acc + 1, s2);
    }
}
// out[y, x] = in[(y - sr) mod Nr, (x - sc) mod Nc]  (common.cu:202-211)
__global__ void __launch_bounds__(256)
k64_circshift(const double* __restrict__ in, double* __restrict__ out, int Nr, int Nc, int sr, int sc) {
    const long long plane = (long long)Nr * Nc, pb = blockIdx.z * plane;
    for (long long i = blockIdx.x * 256LL + threadIdx.x; i < plane; i += gridDim.x * 256LL) {
        const int y = (int)(i / Nc), x = (int)(i - (long long)y * Nc);
        int ys = y - sr, xs = x - sc;
        if (ys < 0) ys += Nr;
        if (xs < 0) xs += Nc;
        out[pb + i] = in[pb + (long long)ys * Nc + xs];
    }
}

// ---- two-pass separable kernels for double, compile-time F = 4 .. 20 --------------------------------------------
// The generic tile kernels (kernels_generic.cu) reach 0.12-0.21 of the fp64 roofline: every tap is a 64-bit shared-memory
// read.  These run a level as two streaming passes instead -- rows (threads along x, 4 output pairs per thread from one
// window of F + 6 samples read through L1), then columns (a thread owns 2 adjacent columns and a run of consecutive
// outputs; the F input rows live in a rotating register window, two new rows per output) -- 32 instead of 16 B per
// level-input pixel and direction, but at streaming speed.  Same arithmetic order as the reference (taps ascending).
__device__ __forceinline__ int wrap_dwt64(int i, int N) {
    const int Ne = N + (N & 1);
    i %= Ne;
    if (i < 0) i += Ne;
    return i >= N ? N - 1 : i;
}
__device__ __forceinline__ int wrap_per64(int i, int N) {
    i %= N;
    return i < 0 ? i + N : i;
}
// rows, analysis: in [rows][Nc] -> lo, hi [rows][Nc2].  A CTA stages the inputs of TO consecutive outputs of one row in
// shared memory with coalesced loads (periodic wrap resolved while staging), then every thread produces outputs
// tid, tid + 256, ... so that the stores are coalesced too.  (First version: 4 outputs per thread straight from global
// memory -- ncu: 2x the time of the column pass for the same bytes, L1 at 67 %, DRAM at 33 %.)
constexpr int kRowTile64 = 1024;                           // outputs per tile
template <typename T, int F>
__global__ void __launch_bounds__(256)
k64_rows_fwd(const T* __restrict__ in, T* __restrict__ lo, T* __restrict__ hi, long long rows, int Nc,
             const __grid_constant__ PwtFiltersT<T> f) {
    constexpr int C = F / 2 - 1, TO = kRowTile64, TI = 2 * TO + F - 2;
    __shared__ T sx[TI];
    const int Nc2 = (Nc + 1) >> 1, ntile = (Nc2 + TO - 1) / TO;
    for (long long t = blockIdx.x; t < rows * ntile; t += gridDim.x) {
        const long long r = t / ntile;
        const int k0 = (int)(t - r * ntile) * TO, x0 = 2 * k0 - C;
        const int nout = min(TO, Nc2 - k0), nin = 2 * nout + F - 2;
        const T* row = in + r * Nc;
        if (x0 >= 0 && x0 + nin <= Nc) {
            for (int i = threadIdx.x; i < nin; i += 256) sx[i] = __ldg(row + x0 + i);
        } else {
            for (int i = threadIdx.x; i < nin; i += 256) sx[i] = __ldg(row + wrap_dwt64(x0 + i, Nc));
        }
        __syncthreads();
        for (int o = threadIdx.x; o < nout; o += 256) {
            T a = 0, d = 0;
#pragma unroll
            for (int j = 0; j < F; j++) {
                const T x = sx[2 * o + j];
                a = fma(x, f.L[F - 1 - j], a);
                d = fma(x, f.H[F - 1 - j], d);
            }
            lo[r * Nc2 + k0 + o] = a;
            hi[r * Nc2 + k0 + o] = d;
        }
        __syncthreads();
    }
}
// rows, synthesis: t1, t2 [rows][nc] -> out [rows][Nc_out];  x[n] = sum_jj IL[2 jj + t0] t1[kb - jj] + IH[2 jj + t0] t2[kb - jj],
// n = 2 j + b, t0 = (b + P) & 1, kb = j + ((b + P) >> 1), indices modulo nc (separable.cu:293-328).  Same tiling.
template <typename T, int F>
__global__ void __launch_bounds__(256)
k64_rows_inv(const T* __restrict__ t1, const T* __restrict__ t2, T* __restrict__ out, long long rows, int nc,
             int Nc_out, const __grid_constant__ PwtFiltersT<T> f) {
    constexpr int P = F / 2 - 1, HALF = F / 2, TO = kRowTile64 / 2, HB = HALF - 1 - (P >> 1), TI = TO + HALF;   // HB: samples needed below j0
    __shared__ T sa[TI], sd[TI];
    const int ntile = (nc + TO - 1) / TO;
    for (long long t = blockIdx.x; t < rows * ntile; t += gridDim.x) {
        const long long r = t / ntile;
        const int j0 = (int)(t - r * ntile) * TO, kmin = j0 - HB;
        const int npair = min(TO, nc - j0), nin = npair + HALF;
        const T* a = t1 + r * nc;
        const T* d = t2 + r * nc;
        if (kmin >= 0 && kmin + nin <= nc) {
            for (int i = threadIdx.x; i < nin; i += 256) { sa[i] = __ldg(a + kmin + i); sd[i] = __ldg(d + kmin + i); }
        } else {
            for (int i = threadIdx.x; i < nin; i += 256) { const int k = wrap_per64(kmin + i, nc); sa[i] = __ldg(a + k); sd[i] = __ldg(d + k); }
        }
        __syncthreads();
        for (int o = threadIdx.x; o < npair; o += 256) {
            T x[2];
#pragma unroll
            for (int b = 0; b < 2; b++) {
                const int t0 = (b + P) & 1, kb = o + ((b + P) >> 1) + HB;            // window index of tap jj = 0
                T v = 0;
#pragma unroll
                for (int jj = 0; jj < HALF; jj++) {
                    v = fma(sa[kb - jj], f.IL[2 * jj + t0], v);
                    v = fma(sd[kb - jj], f.IH[2 * jj + t0], v);
                }
                x[b] = v;
            }
            const int n = 2 * (j0 + o);
            T* op = out + r * Nc_out + n;
            if (n + 1 < Nc_out && ((((uintptr_t)op) & (2 * sizeof(T) - 1)) == 0)) {
                if constexpr (sizeof(T) == 8) *reinterpret_cast<double2*>(op) = make_double2(x[0], x[1]);
                else *reinterpret_cast<float2*>(op) = make_float2(x[0], x[1]);
            }
            else {
                if (n < Nc_out) op[0] = x[0];
                if (n + 1 < Nc_out) op[1] = x[1];
            }
        }
        __syncthreads();
    }
}
struct ColJobs64 {
    const double* a[2];
    const double* b[2];
    double* o0[2];
    double* o1[2];
};
// columns, analysis: plane [Nr][P] (batch stride in_bs) -> lo, hi [Nr2][P] (batch stride out_bs); blockIdx.y = job, z = image
template <int F, int VEC>
__global__ void __launch_bounds__(128)
k64_cols_fwd(const __grid_constant__ ColJobs64 jb, int Nr, int P, long long in_bs, long long out_bs, int KS,
             const __grid_constant__ PwtFilters64 f) {
    constexpr int C = F / 2 - 1;
    const double* __restrict__ in = jb.a[blockIdx.y] + blockIdx.z * in_bs;
    double* __restrict__ lo = jb.o0[blockIdx.y] + blockIdx.z * out_bs;
    double* __restrict__ hi = jb.o1[blockIdx.y] + blockIdx.z * out_bs;
    const int Nr2 = (Nr + 1) >> 1, PV = P / VEC, nseg = (Nr2 + KS - 1) / KS;
    for (long long i = blockIdx.x * 128LL + threadIdx.x; i < (long long)PV * nseg; i += gridDim.x * 128LL) {
        const int seg = (int)(i / PV), p = (int)(i - (long long)seg * PV) * VEC;
        const int k0 = seg * KS, kend = min(k0 + KS, Nr2);
        double w[F][VEC];
        auto ld = [&](int slot, int row) {
            const double* q = in + (long long)wrap_dwt64(row, Nr) * P + p;
            if (VEC == 2) { const double2 v = __ldg(reinterpret_cast<const double2*>(q)); w[slot][0] = v.x; w[slot][VEC - 1] = v.y; }
            else w[slot][0] = __ldg(q);
        };
#pragma unroll
        for (int j = 0; j < F - 2; j++) ld(j, 2 * k0 - C + j);
        for (int kb = k0; kb < kend; kb += F / 2) {
#pragma unroll
            for (int u = 0; u < F / 2; u++) {
                const int k = kb + u;
                if (k < kend) {
                    ld((F - 2 + 2 * u) % F, 2 * k - C + F - 2);
                    ld((F - 1 + 2 * u) % F, 2 * k - C + F - 1);
                    double a[VEC], d[VEC];
#pragma unroll
                    for (int v = 0; v < VEC; v++) a[v] = d[v] = 0.0;
#pragma unroll
                    for (int j = 0; j < F; j++)
#pragma unroll
                        for (int v = 0; v < VEC; v++) {
                            a[v] = fma(w[(j + 2 * u) % F][v], f.L[F - 1 - j], a[v]);
                            d[v] = fma(w[(j + 2 * u) % F][v], f.H[F - 1 - j], d[v]);
                        }
                    double* ol = lo + (long long)k * P + p;
                    double* oh = hi + (long long)k * P + p;
                    if (VEC == 2) {
                        *reinterpret_cast<double2*>(ol) = make_double2(a[0], a[VEC - 1]);
                        *reinterpret_cast<double2*>(oh) = make_double2(d[0], d[VEC - 1]);
                    } else {
                        ol[0] = a[0];
                        oh[0] = d[0];
                    }
                }
            }
        }
    }
}
// columns, synthesis: lo, hi [nr2][P] -> out [Nr_out][P]: output pair j (rows 2j, 2j + 1) from band rows j - S1 .. j - S1 + F/2
template <int F, int VEC>
__global__ void __launch_bounds__(128)
k64_cols_inv(const __grid_constant__ ColJobs64 jb, int nr2, int Nr_out, int P, long long in_bs, long long out_bs, int KS,
             const __grid_constant__ PwtFilters64 f) {
    constexpr int PP = F / 2 - 1, S1 = (PP + 1) >> 1, W = F / 2 + 1;
    const double* __restrict__ lo = jb.a[blockIdx.y] + blockIdx.z * in_bs;
    const double* __restrict__ hi = jb.b[blockIdx.y] + blockIdx.z * in_bs;
    double* __restrict__ out = jb.o0[blockIdx.y] + blockIdx.z * out_bs;
    const int PV = P / VEC, nseg = (nr2 + KS - 1) / KS;
    for (long long i = blockIdx.x * 128LL + threadIdx.x; i < (long long)PV * nseg; i += gridDim.x * 128LL) {
        const int seg = (int)(i / PV), p = (int)(i - (long long)seg * PV) * VEC;
        const int j0 = seg * KS, jend = min(j0 + KS, nr2);
        double wa[W][VEC], wd[W][VEC];
        auto ld = [&](int slot, int row) {
            const long long o = (long long)wrap_per64(row, nr2) * P + p;
            if (VEC == 2) {
                const double2 x = __ldg(reinterpret_cast<const double2*>(lo + o)), y = __ldg(reinterpret_cast<const double2*>(hi + o));
                wa[slot][0] = x.x; wa[slot][VEC - 1] = x.y; wd[slot][0] = y.x; wd[slot][VEC - 1] = y.y;
            } else {
                wa[slot][0] = __ldg(lo + o);
                wd[slot][0] = __ldg(hi + o);
            }
        };
#pragma unroll
        for (int w = 0; w < W - 1; w++) ld(w, j0 - S1 + w);
        for (int jb0 = j0; jb0 < jend; jb0 += W) {
#pragma unroll
            for (int u = 0; u < W; u++) {
                const int j = jb0 + u;
                if (j < jend) {
                    ld((W - 1 + u) % W, j - S1 + W - 1);
                    // window position w holds band row j - S1 + w; output 2j + b uses taps 2 jj + t0 at row j + ((b+PP)>>1) - jj
#pragma unroll
                    for (int b = 0; b < 2; b++) {
                        const int t0 = (b + PP) & 1, wb = ((b + PP) >> 1) + S1;
                        double x[VEC];
#pragma unroll
                        for (int v = 0; v < VEC; v++) x[v] = 0.0;
#pragma unroll
                        for (int jj = 0; jj < F / 2; jj++)
#pragma unroll
                            for (int v = 0; v < VEC; v++) {
                                x[v] = fma(wa[(wb - jj + u) % W][v], f.IL[2 * jj + t0], x[v]);
                                x[v] = fma(wd[(wb - jj + u) % W][v], f.IH[2 * jj + t0], x[v]);
                            }
                        const int n = 2 * j + b;
                        if (n < Nr_out) {
                            double* o = out + (long long)n * P + p;
                            if (VEC == 2) *reinterpret_cast<double2*>(o) = make_double2(x[0], x[VEC - 1]);
                            else o[0] = x[0];
                        }
                    }
                }
            }
        }
    }
}
inline int pick_ks64(int n_out, int PV, int period, int jobs) {
    const long long want_threads = 2LL * 148 * 1024;
    long long nseg = (want_threads + (long long)jobs * PV - 1) / ((long long)jobs * PV);
    if (nseg < 1) nseg = 1;
    int ks = (int)((n_out + nseg - 1) / nseg);
    if (ks < 4 * period) ks = 4 * period;
    return ((ks + period - 1) / period) * period;
}
inline unsigned grid64(long long items, int threads) {
    long long g = (items + threads - 1) / threads;
    const long long cap = (long long)pwt_sm_count() * 64;
    return (unsigned)(g < 1 ? 1 : (g > cap ? cap : g));
}
#define PWT64_CASES(X) X(4) X(6) X(8) X(10) X(12) X(14) X(16) X(18) X(20)
// one 2D level, two passes; tmp holds 2 * batch * Nr * ceil(Nc / 2) doubles.  Returns the launches (0: not covered).
int level_fwd2d_2pass(const double* in, double* A, double* Hb, double* V, double* D, double* tmp, int batch, int Nr, int Nc,
                      long long in_bs, long long out_bs, const PwtFilters64& f, cudaStream_t st) {
    const int F = f.hlen, Nc2 = (Nc + 1) >> 1, Nr2 = (Nr + 1) >> 1;
    if (F < 4 || F > 20 || (F & 1) || in_bs != (long long)Nr * Nc || out_bs != (long long)Nr2 * Nc2) return 0;
    const long long rows = (long long)batch * Nr, half = rows * Nc2;
    double* lo = tmp;
    double* hi = tmp + half;
    ColJobs64 jb = {};
    jb.a[0] = lo; jb.o0[0] = A; jb.o1[0] = Hb;
    jb.a[1] = hi; jb.o0[1] = V; jb.o1[1] = D;
    const bool vec = (Nc2 & 1) == 0 && ((((uintptr_t)lo) | ((uintptr_t)hi) | ((uintptr_t)A) | ((uintptr_t)Hb) | ((uintptr_t)V) | ((uintptr_t)D)) & 15) == 0 &&
                     (((long long)Nr * Nc2) & 1) == 0 && (out_bs & 1) == 0;
    const int PV = vec ? Nc2 / 2 : Nc2;
    const int KS = pick_ks64(Nr2, PV * batch, F / 2, 2);
    const dim3 gc(grid64((long long)PV * ((Nr2 + KS - 1) / KS), 128), 2, batch);
    const unsigned gr = grid64(rows * ((Nc2 + kRowTile64 - 1) / kRowTile64) * 256, 256);
    switch (F) {
#define X(FF) case FF: k64_rows_fwd<double, FF><<<gr, 256, 0, st>>>(in, lo, hi, rows, Nc, f); \
        if (vec) k64_cols_fwd<FF, 2><<<gc, 128, 0, st>>>(jb, Nr, Nc2, (long long)Nr * Nc2, out_bs, KS, f); \
        else k64_cols_fwd<FF, 1><<<gc, 128, 0, st>>>(jb, Nr, Nc2, (long long)Nr * Nc2, out_bs, KS, f); \
        return 2;
        PWT64_CASES(X)
#undef X
    }
    return 0;
}
// batched 1D levels with the same row kernels (lo -> A, hi -> D): 0 when not covered
int level_fwd1d_rows(const double* in, double* A, double* D, long long rows, int Nc, const PwtFilters64& f, cudaStream_t st) {
    const int F = f.hlen, Nc2 = (Nc + 1) >> 1;
    if (F < 4 || F > 20 || (F & 1)) return 0;
    const unsigned gr = grid64(rows * ((Nc2 + kRowTile64 - 1) / kRowTile64) * 256, 256);
    switch (F) {
#define X(FF) case FF: k64_rows_fwd<double, FF><<<gr, 256, 0, st>>>(in, A, D, rows, Nc, f); return 1;
        PWT64_CASES(X)
#undef X
    }
    return 0;
}
int level_inv1d_rows(const double* A, const double* D, double* out, long long rows, int nc, int Nc_out, const PwtFilters64& f,
                     cudaStream_t st) {
    const int F = f.hlen;
    if (F < 4 || F > 20 || (F & 1)) return 0;
    const unsigned gr = grid64(rows * ((nc + kRowTile64 / 2 - 1) / (kRowTile64 / 2)) * 256, 256);
    switch (F) {
#define X(FF) case FF: k64_rows_inv<double, FF><<<gr, 256, 0, st>>>(A, D, out, rows, nc, Nc_out, f); return 1;
        PWT64_CASES(X)
#undef X
    }
    return 0;
}
}  // namespace
// The same tiled row kernels for float: batched 1D levels with FEW, LONG rows (a single 16 M-sample signal ran the strip kernels
// -- one 256-column strip per CTA, a one-row "walk" -- at 0.04 of the roofline).  0: not covered.
int pwt_rows1d_fwd_f32(const float* in, float* A, float* D, long long rows, int Nc, const PwtFilters& f, cudaStream_t st) {
    const int F = f.hlen, Nc2 = (Nc + 1) >> 1;
    if (F < 4 || F > 20 || (F & 1)) return 0;
    const unsigned gr = grid64(rows * ((Nc2 + kRowTile64 - 1) / kRowTile64) * 256, 256);
    switch (F) {
#define X(FF) case FF: k64_rows_fwd<float, FF><<<gr, 256, 0, st>>>(in, A, D, rows, Nc, f); return 1;
        PWT64_CASES(X)
#undef X
    }
    return 0;
}
int pwt_rows1d_inv_f32(const float* A, const float* D, float* out, long long rows, int nc, int Nc_out, const PwtFilters& f, cudaStream_t st) {
    const int F = f.hlen;
    if (F < 4 || F > 20 || (F & 1)) return 0;
    const unsigned gr = grid64(rows * ((nc + kRowTile64 / 2 - 1) / (kRowTile64 / 2)) * 256, 256);
    switch (F) {
#define X(FF) case FF: k64_rows_inv<float, FF><<<gr, 256, 0, st>>>(A, D, out, rows, nc, Nc_out, f); return 1;
        PWT64_CASES(X)
#undef X
    }
    return 0;
}
namespace {
int level_inv2d_2pass(const double* A, const double* Hb, const double* V, const double* D, double* out, double* tmp, int batch,
                      int nr, int nc, int Nro, int Nco, long long in_bs, long long out_bs, const PwtFilters64& f, cudaStream_t st) {
    const int F = f.hlen;
    if (F < 4 || F > 20 || (F & 1) || in_bs != (long long)nr * nc || out_bs != (long long)Nro * Nco) return 0;
    const long long rows = (long long)batch * Nro, half = rows * nc;
    double* t1 = tmp;
    double* t2 = tmp + half;
    ColJobs64 jb = {};
    jb.a[0] = A; jb.b[0] = Hb; jb.o0[0] = t1;
    jb.a[1] = V; jb.b[1] = D; jb.o0[1] = t2;
    const bool vec = (nc & 1) == 0 && ((((uintptr_t)t1) | ((uintptr_t)t2) | ((uintptr_t)A) | ((uintptr_t)Hb) | ((uintptr_t)V) | ((uintptr_t)D)) & 15) == 0 &&
                     (in_bs & 1) == 0 && (((long long)Nro * nc) & 1) == 0;
    const int PV = vec ? nc / 2 : nc;
    const int KS = pick_ks64(nr, PV * batch, F / 2 + 1, 2);
    const dim3 gc(grid64((long long)PV * ((nr + KS - 1) / KS), 128), 2, batch);
    const unsigned gr = grid64(rows * ((nc + kRowTile64 / 2 - 1) / (kRowTile64 / 2)) * 256, 256);
    switch (F) {
#define X(FF) case FF: \
        if (vec) k64_cols_inv<FF, 2><<<gc, 128, 0, st>>>(jb, nr, Nro, nc, in_bs, (long long)Nro * nc, KS, f); \
        else k64_cols_inv<FF, 1><<<gc, 128, 0, st>>>(jb, nr, Nro, nc, in_bs, (long long)Nro * nc, KS, f); \
        k64_rows_inv<double, FF><<<gr, 256, 0, st>>>(t1, t2, out, rows, nc, Nco, f); \
        return 2;
        PWT64_CASES(X)
#undef X
    }
    return 0;
}

inline unsigned grid_for(long long n) {
    long long g = (n + 255) / 256;
    const long long cap = (long long)pwt_sm_count() * 16;
    return (unsigned)(g < 1 ? 1 : (g > cap ? cap : g));
}
}  // namespace

struct pwt64_plan {
    int device;
    cudaStream_t stream;
    cudaEvent_t ev0, ev1;
    int batch, Nr, Nc, ndims, nlevels, hlen, do_swt, do_separable, do_cs;
    int state, shift_r, shift_c;
    int custom_len;      // taps given to pwt64_set_filters_forward (0: a built-in bank); a custom 2-tap bank is NOT the Haar butterfly
    char wname[128];
    PwtFilters64 filt;
    int lvNr[PWT_MAX_LEVELS + 1], lvNc[PWT_MAX_LEVELS + 1];
    double* slab;
    double* d_image;
    double* d_tmp;       // DWT: one image-sized plane; 2D SWT: three
    double* d_tmp2;      // 2D DWT: the two row-pass planes of the two-pass kernels (2 x Nr x ceil(Nc/2))
    int nbands;
    double* d_band[PWT_MAX_BANDS];
    int band_nr[PWT_MAX_BANDS], band_nc[PWT_MAX_BANDS];
    double* d_acc;
    double* h_acc;
    long long launches;
};

namespace {
inline bool is_haar(const pwt64_plan* p) { return p->hlen == 2 && !p->do_swt && !p->custom_len; }
inline long long img_elems(const pwt64_plan* p) { return (long long)p->Nr * p->Nc; }
inline long long lvl_elems(const pwt64_plan* p, int l) { return (long long)p->lvNr[l] * p->lvNc[l]; }
inline long long band_elems(const pwt64_plan* p, int b) { return (long long)p->band_nr[b] * p->band_nc[b]; }

int circshift64(pwt64_plan* p, int sr, int sc) {
    const int Nr = p->Nr, Nc = p->Nc;
    sr %= Nr; sc %= Nc;
    if (sr < 0) sr += Nr;
    if (sc < 0) sc += Nc;
    const size_t n = (size_t)p->batch * img_elems(p);
    k64_circshift<<<dim3(grid_for(img_elems(p)), 1, p->batch), 256, 0, p->stream>>>(p->d_image, p->d_tmp, Nr, Nc, sr, sc);
    CK64(cudaMemcpyAsync(p->d_image, p->d_tmp, n * sizeof(double), cudaMemcpyDeviceToDevice, p->stream));
    p->launches++;
    return PWT_OK;
}
}  // namespace

extern "C" void pwt64_destroy(pwt64_plan* p) {
    if (!p) return;
    cudaSetDevice(p->device);
    if (p->stream) cudaStreamSynchronize(p->stream);
    if (p->slab) cudaFree(p->slab);
    if (p->d_acc) cudaFree(p->d_acc);
    if (p->h_acc) cudaFreeHost(p->h_acc);
    if (p->ev0) cudaEventDestroy(p->ev0);
    if (p->ev1) cudaEventDestroy(p->ev1);
    if (p->stream) cudaStreamDestroy(p->stream);
    free(p);
}

// wt.cu:84-185 with DTYPE = double; `batch` stacked images like pwt_create_batch
extern "C" int pwt64_create(pwt64_plan** out, const double* img, int batch, int Nr, int Nc, const char* wname, int levels,
                            int memisonhost, int do_separable, int do_cycle_spinning, int do_swt, int ndim) {
    if (!out) return fail64(PWT_ERR_ARG, "null output handle");
    *out = nullptr;
    if (!wname || Nr < 1 || Nc < 1 || batch < 1) return fail64(PWT_ERR_ARG, "invalid geometry %dx%dx%d", batch, Nr, Nc);
    if ((long long)Nr * Nc >= (1LL << 31)) return fail64(PWT_ERR_ARG, "one image must hold < 2^31 samples");
    if (ndim > 2) return fail64(PWT_ERR_UNSUPPORTED, "ndim=%d is not implemented", ndim);
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1)
        return fail64(PWT_ERR_CUDA, "no CUDA device available (this library has no CPU fallback)");
    pwt64_plan* p = (pwt64_plan*)calloc(1, sizeof(pwt64_plan));
    if (!p) return fail64(PWT_ERR_NOMEM, "out of host memory");
    cudaGetDevice(&p->device);
    p->batch = batch; p->Nr = Nr; p->Nc = Nc;
    p->do_swt = do_swt ? 1 : 0;
    p->do_cs = do_cycle_spinning ? 1 : 0;
    p->state = PWT_INIT;
    p->ndims = (ndim < 2 || Nr == 1) ? 1 : 2;                       // wt.cu:133-136
    p->do_separable = (p->ndims == 1) ? 1 : (do_separable ? 1 : 0);
    strncpy(p->wname, wname, sizeof(p->wname) - 1);
    if (levels < 1) levels = 1;
    if (p->do_swt && pwt_is_haar_alias(wname) && strcasecmp(wname, "haar")) {
        free(p);
        return fail64(PWT_ERR_UNKNOWN_WAVELET, "unknown wavelet '%s' for the stationary transform", wname);
    }
    const int hlen = pwt_fill_filters64(wname, &p->filt);
    if (hlen < 0) {
        free(p);
        return fail64(PWT_ERR_UNKNOWN_WAVELET, "unknown wavelet name '%s'", wname);
    }
    p->hlen = hlen;
    const int N = p->ndims == 2 ? (Nr < Nc ? Nr : Nc) : Nc;         // wt.cu:156-165
    const int wmaxlev = ilog2i(N / (hlen - 1));
    if (wmaxlev < 1) {
        free(p);
        return fail64(PWT_ERR_TOO_SMALL, "a %dx%d image is too small for wavelet %s (%d taps)", Nr, Nc, wname, hlen);
    }
    if (levels > wmaxlev) {
        printf("Warning: required level (%d) is greater than the maximum possible level for %s (%d) on a %dx%d image.\n",
               levels, wname, wmaxlev, Nc, Nr);
        printf("Forcing nlevels = %d\n", wmaxlev);
        levels = wmaxlev;
    }
    if (levels > PWT_MAX_LEVELS) levels = PWT_MAX_LEVELS;
    p->nlevels = levels;
    if (p->do_cs && p->ndims == 1) {                                // wt.cu:179-183
        free(p);
        return fail64(PWT_ERR_UNSUPPORTED, "cycle spinning is not implemented for 1D. Use SWT instead.");
    }
    // geometry (common.cu:400-445)
    const int L = p->nlevels;
    p->lvNr[0] = Nr; p->lvNc[0] = Nc;
    for (int l = 1; l <= L; l++) {
        p->lvNr[l] = p->do_swt ? Nr : (p->ndims == 2 ? div2i(p->lvNr[l - 1]) : Nr);
        p->lvNc[l] = p->do_swt ? Nc : div2i(p->lvNc[l - 1]);
    }
    const int per = p->ndims == 2 ? 3 : 1;
    p->nbands = per * L + 1;
    for (int i = 0; i < L; i++)
        for (int j = 1; j <= per; j++) {
            p->band_nr[per * i + j] = p->lvNr[i + 1];
            p->band_nc[per * i + j] = p->lvNc[i + 1];
        }
    p->band_nr[0] = p->lvNr[L];
    p->band_nc[0] = p->lvNc[L];

    int rc = PWT_OK;
    cudaError_t e = cudaStreamCreate(&p->stream);
    if (e == cudaSuccess) e = cudaEventCreate(&p->ev0);
    if (e == cudaSuccess) e = cudaEventCreate(&p->ev1);
    if (e != cudaSuccess) rc = fail64(PWT_ERR_CUDA, "stream/event creation failed: %s", cudaGetErrorString(e));
    if (rc == PWT_OK) {
        const size_t B = (size_t)batch, img_n = align32(B * img_elems(p));
        size_t total = img_n, off[PWT_MAX_BANDS];
        for (int b = 0; b < p->nbands; b++) {                       // band 0 keeps the level-1 size (ping-pong plane)
            off[b] = total;
            total += align32(b == 0 ? B * (size_t)lvl_elems(p, 1) : B * (size_t)band_elems(p, b));
        }
        const size_t tmp_off = total;
        total += (p->do_swt && p->ndims == 2) ? 3 * img_n : img_n;
        const size_t tmp2_off = total;
        if (!p->do_swt && p->ndims == 2) total += align32(B * (size_t)p->Nr * (size_t)(p->Nc + 2));
        e = cudaMalloc((void**)&p->slab, total * sizeof(double));
        if (e == cudaSuccess) e = cudaMemsetAsync(p->slab, 0, total * sizeof(double), p->stream);
        if (e == cudaSuccess) e = cudaMalloc((void**)&p->d_acc, 2 * sizeof(double));
        if (e == cudaSuccess) e = cudaMallocHost((void**)&p->h_acc, 2 * sizeof(double));
        if (e != cudaSuccess) rc = fail64(PWT_ERR_CUDA, "allocation of %zu MB failed: %s", total * 8 >> 20, cudaGetErrorString(e));
        else {
            p->d_image = p->slab;
            for (int b = 0; b < p->nbands; b++) p->d_band[b] = p->slab + off[b];
            p->d_tmp = p->slab + tmp_off;
            p->d_tmp2 = (!p->do_swt && p->ndims == 2) ? p->slab + tmp2_off : nullptr;
        }
    }
    if (rc == PWT_OK && img) {
        e = cudaMemcpyAsync(p->d_image, img, (size_t)batch * img_elems(p) * sizeof(double),
                            memisonhost ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice, p->stream);
        if (e != cudaSuccess) rc = fail64(PWT_ERR_CUDA, "image upload failed: %s", cudaGetErrorString(e));
    }
    if (rc == PWT_OK && cudaStreamSynchronize(p->stream) != cudaSuccess) rc = fail64(PWT_ERR_CUDA, "plan initialisation failed");
    if (rc != PWT_OK) {
        pwt64_destroy(p);
        return rc;
    }
    *out = p;
    return PWT_OK;
}

extern "C" int pwt64_get_info(const pwt64_plan* p, pwt_info* info) {
    if (!p || !info) return fail64(PWT_ERR_ARG, "null argument");
    info->batch = p->batch; info->Nr = p->Nr; info->Nc = p->Nc; info->ndims = p->ndims; info->nlevels = p->nlevels;
    info->hlen = p->hlen; info->do_swt = p->do_swt; info->do_separable = p->do_separable; info->do_cycle_spinning = p->do_cs;
    info->state = p->state; info->shift_r = p->shift_r; info->shift_c = p->shift_c; info->nbands = p->nbands;
    info->device = p->device;
    return PWT_OK;
}
extern "C" int pwt64_band_shape(const pwt64_plan* p, int num, int* nr, int* nc) {
    if (!p || num < 0 || num >= p->nbands) return fail64(PWT_ERR_ARG, "bad band index");
    if (nr) *nr = p->band_nr[num];
    if (nc) *nc = p->band_nc[num];
    return PWT_OK;
}

// Wavelets::forward wt.cu:236-269
extern "C" int pwt64_forward(pwt64_plan* p) {
    if (!p) return fail64(PWT_ERR_ARG, "null plan");
    cudaSetDevice(p->device);
    if (p->do_cs) {                                                 // wt.cu:242-246 (same libc rand() stream as the fp32 plans)
        p->shift_r = rand() % p->Nr;
        p->shift_c = rand() % p->Nc;
        int rc = circshift64(p, p->shift_r, p->shift_c);
        if (rc != PWT_OK) return rc;
    }
    const int L = p->nlevels, B = p->batch;
    const bool haar = is_haar(p);
    const bool swap = !p->do_separable && !haar;                    // non-separable slot order (Q1)
    const int sH = swap ? 2 : 1, sV = swap ? 1 : 2;
    cudaStream_t st = p->stream;
    const double* src = p->d_image;
    const long long plane = (long long)B * img_elems(p);
    for (int l = 1; l <= L; l++) {
        double* alt = (p->do_swt && p->ndims == 2) ? p->d_tmp + 2 * plane : p->d_tmp;
        double* dstA = ((L - l) & 1) ? alt : p->d_band[0];          // A_L lands in band 0 without a fix-up copy
        if (p->ndims == 1) {
            const int rows = B * p->Nr;
            if (p->do_swt) p->launches += pwt_launch_swt_fwd1d_f64(src, dstA, p->d_band[l], rows, p->Nc, l, p->filt, st);
            else {
                int n = haar ? 0 : level_fwd1d_rows(src, dstA, p->d_band[l], rows, p->lvNc[l - 1], p->filt, st);
                if (!n) n = pwt_launch_dwt_fwd1d_f64(src, dstA, p->d_band[l], rows, p->lvNc[l - 1], p->filt, haar, st);
                p->launches += n;
            }
        } else {
            double* Hb = p->d_band[3 * (l - 1) + sH];
            double* V = p->d_band[3 * (l - 1) + sV];
            double* D = p->d_band[3 * (l - 1) + 3];
            if (p->do_swt) {
                int n = pwt_swt2p_fwd2d_f64(src, dstA, Hb, V, D, p->d_tmp, B, p->Nr, p->Nc, l, p->filt, st);
                if (!n) n = pwt_launch_swt_fwd2d_f64(src, dstA, Hb, V, D, p->d_tmp, B, p->Nr, p->Nc, l, p->filt, st);
                p->launches += n;
            }
            else {
                int n = haar ? pwt64_haar_fwd2d(src, dstA, Hb, V, D, B, p->lvNr[l - 1], p->lvNc[l - 1], lvl_elems(p, l - 1), lvl_elems(p, l), st)
                             : pwt64_fused_fwd2d(src, dstA, Hb, V, D, B, p->lvNr[l - 1], p->lvNc[l - 1], lvl_elems(p, l - 1),
                                                 lvl_elems(p, l), p->filt, st);
                if (!n && !haar) n = level_fwd2d_2pass(src, dstA, Hb, V, D, p->d_tmp2, B, p->lvNr[l - 1], p->lvNc[l - 1],
                                                       lvl_elems(p, l - 1), lvl_elems(p, l), p->filt, st);
                if (!n) n = pwt_launch_dwt_fwd2d_f64(src, dstA, Hb, V, D, B, p->lvNr[l - 1], p->lvNc[l - 1],
                                                     lvl_elems(p, l - 1), lvl_elems(p, l), p->filt, haar, st);
                p->launches += n;
            }
        }
        src = dstA;
    }
    CK64(cudaGetLastError());
    p->state = PWT_FORWARD;
    return PWT_OK;
}

// Wavelets::inverse wt.cu:271-305
extern "C" int pwt64_inverse(pwt64_plan* p) {
    if (!p) return fail64(PWT_ERR_ARG, "null plan");
    if (p->state == PWT_INVERSE) {
        puts("Warning: W.inverse() has already been run. Inverse is available in W.get_image()");
        return 1;
    }
    cudaSetDevice(p->device);
    const int L = p->nlevels, B = p->batch;
    const bool haar = is_haar(p);
    const bool swap = !p->do_separable && !haar;
    const int sH = swap ? 2 : 1, sV = swap ? 1 : 2;
    cudaStream_t st = p->stream;
    const double* cur = p->d_band[0];
    const long long plane = (long long)B * img_elems(p);
    for (int l = L; l >= 1; l--) {
        double* alt = (p->do_swt && p->ndims == 2) ? p->d_tmp + 2 * plane : p->d_tmp;
        double* dst = (l == 1) ? p->d_image : (cur == p->d_band[0] ? alt : p->d_band[0]);
        if (p->ndims == 1) {
            const int rows = B * p->Nr;
            if (p->do_swt) p->launches += pwt_launch_swt_inv1d_f64(cur, p->d_band[l], dst, rows, p->Nc, l, p->filt, st);
            else {
                int n = haar ? 0 : level_inv1d_rows(cur, p->d_band[l], dst, rows, p->lvNc[l], p->lvNc[l - 1], p->filt, st);
                if (!n) n = pwt_launch_dwt_inv1d_f64(cur, p->d_band[l], dst, rows, p->lvNc[l], p->lvNc[l - 1], p->filt, haar, st);
                p->launches += n;
            }
        } else {
            const double* Hb = p->d_band[3 * (l - 1) + sH];
            const double* V = p->d_band[3 * (l - 1) + sV];
            const double* D = p->d_band[3 * (l - 1) + 3];
            if (p->do_swt) {
                int n = pwt_swt2p_inv2d_f64(cur, Hb, V, D, dst, p->d_tmp, B, p->Nr, p->Nc, l, p->filt, st);
                if (!n) n = pwt_launch_swt_inv2d_f64(cur, Hb, V, D, dst, p->d_tmp, B, p->Nr, p->Nc, l, p->filt, st);
                p->launches += n;
            }
            else {
                int n = haar ? pwt64_haar_inv2d(cur, Hb, V, D, dst, B, p->lvNr[l], p->lvNc[l], p->lvNr[l - 1], p->lvNc[l - 1],
                                                lvl_elems(p, l), lvl_elems(p, l - 1), st)
                             : pwt64_fused_inv2d(cur, Hb, V, D, dst, B, p->lvNr[l], p->lvNc[l], p->lvNr[l - 1], p->lvNc[l - 1],
                                                 lvl_elems(p, l), lvl_elems(p, l - 1), p->filt, st);
                if (!n && !haar) n = level_inv2d_2pass(cur, Hb, V, D, dst, p->d_tmp2, B, p->lvNr[l], p->lvNc[l], p->lvNr[l - 1],
                                                       p->lvNc[l - 1], lvl_elems(p, l), lvl_elems(p, l - 1), p->filt, st);
                if (!n) n = pwt_launch_dwt_inv2d_f64(cur, Hb, V, D, dst, B, p->lvNr[l], p->lvNc[l], p->lvNr[l - 1],
                                                     p->lvNc[l - 1], lvl_elems(p, l), lvl_elems(p, l - 1), p->filt, haar, st);
                p->launches += n;
            }
        }
        cur = dst;
    }
    CK64(cudaGetLastError());
    if (p->do_cs) {                                                 // wt.cu:303
        int rc = circshift64(p, -p->shift_r, -p->shift_c);
        if (rc != PWT_OK) return rc;
    }
    p->state = PWT_INVERSE;
    return PWT_OK;
}

// ---- custom filter banks (wt.cu:558-600 with DTYPE = double; separable banks) --------------------------------------
// Odd lengths are mapped onto the windows the reference's kernels read (separable.cu:98-102, :252-264), like the fp32
// plans do (pwt_plan.cu: load_taps): analysis taps padded in front, synthesis taps of the decimated transform with the first
// tap dropped, of the stationary transform padded at the back.
namespace {
enum { PAD64_FRONT = 0, PAD64_DROP0 = 1, PAD64_BACK = 2 };
void load_taps64(double* dst, const double* src, unsigned len, unsigned padded, int mode) {
    memset(dst, 0, PWT_MAX_TAPS * sizeof(double));
    if (padded == len) mode = PAD64_FRONT;
    const unsigned o = mode == PAD64_FRONT ? padded - len : 0;
    for (unsigned k = (mode == PAD64_DROP0 ? 1 : 0); k < len; k++) dst[o + k] = src[k];
}
}  // namespace
extern "C" int pwt64_set_filters_forward(pwt64_plan* p, const char* name, unsigned len, const double* lo, const double* hi) {
    if (!p || !lo || !hi || len < 2) return fail64(PWT_ERR_ARG, "bad argument");
    const unsigned padded = len + (len & 1);
    if (padded > PWT_MAX_TAPS) {                                    // wt.cu:560-563
        printf("ERROR: Wavelets.set_filters_forward(): filter length (%d) exceeds the maximum size (%d)\n", len, PWT_MAX_TAPS);
        return -1;
    }
    if (!p->do_separable) {
        puts("ERROR: Wavelets.set_filters_forward(): the double-precision plans take separable banks only");
        return -2;
    }
    load_taps64(p->filt.L, lo, len, padded, PAD64_FRONT);
    load_taps64(p->filt.H, hi, len, padded, PAD64_FRONT);
    p->hlen = (int)padded;
    p->filt.hlen = (int)padded;
    p->custom_len = (int)len;
    if (name) {
        memset(p->wname, 0, sizeof(p->wname));
        strncpy(p->wname, name, sizeof(p->wname) - 1);
    }
    return PWT_OK;
}
extern "C" int pwt64_set_filters_inverse(pwt64_plan* p, const double* lo, const double* hi) {
    if (!p || !lo || !hi) return fail64(PWT_ERR_ARG, "bad argument");
    if (!p->do_separable) {
        puts("ERROR: Wavelets.set_filters_inverse(): the double-precision plans take separable banks only");
        return -2;
    }
    const unsigned padded = (unsigned)p->hlen;                      // the length given to set_filters_forward (wt.cu:587)
    const unsigned len = p->custom_len ? (unsigned)p->custom_len : padded;
    const int mode = p->do_swt ? PAD64_BACK : PAD64_DROP0;
    load_taps64(p->filt.IL, lo, len, padded, mode);
    load_taps64(p->filt.IH, hi, len, padded, mode);
    return PWT_OK;
}

// ---- thresholds / shrink (common.cu:219-282, 347-371 with DTYPE = double) -----------------------
namespace {
const double kSqrt2d = 1.4142135623730951;
template <int OP>
void launch_op(pwt64_plan* p, double* ptr, long long n, double beta) {
    if (n <= 0) return;
    k64_eltwise<OP><<<grid_for(n), 256, 0, p->stream>>>(ptr, n, beta);
    p->launches++;
}
int run_op(pwt64_plan* p, int op, double beta, int app, int normalize, const char* what) {
    if (!p) return fail64(PWT_ERR_ARG, "null plan");
    if (p->state == PWT_INVERSE) {                                  // wt.cu:309-312
        printf("Warning: Wavelets(): cannot %s coefficients, as they were modified by W.inverse()\n", what);
        return 1;
    }
    cudaSetDevice(p->device);
    const long long B = p->batch;
    const int L = p->nlevels, per = p->ndims == 2 ? 3 : 1;
    auto apply = [&](double* ptr, long long n, double b) {
        if (op == OP_SOFT) launch_op<OP_SOFT>(p, ptr, n, b);
        else if (op == OP_HARD) launch_op<OP_HARD>(p, ptr, n, b);
        else launch_op<OP_SCALE>(p, ptr, n, b);
    };
    if (op == OP_SCALE) {                                           // common.cu:355: every band *= 1 / (1 + beta)
        const double f = 1.0 / (1.0 + beta);
        if (app) apply(p->d_band[0], B * band_elems(p, 0), f);
        for (int b = 1; b < p->nbands; b++) apply(p->d_band[b], B * band_elems(p, b), f);
    } else {
        if (app) {
            double beta2 = beta;
            if (normalize > 0 && op == OP_SOFT) {                   // hard: the unscaled beta is passed (common.cu:264-270)
                const int n2 = L / 2;
                beta2 /= (double)(1 << n2);
                if (n2 * 2 != L) beta2 /= kSqrt2d;
            }
            apply(p->d_band[0], B * band_elems(p, 0), beta2);
        }
        for (int i = 0; i < L; i++) {
            if (normalize > 0) beta /= kSqrt2d;
            for (int j = 1; j <= per; j++) apply(p->d_band[per * i + j], B * band_elems(p, per * i + j), beta);
        }
    }
    CK64(cudaGetLastError());
    return PWT_OK;
}
}  // namespace
extern "C" int pwt64_soft_threshold(pwt64_plan* p, double beta, int app, int normalize) { return run_op(p, OP_SOFT, beta, app, normalize, "threshold"); }
extern "C" int pwt64_hard_threshold(pwt64_plan* p, double beta, int app, int normalize) { return run_op(p, OP_HARD, beta, app, normalize, "threshold"); }
extern "C" int pwt64_shrink(pwt64_plan* p, double beta, int app) { return run_op(p, OP_SCALE, beta, app, 0, "shrink"); }

// Wavelets::norm1 / norm2sq wt.cu:368-416 (1D: the true sum of squares, like the fp32 plan)
extern "C" int pwt64_norms(pwt64_plan* p, double* n1, double* n2) {
    if (!p) return fail64(PWT_ERR_ARG, "null plan");
    cudaSetDevice(p->device);
    CK64(cudaMemsetAsync(p->d_acc, 0, 2 * sizeof(double), p->stream));
    for (int b = 0; b < p->nbands; b++) {
        const long long n = (long long)p->batch * band_elems(p, b);
        k64_norms<<<grid_for(n), 256, 0, p->stream>>>(p->d_band[b], n, p->d_acc);
        p->launches++;
    }
    CK64(cudaMemcpyAsync(p->h_acc, p->d_acc, 2 * sizeof(double), cudaMemcpyDeviceToHost, p->stream));
    CK64(cudaStreamSynchronize(p->stream));
    if (n1) *n1 = p->h_acc[0];
    if (n2) *n2 = p->h_acc[1];
    return PWT_OK;
}

// ---- data in / out (wt.cu:419-506) ----------------------------------------------------------------
extern "C" int pwt64_get_image(pwt64_plan* p, double* dst) {
    if (!p || !dst) return 0;
    cudaSetDevice(p->device);
    const size_t n = (size_t)p->batch * img_elems(p);
    if (cudaMemcpyAsync(dst, p->d_image, n * sizeof(double), cudaMemcpyDeviceToHost, p->stream) != cudaSuccess ||
        cudaStreamSynchronize(p->stream) != cudaSuccess) {
        fail64(PWT_ERR_CUDA, "get_image failed: %s", cudaGetErrorString(cudaGetLastError()));
        return 0;
    }
    return n > 0x7fffffff ? 0x7fffffff : (int)n;
}
extern "C" int pwt64_set_image(pwt64_plan* p, const double* img, int on_device) {
    if (!p || !img) return fail64(PWT_ERR_ARG, "null argument");
    cudaSetDevice(p->device);
    const size_t n = (size_t)p->batch * img_elems(p);
    CK64(cudaMemcpyAsync(p->d_image, img, n * sizeof(double), on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, p->stream));
    if (!on_device) CK64(cudaStreamSynchronize(p->stream));
    p->state = PWT_INIT;
    return PWT_OK;
}
extern "C" int pwt64_get_coeff(pwt64_plan* p, double* dst, int num) {
    if (!p || !dst || num < 0 || num >= p->nbands) return 0;
    if (p->state == PWT_INVERSE) {                                  // wt.cu:474-477
        puts("Warning: get_coeff(): inverse() has been performed, the coefficients has been modified and do not make sense anymore.");
        return 0;
    }
    cudaSetDevice(p->device);
    const size_t n = (size_t)p->batch * band_elems(p, num);
    if (cudaMemcpyAsync(dst, p->d_band[num], n * sizeof(double), cudaMemcpyDeviceToHost, p->stream) != cudaSuccess ||
        cudaStreamSynchronize(p->stream) != cudaSuccess) {
        fail64(PWT_ERR_CUDA, "get_coeff failed: %s", cudaGetErrorString(cudaGetLastError()));
        return 0;
    }
    return n > 0x7fffffff ? 0x7fffffff : (int)n;
}
extern "C" int pwt64_set_coeff(pwt64_plan* p, const double* src, int num, int on_device) {
    if (!p || !src || num < 0 || num >= p->nbands) return fail64(PWT_ERR_ARG, "bad argument");
    cudaSetDevice(p->device);
    const size_t n = (size_t)p->batch * band_elems(p, num);
    CK64(cudaMemcpyAsync(p->d_band[num], src, n * sizeof(double), on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, p->stream));
    if (!on_device) CK64(cudaStreamSynchronize(p->stream));
    return PWT_OK;
}
extern "C" intptr_t pwt64_image_ptr(pwt64_plan* p) { return p ? (intptr_t)p->d_image : 0; }
extern "C" intptr_t pwt64_coeff_ptr(pwt64_plan* p, int num) { return (p && num >= 0 && num < p->nbands) ? (intptr_t)p->d_band[num] : 0; }
extern "C" int pwt64_sync(pwt64_plan* p) {
    if (!p) return fail64(PWT_ERR_ARG, "null plan");
    cudaSetDevice(p->device);
    CK64(cudaStreamSynchronize(p->stream));
    return PWT_OK;
}
extern "C" int pwt64_timer_start(pwt64_plan* p) {
    if (!p) return fail64(PWT_ERR_ARG, "null plan");
    cudaSetDevice(p->device);
    CK64(cudaEventRecord(p->ev0, p->stream));
    return PWT_OK;
}
extern "C" int pwt64_timer_stop(pwt64_plan* p, float* ms) {
    if (!p || !ms) return fail64(PWT_ERR_ARG, "null argument");
    cudaSetDevice(p->device);
    CK64(cudaEventRecord(p->ev1, p->stream));
    CK64(cudaEventSynchronize(p->ev1));
    CK64(cudaEventElapsedTime(ms, p->ev0, p->ev1));
    return PWT_OK;
}
extern "C" long long pwt64_launch_count(const pwt64_plan* p) { return p ? p->launches : 0; }
extern "C" int pwt64_lookup_filters(const char* wname, double* L, double* H, double* IL, double* IH) {
    PwtFilters64 f;
    if (!wname) return PWT_ERR_ARG;
    const int hlen = pwt_fill_filters64(wname, &f);
    if (hlen < 0) return hlen;
    for (int k = 0; k < hlen; k++) {
        if (L) L[k] = f.L[k];
        if (H) H[k] = f.H[k];
        if (IL) IL[k] = f.IL[k];
        if (IH) IH[k] = f.IH[k];
    }
    return hlen;
}
