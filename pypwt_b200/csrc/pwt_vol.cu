// Volumetric (3D) separable DWT -- SURVEY 8f rank 4.  The reference stops at 2D ("3D is not handled",
// pdwt/README.md:29; pypwt.pyx:155-156 rejects 3D input); this is the natural extension of its transform to volumes with
// the same conventions: periodisation over the size rounded up to even (separable.cu:98-102), ceil halving of every axis
// (utils.cu:24-27), level clip ilog2(min(Nz, Ny, Nx) / (F - 1)) (wt.cu:156-165), x filtered first, then y, then z.
// The result equals pywt.wavedecn(mode="periodization") on the same volume.
//
// One level = (1) the batched 2D level of the fp32 plans over the Nz slices of the current approximation -- the same
// kernels, chosen by pwt_level_fwd2d -- into four half-resolution sub-volumes, then (2) a pass along z of each
// sub-volume (k_vol_z_fwd below: threads along the contiguous x-y plane, 128-bit accesses, (low, high) tap pairs as
// FFMA2) that produces the eight bands.  Band index b = 4 dz + 2 dy + dx (d = 1: high-pass along that axis), so
// b = 0 is the approximation 'aaa' and b = 1 .. 7 are 'aad', 'ada', 'add', 'daa', 'dad', 'dda', 'ddd' in pywt's (z, y, x)
// key order.  Algorithmic traffic 8 B/voxel per direction; this design moves 16 (x-y pass + z pass) per level.
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/pwt_b200.h"
#include "pwt_internal.h"

int pwt_is_haar_alias(const char* wname);
int pwt_fill_filters(const char* wname, PwtFilters* out);
// kernels_vol_fused.cu: one level, x + y + z in one launch (F = 4, 6); bands[b], b = 4 dz + 2 dy + dx.  0: not covered
int pwt_vol_fused_fwd(const float* in, float* const* bands, int Nz, int Ny, int Nx, const PwtFilters& f, cudaStream_t st);
int pwt_vol_fused_inv(const float* const* bands, float* out, int nz2, int ny2, int nx2, int Nz, int Ny, int Nx, const PwtFilters& f,
                      cudaStream_t st);

namespace {
int failv(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    return pwt_set_error(code, buf);
}
#define CKV(call)                                                                                   \
    do {                                                                                            \
        cudaError_t e_ = (call);                                                                    \
        if (e_ != cudaSuccess) return failv(PWT_ERR_CUDA, "%s failed: %s", #call, cudaGetErrorString(e_)); \
    } while (0)
inline int div2i(int n) { return (n + 1) >> 1; }
inline int ilog2i(int i) {
    int l = 0;
    while (i > 1) { i >>= 1; ++l; }
    return l;
}
inline size_t align64(size_t n) { return (n + 63) & ~(size_t)63; }

__device__ __forceinline__ float2 fma2s(float x, float2 t, float2 acc) { return __ffma2_rn(make_float2(x, x), t, acc); }
__device__ __forceinline__ int wrap_dwt(int i, int N) {          // period N rounded up to even, x~[N] = x[N-1] (odd N)
    const int Ne = N + (N & 1);
    i %= Ne;
    if (i < 0) i += Ne;
    return i >= N ? N - 1 : i;
}
__device__ __forceinline__ int wrap_per(int i, int N) {
    i %= N;
    return i < 0 ? i + N : i;
}

// ---- z pass, analysis: in [Nz][P] -> lo, hi [ceil(Nz/2)][P];  VEC = floats per thread (4: P % 4 == 0) ----------
// lo[k] = sum_m L[m] * x~[(2k + F/2 - m) mod Ne]  (separable.cu:135-176 applied along z)
template <int VEC>
__global__ void __launch_bounds__(256)
k_vol_z_fwd(const float* __restrict__ in, float* __restrict__ lo, float* __restrict__ hi, int Nz, long long P, int F,
            const __grid_constant__ PwtTapsFwd tp, int haar) {
    const int Nz2 = (Nz + 1) >> 1;
    const long long PV = P / VEC;
    pwt_pdl_wait();
    for (long long i = blockIdx.x * 256LL + threadIdx.x; i < PV * Nz2; i += gridDim.x * 256LL) {
        const int k = (int)(i / PV);
        const long long p = (i - (long long)k * PV) * VEC;
        float2 acc[VEC];
#pragma unroll
        for (int v = 0; v < VEC; v++) acc[v] = make_float2(0.f, 0.f);
        if (haar) {                                                 // haar.cu:10-42: plain sums, one scale
            const float* r0 = in + (long long)wrap_dwt(2 * k, Nz) * P + p;
            const float* r1 = in + (long long)wrap_dwt(2 * k + 1, Nz) * P + p;
#pragma unroll
            for (int v = 0; v < VEC; v++) {
                const float a = __ldg(r0 + v), b = __ldg(r1 + v);
                acc[v] = make_float2(0.70710678118654746f * (a + b), 0.70710678118654746f * (a - b));
            }
        } else {
            const int c = F / 2 - 1;
            for (int j = 0; j < F; j++) {
                const float* r = in + (long long)wrap_dwt(2 * k - c + j, Nz) * P + p;
                if (VEC == 4) {
                    const float4 x = __ldg(reinterpret_cast<const float4*>(r));
                    acc[0] = fma2s(x.x, tp.t[j], acc[0]);
                    acc[1] = fma2s(x.y, tp.t[j], acc[1]);
                    acc[2] = fma2s(x.z, tp.t[j], acc[2]);
                    acc[3] = fma2s(x.w, tp.t[j], acc[3]);
                } else {
                    acc[0] = fma2s(__ldg(r), tp.t[j], acc[0]);
                }
            }
        }
        float* ol = lo + (long long)k * P + p;
        float* oh = hi + (long long)k * P + p;
        if (VEC == 4) {
            *reinterpret_cast<float4*>(ol) = make_float4(acc[0].x, acc[1].x, acc[2].x, acc[3].x);
            *reinterpret_cast<float4*>(oh) = make_float4(acc[0].y, acc[1].y, acc[2].y, acc[3].y);
        } else {
            ol[0] = acc[0].x;
            oh[0] = acc[0].y;
        }
    }
}

// ---- z pass, synthesis: lo, hi [nz2][P] -> out [Nz_out][P] (polyphase, separable.cu:293-328 along z) --------------
// x[n] = sum_t a[k] IL[t] + d[k] IH[t] over taps t with n + F/2 - 1 - t even, k = ((n + F/2 - 1 - t) / 2) mod nz2
template <int VEC>
__global__ void __launch_bounds__(256)
k_vol_z_inv(const float* __restrict__ lo, const float* __restrict__ hi, float* __restrict__ out, int nz2, int Nz_out,
            long long P, int F, const __grid_constant__ PwtFilters f, int haar) {
    const long long PV = P / VEC;
    pwt_pdl_wait();
    for (long long i = blockIdx.x * 256LL + threadIdx.x; i < PV * Nz_out; i += gridDim.x * 256LL) {
        const int n = (int)(i / PV);
        const long long p = (i - (long long)n * PV) * VEC;
        float r[VEC];
#pragma unroll
        for (int v = 0; v < VEC; v++) r[v] = 0.f;
        if (haar) {
            const float* a = lo + (long long)(n >> 1) * P + p;
            const float* d = hi + (long long)(n >> 1) * P + p;
#pragma unroll
            for (int v = 0; v < VEC; v++)
                r[v] = 0.70710678118654746f * ((n & 1) ? __ldg(a + v) - __ldg(d + v) : __ldg(a + v) + __ldg(d + v));
        } else {
            const int pp = F / 2 - 1, b = n & 1, t0 = (b + pp) & 1, kb = (n >> 1) + ((b + pp) >> 1);
            for (int j = 0; j < F / 2; j++) {
                const long long row = (long long)wrap_per(kb - j, nz2) * P + p;
                const float tl = f.IL[2 * j + t0], th = f.IH[2 * j + t0];
                if (VEC == 4) {
                    const float4 a = __ldg(reinterpret_cast<const float4*>(lo + row));
                    const float4 d = __ldg(reinterpret_cast<const float4*>(hi + row));
                    r[0] = fmaf(a.x, tl, r[0]); r[0] = fmaf(d.x, th, r[0]);
                    r[1] = fmaf(a.y, tl, r[1]); r[1] = fmaf(d.y, th, r[1]);
                    r[2] = fmaf(a.z, tl, r[2]); r[2] = fmaf(d.z, th, r[2]);
                    r[3] = fmaf(a.w, tl, r[3]); r[3] = fmaf(d.w, th, r[3]);
                } else {
                    r[0] = fmaf(__ldg(lo + row), tl, r[0]);
                    r[0] = fmaf(__ldg(hi + row), th, r[0]);
                }
            }
        }
        float* o = out + (long long)n * P + p;
        if (VEC == 4) *reinterpret_cast<float4*>(o) = make_float4(r[0], r[1], r[2], r[3]);
        else o[0] = r[0];
    }
}

// ---- z pass with a sliding register window (planes that are multiples of 4 floats) -----
// (compile-time F = 2 .. 20.)  A thread owns 4 adjacent columns and a run of KS consecutive outputs: the F input rows of an output live in a rotating
// window of F float4 registers, each new output loads TWO new rows (analysis) / ONE new row of each band (synthesis)
// instead of re-reading F (F/2 + 1) rows through L1/L2.  The rotation is unrolled over its period so that every window
// index is static.  The four sub-volumes of a level go in ONE launch (blockIdx.y).
struct ZJobs {
    const float* a[4];     // analysis: input sub-volume;  synthesis: low-pass band
    const float* b[4];     // synthesis: high-pass band
    float* o0[4];          // analysis: low-pass output;   synthesis: output sub-volume
    float* o1[4];          // analysis: high-pass output
};
template <int F>
__global__ void __launch_bounds__(128)
k_vol_z_fwd_win(const __grid_constant__ ZJobs jb, int Nz, long long P, int KS, const __grid_constant__ PwtTapsFwd tp) {
    constexpr int C = F / 2 - 1;
    const float* __restrict__ in = jb.a[blockIdx.y];
    float* __restrict__ lo = jb.o0[blockIdx.y];
    float* __restrict__ hi = jb.o1[blockIdx.y];
    const int Nz2 = (Nz + 1) >> 1;
    const long long PV = P >> 2;
    const int nseg = (Nz2 + KS - 1) / KS;
    pwt_pdl_wait();
    for (long long i = blockIdx.x * 128LL + threadIdx.x; i < PV * nseg; i += gridDim.x * 128LL) {
        const int seg = (int)(i / PV);
        const long long p = (i - (long long)seg * PV) << 2;
        const int k0 = seg * KS, kend = min(k0 + KS, Nz2);
        float4 w[F];
#pragma unroll
        for (int j = 0; j < F - 2; j++) w[j] = __ldg(reinterpret_cast<const float4*>(in + (long long)wrap_dwt(2 * k0 - C + j, Nz) * P + p));
        for (int kb = k0; kb < kend; kb += F / 2) {
#pragma unroll
            for (int u = 0; u < F / 2; u++) {
                const int k = kb + u;
                if (k < kend) {
                    w[(F - 2 + 2 * u) % F] = __ldg(reinterpret_cast<const float4*>(in + (long long)wrap_dwt(2 * k - C + F - 2, Nz) * P + p));
                    w[(F - 1 + 2 * u) % F] = __ldg(reinterpret_cast<const float4*>(in + (long long)wrap_dwt(2 * k - C + F - 1, Nz) * P + p));
                    float2 a0 = make_float2(0.f, 0.f), a1 = a0, a2 = a0, a3 = a0;
#pragma unroll
                    for (int j = 0; j < F; j++) {
                        const float4 x = w[(j + 2 * u) % F];
                        a0 = fma2s(x.x, tp.t[j], a0);
                        a1 = fma2s(x.y, tp.t[j], a1);
                        a2 = fma2s(x.z, tp.t[j], a2);
                        a3 = fma2s(x.w, tp.t[j], a3);
                    }
                    __stcs(reinterpret_cast<float4*>(lo + (long long)k * P + p), make_float4(a0.x, a1.x, a2.x, a3.x));
                    __stcs(reinterpret_cast<float4*>(hi + (long long)k * P + p), make_float4(a0.y, a1.y, a2.y, a3.y));
                }
            }
        }
    }
}
// synthesis: output pair j (rows 2j, 2j + 1) = sum_w a[j - S1 + w] * l[w] + d[j - S1 + w] * h[w], (even, odd) tap pairs
template <int F>
__global__ void __launch_bounds__(128)
k_vol_z_inv_win(const __grid_constant__ ZJobs jb, int nz2, int Nz_out, long long P, int KS, const __grid_constant__ PwtTapsInv tp) {
    constexpr int S1 = (F / 2) >> 1, W = F / 2 + 1;
    const float* __restrict__ lo = jb.a[blockIdx.y];
    const float* __restrict__ hi = jb.b[blockIdx.y];
    float* __restrict__ out = jb.o0[blockIdx.y];
    const long long PV = P >> 2;
    const int nseg = (nz2 + KS - 1) / KS;
    pwt_pdl_wait();
    for (long long i = blockIdx.x * 128LL + threadIdx.x; i < PV * nseg; i += gridDim.x * 128LL) {
        const int seg = (int)(i / PV);
        const long long p = (i - (long long)seg * PV) << 2;
        const int j0 = seg * KS, jend = min(j0 + KS, nz2);
        float4 wa[W], wd[W];
#pragma unroll
        for (int w = 0; w < W - 1; w++) {
            const long long row = (long long)wrap_per(j0 - S1 + w, nz2) * P + p;
            wa[w] = __ldg(reinterpret_cast<const float4*>(lo + row));
            wd[w] = __ldg(reinterpret_cast<const float4*>(hi + row));
        }
        for (int jb0 = j0; jb0 < jend; jb0 += W) {
#pragma unroll
            for (int u = 0; u < W; u++) {
                const int j = jb0 + u;
                if (j < jend) {
                    const long long row = (long long)wrap_per(j - S1 + W - 1, nz2) * P + p;
                    wa[(W - 1 + u) % W] = __ldg(reinterpret_cast<const float4*>(lo + row));
                    wd[(W - 1 + u) % W] = __ldg(reinterpret_cast<const float4*>(hi + row));
                    float2 e0 = make_float2(0.f, 0.f), e1 = e0, e2 = e0, e3 = e0;   // (row 2j, row 2j + 1) of the 4 columns
#pragma unroll
                    for (int w = 0; w < W; w++) {
                        const float4 a = wa[(w + u) % W], d = wd[(w + u) % W];
                        e0 = fma2s(a.x, tp.l[w], e0); e0 = fma2s(d.x, tp.h[w], e0);
                        e1 = fma2s(a.y, tp.l[w], e1); e1 = fma2s(d.y, tp.h[w], e1);
                        e2 = fma2s(a.z, tp.l[w], e2); e2 = fma2s(d.z, tp.h[w], e2);
                        e3 = fma2s(a.w, tp.l[w], e3); e3 = fma2s(d.w, tp.h[w], e3);
                    }
                    __stcs(reinterpret_cast<float4*>(out + (long long)(2 * j) * P + p), make_float4(e0.x, e1.x, e2.x, e3.x));
                    if (2 * j + 1 < Nz_out)
                        __stcs(reinterpret_cast<float4*>(out + (long long)(2 * j + 1) * P + p), make_float4(e0.y, e1.y, e2.y, e3.y));
                }
            }
        }
    }
}
// outputs per thread run: long enough to amortise the F - 2 preloaded rows, short enough to fill the GPU
inline int pick_ks(int n_out, long long PV, int period) {
    const long long want_threads = 4LL * 148 * 1024;
    long long nseg = (want_threads + 4 * PV - 1) / (4 * PV);
    if (nseg < 1) nseg = 1;
    int ks = (int)((n_out + nseg - 1) / nseg);
    if (ks < 4 * period) ks = 4 * period;
    ks = ((ks + period - 1) / period) * period;
    return ks;
}
#define PWT_VOLZ_CASES(X) X(2) X(4) X(6) X(8) X(10) X(12) X(14) X(16) X(18) X(20)
// all four sub-volumes of a level in one launch; returns 0 when the window kernels do not cover the configuration
int launch_z_fwd4(const ZJobs& jb, int Nz, long long P, const PwtFilters& f, cudaStream_t st) {
    if ((P & 3) || f.hlen < 2 || f.hlen > 20 || (f.hlen & 1)) return 0;
    for (int q = 0; q < 4; q++)
        if ((((uintptr_t)jb.a[q]) | ((uintptr_t)jb.o0[q]) | ((uintptr_t)jb.o1[q])) & 15) return 0;
    const int Nz2 = (Nz + 1) >> 1;
    const PwtTapsFwd t = pwt_pack_taps_fwd(f, f.hlen);
    const int KS = pick_ks(Nz2, P >> 2, f.hlen / 2);
    const long long items = (P >> 2) * ((Nz2 + KS - 1) / KS);
    const dim3 grid((unsigned)((items + 127) / 128 < 65535 * 16 ? (items + 127) / 128 : 65535 * 16), 4);
    switch (f.hlen) {
#define X(FF) case FF: pwt_launch_pdl(k_vol_z_fwd_win<FF>, grid, 128, 0, st, jb, Nz, P, KS, t); return 1;
        PWT_VOLZ_CASES(X)
#undef X
    }
    return 0;
}
int launch_z_inv4(const ZJobs& jb, int nz2, int Nz_out, long long P, const PwtFilters& f, cudaStream_t st) {
    if ((P & 3) || f.hlen < 2 || f.hlen > 20 || (f.hlen & 1)) return 0;
    for (int q = 0; q < 4; q++)
        if ((((uintptr_t)jb.a[q]) | ((uintptr_t)jb.b[q]) | ((uintptr_t)jb.o0[q])) & 15) return 0;
    const PwtTapsInv t = pwt_pack_taps_inv(f, f.hlen);
    const int KS = pick_ks(nz2, P >> 2, f.hlen / 2 + 1);
    const long long items = (P >> 2) * ((nz2 + KS - 1) / KS);
    const dim3 grid((unsigned)((items + 127) / 128 < 65535 * 16 ? (items + 127) / 128 : 65535 * 16), 4);
    switch (f.hlen) {
#define X(FF) case FF: pwt_launch_pdl(k_vol_z_inv_win<FF>, grid, 128, 0, st, jb, nz2, Nz_out, P, KS, t); return 1;
        PWT_VOLZ_CASES(X)
#undef X
    }
    return 0;
}

inline unsigned grid_for(long long items) {
    long long g = (items + 255) / 256;
    const long long cap = (long long)pwt_sm_count() * 32;
    return (unsigned)(g < 1 ? 1 : (g > cap ? cap : g));
}
int launch_z_fwd(const float* in, float* lo, float* hi, int Nz, long long P, const PwtFilters& f, bool haar, cudaStream_t st) {
    const PwtTapsFwd t = pwt_pack_taps_fwd(f, f.hlen);
    const int Nz2 = (Nz + 1) >> 1;
    const bool vec = (P & 3) == 0 && ((((uintptr_t)in) | ((uintptr_t)lo) | ((uintptr_t)hi)) & 15) == 0;
    if (vec) pwt_launch_pdl(k_vol_z_fwd<4>, dim3(grid_for(P / 4 * Nz2)), 256, 0, st, in, lo, hi, Nz, P, f.hlen, t, haar ? 1 : 0);
    else pwt_launch_pdl(k_vol_z_fwd<1>, dim3(grid_for(P * Nz2)), 256, 0, st, in, lo, hi, Nz, P, f.hlen, t, haar ? 1 : 0);
    return 1;
}
int launch_z_inv(const float* lo, const float* hi, float* out, int nz2, int Nz_out, long long P, const PwtFilters& f, bool haar,
                 cudaStream_t st) {
    const bool vec = (P & 3) == 0 && ((((uintptr_t)out) | ((uintptr_t)lo) | ((uintptr_t)hi)) & 15) == 0;
    if (vec) pwt_launch_pdl(k_vol_z_inv<4>, dim3(grid_for(P / 4 * Nz_out)), 256, 0, st, lo, hi, out, nz2, Nz_out, P, f.hlen, f, haar ? 1 : 0);
    else pwt_launch_pdl(k_vol_z_inv<1>, dim3(grid_for(P * Nz_out)), 256, 0, st, lo, hi, out, nz2, Nz_out, P, f.hlen, f, haar ? 1 : 0);
    return 1;
}
}  // namespace

struct pwt3_plan {
    int device;
    cudaStream_t stream;
    cudaEvent_t ev0, ev1;
    int Nz, Ny, Nx, nlevels, hlen, haar, state;
    char wname[128];
    PwtFilters filt;
    int lz[PWT_MAX_LEVELS + 1], ly[PWT_MAX_LEVELS + 1], lx[PWT_MAX_LEVELS + 1];
    float* slab;
    float* d_image;
    float* d_sub[4];                 // x-y sub-volumes of the level being processed: (a, H, V, D), Nz_l x Ny_{l+1} x Nx_{l+1} each
    float* d_app[2];                 // ping-pong planes of the intermediate approximations
    float* d_band[PWT_MAX_LEVELS][8];// [level - 1][b], b = 1 .. 7 (b = 0 unused); the final approximation is d_A
    float* d_A;
    double* d_acc;
    double* h_acc;
    long long launches;
};
namespace {
inline long long vox(const pwt3_plan* p, int l) { return (long long)p->lz[l] * p->ly[l] * p->lx[l]; }
}

extern "C" void pwt3_destroy(pwt3_plan* p) {
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

extern "C" int pwt3_create(pwt3_plan** out, const float* vol, int Nz, int Ny, int Nx, const char* wname, int levels,
                           int memisonhost) {
    if (!out) return failv(PWT_ERR_ARG, "null output handle");
    *out = nullptr;
    if (!wname || Nz < 1 || Ny < 1 || Nx < 1) return failv(PWT_ERR_ARG, "invalid geometry %dx%dx%d", Nz, Ny, Nx);
    if ((long long)Ny * Nx >= (1LL << 31)) return failv(PWT_ERR_ARG, "one slice must hold < 2^31 samples");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1)
        return failv(PWT_ERR_CUDA, "no CUDA device available (this library has no CPU fallback)");
    pwt3_plan* p = (pwt3_plan*)calloc(1, sizeof(pwt3_plan));
    if (!p) return failv(PWT_ERR_NOMEM, "out of host memory");
    cudaGetDevice(&p->device);
    p->Nz = Nz; p->Ny = Ny; p->Nx = Nx;
    p->state = PWT_INIT;
    strncpy(p->wname, wname, sizeof(p->wname) - 1);
    const int hlen = pwt_fill_filters(wname, &p->filt);
    if (hlen < 0) {
        free(p);
        return failv(PWT_ERR_UNKNOWN_WAVELET, "unknown wavelet name '%s'", wname);
    }
    p->hlen = hlen;
    p->haar = hlen == 2;
    int N = Nz < Ny ? Nz : Ny;
    N = N < Nx ? N : Nx;
    const int wmaxlev = ilog2i(N / (hlen - 1));                     // wt.cu:156-165 over the three axes
    if (wmaxlev < 1) {
        free(p);
        return failv(PWT_ERR_TOO_SMALL, "a %dx%dx%d volume is too small for wavelet %s (%d taps)", Nz, Ny, Nx, wname, hlen);
    }
    if (levels < 1) levels = 1;
    if (levels > wmaxlev) {
        printf("Warning: required level (%d) is greater than the maximum possible level for %s (%d) on a %dx%dx%d volume.\n",
               levels, wname, wmaxlev, Nx, Ny, Nz);
        printf("Forcing nlevels = %d\n", wmaxlev);
        levels = wmaxlev;
    }
    if (levels > PWT_MAX_LEVELS) levels = PWT_MAX_LEVELS;
    p->nlevels = levels;
    p->lz[0] = Nz; p->ly[0] = Ny; p->lx[0] = Nx;
    for (int l = 1; l <= levels; l++) {
        p->lz[l] = div2i(p->lz[l - 1]);
        p->ly[l] = div2i(p->ly[l - 1]);
        p->lx[l] = div2i(p->lx[l - 1]);
    }
    int rc = PWT_OK;
    cudaError_t e = cudaStreamCreate(&p->stream);
    if (e == cudaSuccess) e = cudaEventCreate(&p->ev0);
    if (e == cudaSuccess) e = cudaEventCreate(&p->ev1);
    if (e != cudaSuccess) rc = failv(PWT_ERR_CUDA, "stream/event creation failed: %s", cudaGetErrorString(e));
    if (rc == PWT_OK) {
        const size_t nimg = align64((size_t)vox(p, 0));
        const size_t nsub = align64((size_t)Nz * p->ly[1] * p->lx[1]);      // largest x-y sub-volume (level 1)
        const size_t napp = align64((size_t)vox(p, 1));
        size_t total = nimg + 4 * nsub + 2 * napp;
        size_t off_band[PWT_MAX_LEVELS][8];
        for (int l = 1; l <= levels; l++)
            for (int b = 1; b < 8; b++) {
                off_band[l - 1][b] = total;
                total += align64((size_t)vox(p, l));
            }
        const size_t off_A = total;
        total += align64((size_t)vox(p, levels));
        e = cudaMalloc((void**)&p->slab, total * sizeof(float));
        if (e == cudaSuccess) e = cudaMemsetAsync(p->slab, 0, total * sizeof(float), p->stream);
        if (e == cudaSuccess) e = cudaMalloc((void**)&p->d_acc, 2 * sizeof(double));
        if (e == cudaSuccess) e = cudaMallocHost((void**)&p->h_acc, 2 * sizeof(double));
        if (e != cudaSuccess) rc = failv(PWT_ERR_CUDA, "allocation of %zu MB failed: %s", total * 4 >> 20, cudaGetErrorString(e));
        else {
            p->d_image = p->slab;
            for (int s = 0; s < 4; s++) p->d_sub[s] = p->slab + nimg + s * nsub;
            p->d_app[0] = p->slab + nimg + 4 * nsub;
            p->d_app[1] = p->d_app[0] + napp;
            for (int l = 1; l <= levels; l++)
                for (int b = 1; b < 8; b++) p->d_band[l - 1][b] = p->slab + off_band[l - 1][b];
            p->d_A = p->slab + off_A;
        }
    }
    if (rc == PWT_OK && vol) {
        e = cudaMemcpyAsync(p->d_image, vol, (size_t)vox(p, 0) * sizeof(float),
                            memisonhost ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice, p->stream);
        if (e != cudaSuccess) rc = failv(PWT_ERR_CUDA, "volume upload failed: %s", cudaGetErrorString(e));
    }
    if (rc == PWT_OK && cudaStreamSynchronize(p->stream) != cudaSuccess) rc = failv(PWT_ERR_CUDA, "plan initialisation failed");
    if (rc != PWT_OK) {
        pwt3_destroy(p);
        return rc;
    }
    *out = p;
    return PWT_OK;
}

extern "C" int pwt3_levels(const pwt3_plan* p) { return p ? p->nlevels : 0; }
// shape of the bands of level `level` (1 = finest); level = nlevels also gives the approximation's shape
extern "C" int pwt3_band_shape(const pwt3_plan* p, int level, int* nz, int* ny, int* nx) {
    if (!p || level < 1 || level > p->nlevels) return failv(PWT_ERR_ARG, "bad level");
    if (nz) *nz = p->lz[level];
    if (ny) *ny = p->ly[level];
    if (nx) *nx = p->lx[level];
    return PWT_OK;
}

extern "C" int pwt3_forward(pwt3_plan* p) {
    if (!p) return failv(PWT_ERR_ARG, "null plan");
    cudaSetDevice(p->device);
    cudaStream_t st = p->stream;
    const int L = p->nlevels;
    const float* src = p->d_image;
    for (int l = 1; l <= L; l++) {
        const int nz = p->lz[l - 1], ny = p->ly[l - 1], nx = p->lx[l - 1];
        const long long P = (long long)p->ly[l] * p->lx[l];
        {                                                           // x, y and z in one launch (short filters; Haar as a 2-tap bank)
            float* fb[8];
            fb[0] = l == L ? p->d_A : p->d_app[l & 1];
            for (int b = 1; b < 8; b++) fb[b] = p->d_band[l - 1][b];
            const int nf = pwt_vol_fused_fwd(src, fb, nz, ny, nx, p->filt, st);
            if (nf) {
                p->launches += nf;
                src = fb[0];
                continue;
            }
        }
        // (1) x then y on every slice: a = (Lx, Ly), H = (Lx, Hy), V = (Hx, Ly), D = (Hx, Hy)   [separable.cu:165-174]
        p->launches += pwt_level_fwd2d(src, p->d_sub[0], p->d_sub[1], p->d_sub[2], p->d_sub[3], nz, ny, nx, (long long)ny * nx, P,
                                       p->filt, p->haar, st);
        // (2) z: band b = 4 dz + 2 dy + dx.  a -> (0, 4), V (dx) -> (1, 5), H (dy) -> (2, 6), D -> (3, 7)
        float* dstA = l == L ? p->d_A : p->d_app[l & 1];
        float** B = p->d_band[l - 1];
        ZJobs jb = {};
        const int sub_of[4] = {0, 2, 1, 3};                         // job q reads a, V, H, D
        for (int q = 0; q < 4; q++) {
            jb.a[q] = p->d_sub[sub_of[q]];
            jb.o0[q] = q == 0 ? dstA : B[q];
            jb.o1[q] = B[4 + q];
        }
        int n = launch_z_fwd4(jb, nz, P, p->filt, st);
        if (!n)
            for (int q = 0; q < 4; q++) n += launch_z_fwd(jb.a[q], jb.o0[q], jb.o1[q], nz, P, p->filt, p->haar, st);
        p->launches += n;
        src = dstA;
    }
    CKV(cudaGetLastError());
    p->state = PWT_FORWARD;
    return PWT_OK;
}

extern "C" int pwt3_inverse(pwt3_plan* p) {
    if (!p) return failv(PWT_ERR_ARG, "null plan");
    if (p->state == PWT_INVERSE) {
        puts("Warning: W.inverse() has already been run. Inverse is available in W.get_image()");
        return 1;
    }
    cudaSetDevice(p->device);
    cudaStream_t st = p->stream;
    const int L = p->nlevels;
    const float* cur = p->d_A;
    for (int l = L; l >= 1; l--) {
        const int nz = p->lz[l - 1], ny = p->ly[l - 1], nx = p->lx[l - 1];
        const long long P = (long long)p->ly[l] * p->lx[l];
        float** B = p->d_band[l - 1];
        {                                                           // z, y and x in one launch (short filters; Haar as a 2-tap bank)
            const float* fb[8];
            fb[0] = cur;
            for (int b = 1; b < 8; b++) fb[b] = B[b];
            float* fdst = l == 1 ? p->d_image : p->d_app[(l - 1) & 1];
            const int nf = pwt_vol_fused_inv(fb, fdst, p->lz[l], p->ly[l], p->lx[l], nz, ny, nx, p->filt, st);
            if (nf) {
                p->launches += nf;
                cur = fdst;
                continue;
            }
        }
        ZJobs jb = {};
        const int sub_of[4] = {0, 2, 1, 3};
        for (int q = 0; q < 4; q++) {
            jb.a[q] = q == 0 ? cur : B[q];
            jb.b[q] = B[4 + q];
            jb.o0[q] = p->d_sub[sub_of[q]];
        }
        int n = launch_z_inv4(jb, p->lz[l], nz, P, p->filt, st);
        if (!n)
            for (int q = 0; q < 4; q++) n += launch_z_inv(jb.a[q], jb.b[q], jb.o0[q], p->lz[l], nz, P, p->filt, p->haar, st);
        p->launches += n;
        float* dst = l == 1 ? p->d_image : p->d_app[(l - 1) & 1];
        p->launches += pwt_level_inv2d(p->d_sub[0], p->d_sub[1], p->d_sub[2], p->d_sub[3], dst, nz, p->ly[l], p->lx[l], ny, nx, P,
                                       (long long)ny * nx, p->filt, p->haar, st);
        cur = dst;
    }
    CKV(cudaGetLastError());
    p->state = PWT_INVERSE;
    return PWT_OK;
}

namespace {
int run_thresh3(pwt3_plan* p, int op, float beta, int app) {
    if (!p) return failv(PWT_ERR_ARG, "null plan");
    if (p->state == PWT_INVERSE) {
        puts("Warning: Wavelets(): cannot threshold coefficients, as they were modified by W.inverse()");
        return 1;
    }
    cudaSetDevice(p->device);
    PwtSegTable t;
    t.nseg = 0;
    auto add = [&](float* ptr, long long n) {
        if (t.nseg >= PWT_MAX_SEGS) {
            p->launches += pwt_launch_eltwise(t, op, p->stream);
            t.nseg = 0;
        }
        t.seg[t.nseg].ptr = ptr; t.seg[t.nseg].n = n; t.seg[t.nseg].beta = beta; t.seg[t.nseg].pad = 0;
        t.nseg++;
    };
    if (app) add(p->d_A, vox(p, p->nlevels));
    for (int l = 1; l <= p->nlevels; l++)
        for (int b = 1; b < 8; b++) add(p->d_band[l - 1][b], vox(p, l));
    if (t.nseg) p->launches += pwt_launch_eltwise(t, op, p->stream);
    CKV(cudaGetLastError());
    return PWT_OK;
}
}  // namespace
extern "C" int pwt3_soft_threshold(pwt3_plan* p, float beta, int app) { return run_thresh3(p, PWT_OP_SOFT, beta, app); }
extern "C" int pwt3_hard_threshold(pwt3_plan* p, float beta, int app) { return run_thresh3(p, PWT_OP_HARD, beta, app); }

extern "C" int pwt3_norms(pwt3_plan* p, double* n1, double* n2) {
    if (!p) return failv(PWT_ERR_ARG, "null plan");
    cudaSetDevice(p->device);
    PwtSegTable t;
    t.nseg = 0;
    bool first = true;
    auto flush = [&]() {
        if (!t.nseg) return;
        if (first) p->launches += pwt_launch_norms(t, p->d_acc, p->stream);     // (zeroes the accumulators)
        first = false;
        t.nseg = 0;
    };
    auto add = [&](float* ptr, long long n) {
        t.seg[t.nseg].ptr = ptr; t.seg[t.nseg].n = n; t.seg[t.nseg].beta = 0.f; t.seg[t.nseg].pad = 0;
        t.nseg++;
    };
    if (7 * p->nlevels + 1 > PWT_MAX_SEGS) return failv(PWT_ERR_UNSUPPORTED, "norms: more than %d bands", PWT_MAX_SEGS);
    add(p->d_A, vox(p, p->nlevels));
    for (int l = 1; l <= p->nlevels; l++)
        for (int b = 1; b < 8; b++) add(p->d_band[l - 1][b], vox(p, l));
    flush();
    CKV(cudaMemcpyAsync(p->h_acc, p->d_acc, 2 * sizeof(double), cudaMemcpyDeviceToHost, p->stream));
    CKV(cudaStreamSynchronize(p->stream));
    if (n1) *n1 = p->h_acc[0];
    if (n2) *n2 = p->h_acc[1];
    return PWT_OK;
}

extern "C" int pwt3_get_image(pwt3_plan* p, float* dst) {
    if (!p || !dst) return failv(PWT_ERR_ARG, "null argument");
    cudaSetDevice(p->device);
    CKV(cudaMemcpyAsync(dst, p->d_image, (size_t)vox(p, 0) * sizeof(float), cudaMemcpyDeviceToHost, p->stream));
    CKV(cudaStreamSynchronize(p->stream));
    return PWT_OK;
}
extern "C" int pwt3_set_image(pwt3_plan* p, const float* vol, int on_device) {
    if (!p || !vol) return failv(PWT_ERR_ARG, "null argument");
    cudaSetDevice(p->device);
    CKV(cudaMemcpyAsync(p->d_image, vol, (size_t)vox(p, 0) * sizeof(float), on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, p->stream));
    if (!on_device) CKV(cudaStreamSynchronize(p->stream));
    p->state = PWT_INIT;
    return PWT_OK;
}
// band b of level `level` (b = 1 .. 7), or the approximation (level = nlevels, b = 0)
namespace {
float* band_ptr(pwt3_plan* p, int level, int b) {
    if (!p || level < 1 || level > p->nlevels || b < 0 || b > 7) return nullptr;
    if (b == 0) return level == p->nlevels ? p->d_A : nullptr;
    return p->d_band[level - 1][b];
}
}
extern "C" int pwt3_get_coeff(pwt3_plan* p, float* dst, int level, int b) {
    float* src = band_ptr(p, level, b);
    if (!src || !dst) return failv(PWT_ERR_ARG, "bad band (level %d, index %d)", level, b);
    if (p->state == PWT_INVERSE) return failv(PWT_ERR_STATE, "the coefficients were consumed by inverse()");
    cudaSetDevice(p->device);
    CKV(cudaMemcpyAsync(dst, src, (size_t)vox(p, level) * sizeof(float), cudaMemcpyDeviceToHost, p->stream));
    CKV(cudaStreamSynchronize(p->stream));
    return PWT_OK;
}
extern "C" int pwt3_set_coeff(pwt3_plan* p, const float* src, int level, int b, int on_device) {
    float* dst = band_ptr(p, level, b);
    if (!dst || !src) return failv(PWT_ERR_ARG, "bad band (level %d, index %d)", level, b);
    cudaSetDevice(p->device);
    CKV(cudaMemcpyAsync(dst, src, (size_t)vox(p, level) * sizeof(float), on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, p->stream));
    if (!on_device) CKV(cudaStreamSynchronize(p->stream));
    return PWT_OK;
}
extern "C" intptr_t pwt3_coeff_ptr(pwt3_plan* p, int level, int b) { return (intptr_t)band_ptr(p, level, b); }
extern "C" intptr_t pwt3_image_ptr(pwt3_plan* p) { return p ? (intptr_t)p->d_image : 0; }
extern "C" int pwt3_sync(pwt3_plan* p) {
    if (!p) return failv(PWT_ERR_ARG, "null plan");
    cudaSetDevice(p->device);
    CKV(cudaStreamSynchronize(p->stream));
    return PWT_OK;
}
extern "C" int pwt3_timer_start(pwt3_plan* p) {
    if (!p) return failv(PWT_ERR_ARG, "null plan");
    cudaSetDevice(p->device);
    CKV(cudaEventRecord(p->ev0, p->stream));
    return PWT_OK;
}
extern "C" int pwt3_timer_stop(pwt3_plan* p, float* ms) {
    if (!p || !ms) return failv(PWT_ERR_ARG, "null argument");
    cudaSetDevice(p->device);
    CKV(cudaEventRecord(p->ev1, p->stream));
    CKV(cudaEventSynchronize(p->ev1));
    CKV(cudaEventElapsedTime(ms, p->ev0, p->ev1));
    return PWT_OK;
}
extern "C" long long pwt3_launch_count(const pwt3_plan* p) { return p ? p->launches : 0; }
