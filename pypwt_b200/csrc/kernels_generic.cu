// Generic shared-memory tiled kernels: valid for every image size (odd sizes, tiny images,
// wrap-around tiles) and every filter length up to PWT_MAX_TAPS.  They are the reference point
// the specialised kernels in kernels_fast.cu are tested against, and the fallback for every
// configuration those do not cover.
//
// Arithmetic contract (what the reference's kernels compute; restated in
// oracle/pdwt_oracle.py):
//   analysis   out[k] = sum_j f[F-1-j] * xe[(2k - c + j) mod Ne], c=(F-1)/2   separable.cu:91-131
//   synthesis  x[n]   = sum_j fl[t]*a[k] + fh[t]*d[k],  b=n&1, p=F/2-1,
//                       t = 2j + ((b+p)&1), k = ((n>>1) + ((b+p)>>1) - j) mod n2   separable.cu:246-328
//   a trous    out[g] = sum_j f[F-1-j] * x[(g + (j-c)*s) mod N]               separable.cu:409-493
//   a trous^-1 x[g]   = sum_j (fl[F-1-j]/2)*a[(g+(j-F/2)*s) mod N] + (fh..)   separable.cu:553-626
// Rows are filtered first and columns second in the forward direction, columns first in the
// inverse direction (same order as the reference, so fp32 rounding follows the same path).
#include "pwt_internal.h"

namespace {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;

// ---- index helpers ---------------------------------------------------------------------
// DWT periodisation: odd N is first extended by one repeated sample, then made periodic.
__device__ __forceinline__ int wrap_dwt(int i, int N) {
    const int Ne = N + (N & 1);
    i %= Ne;
    if (i < 0) i += Ne;
    return i >= N ? N - 1 : i;
}
__device__ __forceinline__ int wrap_per(int i, int N) {
    i %= N;
    return i < 0 ? i + N : i;
}

// ---- precision helpers: the separable kernels below are templates over the sample type (float: the reference's
// build; double: its DOUBLEPRECISION build, filters.h:16-30 -- SURVEY 8f rank 4) -------------------------------
__device__ __forceinline__ float fma_t(float a, float b, float c) { return fmaf(a, b, c); }
__device__ __forceinline__ double fma_t(double a, double b, double c) { return fma(a, b, c); }
template <typename T> __device__ __forceinline__ T sqrt_half();
template <> __device__ __forceinline__ float sqrt_half<float>() { return 0.70710678118654746f; }
template <> __device__ __forceinline__ double sqrt_half<double>() { return 0.70710678118654752440; }

// =========================================================================================
// separable DWT, forward, fused row+column pass
// =========================================================================================
constexpr int GTX = 32;   // output columns per tile
constexpr int GTY = 16;   // output rows per tile

template <typename T, bool HAAR>
__global__ void __launch_bounds__(kThreads)
k_dwt_fwd2d(const T* __restrict__ in, T* __restrict__ A, T* __restrict__ Hb,
            T* __restrict__ V, T* __restrict__ D, int Nr, int Nc, long long in_bs,
            long long out_bs, const __grid_constant__ PwtFiltersT<T> f) {
    extern __shared__ __align__(16) unsigned char sm_raw[];
    T* sm = reinterpret_cast<T*>(sm_raw);
    const int F = f.hlen;
    const int c = (F - 1) / 2;
    const int IH = 2 * GTY + F - 2, IW = 2 * GTX + F - 2;
    const int IWp = IW | 1;                       // odd pitch: the stride-2 row pass stays conflict-light
    int* colidx = reinterpret_cast<int*>(sm);     // IW entries, padded to a multiple of 4
    T* s_in = sm + ((IW + 4) & ~3);
    T* s_lo = s_in + IH * IWp;
    T* s_hi = s_lo + IH * GTX;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int Nr2 = (Nr + 1) >> 1, Nc2 = (Nc + 1) >> 1;
    const int kx0 = blockIdx.x * GTX, ky0 = blockIdx.y * GTY;
    in += blockIdx.z * in_bs;
    const long long ob = blockIdx.z * out_bs;

    for (int i = tid; i < IW; i += kThreads) colidx[i] = wrap_dwt(2 * kx0 - c + i, Nc);
    __syncthreads();
    for (int r = warp; r < IH; r += kWarps) {
        const T* row = in + (long long)wrap_dwt(2 * ky0 - c + r, Nr) * Nc;
        for (int cc = lane; cc < IW; cc += 32) s_in[r * IWp + cc] = __ldg(row + colidx[cc]);
    }
    __syncthreads();
    // row pass (lane = output column)
    for (int r = warp; r < IH; r += kWarps) {
        const T* p = s_in + r * IWp + 2 * lane;
        T lo, hi;
        if (HAAR) {                                     // the butterfly associates down the columns first (haar.cu:27-35): keep
            lo = p[0];                                  // the two samples of the pair, the column pass does all the arithmetic
            hi = p[1];
        } else {
            lo = T(0), hi = T(0);
            for (int j = 0; j < F; j++) {
                const T v = p[j];
                lo = fma_t(v, f.L[F - 1 - j], lo);
                hi = fma_t(v, f.H[F - 1 - j], hi);
            }
        }
        s_lo[r * GTX + lane] = lo;
        s_hi[r * GTX + lane] = hi;
    }
    __syncthreads();
    // column pass
    for (int y = warp; y < GTY; y += kWarps) {
        T a, h, v, d;
        const T* pl = s_lo + (2 * y) * GTX + lane;
        const T* ph = s_hi + (2 * y) * GTX + lane;
        if (HAAR) {                                     // A = ((a + c) + (b + d)) / 2, V = ((a + c) - (b + d)) / 2,
            const T ac = pl[0] + pl[GTX], bd = ph[0] + ph[GTX];   // H = ((a - c) + (b - d)) / 2, D = ((a - c) - (b - d)) / 2
            const T am = pl[0] - pl[GTX], bm = ph[0] - ph[GTX];
            a = T(0.5) * (ac + bd);
            v = T(0.5) * (ac - bd);
            h = T(0.5) * (am + bm);
            d = T(0.5) * (am - bm);
        } else {
            a = h = v = d = T(0);
            for (int j = 0; j < F; j++) {
                const T l = pl[j * GTX], g = ph[j * GTX];
                const T tl = f.L[F - 1 - j], th = f.H[F - 1 - j];
                a = fma_t(l, tl, a);
                h = fma_t(l, th, h);
                v = fma_t(g, tl, v);
                d = fma_t(g, th, d);
            }
        }
        const int ky = ky0 + y, kx = kx0 + lane;
        if (ky < Nr2 && kx < Nc2) {
            const long long o = ob + (long long)ky * Nc2 + kx;
            A[o] = a;
            Hb[o] = h;
            V[o] = v;
            D[o] = d;
        }
    }
}

// =========================================================================================
// separable DWT, inverse, fused column+row pass (polyphase: F/2 taps per output)
// =========================================================================================
constexpr int GOX = 64;   // output columns per tile
constexpr int GOY = 32;   // output rows per tile

template <typename T, bool HAAR>
__global__ void __launch_bounds__(kThreads)
k_dwt_inv2d(const T* __restrict__ A, const T* __restrict__ Hb, const T* __restrict__ V,
            const T* __restrict__ D, T* __restrict__ out, int nr, int nc, int Nr_out,
            int Nc_out, long long in_bs, long long out_bs, const __grid_constant__ PwtFiltersT<T> f) {
    extern __shared__ __align__(16) unsigned char sm_raw[];
    T* sm = reinterpret_cast<T*>(sm_raw);
    const int F = f.hlen;
    const int p = F / 2 - 1, hl = (p + 1) >> 1, half = F / 2;
    const int BH = GOY / 2 + 2 * hl, BW = GOX / 2 + 2 * hl;
    const int BWp = BW | 1;
    int* colidx = reinterpret_cast<int*>(sm);
    T* s_b = sm + ((BW + 4) & ~3);            // 4 band tiles [4][BH][BWp]
    T* s_t = s_b + 4 * BH * BWp;              // t1, t2 : [2][GOY][BWp]

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n0x = blockIdx.x * GOX, n0y = blockIdx.y * GOY;
    const int kx_start = n0x / 2 - hl, ky_start = n0y / 2 - hl;
    const long long ib = blockIdx.z * in_bs;
    out += blockIdx.z * out_bs;
    const T* bands[4] = {A + ib, Hb + ib, V + ib, D + ib};

    for (int i = tid; i < BW; i += kThreads) colidx[i] = wrap_per(kx_start + i, nc);
    __syncthreads();
    for (int r = warp; r < 4 * BH; r += kWarps) {
        const int b = r / BH, rr = r - b * BH;
        const T* row = bands[b] + (long long)wrap_per(ky_start + rr, nr) * nc;
        T* dst = s_b + (b * BH + rr) * BWp;
        for (int cc = lane; cc < BW; cc += 32) dst[cc] = __ldg(row + colidx[cc]);
    }
    __syncthreads();
    // column synthesis: t1 = syn_y(A, H), t2 = syn_y(V, D)
    for (int i = tid; i < GOY * BW; i += kThreads) {
        const int y = i / BW, x = i - y * BW;
        const int b = y & 1;
        const int kb = (y >> 1) + ((b + p) >> 1) + hl;   // local row of tap j = 0
        const int t0 = (b + p) & 1;
        T t1, t2;
        const T* pa = s_b + (0 * BH + kb) * BWp + x;
        const T* ph = s_b + (1 * BH + kb) * BWp + x;
        const T* pv = s_b + (2 * BH + kb) * BWp + x;
        const T* pd = s_b + (3 * BH + kb) * BWp + x;
        if (HAAR) {
            t1 = b ? pa[0] - ph[0] : pa[0] + ph[0];
            t2 = b ? pv[0] - pd[0] : pv[0] + pd[0];
        } else {
            t1 = t2 = T(0);
            for (int j = 0; j < half; j++) {
                const T tl = f.IL[2 * j + t0], th = f.IH[2 * j + t0];
                const int o = -j * BWp;
                t1 = fma_t(pa[o], tl, t1);
                t1 = fma_t(ph[o], th, t1);
                t2 = fma_t(pv[o], tl, t2);
                t2 = fma_t(pd[o], th, t2);
            }
        }
        s_t[y * BWp + x] = t1;
        s_t[(GOY + y) * BWp + x] = t2;
    }
    __syncthreads();
    // row synthesis
    for (int i = tid; i < GOY * GOX; i += kThreads) {
        const int y = i / GOX, x = i - y * GOX;
        const int b = x & 1;
        const int kb = (x >> 1) + ((b + p) >> 1) + hl;
        const int t0 = (b + p) & 1;
        const T* p1 = s_t + y * BWp + kb;
        const T* p2 = s_t + (GOY + y) * BWp + kb;
        T r;
        if (HAAR) {
            r = T(0.5) * (b ? p1[0] - p2[0] : p1[0] + p2[0]);
        } else {
            r = T(0);
            for (int j = 0; j < half; j++) {
                r = fma_t(p1[-j], f.IL[2 * j + t0], r);
                r = fma_t(p2[-j], f.IH[2 * j + t0], r);
            }
        }
        const int gy = n0y + y, gx = n0x + x;
        if (gy < Nr_out && gx < Nc_out) out[(long long)gy * Nc_out + gx] = r;
    }
}

// =========================================================================================
// batched 1D DWT (row pass only), forward and inverse
// =========================================================================================
constexpr int G1X = 256;  // outputs per block (forward) / coefficient columns per block (inverse)

template <typename T, bool HAAR>
__global__ void __launch_bounds__(kThreads)
k_dwt_fwd1d(const T* __restrict__ in, T* __restrict__ A, T* __restrict__ D, int rows,
            int Nc, const __grid_constant__ PwtFiltersT<T> f) {
    extern __shared__ __align__(16) unsigned char sm_raw[];
    T* sm = reinterpret_cast<T*>(sm_raw);
    const int F = f.hlen, c = (F - 1) / 2;
    const int IW = 2 * G1X + F - 2;
    const int Nc2 = (Nc + 1) >> 1;
    const int kx0 = blockIdx.x * G1X, tid = threadIdx.x;
    for (int row = blockIdx.y; row < rows; row += gridDim.y) {
        const T* src = in + (long long)row * Nc;
        for (int i = tid; i < IW; i += kThreads) sm[i] = __ldg(src + wrap_dwt(2 * kx0 - c + i, Nc));
        __syncthreads();
        const T* p = sm + 2 * tid;
        T lo, hi;
        if (HAAR) {
            lo = sqrt_half<T>() * (p[0] + p[1]);
            hi = sqrt_half<T>() * (p[0] - p[1]);
        } else {
            lo = hi = T(0);
            for (int j = 0; j < F; j++) {
                lo = fma_t(p[j], f.L[F - 1 - j], lo);
                hi = fma_t(p[j], f.H[F - 1 - j], hi);
            }
        }
        if (kx0 + tid < Nc2) {
            A[(long long)row * Nc2 + kx0 + tid] = lo;
            D[(long long)row * Nc2 + kx0 + tid] = hi;
        }
        __syncthreads();
    }
}

template <typename T, bool HAAR>
__global__ void __launch_bounds__(kThreads)
k_dwt_inv1d(const T* __restrict__ A, const T* __restrict__ D, T* __restrict__ out,
            int rows, int nc, int Nc_out, const __grid_constant__ PwtFiltersT<T> f) {
    extern __shared__ __align__(16) unsigned char sm_raw[];
    T* sm = reinterpret_cast<T*>(sm_raw);
    const int F = f.hlen, p = F / 2 - 1, hl = (p + 1) >> 1, half = F / 2;
    const int BW = G1X + 2 * hl;
    T* s_a = sm;
    T* s_d = sm + BW;
    const int k0 = blockIdx.x * G1X - hl, tid = threadIdx.x;
    for (int row = blockIdx.y; row < rows; row += gridDim.y) {
        for (int i = tid; i < BW; i += kThreads) {
            const int k = wrap_per(k0 + i, nc);
            s_a[i] = __ldg(A + (long long)row * nc + k);
            s_d[i] = __ldg(D + (long long)row * nc + k);
        }
        __syncthreads();
#pragma unroll
        for (int b = 0; b < 2; b++) {
            const int kb = tid + ((b + p) >> 1) + hl, t0 = (b + p) & 1;
            T r;
            if (HAAR) {
                r = sqrt_half<T>() * (b ? s_a[kb] - s_d[kb] : s_a[kb] + s_d[kb]);
            } else {
                r = T(0);
                for (int j = 0; j < half; j++) {
                    r = fma_t(s_a[kb - j], f.IL[2 * j + t0], r);
                    r = fma_t(s_d[kb - j], f.IH[2 * j + t0], r);
                }
            }
            const int n = 2 * (blockIdx.x * G1X + tid) + b;
            if (n < Nc_out) out[(long long)row * Nc_out + n] = r;
        }
        __syncthreads();
    }
}

// =========================================================================================
// stationary (a trous) transform: unfused passes, one output per thread, lanes along x
// (coalesced for every tap; the F-fold re-reads are served by L1/L2).  A fused tile kernel
// for small dilations lives in kernels_fast.cu.
// =========================================================================================
// rows pass.  FWD: (in) -> (lo, hi) with analysis taps.  INV: (a, d) -> out with synthesis taps / 2.
template <typename T, bool INV>
__global__ void __launch_bounds__(kThreads)
k_swt_rows(const T* __restrict__ in0, const T* __restrict__ in1, T* __restrict__ out0,
           T* __restrict__ out1, long long rows, int Nc, int s,
           const __grid_constant__ PwtFiltersT<T> f) {
    const int F = f.hlen;
    const int c = INV ? F / 2 : (F - 1) / 2;
    const int x = blockIdx.x * kThreads + threadIdx.x;
    if (x >= Nc) return;
    for (long long row = blockIdx.y; row < rows; row += gridDim.y) {
        const T* p0 = in0 + row * Nc;
        if (!INV) {
            T lo = T(0), hi = T(0);
            for (int j = 0; j < F; j++) {
                const T v = __ldg(p0 + wrap_per(x + (j - c) * s, Nc));
                lo = fma_t(v, f.L[F - 1 - j], lo);
                hi = fma_t(v, f.H[F - 1 - j], hi);
            }
            out0[row * Nc + x] = lo;
            out1[row * Nc + x] = hi;
        } else {
            const T* p1 = in1 + row * Nc;
            T r1 = T(0), r2 = T(0);
            for (int j = 0; j < F; j++) {
                const int xx = wrap_per(x + (j - c) * s, Nc);
                r1 = fma_t(__ldg(p0 + xx), T(0.5) * f.IL[F - 1 - j], r1);
                r2 = fma_t(__ldg(p1 + xx), T(0.5) * f.IH[F - 1 - j], r2);
            }
            out0[row * Nc + x] = r1 + r2;
        }
    }
}

// columns pass.  FWD: (lo, hi) -> (A, H, V, D).  INV: (A, H, V, D) -> (t1, t2).
template <typename T, bool INV>
__global__ void __launch_bounds__(kThreads)
k_swt_cols(const T* __restrict__ i0, const T* __restrict__ i1, const T* __restrict__ i2,
           const T* __restrict__ i3, T* __restrict__ o0, T* __restrict__ o1,
           T* __restrict__ o2, T* __restrict__ o3, int Nr, int Nc, int s,
           const __grid_constant__ PwtFiltersT<T> f) {
    const int F = f.hlen;
    const int c = INV ? F / 2 : (F - 1) / 2;
    const int x = blockIdx.x * kThreads + threadIdx.x;
    const long long pb = (long long)blockIdx.z * Nr * Nc;
    if (x >= Nc) return;
    for (int y = blockIdx.y; y < Nr; y += gridDim.y) {
        const long long o = pb + (long long)y * Nc + x;
        if (!INV) {
            T a = T(0), h = T(0), v = T(0), d = T(0);
            for (int j = 0; j < F; j++) {
                const long long q = pb + (long long)wrap_per(y + (j - c) * s, Nr) * Nc + x;
                const T l = __ldg(i0 + q), g = __ldg(i1 + q);
                const T tl = f.L[F - 1 - j], th = f.H[F - 1 - j];
                a = fma_t(l, tl, a);
                h = fma_t(l, th, h);
                v = fma_t(g, tl, v);
                d = fma_t(g, th, d);
            }
            o0[o] = a;
            o1[o] = h;
            o2[o] = v;
            o3[o] = d;
        } else {
            T ra = T(0), rh = T(0), rv = T(0), rd = T(0);
            for (int j = 0; j < F; j++) {
                const long long q = pb + (long long)wrap_per(y + (j - c) * s, Nr) * Nc + x;
                const T tl = T(0.5) * f.IL[F - 1 - j], th = T(0.5) * f.IH[F - 1 - j];
                ra = fma_t(__ldg(i0 + q), tl, ra);
                rh = fma_t(__ldg(i1 + q), th, rh);
                rv = fma_t(__ldg(i2 + q), tl, rv);
                rd = fma_t(__ldg(i3 + q), th, rd);
            }
            o0[o] = ra + rh;
            o1[o] = rv + rd;
        }
    }
}

// =========================================================================================
// non-separable transforms: true 2D stencils with four F x F filters (LL, LH, HL, HH)
// k2d layout: [4][F][F], K[b][i][j] with i the y-tap and j the x-tap (nonseparable.cu:16-25,70-74)
// =========================================================================================
__global__ void __launch_bounds__(kThreads)
k_ns_fwd2d(const float* __restrict__ in, float* __restrict__ A, float* __restrict__ Hb,
           float* __restrict__ V, float* __restrict__ D, int Nr, int Nc, long long in_bs,
           long long out_bs, const float* __restrict__ k2d, int F) {
    extern __shared__ float sm[];
    const int c = (F - 1) / 2;
    const int IH = 2 * GTY + F - 2, IW = 2 * GTX + F - 2;
    const int IWp = IW | 1;
    int* colidx = reinterpret_cast<int*>(sm);
    float* s_in = sm + ((IW + 4) & ~3);
    float4* s_k = reinterpret_cast<float4*>(s_in + ((IH * IWp + 3) & ~3));   // [F*F] interleaved taps

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int Nr2 = (Nr + 1) >> 1, Nc2 = (Nc + 1) >> 1;
    const int kx0 = blockIdx.x * GTX, ky0 = blockIdx.y * GTY;
    in += blockIdx.z * in_bs;
    const long long ob = blockIdx.z * out_bs;
    const int FF = F * F;
    for (int i = tid; i < FF; i += kThreads)
        s_k[i] = make_float4(k2d[i], k2d[FF + i], k2d[2 * FF + i], k2d[3 * FF + i]);
    for (int i = tid; i < IW; i += kThreads) colidx[i] = wrap_dwt(2 * kx0 - c + i, Nc);
    __syncthreads();
    for (int r = warp; r < IH; r += kWarps) {
        const float* row = in + (long long)wrap_dwt(2 * ky0 - c + r, Nr) * Nc;
        for (int cc = lane; cc < IW; cc += 32) s_in[r * IWp + cc] = __ldg(row + colidx[cc]);
    }
    __syncthreads();
    for (int y = warp; y < GTY; y += kWarps) {
        float a = 0.f, h = 0.f, v = 0.f, d = 0.f;
        for (int jy = 0; jy < F; jy++) {
            const float* p = s_in + (2 * y + jy) * IWp + 2 * lane;
            const float4* kk = s_k + (F - 1 - jy) * F + (F - 1);
            for (int jx = 0; jx < F; jx++) {
                const float val = p[jx];
                const float4 t = kk[-jx];
                a = fmaf(val, t.x, a);
                h = fmaf(val, t.y, h);
                v = fmaf(val, t.z, v);
                d = fmaf(val, t.w, d);
            }
        }
        const int ky = ky0 + y, kx = kx0 + lane;
        if (ky < Nr2 && kx < Nc2) {
            const long long o = ob + (long long)ky * Nc2 + kx;
            A[o] = a;
            Hb[o] = h;
            V[o] = v;
            D[o] = d;
        }
    }
}

__global__ void __launch_bounds__(kThreads)
k_ns_inv2d(const float* __restrict__ A, const float* __restrict__ Hb, const float* __restrict__ V,
           const float* __restrict__ D, float* __restrict__ out, int nr, int nc, int Nr_out,
           int Nc_out, long long in_bs, long long out_bs, const float* __restrict__ k2d, int F) {
    extern __shared__ float sm[];
    const int p = F / 2 - 1, hl = (p + 1) >> 1, half = F / 2;
    const int BH = GOY / 2 + 2 * hl, BW = GOX / 2 + 2 * hl;
    const int BWp = BW | 1;
    int* colidx = reinterpret_cast<int*>(sm);
    float* s_b = sm + ((BW + 4) & ~3);
    float4* s_k = reinterpret_cast<float4*>(s_b + ((4 * BH * BWp + 3) & ~3));

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n0x = blockIdx.x * GOX, n0y = blockIdx.y * GOY;
    const int kx_start = n0x / 2 - hl, ky_start = n0y / 2 - hl;
    const long long ib = blockIdx.z * in_bs;
    out += blockIdx.z * out_bs;
    const float* bands[4] = {A + ib, Hb + ib, V + ib, D + ib};
    const int FF = F * F;
    for (int i = tid; i < FF; i += kThreads)
        s_k[i] = make_float4(k2d[i], k2d[FF + i], k2d[2 * FF + i], k2d[3 * FF + i]);
    for (int i = tid; i < BW; i += kThreads) colidx[i] = wrap_per(kx_start + i, nc);
    __syncthreads();
    for (int r = warp; r < 4 * BH; r += kWarps) {
        const int b = r / BH, rr = r - b * BH;
        const float* row = bands[b] + (long long)wrap_per(ky_start + rr, nr) * nc;
        float* dst = s_b + (b * BH + rr) * BWp;
        for (int cc = lane; cc < BW; cc += 32) dst[cc] = __ldg(row + colidx[cc]);
    }
    __syncthreads();
    for (int i = tid; i < GOY * GOX; i += kThreads) {
        const int y = i / GOX, x = i - y * GOX;
        const int by = y & 1, bx = x & 1;
        const int kyb = (y >> 1) + ((by + p) >> 1) + hl, ty0 = (by + p) & 1;
        const int kxb = (x >> 1) + ((bx + p) >> 1) + hl, tx0 = (bx + p) & 1;
        float r = 0.f;
        for (int jy = 0; jy < half; jy++) {
            const int off = (kyb - jy) * BWp + kxb;
            const float4* kk = s_k + (2 * jy + ty0) * F + tx0;
            for (int jx = 0; jx < half; jx++) {
                const float4 t = kk[2 * jx];
                const int o = off - jx;
                r = fmaf(s_b[o], t.x, r);
                r = fmaf(s_b[BH * BWp + o], t.y, r);
                r = fmaf(s_b[2 * BH * BWp + o], t.z, r);
                r = fmaf(s_b[3 * BH * BWp + o], t.w, r);
            }
        }
        const int gy = n0y + y, gx = n0x + x;
        if (gy < Nr_out && gx < Nc_out) out[(long long)gy * Nc_out + gx] = r;
    }
}

// a trous 2D stencils (direct global reads, coalesced along x)
template <bool INV>
__global__ void __launch_bounds__(kThreads)
k_ns_swt2d(const float* __restrict__ i0, const float* __restrict__ i1, const float* __restrict__ i2,
           const float* __restrict__ i3, float* __restrict__ o0, float* __restrict__ o1,
           float* __restrict__ o2, float* __restrict__ o3, int Nr, int Nc, int s,
           const float* __restrict__ k2d, int F) {
    extern __shared__ float sm[];
    float4* s_k = reinterpret_cast<float4*>(sm);
    const int FF = F * F;
    for (int i = threadIdx.x; i < FF; i += kThreads)
        s_k[i] = make_float4(k2d[i], k2d[FF + i], k2d[2 * FF + i], k2d[3 * FF + i]);
    __syncthreads();
    const int c = INV ? F / 2 : (F - 1) / 2;
    const int x = blockIdx.x * kThreads + threadIdx.x;
    const long long pb = (long long)blockIdx.z * Nr * Nc;
    if (x >= Nc) return;
    for (int y = blockIdx.y; y < Nr; y += gridDim.y) {
        float a = 0.f, h = 0.f, v = 0.f, d = 0.f;
        for (int jy = 0; jy < F; jy++) {
            const long long rb = pb + (long long)wrap_per(y + (jy - c) * s, Nr) * Nc;
            const float4* kk = s_k + (F - 1 - jy) * F + (F - 1);
            for (int jx = 0; jx < F; jx++) {
                const long long q = rb + wrap_per(x + (jx - c) * s, Nc);
                const float4 t = kk[-jx];
                if (!INV) {
                    const float val = __ldg(i0 + q);
                    a = fmaf(val, t.x, a);
                    h = fmaf(val, t.y, h);
                    v = fmaf(val, t.z, v);
                    d = fmaf(val, t.w, d);
                } else {
                    a = fmaf(__ldg(i0 + q), 0.25f * t.x, a);
                    h = fmaf(__ldg(i1 + q), 0.25f * t.y, h);
                    v = fmaf(__ldg(i2 + q), 0.25f * t.z, v);
                    d = fmaf(__ldg(i3 + q), 0.25f * t.w, d);
                }
            }
        }
        const long long o = pb + (long long)y * Nc + x;
        if (!INV) {
            o0[o] = a;
            o1[o] = h;
            o2[o] = v;
            o3[o] = d;
        } else {
            o0[o] = a + h + v + d;
        }
    }
}

inline int cdiv(int a, int b) { return (a + b - 1) / b; }
inline int clamp_grid_y(long long n) { return (int)(n < 65535 ? (n < 1 ? 1 : n) : 65535); }

template <typename K>
inline void set_smem(K kernel, size_t bytes) {
    if (bytes > 48 * 1024)
        cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
}

}  // namespace

namespace {
template <typename T>
static int launch_dwt_fwd2d_t(const T* in, T* A, T* Hb, T* V, T* D, int batch, int Nr,
                         int Nc, long long in_bs, long long out_bs, const PwtFiltersT<T>& f, bool haar,
                         cudaStream_t st) {
    const int F = haar ? 2 : f.hlen;
    PwtFiltersT<T> ff = f;
    ff.hlen = F;
    const int IH = 2 * GTY + F - 2, IW = 2 * GTX + F - 2, IWp = IW | 1;
    const size_t smem = sizeof(T) * (size_t)(((IW + 4) & ~3) + IH * IWp + 2 * IH * GTX);
    dim3 grid(cdiv((Nc + 1) / 2, GTX), cdiv((Nr + 1) / 2, GTY), batch);
    if (haar) {
        set_smem(k_dwt_fwd2d<T, true>, smem);
        k_dwt_fwd2d<T, true><<<grid, kThreads, smem, st>>>(in, A, Hb, V, D, Nr, Nc, in_bs, out_bs, ff);
    } else {
        set_smem(k_dwt_fwd2d<T, false>, smem);
        k_dwt_fwd2d<T, false><<<grid, kThreads, smem, st>>>(in, A, Hb, V, D, Nr, Nc, in_bs, out_bs, ff);
    }
    return 1;
}

template <typename T>
static int launch_dwt_inv2d_t(const T* A, const T* Hb, const T* V, const T* D, T* out,
                         int batch, int nr, int nc, int Nr_out, int Nc_out, long long in_bs,
                         long long out_bs, const PwtFiltersT<T>& f, bool haar, cudaStream_t st) {
    const int F = haar ? 2 : f.hlen;
    PwtFiltersT<T> ff = f;
    ff.hlen = F;
    const int p = F / 2 - 1, hl = (p + 1) >> 1;
    const int BH = GOY / 2 + 2 * hl, BW = GOX / 2 + 2 * hl, BWp = BW | 1;
    const size_t smem = sizeof(T) * (size_t)(((BW + 4) & ~3) + 4 * BH * BWp + 2 * GOY * BWp);
    dim3 grid(cdiv(Nc_out, GOX), cdiv(Nr_out, GOY), batch);
    if (haar) {
        set_smem(k_dwt_inv2d<T, true>, smem);
        k_dwt_inv2d<T, true><<<grid, kThreads, smem, st>>>(A, Hb, V, D, out, nr, nc, Nr_out, Nc_out,
                                                        in_bs, out_bs, ff);
    } else {
        set_smem(k_dwt_inv2d<T, false>, smem);
        k_dwt_inv2d<T, false><<<grid, kThreads, smem, st>>>(A, Hb, V, D, out, nr, nc, Nr_out, Nc_out,
                                                         in_bs, out_bs, ff);
    }
    return 1;
}

template <typename T>
static int launch_dwt_fwd1d_t(const T* in, T* A, T* D, int rows, int Nc, const PwtFiltersT<T>& f,
                         bool haar, cudaStream_t st) {
    const int F = haar ? 2 : f.hlen;
    PwtFiltersT<T> ff = f;
    ff.hlen = F;
    const size_t smem = sizeof(T) * (size_t)(2 * G1X + F - 2);
    dim3 grid(cdiv((Nc + 1) / 2, G1X), clamp_grid_y(rows), 1);
    if (haar)
        k_dwt_fwd1d<T, true><<<grid, kThreads, smem, st>>>(in, A, D, rows, Nc, ff);
    else
        k_dwt_fwd1d<T, false><<<grid, kThreads, smem, st>>>(in, A, D, rows, Nc, ff);
    return 1;
}

template <typename T>
static int launch_dwt_inv1d_t(const T* A, const T* D, T* out, int rows, int nc, int Nc_out,
                         const PwtFiltersT<T>& f, bool haar, cudaStream_t st) {
    const int F = haar ? 2 : f.hlen;
    PwtFiltersT<T> ff = f;
    ff.hlen = F;
    const int hl = (F / 2) >> 1;
    const size_t smem = sizeof(T) * (size_t)(2 * (G1X + 2 * hl));
    dim3 grid(cdiv(cdiv(Nc_out, 2), G1X), clamp_grid_y(rows), 1);
    if (haar)
        k_dwt_inv1d<T, true><<<grid, kThreads, smem, st>>>(A, D, out, rows, nc, Nc_out, ff);
    else
        k_dwt_inv1d<T, false><<<grid, kThreads, smem, st>>>(A, D, out, rows, nc, Nc_out, ff);
    return 1;
}

template <typename T>
static int launch_swt_fwd1d_t(const T* in, T* A, T* D, int rows, int Nc, int level,
                         const PwtFiltersT<T>& f, cudaStream_t st) {
    dim3 grid(cdiv(Nc, kThreads), clamp_grid_y(rows), 1);
    k_swt_rows<T, false><<<grid, kThreads, 0, st>>>(in, nullptr, A, D, rows, Nc, 1 << (level - 1), f);
    return 1;
}

template <typename T>
static int launch_swt_inv1d_t(const T* A, const T* D, T* out, int rows, int Nc, int level,
                         const PwtFiltersT<T>& f, cudaStream_t st) {
    dim3 grid(cdiv(Nc, kThreads), clamp_grid_y(rows), 1);
    k_swt_rows<T, true><<<grid, kThreads, 0, st>>>(A, D, out, nullptr, rows, Nc, 1 << (level - 1), f);
    return 1;
}

template <typename T>
static int launch_swt_fwd2d_t(const T* in, T* A, T* Hb, T* V, T* D, T* tmp,
                         int batch, int Nr, int Nc, int level, const PwtFiltersT<T>& f, cudaStream_t st) {
    const long long plane = (long long)batch * Nr * Nc;
    T* lo = tmp;
    T* hi = tmp + plane;
    const int s = 1 << (level - 1);
    dim3 g1(cdiv(Nc, kThreads), clamp_grid_y((long long)batch * Nr), 1);
    k_swt_rows<T, false><<<g1, kThreads, 0, st>>>(in, nullptr, lo, hi, (long long)batch * Nr, Nc, s, f);
    dim3 g2(cdiv(Nc, kThreads), clamp_grid_y(Nr), batch);
    k_swt_cols<T, false><<<g2, kThreads, 0, st>>>(lo, hi, nullptr, nullptr, A, Hb, V, D, Nr, Nc, s, f);
    return 2;
}

template <typename T>
static int launch_swt_inv2d_t(const T* A, const T* Hb, const T* V, const T* D, T* out,
                         T* tmp, int batch, int Nr, int Nc, int level, const PwtFiltersT<T>& f,
                         cudaStream_t st) {
    const long long plane = (long long)batch * Nr * Nc;
    T* t1 = tmp;
    T* t2 = tmp + plane;
    const int s = 1 << (level - 1);
    dim3 g2(cdiv(Nc, kThreads), clamp_grid_y(Nr), batch);
    k_swt_cols<T, true><<<g2, kThreads, 0, st>>>(A, Hb, V, D, t1, t2, nullptr, nullptr, Nr, Nc, s, f);
    dim3 g1(cdiv(Nc, kThreads), clamp_grid_y((long long)batch * Nr), 1);
    k_swt_rows<T, true><<<g1, kThreads, 0, st>>>(t1, t2, out, nullptr, (long long)batch * Nr, Nc, s, f);
    return 2;
}

}  // namespace

// =========================================================================================
// launchers (float: the product path's fallback family; double: the fp64 plans of pwt_plan64.cu)
// =========================================================================================
int pwt_launch_dwt_fwd2d(const float* in, float* A, float* Hb, float* V, float* D, int batch, int Nr, int Nc, long long in_bs, long long out_bs, const PwtFilters& f, bool haar, cudaStream_t st) {
    return launch_dwt_fwd2d_t<float>(in, A, Hb, V, D, batch, Nr, Nc, in_bs, out_bs, f, haar, st);
}
int pwt_launch_dwt_fwd2d_f64(const double* in, double* A, double* Hb, double* V, double* D, int batch, int Nr, int Nc, long long in_bs, long long out_bs, const PwtFilters64& f, bool haar, cudaStream_t st) {
    return launch_dwt_fwd2d_t<double>(in, A, Hb, V, D, batch, Nr, Nc, in_bs, out_bs, f, haar, st);
}
int pwt_launch_dwt_inv2d(const float* A, const float* Hb, const float* V, const float* D, float* out, int batch, int nr, int nc, int Nr_out, int Nc_out, long long in_bs, long long out_bs, const PwtFilters& f, bool haar, cudaStream_t st) {
    return launch_dwt_inv2d_t<float>(A, Hb, V, D, out, batch, nr, nc, Nr_out, Nc_out, in_bs, out_bs, f, haar, st);
}
int pwt_launch_dwt_inv2d_f64(const double* A, const double* Hb, const double* V, const double* D, double* out, int batch, int nr, int nc, int Nr_out, int Nc_out, long long in_bs, long long out_bs, const PwtFilters64& f, bool haar, cudaStream_t st) {
    return launch_dwt_inv2d_t<double>(A, Hb, V, D, out, batch, nr, nc, Nr_out, Nc_out, in_bs, out_bs, f, haar, st);
}
int pwt_launch_dwt_fwd1d(const float* in, float* A, float* D, int rows, int Nc, const PwtFilters& f, bool haar, cudaStream_t st) {
    return launch_dwt_fwd1d_t<float>(in, A, D, rows, Nc, f, haar, st);
}
int pwt_launch_dwt_fwd1d_f64(const double* in, double* A, double* D, int rows, int Nc, const PwtFilters64& f, bool haar, cudaStream_t st) {
    return launch_dwt_fwd1d_t<double>(in, A, D, rows, Nc, f, haar, st);
}
int pwt_launch_dwt_inv1d(const float* A, const float* D, float* out, int rows, int nc, int Nc_out, const PwtFilters& f, bool haar, cudaStream_t st) {
    return launch_dwt_inv1d_t<float>(A, D, out, rows, nc, Nc_out, f, haar, st);
}
int pwt_launch_dwt_inv1d_f64(const double* A, const double* D, double* out, int rows, int nc, int Nc_out, const PwtFilters64& f, bool haar, cudaStream_t st) {
    return launch_dwt_inv1d_t<double>(A, D, out, rows, nc, Nc_out, f, haar, st);
}
int pwt_launch_swt_fwd1d(const float* in, float* A, float* D, int rows, int Nc, int level, const PwtFilters& f, cudaStream_t st) {
    return launch_swt_fwd1d_t<float>(in, A, D, rows, Nc, level, f, st);
}
int pwt_launch_swt_fwd1d_f64(const double* in, double* A, double* D, int rows, int Nc, int level, const PwtFilters64& f, cudaStream_t st) {
    return launch_swt_fwd1d_t<double>(in, A, D, rows, Nc, level, f, st);
}
int pwt_launch_swt_inv1d(const float* A, const float* D, float* out, int rows, int Nc, int level, const PwtFilters& f, cudaStream_t st) {
    return launch_swt_inv1d_t<float>(A, D, out, rows, Nc, level, f, st);
}
int pwt_launch_swt_inv1d_f64(const double* A, const double* D, double* out, int rows, int Nc, int level, const PwtFilters64& f, cudaStream_t st) {
    return launch_swt_inv1d_t<double>(A, D, out, rows, Nc, level, f, st);
}
int pwt_launch_swt_fwd2d(const float* in, float* A, float* Hb, float* V, float* D, float* tmp, int batch, int Nr, int Nc, int level, const PwtFilters& f, cudaStream_t st) {
    return launch_swt_fwd2d_t<float>(in, A, Hb, V, D, tmp, batch, Nr, Nc, level, f, st);
}
int pwt_launch_swt_fwd2d_f64(const double* in, double* A, double* Hb, double* V, double* D, double* tmp, int batch, int Nr, int Nc, int level, const PwtFilters64& f, cudaStream_t st) {
    return launch_swt_fwd2d_t<double>(in, A, Hb, V, D, tmp, batch, Nr, Nc, level, f, st);
}
int pwt_launch_swt_inv2d(const float* A, const float* Hb, const float* V, const float* D, float* out, float* tmp, int batch, int Nr, int Nc, int level, const PwtFilters& f, cudaStream_t st) {
    return launch_swt_inv2d_t<float>(A, Hb, V, D, out, tmp, batch, Nr, Nc, level, f, st);
}
int pwt_launch_swt_inv2d_f64(const double* A, const double* Hb, const double* V, const double* D, double* out, double* tmp, int batch, int Nr, int Nc, int level, const PwtFilters64& f, cudaStream_t st) {
    return launch_swt_inv2d_t<double>(A, Hb, V, D, out, tmp, batch, Nr, Nc, level, f, st);
}

int pwt_launch_ns_fwd2d(const float* in, float* A, float* Hb, float* V, float* D, int batch, int Nr,
                        int Nc, long long in_bs, long long out_bs, const float* k2d, int F,
                        cudaStream_t st) {
    const int IH = 2 * GTY + F - 2, IW = 2 * GTX + F - 2, IWp = IW | 1;
    const size_t smem = sizeof(float) * (size_t)(((IW + 4) & ~3) + ((IH * IWp + 3) & ~3)) +
                        sizeof(float4) * (size_t)F * F;
    dim3 grid(cdiv((Nc + 1) / 2, GTX), cdiv((Nr + 1) / 2, GTY), batch);
    set_smem(k_ns_fwd2d, smem);
    k_ns_fwd2d<<<grid, kThreads, smem, st>>>(in, A, Hb, V, D, Nr, Nc, in_bs, out_bs, k2d, F);
    return 1;
}

int pwt_launch_ns_inv2d(const float* A, const float* Hb, const float* V, const float* D, float* out,
                        int batch, int nr, int nc, int Nr_out, int Nc_out, long long in_bs,
                        long long out_bs, const float* k2d, int F, cudaStream_t st) {
    const int p = F / 2 - 1, hl = (p + 1) >> 1;
    const int BH = GOY / 2 + 2 * hl, BW = GOX / 2 + 2 * hl, BWp = BW | 1;
    const size_t smem = sizeof(float) * (size_t)(((BW + 4) & ~3) + ((4 * BH * BWp + 3) & ~3)) +
                        sizeof(float4) * (size_t)F * F;
    dim3 grid(cdiv(Nc_out, GOX), cdiv(Nr_out, GOY), batch);
    set_smem(k_ns_inv2d, smem);
    k_ns_inv2d<<<grid, kThreads, smem, st>>>(A, Hb, V, D, out, nr, nc, Nr_out, Nc_out, in_bs, out_bs,
                                             k2d, F);
    return 1;
}

int pwt_launch_ns_swt_fwd2d(const float* in, float* A, float* Hb, float* V, float* D, int batch,
                            int Nr, int Nc, int level, const float* k2d, int F, cudaStream_t st) {
    dim3 grid(cdiv(Nc, kThreads), clamp_grid_y(Nr), batch);
    const size_t smem = sizeof(float4) * (size_t)F * F;
    k_ns_swt2d<false><<<grid, kThreads, smem, st>>>(in, nullptr, nullptr, nullptr, A, Hb, V, D, Nr, Nc,
                                                    1 << (level - 1), k2d, F);
    return 1;
}

int pwt_launch_ns_swt_inv2d(const float* A, const float* Hb, const float* V, const float* D,
                            float* out, int batch, int Nr, int Nc, int level, const float* k2d,
                            int F, cudaStream_t st) {
    dim3 grid(cdiv(Nc, kThreads), clamp_grid_y(Nr), batch);
    const size_t smem = sizeof(float4) * (size_t)F * F;
    k_ns_swt2d<true><<<grid, kThreads, smem, st>>>(A, Hb, V, D, out, nullptr, nullptr, nullptr, Nr, Nc,
                                                   1 << (level - 1), k2d, F);
    return 1;
}
