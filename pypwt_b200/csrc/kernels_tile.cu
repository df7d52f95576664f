// Long filters (F = 10 .. 40): shared-memory tile kernels with compile-time filter length and
// register-blocked accumulators.  At these lengths the transform is FMA-bound (2F FMA per pixel and
// direction: db20 = 213 FMA/px for a 3-level fwd+inv), so the design goal is FMA issue efficiency:
// every 128-bit shared-memory read feeds 16-32 FMAs, taps are constant-bank operands (full unroll),
// no sliding-window register shuffling.  Every multiply-add is issued 2-wide (FFMA2, sm_100): one sample times
// a packed pair of taps -- (low-pass, high-pass) in the analysis, (even-phase, odd-phase) in the synthesis.
// FFMA2 has half the issue rate of FFMA (same FMA-pipe throughput, tools/bench/fma2bench.cu), so this
// halves the issue slots the arithmetic needs and leaves them to the shared-memory reads.
//   forward : stage the haloed input tile (wrap / odd sizes resolved while staging), row pass IN PLACE
//             (one warp per tile row: each lane reads its window, then the row is overwritten with its
//             lo | hi halves), column pass streaming over the F+2 rows of an output-row pair.
//   inverse : stage the four haloed band tiles, column synthesis into t1/t2, row synthesis, 128-bit stores.
// Same arithmetic order as the reference (rows then columns forward, columns then rows inverse).
#include <stdlib.h>

#include "pwt_internal.h"

namespace {

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
__device__ __forceinline__ void fma4(float4& acc, const float4& v, float t) {
    acc.x = fmaf(v.x, t, acc.x);
    acc.y = fmaf(v.y, t, acc.y);
    acc.z = fmaf(v.z, t, acc.z);
    acc.w = fmaf(v.w, t, acc.w);
}

// asynchronous global -> shared copies (LDGSTS): the whole haloed tile is requested back to back, so the
// staging phase costs about one memory latency instead of one per row
__device__ __forceinline__ void cp_async16(float* dst, const float* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async4(float* dst, const float* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((unsigned)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

__device__ __forceinline__ float2 fma2s(float x, float2 t, float2 acc) { return __ffma2_rn(make_float2(x, x), t, acc); }

using TapsFwd = PwtTapsFwd;
using TapsInv = PwtTapsInv;

constexpr int NT = 256;
constexpr int TX = 64;        // forward: output columns per tile (one warp row = 32 lanes x 2 outputs)
constexpr int TY = 32;        // forward: output rows per tile (16 row pairs x 16 column groups = 256 items)

template <int F>
struct FwdGeo {
    static constexpr int C = F / 2 - 1;
    static constexpr int CL = (C + 3) & ~3;                // the tile starts CL (aligned) columns left of 2*kx0
    static constexpr int DX = CL - C;                      // offset of the first window inside the tile
    static constexpr int IH = 2 * TY + F - 2;
    static constexpr int IW = 2 * TX + F - 2 + DX;
    static constexpr int P = ((IW + 3) & ~3) + 4;          // row pitch (floats), 16-byte aligned rows
    static constexpr size_t smem = sizeof(float) * ((size_t)IH * P) + sizeof(int) * (size_t)((IW + 3) & ~3);
};

template <int F>
__global__ void __launch_bounds__(NT)
k_tile_fwd(const float* __restrict__ in, float* __restrict__ A, float* __restrict__ Hb, float* __restrict__ V,
           float* __restrict__ D, int Nr, int Nc, long long in_bs, long long out_bs,
           const __grid_constant__ TapsFwd f) {
    using G = FwdGeo<F>;
    constexpr int C = G::C, CL = G::CL, DX = G::DX, IH = G::IH, IW = G::IW, P = G::P;
    constexpr int IW4 = (IW + 3) / 4;
    extern __shared__ __align__(16) float sm[];
    float* s = sm;
    int* colidx = reinterpret_cast<int*>(sm + IH * P);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int Nr2 = (Nr + 1) >> 1, Nc2 = (Nc + 1) >> 1;
    const int kx0 = blockIdx.x * TX, ky0 = blockIdx.y * TY;
    in += blockIdx.z * in_bs;
    const long long ob = blockIdx.z * out_bs;
    const int xs = 2 * kx0 - CL;                                // first tile column (multiple of 4)
    // interior tiles (all columns inside the image, 16-byte aligned rows) are staged with 128-bit loads
    const bool vec = (Nc & 3) == 0 && xs >= 0 && xs + 4 * IW4 <= Nc && (((uintptr_t)in) & 15) == 0 && (in_bs & 3) == 0;
    if (!vec) {
        for (int i = tid; i < IW; i += NT) colidx[i] = wrap_dwt(xs + i, Nc);
        __syncthreads();
    }
    for (int r = warp; r < IH; r += NT / 32) {
        const float* row = in + (long long)wrap_dwt(2 * ky0 - C + r, Nr) * Nc;
        if (vec) {
            for (int c4 = lane; c4 < IW4; c4 += 32) cp_async16(s + r * P + 4 * c4, row + xs + 4 * c4);
        } else {
            for (int cc = lane; cc < IW; cc += 32) cp_async4(s + r * P + cc, row + colidx[cc]);
        }
    }
    cp_async_wait_all();
    __syncthreads();
    // ---- row pass, in place: lane l owns outputs 2l, 2l+1 of the row; window = s[r][4l .. 4l+F+1] ----
    constexpr int NV = (DX + F + 2 + 3) / 4;
    for (int r = warp; r < IH; r += NT / 32) {
        float win[4 * NV];
#pragma unroll
        for (int k = 0; k < NV; k++) {
            const float4 v = *reinterpret_cast<const float4*>(s + r * P + 4 * lane + 4 * k);
            win[4 * k] = v.x; win[4 * k + 1] = v.y; win[4 * k + 2] = v.z; win[4 * k + 3] = v.w;
        }
        float2 p0 = make_float2(0.f, 0.f), p1 = p0;            // (lo, hi) of outputs 2l and 2l+1
#pragma unroll
        for (int j = 0; j < F; j++) {
            p0 = fma2s(win[DX + j], f.t[j], p0);
            p1 = fma2s(win[DX + j + 2], f.t[j], p1);
        }
        __syncwarp();                                           // everybody has read the row
        *reinterpret_cast<float2*>(s + r * P + 2 * lane) = make_float2(p0.x, p1.x);
        *reinterpret_cast<float2*>(s + r * P + TX + 2 * lane) = make_float2(p0.y, p1.y);
    }
    __syncthreads();
    // ---- column pass: item = (4 columns, 2 output rows); streams over the F+2 rows it needs ----
    {
        const int g = tid & 15, yp = tid >> 4;                  // column group, row pair
        // accumulators: (a, h) pairs from the low-pass rows, (v, d) pairs from the high-pass rows, 4 columns, 2 rows
        float2 ah0[4], vd0[4], ah1[4], vd1[4];
#pragma unroll
        for (int c = 0; c < 4; c++) ah0[c] = vd0[c] = ah1[c] = vd1[c] = make_float2(0.f, 0.f);
        const float* pl = s + (4 * yp) * P + 4 * g;
#pragma unroll
        for (int m = 0; m < F + 2; m++) {
            const float4 l = *reinterpret_cast<const float4*>(pl + m * P);
            const float4 h = *reinterpret_cast<const float4*>(pl + m * P + TX);
            const float lc[4] = {l.x, l.y, l.z, l.w}, hc[4] = {h.x, h.y, h.z, h.w};
            if (m < F) {
#pragma unroll
                for (int c = 0; c < 4; c++) {
                    ah0[c] = fma2s(lc[c], f.t[m], ah0[c]);
                    vd0[c] = fma2s(hc[c], f.t[m], vd0[c]);
                }
            }
            if (m >= 2) {
#pragma unroll
                for (int c = 0; c < 4; c++) {
                    ah1[c] = fma2s(lc[c], f.t[m - 2], ah1[c]);
                    vd1[c] = fma2s(hc[c], f.t[m - 2], vd1[c]);
                }
            }
        }
        const float4 a0 = make_float4(ah0[0].x, ah0[1].x, ah0[2].x, ah0[3].x), h0 = make_float4(ah0[0].y, ah0[1].y, ah0[2].y, ah0[3].y);
        const float4 v0 = make_float4(vd0[0].x, vd0[1].x, vd0[2].x, vd0[3].x), d0 = make_float4(vd0[0].y, vd0[1].y, vd0[2].y, vd0[3].y);
        const float4 a1 = make_float4(ah1[0].x, ah1[1].x, ah1[2].x, ah1[3].x), h1 = make_float4(ah1[0].y, ah1[1].y, ah1[2].y, ah1[3].y);
        const float4 v1 = make_float4(vd1[0].x, vd1[1].x, vd1[2].x, vd1[3].x), d1 = make_float4(vd1[0].y, vd1[1].y, vd1[2].y, vd1[3].y);
        const int kx = kx0 + 4 * g;
#pragma unroll
        for (int rr = 0; rr < 2; rr++) {
            const int ky = ky0 + 2 * yp + rr;
            if (ky >= Nr2) break;
            const float4 va = rr ? a1 : a0, vh = rr ? h1 : h0, vv = rr ? v1 : v0, vd = rr ? d1 : d0;
            const long long o = ob + (long long)ky * Nc2 + kx;
            if (kx + 3 < Nc2 && (Nc2 & 3) == 0) {
                *reinterpret_cast<float4*>(A + o) = va;
                *reinterpret_cast<float4*>(Hb + o) = vh;
                *reinterpret_cast<float4*>(V + o) = vv;
                *reinterpret_cast<float4*>(D + o) = vd;
            } else {
                const float ea[4] = {va.x, va.y, va.z, va.w}, eh[4] = {vh.x, vh.y, vh.z, vh.w};
                const float ev[4] = {vv.x, vv.y, vv.z, vv.w}, ed[4] = {vd.x, vd.y, vd.z, vd.w};
#pragma unroll
                for (int c = 0; c < 4; c++)
                    if (kx + c < Nc2) {
                        A[o + c] = ea[c]; Hb[o + c] = eh[c]; V[o + c] = ev[c]; D[o + c] = ed[c];
                    }
            }
        }
    }
}

// ---- inverse ------------------------------------------------------------------------------------
constexpr int BX = 32;        // band columns per tile (64 output columns)
constexpr int BY = 32;        // band rows per tile    (64 output rows)

template <int F>
struct InvGeo {
    static constexpr int Pp = F / 2 - 1, HALF = F / 2;
    static constexpr int S0 = Pp >> 1, E0 = Pp & 1, S1 = (Pp + 1) >> 1, E1 = (Pp + 1) & 1;
    static constexpr int HL = S1, HLr = (HL + 3) & ~3;
    static constexpr int BH = BY + 2 * HL;                    // band tile rows
    static constexpr int BW = BX + 2 * HLr;                   // band tile columns (aligned halo)
    static constexpr int P = BW + 4;                          // pitch
    static constexpr size_t smem = sizeof(float) * ((size_t)4 * BH * P + (size_t)2 * (2 * BY) * P) + sizeof(int) * BW;
};

template <int F>
__global__ void __launch_bounds__(NT)
k_tile_inv(const float* __restrict__ A, const float* __restrict__ Hb, const float* __restrict__ V,
           const float* __restrict__ D, float* __restrict__ out, int nr, int nc, int Nr_out, int Nc_out,
           long long in_bs, long long out_bs, const __grid_constant__ TapsInv f) {
    using G = InvGeo<F>;
    constexpr int HALF = G::HALF, S0 = G::S0, E0 = G::E0, S1 = G::S1, E1 = G::E1;
    constexpr int HL = G::HL, HLr = G::HLr, BH = G::BH, BW = G::BW, P = G::P;
    extern __shared__ __align__(16) float sm[];
    float* sb = sm;                          // [4][BH][P]
    float* st = sm + 4 * BH * P;             // [2][2*BY][P]
    int* colidx = reinterpret_cast<int*>(st + 2 * (2 * BY) * P);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int x0 = blockIdx.x * BX, y0 = blockIdx.y * BY;     // band coordinates of the tile
    const long long ib = blockIdx.z * in_bs;
    out += blockIdx.z * out_bs;
    const float* bands[4] = {A + ib, Hb + ib, V + ib, D + ib};

    const int xs = x0 - HLr;
    const bool vec = (nc & 3) == 0 && xs >= 0 && xs + BW <= nc && (in_bs & 3) == 0 &&
                     ((((uintptr_t)A) | ((uintptr_t)Hb) | ((uintptr_t)V) | ((uintptr_t)D)) & 15) == 0;
    if (!vec) {
        for (int i = tid; i < BW; i += NT) colidx[i] = wrap_per(xs + i, nc);
        __syncthreads();
    }
    for (int r = warp; r < 4 * BH; r += NT / 32) {
        const int b = r / BH, rr = r - b * BH;
        const float* row = bands[b] + (long long)wrap_per(y0 - HL + rr, nr) * nc;
        float* dst = sb + (b * BH + rr) * P;
        if (vec) {
            for (int c4 = lane; c4 < BW / 4; c4 += 32) cp_async16(dst + 4 * c4, row + xs + 4 * c4);
        } else {
            for (int cc = lane; cc < BW; cc += 32) cp_async4(dst + cc, row + colidx[cc]);
        }
    }
    cp_async_wait_all();
    __syncthreads();
    // ---- column synthesis: item = (4 band columns incl. halo, band row q) -> t1/t2 rows 2q, 2q+1 ----
    constexpr int NG = BW / 4;
    for (int it = tid; it < NG * BY; it += NT) {
        const int g = it % NG, q = it / NG;
        float2 t1[4], t2[4];                                    // (even row, odd row) of t1 / t2, 4 columns
#pragma unroll
        for (int c = 0; c < 4; c++) t1[c] = t2[c] = make_float2(0.f, 0.f);
        // band rows q - S1 .. q + S1 (local rows q .. q + 2*HL), window index w <-> band row q - S1 + w
        const float* pa = sb + (0 * BH + q) * P + 4 * g;
        const float* ph = sb + (1 * BH + q) * P + 4 * g;
        const float* pv = sb + (2 * BH + q) * P + 4 * g;
        const float* pd = sb + (3 * BH + q) * P + 4 * g;
#pragma unroll
        for (int w = 0; w <= 2 * HL; w++) {
            // row q - S1 + w is used by parity 0 with jj = S0 + S1 - w and by parity 1 with jj = 2*S1 - w
            const int je = S0 + S1 - w, jo = 2 * S1 - w;
            const bool ue = je >= 0 && je < HALF, uo = jo >= 0 && jo < HALF;
            if (ue || uo) {
                const float4 va = *reinterpret_cast<const float4*>(pa + w * P);
                const float4 vh = *reinterpret_cast<const float4*>(ph + w * P);
                const float4 vv = *reinterpret_cast<const float4*>(pv + w * P);
                const float4 vd = *reinterpret_cast<const float4*>(pd + w * P);
                const float ca[4] = {va.x, va.y, va.z, va.w}, ch[4] = {vh.x, vh.y, vh.z, vh.w};
                const float cv[4] = {vv.x, vv.y, vv.z, vv.w}, cd[4] = {vd.x, vd.y, vd.z, vd.w};
#pragma unroll
                for (int c = 0; c < 4; c++) {                   // a tap that a phase does not use is 0 in the table
                    t1[c] = fma2s(ca[c], f.l[w], t1[c]);
                    t1[c] = fma2s(ch[c], f.h[w], t1[c]);
                    t2[c] = fma2s(cv[c], f.l[w], t2[c]);
                    t2[c] = fma2s(cd[c], f.h[w], t2[c]);
                }
            }
        }
        *reinterpret_cast<float4*>(st + (2 * q) * P + 4 * g) = make_float4(t1[0].x, t1[1].x, t1[2].x, t1[3].x);
        *reinterpret_cast<float4*>(st + (2 * q + 1) * P + 4 * g) = make_float4(t1[0].y, t1[1].y, t1[2].y, t1[3].y);
        *reinterpret_cast<float4*>(st + (2 * BY + 2 * q) * P + 4 * g) = make_float4(t2[0].x, t2[1].x, t2[2].x, t2[3].x);
        *reinterpret_cast<float4*>(st + (2 * BY + 2 * q + 1) * P + 4 * g) = make_float4(t2[0].y, t2[1].y, t2[2].y, t2[3].y);
    }
    __syncthreads();
    // ---- row synthesis: item = (output row n, 4 band columns) -> 8 output columns ----
    constexpr int NV = (4 + 2 * HLr) / 4;
    for (int it = tid; it < (2 * BY) * (BX / 4); it += NT) {
        const int u = it % (BX / 4), n = it / (BX / 4);
        const int gy = 2 * y0 + n;
        if (gy >= Nr_out) continue;
        float v1[4 * NV], v2[4 * NV];
#pragma unroll
        for (int k = 0; k < NV; k++) {
            const float4 a = *reinterpret_cast<const float4*>(st + n * P + 4 * u + 4 * k);
            const float4 b = *reinterpret_cast<const float4*>(st + (2 * BY + n) * P + 4 * u + 4 * k);
            v1[4 * k] = a.x; v1[4 * k + 1] = a.y; v1[4 * k + 2] = a.z; v1[4 * k + 3] = a.w;
            v2[4 * k] = b.x; v2[4 * k + 1] = b.y; v2[4 * k + 2] = b.z; v2[4 * k + 3] = b.w;
        }
        float o[8];
#pragma unroll
        for (int c = 0; c < 4; c++) {
            // sample at band offset k = w - S1 feeds the even output with jj = S0 - k and the odd one with
            // jj = S1 - k; walking w downwards keeps the reference's summation order (jj ascending)
            float2 eo = make_float2(0.f, 0.f);
#pragma unroll
            for (int w = 2 * HL; w >= 0; w--) {
                const int je = S0 + S1 - w, jo = 2 * S1 - w;
                if ((je >= 0 && je < HALF) || (jo >= 0 && jo < HALF)) {
                    eo = fma2s(v1[HLr + c + w - S1], f.l[w], eo);
                    eo = fma2s(v2[HLr + c + w - S1], f.h[w], eo);
                }
            }
            o[2 * c] = eo.x;
            o[2 * c + 1] = eo.y;
        }
        const int gx = 2 * (x0 + 4 * u);
        float* dst = out + (long long)gy * Nc_out + gx;
        if (gx + 7 < Nc_out && (Nc_out & 3) == 0) {
            *reinterpret_cast<float4*>(dst) = make_float4(o[0], o[1], o[2], o[3]);
            *reinterpret_cast<float4*>(dst + 4) = make_float4(o[4], o[5], o[6], o[7]);
        } else {
#pragma unroll
            for (int c = 0; c < 8; c++)
                if (gx + c < Nc_out) dst[c] = o[c];
        }
    }
}

inline int cdiv(int a, int b) { return (a + b - 1) / b; }

template <int F>
int launch_fwd(const float* in, float* A, float* Hb, float* V, float* D, int batch, int Nr, int Nc,
               long long in_bs, long long out_bs, const PwtFilters& f, cudaStream_t st) {
    static PwtKernelOnce once;
    if (!pwt_kernel_once(once, k_tile_fwd<F>, NT, FwdGeo<F>::smem, FwdGeo<F>::smem)) return 0;
    dim3 grid(cdiv((Nc + 1) / 2, TX), cdiv((Nr + 1) / 2, TY), batch);
    const TapsFwd t = pwt_pack_taps_fwd(f, F);
    k_tile_fwd<F><<<grid, NT, FwdGeo<F>::smem, st>>>(in, A, Hb, V, D, Nr, Nc, in_bs, out_bs, t);
    return 1;
}
template <int F>
int launch_inv(const float* A, const float* Hb, const float* V, const float* D, float* out, int batch, int nr,
               int nc, int Nr_out, int Nc_out, long long in_bs, long long out_bs, const PwtFilters& f,
               cudaStream_t st) {
    static PwtKernelOnce once;
    if (!pwt_kernel_once(once, k_tile_inv<F>, NT, InvGeo<F>::smem, InvGeo<F>::smem)) return 0;
    dim3 grid(cdiv(nc, BX), cdiv(nr, BY), batch);
    const TapsInv t = pwt_pack_taps_inv(f, F);
    k_tile_inv<F><<<grid, NT, InvGeo<F>::smem, st>>>(A, Hb, V, D, out, nr, nc, Nr_out, Nc_out, in_bs, out_bs, t);
    return 1;
}

}  // namespace

#define PWT_TILE_CASES(X) X(10) X(12) X(14) X(16) X(18) X(20) X(22) X(24) X(26) X(28) X(30) X(32) X(34) X(36) X(38) X(40)

int pwt_tile_dwt_fwd2d(const float* in, float* A, float* Hb, float* V, float* D, int batch, int Nr, int Nc,
                       long long in_bs, long long out_bs, const PwtFilters& f, cudaStream_t st) {
    if (batch > 65535 || ((uintptr_t)A & 15) || ((uintptr_t)Hb & 15) || ((uintptr_t)V & 15) || ((uintptr_t)D & 15) ||
        (out_bs & 3))
        return 0;
    switch (f.hlen) {
#define X(FF) case FF: return launch_fwd<FF>(in, A, Hb, V, D, batch, Nr, Nc, in_bs, out_bs, f, st);
        PWT_TILE_CASES(X)
#undef X
        default: return 0;
    }
}

int pwt_tile_dwt_inv2d(const float* A, const float* Hb, const float* V, const float* D, float* out, int batch,
                       int nr, int nc, int Nr_out, int Nc_out, long long in_bs, long long out_bs,
                       const PwtFilters& f, cudaStream_t st) {
    if (batch > 65535 || ((uintptr_t)out & 15) || (out_bs & 3)) return 0;
    switch (f.hlen) {
#define X(FF) case FF: return launch_inv<FF>(A, Hb, V, D, out, batch, nr, nc, Nr_out, Nc_out, in_bs, out_bs, f, st);
        PWT_TILE_CASES(X)
#undef X
        default: return 0;
    }
}
