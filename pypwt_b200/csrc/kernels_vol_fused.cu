// Volumetric DWT level with the x, y and z passes FUSED in one launch (short filters: F = 2, 4, 6).
//
// The volumetric plan ran a level as a batched 2D launch over the slices plus a z pass over the four sub-volumes: 16 B per
// voxel and direction against 8 compulsory, at the streaming limit for that traffic (512^3 db2, 3 levels, fwd+inv 0.82 ms =
// 6.0 TB/s moved, 0.40 of the roofline).  Here a CTA owns a tile of 32 x 32 half-resolution positions (64 x 64 voxels of a
// slice + the filter reach) and walks along z:
//   analysis   slice n of the tile is staged with cp.async (double buffer, 16-byte groups, periodic wrap per group / row /
//              slice); row pass x from shared memory -> low / high plane in shared memory; column pass y: a thread owns one
//              column of one x-plane and 8 consecutive half-rows -> 16 values of the slice's four 2D bands; z in TRANSPOSED
//              form: every value is added to the F/2 pending outputs of its position (rotating register accumulators with
//              static indices, (low, high) tap pairs as FFMA2); every second slice completes one output row of all eight
//              bands (128-byte coalesced stores).
//   synthesis  the mirror image: the eight band tiles of band slice k are staged; x synthesis -> four planes in shared
//              memory; y synthesis -> the thread's 8 rows x 2 columns of the low-z and the high-z plane; z synthesis in
//              transposed form into the pending output slices; every band slice completes two output slices (64-bit stores,
//              256 bytes per warp).
// Conventions as in pwt_vol.cu (periodisation over the size rounded up to even, band b = 4 dz + 2 dy + dx); the sums are the
// same as in the two-launch path, the order of the passes differs in the synthesis (rounding only).
#include <stdlib.h>

#include "pwt_internal.h"

namespace {

constexpr int TH = 32;            // half-resolution positions per tile side
constexpr int TX = 2 * TH;        // voxels per tile side

__device__ __forceinline__ int mod_pos(int i, int n) {
    i %= n;
    return i < 0 ? i + n : i;
}
__device__ __forceinline__ int wrap_dwt(int i, int N) {          // period N rounded up to even, x~[N] = x[N-1] (odd N)
    i = mod_pos(i, N + (N & 1));
    return i >= N ? N - 1 : i;
}
__device__ __forceinline__ void cp_async16(float* dst, const float* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ float2 fma2s(float x, float2 t, float2 acc) { return __ffma2_rn(make_float2(x, x), t, acc); }

struct VolBands {
    float* b[8];                  // b = 4 dz + 2 dy + dx; b[0] = the approximation of this level
};

// ---- analysis -----------------------------------------------------------------------------------------------------------
template <int F>
struct FwdGeo {
    static constexpr int C = F / 2 - 1, HALF = F / 2;
    static constexpr int TYS = TX + F - 2;                 // staged rows of a slice tile
    static constexpr int CL = 4;                           // the staged row starts CL (aligned) columns left of the tile
    static constexpr int NG = (CL + TX + 4 + 3) / 4;       // 16-byte groups per staged row (F <= 10: CL - C >= 0, reach right < 8)
    static constexpr int PX = 4 * NG;
    static constexpr int MH = F <= 6 ? 8 : 4;              // half-rows per thread in the y / z pass (z accumulators: 2 MH x F/2 pairs)
    static constexpr int NTF = 64 * (TH / MH);             // threads: 32 columns x 2 x-planes x TH / MH row groups
    static constexpr int MINB = F <= 6 ? 2 : 1;
    static constexpr int NS = (TYS * NG + NTF - 1) / NTF;
    static constexpr size_t smem = sizeof(float) * ((size_t)2 * TYS * PX + (size_t)TYS * TX);
};

template <int F>
__global__ void __launch_bounds__(FwdGeo<F>::NTF, FwdGeo<F>::MINB)
k_vol3_fwd(const float* __restrict__ in, const __grid_constant__ VolBands out, int Nz, int Ny, int Nx, int tiles_x, int KS,
           const __grid_constant__ PwtTapsFwd tp) {
    using G = FwdGeo<F>;
    constexpr int C = G::C, HALF = G::HALF, TYS = G::TYS, CL = G::CL, NG = G::NG, PX = G::PX, NS = G::NS, MH = G::MH, NT = G::NTF;
    extern __shared__ __align__(16) float sm[];
    float* raw = sm;                                       // [2][TYS][PX]
    float* rp = sm + 2 * TYS * PX;                         // [TYS][TX]: low-pass half | high-pass half of every row
    const int tid = threadIdx.x;
    const int Nz2 = (Nz + 1) >> 1, Ny2 = (Ny + 1) >> 1, Nx2 = (Nx + 1) >> 1, NzE = Nz + (Nz & 1);
    const int ty = blockIdx.x / tiles_x, tx = blockIdx.x - ty * tiles_x;
    const int hx0 = tx * TH, hy0 = ty * TH;
    const int k0 = blockIdx.y * KS, k1 = min(k0 + KS, Nz2);
    if (k0 >= k1) return;
    const int nsl = 2 * (k1 - k0) + F - 2;                 // stream slices: volume slices 2 k0 - C ... (wrapped); even
    const long long slice = (long long)Ny * Nx;

    // staging slots of this thread: (row of the tile, 16-byte group) -> shared offset, offset inside a slice
    int s_off[NS], s_img[NS];
#pragma unroll
    for (int s = 0; s < NS; s++) {
        const int idx = tid + s * NT, r = idx / NG, q = idx - r * NG;
        s_off[s] = r * PX + 4 * q;
        s_img[s] = wrap_dwt(2 * hy0 - C + r, Ny) * Nx + mod_pos(2 * hx0 - CL + 4 * q, Nx);
    }
    int sz = mod_pos(2 * k0 - C, NzE);                      // (even-extended) slice of the next stage
    auto stage = [&](int n) {
        if (n < nsl) {
            const float* src = in + (long long)(sz >= Nz ? Nz - 1 : sz) * slice;
            if (++sz == NzE) sz = 0;
            float* dst = raw + (n & 1) * TYS * PX;
#pragma unroll
            for (int s = 0; s < NS; s++)
                if (s < NS - 1 || tid + s * NT < TYS * NG) cp_async16(dst + s_off[s], src + s_img[s]);
        }
        cp_async_commit();
    };
    // y / z pass ownership: column cx of x-plane pl (0: low-pass along x, 1: high-pass), half-rows MH g .. MH g + MH - 1
    const int cx = tid & 31, pl = (tid >> 5) & 1, g = tid >> 6;
    const int hx = hx0 + cx;
    const bool colok = hx < Nx2;
    const float2 zero2 = make_float2(0.f, 0.f);
    float2 acc[2 * MH][HALF];                              // [2 m + dy][pending output]: (low-pass, high-pass) along z
#pragma unroll
    for (int v = 0; v < 2 * MH; v++)
#pragma unroll
        for (int a = 0; a < HALF; a++) acc[v][a] = zero2;
    float* ob[4];                                          // [2 dz + dy]
#pragma unroll
    for (int i = 0; i < 4; i++) ob[i] = out.b[4 * (i >> 1) + 2 * (i & 1) + pl] + (long long)(hy0 + MH * g) * Nx2 + hx;
    const long long oslice = (long long)Ny2 * Nx2;

    pwt_pdl_wait();
    stage(0);
    for (int n0 = 0; n0 < nsl; n0 += F) {
#pragma unroll
        for (int s = 0; s < F; s++) {
            const int n = n0 + s;
            if (n < nsl) {                                 // uniform over the CTA
                cp_async_wait<0>();
                __syncthreads();                           // slice n landed; everybody is done with rp and the other buffer
                stage(n + 1);
                if (n == nsl - 1) pwt_pdl_trigger();
                // ---- x: half a warp per row, lane -> two adjacent half-resolution columns from three 128-bit loads ----
#pragma unroll
                for (int i = 0; i < (TYS + NT / 16 - 1) / (NT / 16); i++) {
                    const int r = (tid >> 4) + (NT / 16) * i, l = tid & 15;
                    if (TYS % (NT / 16) == 0 || r < TYS) {
                        const float4* w4 = reinterpret_cast<const float4*>(raw + (n & 1) * TYS * PX + r * PX + 4 * l);
                        const float4 v0 = w4[0], v1 = w4[1], v2 = w4[2];
                        const float x[12] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w, v2.x, v2.y, v2.z, v2.w};
                        float2 p0 = zero2, p1 = zero2;
#pragma unroll
                        for (int j = 0; j < F; j++) {
                            p0 = fma2s(x[CL - C + j], tp.t[j], p0);
                            p1 = fma2s(x[CL - C + 2 + j], tp.t[j], p1);
                        }
                        *reinterpret_cast<float2*>(rp + r * TX + 2 * l) = make_float2(p0.x, p1.x);
                        *reinterpret_cast<float2*>(rp + r * TX + TH + 2 * l) = make_float2(p0.y, p1.y);
                    }
                }
                __syncthreads();
                // ---- y: MH half-rows of this thread's column ----
                float col[2 * MH + F - 2];
#pragma unroll
                for (int i = 0; i < 2 * MH + F - 2; i++) col[i] = rp[(2 * MH * g + i) * TX + pl * TH + cx];
#pragma unroll
                for (int m = 0; m < MH; m++) {
                    float2 q = zero2;
#pragma unroll
                    for (int j = 0; j < F; j++) q = fma2s(col[2 * m + j], tp.t[j], q);
                    // ---- z, transposed form: slice n feeds outputs u = n/2 - d with tap (n & 1) + 2 d ----
#pragma unroll
                    for (int dy = 0; dy < 2; dy++) {
                        const float v = dy ? q.y : q.x;
#pragma unroll
                        for (int d = 0; d < HALF; d++) {
                            const int a = (((s >> 1) - d) % HALF + HALF) % HALF, tj = (s & 1) + 2 * d;
                            acc[2 * m + dy][a] = fma2s(v, tp.t[tj], tj == 0 ? zero2 : acc[2 * m + dy][a]);
                        }
                    }
                }
                if (s & 1) {                               // completes output k0 + (n + 1) / 2 - F / 2
                    const int u = ((n + 1) >> 1) - HALF, a = ((s + 1) >> 1) % HALF;
                    if (u >= 0 && colok) {
                        const long long o = (long long)(k0 + u) * oslice;
#pragma unroll
                        for (int m = 0; m < MH; m++) {
                            if (hy0 + MH * g + m < Ny2) {
#pragma unroll
                                for (int dy = 0; dy < 2; dy++) {
                                    ob[dy][o + (long long)m * Nx2] = acc[2 * m + dy][a].x;
                                    ob[2 + dy][o + (long long)m * Nx2] = acc[2 * m + dy][a].y;
                                }
                            }
                        }
                    }
                }
            }
        }
    }
}

// ---- synthesis ----------------------------------------------------------------------------------------------------------
// Along every axis: output pair j (samples 2 j, 2 j + 1) = sum_w lo[j - S1 + w] * l[w] + hi[j - S1 + w] * h[w], w = 0 .. 2 S1, with
// the (even, odd) tap pairs of pwt_pack_taps_inv (F = 4, 6: S1 = 1, three window positions; F = 8, 10: S1 = 2, five) and
// periodic band indices.
constexpr int PXB = TH + 8;       // staged band columns: origin 4 columns left of the tile (16-byte groups)
constexpr int NGB = PXB / 4;
struct VolBandsIn {
    const float* b[8];
};
template <int F>
struct InvGeo {
    static constexpr int S1 = F == 2 ? 1 : (F / 2) >> 1;   // F = 2: its single window position is shifted to w = 1 by the launcher
    static constexpr int NW = 2 * S1 + 1;
    static constexpr int NRB = TH + 2 * S1;                // staged band rows (and used columns) of a tile
    static constexpr int MB = F <= 6 ? 4 : 2;              // band rows per thread in the y / z pass (z accumulators: 4 MB x NW pairs)
    static constexpr int NTI = 32 * (TH / MB);
    static constexpr int MINB = F <= 6 ? 2 : 1;
    static constexpr int NSB = (NRB * NGB + NTI - 1) / NTI;
    static constexpr size_t smem = sizeof(float) * ((size_t)8 * NRB * PXB + (size_t)4 * NRB * TX);
};

template <int F>
__global__ void __launch_bounds__(InvGeo<F>::NTI, InvGeo<F>::MINB)
k_vol3_inv(const __grid_constant__ VolBandsIn bands, float* __restrict__ out, int nz2, int ny2, int nx2, int Nz, int Ny, int Nx,
           int tiles_x, int KS, const __grid_constant__ PwtTapsInv tp) {
    using G = InvGeo<F>;
    constexpr int S1 = G::S1, NW = G::NW, NRB = G::NRB, MB = G::MB, NT = G::NTI, NSB = G::NSB;
    static_assert(S1 <= 4, "the staged rows start 4 columns left of the tile");
    extern __shared__ __align__(16) float sm[];
    float* raw = sm;                                       // [8 bands][NRB][PXB]
    float* up = sm + 8 * NRB * PXB;                        // [2 dz + dy][NRB][TX]: x-synthesised planes
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int ty = blockIdx.x / tiles_x, tx = blockIdx.x - ty * tiles_x;
    const int x0 = tx * TH, y0 = ty * TH;                  // first band column / row of the tile
    const int j0 = blockIdx.y * KS, j1 = min(j0 + KS, nz2);
    if (j0 >= j1) return;
    const int nsl = (j1 - j0) + NW - 1;                    // stream: band slices j0 - S1 ... j1 - 1 + S1 (periodic)
    const long long bslice = (long long)ny2 * nx2;
    // staging: (row, group) positions of this thread (NRB * NGB = 340 / 360 on 256 / 512 threads), all eight bands each
    int s_off[NSB], s_img[NSB];
#pragma unroll
    for (int s = 0; s < NSB; s++) {
        const int idx = tid + s * NT, r = idx / NGB, q = idx - r * NGB;
        s_off[s] = r * PXB + 4 * q;
        s_img[s] = mod_pos(y0 - S1 + r, ny2) * nx2 + mod_pos(x0 - 4 + 4 * q, nx2);
    }
    int sz = mod_pos(j0 - S1, nz2);
    auto stage = [&](int n) {
        if (n < nsl) {
            const long long so = (long long)sz * bslice;
            if (++sz == nz2) sz = 0;
#pragma unroll
            for (int b = 0; b < 8; b++)
#pragma unroll
                for (int s = 0; s < NSB; s++)
                    if (tid + s * NT < NRB * NGB) cp_async16(raw + b * NRB * PXB + s_off[s], bands.b[b] + so + s_img[s]);
        }
        cp_async_commit();
    };
    // y / z ownership: output columns 2 cxp, 2 cxp + 1 of the tile, band rows MB g .. MB g + MB - 1 (2 MB output rows)
    const int cxp = lane, g = warp;
    const float2 zero2 = make_float2(0.f, 0.f);
    float2 acc[4 * MB][NW];                                // [4 jy + 2 by + col][pending slice pair]: (slice 2 j, slice 2 j + 1)
#pragma unroll
    for (int v = 0; v < 4 * MB; v++)
#pragma unroll
        for (int a = 0; a < NW; a++) acc[v][a] = zero2;
    const int ocol = 2 * (x0 + cxp), orow = 2 * (y0 + MB * g);
    const bool colok = ocol < Nx;
    float* op = out + (long long)orow * Nx + ocol;
    const long long oslice = (long long)Ny * Nx;

    pwt_pdl_wait();
    stage(0);
    for (int n0 = 0; n0 < nsl; n0 += NW) {
#pragma unroll
        for (int s = 0; s < NW; s++) {
            const int n = n0 + s;
            if (n < nsl) {                                 // uniform over the CTA
                cp_async_wait<0>();
                __syncthreads();                           // band slice n landed; everybody is done with `up`
                // ---- x: warp -> staged rows warp, warp + 8, ...; lane -> band column o -> output columns 2 o, 2 o + 1 ----
#pragma unroll
                for (int i = 0; i < (NRB + NT / 32 - 1) / (NT / 32); i++) {
                    const int r = warp + (NT / 32) * i;
                    if (r < NRB) {
#pragma unroll
                        for (int q = 0; q < 4; q++) {      // q = 2 dz + dy: bands 2 q (low-pass along x) and 2 q + 1
                            const float* lo = raw + (2 * q) * NRB * PXB + r * PXB + lane + (4 - S1);
                            const float* hi = lo + NRB * PXB;
                            float2 e = zero2;
#pragma unroll
                            for (int w = 0; w < NW; w++) {
                                e = fma2s(lo[w], tp.l[w], e);
                                e = fma2s(hi[w], tp.h[w], e);
                            }
                            *reinterpret_cast<float2*>(up + q * NRB * TX + r * TX + 2 * lane) = e;
                        }
                    }
                }
                __syncthreads();                           // `up` complete, raw free
                stage(n + 1);
                if (n == nsl - 1) pwt_pdl_trigger();
                // ---- y, then z in transposed form: band slice n feeds the slice pairs j0 + n - w with z tap pair w ----
#pragma unroll
                for (int dz = 0; dz < 2; dz++) {
                    float2 wl[MB + NW - 1], wh[MB + NW - 1];   // staged rows MB g ... of the low-y and the high-y plane
#pragma unroll
                    for (int i = 0; i < MB + NW - 1; i++) {
                        wl[i] = *reinterpret_cast<const float2*>(up + (2 * dz) * NRB * TX + (MB * g + i) * TX + 2 * cxp);
                        wh[i] = *reinterpret_cast<const float2*>(up + (2 * dz + 1) * NRB * TX + (MB * g + i) * TX + 2 * cxp);
                    }
#pragma unroll
                    for (int jy = 0; jy < MB; jy++) {
                        float2 c0 = zero2, c1 = zero2;     // (row 2 jy, row 2 jy + 1) of the two columns
#pragma unroll
                        for (int w = 0; w < NW; w++) {
                            c0 = fma2s(wl[jy + w].x, tp.l[w], c0); c0 = fma2s(wh[jy + w].x, tp.h[w], c0);
                            c1 = fma2s(wl[jy + w].y, tp.l[w], c1); c1 = fma2s(wh[jy + w].y, tp.h[w], c1);
                        }
                        const float v[4] = {c0.x, c1.x, c0.y, c1.y};   // [2 by + col]
#pragma unroll
                        for (int e = 0; e < 4; e++)
#pragma unroll
                            for (int w = 0; w < NW; w++) {
                                const int a = ((s - w) % NW + NW) % NW;
                                const float2 t = dz ? tp.h[w] : tp.l[w];
                                acc[4 * jy + e][a] = fma2s(v[e], t, (w == 0 && dz == 0) ? zero2 : acc[4 * jy + e][a]);
                            }
                    }
                }
                {                                          // completes the slice pair j = j0 + n - (NW - 1)
                    const int j = j0 + n - (NW - 1), a = ((s - (NW - 1)) % NW + NW) % NW;
                    if (n >= NW - 1 && colok) {
#pragma unroll
                        for (int bz = 0; bz < 2; bz++) {
                            if (2 * j + bz < Nz) {
                                float* o = op + (long long)(2 * j + bz) * oslice;
#pragma unroll
                                for (int jy = 0; jy < MB; jy++)
#pragma unroll
                                    for (int by = 0; by < 2; by++)
                                        if (orow + 2 * jy + by < Ny)
                                            *reinterpret_cast<float2*>(o + (long long)(2 * jy + by) * Nx) =
                                                bz ? make_float2(acc[4 * jy + 2 * by][a].y, acc[4 * jy + 2 * by + 1][a].y)
                                                   : make_float2(acc[4 * jy + 2 * by][a].x, acc[4 * jy + 2 * by + 1][a].x);
                            }
                        }
                    }
                }
            }
        }
    }
}

// Rows per segment: whole waves of the resident CTAs; every segment re-reads `halo` stream slices.
inline int pick_ks(int n_out, long long units, int cap, int rows_per_out, int halo, int min_ks) {
    int best_ks = n_out, max_seg = n_out / min_ks;
    if (max_seg < 1) max_seg = 1;
    if (max_seg > 256) max_seg = 256;
    long long best = -1;
    for (int want = 1; want <= max_seg; want++) {
        const int ks = (n_out + want - 1) / want, nseg = (n_out + ks - 1) / ks;
        const long long waves = (units * nseg + cap - 1) / cap;
        const long long cost = waves * ((long long)rows_per_out * ks + halo + 2);
        if (best < 0 || cost < best) { best = cost; best_ks = ks; }
    }
    return best_ks;
}
inline bool vol_fused_enabled() {                          // PWT_VOL_FUSED=0: batched 2D level + z pass (A/B, tests)
    static const bool on = [] { const char* e = getenv("PWT_VOL_FUSED"); return !(e && *e == '0'); }();
    return on;
}

template <int F>
int launch_fwd(const float* in, const VolBands& out, int Nz, int Ny, int Nx, const PwtFilters& f, cudaStream_t st) {
    using G = FwdGeo<F>;
    static PwtKernelOnce once;
    const int per_sm = pwt_kernel_once(once, k_vol3_fwd<F>, G::NTF, G::smem, G::smem);
    if (!per_sm) return 0;
    const int Nz2 = (Nz + 1) >> 1, Ny2 = (Ny + 1) >> 1, Nx2 = (Nx + 1) >> 1;
    const int tiles_x = (Nx2 + TH - 1) / TH, tiles_y = (Ny2 + TH - 1) / TH;
    const int KS = pick_ks(Nz2, (long long)tiles_x * tiles_y, per_sm * pwt_sm_count(), 2, F - 2, 2 * F);
    const int nseg = (Nz2 + KS - 1) / KS;
    if (nseg > 65535) return 0;
    pwt_launch_pdl(k_vol3_fwd<F>, dim3(tiles_x * tiles_y, nseg), G::NTF, G::smem, st, in, out, Nz, Ny, Nx, tiles_x, KS, pwt_pack_taps_fwd(f, F));
    return 1;
}
template <int F>
int launch_inv(const VolBandsIn& bands, float* out, int nz2, int ny2, int nx2, int Nz, int Ny, int Nx, const PwtFilters& f, cudaStream_t st) {
    using G = InvGeo<F>;
    static PwtKernelOnce once;
    const int per_sm = pwt_kernel_once(once, k_vol3_inv<F>, G::NTI, G::smem, G::smem);
    if (!per_sm) return 0;
    const int tiles_x = (nx2 + TH - 1) / TH, tiles_y = (ny2 + TH - 1) / TH;
    const int KS = pick_ks(nz2, (long long)tiles_x * tiles_y, per_sm * pwt_sm_count(), 1, G::NW - 1, 2 * G::NW);
    const int nseg = (nz2 + KS - 1) / KS;
    if (nseg > 65535) return 0;
    PwtTapsInv t = pwt_pack_taps_inv(f, F);
    if (F == 2) {                                          // S1 = 0: the window of pair j is band sample j alone -> position w = 1
        t.l[1] = t.l[0]; t.h[1] = t.h[0];
        t.l[0] = t.h[0] = make_float2(0.f, 0.f);
    }
    pwt_launch_pdl(k_vol3_inv<F>, dim3(tiles_x * tiles_y, nseg), G::NTI, G::smem, st, bands, out, nz2, ny2, nx2, Nz, Ny, Nx, tiles_x, KS, t);
    return 1;
}
}  // namespace

// One analysis level of a volume: in [Nz][Ny][Nx] -> eight bands [ceil(Nz/2)][ceil(Ny/2)][ceil(Nx/2)], bands[b], b = 4 dz + 2 dy + dx.
// Returns the launches (1), 0 when not covered (filter length, width not a multiple of 4, small or misaligned volumes).
int pwt_vol_fused_fwd(const float* in, float* const* bands, int Nz, int Ny, int Nx, const PwtFilters& f, cudaStream_t st) {
    if (!vol_fused_enabled() || (Nx & 3) || Nx < 80 || Ny < 8 || Nz < 2 || (((uintptr_t)in) & 15) || (long long)Ny * Nx >= (1LL << 31)) return 0;
    VolBands vb;
    for (int b = 0; b < 8; b++) vb.b[b] = bands[b];
    switch (f.hlen) {
        case 2: return launch_fwd<2>(in, vb, Nz, Ny, Nx, f, st);
        case 4: return launch_fwd<4>(in, vb, Nz, Ny, Nx, f, st);
        case 6: return launch_fwd<6>(in, vb, Nz, Ny, Nx, f, st);
        // F = 8, 10 (512 threads, 4 half-rows per thread, one CTA per SM) were built and measured: 512^3 db4 1.015 ms against
        // 0.986 ms for the batched 2D launch + z pass, db5 1.074 against 1.055 -- not dispatched
    }
    return 0;
}
// One synthesis level: eight bands [nz2][ny2][nx2] -> out [Nz][Ny][Nx] (Nz in {2 nz2 - 1, 2 nz2}, Ny likewise, Nx = 2 nx2).
int pwt_vol_fused_inv(const float* const* bands, float* out, int nz2, int ny2, int nx2, int Nz, int Ny, int Nx, const PwtFilters& f,
                      cudaStream_t st) {
    if (!vol_fused_enabled() || (nx2 & 3) || Nx != 2 * nx2 || nx2 < 40 || ny2 < 4 || nz2 < 2 || (((uintptr_t)out) & 7) ||
        (long long)Ny * Nx >= (1LL << 31))
        return 0;
    VolBandsIn vb;
    for (int b = 0; b < 8; b++) {
        if (((uintptr_t)bands[b]) & 15) return 0;
        vb.b[b] = bands[b];
    }
    switch (f.hlen) {
        case 2: return launch_inv<2>(vb, out, nz2, ny2, nx2, Nz, Ny, Nx, f, st);
        case 4: return launch_inv<4>(vb, out, nz2, ny2, nx2, Nz, Ny, Nx, f, st);
        case 6: return launch_inv<6>(vb, out, nz2, ny2, nx2, Nz, Ny, Nx, f, st);
    }
    return 0;
}
