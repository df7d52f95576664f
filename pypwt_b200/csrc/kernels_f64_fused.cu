// Double precision, one 2D DWT level per launch with the row and the column pass FUSED (F = 4 .. 40, any size).
//
// The two-pass kernels of pwt_plan64.cu move 32 B per level-input sample and direction (the row-filtered planes make a
// round trip through HBM) and sit at the streaming limit for that traffic (8192^2 db2 3 levels fwd+inv 0.95 ms = 0.35 of
// the 32 B/px roofline of the double-precision transform).  Here a CTA owns a strip of 128 half-resolution columns and
// walks DOWN a segment of rows:
//   analysis   the input rows of the strip (256 + F - 2 samples each) are staged with cp.async through a ring; thread t
//              filters row n along x for its column (F/2 128-bit shared loads -> low-pass and high-pass sample), pushes the
//              pair into two rotating register windows of F rows, and every second row emits A, H, V, D of its column
//              (coalesced 8-byte stores).  Reference: w_kern_forward_pass1/2, separable.cu:98-197 with DTYPE = double.
//   synthesis  the rows of the four bands (128 + F/2 columns each) are staged the same way; thread t synthesises along x the
//              two output columns 2t, 2t + 1 of the two row planes (A, V -> low-pass plane; H, D -> high-pass plane), keeps
//              F/2 + 1 rows of them in rotating register windows, and every band row emits two output rows (128-bit
//              stores).  Reference: w_kern_inverse_pass1/2, separable.cu:252-361 (columns first there: same sums, the
//              result differs by fp64 rounding only).
// 16 B per level-input sample and direction: 8192^2 3 levels fwd+inv db2 0.95 -> 0.52 ms (0.63 of the roofline; 0.75 is the
// ceiling of one launch per level), sym8 1.33 -> 0.68, haar 0.94 -> 0.45.  The analysis keeps the summation order of the two-pass kernels exactly
// (bit-identical results); rotations are unrolled over their period so every register index is static.
#include <stdlib.h>

#include <type_traits>

#include "pwt_internal.h"

namespace {

constexpr int NT = 128;          // threads per CTA = half-resolution columns per strip
constexpr int TB = NT;

__device__ __forceinline__ int mod_pos(int i, int n) {
    i %= n;
    return i < 0 ? i + n : i;
}
// analysis extension (separable.cu:98-131): periodic over the size rounded up to even, the extra sample of an odd size
// repeats the last one
__device__ __forceinline__ int wrap_dwt64(int i, int N) {
    i = mod_pos(i, N + (N & 1));
    return i >= N ? N - 1 : i;
}
__device__ __forceinline__ void cp_async8(double* dst, const double* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// Staging: chunks of RC stream rows in a ring of NST buffers, chunk c + NST - 1 requested while chunk c is consumed.
// F <= 20: the register windows rotate (loops unrolled over the period, static indices).  F >= 22: that code would not
// fit the instruction cache (4 F^2 DFMA per period; ncu on the first version: 3.4 no-instruction stalls per issue for
// F = 40), so the windows SHIFT by register moves instead (+25 % issue slots, a loop body of one row pair).
template <int F>
struct FwdGeo {
    static constexpr int C = F / 2 - 1;
    static constexpr bool SHIFT = F >= 22;
    static constexpr int TI = 2 * TB + F - 2;               // staged samples per row
    static constexpr int PITCH = (TI + 1) & ~1;             // rows stay 16-byte aligned
    static constexpr int RC = (SHIFT || F % 4 == 0) ? 4 : 2;
    static constexpr int NST = RC == 2 ? 4 : 3;
    static constexpr int NS = (TI + NT - 1) / NT;
    static constexpr size_t smem = sizeof(double) * NST * RC * PITCH;
};

template <int F>
__global__ void __launch_bounds__(NT)
k64_fused_fwd(const double* __restrict__ in, double* __restrict__ A, double* __restrict__ Hb, double* __restrict__ V,
              double* __restrict__ D, int Nr, int Nc, long long in_bs, long long out_bs, int KS,
              const __grid_constant__ PwtFilters64 f) {
    using G = FwdGeo<F>;
    constexpr int C = G::C, TI = G::TI, PITCH = G::PITCH, RC = G::RC, NST = G::NST, NS = G::NS;
    extern __shared__ __align__(16) double smd[];
    const int tid = threadIdx.x;
    const int Nr2 = (Nr + 1) >> 1, Nc2 = (Nc + 1) >> 1, NrE = Nr + (Nr & 1);
    const int kx0 = blockIdx.x * TB;
    const int k0 = blockIdx.y * KS, kend = min(k0 + KS, Nr2);
    if (k0 >= kend) return;
    in += blockIdx.z * in_bs;
    const long long ob = blockIdx.z * out_bs;
    const int nrows = 2 * (kend - k0) + F - 2;             // stream rows: image rows 2 k0 - C ... (wrapped); even
    const int nchunks = (nrows + RC - 1) / RC;
    int colidx[NS];
#pragma unroll
    for (int s = 0; s < NS; s++) colidx[s] = wrap_dwt64(2 * kx0 - C + tid + s * NT, Nc);
    int se = mod_pos(2 * k0 - C, NrE);                      // image row (even extension) of the next row to stage
    auto stage = [&](int c) {                              // chunks are staged in order; always commits a group
        if (c < nchunks) {
            double* dst = smd + (c % NST) * RC * PITCH;
#pragma unroll
            for (int r = 0; r < RC; r++) {
                if (c * RC + r < nrows) {
                    const double* row = in + (long long)(se >= Nr ? Nr - 1 : se) * Nc;
                    if (++se == NrE) se = 0;
#pragma unroll
                    for (int s = 0; s < NS; s++)
                        if (s < NS - 1 || tid + s * NT < TI) cp_async8(dst + r * PITCH + tid + s * NT, row + colidx[s]);
                }
            }
        }
        cp_async_commit();
    };
    const int k = kx0 + tid;
    const bool colok = k < Nc2;
    double* oA = A + ob + k;
    double* oH = Hb + ob + k;
    double* oV = V + ob + k;
    double* oD = D + ob + k;
    double wl[F], wh[F];
#pragma unroll
    for (int j = 0; j < F; j++) wl[j] = wh[j] = 0.0;
    auto rowpass = [&](const double* rb, double& a, double& d) {       // taps ascending (the two-pass kernels' order)
        a = 0.0;
        d = 0.0;
#pragma unroll
        for (int j2 = 0; j2 < F / 2; j2++) {
            const double2 v = *reinterpret_cast<const double2*>(rb + 2 * j2);
            a = fma(v.x, f.L[F - 1 - 2 * j2], a);
            d = fma(v.x, f.H[F - 1 - 2 * j2], d);
            a = fma(v.y, f.L[F - 2 - 2 * j2], a);
            d = fma(v.y, f.H[F - 2 - 2 * j2], d);
        }
    };
    auto chunk_ready = [&](int c) {                        // chunk c landed and is visible; request chunk c + NST - 1
        cp_async_wait<NST - 2>();
        __syncthreads();                                   // ... and everybody is done with the buffer of chunk c - 1
        stage(c + NST - 1);
        if (c == nchunks - 1) pwt_pdl_trigger();
    };

    pwt_pdl_wait();
#pragma unroll
    for (int c = 0; c < NST - 1; c++) stage(c);
    for (int n0 = 0; n0 < nrows; n0 += F) {
#pragma unroll
        for (int u = 0; u < F; u++) {
            const int n = n0 + u;
            if (n < nrows) {                           // uniform over the CTA
                if (u % RC == 0) chunk_ready(n / RC);
                rowpass(smd + ((n / RC) % NST) * RC * PITCH + (u % RC) * PITCH + 2 * tid, wl[u], wh[u]);
                if ((u & 1) && n >= F - 1) {           // stream row n completes output k0 + (n - (F - 1)) / 2
                    double xa = 0.0, xh = 0.0, xv = 0.0, xd = 0.0;
#pragma unroll
                    for (int j = 0; j < F; j++) {
                        const int sl = (u + 1 + j) % F;
                        xa = fma(wl[sl], f.L[F - 1 - j], xa);
                        xh = fma(wl[sl], f.H[F - 1 - j], xh);
                        xv = fma(wh[sl], f.L[F - 1 - j], xv);
                        xd = fma(wh[sl], f.H[F - 1 - j], xd);
                    }
                    if (colok) {
                        const long long o = (long long)(k0 + ((n - (F - 1)) >> 1)) * Nc2;
                        oA[o] = xa;
                        oH[o] = xh;
                        oV[o] = xv;
                        oD[o] = xd;
                    }
                }
            }
        }
    }
}

// F >= 22: 256 threads, thread (plane, column) -- the low-pass and the high-pass row plane of a column are owned by two
// different threads, so a window is F doubles per thread (2 F would spill from F = 30 on) and twice the warps hide the
// latencies; the window shifts by register moves, the loop body is one row pair.  Same sums, same order.
template <int F>
__global__ void __launch_bounds__(2 * NT)
k64_fused_fwd_long(const double* __restrict__ in, double* __restrict__ A, double* __restrict__ Hb, double* __restrict__ V,
                   double* __restrict__ D, int Nr, int Nc, long long in_bs, long long out_bs, int KS,
                   const __grid_constant__ PwtFilters64 f) {
    using G = FwdGeo<F>;
    constexpr int C = G::C, TI = G::TI, PITCH = G::PITCH, RC = G::RC, NST = G::NST, NS = (TI + 2 * NT - 1) / (2 * NT);
    extern __shared__ __align__(16) double smd[];
    const int tid = threadIdx.x, pl = tid >> 7, t = tid & (NT - 1);
    const int Nr2 = (Nr + 1) >> 1, Nc2 = (Nc + 1) >> 1, NrE = Nr + (Nr & 1);
    const int kx0 = blockIdx.x * TB;
    const int k0 = blockIdx.y * KS, kend = min(k0 + KS, Nr2);
    if (k0 >= kend) return;
    in += blockIdx.z * in_bs;
    const long long ob = blockIdx.z * out_bs;
    const int nrows = 2 * (kend - k0) + F - 2;
    const int nchunks = (nrows + RC - 1) / RC;
    int colidx[NS];
#pragma unroll
    for (int s = 0; s < NS; s++) colidx[s] = wrap_dwt64(2 * kx0 - C + tid + s * 2 * NT, Nc);
    int se = mod_pos(2 * k0 - C, NrE);
    auto stage = [&](int c) {
        if (c < nchunks) {
            double* dst = smd + (c % NST) * RC * PITCH;
#pragma unroll
            for (int r = 0; r < RC; r++) {
                if (c * RC + r < nrows) {
                    const double* row = in + (long long)(se >= Nr ? Nr - 1 : se) * Nc;
                    if (++se == NrE) se = 0;
#pragma unroll
                    for (int s = 0; s < NS; s++)
                        if (tid + s * 2 * NT < TI) cp_async8(dst + r * PITCH + tid + s * 2 * NT, row + colidx[s]);
                }
            }
        }
        cp_async_commit();
    };
    const int k = kx0 + t;
    const bool colok = k < Nc2;
    double* o0 = (pl ? V : A) + ob + k;                    // low-pass down the column
    double* o1 = (pl ? D : Hb) + ob + k;                   // high-pass down the column
    double w[F];
#pragma unroll
    for (int j = 0; j < F; j++) w[j] = 0.0;
    auto rowpass1 = [&](const double* rb, auto hi) {        // this thread's row filter: taps stay constant-bank operands
        constexpr bool HI = decltype(hi)::value;
        double a = 0.0;
#pragma unroll
        for (int j2 = 0; j2 < F / 2; j2++) {
            const double2 v = *reinterpret_cast<const double2*>(rb + 2 * j2);
            a = fma(v.x, HI ? f.H[F - 1 - 2 * j2] : f.L[F - 1 - 2 * j2], a);
            a = fma(v.y, HI ? f.H[F - 2 - 2 * j2] : f.L[F - 2 - 2 * j2], a);
        }
        return a;
    };
    auto rowpass = [&](const double* rb) {                 // warp-uniform branch
        return pl ? rowpass1(rb, std::true_type()) : rowpass1(rb, std::false_type());
    };

    pwt_pdl_wait();
#pragma unroll
    for (int c = 0; c < NST - 1; c++) stage(c);
    for (int c = 0; c < nchunks; c++) {
        cp_async_wait<NST - 2>();
        __syncthreads();
        stage(c + NST - 1);
        if (c == nchunks - 1) pwt_pdl_trigger();
        const double* cb = smd + (c % NST) * RC * PITCH + 2 * t;
#pragma unroll
        for (int u = 0; u < RC; u += 2) {
            const int n = c * RC + u;                      // rows n, n + 1 (nrows is even)
            if (n < nrows) {
#pragma unroll
                for (int j = 0; j < F - 2; j++) w[j] = w[j + 2];
                w[F - 2] = rowpass(cb + u * PITCH);
                w[F - 1] = rowpass(cb + (u + 1) * PITCH);
                if (n + 1 >= F - 1) {
                    double x0 = 0.0, x1 = 0.0;
#pragma unroll
                    for (int j = 0; j < F; j++) {
                        x0 = fma(w[j], f.L[F - 1 - j], x0);
                        x1 = fma(w[j], f.H[F - 1 - j], x1);
                    }
                    if (colok) {
                        const long long o = (long long)(k0 + ((n + 1 - (F - 1)) >> 1)) * Nc2;
                        o0[o] = x0;
                        o1[o] = x1;
                    }
                }
            }
        }
    }
}

template <int F>
struct InvGeo {
    static constexpr int P = F / 2 - 1, HALF = F / 2;
    static constexpr bool SHIFT = F >= 22;
    static constexpr int HB = HALF - 1 - (P >> 1);          // band columns needed left of the strip
    static constexpr int S1 = (P + 1) >> 1, W = HALF + 1;   // column pass: window of W band rows starting at j - S1
    static constexpr int TI = TB + HALF;
    static constexpr int PITCH = TI;
    static constexpr int NST = 4;                           // one band row (of the four bands) per stage
    static constexpr int NS = (TI + NT - 1) / NT;
    static constexpr size_t smem = sizeof(double) * NST * 4 * PITCH;
};

template <int F>
__global__ void __launch_bounds__(NT)
k64_fused_inv(const double* __restrict__ A, const double* __restrict__ Hb, const double* __restrict__ V,
              const double* __restrict__ D, double* __restrict__ out, int nr, int nc, int Nro, int Nco, long long in_bs,
              long long out_bs, int KS, const __grid_constant__ PwtFilters64 f) {
    using G = InvGeo<F>;
    constexpr int P = G::P, HALF = G::HALF, HB = G::HB, S1 = G::S1, W = G::W, TI = G::TI, PITCH = G::PITCH, NST = G::NST, NS = G::NS;
    extern __shared__ __align__(16) double smd[];
    const int tid = threadIdx.x;
    const int x0 = blockIdx.x * TB;
    const int j0 = blockIdx.y * KS, jend = min(j0 + KS, nr);
    if (j0 >= jend) return;
    const long long ib = blockIdx.z * in_bs;
    A += ib; Hb += ib; V += ib; D += ib;
    out += blockIdx.z * out_bs;
    const int nrows = (jend - j0) + W - 1;                 // stream rows: band rows j0 - S1 ... (periodic)
    int colidx[NS];
#pragma unroll
    for (int s = 0; s < NS; s++) colidx[s] = mod_pos(x0 - HB + tid + s * NT, nc);
    int se = mod_pos(j0 - S1, nr);
    auto stage = [&](int n) {                              // rows are staged in order; always commits a group
        if (n < nrows) {
            double* dst = smd + (n % NST) * 4 * PITCH;
            const long long ro = (long long)se * nc;
            if (++se == nr) se = 0;
#pragma unroll
            for (int s = 0; s < NS; s++)
                if (s < NS - 1 || tid + s * NT < TI) {
                    double* q = dst + tid + s * NT;
                    const long long o = ro + colidx[s];
                    cp_async8(q, A + o);
                    cp_async8(q + PITCH, Hb + o);
                    cp_async8(q + 2 * PITCH, V + o);
                    cp_async8(q + 3 * PITCH, D + o);
                }
        }
        cp_async_commit();
    };
    const int m = 2 * (x0 + tid);
    const bool c0ok = m < Nco, c1ok = m + 1 < Nco;
    double* op = out + m;
    const bool v2 = c1ok && ((((uintptr_t)op) | ((uintptr_t)Nco * 8)) & 15) == 0;
    double wa[W][2], wd[W][2];
#pragma unroll
    for (int w = 0; w < W; w++) wa[w][0] = wa[w][1] = wd[w][0] = wd[w][1] = 0.0;
    // row synthesis of stream row n (separable.cu:293-328): output column 2 o + b reads band columns o + ((b + P) >> 1) - jj
    auto rowpass = [&](int n, double (&ua)[2], double (&ud)[2]) {
        cp_async_wait<NST - 2>();
        __syncthreads();
        stage(n + NST - 1);
        if (n == nrows - 1) pwt_pdl_trigger();
        const double* sa = smd + (n % NST) * 4 * PITCH + tid + HB;
#pragma unroll
        for (int b = 0; b < 2; b++) {
            const int t0 = (b + P) & 1, kb = (b + P) >> 1;
            double ul = 0.0, uh = 0.0;
#pragma unroll
            for (int jj = 0; jj < HALF; jj++) {
                ul = fma(sa[kb - jj], f.IL[2 * jj + t0], ul);
                ul = fma(sa[2 * PITCH + kb - jj], f.IH[2 * jj + t0], ul);
                uh = fma(sa[PITCH + kb - jj], f.IL[2 * jj + t0], uh);
                uh = fma(sa[3 * PITCH + kb - jj], f.IH[2 * jj + t0], uh);
            }
            ua[b] = ul;
            ud[b] = uh;
        }
    };
    auto store2 = [&](int row, double x0v, double x1v) {
        if (row < Nro) {
            double* o = op + (long long)row * Nco;
            if (v2) *reinterpret_cast<double2*>(o) = make_double2(x0v, x1v);
            else {
                if (c0ok) o[0] = x0v;
                if (c1ok) o[1] = x1v;
            }
        }
    };

    pwt_pdl_wait();
#pragma unroll
    for (int n = 0; n < NST - 1; n++) stage(n);
    if (!G::SHIFT) {
        for (int n0 = 0; n0 < nrows; n0 += W) {
#pragma unroll
            for (int u = 0; u < W; u++) {
                const int n = n0 + u;
                if (n < nrows) {                           // uniform over the CTA
                    rowpass(n, wa[u], wd[u]);
                    if (n >= W - 1) {                      // stream row n completes band row j0 + n - (W - 1): two output rows
#pragma unroll
                        for (int b = 0; b < 2; b++) {
                            const int t0 = (b + P) & 1, wb = ((b + P) >> 1) + S1;
                            double x0v = 0.0, x1v = 0.0;
#pragma unroll
                            for (int jj = 0; jj < HALF; jj++) {
                                const int sl = (wb - jj + u + 1) % W;
                                x0v = fma(wa[sl][0], f.IL[2 * jj + t0], x0v);
                                x0v = fma(wd[sl][0], f.IH[2 * jj + t0], x0v);
                                x1v = fma(wa[sl][1], f.IL[2 * jj + t0], x1v);
                                x1v = fma(wd[sl][1], f.IH[2 * jj + t0], x1v);
                            }
                            store2(2 * (j0 + n - (W - 1)) + b, x0v, x1v);
                        }
                    }
                }
            }
        }
    } else {
        for (int n = 0; n < nrows; n++) {
#pragma unroll
            for (int w = 0; w < W - 1; w++) {
                wa[w][0] = wa[w + 1][0]; wa[w][1] = wa[w + 1][1];
                wd[w][0] = wd[w + 1][0]; wd[w][1] = wd[w + 1][1];
            }
            rowpass(n, wa[W - 1], wd[W - 1]);
            if (n >= W - 1) {
#pragma unroll
                for (int b = 0; b < 2; b++) {
                    const int t0 = (b + P) & 1, wb = ((b + P) >> 1) + S1;
                    double x0v = 0.0, x1v = 0.0;
#pragma unroll
                    for (int jj = 0; jj < HALF; jj++) {
                        x0v = fma(wa[wb - jj][0], f.IL[2 * jj + t0], x0v);
                        x0v = fma(wd[wb - jj][0], f.IH[2 * jj + t0], x0v);
                        x1v = fma(wa[wb - jj][1], f.IL[2 * jj + t0], x1v);
                        x1v = fma(wd[wb - jj][1], f.IH[2 * jj + t0], x1v);
                    }
                    store2(2 * (j0 + n - (W - 1)) + b, x0v, x1v);
                }
            }
        }
    }
}

// ---- Haar (haar.cu:10-58 with DTYPE = double: the exact 1/2 butterfly, not the filter bank) ------------------------------
// One thread per pair of adjacent band columns: 2 x 32 bytes in, 4 x 16 bytes out (analysis) and the converse.  VEC: sizes
// and pointers allow 128-bit accesses; otherwise one band column per thread, scalar, with the odd-size rules (the analysis
// repeats the last row / column, the synthesis drops the extra ones).
template <bool VEC>
__global__ void __launch_bounds__(256)
k64_haar_fwd(const double* __restrict__ in, double* __restrict__ A, double* __restrict__ Hb, double* __restrict__ V,
             double* __restrict__ D, int Nr, int Nc, long long in_bs, long long out_bs) {
    const int Nr2 = (Nr + 1) >> 1, Nc2 = (Nc + 1) >> 1, PV = VEC ? Nc2 >> 1 : Nc2;
    in += blockIdx.z * in_bs;
    const long long ob = blockIdx.z * out_bs;
    pwt_pdl_wait();
    for (long long i = blockIdx.x * 256LL + threadIdx.x; i < (long long)Nr2 * PV; i += gridDim.x * 256LL) {
        const int k = (int)(i / PV), p = (int)(i - (long long)k * PV);
        const double* r0 = in + (long long)(2 * k) * Nc;
        const double* r1 = in + (long long)min(2 * k + 1, Nr - 1) * Nc;
        if (VEC) {
            const double2 u0 = __ldcs(reinterpret_cast<const double2*>(r0 + 4 * p)), u1 = __ldcs(reinterpret_cast<const double2*>(r0 + 4 * p + 2));
            const double2 l0 = __ldcs(reinterpret_cast<const double2*>(r1 + 4 * p)), l1 = __ldcs(reinterpret_cast<const double2*>(r1 + 4 * p + 2));
            const double s0 = u0.x + l0.x, t0 = u0.y + l0.y, d0 = u0.x - l0.x, e0 = u0.y - l0.y;
            const double s1 = u1.x + l1.x, t1 = u1.y + l1.y, d1 = u1.x - l1.x, e1 = u1.y - l1.y;
            const long long o = ob + (long long)k * Nc2 + 2 * p;
            *reinterpret_cast<double2*>(A + o) = make_double2(0.5 * (s0 + t0), 0.5 * (s1 + t1));
            *reinterpret_cast<double2*>(V + o) = make_double2(0.5 * (s0 - t0), 0.5 * (s1 - t1));
            *reinterpret_cast<double2*>(Hb + o) = make_double2(0.5 * (d0 + e0), 0.5 * (d1 + e1));
            *reinterpret_cast<double2*>(D + o) = make_double2(0.5 * (d0 - e0), 0.5 * (d1 - e1));
        } else {
            const int c0 = 2 * p, c1 = min(2 * p + 1, Nc - 1);
            const double a = r0[c0], b = r0[c1], c = r1[c0], d = r1[c1];
            const long long o = ob + (long long)k * Nc2 + p;
            A[o] = 0.5 * ((a + c) + (b + d));
            V[o] = 0.5 * ((a + c) - (b + d));
            Hb[o] = 0.5 * ((a - c) + (b - d));
            D[o] = 0.5 * ((a - c) - (b - d));
        }
    }
}
template <bool VEC>
__global__ void __launch_bounds__(256)
k64_haar_inv(const double* __restrict__ A, const double* __restrict__ Hb, const double* __restrict__ V,
             const double* __restrict__ D, double* __restrict__ out, int nr, int nc, int Nro, int Nco, long long in_bs,
             long long out_bs) {
    const int PV = VEC ? nc >> 1 : nc;
    const long long ib = blockIdx.z * in_bs;
    out += blockIdx.z * out_bs;
    pwt_pdl_wait();
    for (long long i = blockIdx.x * 256LL + threadIdx.x; i < (long long)nr * PV; i += gridDim.x * 256LL) {
        const int j = (int)(i / PV), p = (int)(i - (long long)j * PV);
        double* o0 = out + (long long)(2 * j) * Nco;
        double* o1 = o0 + Nco;
        const bool row1 = 2 * j + 1 < Nro;
        if (VEC) {
            const long long o = ib + (long long)j * nc + 2 * p;
            const double2 a = __ldcs(reinterpret_cast<const double2*>(A + o)), b = __ldcs(reinterpret_cast<const double2*>(V + o));
            const double2 c = __ldcs(reinterpret_cast<const double2*>(Hb + o)), d = __ldcs(reinterpret_cast<const double2*>(D + o));
            const double s0 = a.x + c.x, t0 = b.x + d.x, u0 = a.x - c.x, v0 = b.x - d.x;
            const double s1 = a.y + c.y, t1 = b.y + d.y, u1 = a.y - c.y, v1 = b.y - d.y;
            *reinterpret_cast<double2*>(o0 + 4 * p) = make_double2(0.5 * (s0 + t0), 0.5 * (s0 - t0));
            *reinterpret_cast<double2*>(o0 + 4 * p + 2) = make_double2(0.5 * (s1 + t1), 0.5 * (s1 - t1));
            if (row1) {
                *reinterpret_cast<double2*>(o1 + 4 * p) = make_double2(0.5 * (u0 + v0), 0.5 * (u0 - v0));
                *reinterpret_cast<double2*>(o1 + 4 * p + 2) = make_double2(0.5 * (u1 + v1), 0.5 * (u1 - v1));
            }
        } else {
            const long long o = ib + (long long)j * nc + p;
            const double a = A[o], b = V[o], c = Hb[o], d = D[o];
            const bool col1 = 2 * p + 1 < Nco;
            o0[2 * p] = 0.5 * ((a + c) + (b + d));
            if (col1) o0[2 * p + 1] = 0.5 * ((a + c) - (b + d));
            if (row1) {
                o1[2 * p] = 0.5 * ((a - c) + (b - d));
                if (col1) o1[2 * p + 1] = 0.5 * ((a - c) - (b - d));
            }
        }
    }
}
inline unsigned haar_grid(long long items) {
    long long g = (items + 255) / 256;
    const long long cap = (long long)pwt_sm_count() * 32;
    return (unsigned)(g < 1 ? 1 : (g > cap ? cap : g));
}
inline bool al16(const void* a, const void* b = nullptr, const void* c = nullptr, const void* d = nullptr, const void* e = nullptr) {
    return ((((uintptr_t)a) | ((uintptr_t)b) | ((uintptr_t)c) | ((uintptr_t)d) | ((uintptr_t)e)) & 15) == 0;
}

// Rows per segment: the grid (units x segments) should fill whole waves of the resident CTAs (cap); every segment
// re-reads `halo` stream rows.  Cost model: waves x (rows a CTA streams).
inline int pick_ks(int n_out, long long units, int cap, int rows_per_out, int halo, int min_ks) {
    int best_ks = n_out, max_seg = n_out / min_ks;
    if (max_seg < 1) max_seg = 1;
    if (max_seg > 256) max_seg = 256;
    long long best = -1;
    for (int want = 1; want <= max_seg; want++) {
        const int ks = (n_out + want - 1) / want, nseg = (n_out + ks - 1) / ks;
        const long long waves = (units * nseg + cap - 1) / cap;
        const long long cost = waves * ((long long)rows_per_out * ks + halo + 8);
        if (best < 0 || cost < best) { best = cost; best_ks = ks; }
    }
    return best_ks;
}
inline bool fused_enabled() {                              // PWT_F64_FUSED=0: the two-pass kernels (A/B, tests)
    static const bool on = [] { const char* e = getenv("PWT_F64_FUSED"); return !(e && *e == '0'); }();
    return on;
}
template <int F>
int launch_fwd(const double* in, double* A, double* Hb, double* V, double* D, int batch, int Nr, int Nc, long long in_bs,
               long long out_bs, const PwtFilters64& f, cudaStream_t st) {
    using G = FwdGeo<F>;
    static PwtKernelOnce once;
    constexpr int threads = G::SHIFT ? 2 * NT : NT;
    void (*kern)(const double*, double*, double*, double*, double*, int, int, long long, long long, int, PwtFilters64);
    if constexpr (G::SHIFT) kern = k64_fused_fwd_long<F>;
    else kern = k64_fused_fwd<F>;
    const int per_sm = pwt_kernel_once(once, kern, threads, G::smem, G::smem);
    if (!per_sm) return 0;
    const int Nr2 = (Nr + 1) >> 1, Nc2 = (Nc + 1) >> 1, strips = (Nc2 + TB - 1) / TB;
    const int KS = pick_ks(Nr2, (long long)strips * batch, per_sm * pwt_sm_count(), 2, F - 2, 2 * F);
    pwt_launch_pdl(kern, dim3(strips, (Nr2 + KS - 1) / KS, batch), threads, G::smem, st, in, A, Hb, V, D, Nr, Nc, in_bs, out_bs, KS, f);
    return 1;
}
template <int F>
int launch_inv(const double* A, const double* Hb, const double* V, const double* D, double* out, int batch, int nr, int nc,
               int Nro, int Nco, long long in_bs, long long out_bs, const PwtFilters64& f, cudaStream_t st) {
    using G = InvGeo<F>;
    static PwtKernelOnce once;
    const int per_sm = pwt_kernel_once(once, k64_fused_inv<F>, NT, G::smem, G::smem);
    if (!per_sm) return 0;
    const int strips = (nc + TB - 1) / TB;
    const int KS = pick_ks(nr, (long long)strips * batch, per_sm * pwt_sm_count(), 1, G::W - 1, 2 * G::W);
    pwt_launch_pdl(k64_fused_inv<F>, dim3(strips, (nr + KS - 1) / KS, batch), NT, G::smem, st, A, Hb, V, D, out, nr, nc, Nro, Nco, in_bs, out_bs, KS, f);
    return 1;
}
}  // namespace

#define PWT64F_CASES(X) X(4) X(6) X(8) X(10) X(12) X(14) X(16) X(18) X(20) X(22) X(24) X(26) X(28) X(30) X(32) X(34) X(36) X(38) X(40)

int pwt64_haar_fwd2d(const double* in, double* A, double* Hb, double* V, double* D, int batch, int Nr, int Nc, long long in_bs,
                     long long out_bs, cudaStream_t st) {
    if (!fused_enabled() || batch < 1 || batch > 65535 || Nr < 1 || Nc < 1) return 0;
    const int Nr2 = (Nr + 1) >> 1, Nc2 = (Nc + 1) >> 1;
    const bool vec = (Nc & 3) == 0 && al16(in, A, Hb, V, D) && (in_bs & 1) == 0 && (out_bs & 1) == 0;
    const dim3 grid(haar_grid((long long)Nr2 * (vec ? Nc2 / 2 : Nc2)), 1, batch);
    if (vec) pwt_launch_pdl(k64_haar_fwd<true>, grid, 256, 0, st, in, A, Hb, V, D, Nr, Nc, in_bs, out_bs);
    else pwt_launch_pdl(k64_haar_fwd<false>, grid, 256, 0, st, in, A, Hb, V, D, Nr, Nc, in_bs, out_bs);
    return 1;
}
int pwt64_haar_inv2d(const double* A, const double* Hb, const double* V, const double* D, double* out, int batch, int nr, int nc,
                     int Nro, int Nco, long long in_bs, long long out_bs, cudaStream_t st) {
    if (!fused_enabled() || batch < 1 || batch > 65535 || nr < 1 || nc < 1) return 0;
    const bool vec = (nc & 1) == 0 && Nco == 2 * nc && al16(out, A, Hb, V, D) && (in_bs & 1) == 0 && (out_bs & 1) == 0;
    const dim3 grid(haar_grid((long long)nr * (vec ? nc / 2 : nc)), 1, batch);
    if (vec) pwt_launch_pdl(k64_haar_inv<true>, grid, 256, 0, st, A, Hb, V, D, out, nr, nc, Nro, Nco, in_bs, out_bs);
    else pwt_launch_pdl(k64_haar_inv<false>, grid, 256, 0, st, A, Hb, V, D, out, nr, nc, Nro, Nco, in_bs, out_bs);
    return 1;
}
// One analysis level, [batch] images of Nr x Nc -> four bands of ceil(Nr/2) x ceil(Nc/2).  Returns the launches (1), 0: not covered.
int pwt64_fused_fwd2d(const double* in, double* A, double* Hb, double* V, double* D, int batch, int Nr, int Nc, long long in_bs,
                      long long out_bs, const PwtFilters64& f, cudaStream_t st) {
    if (!fused_enabled() || batch < 1 || batch > 65535 || Nr < 1 || Nc < 1) return 0;
    switch (f.hlen) {
#define X(FF) case FF: return launch_fwd<FF>(in, A, Hb, V, D, batch, Nr, Nc, in_bs, out_bs, f, st);
        PWT64F_CASES(X)
#undef X
    }
    return 0;
}
// One synthesis level, bands of nr x nc -> [batch] images of Nro x Nco (Nro in {2 nr - 1, 2 nr}, Nco likewise).
int pwt64_fused_inv2d(const double* A, const double* Hb, const double* V, const double* D, double* out, int batch, int nr, int nc,
                      int Nro, int Nco, long long in_bs, long long out_bs, const PwtFilters64& f, cudaStream_t st) {
    if (!fused_enabled() || batch < 1 || batch > 65535 || nr < 1 || nc < 1) return 0;
    switch (f.hlen) {
#define X(FF) case FF: return launch_inv<FF>(A, Hb, V, D, out, batch, nr, nc, Nro, Nco, in_bs, out_bs, f, st);
        PWT64F_CASES(X)
#undef X
    }
    return 0;
}
