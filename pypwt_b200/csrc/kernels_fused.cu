// Three decomposition levels in ONE launch, entirely in registers (forward and inverse).
//
// With one launch per level the intermediate approximations A1 and A2 make a round trip through
// memory (10.5 B per pixel and per direction instead of the compulsory 8).  Here a warp streams down
// a strip of the image and runs the three levels as a cascade:
//
//   forward   input row pair -> A1 row (2 columns per lane)  -> level-1 details stored
//             A1 row pair    -> A2 row (1 column per lane)   -> level-2 details stored
//             A2 row pair    -> A3 row (every other lane)    -> level-3 details + A3 stored
//   inverse   the mirror image: A3 + details3 -> A2 rows -> (+ details2) A1 rows -> (+ details1) image
//
// Horizontal neighbours come from warp shuffles; what a warp cannot get from its own lanes (the
// cascade of the neighbouring strip) is recomputed: strips overlap, and a warp OWNS (stores) fewer
// columns than it loads (forward db2: 112 of 128).  The same holds vertically: a task warms its
// sliding windows up on 8+6 extra input rows.  The arithmetic order inside each level is the same as
// in the single-level kernels of kernels_reg.cu, so results are bit-identical to three launches.
//
// Requirements (checked by the dispatcher): even filter length <= 8, rows and columns multiples of 8
// (every level then has an even size: pure periodic wrap, no odd-size extension), >= 3 levels.
#include <stdlib.h>

#include <type_traits>

#include "pwt_internal.h"

namespace {

__device__ __forceinline__ int wrap1_per(int i, int N) {   // -N <= i < 2N
    if (i < 0) i += N;
    if (i >= N) i -= N;
    return i;
}
__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ void stg2(float* p, float a, float b) { *reinterpret_cast<float2*>(p) = make_float2(a, b); }

constexpr unsigned FULL = 0xffffffffu;
constexpr int kWarps = 4;

// ---- asynchronous bulk copies (TMA, SASS UBLKCP) global -> shared, completion on an mbarrier ----
// Used by the PF ("prefetch") variant of the forward cascade: the input rows of the next two iterations are
// in flight in a per-warp shared-memory ring while the warp computes, at no register cost.
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned mb, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mb), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned mb, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mb), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(unsigned dst, const void* src, unsigned bytes, unsigned mb) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(mb) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned mb, unsigned parity) {
    unsigned ok;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(mb), "r"(parity) : "memory");
    } while (!ok);
}
__device__ __forceinline__ float4 lds4(unsigned addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ float lds1(unsigned addr) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ void cp_async16(unsigned dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async4(unsigned dst, const void* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
constexpr int kStages = 2;          // iterations of input rows in flight per warp

__device__ __forceinline__ float comp(const float4& v, int c) {
    return c == 0 ? v.x : (c == 1 ? v.y : (c == 2 ? v.z : v.w));
}

struct Fwd3Args {
    const float* in;
    float* A3;
    float* H[3];
    float* V[3];
    float* D[3];
    int Nr, Nc;            // level-0 size
    int n3;                // level-3 columns owned by one warp
    int T3;                // level-3 rows per task
    long long in_bs;       // batch strides (elements) of the input and of the level-1/2/3 planes
    long long bs[3];
    unsigned* counter;     // dynamic task queue: task = atomicAdd(counter, 1) - base
    unsigned base;
    int ntasks;            // tasks per image (strips * bands); total = ntasks * batch
    int batch;
    double* partials;      // [ntasks*batch][2]: per-task sum |c| and sum c^2 of everything the task stored (or null)
    int count_a3;          // A3 is the final approximation (3-level transform): include it in the sums
    long long dV[3], dD[3], dA3;   // byte distance from the H plane of a level to its V / D plane (and level-3 H -> A3)
};

// analysis taps packed for the 2-wide FMA (FFMA2): t[j] = (L[F-1-j], H[F-1-j]).  One FFMA2 multiplies a
// broadcast sample by the (low-pass, high-pass) pair, so every multiply-add of the cascade is issued 2-wide.
struct TapsLH {
    float2 t[8];
};
__device__ __forceinline__ float2 fma2s(float x, float2 t, float2 acc) { return __ffma2_rn(make_float2(x, x), t, acc); }

// geometry shared by host and device
template <int F>
struct Geo {
    static constexpr int C = F / 2 - 1;                 // left reach of an analysis window
    static constexpr int O1 = (3 * C + 1) & ~1;         // level-1 columns loaded left of the owned region (even)
    static constexpr int D2 = O1 / 2;                   // level-2 lane offset of the owned region
    // owned level-3 columns per warp: 64 level-1 columns must cover [4c3 - 3C, 4c3 + 4n - 4 + 3F/2]
    static constexpr int N3 = (63 - O1 - 3 * F / 2 + 4) / 4;
};

template <int F, bool HAAR, int MINB, int PF, bool NRM>      // PF: 0 direct loads, 1 bulk-copy (TMA) ring, 2 cp.async ring, 3 direct loads pipelined in place; NRM: accumulate norms
__global__ void __launch_bounds__(32 * kWarps, MINB)
k_fwd3(const __grid_constant__ Fwd3Args a, const __grid_constant__ TapsLH f) {
    using G = Geo<F>;
    constexpr int C = G::C;
    pwt_pdl_wait();                       // programmatic dependent launch (pwt_internal.h): nothing global is touched before
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int Nr = a.Nr, Nc = a.Nc;
    const int W1 = Nc >> 1, W2 = Nc >> 2, W3 = Nc >> 3;
    const int R3 = Nr >> 3;
    const int strips = (W3 + a.n3 - 1) / a.n3;
    // PF: per-warp ring of kStages x 8 input rows of HB + 128 + HB samples, filled by bulk copies
    constexpr bool RING = PF == 1 || PF == 2;
    constexpr int HB = (RING && C > 0) ? 4 : 0;              // halo block (16 B aligned) on either side
    constexpr int ROWB = (128 + 2 * HB) * 4;                 // bytes per staged row
    __shared__ __align__(128) float ring[RING ? kWarps * kStages * 8 * (128 + 2 * HB) : 1];
    __shared__ __align__(8) unsigned long long mbars[RING ? kWarps * kStages : 1];
    const unsigned ring0 = smem_u32(ring) + warp * (kStages * 8 * ROWB);
    const unsigned mbar0 = smem_u32(mbars) + warp * (kStages * 8);
    unsigned uses = 0;                                       // stage uses so far: stage = uses % kStages, parity from uses / kStages
    if (PF == 1) {
        if (lane == 0) {
#pragma unroll
            for (int s = 0; s < kStages; s++) mbar_init(mbar0 + 8 * s, 8);     // lanes 0..7 arrive, one row each
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        }
        __syncwarp();
    }
  for (;;) {
    // persistent warp: pull the next (image, band, strip) task from the queue
    unsigned t_ = 0;
    if (lane == 0) t_ = atomicAdd(a.counter, 1u) - a.base;
    t_ = __shfl_sync(FULL, t_, 0);
    // last wave of tasks (or none left): the next launch may start being scheduled into the slots that free up
    if (t_ + gridDim.x * kWarps >= (unsigned)a.ntasks * (unsigned)a.batch) pwt_pdl_trigger();
    if (t_ >= (unsigned)a.ntasks * (unsigned)a.batch) return;
    const int img = t_ / a.ntasks, task = t_ - img * a.ntasks;
    const int strip = task % strips, band = task / strips;
    const int n0 = band * a.T3;
    const int n1 = min(n0 + a.T3, R3);
    const int c3 = strip * a.n3;                         // first owned level-3 column
    const int n3e = min(a.n3, W3 - c3);                  // owned level-3 columns (last strip may be short)
    const int K1 = 4 * c3 - G::O1;                       // level-1 column of lane 0 (may be negative: wraps)
    const int X0 = 2 * K1;                               // input column of lane 0

    const float* in = a.in + img * a.in_bs;
    const int xcol = wrap1_per(X0 + 4 * lane, Nc);       // Nc % 4 == 0: a lane's 4 samples never straddle the wrap
    // halo samples of the warp's strip: lane 0 fetches the C samples left of it, lane 31 the C samples
    // right of it; every other lane re-reads its own first sample (same sector as its 128-bit load), so
    // the extra load is branch-free.
    int ecol[C > 0 ? C : 1];
#pragma unroll
    for (int i = 0; i < C; i++)
        ecol[i] = lane == 0 ? wrap1_per(X0 - C + i, Nc) : (lane == 31 ? wrap1_per(wrap1_per(X0 + 128, Nc) + i, Nc) : xcol);
    // ownership of this lane's columns
    const int k1 = K1 + 2 * lane;                        // level-1 columns k1, k1+1
    const bool own1 = k1 >= 4 * c3 && k1 < 4 * (c3 + n3e);
    const int k2 = (K1 >> 1) + lane;                     // level-2 column (K1 is even)
    const bool own2 = k2 >= 2 * c3 && k2 < 2 * (c3 + n3e);
    const int i3 = lane - G::D2;                         // level-3: column c3 + i3/2 lives on lane D2 + 2*i
    const int k3 = c3 + (i3 >> 1);
    const bool own3 = i3 >= 0 && !(i3 & 1) && (i3 >> 1) < n3e;
    // one pointer per level: the H plane at this lane's column, row 0; a row is reached with a single
    // multiply-add (IMAD.WIDE), the V / D / A planes with a warp-uniform byte distance
    char* const q1 = reinterpret_cast<char*>(a.H[0] + img * a.bs[0] + k1);
    char* const q2 = reinterpret_cast<char*>(a.H[1] + img * a.bs[1] + k2);
    char* const q3 = reinterpret_cast<char*>(a.H[2] + img * a.bs[2] + k3);
    const int W1b = W1 * 4, W2b = W2 * 4, W3b = W3 * 4;

    // ---- level 1: horizontal pass of one input row (same arithmetic as k_fwd_reg) ----
    // eight consecutive rows starting at rb; the wrap test is hoisted out (uniform)
    auto load_rows8 = [&](int rb, float4* v, float (*e)[C > 0 ? C : 1]) {
        if (rb >= 0 && rb + 8 <= Nr) {
            const float* p = in + (long long)rb * Nc;
#pragma unroll
            for (int i = 0; i < 8; i++, p += Nc) {
                v[i] = ldg4(p + xcol);
#pragma unroll
                for (int c = 0; c < C; c++) e[i][c] = __ldg(p + ecol[c]);
            }
        } else {
#pragma unroll
            for (int i = 0; i < 8; i++) {
                const float* p = in + (long long)wrap1_per(rb + i, Nr) * Nc;
                v[i] = ldg4(p + xcol);
#pragma unroll
                for (int c = 0; c < C; c++) e[i][c] = __ldg(p + ecol[c]);
            }
        }
    };
    const bool first_lane = lane == 0, last_lane = lane == 31;
    auto hpass1 = [&](const float4& v, const float* e) -> float4 {
        if (HAAR) return v;      // raw samples: the Haar butterfly combines rows first (haar.cu:27-35)
        float ext[F + 2];
#pragma unroll
        for (int c = 0; c < 4; c++) ext[C + c] = comp(v, c);
#pragma unroll
        for (int i = 0; i < C; i++) {
            const float l = __shfl_up_sync(FULL, comp(v, 4 - C + i), 1);
            const float r = __shfl_down_sync(FULL, comp(v, i), 1);
            ext[i] = first_lane ? e[i] : l;
            ext[C + 4 + i] = last_lane ? e[i] : r;
        }
        float2 p0 = make_float2(0.f, 0.f), p1 = p0;          // (lo0, hi0), (lo1, hi1)
#pragma unroll
        for (int j = 0; j < F; j++) {
            p0 = fma2s(ext[j], f.t[j], p0);
            p1 = fma2s(ext[j + 2], f.t[j], p1);
        }
        return make_float4(p0.x, p0.y, p1.x, p1.y);
    };
    // ---- levels 2 and 3: horizontal pass of an approximation row held one/two columns per lane ----
    // level 2: this lane's output column is 2*k2 = its own pair (a0, a1); offset d lives in lane + floor(d/2)
    auto hpass2 = [&](float a0, float a1) -> float2 {
        if (HAAR) return make_float2(a0, a1);
        float2 p = make_float2(0.f, 0.f);                    // (lo, hi)
#pragma unroll
        for (int j = 0; j < F; j++) {
            const int d = j - C;                         // column offset relative to a0
            const int dl = d >= 0 ? d / 2 : -((1 - d) / 2);   // floor(d / 2)
            const float src = (d & 1) ? a1 : a0;
            const float val = dl == 0 ? src : __shfl_sync(FULL, src, lane + dl);
            p = fma2s(val, f.t[j], p);
        }
        return p;
    };
    // level 3: one column per lane; column offset d lives in lane + d
    auto hpass3 = [&](float a2) -> float2 {
        if (HAAR) {
            const float nb = __shfl_down_sync(FULL, a2, 1);
            return make_float2(a2, nb);
        }
        float2 p = make_float2(0.f, 0.f);
#pragma unroll
        for (int j = 0; j < F; j++) {
            const int d = j - C;
            const float val = d == 0 ? a2 : __shfl_sync(FULL, a2, lane + d);
            p = fma2s(val, f.t[j], p);
        }
        return p;
    };

    float nrm1 = 0.f, nrm2 = 0.f;      // this lane's share of sum |c|, sum c^2 over the coefficients it stores
    auto acc = [&](float c) {
        nrm1 += fabsf(c);
        nrm2 = fmaf(c, c, nrm2);
    };
    constexpr int FW = HAAR ? 2 : F;
    float4 w1[FW];     // level-1 window of horizontally filtered rows: (lo0, hi0, lo1, hi1); Haar: raw samples
    float2 w2[FW];     // level-2 window: (lo, hi)
    float2 w3[FW];     // level-3 window
#pragma unroll
    for (int j = 0; j < FW; j++) {
        w1[j] = make_float4(0.f, 0.f, 0.f, 0.f);
        w2[j] = make_float2(0.f, 0.f);
        w3[j] = make_float2(0.f, 0.f);
    }

    // vertical steps: combine the window into one output row (Haar: exact 1/2 butterfly, haar.cu:27-35 order
    // differs from the single-level kernel only in that rows are combined after columns)
    auto vstep1 = [&](int krow, bool row_owned, float& a0, float& a1) {
        float h0, h1, v0, v1, d0, d1;
        if (HAAR) {
            const float sx = w1[0].x + w1[1].x, sy = w1[0].y + w1[1].y, sz = w1[0].z + w1[1].z, sw = w1[0].w + w1[1].w;
            const float dx = w1[0].x - w1[1].x, dy = w1[0].y - w1[1].y, dz = w1[0].z - w1[1].z, dw = w1[0].w - w1[1].w;
            a0 = 0.5f * (sx + sy); a1 = 0.5f * (sz + sw);
            v0 = 0.5f * (sx - sy); v1 = 0.5f * (sz - sw);
            h0 = 0.5f * (dx + dy); h1 = 0.5f * (dz + dw);
            d0 = 0.5f * (dx - dy); d1 = 0.5f * (dz - dw);
        } else {
            float2 ah0 = make_float2(0.f, 0.f), vd0 = ah0, ah1 = ah0, vd1 = ah0;      // (a, h) and (v, d) of both columns
#pragma unroll
            for (int j = 0; j < F; j++) {
                ah0 = fma2s(w1[j].x, f.t[j], ah0);
                vd0 = fma2s(w1[j].y, f.t[j], vd0);
                ah1 = fma2s(w1[j].z, f.t[j], ah1);
                vd1 = fma2s(w1[j].w, f.t[j], vd1);
            }
            a0 = ah0.x; h0 = ah0.y; v0 = vd0.x; d0 = vd0.y;
            a1 = ah1.x; h1 = ah1.y; v1 = vd1.x; d1 = vd1.y;
        }
        if (row_owned && own1) {
            char* r = q1 + (long long)krow * (long long)W1b;
            stg2(reinterpret_cast<float*>(r), h0, h1);
            stg2(reinterpret_cast<float*>(r + a.dV[0]), v0, v1);
            stg2(reinterpret_cast<float*>(r + a.dD[0]), d0, d1);
            if (NRM) { acc(h0); acc(h1); acc(v0); acc(v1); acc(d0); acc(d1); }
        }
#pragma unroll
        for (int j = 0; j < FW - 2; j++) w1[j] = w1[j + 2];
    };
    auto vstep23 = [&](float2* w, float& av, float& hv, float& vv, float& dv) {
        if (HAAR) {
            const float sp = w[0].x + w[1].x, sq = w[0].y + w[1].y, dp = w[0].x - w[1].x, dq = w[0].y - w[1].y;
            av = 0.5f * (sp + sq); vv = 0.5f * (sp - sq);
            hv = 0.5f * (dp + dq); dv = 0.5f * (dp - dq);
        } else {
            float2 ah = make_float2(0.f, 0.f), vd = ah;
#pragma unroll
            for (int j = 0; j < F; j++) {
                ah = fma2s(w[j].x, f.t[j], ah);
                vd = fma2s(w[j].y, f.t[j], vd);
            }
            av = ah.x; hv = ah.y; vv = vd.x; dv = vd.y;
        }
#pragma unroll
        for (int j = 0; j < FW - 2; j++) w[j] = w[j + 2];
    };

    // ---- warm-up: the F-2 input rows that precede the first full iteration, then (for F > 2) enough
    // level-1 rows to fill the level-2 window.  Input rows are numbered so that level-1 row k uses
    // rows 2k - C .. 2k - C + F - 1, level-2 row m uses level-1 rows 2m - C .., and so on.
    // Steady state, iteration n (level-3 row n): consumes input rows r0(n) .. r0(n)+7 with
    //   r0(n) = 8n + 7*(F/2) - 7*C ... derived below from the newest rows each window needs.
    // newest level-2 rows for level-3 row n : 2n - C + F - 2, 2n - C + F - 1           =: m_a, m_b
    // newest level-1 rows for level-2 row m : 2m - C + F - 2, 2m - C + F - 1
    // newest input   rows for level-1 row k : 2k - C + F - 2, 2k - C + F - 1
    constexpr int E = F - 1 - C;                 // newest row offset: level row k needs source rows up to 2k + E
    // level-1 rows consumed by iteration n: k = 4n + 3E - 2 - 1 ... explicit list below
    // input rows of iteration n start at rbase(n); PREFETCH: the 8 rows of iteration n+1 are requested
    // before iteration n is computed (template parameter PF), so a warp always has loads in flight.
    auto rbase_of = [&](int n) { return 2 * (2 * (2 * n + E - 1) + E - 1) + E - 1; };
    // PF: the strip's HB + 128 + HB columns of a row are one contiguous piece of memory, or two when the
    // strip straddles the periodic wrap (first / last strip); all pieces are multiples of 16 bytes
    const int s0 = wrap1_per(X0 - HB, Nc);
    const int len0 = min(128 + 2 * HB, Nc - s0);
    const unsigned eoff = lane == 0 ? (HB - C) * 4 : (lane == 31 ? (HB + 128) * 4 : 0);   // other lanes: one broadcast word
    auto issue_rows = [&](int n, unsigned stage) {           // lanes 0..7 request one row each of iteration n
        if (PF == 2) {                                       // cp.async: every lane copies exactly the bytes it will read
            const int rb = rbase_of(n);
            const bool inside = rb >= 0 && rb + 8 <= Nr;
            const unsigned dst = ring0 + stage * (8 * ROWB);
#pragma unroll
            for (int i = 0; i < 8; i++) {
                const float* src = in + (long long)(inside ? rb + i : wrap1_per(rb + i, Nr)) * Nc;
                cp_async16(dst + i * ROWB + HB * 4 + 16 * lane, src + xcol);
                if (C > 0 && (lane == 0 || lane == 31)) {
#pragma unroll
                    for (int c = 0; c < C; c++) cp_async4(dst + i * ROWB + eoff + 4 * c, src + ecol[c]);
                }
            }
        } else if (lane < 8) {
            const float* src = in + (long long)wrap1_per(rbase_of(n) + lane, Nr) * Nc;
            const unsigned dst = ring0 + (stage * 8 + lane) * ROWB;
            const unsigned mb = mbar0 + 8 * stage;
            mbar_expect_tx(mb, ROWB);
            bulk_g2s(dst, src + s0, len0 * 4, mb);
            if (len0 < 128 + 2 * HB) bulk_g2s(dst + len0 * 4, src, ROWB - len0 * 4, mb);
        }
    };
    // PF == 0 (default): all eight input rows of an iteration are requested at its top.
    // PF == 3: the loads software-pipelined IN PLACE -- the registers of input rows 0-3 are re-loaded with the next
    // iteration's rows as soon as level-1 rows 0, 1 have consumed them, rows 4-7 after level-1 rows 2, 3, so a warp
    // always has four rows (2 KB) in flight while it computes, at no register cost (114 vs 108 registers, same 16 warps
    // per SM).  Built because the ncu profile shows long-scoreboard stalls (6.7 per issue, issue slots 39 % busy);
    // measured on B200 (profiles/r02_notes.md): 4096^2 forward 0.0266 vs 0.0275 ms, but 8192^2 forward 0.118 vs
    // 0.096 ms back to back and 0.200 vs 0.198 ms per forward + inverse -- more requests in flight do not help a kernel
    // that already runs at 85 % of the copy bandwidth, they disturb the DRAM access order.  Kept for the A/B only.
    float4 v[8];
    float e[8][C > 0 ? C : 1];
    auto load_rows4 = [&](int rb, int h) {                   // rows rb + 4h .. rb + 4h + 3 -> v[4h ..], e[4h ..]
        if (rb >= 0 && rb + 8 <= Nr) {
            const float* p = in + (long long)(rb + 4 * h) * Nc;
#pragma unroll
            for (int i = 0; i < 4; i++, p += Nc) {
                v[4 * h + i] = ldg4(p + xcol);
#pragma unroll
                for (int c = 0; c < C; c++) e[4 * h + i][c] = __ldg(p + ecol[c]);
            }
        } else {
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const float* p = in + (long long)wrap1_per(rb + 4 * h + i, Nr) * Nc;
                v[4 * h + i] = ldg4(p + xcol);
#pragma unroll
                for (int c = 0; c < C; c++) e[4 * h + i][c] = __ldg(p + ecol[c]);
            }
        }
    };
    auto iteration = [&](int n, bool store_ok) {
        // level-2 rows produced here: m0 = 2n + E - 1, m1 = 2n + E ; level-1 rows: k = 2*m0 + E - 1 .. 2*m1 + E
        const int m0 = 2 * n + E - 1;
        const int kbase = 2 * m0 + E - 1;                 // four level-1 rows kbase .. kbase+3
        const int rb_next = rbase_of(n + 1);
        const bool more = n + 1 < n1;
        if (PF == 1 || PF == 2) {
            const unsigned stage = uses % kStages, parity = (uses / kStages) & 1;
            const unsigned base = ring0 + stage * (8 * ROWB);
            if (PF == 1) mbar_wait(mbar0 + 8 * stage, parity);
            else cp_async_wait<kStages - 1>();                // this lane's own copies of iteration n have landed
#pragma unroll
            for (int i = 0; i < 8; i++) {
                v[i] = lds4(base + i * ROWB + HB * 4 + 16 * lane);
#pragma unroll
                for (int c = 0; c < C; c++) e[i][c] = lds1(base + i * ROWB + eoff + 4 * c);
            }
            if (PF == 1) __syncwarp();                        // every lane has its copy: the stage may be refilled
            if (n + kStages < n1) issue_rows(n + kStages, stage);
            if (PF == 2) cp_async_commit();                   // (possibly empty) group: keeps the group count uniform
            uses++;
        } else if (PF == 0) {
            load_rows8(rbase_of(n), v, e);
        }
#pragma unroll
        for (int t = 0; t < 4; t++) {
            w1[FW - 2] = hpass1(v[2 * t], e[2 * t]);
            w1[FW - 1] = hpass1(v[2 * t + 1], e[2 * t + 1]);
            if (PF == 3 && (t & 1) && more) load_rows4(rb_next, t >> 1);      // rows 4(t>>1) .. +3 are dead now
            const int k = kbase + t;
            float a0, a1;
            vstep1(k, store_ok && k >= 4 * n0 && k < 4 * n1, a0, a1);
            w2[FW - 2 + (t & 1)] = hpass2(a0, a1);
            if (t & 1) {
                const int m = m0 + (t >> 1);
                float a2, h2, v2, d2;
                vstep23(w2, a2, h2, v2, d2);
                if (store_ok && own2 && m >= 2 * n0 && m < 2 * n1) {
                    char* r = q2 + (long long)m * (long long)W2b;
                    *reinterpret_cast<float*>(r) = h2;
                    *reinterpret_cast<float*>(r + a.dV[1]) = v2;
                    *reinterpret_cast<float*>(r + a.dD[1]) = d2;
                    if (NRM) { acc(h2); acc(v2); acc(d2); }
                }
                w3[FW - 2 + (t >> 1)] = hpass3(a2);
            }
        }
        float a3, h3, v3, d3;
        vstep23(w3, a3, h3, v3, d3);
        if (store_ok && own3 && n >= n0 && n < n1) {
            char* r = q3 + (long long)n * (long long)W3b;
            *reinterpret_cast<float*>(r + a.dA3) = a3;
            *reinterpret_cast<float*>(r) = h3;
            *reinterpret_cast<float*>(r + a.dV[2]) = v3;
            *reinterpret_cast<float*>(r + a.dD[2]) = d3;
            if (NRM) {
                acc(h3); acc(v3); acc(d3);
                if (a.count_a3) acc(a3);
            }
        }
    };

    // Warm-up: J extra iterations fill the three sliding windows; whatever they compute from a still
    // incomplete window belongs to rows this task does not own, and the ownership tests inside
    // iteration() keep it from being stored.  J = ceil((3E + 4C - 3) / 4): 0 (haar), 2 (F=4), 4 (F=6), 6 (F=8).
    constexpr int J = HAAR ? 0 : (3 * E + 4 * C - 3 + 3) / 4;
    if (PF == 1 || PF == 2) {
#pragma unroll
        for (int s = 0; s < kStages; s++) {
            if (n0 - J + s < n1) issue_rows(n0 - J + s, (uses + s) % kStages);
            if (PF == 2) cp_async_commit();
        }
    }
    if (PF == 3) {                                            // prologue of the in-place pipeline
        load_rows4(rbase_of(n0 - J), 0);
        load_rows4(rbase_of(n0 - J), 1);
    }
    for (int n = n0 - J; n < n1; n++) iteration(n, true);
    if (NRM) {             // fused norm reduction: warp shuffle, one pair of plain stores per task (no atomics, no memset)
        double d1 = (double)nrm1, d2 = (double)nrm2;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            d1 += __shfl_xor_sync(FULL, d1, o);
            d2 += __shfl_xor_sync(FULL, d2, o);
        }
        if (lane == 0) {
            a.partials[2 * (size_t)t_] = d1;
            a.partials[2 * (size_t)t_ + 1] = d2;
        }
    }
  }
}

inline int cdiv(int a, int b) { return (a + b - 1) / b; }

// Task height (level-3 rows per task): the tasks are long, so a partially filled last wave costs a lot
// (measured: 1.26 waves -> +15 %).  Pick the height that makes the task count an integer number k of
// full waves of resident warps, with the smallest k that keeps a task <= 48 level-3 rows (warm-up
// overhead ~2/T3) -- e.g. forward 8192^2: 74 strips x 32 bands = 2368 tasks = exactly one wave.
int pick_t3(int R3, int strips, int batch, int slots) {
    const long long per_band = (long long)strips * batch;
    for (int k = 1; k <= 64; k++) {
        const long long bands_max = (long long)k * slots / per_band;
        if (bands_max < 1) continue;
        int t3 = (int)((R3 + bands_max - 1) / bands_max);
        if (t3 < 4) t3 = 4;
        if (t3 <= 48) return t3;
    }
    return 16;
}

struct NormSink {     // where the launcher reports how many per-task partial sums were written (travels with the call:
    int cap;          // one plan per host thread may be launching at the same time)
    int* ntasks_out;
};

template <int F, bool HAAR, int MINB, int PF, bool NRM>
int launch_fwd3n(Fwd3Args a, int batch, const PwtFilters& f, PwtTaskQueue* q, const NormSink& sink, cudaStream_t st) {
    a.n3 = HAAR ? 16 : Geo<F>::N3;
    const int W3 = a.Nc / 8, R3 = a.Nr / 8;
    static PwtKernelOnce once;
    const int per_sm = pwt_kernel_once(once, k_fwd3<F, HAAR, MINB, PF, NRM>, 32 * kWarps, 0, 0);
    if (!per_sm) return 0;
    const int resident = pwt_sm_count() * per_sm;
    // task height: as tall as possible (less warm-up) while every resident warp still gets a task
    a.T3 = pwt_tuning().fused_t3;
    // Haar has no halo, hence no warm-up rows: short tasks cost nothing and balance best through the dynamic queue
    // (A/B at 8192^2: T3 = 4 -> 0.0891 ms, 8 -> 0.0905, wave-exact 28 -> 0.1027)
    if (a.T3 <= 0) a.T3 = HAAR ? 4 : pick_t3(R3, cdiv(W3, a.n3), batch, resident * kWarps);
    a.ntasks = cdiv(W3, a.n3) * cdiv(R3, a.T3);
    a.batch = batch;
    long long total = (long long)a.ntasks * batch;
    if (NRM && total > sink.cap) {                     // never write past the plan's buffer
        a.partials = nullptr;
        if constexpr (NRM) return launch_fwd3n<F, HAAR, MINB, PF, false>(a, batch, f, q, sink, st);
    }
    if (sink.ntasks_out) *sink.ntasks_out = NRM ? (int)total : 0;
    int grid = (int)(total < (long long)resident * kWarps ? (total + kWarps - 1) / kWarps : resident);
    a.counter = q->counter;
    a.base = q->base;
    q->base += (unsigned)total + (unsigned)grid * kWarps;      // every warp makes exactly one failing pull
    TapsLH taps;
    for (int j = 0; j < 8; j++) taps.t[j] = (!HAAR && j < F) ? make_float2(f.L[F - 1 - j], f.H[F - 1 - j]) : make_float2(0.f, 0.f);
    // programmatic dependent launch (the kernel triggers when its last wave of tasks starts): 3-level fwd+inv db2
    // 0.0226 -> 0.0178 ms at 512^2, 0.0249 -> 0.0216 at 2048^2, 0.0612 -> 0.0567 at 4096^2, 0.198 -> 0.196 at 8192^2.
    // Not for Haar: 4096^2 got slower (0.0502 -> 0.0594 ms) although 2048^2 and 8192^2 gained.
    if (!HAAR && (pwt_tuning().fused_pdl & 1)) pwt_launch_pdl(k_fwd3<F, HAAR, MINB, PF, NRM>, dim3(grid), 32 * kWarps, 0, st, a, taps);
    else k_fwd3<F, HAAR, MINB, PF, NRM><<<grid, 32 * kWarps, 0, st>>>(a, taps);
    return 1;
}

template <int F, bool HAAR, int MINB, int PF>
int launch_fwd3(Fwd3Args a, int batch, const PwtFilters& f, PwtTaskQueue* q, const NormSink& sink, cudaStream_t st) {
    // the norm accumulation is compiled in only when the plan asked for it (~2 % of the pass)
    return a.partials ? launch_fwd3n<F, HAAR, MINB, PF, true>(a, batch, f, q, sink, st)
                      : launch_fwd3n<F, HAAR, MINB, PF, false>(a, batch, f, q, sink, st);
}

}  // namespace

// Levels 1..3 of the forward transform in one launch.  band pointers: H/V/D of levels 1, 2, 3; A3.
int pwt_fused_fwd3_max_tasks(int batch, int Nr, int Nc) {
    // upper bound of the task count of one launch (smallest strips of 11 columns, task height >= 4 rows)
    const long long W3 = Nc / 8, R3 = Nr / 8;
    return (int)(((W3 + 10) / 11) * ((R3 + 3) / 4) * batch);
}

int pwt_fused_dwt_fwd3(const float* in, float* A3, float* const* H, float* const* V, float* const* D,
                       int batch, int Nr, int Nc, const PwtFilters& f, bool haar, PwtTaskQueue* q,
                       double* partials, int partials_cap, int count_a3, int* ntasks_out, cudaStream_t st) {
    const int F = haar ? 2 : f.hlen;
    if (pwt_tuning().no_fused) return 0;
    if (F > 8 || (F & 1) || Nr % 8 != 0 || Nc % 8 != 0 || Nc < 512 || Nr < 64 || batch > 65535) return 0;
    if (((uintptr_t)in & 15) != 0) return 0;
    Fwd3Args a;
    a.in = in;
    a.A3 = A3;
    for (int l = 0; l < 3; l++) {
        a.H[l] = H[l];
        a.V[l] = V[l];
        a.D[l] = D[l];
        a.bs[l] = (long long)(Nr >> (l + 1)) * (Nc >> (l + 1));
        a.dV[l] = (long long)((const char*)V[l] - (const char*)H[l]);
        a.dD[l] = (long long)((const char*)D[l] - (const char*)H[l]);
        if (((uintptr_t)H[l] | (uintptr_t)V[l] | (uintptr_t)D[l]) & 7) return 0;
    }
    a.dA3 = (long long)((const char*)A3 - (const char*)H[2]);
    a.Nr = Nr;
    a.Nc = Nc;
    a.in_bs = (long long)Nr * Nc;
    a.partials = partials;
    a.count_a3 = count_a3;
    if (ntasks_out) *ntasks_out = 0;
    const NormSink sink = {partials ? partials_cap : 0, ntasks_out};
    const int variant = pwt_tuning().fused_variant;
    if (haar) {
        if (variant == 1) return launch_fwd3<2, true, 6, 0>(a, batch, f, q, sink, st);
        if (variant == 8) return launch_fwd3<2, true, 4, 3>(a, batch, f, q, sink, st);
        if (variant == 2) return launch_fwd3<2, true, 5, 0>(a, batch, f, q, sink, st);
        return launch_fwd3<2, true, 4, 0>(a, batch, f, q, sink, st);
    }
    switch (F) {
        case 4:
            if (variant == 2) return launch_fwd3<4, false, 5, 0>(a, batch, f, q, sink, st);
            if (variant == 3) return launch_fwd3<4, false, 6, 0>(a, batch, f, q, sink, st);
            if (variant == 4) return launch_fwd3<4, false, 4, 1>(a, batch, f, q, sink, st);
            if (variant == 5) return launch_fwd3<4, false, 4, 2>(a, batch, f, q, sink, st);
            if (variant == 6) return launch_fwd3<4, false, 3, 0>(a, batch, f, q, sink, st);
            if (variant == 7) return launch_fwd3<4, false, 2, 0>(a, batch, f, q, sink, st);
            if (variant == 8) return launch_fwd3<4, false, 4, 3>(a, batch, f, q, sink, st);
            return launch_fwd3<4, false, 4, 0>(a, batch, f, q, sink, st);
        case 6: return launch_fwd3<6, false, 3, 0>(a, batch, f, q, sink, st);
        case 8: return launch_fwd3<8, false, 3, 0>(a, batch, f, q, sink, st);
        default: return 0;
    }
}

// =========================================================================================
// inverse cascade: levels 3 -> 2 -> 1 -> image in one launch
// =========================================================================================
namespace {

struct Inv3Args {
    int plain;                     // host only: launch without the programmatic-dependent-launch attribute
    const float* A3;
    const float* H[3];     // index 0 = level 1 (finest)
    const float* V[3];
    const float* D[3];
    float* out;
    int Nr, Nc;            // image size
    int n3;                // level-3 columns owned by one warp
    int T3;                // level-3 rows (= 8 image rows each) per task
    long long out_bs;
    long long bs[3];
    unsigned* counter;
    unsigned base;
    int ntasks, batch;
    // deferred coefficient operator applied to the coefficients as they are loaded (0 extra bytes):
    int thr_op;            // -1 none, PWT_OP_SOFT, PWT_OP_HARD
    int thr_app;           // also threshold A3
    float thr_beta[3];     // per level (index 0 = level 1)
    float thr_beta_app;
};

template <int THR>
__device__ __forceinline__ float thr1(float v, float beta) {
    // common.cu:19 (soft) / common.cu:63 (hard, strict >)
    if (THR == 1) return copysignf(fmaxf(fabsf(v) - beta, 0.0f), v);
    return (fabsf(v) - beta > 0.0f) ? v : 0.0f * v;
}
template <int THR>
__device__ __forceinline__ float2 thr2(float2 v, float b) { return make_float2(thr1<THR>(v.x, b), thr1<THR>(v.y, b)); }
template <int THR>
__device__ __forceinline__ float4 thr4(float4 v, float b) {
    return make_float4(thr1<THR>(v.x, b), thr1<THR>(v.y, b), thr1<THR>(v.z, b), thr1<THR>(v.w, b));
}

// one horizontally synthesised band row held CW band columns per lane: u1 = syn_x(A, V), u2 = syn_x(H, D)
template <int CW>
struct URow {
    float u1[2 * CW], u2[2 * CW];
};

template <int F, bool HAAR, int MINB, int THR>   // THR: 0 none, 1 soft, 2 hard threshold applied on load
__global__ void __launch_bounds__(32 * kWarps, MINB)
k_inv3(const __grid_constant__ Inv3Args a, const __grid_constant__ PwtFilters f) {
    constexpr int P = F / 2 - 1, HALF = F / 2;
    constexpr int S0 = P >> 1, E0 = P & 1;
    constexpr int S1 = (P + 1) >> 1, E1 = (P + 1) & 1;
    constexpr int WIN = HALF + (S1 - S0);
    constexpr int HW = S1;                               // 0 (haar) or 1 (F = 4, 6)
    static_assert(HW <= 1, "fused inverse supports a horizontal reach of one band sample");
    pwt_pdl_wait();
    constexpr int OWN0 = (6 * S1 + 7) & ~7;              // image columns given up on each side of the 256 loaded
    constexpr int DT0 = S1 ? 2 : 0, DT1 = S1 ? 1 : -1;   // iterations before n0 / after n1-1
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int Nr = a.Nr, Nc = a.Nc;
    const int W1 = Nc >> 1, W2 = Nc >> 2, W3 = Nc >> 3;
    const int R1 = Nr >> 1, R2 = Nr >> 2, R3 = Nr >> 3;
    const int strips = (W3 + a.n3 - 1) / a.n3;
    const bool first_lane = lane == 0, last_lane = lane == 31;
  for (;;) {
    unsigned t_ = 0;
    if (lane == 0) t_ = atomicAdd(a.counter, 1u) - a.base;
    t_ = __shfl_sync(FULL, t_, 0);
    // last wave of tasks (or none left): the next launch may start being scheduled into the slots that free up
    if (t_ + gridDim.x * kWarps >= (unsigned)a.ntasks * (unsigned)a.batch) pwt_pdl_trigger();
    if (t_ >= (unsigned)a.ntasks * (unsigned)a.batch) return;
    const int img = t_ / a.ntasks, task = t_ - img * a.ntasks;
    const int strip = task % strips, band = task / strips;
    const int n0 = band * a.T3;
    const int n1 = min(n0 + a.T3, R3);
    const int c3 = strip * a.n3;
    const int n3e = min(a.n3, W3 - c3);
    const int K3 = c3 - OWN0 / 8;                        // level-3 column of lane 0 (may be -1: wraps)

    const float* A3 = a.A3 + img * a.bs[2];
    const float* H3 = a.H[2] + img * a.bs[2]; const float* V3 = a.V[2] + img * a.bs[2]; const float* D3 = a.D[2] + img * a.bs[2];
    const float* H2 = a.H[1] + img * a.bs[1]; const float* V2 = a.V[1] + img * a.bs[1]; const float* D2 = a.D[1] + img * a.bs[1];
    const float* H1 = a.H[0] + img * a.bs[0]; const float* V1 = a.V[0] + img * a.bs[0]; const float* D1 = a.D[0] + img * a.bs[0];
    float* out = a.out + img * a.out_bs;

    const int x3 = wrap1_per(K3 + lane, W3);
    const int x2 = wrap1_per(2 * (K3 + lane), W2);
    const int x1 = wrap1_per(4 * (K3 + lane), W1);
    const int e3col = first_lane ? wrap1_per(K3 - 1, W3) : (last_lane ? wrap1_per(K3 + 32, W3) : x3);
    const int px = 8 * (K3 + lane);                      // first image column of this lane
    const bool own = px >= 8 * c3 && px < 8 * (c3 + n3e);

    // ---- horizontal synthesis of a band row held CW columns per lane (same arithmetic as k_inv_reg) ----
    // neighbours: lane-1's last column / lane+1's first column by shuffle; `edge` supplies them for the strip edge
    auto hsyn = [&](auto cw, const float* ba, const float* bh, const float* bv, const float* bd, const float* edge,
                    float* u1, float* u2) {
        constexpr int CW = decltype(cw)::value;
        if (HAAR) {
            // haar.cu:41-58 order: (a + h) and (v + d) first; rows are split in vsyn
#pragma unroll
            for (int c = 0; c < CW; c++) {
                u1[2 * c] = ba[c] + bh[c];       // ac
                u1[2 * c + 1] = bv[c] + bd[c];   // bd
                u2[2 * c] = ba[c] - bh[c];       // am
                u2[2 * c + 1] = bv[c] - bd[c];   // bm
            }
            return;
        }
        float xa[CW + 2 * HW], xh[CW + 2 * HW], xv[CW + 2 * HW], xd[CW + 2 * HW];
        const float* src[4] = {ba, bh, bv, bd};
        float* dst[4] = {xa, xh, xv, xd};
#pragma unroll
        for (int k = 0; k < 4; k++) {
#pragma unroll
            for (int c = 0; c < CW; c++) dst[k][HW + c] = src[k][c];
            if (HW > 0) {
                const float l = __shfl_up_sync(FULL, src[k][CW - 1], 1);
                const float r = __shfl_down_sync(FULL, src[k][0], 1);
                dst[k][0] = (edge && first_lane) ? edge[k] : l;
                dst[k][HW + CW] = (edge && last_lane) ? edge[k] : r;
            }
        }
#pragma unroll
        for (int c = 0; c < CW; c++) {
            float e1 = 0.f, o1 = 0.f, e2 = 0.f, o2 = 0.f;
#pragma unroll
            for (int jj = 0; jj < HALF; jj++) {
                const float le = f.IL[2 * jj + E0], he = f.IH[2 * jj + E0];
                const float lo = f.IL[2 * jj + E1], ho = f.IH[2 * jj + E1];
                e1 = fmaf(xa[HW + c + S0 - jj], le, e1);
                e1 = fmaf(xv[HW + c + S0 - jj], he, e1);
                o1 = fmaf(xa[HW + c + S1 - jj], lo, o1);
                o1 = fmaf(xv[HW + c + S1 - jj], ho, o1);
                e2 = fmaf(xh[HW + c + S0 - jj], le, e2);
                e2 = fmaf(xd[HW + c + S0 - jj], he, e2);
                o2 = fmaf(xh[HW + c + S1 - jj], lo, o2);
                o2 = fmaf(xd[HW + c + S1 - jj], ho, o2);
            }
            u1[2 * c] = e1;
            u1[2 * c + 1] = o1;
            u2[2 * c] = e2;
            u2[2 * c + 1] = o2;
        }
    };
    // ---- vertical synthesis: window of WIN rows -> output rows 2q (ev) and 2q+1 (od), then shift ----
    auto vsyn = [&](auto nv, auto* w, float* ev, float* od) {
        constexpr int NV = decltype(nv)::value;
        if (HAAR) {
#pragma unroll
            for (int c = 0; c < NV; c += 2) {
                const float ac = w[0].u1[c], bd = w[0].u1[c + 1], am = w[0].u2[c], bm = w[0].u2[c + 1];
                ev[c] = 0.5f * (ac + bd);
                ev[c + 1] = 0.5f * (ac - bd);
                od[c] = 0.5f * (am + bm);
                od[c + 1] = 0.5f * (am - bm);
            }
            return;
        }
#pragma unroll
        for (int c = 0; c < NV; c++) {
            float s0 = 0.f, s1 = 0.f;
#pragma unroll
            for (int jj = 0; jj < HALF; jj++) {
                const int je = HALF - 1 - jj, jo = HALF - 1 - jj + (S1 - S0);
                s0 = fmaf(w[je].u1[c], f.IL[2 * jj + E0], s0);
                s0 = fmaf(w[je].u2[c], f.IH[2 * jj + E0], s0);
                s1 = fmaf(w[jo].u1[c], f.IL[2 * jj + E1], s1);
                s1 = fmaf(w[jo].u2[c], f.IH[2 * jj + E1], s1);
            }
            ev[c] = s0;
            od[c] = s1;
        }
#pragma unroll
        for (int j = 0; j < WIN - 1; j++) w[j] = w[j + 1];
    };

    URow<1> w3[WIN];
    URow<2> w2[WIN];
    URow<4> w1[WIN];

    for (int t = n0 - DT0; t <= n1 + DT1; t++) {
        const bool do2 = HAAR || t >= n0;            // level-2 section (and level-3 emission) needed
        const bool do1 = HAAR || t >= n0 + 1;        // level-1 section needed
        // ---- loads of this iteration, issued up front ----
        float b3[4], e3[4];
        {
            const long long ro = (long long)wrap1_per(t, R3) * W3;
            b3[0] = __ldg(A3 + ro + x3); b3[1] = __ldg(H3 + ro + x3); b3[2] = __ldg(V3 + ro + x3); b3[3] = __ldg(D3 + ro + x3);
            if (HW > 0) {
                e3[0] = __ldg(A3 + ro + e3col); e3[1] = __ldg(H3 + ro + e3col);
                e3[2] = __ldg(V3 + ro + e3col); e3[3] = __ldg(D3 + ro + e3col);
            }
            if (THR) {
#pragma unroll
                for (int k = 1; k < 4; k++) {
                    b3[k] = thr1<THR>(b3[k], a.thr_beta[2]);
                    if (HW > 0) e3[k] = thr1<THR>(e3[k], a.thr_beta[2]);
                }
                if (a.thr_app) {
                    b3[0] = thr1<THR>(b3[0], a.thr_beta_app);
                    if (HW > 0) e3[0] = thr1<THR>(e3[0], a.thr_beta_app);
                }
            }
        }
        float2 h2[2], v2[2], d2[2];
        float4 h1[4], v1[4], d1[4];
        if (do2) {
            const long long ro = (long long)wrap1_per(2 * (t - S1), R2) * W2 + x2;
#pragma unroll
            for (int i = 0; i < 2; i++) {
                h2[i] = __ldg(reinterpret_cast<const float2*>(H2 + ro + (long long)i * W2));
                v2[i] = __ldg(reinterpret_cast<const float2*>(V2 + ro + (long long)i * W2));
                d2[i] = __ldg(reinterpret_cast<const float2*>(D2 + ro + (long long)i * W2));
                if (THR) {
                    h2[i] = thr2<THR>(h2[i], a.thr_beta[1]);
                    v2[i] = thr2<THR>(v2[i], a.thr_beta[1]);
                    d2[i] = thr2<THR>(d2[i], a.thr_beta[1]);
                }
            }
        }
        if (do1) {
#pragma unroll
            for (int pr = 0; pr < 2; pr++) {
                const long long ro = (long long)wrap1_per(4 * t - 6 * S1 + 2 * pr, R1) * W1 + x1;
#pragma unroll
                for (int i = 0; i < 2; i++) {
                    h1[2 * pr + i] = ldg4(H1 + ro + (long long)i * W1);
                    v1[2 * pr + i] = ldg4(V1 + ro + (long long)i * W1);
                    d1[2 * pr + i] = ldg4(D1 + ro + (long long)i * W1);
                    if (THR) {
                        h1[2 * pr + i] = thr4<THR>(h1[2 * pr + i], a.thr_beta[0]);
                        v1[2 * pr + i] = thr4<THR>(v1[2 * pr + i], a.thr_beta[0]);
                        d1[2 * pr + i] = thr4<THR>(d1[2 * pr + i], a.thr_beta[0]);
                    }
                }
            }
        }
        // ---- level 3: push row t ----
        hsyn(std::integral_constant<int, 1>{}, &b3[0], &b3[1], &b3[2], &b3[3], HW > 0 ? e3 : nullptr,
             w3[WIN - 1].u1, w3[WIN - 1].u2);
        if (!do2) {
#pragma unroll
            for (int j = 0; j < WIN - 1; j++) w3[j] = w3[j + 1];
            continue;
        }
        float a2[2][2];                                // [row 2q3 / 2q3+1][2 columns]
        vsyn(std::integral_constant<int, 2>{}, w3, a2[0], a2[1]);
        // ---- level 2: two A2 rows ----
#pragma unroll
        for (int i2 = 0; i2 < 2; i2++) {
            const float bh[2] = {h2[i2].x, h2[i2].y}, bv[2] = {v2[i2].x, v2[i2].y}, bd[2] = {d2[i2].x, d2[i2].y};
            hsyn(std::integral_constant<int, 2>{}, a2[i2], bh, bv, bd, nullptr, w2[WIN - 1].u1, w2[WIN - 1].u2);
            float a1[2][4];
            vsyn(std::integral_constant<int, 4>{}, w2, a1[0], a1[1]);
            if (!do1) continue;
            // ---- level 1: two A1 rows -> four image rows ----
#pragma unroll
            for (int i1 = 0; i1 < 2; i1++) {
                const int r = 2 * i2 + i1;               // index of the A1 row inside this iteration
                const float bh1[4] = {h1[r].x, h1[r].y, h1[r].z, h1[r].w};
                const float bv1[4] = {v1[r].x, v1[r].y, v1[r].z, v1[r].w};
                const float bd1[4] = {d1[r].x, d1[r].y, d1[r].z, d1[r].w};
                hsyn(std::integral_constant<int, 4>{}, a1[i1], bh1, bv1, bd1, nullptr, w1[WIN - 1].u1, w1[WIN - 1].u2);
                float ev[8], od[8];
                vsyn(std::integral_constant<int, 8>{}, w1, ev, od);
                const int y = 8 * t - 14 * S1 + 2 * r;   // image rows y, y+1
                if (own && y >= 8 * n0 && y < 8 * n1) {
                    float* p = out + (long long)y * Nc + px;
                    *reinterpret_cast<float4*>(p) = make_float4(ev[0], ev[1], ev[2], ev[3]);
                    *reinterpret_cast<float4*>(p + 4) = make_float4(ev[4], ev[5], ev[6], ev[7]);
                    *reinterpret_cast<float4*>(p + Nc) = make_float4(od[0], od[1], od[2], od[3]);
                    *reinterpret_cast<float4*>(p + Nc + 4) = make_float4(od[4], od[5], od[6], od[7]);
                }
            }
        }
    }
  }
}

template <int F, bool HAAR, int MINB, int THR>
int launch_inv3t(Inv3Args a, int batch, const PwtFilters& f, PwtTaskQueue* q, cudaStream_t st) {
    constexpr int S1 = (F / 2) >> 1;
    constexpr int OWN0 = (6 * S1 + 7) & ~7;
    a.n3 = (256 - 2 * OWN0) / 8;
    const int W3 = a.Nc / 8, R3 = a.Nr / 8;
    static PwtKernelOnce once;
    const int per_sm = pwt_kernel_once(once, k_inv3<F, HAAR, MINB, THR>, 32 * kWarps, 0, 0);
    if (!per_sm) return 0;
    const int resident = pwt_sm_count() * per_sm;
    a.T3 = pwt_tuning().fused_inv_t3;
    if (a.T3 <= 0) a.T3 = HAAR ? 8 : pick_t3(R3, cdiv(W3, a.n3), batch, resident * kWarps);   // Haar: see the forward
    a.ntasks = cdiv(W3, a.n3) * cdiv(R3, a.T3);
    a.batch = batch;
    const long long total = (long long)a.ntasks * batch;
    const int grid = (int)(total < (long long)resident * kWarps ? (total + kWarps - 1) / kWarps : resident);
    a.counter = q->counter;
    a.base = q->base;
    q->base += (unsigned)total + (unsigned)grid * kWarps;
    if (!HAAR && !a.plain && (pwt_tuning().fused_pdl & 2)) pwt_launch_pdl(k_inv3<F, HAAR, MINB, THR>, dim3(grid), 32 * kWarps, 0, st, a, f);
    else k_inv3<F, HAAR, MINB, THR><<<grid, 32 * kWarps, 0, st>>>(a, f);
    return 1;
}

template <int F, bool HAAR, int MINB>
int launch_inv3(const Inv3Args& a, int batch, const PwtFilters& f, PwtTaskQueue* q, cudaStream_t st) {
    if (a.thr_op == PWT_OP_SOFT) return launch_inv3t<F, HAAR, MINB, 1>(a, batch, f, q, st);
    if (a.thr_op == PWT_OP_HARD) return launch_inv3t<F, HAAR, MINB, 2>(a, batch, f, q, st);
    return launch_inv3t<F, HAAR, MINB, 0>(a, batch, f, q, st);
}

}  // namespace

// Levels 3..1 of the inverse transform in one launch.  H/V/D[0] = level 1 (finest).
int pwt_fused_dwt_inv3(const float* A3, const float* const* H, const float* const* V, const float* const* D,
                       float* out, int batch, int Nr, int Nc, const PwtFilters& f, bool haar, PwtTaskQueue* q,
                       const PwtDeferredOp* op, int plain_launch, cudaStream_t st) {
    const int F = haar ? 2 : f.hlen;
    if (pwt_tuning().no_fused || pwt_tuning().no_fused_inv) return 0;
    if (F > 6 || (F & 1) || Nr % 8 != 0 || Nc % 8 != 0 || Nc < 512 || Nr < 64 || batch > 65535) return 0;
    if (((uintptr_t)out & 15) != 0) return 0;
    Inv3Args a;
    a.plain = plain_launch;
    a.A3 = A3;
    a.out = out;
    for (int l = 0; l < 3; l++) {
        a.H[l] = H[l];
        a.V[l] = V[l];
        a.D[l] = D[l];
        a.bs[l] = (long long)(Nr >> (l + 1)) * (Nc >> (l + 1));
        if (((uintptr_t)H[l] | (uintptr_t)V[l] | (uintptr_t)D[l]) & 15) return 0;
    }
    a.Nr = Nr;
    a.Nc = Nc;
    a.out_bs = (long long)Nr * Nc;
    a.thr_op = -1;
    a.thr_app = 0;
    a.thr_beta[0] = a.thr_beta[1] = a.thr_beta[2] = a.thr_beta_app = 0.f;
    if (op && op->op >= 0) {
        a.thr_op = op->op;
        a.thr_app = op->app;
        a.thr_beta_app = op->beta_app;
        for (int l = 0; l < 3; l++) a.thr_beta[l] = op->beta[l];
    }
    if (haar) return launch_inv3<2, true, 4>(a, batch, f, q, st);
    switch (F) {
        case 4: return launch_inv3<4, false, 3>(a, batch, f, q, st);
        case 6: return launch_inv3<6, false, 2>(a, batch, f, q, st);
        default: return 0;
    }
}
