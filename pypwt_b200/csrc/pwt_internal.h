// Internal declarations shared by the plan/dispatch code and the kernel files.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

#define PWT_MAX_TAPS 40      // same cap as the reference (common.h:15)
#define PWT_MAX_LEVELS 31
#define PWT_MAX_BANDS (3 * PWT_MAX_LEVELS + 1)

// 1D filter bank, passed BY VALUE to kernels (lands in the constant bank as kernel
// parameters, so every instance has its own filters -- fixes reference quirk Q4).
template <typename T>
struct PwtFiltersT {
    T L[PWT_MAX_TAPS];        // analysis low-pass   (dec_lo)
    T H[PWT_MAX_TAPS];        // analysis high-pass  (dec_hi)
    T IL[PWT_MAX_TAPS];       // synthesis low-pass  (rec_lo)
    T IH[PWT_MAX_TAPS];       // synthesis high-pass (rec_hi)
    int hlen;
};
using PwtFilters = PwtFiltersT<float>;
using PwtFilters64 = PwtFiltersT<double>;   // the reference's DOUBLEPRECISION build (filters.h:16-30): taps not rounded to fp32

// Taps packed for the 2-wide FMA of sm_100 (FFMA2: one sample times a pair of taps).  FFMA2 issues at half the
// rate of FFMA with the same FMA-pipe throughput (tools/bench/fma2bench.cu): it halves the issue slots the
// arithmetic needs.  The summation order of every output is the reference's; taps a phase does not use are 0.
struct PwtTapsFwd {           // analysis, reversed and interleaved: t[j] = (L[F-1-j], H[F-1-j])
    float2 t[PWT_MAX_TAPS];
};
struct PwtTapsInv {           // synthesis by window position w (band offset + S1), (even phase, odd phase):
    float2 l[PWT_MAX_TAPS / 2 + 2];   //   l[w] = (IL[2*je+E0], IL[2*jo+E1]),  je = S0+S1-w, jo = 2*S1-w
    float2 h[PWT_MAX_TAPS / 2 + 2];   //   h[w] likewise from IH
};
static inline PwtTapsFwd pwt_pack_taps_fwd(const PwtFilters& f, int F) {
    PwtTapsFwd t;
    for (int j = 0; j < PWT_MAX_TAPS; j++) {
        t.t[j].x = j < F ? f.L[F - 1 - j] : 0.f;
        t.t[j].y = j < F ? f.H[F - 1 - j] : 0.f;
    }
    return t;
}
static inline PwtTapsInv pwt_pack_taps_inv(const PwtFilters& f, int F) {
    const int P = F / 2 - 1, HALF = F / 2, S0 = P >> 1, E0 = P & 1, S1 = (P + 1) >> 1, E1 = (P + 1) & 1;
    PwtTapsInv t;
    for (int w = 0; w < PWT_MAX_TAPS / 2 + 2; w++) {
        const int je = S0 + S1 - w, jo = 2 * S1 - w;
        const bool ue = je >= 0 && je < HALF, uo = jo >= 0 && jo < HALF;
        t.l[w].x = ue ? f.IL[2 * je + E0] : 0.f;
        t.l[w].y = uo ? f.IL[2 * jo + E1] : 0.f;
        t.h[w].x = ue ? f.IH[2 * je + E0] : 0.f;
        t.h[w].y = uo ? f.IH[2 * jo + E1] : 0.f;
    }
    return t;
}

// ---- host-side launcher state: everything below is safe to call from any number of host threads (one plan per
// thread, the Cython wrapper releases the GIL around the transforms) and keeps its caches PER DEVICE ----
#define PWT_MAX_DEVICES 64
#ifdef __cplusplus
#include <mutex>
// Developer A/B knobs, read from the environment ONCE per process (never on the launch path).
struct PwtTuning {
    int no_pdl;            // PWT_NO_PDL=1: plain launches instead of programmatic dependent launch
    int no_fused;          // PWT_NO_FUSED=1: no fused 3-level cascade
    int no_fused_inv;      // PWT_NO_FUSED_INV=1
    int fused_variant;     // PWT_FUSED_VARIANT (occupancy / staging variants of k_fwd3)
    int fused_t3;          // PWT_FUSED_T3: forced task height of k_fwd3 (0 = automatic)
    int fused_inv_t3;      // PWT_FUSED_INV_T3
    int fused_pdl;         // PWT_FUSED_PDL (default 1)
    int no_fused_norms;    // PWT_NO_FUSED_NORMS=1
    int no_defer;          // PWT_NO_DEFER=1: thresholds always applied through memory
    int tile_min_f;        // PWT_TILE_MIN_F (default 22)
    int strip_min_f;       // PWT_STRIP_MIN_F (default 8)
    int ns_direct;         // PWT_NS_DIRECT=1: true F x F stencils even for rank-1 banks
    int l2_persist_mb;     // PWT_L2_PERSIST_MB (default 0)
    int verbose;           // PWT_VERBOSE
    int strip_segs;        // PWT_STRIP_SEGS: forced segment count of the strip kernels (0 = automatic)
    int strip_occ_fwd;     // PWT_STRIP_OCC_FWD / _INV: forced CTAs per SM variant
    int strip_occ_inv;
    int swt_nbuf;          // PWT_SWT_NBUF (default 1)
    int no_strip_swt;      // PWT_NO_STRIP_SWT=1
    int no_fast_swt;       // PWT_NO_FAST_SWT=1
    int swt_tq;            // PWT_SWT_TQ
    int fast_tile_rows;    // PWT_FAST_TILE_ROWS
    int fwd_variant;       // PWT_FWD_VARIANT (-1 = automatic)
    int reg_tile_rows;     // PWT_REG_TILE_ROWS (default 16)
    int use_hints;         // PWT_USE_HINTS
    int reg_fwd_variant;   // PWT_REG_FWD_VARIANT (default 2)
    int no_fold_cs;        // PWT_NO_FOLD_CS=1: cycle-spinning shifts as separate gather passes
    int no_cascade8;       // PWT_NO_CASCADE8=1: no level-fused strip kernels (F >= 8)
    int no_fused1d;        // PWT_NO_FUSED1D=1: batched 1D one launch per level
    int strip_thr_occ3;    // PWT_STRIP_THR_OCC3 (default 1): thresholding strip inverse at 3 CTAs/SM for F = 14
    int tail_strip;        // PWT_TAIL_STRIP (default 1): levels >= 4 of short-filter transforms run the strip kernels
};
const PwtTuning& pwt_tuning();       // pwt_plan.cu
int pwt_sm_count();                  // SM count of the CURRENT device (cached per device; pwt_plan.cu)

// One-time per-device set-up of ONE kernel instantiation: raises the dynamic shared memory limit to `smem_attr`
// and caches the resident CTAs per SM for (threads, smem_occ).  Returns that count (>= 1), or 0 when the set-up
// failed -- nothing is cached then, so the next call retries instead of launching with a limit that was never raised.
struct PwtKernelOnce {
    std::mutex m;
    int per_sm[PWT_MAX_DEVICES] = {};
};
template <typename K>
static inline int pwt_kernel_once(PwtKernelOnce& o, K kernel, int threads, size_t smem_attr, size_t smem_occ) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= PWT_MAX_DEVICES) return 0;
    int v = __atomic_load_n(&o.per_sm[dev], __ATOMIC_ACQUIRE);
    if (v) return v;
    std::lock_guard<std::mutex> g(o.m);
    v = o.per_sm[dev];
    if (v) return v;
    if (smem_attr > 0 && cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_attr) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&v, kernel, threads, smem_occ) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    if (v < 1) v = 1;
    __atomic_store_n(&o.per_sm[dev], v, __ATOMIC_RELEASE);
    return v;
}
#endif

#ifdef __CUDACC__
// Programmatic dependent launch (sm_90+): a kernel launched through pwt_launch_pdl may start being scheduled while
// the previous kernel of the stream drains (after all its CTAs called pwt_pdl_trigger() or exited).  Such a kernel
// MUST call pwt_pdl_wait() before its first global-memory access: the wait returns once the previous grid has
// completed and its writes are visible (immediately for a normally launched grid).  Transforms are chains of
// 2-14 dependent launches, the small levels only a few microseconds long: this hides their launch latency
// (5-level db4 / sym8 / db12 forward + inverse at 8192^2: 0.307 -> 0.287, 0.352 -> 0.328, 0.480 -> 0.454 ms).
__device__ __forceinline__ void pwt_pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pwt_pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;"); }

#include <stdlib.h>
template <typename... KArgs, typename... Args>
static inline void pwt_launch_pdl(void (*kernel)(KArgs...), dim3 grid, int threads, size_t smem, cudaStream_t st, Args... args) {
    const int use = pwt_tuning().no_pdl ? 0 : 1;                    // PWT_NO_PDL=1: plain launches (A/B)
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = dim3(threads, 1, 1);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = use ? 1 : 0;
    cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}
#endif

// Geometry of one 2D plane set processed by a launch (all strides in elements).
struct PwtPlane {
    int nr, nc;               // rows, cols of ONE image of the stack
    long long bstride;        // distance between consecutive images of the stack
};

// element-wise operator codes (kernels_ops.cu)
enum { PWT_OP_SOFT = 0, PWT_OP_HARD = 1, PWT_OP_PROJ = 2, PWT_OP_SCALE = 3 };

struct PwtSeg {               // one contiguous run of coefficients + its parameter
    float* ptr;
    long long n;
    float beta;
    int pad;
};
#define PWT_MAX_SEGS PWT_MAX_BANDS
struct PwtSegTable {
    PwtSeg seg[PWT_MAX_SEGS];
    int nseg;
};

// ---- kernel launchers (each returns the number of kernels it launched) ----------------
// kernels_generic.cu : tiled kernels valid for every size / filter length
int pwt_launch_dwt_fwd2d(const float* in, float* A, float* Hb, float* V, float* D, int batch, int Nr,
                         int Nc, long long in_bs, long long out_bs, const PwtFilters& f, bool haar,
                         cudaStream_t st);
int pwt_launch_dwt_inv2d(const float* A, const float* Hb, const float* V, const float* D, float* out,
                         int batch, int nr, int nc, int Nr_out, int Nc_out, long long in_bs,
                         long long out_bs, const PwtFilters& f, bool haar, cudaStream_t st);
// batched 1D over `rows` rows (rows = batch*Nr)
int pwt_launch_dwt_fwd1d(const float* in, float* A, float* D, int rows, int Nc, const PwtFilters& f,
                         bool haar, cudaStream_t st);
int pwt_launch_dwt_inv1d(const float* A, const float* D, float* out, int rows, int nc, int Nc_out,
                         const PwtFilters& f, bool haar, cudaStream_t st);
// SWT (a trous), level >= 1.  tmp must hold 2*batch*Nr*Nc floats (2D only).
int pwt_launch_swt_fwd2d(const float* in, float* A, float* Hb, float* V, float* D, float* tmp,
                         int batch, int Nr, int Nc, int level, const PwtFilters& f, cudaStream_t st);
int pwt_launch_swt_inv2d(const float* A, const float* Hb, const float* V, const float* D, float* out,
                         float* tmp, int batch, int Nr, int Nc, int level, const PwtFilters& f,
                         cudaStream_t st);
int pwt_launch_swt_fwd1d(const float* in, float* A, float* D, int rows, int Nc, int level,
                         const PwtFilters& f, cudaStream_t st);
int pwt_launch_swt_inv1d(const float* A, const float* D, float* out, int rows, int Nc, int level,
                         const PwtFilters& f, cudaStream_t st);
// the same kernels instantiated for double (pwt_plan64.cu)
int pwt_launch_dwt_fwd2d_f64(const double* in, double* A, double* Hb, double* V, double* D, int batch, int Nr,
                             int Nc, long long in_bs, long long out_bs, const PwtFilters64& f, bool haar, cudaStream_t st);
int pwt_launch_dwt_inv2d_f64(const double* A, const double* Hb, const double* V, const double* D, double* out,
                             int batch, int nr, int nc, int Nr_out, int Nc_out, long long in_bs, long long out_bs,
                             const PwtFilters64& f, bool haar, cudaStream_t st);
int pwt_launch_dwt_fwd1d_f64(const double* in, double* A, double* D, int rows, int Nc, const PwtFilters64& f, bool haar,
                             cudaStream_t st);
int pwt_launch_dwt_inv1d_f64(const double* A, const double* D, double* out, int rows, int nc, int Nc_out,
                             const PwtFilters64& f, bool haar, cudaStream_t st);
int pwt_launch_swt_fwd2d_f64(const double* in, double* A, double* Hb, double* V, double* D, double* tmp, int batch, int Nr,
                             int Nc, int level, const PwtFilters64& f, cudaStream_t st);
int pwt_launch_swt_inv2d_f64(const double* A, const double* Hb, const double* V, const double* D, double* out, double* tmp,
                             int batch, int Nr, int Nc, int level, const PwtFilters64& f, cudaStream_t st);
int pwt_launch_swt_fwd1d_f64(const double* in, double* A, double* D, int rows, int Nc, int level, const PwtFilters64& f,
                             cudaStream_t st);
int pwt_launch_swt_inv1d_f64(const double* A, const double* D, double* out, int rows, int Nc, int level,
                             const PwtFilters64& f, cudaStream_t st);
// kernels_f64_fused.cu : one double-precision 2D DWT level per launch, row and column pass fused (F = 4 .. 40).  0: not covered.
int pwt64_fused_fwd2d(const double* in, double* A, double* Hb, double* V, double* D, int batch, int Nr, int Nc, long long in_bs,
                      long long out_bs, const PwtFilters64& f, cudaStream_t st);
int pwt64_fused_inv2d(const double* A, const double* Hb, const double* V, const double* D, double* out, int batch, int nr, int nc,
                      int Nro, int Nco, long long in_bs, long long out_bs, const PwtFilters64& f, cudaStream_t st);
// non-separable (true 2D stencils).  k2d = device array of 4*hlen*hlen taps: LL, LH, HL, HH.
int pwt_launch_ns_fwd2d(const float* in, float* A, float* Hb, float* V, float* D, int batch, int Nr,
                        int Nc, long long in_bs, long long out_bs, const float* k2d, int hlen,
                        cudaStream_t st);
int pwt_launch_ns_inv2d(const float* A, const float* Hb, const float* V, const float* D, float* out,
                        int batch, int nr, int nc, int Nr_out, int Nc_out, long long in_bs,
                        long long out_bs, const float* k2d, int hlen, cudaStream_t st);
int pwt_launch_ns_swt_fwd2d(const float* in, float* A, float* Hb, float* V, float* D, int batch,
                            int Nr, int Nc, int level, const float* k2d, int hlen, cudaStream_t st);
int pwt_launch_ns_swt_inv2d(const float* A, const float* Hb, const float* V, const float* D,
                            float* out, int batch, int Nr, int Nc, int level, const float* k2d,
                            int hlen, cudaStream_t st);

// kernels_ops.cu
int pwt_launch_eltwise(const PwtSegTable& t, int op, cudaStream_t st);
int pwt_launch_group_soft(float* h, float* v, float* d, float* a, long long n, float beta,
                          cudaStream_t st);
int pwt_launch_norms(const PwtSegTable& t, double* d_acc /* [2]: l1, l2sq */, cudaStream_t st);
int pwt_launch_axpy(const PwtSegTable& dst, const PwtSegTable& src, float alpha, cudaStream_t st);
int pwt_launch_circshift(const float* in, float* out, int batch, int Nr, int Nc, int sr, int sc,
                         cudaStream_t st);
int pwt_launch_fill(float* p, long long n, float v, cudaStream_t st);

// kernels_fast.cu : register/shuffle-blocked kernels for short filters on the headline path.
// Return 0 when the configuration is not covered (caller falls back to the generic kernels).
// hint_flags: PWT_HINT_OUT_FEEDS_NEXT = the approximation / image written by this launch is the input
// of the next launch (keep it L2-resident), PWT_HINT_IN_FROM_PREV = the input came from the previous one.
#define PWT_HINT_OUT_FEEDS_NEXT 4
#define PWT_HINT_IN_FROM_PREV 8
int pwt_fast_dwt_fwd2d(const float* in, float* A, float* Hb, float* V, float* D, int batch, int Nr,
                       int Nc, long long in_bs, long long out_bs, const PwtFilters& f, bool haar,
                       int hint_flags, cudaStream_t st);
int pwt_fast_dwt_inv2d(const float* A, const float* Hb, const float* V, const float* D, float* out,
                       int batch, int nr, int nc, int Nr_out, int Nc_out, long long in_bs,
                       long long out_bs, const PwtFilters& f, bool haar, int hint_flags, cudaStream_t st);

// kernels_reg.cu : register-resident kernels (warp shuffles, no shared memory) for F <= 10 on
// 128-column-aligned planes.  Return 0 when the configuration is not covered.
int pwt_reg_dwt_fwd2d(const float* in, float* A, float* Hb, float* V, float* D, int batch, int Nr,
                      int Nc, long long in_bs, long long out_bs, const PwtFilters& f, bool haar,
                      int hint_flags, cudaStream_t st);
int pwt_reg_dwt_inv2d(const float* A, const float* Hb, const float* V, const float* D, float* out,
                      int batch, int nr, int nc, int Nr_out, int Nc_out, long long in_bs,
                      long long out_bs, const PwtFilters& f, bool haar, int hint_flags, cudaStream_t st);

// kernels_fused.cu : levels 1..3 in one launch (register cascade).  Return 0 when not covered.
// Dynamic task queue of the persistent kernels: a device counter that only ever grows; the host
// tracks how far each launch advances it (tasks + one failing pull per warp), so no reset is needed.
struct PwtTaskQueue {
    unsigned* counter;     // device memory, zero-initialised at plan creation
    unsigned base;         // value of *counter when the next launch starts
};
// partials (may be null): [tasks][2] doubles receiving each task's sum |c| and sum c^2 of the coefficients it
// stored (levels 1-3 details, plus A3 when count_a3) -- the norm reduction fused into the transform pass.
int pwt_fused_fwd3_max_tasks(int batch, int Nr, int Nc);
int pwt_fused_dwt_fwd3(const float* in, float* A3, float* const* H, float* const* V, float* const* D,
                       int batch, int Nr, int Nc, const PwtFilters& f, bool haar, PwtTaskQueue* q,
                       double* partials, int partials_cap, int count_a3, int* ntasks_out, cudaStream_t st);
int pwt_launch_reduce_partials(const double* partials, int n, double* d_acc, int add, cudaStream_t st);
// A threshold recorded by soft/hard_threshold() and not yet applied to memory: the fused inverse applies it
// to the coefficients while loading them (the denoising loop forward -> threshold -> inverse then moves
// no extra bytes).  beta[l] = threshold of level l+1 details, beta_app = threshold of A (if app).
struct PwtDeferredOp {
    int op;                // -1: nothing pending, else PWT_OP_SOFT / PWT_OP_HARD
    int app;
    float beta[PWT_MAX_LEVELS];
    float beta_app;
};
int pwt_fused_dwt_inv3(const float* A3, const float* const* H, const float* const* V, const float* const* D,
                       float* out, int batch, int Nr, int Nc, const PwtFilters& f, bool haar, PwtTaskQueue* q,
                       const PwtDeferredOp* op, int plain_launch, cudaStream_t st);

// kernels_swt.cu : fused (row + column) a-trous level in registers.  Return 0 when not covered.
int pwt_fast_swt_fwd2d(const float* in, float* A, float* Hb, float* V, float* D, int batch, int Nr, int Nc,
                       int level, const PwtFilters& f, cudaStream_t st);
int pwt_fast_swt_inv2d_covers(int batch, int Nr, int Nc, int level, const PwtFilters& f, const void* A,
                              const void* out);
int pwt_fast_swt_inv2d(const float* A, const float* Hb, const float* V, const float* D, float* out, int batch,
                       int Nr, int Nc, int level, const PwtFilters& f, int thr_op, float beta, int app,
                       float beta_app, cudaStream_t st);

// kernels_tile.cu : compile-time F = 10..40 shared-memory tile kernels (FMA-bound regime), any size.
int pwt_tile_dwt_fwd2d(const float* in, float* A, float* Hb, float* V, float* D, int batch, int Nr, int Nc,
                       long long in_bs, long long out_bs, const PwtFilters& f, cudaStream_t st);
int pwt_tile_dwt_inv2d(const float* A, const float* Hb, const float* V, const float* D, float* out, int batch,
                       int nr, int nc, int Nr_out, int Nc_out, long long in_bs, long long out_bs,
                       const PwtFilters& f, cudaStream_t st);

// kernels_strip.cu : compile-time F = 6..40 streaming strip kernels (row pass from shared memory, transposed-form
// column pass in registers), any size.  Return 0 when not covered.
int pwt_strip_dwt_fwd2d(const float* in, float* A, float* Hb, float* V, float* D, int batch, int Nr, int Nc,
                        long long in_bs, long long out_bs, const PwtFilters& f, cudaStream_t st);
int pwt_strip_dwt_inv2d(const float* A, const float* Hb, const float* V, const float* D, float* out, int batch,
                        int nr, int nc, int Nr_out, int Nc_out, long long in_bs, long long out_bs,
                        const PwtFilters& f, cudaStream_t st);
// batched 1D (every row an independent signal): F = 4..40
int pwt_strip_dwt_fwd1d(const float* in, float* A, float* D, int rows, int Nc, const PwtFilters& f, cudaStream_t st);
int pwt_strip_dwt_inv1d(const float* A, const float* D, float* out, int rows, int nc, int Nc_out, const PwtFilters& f,
                        cudaStream_t st);
// Haar, batched 1D, width multiple of 8: flat streaming butterfly.  Return 0 when not covered.
int pwt_haar_fwd1d_flat(const float* in, float* A, float* D, int rows, int Nc, cudaStream_t st);
int pwt_haar_inv1d_flat(const float* A, const float* D, float* out, int rows, int nc, int Nc_out, cudaStream_t st);
// pwt_plan.cu : one separable 2D level over a stack with the automatic kernel choice (for pwt_vol.cu)
int pwt_level_fwd2d(const float* src, float* A, float* Hb, float* V, float* D, int batch, int nr, int nc, long long in_bs,
                    long long out_bs, const PwtFilters& f, bool haar, cudaStream_t st);
int pwt_level_inv2d(const float* A, const float* Hb, const float* V, const float* D, float* dst, int batch, int nr, int nc,
                    int Nro, int Nco, long long in_bs, long long out_bs, const PwtFilters& f, bool haar, cudaStream_t st);
int pwt_set_error(int code, const char* msg);
// kernels_swt2p.cu : 2D a-trous level as two streaming passes (any even F <= 40, any size): the fallback behind the fused
// SWT kernels.  tmp holds 2 * batch * Nr * Nc floats.  Return 0 when not covered.
int pwt_swt2p_fwd2d(const float* in, float* A, float* Hb, float* V, float* D, float* tmp, int batch, int Nr, int Nc, int level,
                    const PwtFilters& f, cudaStream_t st);
int pwt_swt2p_inv2d(const float* A, const float* Hb, const float* V, const float* D, float* out, float* tmp, int batch, int Nr,
                    int Nc, int level, const PwtFilters& f, cudaStream_t st);
// kernels_row1d.cu : batched 1D DWT / IDWT, every level in ONE launch (rows staged once in shared memory).  D[l] = detail
// band of level l + 1.  Return 0 when not covered (row too long for a CTA's shared memory, odd filter length).
int pwt_row_dwt_fwd1d_all(const float* in, float* A, float* const* D, int rows, int Nc, int L, const PwtFilters& f,
                          cudaStream_t st);
int pwt_row_swt_fwd1d_all(const float* in, float* A, float* const* D, int rows, int Nc, int L, const PwtFilters& f,
                          cudaStream_t st);
int pwt_row_swt_inv1d_all(const float* A, float* const* D, float* out, int rows, int Nc, int L, const PwtFilters& f,
                          cudaStream_t st);
int pwt_row_dwt_inv1d_all(const float* A, float* const* D, float* out, int rows, int Nc, int L, const PwtFilters& f,
                          cudaStream_t st);
// kernels_swt1d.cu : batched 1D a-trous level from a staged shared-memory row.  Return 0 when not covered.
int pwt_fast_swt_fwd1d(const float* in, float* A, float* D, int rows, int Nc, int level, const PwtFilters& f,
                       cudaStream_t st);
int pwt_fast_swt_inv1d(const float* A, const float* D, float* out, int rows, int Nc, int level, const PwtFilters& f,
                       cudaStream_t st);
// kernels_swt_strip.cu : 2D a-trous forward level, streaming strip design (F <= 16).  Return 0 when not covered.
int pwt_strip_swt_fwd2d(const float* in, float* A, float* Hb, float* V, float* D, int batch, int Nr, int Nc,
                        int level, const PwtFilters& f, cudaStream_t st);
int pwt_strip_swt_inv2d(const float* A, const float* Hb, const float* V, const float* D, float* out, int batch,
                        int Nr, int Nc, int level, const PwtFilters& f, int thr_op, float beta, int app,
                        float beta_app, cudaStream_t st);
int pwt_strip_swt_inv2d_covers(int batch, int Nr, int Nc, int level, const PwtFilters& f, const void* A, const void* out);
int pwt_strip_dwt_fwd2d_norms(const float* in, float* A, float* Hb, float* V, float* D, int batch, int Nr, int Nc,
                              long long in_bs, long long out_bs, const PwtFilters& f, double* partials, int cap,
                              int count_a, int* written, cudaStream_t st);
int pwt_strip_dwt_inv2d_thr(const float* A, const float* Hb, const float* V, const float* D, float* out, int batch,
                            int nr, int nc, int Nr_out, int Nc_out, long long in_bs, long long out_bs,
                            const PwtFilters& f, int thr_op, float beta, int app, float beta_app, cudaStream_t st);
