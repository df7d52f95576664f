// Plan object + C ABI (include/pwt_b200.h).  Host-side driver of the hot path: owns the device
// slab, the per-instance filters, the stream, and the reference's state machine and dispatch
// (pdwt/src/wt.cu:84-665), re-designed around fused single-launch-per-level kernels.
//
// Device memory layout (one cudaMalloc per plan, 256-byte aligned sub-buffers):
//   [ image  B*Nr*Nc ][ image2 (cycle spinning only) ][ band 0 = A ][ band 1 ] ... [ tmp ]
// band 0 is allocated at level-1 size like the reference (common.cu:421-423) because it doubles as
// the ping-pong buffer of the intermediate approximations; every band is a dense row-major array
// (per image of the stack), so pwt_coeff_ptr() exposes the same layout as the reference.
#include <dlfcn.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <strings.h>

#include "../../include/pwt_b200.h"
#include "pwt_internal.h"

// Short filters outside the 3-level cascade (1- and 2-level plans): levels of at most this many samples (over the whole stack) take the strip kernels
// instead of the register kernels.  Measured (tools/gpu_levels12.py, fwd+inv): 2048^2 db2 1 level 0.0229 -> 0.0150 ms, 2 levels
// 0.0409 -> 0.0222; 4096^2 db2 2 levels 0.0711 -> 0.0620 but 1 level 0.0461 -> 0.0529 (F = 4: up to 2048^2); 4096^2 db3 1 level
// 0.0593 -> 0.0528, 2 levels 0.0879 -> 0.0709 (F = 6: up to 4096^2).  PWT_TAIL_PX overrides both (A/B knob, read once).
// (8192^2 db3 2 levels: a 4096^2 strip level behind the register kernel of level 1 was slower, 0.260 -> 0.294 ms: the larger
// threshold of F = 6 only applies when the whole image is at most 4096^2.)
static long long tail_plane_px(int hlen, long long image_px) {
    static const long long v = [] { const char* e = getenv("PWT_TAIL_PX"); return e && *e ? atoll(e) : -1LL; }();
    if (v >= 0) return v;
    return (hlen > 4 && image_px <= (1LL << 24)) ? (1LL << 24) : (1LL << 22);
}

// Haar planes of at most this many samples take the flat butterfly kernels ahead of the register kernels (PWT_HAAR_FLAT_PX;
// default 4096^2).  A/B, fwd+inv: 2048^2 1 level 0.0201 -> 0.0106 ms, 2 levels 0.0320 -> 0.0164; 4096^2 1 level 0.0434 -> 0.0278,
// 2 levels 0.0710 -> 0.0394; 8192^2 2 levels 0.2189 -> 0.2060, 4 levels 0.2001 -> 0.1887; 8192^2 1 level: equal (0.172).
static long long haar_flat_px() {
    static const long long v = [] { const char* e = getenv("PWT_HAAR_FLAT_PX"); return e && *e ? atoll(e) : (1LL << 24); }();
    return v;
}
// pwt_plan64.cu: the tiled row kernels of the double-precision plans instantiated for float -- batched 1D levels with few, long rows
int pwt_rows1d_fwd_f32(const float* in, float* A, float* D, long long rows, int Nc, const PwtFilters& f, cudaStream_t st);
int pwt_rows1d_inv_f32(const float* A, const float* D, float* out, long long rows, int nc, int Nc_out, const PwtFilters& f, cudaStream_t st);
// kernels_haar2d.cu: the Haar butterfly of a 2D level for the sizes the register kernels do not take (any size; 1 launch)
int pwt_haar2d_fwd_flat(const float* in, float* A, float* Hb, float* V, float* D, int batch, int Nr, int Nc, long long in_bs,
                        long long out_bs, cudaStream_t st);
int pwt_haar2d_inv_flat(const float* A, const float* Hb, const float* V, const float* D, float* out, int batch, int nr, int nc,
                        int Nro, int Nco, long long in_bs, long long out_bs, cudaStream_t st);

int pwt_is_haar_alias(const char* wname);
int pwt_fill_filters(const char* wname, PwtFilters* out);

// ---- error plumbing --------------------------------------------------------------------------
static thread_local char g_err[512] = "";
static int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}
int pwt_set_error(int code, const char* msg) {       // for the other translation units (pwt_plan64.cu, pwt_vol.cu)
    snprintf(g_err, sizeof(g_err), "%s", msg);
    return code;
}
#define CK(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess)                                                                     \
            return fail(PWT_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_),      \
                        __FILE__, __LINE__);                                                       \
    } while (0)
#define CK_LAUNCH()                                                                                \
    do {                                                                                           \
        cudaError_t e_ = cudaGetLastError();                                                       \
        if (e_ != cudaSuccess)                                                                     \
            return fail(PWT_ERR_CUDA, "kernel launch failed: %s (%s:%d)", cudaGetErrorString(e_),  \
                        __FILE__, __LINE__);                                                       \
    } while (0)

// ---- process-wide tuning knobs and per-device facts (declared in pwt_internal.h) ---------------
static int env_i(const char* name, int dflt) {
    const char* e = getenv(name);
    return (e && *e) ? atoi(e) : dflt;
}
const PwtTuning& pwt_tuning() {
    static const PwtTuning t = [] {          // C++11 magic static: initialised once, thread-safe
        PwtTuning k;
        k.no_pdl = env_i("PWT_NO_PDL", 0);
        k.no_fused = env_i("PWT_NO_FUSED", 0);
        k.no_fused_inv = env_i("PWT_NO_FUSED_INV", 0);
        k.fused_variant = env_i("PWT_FUSED_VARIANT", 0);
        k.fused_t3 = env_i("PWT_FUSED_T3", 0);
        k.fused_inv_t3 = env_i("PWT_FUSED_INV_T3", 0);
        k.fused_pdl = env_i("PWT_FUSED_PDL", 3);
        k.no_fused_norms = getenv("PWT_NO_FUSED_NORMS") ? 1 : 0;
        k.no_defer = getenv("PWT_NO_DEFER") ? 1 : 0;
        k.tile_min_f = env_i("PWT_TILE_MIN_F", 22);
        k.strip_min_f = env_i("PWT_STRIP_MIN_F", 8);
        k.ns_direct = getenv("PWT_NS_DIRECT") ? 1 : 0;
        k.l2_persist_mb = env_i("PWT_L2_PERSIST_MB", 0);
        k.verbose = getenv("PWT_VERBOSE") ? 1 : 0;
        k.strip_segs = env_i("PWT_STRIP_SEGS", 0);
        k.strip_occ_fwd = env_i("PWT_STRIP_OCC_FWD", 0);
        k.strip_occ_inv = env_i("PWT_STRIP_OCC_INV", 0);
        k.swt_nbuf = env_i("PWT_SWT_NBUF", 1);
        k.no_strip_swt = getenv("PWT_NO_STRIP_SWT") ? 1 : 0;
        k.no_fast_swt = getenv("PWT_NO_FAST_SWT") ? 1 : 0;
        k.swt_tq = env_i("PWT_SWT_TQ", 0);
        k.fast_tile_rows = env_i("PWT_FAST_TILE_ROWS", 0);
        k.fwd_variant = env_i("PWT_FWD_VARIANT", -1);
        k.reg_tile_rows = env_i("PWT_REG_TILE_ROWS", 16);
        k.use_hints = env_i("PWT_USE_HINTS", 0);
        k.reg_fwd_variant = env_i("PWT_REG_FWD_VARIANT", 2);
        k.no_fold_cs = env_i("PWT_NO_FOLD_CS", 0);
        k.no_cascade8 = env_i("PWT_NO_CASCADE8", 0);
        k.no_fused1d = env_i("PWT_NO_FUSED1D", 0);
        k.tail_strip = env_i("PWT_TAIL_STRIP", 1);
        k.strip_thr_occ3 = env_i("PWT_STRIP_THR_OCC3", 1);
        return k;
    }();
    return t;
}
int pwt_sm_count() {
    static int sms[PWT_MAX_DEVICES];         // 0 = not queried yet; racing first calls store the same value
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= PWT_MAX_DEVICES) return 148;
    int v = __atomic_load_n(&sms[dev], __ATOMIC_RELAXED);
    if (!v) {
        if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v < 1) v = 148;
        __atomic_store_n(&sms[dev], v, __ATOMIC_RELAXED);
    }
    return v;
}

// ---- minimal NCCL binding (dlopen: the product has no link-time dependency on NCCL) -----------
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId_t;
struct NcclApi {
    void* lib;
    int (*GetUniqueId)(ncclUniqueId_t*);
    int (*CommInitRank)(ncclComm_t*, int, ncclUniqueId_t, int);
    int (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t);
    int (*CommDestroy)(ncclComm_t);
    const char* (*GetErrorString)(int);
    int (*CommInitAll)(ncclComm_t*, int, const int*);
    int (*GroupStart)();
    int (*GroupEnd)();
};
static NcclApi g_nccl = {};
static int load_nccl() {
    static std::mutex m;
    std::lock_guard<std::mutex> guard(m);
    if (g_nccl.lib) return PWT_OK;
    const char* cands[] = {getenv("PWT_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
    void* lib = nullptr;
    for (const char* c : cands) {
        if (!c || !*c) continue;
        lib = dlopen(c, RTLD_NOW | RTLD_GLOBAL);
        if (lib) break;
    }
    if (!lib) return fail(PWT_ERR_COMM, "cannot dlopen libnccl.so.2 (set PWT_NCCL_LIB): %s", dlerror());
    g_nccl.GetUniqueId = (int (*)(ncclUniqueId_t*))dlsym(lib, "ncclGetUniqueId");
    g_nccl.CommInitRank = (int (*)(ncclComm_t*, int, ncclUniqueId_t, int))dlsym(lib, "ncclCommInitRank");
    g_nccl.AllReduce = (int (*)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t))dlsym(
        lib, "ncclAllReduce");
    g_nccl.CommDestroy = (int (*)(ncclComm_t))dlsym(lib, "ncclCommDestroy");
    g_nccl.GetErrorString = (const char* (*)(int))dlsym(lib, "ncclGetErrorString");
    g_nccl.CommInitAll = (int (*)(ncclComm_t*, int, const int*))dlsym(lib, "ncclCommInitAll");
    g_nccl.GroupStart = (int (*)())dlsym(lib, "ncclGroupStart");
    g_nccl.GroupEnd = (int (*)())dlsym(lib, "ncclGroupEnd");
    if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.AllReduce || !g_nccl.CommDestroy)
        return fail(PWT_ERR_COMM, "libnccl is missing required symbols");
    g_nccl.lib = lib;
    return PWT_OK;
}
enum { kNcclFloat64 = 8, kNcclSum = 0 };

#define PWT_PROF_CAP 512
// ---- the plan --------------------------------------------------------------------------------
struct pwt_plan {
    int device;
    cudaStream_t stream;
    cudaEvent_t ev0, ev1;
    int batch, Nr, Nc, ndims, nlevels, hlen, do_swt, do_separable, do_cs;
    int state, shift_r, shift_c;
    // cycle spinning of a 2D SWT plan: the a-trous transform commutes EXACTLY with circular shifts (period N, no
    // decimation), so the shifts of wt.cu:242-246 / :303 are recorded instead of executed: the logical image is
    // circshift(d_image, vs_img), every logical band circshift(d_band[b], vs_coef).  Thresholds, shrink and norms do not
    // look at positions; whoever does (get/set of image or bands, raw pointers, add_wavelet, clone) materialises first.
    int vs_img_r, vs_img_c, vs_coef_r, vs_coef_c;
    char wname[128];
    PwtFilters filt;
    float* d_k2d_fwd;   // non-separable analysis filters  [LL, LH, HL, HH], hlen*hlen each
    float* d_k2d_inv;   // non-separable synthesis filters
    int lvNr[PWT_MAX_LEVELS + 1], lvNc[PWT_MAX_LEVELS + 1];
    float* slab;
    size_t slab_floats;
    float* d_image;
    float* d_image2;    // second image plane (cycle spinning): shifts are a kernel + pointer swap
    float* d_tmp;
    size_t tmp_floats;
    int nbands;
    float* d_band[PWT_MAX_BANDS];
    size_t coef_base, coef_floats;   // [coef_base, coef_base + coef_floats) of the slab holds every band at its final size
    int band_nr[PWT_MAX_BANDS], band_nc[PWT_MAX_BANDS];
    double* d_acc;      // device accumulators for the norms
    PwtTaskQueue queue; // dynamic task queue of the persistent kernels
    double* d_partials; // per-task |c|, c^2 sums written by the fused forward (norm reduction fused into the pass)
    int partials_n;     // > 0: d_partials holds the sums of bands 1..9 (+A if 3 levels) of the CURRENT coefficients
    int partials_cap;
    unsigned norm_lvl_mask;  // bit l-1: the detail bands of level l are covered by d_partials
    int norm_a;              // the approximation band is covered by d_partials
    int want_norms;     // norms were requested after a forward: later forwards accumulate them in-kernel (~2 % of the pass)
    PwtDeferredOp pend; // threshold recorded but not yet applied to memory (pend.op < 0: none)
    int defer_ok;       // plan shape for which thresholds may be deferred into the fused inverse
    int ns_rank1;       // non-separable plan whose four 2D filters are outer products of the 1D bank (always, unless custom
                        // 2D filters were loaded): evaluated with the separable kernels, detail slots 1/2 swapped (quirk Q1)
    int defer_strip_ok; // 2D DWT plan whose finest level the strip inverse serves: thresholds applied as it stages the bands
    int defer_swt_ok;   // SWT plan whose every level is served by the fused SWT inverse (threshold applied on load)
    double* h_acc;      // pinned mirror
    void* d_flush;
    size_t flush_bytes;
    long long launches;
    int kernel_mode;
    int tile_min_f;        // filter length from which the FMA-oriented tile kernels are preferred
    int strip_min_f;       // filter length from which the streaming strip kernels are preferred
    unsigned custom_len;   // taps given to set_filters_forward (0: built-in bank)
    int prof_on, prof_n;
    cudaEvent_t* prof_ev;  // 2 events per record
    int prof_tag[PWT_PROF_CAP];
    ncclComm_t comm;
    int comm_nranks;
};

// ---- deferred thresholds (see PwtDeferredOp) -------------------------------------------------
// flush: apply the pending threshold to memory for levels >= first_level (1 = everything) and, if
// with_app, to the approximation band; clears the pending state when everything was flushed.
static int flush_pending(pwt_plan* p, int first_level, bool with_app);
static int materialize_image(pwt_plan* p);     // recorded cycle-spinning shifts (2D SWT plans), see pwt_plan::vs_img_r
static int materialize_coeffs(pwt_plan* p);
// thresholds of this plan can be applied by the strip inverse while it stages the coefficients
static inline bool strip_defer_capable(const pwt_plan* p) {
    return p->defer_strip_ok && p->kernel_mode == 0 && p->do_separable && p->hlen >= p->strip_min_f && p->hlen <= 40 &&
           (p->hlen & 1) == 0 && p->hlen != 2;
}

static inline int div2i(int n) { return (n + 1) >> 1; }            // utils.cu:24-27
static inline int ilog2i(int i) {                                   // utils.cu:14-20 (guarded)
    int l = 0;
    while (i > 1) {
        i >>= 1;
        ++l;
    }
    return l;
}
static inline size_t align64(size_t nfloats) { return (nfloats + 63) & ~(size_t)63; }
static inline bool is_haar(const pwt_plan* p) { return p->hlen == 2 && !p->do_swt; }  // wt.cu:248,255
static inline void clear_partials(pwt_plan* p) { p->partials_n = 0; p->norm_lvl_mask = 0; p->norm_a = 0; }
static inline long long img_elems(const pwt_plan* p) { return (long long)p->Nr * p->Nc; }
static inline long long band_elems(const pwt_plan* p, int b) {
    return (long long)p->band_nr[b] * p->band_nc[b];
}
static inline long long lvl_elems(const pwt_plan* p, int l) { return (long long)p->lvNr[l] * p->lvNc[l]; }

static int build_k2d(pwt_plan* p) {
    // nonseparable.cu:70-74: LL = lo(x)lo, LH = lo(x)hi, HL = hi(x)lo, HH = hi(x)hi with
    // res[i*len+j] = a[i]*b[j]  (i = y tap, j = x tap)
    const int F = p->hlen;
    const size_t n = (size_t)4 * F * F;
    float* h = (float*)malloc(2 * n * sizeof(float));
    if (!h) return fail(PWT_ERR_NOMEM, "out of host memory");
    for (int dir = 0; dir < 2; dir++) {
        const float* lo = dir ? p->filt.IL : p->filt.L;
        const float* hi = dir ? p->filt.IH : p->filt.H;
        float* o = h + dir * n;
        for (int i = 0; i < F; i++)
            for (int j = 0; j < F; j++) {
                o[0 * F * F + i * F + j] = lo[i] * lo[j];
                o[1 * F * F + i * F + j] = lo[i] * hi[j];
                o[2 * F * F + i * F + j] = hi[i] * lo[j];
                o[3 * F * F + i * F + j] = hi[i] * hi[j];
            }
    }
    cudaError_t e = cudaMemcpyAsync(p->d_k2d_fwd, h, n * sizeof(float), cudaMemcpyHostToDevice, p->stream);
    if (e == cudaSuccess)
        e = cudaMemcpyAsync(p->d_k2d_inv, h + n, n * sizeof(float), cudaMemcpyHostToDevice, p->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(p->stream);
    free(h);
    if (e != cudaSuccess) return fail(PWT_ERR_CUDA, "2D filter upload failed: %s", cudaGetErrorString(e));
    return PWT_OK;
}

static void compute_geometry(pwt_plan* p) {
    p->lvNr[0] = p->Nr;
    p->lvNc[0] = p->Nc;
    for (int l = 1; l <= p->nlevels; l++) {
        if (p->do_swt) {
            p->lvNr[l] = p->Nr;
            p->lvNc[l] = p->Nc;
        } else {
            p->lvNr[l] = p->ndims == 2 ? div2i(p->lvNr[l - 1]) : p->Nr;
            p->lvNc[l] = div2i(p->lvNc[l - 1]);
        }
    }
    const int L = p->nlevels;
    if (p->ndims == 2) {
        p->nbands = 3 * L + 1;
        for (int i = 0; i < L; i++)
            for (int j = 1; j <= 3; j++) {
                p->band_nr[3 * i + j] = p->lvNr[i + 1];
                p->band_nc[3 * i + j] = p->lvNc[i + 1];
            }
    } else {
        p->nbands = L + 1;
        for (int i = 0; i < L; i++) {
            p->band_nr[i + 1] = p->lvNr[i + 1];
            p->band_nc[i + 1] = p->lvNc[i + 1];
        }
    }
    p->band_nr[0] = p->lvNr[L];
    p->band_nc[0] = p->lvNc[L];
}

// every level of a separable 2D SWT plan is served by the fused SWT inverse (re-evaluated when filters change)
static int swt_all_levels_fused(const pwt_plan* p) {
    if (!(p->ndims == 2 && p->do_swt && p->do_separable)) return 0;
    for (int l = 1; l <= p->nlevels; l++)
        if (!pwt_strip_swt_inv2d_covers(p->batch, p->Nr, p->Nc, l, p->filt, p->d_band[0], p->d_image) &&
            !pwt_fast_swt_inv2d_covers(p->batch, p->Nr, p->Nc, l, p->filt, p->d_band[0], p->d_image))
            return 0;
    return 1;
}

static int alloc_plan(pwt_plan* p) {
    const size_t B = (size_t)p->batch;
    const size_t img = align64(B * img_elems(p));
    size_t total = img + (p->do_cs ? img : 0);
    // coefficient region: detail bands first, the approximation LAST, so that [band 1 .. band N-1][A at its final size]
    // is one contiguous piece of memory -- pwt_get_coeffs moves every band to the host with ONE copy.  Band 0 keeps its
    // level-1-sized allocation (ping-pong plane of the intermediate approximations) behind that piece.
    size_t off[PWT_MAX_BANDS];
    p->coef_base = total;
    for (int b = 1; b <= p->nbands; b++) {
        const int bb = b == p->nbands ? 0 : b;
        off[bb] = total;
        if (bb == 0) p->coef_floats = total + B * (size_t)band_elems(p, 0) - p->coef_base;
        const size_t n = bb == 0 ? B * (size_t)lvl_elems(p, 1) : B * (size_t)band_elems(p, bb);
        total += align64(n);
    }
    // scratch: DWT needs one approximation plane (we keep a full image so circshift can use it);
    // the unfused SWT path needs two full planes for (lo, hi)/(t1, t2) plus the A ping-pong plane.
    p->tmp_floats = (p->do_swt && p->ndims == 2) ? 3 * img : img;
    const size_t tmp_off = total;
    total += p->tmp_floats;
    p->slab_floats = total;
    CK(cudaMalloc((void**)&p->slab, total * sizeof(float)));
    CK(cudaMemsetAsync(p->slab, 0, total * sizeof(float), p->stream));
    p->d_image = p->slab;
    p->d_image2 = p->do_cs ? p->slab + img : nullptr;
    for (int b = 0; b < p->nbands; b++) p->d_band[b] = p->slab + off[b];
    p->d_tmp = p->slab + tmp_off;
    const size_t k2d = (size_t)4 * PWT_MAX_TAPS * PWT_MAX_TAPS * sizeof(float);
    CK(cudaMalloc((void**)&p->d_k2d_fwd, k2d));
    CK(cudaMalloc((void**)&p->d_k2d_inv, k2d));
    CK(cudaMalloc((void**)&p->d_acc, 2 * sizeof(double)));
    p->partials_cap = (p->ndims == 2 && !p->do_swt && p->nlevels >= 3 && p->Nr % 8 == 0 && p->Nc % 8 == 0)
                          ? pwt_fused_fwd3_max_tasks(p->batch, p->Nr, p->Nc) : 0;
    if (p->ndims == 2 && !p->do_swt) p->partials_cap += 32768;      // one pair per CTA of the strip forward launches
    p->d_partials = nullptr;
    clear_partials(p);
    p->want_norms = 0;
    if (pwt_tuning().no_fused_norms) p->partials_cap = 0;
    if (p->partials_cap > 0) CK(cudaMalloc((void**)&p->d_partials, (size_t)p->partials_cap * 2 * sizeof(double)));
    CK(cudaMalloc((void**)&p->queue.counter, sizeof(unsigned)));
    CK(cudaMemsetAsync(p->queue.counter, 0, sizeof(unsigned), p->stream));
    p->queue.base = 0;
    CK(cudaMallocHost((void**)&p->h_acc, 2 * sizeof(double)));
    // SWT plans whose every level the fused inverse serves may defer thresholds into it (pointers are known now)
    p->defer_swt_ok = swt_all_levels_fused(p) && !pwt_tuning().no_defer;
    return PWT_OK;
}

extern "C" void pwt_destroy(pwt_plan* p) {
    if (!p) return;
    cudaSetDevice(p->device);
    if (p->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(p->comm);
    if (p->stream) cudaStreamSynchronize(p->stream);
    if (p->slab) cudaFree(p->slab);
    if (p->d_k2d_fwd) cudaFree(p->d_k2d_fwd);
    if (p->d_k2d_inv) cudaFree(p->d_k2d_inv);
    if (p->d_acc) cudaFree(p->d_acc);
    if (p->queue.counter) cudaFree(p->queue.counter);
    if (p->d_partials) cudaFree(p->d_partials);
    if (p->h_acc) cudaFreeHost(p->h_acc);
    if (p->d_flush) cudaFree(p->d_flush);
    if (p->prof_ev) {
        for (int i = 0; i < 2 * PWT_PROF_CAP; i++)
            if (p->prof_ev[i]) cudaEventDestroy(p->prof_ev[i]);
        free(p->prof_ev);
    }
    if (p->ev0) cudaEventDestroy(p->ev0);
    if (p->ev1) cudaEventDestroy(p->ev1);
    if (p->stream) cudaStreamDestroy(p->stream);
    free(p);
}

extern "C" int pwt_create_batch(pwt_plan** out, const float* img, int batch, int Nr, int Nc,
                                const char* wname, int levels, int memisonhost, int do_separable,
                                int do_cycle_spinning, int do_swt, int ndim) {
    if (!out) return fail(PWT_ERR_ARG, "null output handle");
    *out = nullptr;
    if (!wname || Nr < 1 || Nc < 1 || batch < 1) return fail(PWT_ERR_ARG, "invalid geometry %dx%dx%d", batch, Nr, Nc);
    if ((long long)Nr * Nc >= (1LL << 31)) return fail(PWT_ERR_ARG, "one image must hold < 2^31 samples");
    if (ndim > 2) return fail(PWT_ERR_UNSUPPORTED, "ndim=%d is not implemented", ndim);   // wt.cu:171-175
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1)
        return fail(PWT_ERR_CUDA, "no CUDA device available (this library has no CPU fallback)");

    pwt_plan* p = (pwt_plan*)calloc(1, sizeof(pwt_plan));
    if (!p) return fail(PWT_ERR_NOMEM, "out of host memory");
    cudaGetDevice(&p->device);
    p->batch = batch;
    p->Nr = Nr;
    p->Nc = Nc;
    p->do_swt = do_swt ? 1 : 0;
    p->do_cs = do_cycle_spinning ? 1 : 0;
    p->state = PWT_INIT;
    p->pend.op = -1;
    p->ndims = (ndim < 2 || Nr == 1) ? 1 : 2;                       // wt.cu:133-136
    p->do_separable = (p->ndims == 1) ? 1 : (do_separable ? 1 : 0); // wt.cu:138-142
    strncpy(p->wname, wname, sizeof(p->wname) - 1);
    if (levels < 1) levels = 1;                                     // wt.cu:111-114

    // filters (separable.cu:19-54).  With do_swt only the table names are known (separable.cu:24-28).
    if (p->do_swt && pwt_is_haar_alias(wname) && strcasecmp(wname, "haar")) {
        free(p);
        return fail(PWT_ERR_UNKNOWN_WAVELET, "unknown wavelet '%s' for the stationary transform", wname);
    }
    const int hlen = pwt_fill_filters(wname, &p->filt);
    if (hlen < 0) {
        free(p);
        return fail(PWT_ERR_UNKNOWN_WAVELET, "unknown wavelet name '%s'", wname);
    }
    p->hlen = hlen;
    // level clipping (wt.cu:156-165)
    const int N = p->ndims == 2 ? (Nr < Nc ? Nr : Nc) : Nc;
    const int wmaxlev = ilog2i(N / (hlen - 1));
    if (wmaxlev < 1) {
        free(p);
        return fail(PWT_ERR_TOO_SMALL, "a %dx%d image is too small for wavelet %s (%d taps)", Nr, Nc, wname, hlen);
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
        return fail(PWT_ERR_UNSUPPORTED, "cycle spinning is not implemented for 1D. Use SWT instead.");
    }
    if (p->do_cs && p->do_swt)
        puts("Warning: makes little sense to use Cycle spinning with stationary Wavelet transform");
    compute_geometry(p);
    p->defer_ok = p->ndims == 2 && !p->do_swt && p->do_separable && p->nlevels >= 3 && Nr % 8 == 0 && Nc % 8 == 0 &&
                  Nc >= 512 && Nr >= 64 && (p->hlen <= 6) && !pwt_tuning().no_defer;
    p->tile_min_f = pwt_tuning().tile_min_f;
    p->strip_min_f = pwt_tuning().strip_min_f;
    p->defer_strip_ok = p->ndims == 2 && !p->do_swt && p->do_separable && p->nlevels >= 1 && p->lvNr[1] >= 32 &&
                        p->lvNc[1] >= 128 && !pwt_tuning().no_defer;     // + filter length, checked when used

    // Persisting-L2 carve-out: measured on B200 (profiles/r01_notes.md) to SLOW the level-1 kernels (0.10 -> 0.15-0.18 ms)
    // without speeding up level 2, so it is off unless PWT_L2_PERSIST_MB asks for it (once per process).
    if (pwt_tuning().l2_persist_mb > 0) {
        static std::once_flag l2_once;
        const int dev = p->device;
        std::call_once(l2_once, [dev] {
            int max_persist = 0;
            cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, dev);
            size_t want = (size_t)pwt_tuning().l2_persist_mb << 20;
            if (want > (size_t)max_persist) want = (size_t)max_persist;
            if (want > 0) cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, want);
            if (pwt_tuning().verbose) printf("pwt: persisting L2 carve-out %zu MB (max %d MB)\n", want >> 20, max_persist >> 20);
        });
    }

    int rc = PWT_OK;
    cudaError_t e = cudaStreamCreate(&p->stream);
    if (e == cudaSuccess) e = cudaEventCreate(&p->ev0);
    if (e == cudaSuccess) e = cudaEventCreate(&p->ev1);
    if (e != cudaSuccess) rc = fail(PWT_ERR_CUDA, "stream/event creation failed: %s", cudaGetErrorString(e));
    if (rc == PWT_OK) rc = alloc_plan(p);
    if (rc == PWT_OK && img) {
        e = cudaMemcpyAsync(p->d_image, img, (size_t)batch * img_elems(p) * sizeof(float),
                            memisonhost ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice, p->stream);
        if (e != cudaSuccess) rc = fail(PWT_ERR_CUDA, "image upload failed: %s", cudaGetErrorString(e));
    }
    if (rc == PWT_OK && !p->do_separable) rc = build_k2d(p);
    p->ns_rank1 = !p->do_separable && !pwt_tuning().ns_direct;
    if (rc == PWT_OK) {
        e = cudaStreamSynchronize(p->stream);
        if (e != cudaSuccess) rc = fail(PWT_ERR_CUDA, "plan initialisation failed: %s", cudaGetErrorString(e));
    }
    if (rc != PWT_OK) {
        pwt_destroy(p);
        return rc;
    }
    *out = p;
    return PWT_OK;
}

extern "C" int pwt_create(pwt_plan** out, const float* img, int Nr, int Nc, const char* wname,
                          int levels, int memisonhost, int do_separable, int do_cycle_spinning,
                          int do_swt, int ndim) {
    return pwt_create_batch(out, img, 1, Nr, Nc, wname, levels, memisonhost, do_separable,
                            do_cycle_spinning, do_swt, ndim);
}

extern "C" int pwt_clone(pwt_plan** out, const pwt_plan* src) {
    if (!out || !src) return fail(PWT_ERR_ARG, "null argument");
    pwt_plan* p = (pwt_plan*)malloc(sizeof(pwt_plan));
    if (!p) return fail(PWT_ERR_NOMEM, "out of host memory");
    memcpy(p, src, sizeof(pwt_plan));
    p->stream = nullptr;
    p->ev0 = p->ev1 = nullptr;
    p->slab = nullptr;
    p->d_k2d_fwd = p->d_k2d_inv = nullptr;
    p->d_acc = nullptr;
    p->queue.counter = nullptr;
    p->d_partials = nullptr;
    clear_partials(p);
    p->h_acc = nullptr;
    p->d_flush = nullptr;
    p->flush_bytes = 0;
    p->comm = nullptr;
    p->comm_nranks = 0;
    p->launches = 0;
    p->prof_ev = nullptr;
    p->prof_on = p->prof_n = 0;
    *out = nullptr;
    cudaSetDevice(src->device);
    flush_pending(const_cast<pwt_plan*>(src), 1, true);
    p->pend.op = -1;
    materialize_image(const_cast<pwt_plan*>(src));
    materialize_coeffs(const_cast<pwt_plan*>(src));
    p->vs_img_r = p->vs_img_c = p->vs_coef_r = p->vs_coef_c = 0;
    cudaStreamSynchronize(src->stream);
    int rc = PWT_OK;
    cudaError_t e = cudaStreamCreate(&p->stream);
    if (e == cudaSuccess) e = cudaEventCreate(&p->ev0);
    if (e == cudaSuccess) e = cudaEventCreate(&p->ev1);
    if (e != cudaSuccess) rc = fail(PWT_ERR_CUDA, "stream/event creation failed: %s", cudaGetErrorString(e));
    if (rc == PWT_OK) rc = alloc_plan(p);
    if (rc == PWT_OK) {
        // alloc_plan placed image/image2 in canonical order; keep the source's current roles
        if (src->d_image2 && src->d_image != src->slab) {
            float* t = p->d_image;
            p->d_image = p->d_image2;
            p->d_image2 = t;
        }
        const size_t k2d = (size_t)4 * PWT_MAX_TAPS * PWT_MAX_TAPS * sizeof(float);
        e = cudaMemcpyAsync(p->slab, src->slab, src->slab_floats * sizeof(float), cudaMemcpyDeviceToDevice, p->stream);
        if (e == cudaSuccess) e = cudaMemcpyAsync(p->d_k2d_fwd, src->d_k2d_fwd, k2d, cudaMemcpyDeviceToDevice, p->stream);
        if (e == cudaSuccess) e = cudaMemcpyAsync(p->d_k2d_inv, src->d_k2d_inv, k2d, cudaMemcpyDeviceToDevice, p->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(p->stream);
        if (e != cudaSuccess) rc = fail(PWT_ERR_CUDA, "plan copy failed: %s", cudaGetErrorString(e));
    }
    if (rc != PWT_OK) {
        pwt_destroy(p);
        return rc;
    }
    *out = p;
    return PWT_OK;
}

extern "C" int pwt_get_info(const pwt_plan* p, pwt_info* info) {
    if (!p || !info) return fail(PWT_ERR_ARG, "null argument");
    info->batch = p->batch;
    info->Nr = p->Nr;
    info->Nc = p->Nc;
    info->ndims = p->ndims;
    info->nlevels = p->nlevels;
    info->hlen = p->hlen;
    info->do_swt = p->do_swt;
    info->do_separable = p->do_separable;
    info->do_cycle_spinning = p->do_cs;
    info->state = p->state;
    info->shift_r = p->shift_r;
    info->shift_c = p->shift_c;
    info->nbands = p->nbands;
    info->device = p->device;
    return PWT_OK;
}

extern "C" int pwt_band_shape(const pwt_plan* p, int num, int* nr, int* nc) {
    if (!p || num < 0 || num >= p->nbands) return fail(PWT_ERR_ARG, "band %d out of range", num);
    if (nr) *nr = p->band_nr[num];
    if (nc) *nc = p->band_nc[num];
    return PWT_OK;
}

// ---- circular shift ---------------------------------------------------------------------------
static int do_circshift(pwt_plan* p, int sr, int sc, int inplace) {
    // common.cu:378-396
    const int Nr = p->Nr, Nc = p->Nc;
    if (sr < 0) sr += Nr;
    if (sc < 0) sc += Nc;
    sr %= Nr;
    sc %= Nc;
    if (sr < 0) sr += Nr;
    if (sc < 0) sc += Nc;
    if (p->ndims == 1) sr = 0;
    if (!inplace) {
        p->launches += pwt_launch_circshift(p->d_image, p->d_tmp, p->batch, Nr, Nc, sr, sc, p->stream);
    } else if (p->d_image2) {
        // one gather pass into the spare plane, then swap roles (8 B/px instead of the
        // reference's memcpy + kernel = 16 B/px)
        p->launches += pwt_launch_circshift(p->d_image, p->d_image2, p->batch, Nr, Nc, sr, sc, p->stream);
        float* t = p->d_image;
        p->d_image = p->d_image2;
        p->d_image2 = t;
    } else {
        p->launches += pwt_launch_circshift(p->d_image, p->d_tmp, p->batch, Nr, Nc, sr, sc, p->stream);
        CK(cudaMemcpyAsync(p->d_image, p->d_tmp, (size_t)p->batch * img_elems(p) * sizeof(float),
                           cudaMemcpyDeviceToDevice, p->stream));
    }
    CK_LAUNCH();
    return PWT_OK;
}

// recorded (lazy) cycle-spinning shifts, see pwt_plan::vs_img_r
static inline bool lazy_cs(const pwt_plan* p) {
    return p->do_cs && p->do_swt && p->ndims == 2 && !pwt_tuning().no_fold_cs;
}
static inline int modn(int a, int n) { a %= n; return a < 0 ? a + n : a; }
static int materialize_image(pwt_plan* p) {
    if (!(p->vs_img_r | p->vs_img_c)) return PWT_OK;
    const int r = p->vs_img_r, c = p->vs_img_c;
    p->vs_img_r = p->vs_img_c = 0;
    return do_circshift(p, r, c, 1);
}
static int materialize_coeffs(pwt_plan* p) {
    if (!(p->vs_coef_r | p->vs_coef_c)) return PWT_OK;
    const int r = p->vs_coef_r, c = p->vs_coef_c;
    p->vs_coef_r = p->vs_coef_c = 0;
    const size_t n = (size_t)p->batch * img_elems(p);               // SWT: every band is image-sized
    for (int b = 0; b < p->nbands; b++) {
        p->launches += pwt_launch_circshift(p->d_band[b], p->d_tmp, p->batch, p->Nr, p->Nc, r, c, p->stream);
        CK(cudaMemcpyAsync(p->d_band[b], p->d_tmp, n * sizeof(float), cudaMemcpyDeviceToDevice, p->stream));
    }
    CK_LAUNCH();
    return PWT_OK;
}

extern "C" int pwt_circshift(pwt_plan* p, int sr, int sc, int inplace) {
    if (!p) return fail(PWT_ERR_ARG, "null plan");
    cudaSetDevice(p->device);
    int rc = materialize_image(p);
    if (rc != PWT_OK) return rc;
    return do_circshift(p, sr, sc, inplace);
}

// ---- per-launch profiling --------------------------------------------------------------------
static inline void prof_begin(pwt_plan* p, int tag) {
    if (!p->prof_on || p->prof_n >= PWT_PROF_CAP) return;
    p->prof_tag[p->prof_n] = tag;
    cudaEventRecord(p->prof_ev[2 * p->prof_n], p->stream);
}
static inline void prof_end(pwt_plan* p) {
    if (!p->prof_on || p->prof_n >= PWT_PROF_CAP) return;
    cudaEventRecord(p->prof_ev[2 * p->prof_n + 1], p->stream);
    p->prof_n++;
}

extern "C" int pwt_profile_enable(pwt_plan* p, int on) {
    if (!p) return fail(PWT_ERR_ARG, "null plan");
    cudaSetDevice(p->device);
    if (on && !p->prof_ev) {
        p->prof_ev = (cudaEvent_t*)calloc(2 * PWT_PROF_CAP, sizeof(cudaEvent_t));
        if (!p->prof_ev) return fail(PWT_ERR_NOMEM, "out of host memory");
        for (int i = 0; i < 2 * PWT_PROF_CAP; i++) CK(cudaEventCreate(&p->prof_ev[i]));
    }
    p->prof_on = on ? 1 : 0;
    p->prof_n = 0;
    return PWT_OK;
}

extern "C" int pwt_profile_read(pwt_plan* p, float* ms, int* tags, int cap) {
    if (!p || !ms || !tags) return fail(PWT_ERR_ARG, "null argument");
    cudaSetDevice(p->device);
    CK(cudaStreamSynchronize(p->stream));
    int n = p->prof_n < cap ? p->prof_n : cap;
    for (int i = 0; i < n; i++) {
        CK(cudaEventElapsedTime(&ms[i], p->prof_ev[2 * i], p->prof_ev[2 * i + 1]));
        tags[i] = p->prof_tag[i];
    }
    p->prof_n = 0;
    return n;
}

// ---- forward ----------------------------------------------------------------------------------
// Destination of the level-l approximation so that A_L ends in band 0 without a fix-up copy
// (the reference ping-pongs and memcpy's back for even level counts, e.g. haar.cu:83).
static inline float* approx_dst(pwt_plan* p, int l, float* alt) {
    return ((p->nlevels - l) & 1) ? alt : p->d_band[0];
}

extern "C" int pwt_forward(pwt_plan* p) {
    if (!p) return fail(PWT_ERR_ARG, "null plan");
    if (p->state == PWT_CREATION_ERROR) return fail(PWT_ERR_STATE, "plan is in creation-error state");
    cudaSetDevice(p->device);
    p->pend.op = -1;                                                // the coefficients are about to be overwritten
    clear_partials(p);
    if (p->do_cs) {                                                 // wt.cu:242-246
        p->shift_r = rand() % p->Nr;
        p->shift_c = rand() % p->Nc;
        if (lazy_cs(p)) {                                           // recorded: the bands come out shifted by the same amount
            p->vs_img_r = modn(p->vs_img_r + p->shift_r, p->Nr);
            p->vs_img_c = modn(p->vs_img_c + p->shift_c, p->Nc);
            p->vs_coef_r = p->vs_img_r;
            p->vs_coef_c = p->vs_img_c;
        } else {
            int rc = do_circshift(p, p->shift_r, p->shift_c, 1);
            if (rc != PWT_OK) return rc;
        }
    }
    const int L = p->nlevels, B = p->batch;
    const bool haar = is_haar(p);
    // rank-1 non-separable filters: same numbers as the separable transform (up to fp32 rounding of the tap
    // products), F^2 -> 2F multiply-adds per sample; the reference's slot order differs (Q1: bands 1 and 2 swapped)
    const bool ns_sep = !p->do_separable && !haar && p->ns_rank1 && p->kernel_mode != 1;
    const bool sep = haar || p->do_separable || ns_sep;
    const int sH = ns_sep ? 2 : 1, sV = ns_sep ? 1 : 2;            // band slots that receive the separable H and V
    cudaStream_t st = p->stream;
    const float* src = p->d_image;
    if (p->ndims == 1) {
        const int rows = B * p->Nr;
        int l_first = 1;
        // every level in one launch, the rows staged once in shared memory (kernels_row1d.cu)
        if (p->kernel_mode == 0 && !pwt_tuning().no_fused1d) {
            float* Ds[PWT_MAX_LEVELS];
            for (int l = 1; l <= L; l++) Ds[l - 1] = p->d_band[l];
            prof_begin(p, 100 * L + 21);
            const int n = p->do_swt ? pwt_row_swt_fwd1d_all(src, p->d_band[0], Ds, rows, p->Nc, L, p->filt, st)
                                    : pwt_row_dwt_fwd1d_all(src, p->d_band[0], Ds, rows, p->Nc, L, p->filt, st);
            if (n) {
                prof_end(p);
                p->launches += n;
                l_first = L + 1;
            }
        }
        for (int l = l_first; l <= L; l++) {
            float* dstA = approx_dst(p, l, p->d_tmp);
            prof_begin(p, 100 * l + 1);
            if (p->do_swt) {
                int n = p->kernel_mode == 1 ? 0 : pwt_fast_swt_fwd1d(src, dstA, p->d_band[l], rows, p->Nc, l, p->filt, st);
                if (!n) n = pwt_launch_swt_fwd1d(src, dstA, p->d_band[l], rows, p->Nc, l, p->filt, st);
                p->launches += n;
            }
            else {
                int n = 0;
                if (!haar && p->kernel_mode == 0 && rows <= 32 && p->lvNc[l - 1] >= 8192)     // few long rows: tiles along the row
                    n = pwt_rows1d_fwd_f32(src, dstA, p->d_band[l], rows, p->lvNc[l - 1], p->filt, st);
                if (!n && !haar && ((p->kernel_mode == 0 && p->lvNc[l - 1] >= 256) || p->kernel_mode == 4))
                    n = pwt_strip_dwt_fwd1d(src, dstA, p->d_band[l], rows, p->lvNc[l - 1], p->filt, st);
                if (haar && p->kernel_mode != 1) n = pwt_haar_fwd1d_flat(src, dstA, p->d_band[l], rows, p->lvNc[l - 1], st);
                if (!n) n = pwt_launch_dwt_fwd1d(src, dstA, p->d_band[l], rows, p->lvNc[l - 1], p->filt, haar, st);
                p->launches += n;
            }
            prof_end(p);
            src = dstA;
        }
    } else {
        const long long plane = (long long)B * img_elems(p);
        int l_first = 1;
        // levels 1..3 in one launch when the fused register cascade covers the configuration
        if (!p->do_swt && L >= 3 && sep && p->kernel_mode == 0 && (haar || p->hlen < p->strip_min_f)) {
            float* Hs[3] = {p->d_band[sH], p->d_band[3 + sH], p->d_band[6 + sH]};
            float* Vs[3] = {p->d_band[sV], p->d_band[3 + sV], p->d_band[6 + sV]};
            float* Ds[3] = {p->d_band[3], p->d_band[6], p->d_band[9]};
            float* dstA = approx_dst(p, 3, p->d_tmp);
            prof_begin(p, 100 * 3 + 1 + 10);      // tag x1y: fused levels 1..3
            if (p->queue.base > 0x70000000u) {   // far from wrapping: re-arm the queue
                cudaMemsetAsync(p->queue.counter, 0, sizeof(unsigned), st);
                p->queue.base = 0;
            }
            int ntasks = 0;
            const int n = pwt_fused_dwt_fwd3(src, dstA, Hs, Vs, Ds, B, p->Nr, p->Nc, p->filt, haar, &p->queue,
                                             p->want_norms ? p->d_partials : nullptr, p->partials_cap, L == 3, &ntasks, st);
            if (n) {
                prof_end(p);
                p->launches += n;
                src = dstA;
                l_first = 4;
                p->partials_n = ntasks;
                if (ntasks > 0) { p->norm_lvl_mask = 7u; p->norm_a = (L == 3); }
            }
        }
        for (int l = l_first; l <= L; l++) {
            float* Hb = p->d_band[3 * (l - 1) + sH];
            float* V = p->d_band[3 * (l - 1) + sV];
            float* D = p->d_band[3 * (l - 1) + 3];
            prof_begin(p, 100 * l + 1);
            if (p->do_swt) {
                float* dstA = approx_dst(p, l, p->d_tmp + 2 * plane);
                if (p->do_separable || ns_sep) {
                    int n = p->kernel_mode == 0 ? pwt_strip_swt_fwd2d(src, dstA, Hb, V, D, B, p->Nr, p->Nc, l, p->filt, st) : 0;
                    if (!n && p->kernel_mode != 1) n = pwt_fast_swt_fwd2d(src, dstA, Hb, V, D, B, p->Nr, p->Nc, l, p->filt, st);
                    if (!n && p->kernel_mode != 1) n = pwt_swt2p_fwd2d(src, dstA, Hb, V, D, p->d_tmp, B, p->Nr, p->Nc, l, p->filt, st);
                    if (!n) n = pwt_launch_swt_fwd2d(src, dstA, Hb, V, D, p->d_tmp, B, p->Nr, p->Nc, l, p->filt, st);
                    p->launches += n;
                }
                else
                    p->launches += pwt_launch_ns_swt_fwd2d(src, dstA, Hb, V, D, B, p->Nr, p->Nc, l, p->d_k2d_fwd, p->hlen, st);
                src = dstA;
            } else {
                float* dstA = approx_dst(p, l, p->d_tmp);
                const long long in_bs = lvl_elems(p, l - 1), out_bs = lvl_elems(p, l);
                const int nr = p->lvNr[l - 1], nc = p->lvNc[l - 1];
                if (sep) {
                    int n = 0;
                    const int hints = (l < L ? PWT_HINT_OUT_FEEDS_NEXT : 0) | (l > 1 ? PWT_HINT_IN_FROM_PREV : 0);
                    // F >= 8 always; shorter filters on the small planes that follow the fused cascade (levels >= 4: the
                    // register kernels need 9-11 us per launch there, the strip kernels 3-4)
                    // ... and on widths that are not multiples of 4, where the register kernels do not apply (1001 x 777 db2: 2.4x)
                    const bool small = l >= 4 || (p->kernel_mode == 0 && (long long)B * nr * nc <= tail_plane_px(p->hlen, (long long)B * p->Nr * p->Nc));
                    const bool tail = ((small && (p->kernel_mode == 0 || p->kernel_mode == 3)) || ((nc & 3) && p->kernel_mode == 0)) &&
                                      p->hlen >= 4 && pwt_tuning().tail_strip;
                    // thin and wide planes (fewer rows than the register kernels and the cascade take, but a lot of samples): the strip
                    // kernels walk their few rows at streaming speed (16 x 1 M db2 2 levels fwd+inv: 0.62 ms on the tile kernels)
                    const bool thin = p->kernel_mode == 0 && p->hlen >= 4 && nr >= 8 && nr < 64 && nc >= 256 && (long long)nr * nc >= (1LL << 20);
                    if (!haar && ((((p->kernel_mode == 0 && p->hlen >= p->strip_min_f) || tail) && nr >= 64 && nc >= 256) || thin || p->kernel_mode == 4)) {
                        // norms requested after an earlier forward: the strip kernel reduces |c|, c^2 of what it stores
                        const bool nrm = p->want_norms && p->d_partials && p->do_separable && p->kernel_mode == 0 && l <= 32;
                        int wr = 0;
                        n = pwt_strip_dwt_fwd2d_norms(src, dstA, Hb, V, D, B, nr, nc, in_bs, out_bs, p->filt,
                                                      nrm ? p->d_partials + 2 * (size_t)p->partials_n : nullptr,
                                                      p->partials_cap - p->partials_n, l == L, &wr, st);
                        if (n && wr > 0) {
                            p->partials_n += wr;
                            p->norm_lvl_mask |= 1u << (l - 1);
                            if (l == L) p->norm_a = 1;
                        }
                    }
                    if (!n && haar && p->kernel_mode == 0 && (long long)nr * nc <= haar_flat_px())
                        n = pwt_haar2d_fwd_flat(src, dstA, Hb, V, D, B, nr, nc, in_bs, out_bs, st);
                    if (!n && (p->kernel_mode == 0 || p->kernel_mode == 3))
                        n = pwt_reg_dwt_fwd2d(src, dstA, Hb, V, D, B, nr, nc, in_bs, out_bs, p->filt, haar, hints, st);
                    if (!n && haar && (p->kernel_mode == 0 || p->kernel_mode == 3)) n = pwt_haar2d_fwd_flat(src, dstA, Hb, V, D, B, nr, nc, in_bs, out_bs, st);
                    if (!n && p->kernel_mode != 1 && !haar && p->hlen >= p->tile_min_f && nr >= 64 && nc >= 64)
                        n = pwt_tile_dwt_fwd2d(src, dstA, Hb, V, D, B, nr, nc, in_bs, out_bs, p->filt, st);
                    if (!n && p->kernel_mode != 1)
                        n = pwt_fast_dwt_fwd2d(src, dstA, Hb, V, D, B, nr, nc, in_bs, out_bs, p->filt, haar,
                                               (l < L ? PWT_HINT_OUT_FEEDS_NEXT : 0) | (l > 1 ? PWT_HINT_IN_FROM_PREV : 0), st);
                    if (!n) n = pwt_launch_dwt_fwd2d(src, dstA, Hb, V, D, B, nr, nc, in_bs, out_bs, p->filt, haar, st);
                    p->launches += n;
                } else {
                    p->launches += pwt_launch_ns_fwd2d(src, dstA, Hb, V, D, B, nr, nc, in_bs, out_bs, p->d_k2d_fwd, p->hlen, st);
                }
                src = dstA;
            }
            prof_end(p);
        }
    }
    CK_LAUNCH();
    p->state = PWT_FORWARD;
    return PWT_OK;
}

// One separable 2D level over a stack of planes with the automatic kernel choice of pwt_forward / pwt_inverse (strip
// kernels for F >= 8 on large planes, register kernels, tile / shared-memory kernels, generic).  Used by the volumetric
// plans (pwt_vol.cu), whose x-y passes are batched 2D levels.  Returns the number of launches.
int pwt_level_fwd2d(const float* src, float* A, float* Hb, float* V, float* D, int batch, int nr, int nc, long long in_bs,
                    long long out_bs, const PwtFilters& f, bool haar, cudaStream_t st) {
    const PwtTuning& k = pwt_tuning();
    int n = 0;
    if (!haar && f.hlen >= k.strip_min_f && nr >= 64 && nc >= 256)
        n = pwt_strip_dwt_fwd2d(src, A, Hb, V, D, batch, nr, nc, in_bs, out_bs, f, st);
    if (!n) n = pwt_reg_dwt_fwd2d(src, A, Hb, V, D, batch, nr, nc, in_bs, out_bs, f, haar, 0, st);
    if (!n && haar) n = pwt_haar2d_fwd_flat(src, A, Hb, V, D, batch, nr, nc, in_bs, out_bs, st);
    if (!n && !haar && f.hlen >= k.tile_min_f && nr >= 64 && nc >= 64)
        n = pwt_tile_dwt_fwd2d(src, A, Hb, V, D, batch, nr, nc, in_bs, out_bs, f, st);
    if (!n) n = pwt_fast_dwt_fwd2d(src, A, Hb, V, D, batch, nr, nc, in_bs, out_bs, f, haar, 0, st);
    if (!n) n = pwt_launch_dwt_fwd2d(src, A, Hb, V, D, batch, nr, nc, in_bs, out_bs, f, haar, st);
    return n;
}
int pwt_level_inv2d(const float* A, const float* Hb, const float* V, const float* D, float* dst, int batch, int nr, int nc,
                    int Nro, int Nco, long long in_bs, long long out_bs, const PwtFilters& f, bool haar, cudaStream_t st) {
    const PwtTuning& k = pwt_tuning();
    int n = 0;
    if (!haar && f.hlen >= k.strip_min_f && nr >= 32 && nc >= 128)
        n = pwt_strip_dwt_inv2d(A, Hb, V, D, dst, batch, nr, nc, Nro, Nco, in_bs, out_bs, f, st);
    if (!n) n = pwt_reg_dwt_inv2d(A, Hb, V, D, dst, batch, nr, nc, Nro, Nco, in_bs, out_bs, f, haar, 0, st);
    if (!n && haar) n = pwt_haar2d_inv_flat(A, Hb, V, D, dst, batch, nr, nc, Nro, Nco, in_bs, out_bs, st);
    if (!n && !haar && f.hlen >= k.tile_min_f && nr >= 32 && nc >= 32)
        n = pwt_tile_dwt_inv2d(A, Hb, V, D, dst, batch, nr, nc, Nro, Nco, in_bs, out_bs, f, st);
    if (!n) n = pwt_fast_dwt_inv2d(A, Hb, V, D, dst, batch, nr, nc, Nro, Nco, in_bs, out_bs, f, haar, 0, st);
    if (!n) n = pwt_launch_dwt_inv2d(A, Hb, V, D, dst, batch, nr, nc, Nro, Nco, in_bs, out_bs, f, haar, st);
    return n;
}

// ---- inverse ----------------------------------------------------------------------------------
extern "C" int pwt_inverse(pwt_plan* p) {
    if (!p) return fail(PWT_ERR_ARG, "null plan");
    if (p->state == PWT_INVERSE) {                                  // wt.cu:272-275
        puts("Warning: W.inverse() has already been run. Inverse is available in W.get_image()");
        return 1;
    }
    if (p->state == PWT_CREATION_ERROR) return fail(PWT_ERR_STATE, "plan is in creation-error state");
    cudaSetDevice(p->device);
    const int L = p->nlevels, B = p->batch;
    const bool haar = is_haar(p);
    const bool ns_sep = !p->do_separable && !haar && p->ns_rank1 && p->kernel_mode != 1;     // see pwt_forward
    const bool sep = haar || p->do_separable || ns_sep;
    const int sH = ns_sep ? 2 : 1, sV = ns_sep ? 1 : 2;
    cudaStream_t st = p->stream;
    const float* cur = p->d_band[0];
    if (p->ndims == 1) {
        const int rows = B * p->Nr;
        int l_top = L;
        if (p->kernel_mode == 0 && !pwt_tuning().no_fused1d) {                    // every level in one launch
            float* Ds[PWT_MAX_LEVELS];
            for (int l = 1; l <= L; l++) Ds[l - 1] = p->d_band[l];
            prof_begin(p, 100 * L + 22);
            const int n = p->do_swt ? pwt_row_swt_inv1d_all(cur, Ds, p->d_image, rows, p->Nc, L, p->filt, st)
                                    : pwt_row_dwt_inv1d_all(cur, Ds, p->d_image, rows, p->Nc, L, p->filt, st);
            if (n) {
                prof_end(p);
                p->launches += n;
                l_top = 0;
            }
        }
        for (int l = l_top; l >= 1; l--) {
            float* dst = (l == 1) ? p->d_image : (cur == p->d_band[0] ? p->d_tmp : p->d_band[0]);
            prof_begin(p, 100 * l + 2);
            if (p->do_swt) {
                int n = p->kernel_mode == 1 ? 0 : pwt_fast_swt_inv1d(cur, p->d_band[l], dst, rows, p->Nc, l, p->filt, st);
                if (!n) n = pwt_launch_swt_inv1d(cur, p->d_band[l], dst, rows, p->Nc, l, p->filt, st);
                p->launches += n;
            }
            else {
                int n = 0;
                if (!haar && p->kernel_mode == 0 && rows <= 32 && p->lvNc[l - 1] >= 8192)
                    n = pwt_rows1d_inv_f32(cur, p->d_band[l], dst, rows, p->lvNc[l], p->lvNc[l - 1], p->filt, st);
                if (!n && !haar && ((p->kernel_mode == 0 && p->lvNc[l] >= 128) || p->kernel_mode == 4))
                    n = pwt_strip_dwt_inv1d(cur, p->d_band[l], dst, rows, p->lvNc[l], p->lvNc[l - 1], p->filt, st);
                if (haar && p->kernel_mode != 1) n = pwt_haar_inv1d_flat(cur, p->d_band[l], dst, rows, p->lvNc[l], p->lvNc[l - 1], st);
                if (!n) n = pwt_launch_dwt_inv1d(cur, p->d_band[l], dst, rows, p->lvNc[l], p->lvNc[l - 1], p->filt, haar, st);
                p->launches += n;
            }
            prof_end(p);
            cur = dst;
        }
    } else {
        const long long plane = (long long)B * img_elems(p);
        PwtDeferredOp fop = p->pend;                                // what the fused launch applies while loading
        const bool swt_defer = p->pend.op >= 0 && p->defer_swt_ok && p->kernel_mode == 0 && swt_all_levels_fused(p);
        if (p->pend.op >= 0 && p->do_swt && !swt_defer) {
            int rc = flush_pending(p, 1, true);
            if (rc != PWT_OK) return rc;
        }
        // strip inverse (F >= 8): every level it serves applies its own part of a pending threshold while staging
        // the bands; the coarser levels it does not serve (planes < 32 x 128) and their A go through memory first
        bool strip_defer = false;
        int strip_lmax = 0;
        const bool cascade_ok = sep && p->kernel_mode == 0 && (haar || p->hlen < p->strip_min_f);
        if (p->pend.op >= 0 && !p->do_swt) {
            if (strip_defer_capable(p) && !haar) {
                for (int l = 1; l <= L && p->lvNr[l] >= 32 && p->lvNc[l] >= 128; l++) strip_lmax = l;
                strip_defer = strip_lmax >= 1;
                if (strip_defer && strip_lmax < L) {
                    int rc = flush_pending(p, strip_lmax + 1, true);
                    if (rc != PWT_OK) return rc;
                }
            } else if (!(cascade_ok && p->defer_ok && L >= 3)) {     // nobody will consume it on load
                int rc = flush_pending(p, 1, true);
                if (rc != PWT_OK) return rc;
            }
        }
        if (p->pend.op >= 0 && !p->do_swt && L > 3 && !strip_defer) {
            int rc = flush_pending(p, 4, true);                     // coarser levels + A go through memory (1/64 of the data)
            if (rc != PWT_OK) return rc;
            fop.app = 0;
        }
        for (int l = L; l >= 1; l--) {
            if (l == 3 && p->pend.op >= 0 && !p->do_swt && !strip_defer && !cascade_ok) {
                int rc = flush_pending(p, 1, L == 3);
                if (rc != PWT_OK) return rc;
            }
            // levels 3..1 in one launch when the fused register cascade covers the configuration
            if (l == 3 && !p->do_swt && sep && p->kernel_mode == 0 && (haar || p->hlen < p->strip_min_f)) {
                const float* Hs[3] = {p->d_band[sH], p->d_band[3 + sH], p->d_band[6 + sH]};
                const float* Vs[3] = {p->d_band[sV], p->d_band[3 + sV], p->d_band[6 + sV]};
                const float* Ds[3] = {p->d_band[3], p->d_band[6], p->d_band[9]};
                if (p->queue.base > 0x70000000u) {
                    cudaMemsetAsync(p->queue.counter, 0, sizeof(unsigned), st);
                    p->queue.base = 0;
                }
                prof_begin(p, 100 * 3 + 2 + 10);      // tag 312: fused levels 3..1
                // F = 6 behind a strip launch of level 4: launched plainly -- started early by programmatic dependent launch,
                // its persistent CTAs get placed unevenly next to the draining strip kernel (8192^2 db3 L5: 0.346 vs 0.255 ms)
                const int plain = L > 3 && p->hlen == 6 && pwt_tuning().tail_strip;
                const int n = pwt_fused_dwt_inv3(cur, Hs, Vs, Ds, p->d_image, B, p->Nr, p->Nc, p->filt, haar, &p->queue,
                                                 p->pend.op >= 0 ? &fop : nullptr, plain, st);
                if (n) {
                    prof_end(p);
                    p->launches += n;
                    cur = p->d_image;
                    p->pend.op = -1;
                    break;
                }
                if (p->pend.op >= 0) {                              // not covered after all: apply through memory
                    int rc = flush_pending(p, 1, L == 3);
                    if (rc != PWT_OK) return rc;
                }
            }
            const float* Hb = p->d_band[3 * (l - 1) + sH];
            const float* V = p->d_band[3 * (l - 1) + sV];
            const float* D = p->d_band[3 * (l - 1) + 3];
            prof_begin(p, 100 * l + 2);
            if (p->do_swt) {
                float* alt = p->d_tmp + 2 * plane;
                float* dst = (l == 1) ? p->d_image : (cur == p->d_band[0] ? alt : p->d_band[0]);
                if (p->do_separable || ns_sep) {
                    int n = 0;
                    if (p->kernel_mode == 0) {
                        if (swt_defer)
                            n = pwt_strip_swt_inv2d(cur, Hb, V, D, dst, B, p->Nr, p->Nc, l, p->filt, fop.op, fop.beta[l - 1],
                                                    l == L && fop.app, fop.beta_app, st);
                        else
                            n = pwt_strip_swt_inv2d(cur, Hb, V, D, dst, B, p->Nr, p->Nc, l, p->filt, -1, 0.f, 0, 0.f, st);
                    }
                    if (!n && p->kernel_mode != 1) {
                        if (swt_defer)      // the deferred threshold of this level is applied while its bands are loaded
                            n = pwt_fast_swt_inv2d(cur, Hb, V, D, dst, B, p->Nr, p->Nc, l, p->filt, fop.op, fop.beta[l - 1],
                                                   l == L && fop.app, fop.beta_app, st);
                        else
                            n = pwt_fast_swt_inv2d(cur, Hb, V, D, dst, B, p->Nr, p->Nc, l, p->filt, -1, 0.f, 0, 0.f, st);
                    }
                    if (!n && swt_defer) return fail(PWT_ERR_CUDA, "fused SWT inverse declined a level it had accepted");
                    if (!n && p->kernel_mode != 1) n = pwt_swt2p_inv2d(cur, Hb, V, D, dst, p->d_tmp, B, p->Nr, p->Nc, l, p->filt, st);
                    if (!n) n = pwt_launch_swt_inv2d(cur, Hb, V, D, dst, p->d_tmp, B, p->Nr, p->Nc, l, p->filt, st);
                    p->launches += n;
                }
                else
                    p->launches += pwt_launch_ns_swt_inv2d(cur, Hb, V, D, dst, B, p->Nr, p->Nc, l, p->d_k2d_inv, p->hlen, st);
                cur = dst;
            } else {
                float* dst = (l == 1) ? p->d_image : (cur == p->d_band[0] ? p->d_tmp : p->d_band[0]);
                const long long in_bs = lvl_elems(p, l), out_bs = lvl_elems(p, l - 1);
                const int nr = p->lvNr[l], nc = p->lvNc[l], Nro = p->lvNr[l - 1], Nco = p->lvNc[l - 1];
                if (sep) {
                    int n = 0;
                    const int hints = (l > 1 ? PWT_HINT_OUT_FEEDS_NEXT : 0) | (l < L ? PWT_HINT_IN_FROM_PREV : 0);
                    const bool small = l >= 4 || (p->kernel_mode == 0 && (long long)B * Nro * Nco <= tail_plane_px(p->hlen, (long long)B * p->Nr * p->Nc));
                    const bool tail = ((small && (p->kernel_mode == 0 || p->kernel_mode == 3)) || ((Nco & 3) && p->kernel_mode == 0)) &&
                                      p->hlen >= 4 && pwt_tuning().tail_strip;
                    const bool thin = p->kernel_mode == 0 && p->hlen >= 4 && Nro >= 8 && Nro < 64 && Nco >= 256 && (long long)Nro * Nco >= (1LL << 20);
                    if (!haar && ((((p->kernel_mode == 0 && p->hlen >= p->strip_min_f) || tail) && nr >= 32 && nc >= 128) || thin || p->kernel_mode == 4)) {
                        if (strip_defer && l <= strip_lmax)
                            n = pwt_strip_dwt_inv2d_thr(cur, Hb, V, D, dst, B, nr, nc, Nro, Nco, in_bs, out_bs, p->filt, p->pend.op,
                                                        p->pend.beta[l - 1], l == L && p->pend.app, p->pend.beta_app, st);
                        else
                            n = pwt_strip_dwt_inv2d(cur, Hb, V, D, dst, B, nr, nc, Nro, Nco, in_bs, out_bs, p->filt, st);
                    }
                    if (!n && strip_defer && l <= strip_lmax) {      // declined after all: apply the rest through memory
                        int rc = flush_pending(p, 1, l == L);
                        if (rc != PWT_OK) return rc;
                        strip_defer = false;
                    }
                    if (!n && haar && p->kernel_mode == 0 && (long long)Nro * Nco <= haar_flat_px())
                        n = pwt_haar2d_inv_flat(cur, Hb, V, D, dst, B, nr, nc, Nro, Nco, in_bs, out_bs, st);
                    if (!n && (p->kernel_mode == 0 || p->kernel_mode == 3))
                        n = pwt_reg_dwt_inv2d(cur, Hb, V, D, dst, B, nr, nc, Nro, Nco, in_bs, out_bs, p->filt, haar, hints, st);
                    if (!n && haar && (p->kernel_mode == 0 || p->kernel_mode == 3)) n = pwt_haar2d_inv_flat(cur, Hb, V, D, dst, B, nr, nc, Nro, Nco, in_bs, out_bs, st);
                    if (!n && p->kernel_mode != 1 && !haar && p->hlen >= p->tile_min_f && nr >= 32 && nc >= 32)
                        n = pwt_tile_dwt_inv2d(cur, Hb, V, D, dst, B, nr, nc, Nro, Nco, in_bs, out_bs, p->filt, st);
                    if (!n && p->kernel_mode != 1)
                        n = pwt_fast_dwt_inv2d(cur, Hb, V, D, dst, B, nr, nc, Nro, Nco, in_bs, out_bs, p->filt, haar,
                                               (l > 1 ? PWT_HINT_OUT_FEEDS_NEXT : 0) | (l < L ? PWT_HINT_IN_FROM_PREV : 0), st);
                    if (!n) n = pwt_launch_dwt_inv2d(cur, Hb, V, D, dst, B, nr, nc, Nro, Nco, in_bs, out_bs, p->filt, haar, st);
                    p->launches += n;
                } else {
                    p->launches += pwt_launch_ns_inv2d(cur, Hb, V, D, dst, B, nr, nc, Nro, Nco, in_bs, out_bs, p->d_k2d_inv, p->hlen, st);
                }
                cur = dst;
            }
            prof_end(p);
        }
        if (strip_defer) p->pend.op = -1;                           // consumed by the strip inverse launches
    }
    CK_LAUNCH();
    if (p->do_swt) p->pend.op = -1;                                 // consumed by the fused SWT inverse (or flushed above)
    if (p->do_cs) {                                                 // wt.cu:303
        if (lazy_cs(p)) {                                           // normally (0, 0): the two shifts cancel
            p->vs_img_r = modn(p->vs_coef_r - p->shift_r, p->Nr);
            p->vs_img_c = modn(p->vs_coef_c - p->shift_c, p->Nc);
        } else {
            int rc = do_circshift(p, -p->shift_r, -p->shift_c, 1);
            if (rc != PWT_OK) return rc;
        }
    }
    clear_partials(p);
    p->state = PWT_INVERSE;
    return PWT_OK;
}

// ---- thresholds / shrink ----------------------------------------------------------------------
static const double kSqrt2 = 1.4142135623730951;   // common.cu:8

static void add_seg(PwtSegTable* t, float* ptr, long long n, float beta) {
    if (n <= 0 || t->nseg >= PWT_MAX_SEGS) return;
    t->seg[t->nseg].ptr = ptr;
    t->seg[t->nseg].n = n;
    t->seg[t->nseg].beta = beta;
    t->seg[t->nseg].pad = 0;
    t->nseg++;
}

// Per-band parameter schedule of the reference callers (common.cu:219-282, 347-371):
//  - details of level i+1 use beta / sqrt(2)^(i+1), computed by cumulative fp32 division by the
//    double constant SQRT_2 when normalize > 0;
//  - the approximation uses beta / sqrt(2)^L for the soft threshold, but the UNscaled beta for the
//    hard threshold (common.cu:264-270 computes beta2 and then passes beta) -- replicated.
static void build_thresh_table(pwt_plan* p, PwtSegTable* t, float beta, int app, int normalize,
                               bool app_scaled, bool const_param, float cparam) {
    t->nseg = 0;
    const long long B = p->batch;
    const int L = p->nlevels;
    if (app) {
        float beta2 = beta;
        if (normalize > 0 && app_scaled) {
            const int n2 = L / 2;
            beta2 /= (float)(1 << n2);
            if (n2 * 2 != L) beta2 = (float)(beta2 / kSqrt2);
        }
        add_seg(t, p->d_band[0], B * band_elems(p, 0), const_param ? cparam : beta2);
    }
    for (int i = 0; i < L; i++) {
        if (normalize > 0) beta = (float)(beta / kSqrt2);
        const float b = const_param ? cparam : beta;
        if (p->ndims == 2) {
            for (int j = 1; j <= 3; j++) add_seg(t, p->d_band[3 * i + j], B * band_elems(p, 3 * i + j), b);
        } else {
            add_seg(t, p->d_band[i + 1], B * band_elems(p, i + 1), b);
        }
    }
}

static int run_thresh(pwt_plan* p, int op, float beta, int app, int normalize, bool app_scaled,
                      const char* what) {
    if (!p) return fail(PWT_ERR_ARG, "null plan");
    if (p->state == PWT_INVERSE) {                                  // wt.cu:309-312
        printf("Warning: Wavelets(): cannot %s coefficients, as they were modified by W.inverse()\n", what);
        return 1;
    }
    cudaSetDevice(p->device);
    int rc = flush_pending(p, 1, true);                             // an older pending threshold comes first
    if (rc != PWT_OK) return rc;
    PwtSegTable t;
    if (op == PWT_OP_SCALE)
        build_thresh_table(p, &t, beta, app, 0, false, true, 1.0f / (1.0f + beta));   // common.cu:355
    else
        build_thresh_table(p, &t, beta, app, normalize, app_scaled, false, 0.f);
    if ((op == PWT_OP_SOFT || op == PWT_OP_HARD) && (p->defer_ok || p->defer_swt_ok || strip_defer_capable(p)) && p->kernel_mode == 0) {
        // record instead of launching: the fused inverse applies it on load; any observer of the
        // coefficients (coeffs, norms, pointers, another operator) flushes it to memory first
        p->pend.op = op;
        p->pend.app = app ? 1 : 0;
        int k = 0;
        p->pend.beta_app = app ? t.seg[k++].beta : 0.f;
        for (int i = 0; i < p->nlevels; i++, k += 3) p->pend.beta[i] = t.seg[k].beta;
        return PWT_OK;
    }
    clear_partials(p);
    p->launches += pwt_launch_eltwise(t, op, p->stream);
    CK_LAUNCH();
    return PWT_OK;
}

static int flush_pending(pwt_plan* p, int first_level, bool with_app) {
    if (p->pend.op < 0) return PWT_OK;
    PwtSegTable t;
    t.nseg = 0;
    const long long B = p->batch;
    if (with_app && p->pend.app) add_seg(&t, p->d_band[0], B * band_elems(p, 0), p->pend.beta_app);
    for (int i = first_level - 1; i < p->nlevels; i++)
        for (int j = 1; j <= 3; j++) add_seg(&t, p->d_band[3 * i + j], B * band_elems(p, 3 * i + j), p->pend.beta[i]);
    clear_partials(p);
    p->launches += pwt_launch_eltwise(t, p->pend.op, p->stream);
    CK_LAUNCH();
    if (first_level <= 1) p->pend.op = -1;
    return PWT_OK;
}

extern "C" int pwt_soft_threshold(pwt_plan* p, float beta, int app, int normalize) {
    return run_thresh(p, PWT_OP_SOFT, beta, app, normalize, true, "threshold");
}
extern "C" int pwt_hard_threshold(pwt_plan* p, float beta, int app, int normalize) {
    return run_thresh(p, PWT_OP_HARD, beta, app, normalize, false, "threshold");
}
extern "C" int pwt_proj_linf(pwt_plan* p, float beta, int app) {
    return run_thresh(p, PWT_OP_PROJ, beta, app, 0, false, "project");
}
extern "C" int pwt_shrink(pwt_plan* p, float beta, int app) {
    return run_thresh(p, PWT_OP_SCALE, beta, app, 0, false, "shrink");
}

extern "C" int pwt_group_soft_threshold(pwt_plan* p, float beta, int app, int normalize) {
    // common.cu:311-341: A joins the group of the LAST level only
    if (!p) return fail(PWT_ERR_ARG, "null plan");
    if (p->state == PWT_INVERSE) {
        puts("Warning: Wavelets(): cannot threshold coefficients, as they were modified by W.inverse()");
        return 1;
    }
    cudaSetDevice(p->device);
    {
        int rc0 = flush_pending(p, 1, true);
        if (rc0 != PWT_OK) return rc0;
    }
    clear_partials(p);
    const int L = p->nlevels;
    for (int i = 0; i < L; i++) {
        if (normalize > 0) beta = (float)(beta / kSqrt2);
        float* a = (app && i == L - 1) ? p->d_band[0] : nullptr;
        if (p->ndims == 2) {
            const long long n = (long long)p->batch * band_elems(p, 3 * i + 1);
            p->launches += pwt_launch_group_soft(p->d_band[3 * i + 1], p->d_band[3 * i + 2], p->d_band[3 * i + 3], a, n, beta, p->stream);
        } else {
            const long long n = (long long)p->batch * band_elems(p, i + 1);
            p->launches += pwt_launch_group_soft(nullptr, nullptr, p->d_band[i + 1], a, n, beta, p->stream);
        }
    }
    CK_LAUNCH();
    return PWT_OK;
}

// ---- norms ------------------------------------------------------------------------------------
static int local_norms_async(pwt_plan* p) {
    int rc0 = flush_pending(p, 1, true);
    if (rc0 != PWT_OK) return rc0;
    PwtSegTable t;
    t.nseg = 0;
    // bands 1..9 (and A of a 3-level transform) were already reduced inside the fused forward kernel
    const bool fusedn = p->partials_n > 0 && p->state == PWT_FORWARD;
    if (p->state == PWT_FORWARD) p->want_norms = 1;
    for (int b = 0; b < p->nbands; b++) {
        // 2D DWT band b > 0 belongs to level (b-1)/3 + 1; bands the forward kernels reduced themselves are skipped
        const bool covered = fusedn && (b == 0 ? p->norm_a != 0 : (((b - 1) / 3) < 32 && ((p->norm_lvl_mask >> ((b - 1) / 3)) & 1u)));
        if (!covered) add_seg(&t, p->d_band[b], (long long)p->batch * band_elems(p, b), 0.f);
    }
    if (t.nseg > 0) p->launches += pwt_launch_norms(t, p->d_acc, p->stream);
    if (fusedn) p->launches += pwt_launch_reduce_partials(p->d_partials, p->partials_n, p->d_acc, t.nseg > 0, p->stream);
    CK_LAUNCH();
    return PWT_OK;
}

extern "C" int pwt_norms(pwt_plan* p, double* n1, double* n2) {
    if (!p) return fail(PWT_ERR_ARG, "null plan");
    cudaSetDevice(p->device);
    int rc = local_norms_async(p);
    if (rc != PWT_OK) return rc;
    CK(cudaMemcpyAsync(p->h_acc, p->d_acc, 2 * sizeof(double), cudaMemcpyDeviceToHost, p->stream));
    CK(cudaStreamSynchronize(p->stream));
    if (n1) *n1 = p->h_acc[0];
    if (n2) *n2 = p->h_acc[1];
    return PWT_OK;
}
extern "C" int pwt_norm1(pwt_plan* p, float* out) {
    double a = 0, b = 0;
    int rc = pwt_norms(p, &a, &b);
    if (rc == PWT_OK && out) *out = (float)a;
    return rc;
}
extern "C" int pwt_norm2sq(pwt_plan* p, float* out) {
    double a = 0, b = 0;
    int rc = pwt_norms(p, &a, &b);
    if (rc == PWT_OK && out) *out = (float)b;
    return rc;
}

// ---- add_wavelet ------------------------------------------------------------------------------
extern "C" int pwt_add_wavelet(pwt_plan* d, const pwt_plan* s, float alpha) {
    if (!d || !s) return fail(PWT_ERR_ARG, "null plan");
    // wt.cu:625-650
    if (d->nlevels != s->nlevels || strcasecmp(d->wname, s->wname)) {
        puts("ERROR: add_wavelet(): right operand is not the same transform (wname, level)");
        return -1;
    }
    if (d->state == PWT_INVERSE || s->state == PWT_INVERSE) {
        puts("WARNING: add_wavelet(): this operation makes no sense when wavelet has just been inverted");
        return 1;
    }
    if (d->Nr != s->Nr || d->Nc != s->Nc || d->ndims != s->ndims || d->batch != s->batch) {
        puts("ERROR: add_wavelet(): operands do not have the same geometry");
        return -2;
    }
    if ((d->do_swt != 0) != (s->do_swt != 0)) {
        puts("ERROR: add_wavelet(): operands should both use SWT or DWT");
        return -3;
    }
    if (d->do_cs && s->do_cs && (d->shift_r != s->shift_r || d->shift_c != s->shift_c)) {
        puts("ERROR: add_wavelet(): operands do not have the same current shift");
        return -4;
    }
    cudaSetDevice(d->device);
    if (flush_pending(d, 1, true) != PWT_OK || flush_pending(const_cast<pwt_plan*>(s), 1, true) != PWT_OK) return PWT_ERR_CUDA;
    if (materialize_coeffs(d) != PWT_OK) return PWT_ERR_CUDA;
    cudaSetDevice(s->device);
    if (materialize_coeffs(const_cast<pwt_plan*>(s)) != PWT_OK) return PWT_ERR_CUDA;
    cudaSetDevice(d->device);
    cudaStreamSynchronize(s->stream);   // the source's pending work must be visible
    PwtSegTable td, ts;
    td.nseg = ts.nseg = 0;
    for (int b = 0; b < d->nbands; b++) {
        const long long n = (long long)d->batch * band_elems(d, b);
        add_seg(&td, d->d_band[b], n, 0.f);
        add_seg(&ts, s->d_band[b], n, 0.f);
    }
    clear_partials(d);
    d->launches += pwt_launch_axpy(td, ts, alpha, d->stream);
    CK_LAUNCH();
    return 0;
}

// ---- data in / out ----------------------------------------------------------------------------
extern "C" int pwt_get_image(pwt_plan* p, float* dst) {
    if (!p || !dst) return 0;
    cudaSetDevice(p->device);
    const size_t n = (size_t)p->batch * img_elems(p);
    if (materialize_image(p) != PWT_OK) return 0;
    if (cudaMemcpyAsync(dst, p->d_image, n * sizeof(float), cudaMemcpyDeviceToHost, p->stream) != cudaSuccess ||
        cudaStreamSynchronize(p->stream) != cudaSuccess) {
        fail(PWT_ERR_CUDA, "get_image failed: %s", cudaGetErrorString(cudaGetLastError()));
        return 0;
    }
    return n > 0x7fffffff ? 0x7fffffff : (int)n;
}

extern "C" int pwt_set_image(pwt_plan* p, const float* img, int on_device) {
    if (!p || !img) return fail(PWT_ERR_ARG, "null argument");
    cudaSetDevice(p->device);
    const size_t n = (size_t)p->batch * img_elems(p);
    p->vs_img_r = p->vs_img_c = 0;                                  // a new image: nothing recorded against it
    CK(cudaMemcpyAsync(p->d_image, img, n * sizeof(float),
                       on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, p->stream));
    if (!on_device) CK(cudaStreamSynchronize(p->stream));   // the host buffer may be reused by the caller
    p->state = PWT_INIT;                                            // wt.cu:430
    return PWT_OK;
}

extern "C" int pwt_get_coeff(pwt_plan* p, float* dst, int num) {
    if (!p || !dst || num < 0 || num >= p->nbands) return 0;
    if (p->state == PWT_INVERSE) {                                  // wt.cu:474-477
        puts("Warning: get_coeff(): inverse() has been performed, the coefficients has been modified and do not make sense anymore.");
        return 0;
    }
    cudaSetDevice(p->device);
    if (flush_pending(p, 1, true) != PWT_OK || materialize_coeffs(p) != PWT_OK) return 0;
    const size_t n = (size_t)p->batch * band_elems(p, num);
    if (cudaMemcpyAsync(dst, p->d_band[num], n * sizeof(float), cudaMemcpyDeviceToHost, p->stream) != cudaSuccess ||
        cudaStreamSynchronize(p->stream) != cudaSuccess) {
        fail(PWT_ERR_CUDA, "get_coeff failed: %s", cudaGetErrorString(cudaGetLastError()));
        return 0;
    }
    return n > 0x7fffffff ? 0x7fffffff : (int)n;
}

extern "C" int pwt_set_coeff(pwt_plan* p, const float* src, int num, int on_device) {
    if (!p || !src || num < 0 || num >= p->nbands) return fail(PWT_ERR_ARG, "bad argument");
    cudaSetDevice(p->device);
    {
        int rc0 = flush_pending(p, 1, true);
        if (rc0 != PWT_OK) return rc0;
    }
    clear_partials(p);
    {
        int rc1 = materialize_coeffs(p);
        if (rc1 != PWT_OK) return rc1;
    }
    const size_t n = (size_t)p->batch * band_elems(p, num);
    CK(cudaMemcpyAsync(p->d_band[num], src, n * sizeof(float),
                       on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, p->stream));
    if (!on_device) CK(cudaStreamSynchronize(p->stream));
    return PWT_OK;                                                  // state untouched (wt.cu:465)
}

extern "C" intptr_t pwt_image_ptr(pwt_plan* p) {
    if (!p) return 0;
    if (p->vs_img_r | p->vs_img_c) {
        cudaSetDevice(p->device);
        materialize_image(p);
    }
    return (intptr_t)p->d_image;
}
extern "C" intptr_t pwt_coeff_ptr(pwt_plan* p, int num) {
    if (!p || num < 0 || num >= p->nbands) return 0;
    cudaSetDevice(p->device);
    flush_pending(p, 1, true);          // a raw pointer lets the caller see memory: make it current
    materialize_coeffs(p);
    clear_partials(p);                  // ... and lets the caller WRITE it: the fused per-task norm sums may go stale
    return (intptr_t)p->d_band[num];
}

// ---- every band with one copy (the reference's `coeffs` is 3L+1 cudaMemcpy, pypwt.pyx:290-306) ----
extern "C" long long pwt_coeffs_slab_floats(const pwt_plan* p) { return p ? (long long)p->coef_floats : 0; }
extern "C" long long pwt_coeff_offset(const pwt_plan* p, int num) {
    if (!p || num < 0 || num >= p->nbands) return -1;
    return (long long)((p->d_band[num] - p->slab) - (ptrdiff_t)p->coef_base);
}
extern "C" int pwt_get_coeffs(pwt_plan* p, float* dst) {
    if (!p || !dst) return fail(PWT_ERR_ARG, "null argument");
    if (p->state == PWT_INVERSE) {                                  // wt.cu:474-477
        puts("Warning: get_coeff(): inverse() has been performed, the coefficients has been modified and do not make sense anymore.");
        return 1;
    }
    cudaSetDevice(p->device);
    int rc = flush_pending(p, 1, true);
    if (rc == PWT_OK) rc = materialize_coeffs(p);
    if (rc != PWT_OK) return rc;
    CK(cudaMemcpyAsync(dst, p->slab + p->coef_base, p->coef_floats * sizeof(float), cudaMemcpyDeviceToHost, p->stream));
    CK(cudaStreamSynchronize(p->stream));
    return PWT_OK;
}
extern "C" intptr_t pwt_stream_ptr(pwt_plan* p) { return p ? (intptr_t)p->stream : 0; }
extern "C" int pwt_wait_stream(pwt_plan* p, intptr_t producer) {
    if (!p) return fail(PWT_ERR_ARG, "null plan");
    cudaSetDevice(p->device);
    CK(cudaEventRecord(p->ev1, (cudaStream_t)producer));
    CK(cudaStreamWaitEvent(p->stream, p->ev1, 0));
    return PWT_OK;
}

// ---- custom filters ---------------------------------------------------------------------------
// Every kernel here is written for an EVEN number of taps.  The reference also accepts odd lengths (no built-in bank has
// one: CDF 9/7 or LeGall 5/3 given without the padding of demo.cpp:83-179); its kernels then centre the windows
// differently, and each odd-length case equals an even-length bank of len + 1 taps with one particular zero padding:
//   analysis, DWT and SWT (separable.cu:98-102, 416-420: centre len/2, taps len-1-j)      -> [0, k0 .. k(len-1)]
//   DWT synthesis (separable.cu:251-264: half length len/2, taps len-1-(2j+off); tap 0 is
//                  never read, whichever the parity of len/2)                            -> [0, k1 .. k(len-1), 0]
//   SWT synthesis (separable.cu:559-568: centre len/2, taps len-1-j)                      -> [k0 .. k(len-1), 0]
// (derivation in DESIGN.md; checked against the reference's own CUDA build in tests/test_gpu_vs_pdwt.py).
enum { PAD_FRONT = 0, PAD_DROP0 = 1, PAD_BACK = 2 };
static inline int pad_shift(int mode, unsigned len, unsigned padded) { return mode == PAD_FRONT ? (int)(padded - len) : 0; }
static void load_taps(float* dst, const float* src, unsigned len, unsigned padded, int mode) {
    memset(dst, 0, PWT_MAX_TAPS * sizeof(float));
    if (padded == len) mode = PAD_FRONT;                            // even length: taken as given
    const int o = pad_shift(mode, len, padded);
    for (unsigned k = (mode == PAD_DROP0 ? 1 : 0); k < len; k++) dst[o + k] = src[k];
}

static int upload_k2d(pwt_plan* p, float* d_dst, const float* f[4], unsigned len, unsigned padded, int mode) {
    const size_t n = (size_t)padded * padded;
    float* h = (float*)calloc(4 * n, sizeof(float));
    if (!h) return -3;
    if (padded == len) mode = PAD_FRONT;
    const unsigned o = (unsigned)pad_shift(mode, len, padded), k0 = mode == PAD_DROP0 ? 1 : 0;
    for (int b = 0; b < 4; b++)
        for (unsigned i = k0; i < len; i++)
            for (unsigned j = k0; j < len; j++) h[b * n + (o + i) * padded + (o + j)] = f[b][i * len + j];
    cudaError_t e = cudaMemcpyAsync(d_dst, h, 4 * n * sizeof(float), cudaMemcpyHostToDevice, p->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(p->stream);
    free(h);
    return e == cudaSuccess ? 0 : -3;
}

extern "C" int pwt_set_filters_forward(pwt_plan* p, const char* name, unsigned len, const float* f1,
                                       const float* f2, const float* f3, const float* f4) {
    if (!p || !f1 || !f2 || len < 2) return fail(PWT_ERR_ARG, "bad argument");
    const unsigned padded = len + (len & 1);
    if (padded > PWT_MAX_TAPS) {                                    // wt.cu:560-563
        printf("ERROR: Wavelets.set_filters_forward(): filter length (%d) exceeds the maximum size (%d)\n", len, PWT_MAX_TAPS);
        return -1;
    }
    cudaSetDevice(p->device);
    int res = 0;
    if (p->do_separable) {
        load_taps(p->filt.L, f1, len, padded, PAD_FRONT);
        load_taps(p->filt.H, f2, len, padded, PAD_FRONT);
    } else {
        if (!f3 || !f4) {
            puts("ERROR: Wavelets.set_filters_forward(): expected argument 4 and 5 for non-separable filtering");
            return -2;
        }
        const float* f[4] = {f1, f2, f3, f4};
        res = upload_k2d(p, p->d_k2d_fwd, f, len, padded, PAD_FRONT);
        p->ns_rank1 = 0;                                             // arbitrary 2D filters: direct F x F kernels from now on
    }
    p->hlen = (int)padded;
    p->filt.hlen = (int)padded;
    p->custom_len = len;
    if (name) {
        memset(p->wname, 0, sizeof(p->wname));
        strncpy(p->wname, name, sizeof(p->wname) - 1);
    }
    return res;
}

extern "C" int pwt_set_filters_inverse(pwt_plan* p, const float* f1, const float* f2, const float* f3,
                                       const float* f4) {
    if (!p || !f1 || !f2) return fail(PWT_ERR_ARG, "bad argument");
    cudaSetDevice(p->device);
    // the reference assumes the length given to set_filters_forward (wt.cu:587); odd lengths were
    // padded there, the caller still passes the original number of taps
    const unsigned padded = (unsigned)p->hlen;
    const unsigned len = p->custom_len ? p->custom_len : padded;
    const int mode = p->do_swt ? PAD_BACK : PAD_DROP0;              // odd lengths only (see above)
    if (p->do_separable) {
        load_taps(p->filt.IL, f1, len, padded, mode);
        load_taps(p->filt.IH, f2, len, padded, mode);
        return 0;
    }
    if (!f3 || !f4) {
        puts("ERROR: Wavelets.set_filters_inverse(): expected argument 4 and 5 for non-separable filtering");
        return -2;
    }
    const float* f[4] = {f1, f2, f3, f4};
    return upload_k2d(p, p->d_k2d_inv, f, len, padded, mode);
}

// ---- misc -------------------------------------------------------------------------------------
extern "C" int pwt_print_informations(pwt_plan* p) {
    if (!p) return fail(PWT_ERR_ARG, "null plan");
    // wt.cu:511-550
    const char* yn[2] = {"no", "yes"};
    puts("------------- Wavelet transform infos ------------");
    printf("Data dimensions : ");
    if (p->ndims == 2) printf("(%d, %d)\n", p->Nr, p->Nc);
    else if (p->Nr == 1) printf("%d\n", p->Nc);
    else printf("(%d, %d) [batched 1D transform]\n", p->Nr, p->Nc);
    if (p->batch > 1) printf("Stack size : %d\n", p->batch);
    printf("Wavelet name : %s\n", p->wname);
    printf("Number of levels : %d\n", p->nlevels);
    printf("Stationary WT : %s\n", yn[p->do_swt]);
    printf("Cycle spinning : %s\n", yn[p->do_cs]);
    printf("Separable transform : %s\n", yn[p->do_separable]);
    printf("Estimated memory footprint : %.2f MB\n", p->slab_floats * sizeof(float) / 1e6);
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, p->device) == cudaSuccess) printf("Running on device : %s\n", prop.name);
    puts("--------------------------------------------------");
    fflush(stdout);
    return PWT_OK;
}

extern "C" int pwt_sync(pwt_plan* p) {
    if (!p) return fail(PWT_ERR_ARG, "null plan");
    cudaSetDevice(p->device);
    CK(cudaStreamSynchronize(p->stream));
    return PWT_OK;
}

extern "C" const char* pwt_last_error(void) { return g_err; }
extern "C" const char* pwt_version(void) { return "1.0.3"; }
extern "C" int pwt_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

extern "C" int pwt_set_device(int device) {
    CK(cudaSetDevice(device));
    return PWT_OK;
}

extern "C" int pwt_host_alloc(void** ptr, size_t bytes) {
    if (!ptr) return fail(PWT_ERR_ARG, "null argument");
    CK(cudaMallocHost(ptr, bytes ? bytes : 1));
    return PWT_OK;
}
extern "C" int pwt_host_free(void* ptr) {
    if (ptr) CK(cudaFreeHost(ptr));
    return PWT_OK;
}

extern "C" int pwt_timer_start(pwt_plan* p) {
    if (!p) return fail(PWT_ERR_ARG, "null plan");
    cudaSetDevice(p->device);
    CK(cudaEventRecord(p->ev0, p->stream));
    return PWT_OK;
}
extern "C" int pwt_timer_stop(pwt_plan* p, float* ms) {
    if (!p) return fail(PWT_ERR_ARG, "null plan");
    cudaSetDevice(p->device);
    CK(cudaEventRecord(p->ev1, p->stream));
    CK(cudaEventSynchronize(p->ev1));
    float t = 0.f;
    CK(cudaEventElapsedTime(&t, p->ev0, p->ev1));
    if (ms) *ms = t;
    return PWT_OK;
}
extern "C" int pwt_flush_l2(pwt_plan* p) {
    if (!p) return fail(PWT_ERR_ARG, "null plan");
    cudaSetDevice(p->device);
    if (!p->d_flush) {
        p->flush_bytes = (size_t)256 << 20;   // > 126 MB L2
        CK(cudaMalloc(&p->d_flush, p->flush_bytes));
    }
    CK(cudaMemsetAsync(p->d_flush, 0, p->flush_bytes, p->stream));
    return PWT_OK;
}
extern "C" long long pwt_launch_count(const pwt_plan* p) { return p ? p->launches : 0; }
extern "C" int pwt_set_kernel_mode(pwt_plan* p, int mode) {
    if (!p) return fail(PWT_ERR_ARG, "null plan");
    p->kernel_mode = mode;
    return PWT_OK;
}

// ---- multi-GPU ---------------------------------------------------------------------------------
extern "C" int pwt_comm_unique_id(unsigned char id[128]) {
    int rc = load_nccl();
    if (rc != PWT_OK) return rc;
    ncclUniqueId_t u;
    const int r = g_nccl.GetUniqueId(&u);
    if (r != 0) return fail(PWT_ERR_COMM, "ncclGetUniqueId: %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "?");
    memcpy(id, u.internal, 128);
    return PWT_OK;
}

extern "C" int pwt_comm_init(pwt_plan* p, int nranks, int rank, const unsigned char id[128]) {
    if (!p || !id) return fail(PWT_ERR_ARG, "null argument");
    int rc = load_nccl();
    if (rc != PWT_OK) return rc;
    cudaSetDevice(p->device);
    ncclUniqueId_t u;
    memcpy(u.internal, id, 128);
    const int r = g_nccl.CommInitRank(&p->comm, nranks, u, rank);
    if (r != 0) return fail(PWT_ERR_COMM, "ncclCommInitRank: %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "?");
    p->comm_nranks = nranks;
    return PWT_OK;
}

// Single-process variant (`ncclCommInitAll`): one plan per GPU of this process, no launcher and no id
// exchange.  The plans must live on distinct devices.
extern "C" int pwt_comm_init_all(pwt_plan** plans, int n) {
    if (!plans || n < 1 || n > PWT_MAX_DEVICES) return fail(PWT_ERR_ARG, "bad plan list");
    int rc = load_nccl();
    if (rc != PWT_OK) return rc;
    if (!g_nccl.CommInitAll || !g_nccl.GroupStart || !g_nccl.GroupEnd) return fail(PWT_ERR_COMM, "libnccl lacks ncclCommInitAll / ncclGroupStart");
    int devs[PWT_MAX_DEVICES];
    ncclComm_t comms[PWT_MAX_DEVICES];
    for (int i = 0; i < n; i++) {
        if (!plans[i] || plans[i]->comm) return fail(PWT_ERR_ARG, "plan %d is null or already has a communicator", i);
        devs[i] = plans[i]->device;
        for (int j = 0; j < i; j++)
            if (devs[j] == devs[i]) return fail(PWT_ERR_ARG, "plans %d and %d share device %d", j, i, devs[i]);
    }
    const int r = g_nccl.CommInitAll(comms, n, devs);
    if (r != 0) return fail(PWT_ERR_COMM, "ncclCommInitAll: %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "?");
    for (int i = 0; i < n; i++) {
        plans[i]->comm = comms[i];
        plans[i]->comm_nranks = n;
    }
    return PWT_OK;
}

// Global norms of the plans of one communicator created by pwt_comm_init_all: every plan's fused local reduction, then
// the n all-reduces as ONE NCCL group (a single host thread drives all GPUs), each on its plan's stream.
extern "C" int pwt_norms_allreduce_group(pwt_plan** plans, int n, double* n1, double* n2) {
    if (!plans || n < 1) return fail(PWT_ERR_ARG, "bad plan list");
    for (int i = 0; i < n; i++)
        if (!plans[i] || !plans[i]->comm || plans[i]->comm_nranks != n) return fail(PWT_ERR_COMM, "plan %d is not part of an %d-rank communicator", i, n);
    for (int i = 0; i < n; i++) {
        cudaSetDevice(plans[i]->device);
        int rc = local_norms_async(plans[i]);
        if (rc != PWT_OK) return rc;
    }
    int r = g_nccl.GroupStart();
    for (int i = 0; i < n && r == 0; i++)
        r = g_nccl.AllReduce(plans[i]->d_acc, plans[i]->d_acc, 2, kNcclFloat64, kNcclSum, plans[i]->comm, plans[i]->stream);
    const int r2 = g_nccl.GroupEnd();
    if (r != 0 || r2 != 0) return fail(PWT_ERR_COMM, "grouped ncclAllReduce: %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(r ? r : r2) : "?");
    for (int i = 0; i < n; i++) {
        cudaSetDevice(plans[i]->device);
        CK(cudaMemcpyAsync(plans[i]->h_acc, plans[i]->d_acc, 2 * sizeof(double), cudaMemcpyDeviceToHost, plans[i]->stream));
    }
    for (int i = 0; i < n; i++) {
        cudaSetDevice(plans[i]->device);
        CK(cudaStreamSynchronize(plans[i]->stream));
    }
    if (n1) *n1 = plans[0]->h_acc[0];
    if (n2) *n2 = plans[0]->h_acc[1];
    return PWT_OK;
}

extern "C" int pwt_comm_destroy(pwt_plan* p) {
    if (p && p->comm && g_nccl.CommDestroy) {
        cudaSetDevice(p->device);
        cudaStreamSynchronize(p->stream);
        g_nccl.CommDestroy(p->comm);
        p->comm = nullptr;
        p->comm_nranks = 0;
    }
    return PWT_OK;
}

extern "C" int pwt_norms_allreduce(pwt_plan* p, double* n1, double* n2) {
    if (!p) return fail(PWT_ERR_ARG, "null plan");
    if (!p->comm) return fail(PWT_ERR_COMM, "no communicator: call pwt_comm_init first");
    cudaSetDevice(p->device);
    int rc = local_norms_async(p);
    if (rc != PWT_OK) return rc;
    // the all-reduce is enqueued on the same stream right behind the reduction kernel: no host sync
    const int r = g_nccl.AllReduce(p->d_acc, p->d_acc, 2, kNcclFloat64, kNcclSum, p->comm, p->stream);
    if (r != 0) return fail(PWT_ERR_COMM, "ncclAllReduce: %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "?");
    CK(cudaMemcpyAsync(p->h_acc, p->d_acc, 2 * sizeof(double), cudaMemcpyDeviceToHost, p->stream));
    CK(cudaStreamSynchronize(p->stream));
    if (n1) *n1 = p->h_acc[0];
    if (n2) *n2 = p->h_acc[1];
    return PWT_OK;
}
