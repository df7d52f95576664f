// Double-precision plans: the reference's DOUBLEPRECISION build (pdwt/src/filters.h:16-30 `#define DTYPE double`,
// pdwt/Makefile:36-39 libpdwtd.so), which its Python wrapper never reaches.  Same class surface as
// the fp32 plan -- wt.h:42-75 with DTYPE = double -- behind `pwt64_*` entry points (include/pwt_b200.h).
//
// Kernels: the separable tile kernels of kernels_generic.cu instantiated for double (fused row + column pass per
// level, polyphase synthesis, a-trous passes), taps at the table's full precision.  The roofline doubles per pixel
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
    char wname[128];
    PwtFilters64 filt;
    int lvNr[PWT_MAX_LEVELS + 1], lvNc[PWT_MAX_LEVELS + 1];
    double* slab;
    double* d_image;
    double* d_tmp;       // DWT: one image-sized plane; 2D SWT: three
    int nbands;
    double* d_band[PWT_MAX_BANDS];
    int band_nr[PWT_MAX_BANDS], band_nc[PWT_MAX_BANDS];
    double* d_acc;
    double* h_acc;
    long long launches;
};

namespace {
inline bool is_haar(const pwt64_plan* p) { return p->hlen == 2 && !p->do_swt; }
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
        e = cudaMalloc((void**)&p->slab, total * sizeof(double));
        if (e == cudaSuccess) e = cudaMemsetAsync(p->slab, 0, total * sizeof(double), p->stream);
        if (e == cudaSuccess) e = cudaMalloc((void**)&p->d_acc, 2 * sizeof(double));
        if (e == cudaSuccess) e = cudaMallocHost((void**)&p->h_acc, 2 * sizeof(double));
        if (e != cudaSuccess) rc = fail64(PWT_ERR_CUDA, "allocation of %zu MB failed: %s", total * 8 >> 20, cudaGetErrorString(e));
        else {
            p->d_image = p->slab;
            for (int b = 0; b < p->nbands; b++) p->d_band[b] = p->slab + off[b];
            p->d_tmp = p->slab + tmp_off;
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
            else p->launches += pwt_launch_dwt_fwd1d_f64(src, dstA, p->d_band[l], rows, p->lvNc[l - 1], p->filt, haar, st);
        } else {
            double* Hb = p->d_band[3 * (l - 1) + sH];
            double* V = p->d_band[3 * (l - 1) + sV];
            double* D = p->d_band[3 * (l - 1) + 3];
            if (p->do_swt)
                p->launches += pwt_launch_swt_fwd2d_f64(src, dstA, Hb, V, D, p->d_tmp, B, p->Nr, p->Nc, l, p->filt, st);
            else
                p->launches += pwt_launch_dwt_fwd2d_f64(src, dstA, Hb, V, D, B, p->lvNr[l - 1], p->lvNc[l - 1],
                                                        lvl_elems(p, l - 1), lvl_elems(p, l), p->filt, haar, st);
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
            else p->launches += pwt_launch_dwt_inv1d_f64(cur, p->d_band[l], dst, rows, p->lvNc[l], p->lvNc[l - 1], p->filt, haar, st);
        } else {
            const double* Hb = p->d_band[3 * (l - 1) + sH];
            const double* V = p->d_band[3 * (l - 1) + sV];
            const double* D = p->d_band[3 * (l - 1) + 3];
            if (p->do_swt)
                p->launches += pwt_launch_swt_inv2d_f64(cur, Hb, V, D, dst, p->d_tmp, B, p->Nr, p->Nc, l, p->filt, st);
            else
                p->launches += pwt_launch_dwt_inv2d_f64(cur, Hb, V, D, dst, B, p->lvNr[l], p->lvNc[l], p->lvNr[l - 1],
                                                        p->lvNc[l - 1], lvl_elems(p, l), lvl_elems(p, l - 1), p->filt, haar, st);
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
