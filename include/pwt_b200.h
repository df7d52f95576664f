/*
 * pwt_b200.h -- C ABI of the B200-native wavelet hot path (libpwt_b200.so).
 *
 * This is the drop-in boundary: every entry point replaces one method of the
 * reference's C++ `class Wavelets` (pdwt/src/wt.h:20-76, implemented in
 * pdwt/src/wt.cu) that the Cython wrapper binds in src/pypwt.pyx:29-61.
 * Plain pointers and sizes only; no C++ / torch types cross this boundary.
 *
 * Conventions
 *   - every function returns an int: PWT_OK (0) or a negative PWT_ERR_* code,
 *     except the *_ptr accessors (address or 0) and the two `get_*` copies that
 *     keep the reference's "number of elements copied, 0 on refusal" convention.
 *   - images are row-major float32 `Nr x Nc`; bands are dense row-major float32.
 *   - band numbering (wt.cu:435-506): 2D  0:A  1:H1 2:V1 3:D1 4:H2 ...   1D  0:A 1:D1 2:D2 ...
 *     level 1 = finest.  Band shapes: div2^l (ceil halving, utils.cu:24) for the DWT, Nr x Nc for the SWT.
 *   - all work is enqueued on the plan's stream; like the reference (wt.cu:236-305, no sync)
 *     forward/inverse/threshold calls return without synchronising.  Copies to host synchronise.
 *   - a plan may hold a STACK of `batch` independent images (extension used for multi-GPU
 *     sharding of 3D stacks; the reference handles one image per object).  With batch == 1
 *     the layout is exactly the reference's.
 */
#ifndef PWT_B200_H
#define PWT_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct pwt_plan pwt_plan; /* opaque; replaces `Wavelets*` (wt.h:20) */

/* error codes */
#define PWT_OK 0
#define PWT_ERR_ARG (-1)            /* bad argument / too-long filter (wt.cu:560-563)     */
#define PWT_ERR_UNKNOWN_WAVELET (-2)/* reference: prints + hangs in w_ilog2 (separable.cu:44, utils.cu:14) */
#define PWT_ERR_CUDA (-3)           /* a CUDA runtime call failed (reference ignores all of them) */
#define PWT_ERR_STATE (-4)          /* creation error state (wt.cu:237)                    */
#define PWT_ERR_NOMEM (-5)
#define PWT_ERR_UNSUPPORTED (-6)    /* e.g. cycle spinning in 1D (wt.cu:179-183), ndim > 2 */
#define PWT_ERR_TOO_SMALL (-7)      /* image smaller than the filter: no valid level       */
#define PWT_ERR_COMM (-8)           /* NCCL communicator problem                           */

/* plan states -- same meaning and values as `w_state` (wt.h:8-17) */
enum { PWT_INIT = 0, PWT_FORWARD = 1, PWT_INVERSE = 2, PWT_THRESHOLD = 3, PWT_CREATION_ERROR = 4 };

/* read-only description of a plan; replaces direct reads of `winfos`, `do_separable`, `state`
 * (pypwt.pyx:181-183) */
typedef struct pwt_info {
    int batch;             /* number of stacked images (1 for the reference API)  */
    int Nr, Nc;            /* per-image rows / columns (1D: Nr = 1 or batch rows) */
    int ndims;             /* 1 or 2 after the reference's coercions (wt.cu:133)  */
    int nlevels;           /* after clipping (wt.cu:156-165)                      */
    int hlen;              /* filter length                                       */
    int do_swt, do_separable, do_cycle_spinning;
    int state;             /* PWT_INIT ...                                        */
    int shift_r, shift_c;  /* current cycle-spinning shift (wt.h:31-32)           */
    int nbands;            /* 3*nlevels+1 (2D) or nlevels+1 (1D)                  */
    int device;            /* CUDA device ordinal the plan lives on               */
} pwt_info;

/* ---- life cycle ---------------------------------------------------------------------- */
/* Wavelets::Wavelets(img,Nr,Nc,wname,levels,memisonhost,do_separable,do_cycle_spinning,do_swt,ndim)
 * wt.cu:84-185.  img may be NULL (zero image).  Unknown wavelet -> PWT_ERR_UNKNOWN_WAVELET. */
int pwt_create(pwt_plan** out, const float* img, int Nr, int Nc, const char* wname, int levels,
               int memisonhost, int do_separable, int do_cycle_spinning, int do_swt, int ndim);
/* same, for a stack of `batch` images laid out [batch][Nr][Nc] (extension) */
int pwt_create_batch(pwt_plan** out, const float* img, int batch, int Nr, int Nc, const char* wname,
                     int levels, int memisonhost, int do_separable, int do_cycle_spinning,
                     int do_swt, int ndim);
/* Wavelets::Wavelets(const Wavelets&) wt.cu:191-222 (deep copy of device state) */
int pwt_clone(pwt_plan** out, const pwt_plan* src);
/* Wavelets::~Wavelets wt.cu:226-233 */
void pwt_destroy(pwt_plan* p);
int pwt_get_info(const pwt_plan* p, pwt_info* info);
/* shape of band `num` (what wt.cu:478-502 recomputes on every get_coeff) */
int pwt_band_shape(const pwt_plan* p, int num, int* nr, int* nc);

/* ---- transforms ---------------------------------------------------------------------- */
int pwt_forward(pwt_plan* p);   /* Wavelets::forward wt.cu:236-269 */
int pwt_inverse(pwt_plan* p);   /* Wavelets::inverse wt.cu:271-305; returns 1 if refused (already inverted) */

/* ---- coefficient operators ------------------------------------------------------------ */
int pwt_soft_threshold(pwt_plan* p, float beta, int do_thresh_appcoeffs, int normalize); /* wt.cu:308 */
int pwt_hard_threshold(pwt_plan* p, float beta, int do_thresh_appcoeffs, int normalize); /* wt.cu:318 */
int pwt_group_soft_threshold(pwt_plan* p, float beta, int do_thresh_appcoeffs, int normalize); /* wt.cu:329 */
int pwt_shrink(pwt_plan* p, float beta, int do_thresh_appcoeffs);                        /* wt.cu:340 */
int pwt_proj_linf(pwt_plan* p, float beta, int do_thresh_appcoeffs);                     /* wt.cu:349 */
int pwt_circshift(pwt_plan* p, int sr, int sc, int inplace);                             /* wt.cu:364 */
int pwt_norm1(pwt_plan* p, float* out);      /* Wavelets::norm1   wt.cu:396-416 */
int pwt_norm2sq(pwt_plan* p, float* out);    /* Wavelets::norm2sq wt.cu:368-393 (1D: true sum of squares) */
/* both norms in one pass, double precision, per plan (local to this GPU) */
int pwt_norms(pwt_plan* p, double* norm1, double* norm2sq);
int pwt_add_wavelet(pwt_plan* dst, const pwt_plan* src, float alpha); /* wt.cu:622-655; same return codes */

/* ---- data in / out -------------------------------------------------------------------- */
int pwt_get_image(pwt_plan* p, float* dst);                              /* wt.cu:419 -> batch*Nr*Nc */
int pwt_set_image(pwt_plan* p, const float* img, int mem_is_on_device);  /* wt.cu:425 */
int pwt_get_coeff(pwt_plan* p, float* dst, int num);                     /* wt.cu:473 -> count or 0 */
int pwt_set_coeff(pwt_plan* p, const float* src, int num, int mem_is_on_device); /* wt.cu:435 */
intptr_t pwt_image_ptr(pwt_plan* p);                                     /* wt.cu:658 */
intptr_t pwt_coeff_ptr(pwt_plan* p, int num);                            /* wt.cu:663 */
/* Every band with ONE device->host copy (the reference's `coeffs` property issues 3L+1 blocking copies,
 * pypwt.pyx:290-306 -> wt.cu:473-506).  On the device the bands are contiguous: [band 1 .. band N-1][A].
 * pwt_coeffs_slab_floats = floats of that region, pwt_coeff_offset = where band `num` starts in it (floats, -1 on
 * a bad index); pwt_get_coeffs copies the region into `dst` (>= slab_floats floats, ideally pinned) and
 * synchronises.  Returns PWT_OK, 1 when refused after inverse() (wt.cu:474-477), or an error code. */
long long pwt_coeffs_slab_floats(const pwt_plan* p);
long long pwt_coeff_offset(const pwt_plan* p, int num);
int pwt_get_coeffs(pwt_plan* p, float* dst);
/* the plan's CUDA stream (a cudaStream_t) as an integer, for __cuda_array_interface__ / DLPack consumers */
intptr_t pwt_stream_ptr(pwt_plan* p);
/* order the plan's stream after everything queued so far on `producer_stream` (a cudaStream_t of the same device as an
 * integer): used before a device-to-device set_image / set_coeff of memory another library is still writing */
int pwt_wait_stream(pwt_plan* p, intptr_t producer_stream);

/* ---- custom filter banks --------------------------------------------------------------- */
/* wt.cu:558-581.  separable: f1=L, f2=H (f3,f4 ignored).  non-separable: LL, LH, HL, HH (len x len). */
int pwt_set_filters_forward(pwt_plan* p, const char* name, unsigned len, const float* f1,
                            const float* f2, const float* f3, const float* f4);
/* wt.cu:586-600 */
int pwt_set_filters_inverse(pwt_plan* p, const float* f1, const float* f2, const float* f3,
                            const float* f4);

/* ---- misc ------------------------------------------------------------------------------ */
int pwt_print_informations(pwt_plan* p);   /* wt.cu:511-550 */
int pwt_sync(pwt_plan* p);                 /* cudaStreamSynchronize on the plan's stream */
const char* pwt_last_error(void);          /* thread-local message of the last failure */
const char* pwt_version(void);             /* "1.0.3" -- pypwt.pyx:608-615 */
int pwt_device_count(void);                /* 0 if no usable CUDA device */
int pwt_set_device(int device);            /* device used by subsequently created plans (cudaSetDevice) */

/* look up a built-in bank (filters.cpp:5919-6002): writes hlen taps into each non-NULL array
 * (capacity >= 40) and returns hlen, or PWT_ERR_UNKNOWN_WAVELET.  Host-only, no CUDA call. */
int pwt_lookup_filters(const char* wname, float* L, float* H, float* IL, float* IH);

/* pinned host memory for fast H<->D transfers of numpy buffers */
int pwt_host_alloc(void** ptr, size_t bytes);
int pwt_host_free(void* ptr);

/* ---- measurement helpers (CUDA events on the plan's own stream) ------------------------ */
int pwt_timer_start(pwt_plan* p);
int pwt_timer_stop(pwt_plan* p, float* ms);      /* records, synchronises, returns elapsed ms */
int pwt_flush_l2(pwt_plan* p);                   /* overwrites a >L2-sized scratch buffer on the stream */
long long pwt_launch_count(const pwt_plan* p);   /* kernels launched by this plan so far */
/* per-launch device timing: when enabled every transform kernel launched by forward/inverse is
 * bracketed by CUDA events on the plan's stream.  pwt_profile_read synchronises and returns, in launch
 * order since the last call to pwt_profile_enable, the duration (ms) and a tag of each launch:
 * tag = 100*level + kind, kind 1 = forward (analysis) kernel, 2 = inverse (synthesis) kernel,
 * 3 = other.  Returns the number of records written (<= cap). */
int pwt_profile_enable(pwt_plan* p, int on);
int pwt_profile_read(pwt_plan* p, float* ms, int* tags, int cap);
/* choose kernel family: 0 = auto (default), 1 = force the generic tiled kernels,
 * 2 = shared-memory fast kernels + generic (skip the register-resident kernels),
 * 3 = everything except the fused 3-level cascade,
 * 4 = streaming strip kernels at every level and size (2D and batched 1D DWT, filter length >= 4) */
int pwt_set_kernel_mode(pwt_plan* p, int mode);

/* ---- multi-GPU: one process per GPU, NCCL only for the scalar all-reduce --------------- */
/* 128-byte NCCL unique id, created by rank 0 and distributed by the host (torch.distributed / files) */
int pwt_comm_unique_id(unsigned char id[128]);
int pwt_comm_init(pwt_plan* p, int nranks, int rank, const unsigned char id[128]);
int pwt_comm_destroy(pwt_plan* p);
/* single-process variant (ncclCommInitAll): `plans` = one plan per GPU of this process, on distinct devices */
int pwt_comm_init_all(pwt_plan** plans, int n);
/* global norms over the plans of such a communicator: local fused reductions + the n all-reduces as one NCCL group */
int pwt_norms_allreduce_group(pwt_plan** plans, int n, double* norm1, double* norm2sq);
/* global norms over all ranks' shards: fused local reduction + ncclAllReduce on the plan's stream */
int pwt_norms_allreduce(pwt_plan* p, double* norm1, double* norm2sq);

/* ---- double precision ------------------------------------------------------ */
/* The reference's DOUBLEPRECISION build (pdwt/src/filters.h:16-30 `DTYPE double`, pdwt/Makefile:36-39 libpdwtd.so):
 * the same class (wt.h:20-76) with double samples and the filter table at full precision.  One entry point per
 * method, same argument meaning, return codes and band numbering as the float functions above; `batch` stacked
 * images like pwt_create_batch.  Non-separable plans use the rank-1 identity of the built-in banks (separable kernels,
 * reference slot order).  Not carried over: custom filter banks, add_wavelet, group_soft / proj_linf, NCCL norms. */
typedef struct pwt64_plan pwt64_plan;
int pwt64_create(pwt64_plan** out, const double* img, int batch, int Nr, int Nc, const char* wname, int levels,
                 int memisonhost, int do_separable, int do_cycle_spinning, int do_swt, int ndim);   /* wt.cu:84-185 */
void pwt64_destroy(pwt64_plan* p);                                                                  /* wt.cu:226-233 */
int pwt64_get_info(const pwt64_plan* p, pwt_info* info);
int pwt64_band_shape(const pwt64_plan* p, int num, int* nr, int* nc);
int pwt64_forward(pwt64_plan* p);                                                                   /* wt.cu:236-269 */
int pwt64_inverse(pwt64_plan* p);                                                                   /* wt.cu:271-305 */
int pwt64_soft_threshold(pwt64_plan* p, double beta, int do_thresh_appcoeffs, int normalize);       /* wt.cu:308 */
int pwt64_hard_threshold(pwt64_plan* p, double beta, int do_thresh_appcoeffs, int normalize);       /* wt.cu:318 */
int pwt64_shrink(pwt64_plan* p, double beta, int do_thresh_appcoeffs);                              /* wt.cu:340 */
int pwt64_norms(pwt64_plan* p, double* norm1, double* norm2sq);                                     /* wt.cu:368-416 */
int pwt64_get_image(pwt64_plan* p, double* dst);                                                    /* wt.cu:419 */
int pwt64_set_image(pwt64_plan* p, const double* img, int mem_is_on_device);                        /* wt.cu:425 */
int pwt64_get_coeff(pwt64_plan* p, double* dst, int num);                                           /* wt.cu:473 */
int pwt64_set_coeff(pwt64_plan* p, const double* src, int num, int mem_is_on_device);               /* wt.cu:435 */
intptr_t pwt64_image_ptr(pwt64_plan* p);                                                            /* wt.cu:658 */
intptr_t pwt64_coeff_ptr(pwt64_plan* p, int num);                                                   /* wt.cu:663 */
int pwt64_sync(pwt64_plan* p);
int pwt64_timer_start(pwt64_plan* p);
int pwt64_timer_stop(pwt64_plan* p, float* ms);
long long pwt64_launch_count(const pwt64_plan* p);
int pwt64_lookup_filters(const char* wname, double* L, double* H, double* IL, double* IH);
/* custom separable banks: Wavelets::set_filters_forward / set_filters_inverse wt.cu:558-600 in the DOUBLEPRECISION build
 * (len <= 40 taps; odd lengths mapped like pwt_set_filters_*; non-separable plans -> -2) */
int pwt64_set_filters_forward(pwt64_plan* p, const char* name, unsigned len, const double* lowpass, const double* highpass);
int pwt64_set_filters_inverse(pwt64_plan* p, const double* lowpass, const double* highpass);

/* ---- volumetric (3D) separable DWT ------------------------------------------ */
/* The reference stops at 2D ("3D is not handled", pdwt/README.md:29; pypwt.pyx:155-156 raises on 3D input).  Same
 * conventions carried to volumes [Nz][Ny][Nx]: periodisation (separable.cu:98-102), ceil halving per axis (utils.cu:24),
 * level clip over the smallest axis (wt.cu:156-165); x is filtered first, then y, then z.  Bands of level l (1 = finest)
 * are indexed b = 4 dz + 2 dy + dx, d = 1 meaning the high-pass along that axis: b = 1..7 are pywt.wavedecn's
 * 'aad','ada','add','daa','dad','dda','ddd'; the approximation is (level = nlevels, b = 0). */
typedef struct pwt3_plan pwt3_plan;
int pwt3_create(pwt3_plan** out, const float* vol, int Nz, int Ny, int Nx, const char* wname, int levels, int memisonhost);
void pwt3_destroy(pwt3_plan* p);
int pwt3_levels(const pwt3_plan* p);
int pwt3_band_shape(const pwt3_plan* p, int level, int* nz, int* ny, int* nx);
int pwt3_forward(pwt3_plan* p);
int pwt3_inverse(pwt3_plan* p);                                   /* returns 1 if refused (already inverted) */
int pwt3_soft_threshold(pwt3_plan* p, float beta, int do_thresh_appcoeffs);   /* common.cu:13-22 on every detail band */
int pwt3_hard_threshold(pwt3_plan* p, float beta, int do_thresh_appcoeffs);   /* common.cu:56-64 */
int pwt3_norms(pwt3_plan* p, double* norm1, double* norm2sq);
int pwt3_get_image(pwt3_plan* p, float* dst);
int pwt3_set_image(pwt3_plan* p, const float* vol, int mem_is_on_device);
int pwt3_get_coeff(pwt3_plan* p, float* dst, int level, int b);
int pwt3_set_coeff(pwt3_plan* p, const float* src, int level, int b, int mem_is_on_device);
intptr_t pwt3_coeff_ptr(pwt3_plan* p, int level, int b);
intptr_t pwt3_image_ptr(pwt3_plan* p);
int pwt3_sync(pwt3_plan* p);
int pwt3_timer_start(pwt3_plan* p);
int pwt3_timer_stop(pwt3_plan* p, float* ms);
long long pwt3_launch_count(const pwt3_plan* p);

#ifdef __cplusplus
}
#endif
#endif /* PWT_B200_H */
