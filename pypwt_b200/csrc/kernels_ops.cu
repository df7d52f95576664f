// Coefficient operators: thresholds, shrink, norms, axpy, circular shift.
// All are streaming (HBM-bound) kernels: one launch covers every band of the pyramid through a
// small segment table, 128-bit accesses on the 16-byte aligned body of each band, scalar head/tail,
// persistent grid sized from the SM count.
// Reference semantics: pdwt/src/common.cu:13-211 (kernels), :219-396 (callers), wt.cu:368-416 (norms).
#include "pwt_internal.h"

namespace {

constexpr int kThreads = 256;

__device__ __forceinline__ float apply_op(float v, float beta, int op) {
    switch (op) {
        case PWT_OP_SOFT:   // common.cu:19  copysignf(max(|v|-beta,0), v)
            return copysignf(fmaxf(fabsf(v) - beta, 0.0f), v);
        case PWT_OP_HARD:   // common.cu:63  max(W_SIGN(|v|-beta),0)*v  (strict >)
            return (fabsf(v) - beta > 0.0f) ? v : 0.0f * v;
        case PWT_OP_PROJ:   // common.cu:107 copysignf(min(|v|,beta), v)
            return copysignf(fminf(fabsf(v), beta), v);
        default:            // PWT_OP_SCALE: cublasSscal(alpha = beta), common.cu:355
            return v * beta;
    }
}

template <int OP>
__global__ void __launch_bounds__(kThreads)
k_eltwise(const __grid_constant__ PwtSegTable t) {
    const long long gtid = (long long)blockIdx.x * kThreads + threadIdx.x;
    const long long gsz = (long long)gridDim.x * kThreads;
    for (int s = 0; s < t.nseg; s++) {
        float* p = t.seg[s].ptr;
        const long long n = t.seg[s].n;
        const float beta = t.seg[s].beta;
        // head up to 16-byte alignment
        long long head = ((16 - ((uintptr_t)p & 15)) & 15) / 4;
        if (head > n) head = n;
        const long long nvec = (n - head) / 4;
        float4* pv = reinterpret_cast<float4*>(p + head);
        for (long long i = gtid; i < nvec; i += gsz) {
            float4 v = pv[i];
            v.x = apply_op(v.x, beta, OP);
            v.y = apply_op(v.y, beta, OP);
            v.z = apply_op(v.z, beta, OP);
            v.w = apply_op(v.w, beta, OP);
            pv[i] = v;
        }
        const long long tail0 = head + nvec * 4;
        const long long nscal = head + (n - tail0);
        for (long long i = gtid; i < nscal; i += gsz) {
            const long long j = i < head ? i : tail0 + (i - head);
            p[j] = apply_op(p[j], beta, OP);
        }
    }
}

// common.cu:145-198: joint shrink of (h, v, d [, a]) by max(1 - beta/||.||_2, 0)
__global__ void __launch_bounds__(kThreads)
k_group_soft(float* __restrict__ h, float* __restrict__ v, float* __restrict__ d,
             float* __restrict__ a, long long n, float beta) {
    for (long long i = (long long)blockIdx.x * kThreads + threadIdx.x; i < n;
         i += (long long)gridDim.x * kThreads) {
        const float vh = h ? h[i] : 0.f, vv = v ? v[i] : 0.f, vd = d[i], va = a ? a[i] : 0.f;
        float nrm = vh * vh + vv * vv + vd * vd;
        if (a) nrm += va * va;
        nrm = sqrtf(nrm);
        const float res = (nrm == 0.f) ? 0.f : fmaxf(1.0f - beta / nrm, 0.0f);
        if (h) h[i] = vh * res;
        if (v) v[i] = vv * res;
        d[i] = vd * res;
        if (a) a[i] = va * res;
    }
}

// sum |c| and sum c^2 over all segments: fp32 loads, fp64 accumulation per thread, warp shuffle,
// one atomicAdd(double) pair per block.
__global__ void __launch_bounds__(kThreads)
k_norms(const __grid_constant__ PwtSegTable t, double* __restrict__ acc) {
    const long long gtid = (long long)blockIdx.x * kThreads + threadIdx.x;
    const long long gsz = (long long)gridDim.x * kThreads;
    double l1 = 0.0, l2 = 0.0;
    for (int s = 0; s < t.nseg; s++) {
        const float* p = t.seg[s].ptr;
        const long long n = t.seg[s].n;
        long long head = ((16 - ((uintptr_t)p & 15)) & 15) / 4;
        if (head > n) head = n;
        const long long nvec = (n - head) / 4;
        const float4* pv = reinterpret_cast<const float4*>(p + head);
        float f1 = 0.f, f2 = 0.f;   // short fp32 partials, flushed to fp64 every iteration
        for (long long i = gtid; i < nvec; i += gsz) {
            const float4 v = pv[i];
            f1 = (fabsf(v.x) + fabsf(v.y)) + (fabsf(v.z) + fabsf(v.w));
            f2 = fmaf(v.x, v.x, v.y * v.y) + fmaf(v.z, v.z, v.w * v.w);
            l1 += (double)f1;
            l2 += (double)f2;
        }
        const long long tail0 = head + nvec * 4;
        const long long nscal = head + (n - tail0);
        for (long long i = gtid; i < nscal; i += gsz) {
            const long long j = i < head ? i : tail0 + (i - head);
            const float v = p[j];
            l1 += (double)fabsf(v);
            l2 += (double)v * (double)v;
        }
    }
    for (int o = 16; o > 0; o >>= 1) {
        l1 += __shfl_xor_sync(0xffffffffu, l1, o);
        l2 += __shfl_xor_sync(0xffffffffu, l2, o);
    }
    __shared__ double s1[kThreads / 32], s2[kThreads / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) {
        s1[warp] = l1;
        s2[warp] = l2;
    }
    __syncthreads();
    if (warp == 0) {
        l1 = lane < kThreads / 32 ? s1[lane] : 0.0;
        l2 = lane < kThreads / 32 ? s2[lane] : 0.0;
        for (int o = 4; o > 0; o >>= 1) {
            l1 += __shfl_xor_sync(0xffffffffu, l1, o);
            l2 += __shfl_xor_sync(0xffffffffu, l2, o);
        }
        if (lane == 0) {
            atomicAdd(acc, l1);
            atomicAdd(acc + 1, l2);
        }
    }
}

// dst += alpha * src (cublasSaxpy, common.cu:499-526)
__global__ void __launch_bounds__(kThreads)
k_axpy(const __grid_constant__ PwtSegTable dst, const __grid_constant__ PwtSegTable src, float alpha) {
    const long long gtid = (long long)blockIdx.x * kThreads + threadIdx.x;
    const long long gsz = (long long)gridDim.x * kThreads;
    for (int s = 0; s < dst.nseg; s++) {
        float* d = dst.seg[s].ptr;
        const float* q = src.seg[s].ptr;
        const long long n = dst.seg[s].n;
        for (long long i = gtid; i < n; i += gsz) d[i] = fmaf(alpha, q[i], d[i]);
    }
}

// out[y,x] = in[(y-sr) mod Nr, (x-sc) mod Nc]  (common.cu:202-211); sr, sc already in [0,N)
__global__ void __launch_bounds__(kThreads)
k_circshift(const float* __restrict__ in, float* __restrict__ out, int Nr, int Nc, int sr, int sc) {
    const int x = blockIdx.x * kThreads + threadIdx.x;
    const long long pb = (long long)blockIdx.z * Nr * Nc;
    if (x >= Nc) return;
    int cx = x - sc;
    if (cx < 0) cx += Nc;
    for (int y = blockIdx.y; y < Nr; y += gridDim.y) {
        int r = y - sr;
        if (r < 0) r += Nr;
        out[pb + (long long)y * Nc + x] = __ldg(in + pb + (long long)r * Nc + cx);
    }
}

// Same shift for widths that are a multiple of 4 (16-byte aligned planes): a thread stores one aligned 128-bit
// group and builds it from the two aligned groups that cover its (misaligned by sc & 3, uniform) source window.
template <int K>
__device__ __forceinline__ float4 shift_pick(const float4& a, const float4& b) {
    if (K == 0) return a;
    if (K == 1) return make_float4(a.y, a.z, a.w, b.x);
    if (K == 2) return make_float4(a.z, a.w, b.x, b.y);
    return make_float4(a.w, b.x, b.y, b.z);
}
template <int K>
__global__ void __launch_bounds__(kThreads)
k_circshift4(const float4* __restrict__ in, float4* __restrict__ out, int Nr, int G, int sr, int g0, long long total) {
    // G = Nc/4 groups per row; source group of output group g is (g + g0) mod G (and the next one when K != 0)
    for (long long i = blockIdx.x * (long long)kThreads + threadIdx.x; i < total; i += (long long)gridDim.x * kThreads) {
        const long long row = i / G;
        const int g = (int)(i - row * G);
        const int y = (int)(row % Nr);
        const long long pb = (row - y) * G;                  // first group of this image
        int r = y - sr;
        if (r < 0) r += Nr;
        int ga = g + g0;
        if (ga >= G) ga -= G;
        int gb = ga + 1;
        if (gb >= G) gb -= G;
        const float4* src = in + pb + (long long)r * G;
        const float4 a = __ldg(src + ga);
        const float4 b = K ? __ldg(src + gb) : a;
        out[i] = shift_pick<K>(a, b);
    }
}

// adds the per-task partial sums written by the fused forward kernel to acc[0..1]
__global__ void __launch_bounds__(kThreads) k_reduce_partials(const double* __restrict__ part, int n, double* __restrict__ acc, int add) {
    double l1 = 0.0, l2 = 0.0;
    for (int i = threadIdx.x; i < n; i += kThreads) {
        l1 += part[2 * i];
        l2 += part[2 * i + 1];
    }
    for (int o = 16; o > 0; o >>= 1) {
        l1 += __shfl_xor_sync(0xffffffffu, l1, o);
        l2 += __shfl_xor_sync(0xffffffffu, l2, o);
    }
    __shared__ double s1[kThreads / 32], s2[kThreads / 32];
    if ((threadIdx.x & 31) == 0) {
        s1[threadIdx.x >> 5] = l1;
        s2[threadIdx.x >> 5] = l2;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < kThreads / 32; w++) {
            l1 += s1[w];
            l2 += s2[w];
        }
        // single thread, stream-ordered after the k_norms of this call when there is one (add != 0)
        acc[0] = add ? acc[0] + l1 : l1;
        acc[1] = add ? acc[1] + l2 : l2;
    }
}

__global__ void __launch_bounds__(kThreads) k_fill(float* p, long long n, float v) {
    for (long long i = (long long)blockIdx.x * kThreads + threadIdx.x; i < n;
         i += (long long)gridDim.x * kThreads)
        p[i] = v;
}

int persistent_grid() {
    return pwt_sm_count() * 8;   // 8 x 256 threads = full occupancy, grid = multiple of the SM count
}

long long total_elems(const PwtSegTable& t) {
    long long n = 0;
    for (int i = 0; i < t.nseg; i++) n += t.seg[i].n;
    return n;
}

int grid_for(long long n) {
    long long need = (n / 4 + kThreads - 1) / kThreads;
    if (need < 1) need = 1;
    const int cap = persistent_grid();
    return (int)(need < cap ? need : cap);
}

}  // namespace

int pwt_launch_eltwise(const PwtSegTable& t, int op, cudaStream_t st) {
    if (t.nseg == 0) return 0;
    const int grid = grid_for(total_elems(t));
    switch (op) {
        case PWT_OP_SOFT: k_eltwise<PWT_OP_SOFT><<<grid, kThreads, 0, st>>>(t); break;
        case PWT_OP_HARD: k_eltwise<PWT_OP_HARD><<<grid, kThreads, 0, st>>>(t); break;
        case PWT_OP_PROJ: k_eltwise<PWT_OP_PROJ><<<grid, kThreads, 0, st>>>(t); break;
        default: k_eltwise<PWT_OP_SCALE><<<grid, kThreads, 0, st>>>(t); break;
    }
    return 1;
}

int pwt_launch_group_soft(float* h, float* v, float* d, float* a, long long n, float beta,
                          cudaStream_t st) {
    k_group_soft<<<grid_for(n * 4), kThreads, 0, st>>>(h, v, d, a, n, beta);
    return 1;
}

int pwt_launch_norms(const PwtSegTable& t, double* d_acc, cudaStream_t st) {
    cudaMemsetAsync(d_acc, 0, 2 * sizeof(double), st);
    if (t.nseg == 0) return 0;
    k_norms<<<grid_for(total_elems(t)), kThreads, 0, st>>>(t, d_acc);
    return 1;
}

int pwt_launch_axpy(const PwtSegTable& dst, const PwtSegTable& src, float alpha, cudaStream_t st) {
    if (dst.nseg == 0) return 0;
    k_axpy<<<grid_for(total_elems(dst) * 4), kThreads, 0, st>>>(dst, src, alpha);
    return 1;
}

int pwt_launch_circshift(const float* in, float* out, int batch, int Nr, int Nc, int sr, int sc,
                         cudaStream_t st) {
    if ((Nc & 3) == 0 && ((((uintptr_t)in) | ((uintptr_t)out)) & 15) == 0) {
        // out[x] = in[x - sc]: the source of output group g starts at column 4g - sc = 4(g + g0) + k, k = (-sc) & 3
        const int G = Nc / 4;
        int c0 = (Nc - sc) % Nc;                             // source column of output column 0
        const int k = c0 & 3, g0 = c0 >> 2;
        const long long total = (long long)batch * Nr * G;
        const unsigned blocks = (unsigned)grid_for(total * 4);
        const float4* i4 = reinterpret_cast<const float4*>(in);
        float4* o4 = reinterpret_cast<float4*>(out);
        if (k == 0) k_circshift4<0><<<blocks, kThreads, 0, st>>>(i4, o4, Nr, G, sr, g0, total);
        else if (k == 1) k_circshift4<1><<<blocks, kThreads, 0, st>>>(i4, o4, Nr, G, sr, g0, total);
        else if (k == 2) k_circshift4<2><<<blocks, kThreads, 0, st>>>(i4, o4, Nr, G, sr, g0, total);
        else k_circshift4<3><<<blocks, kThreads, 0, st>>>(i4, o4, Nr, G, sr, g0, total);
        return 1;
    }
    dim3 grid((Nc + kThreads - 1) / kThreads, Nr < 65535 ? Nr : 65535, batch);
    k_circshift<<<grid, kThreads, 0, st>>>(in, out, Nr, Nc, sr, sc);
    return 1;
}

int pwt_launch_reduce_partials(const double* partials, int n, double* d_acc, int add, cudaStream_t st) {
    k_reduce_partials<<<1, kThreads, 0, st>>>(partials, n, d_acc, add);
    return 1;
}

int pwt_launch_fill(float* p, long long n, float v, cudaStream_t st) {
    k_fill<<<grid_for(n * 4), kThreads, 0, st>>>(p, n, v);
    return 1;
}
