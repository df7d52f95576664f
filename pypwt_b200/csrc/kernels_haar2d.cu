// Haar, one 2D level, sizes the register kernels do not take (widths that are not multiples of 128, odd sizes): the exact
// 1/2 butterfly of the reference (haar.cu:10-58: A = ((a + c) + (b + d)) / 2, ... in that order), one thread per pair of
// adjacent band columns with 128-bit / 64-bit accesses where the sizes allow it, one band column per thread otherwise (the
// analysis repeats the last row / column of an odd size, the synthesis drops the extra ones).  No halo, no shared memory:
// 8 B/px per direction.  The shared-memory tile kernels that served these sizes made Haar SLOWER than db2 on them
// (1500^2 3 levels fwd+inv: 0.077 ms against 0.045; 8188^2: 0.408 against 0.372).
#include "pwt_internal.h"

namespace {
template <bool VEC>
__global__ void __launch_bounds__(256)
k_haar2d_fwd(const float* __restrict__ in, float* __restrict__ A, float* __restrict__ Hb, float* __restrict__ V,
             float* __restrict__ D, int Nr, int Nc, long long in_bs, long long out_bs) {
    const int Nr2 = (Nr + 1) >> 1, Nc2 = (Nc + 1) >> 1, PV = VEC ? Nc2 >> 1 : Nc2;
    in += blockIdx.z * in_bs;
    const long long ob = blockIdx.z * out_bs;
    pwt_pdl_wait();
    for (long long i = blockIdx.x * 256LL + threadIdx.x; i < (long long)Nr2 * PV; i += gridDim.x * 256LL) {
        const int k = (int)(i / PV), p = (int)(i - (long long)k * PV);
        const float* r0 = in + (long long)(2 * k) * Nc;
        const float* r1 = in + (long long)min(2 * k + 1, Nr - 1) * Nc;
        if (VEC) {
            const float4 u = __ldcs(reinterpret_cast<const float4*>(r0 + 4 * p)), l = __ldcs(reinterpret_cast<const float4*>(r1 + 4 * p));
            const float s0 = u.x + l.x, t0 = u.y + l.y, d0 = u.x - l.x, e0 = u.y - l.y;
            const float s1 = u.z + l.z, t1 = u.w + l.w, d1 = u.z - l.z, e1 = u.w - l.w;
            const long long o = ob + (long long)k * Nc2 + 2 * p;
            *reinterpret_cast<float2*>(A + o) = make_float2(0.5f * (s0 + t0), 0.5f * (s1 + t1));
            *reinterpret_cast<float2*>(V + o) = make_float2(0.5f * (s0 - t0), 0.5f * (s1 - t1));
            *reinterpret_cast<float2*>(Hb + o) = make_float2(0.5f * (d0 + e0), 0.5f * (d1 + e1));
            *reinterpret_cast<float2*>(D + o) = make_float2(0.5f * (d0 - e0), 0.5f * (d1 - e1));
        } else {
            const int c0 = 2 * p, c1 = min(2 * p + 1, Nc - 1);
            const float a = r0[c0], b = r0[c1], c = r1[c0], d = r1[c1];
            const long long o = ob + (long long)k * Nc2 + p;
            A[o] = 0.5f * ((a + c) + (b + d));
            V[o] = 0.5f * ((a + c) - (b + d));
            Hb[o] = 0.5f * ((a - c) + (b - d));
            D[o] = 0.5f * ((a - c) - (b - d));
        }
    }
}
template <bool VEC>
__global__ void __launch_bounds__(256)
k_haar2d_inv(const float* __restrict__ A, const float* __restrict__ Hb, const float* __restrict__ V, const float* __restrict__ D,
             float* __restrict__ out, int nr, int nc, int Nro, int Nco, long long in_bs, long long out_bs) {
    const int PV = VEC ? nc >> 1 : nc;
    const long long ib = blockIdx.z * in_bs;
    out += blockIdx.z * out_bs;
    pwt_pdl_wait();
    for (long long i = blockIdx.x * 256LL + threadIdx.x; i < (long long)nr * PV; i += gridDim.x * 256LL) {
        const int j = (int)(i / PV), p = (int)(i - (long long)j * PV);
        float* o0 = out + (long long)(2 * j) * Nco;
        float* o1 = o0 + Nco;
        const bool row1 = 2 * j + 1 < Nro;
        if (VEC) {
            const long long o = ib + (long long)j * nc + 2 * p;
            const float2 a = __ldcs(reinterpret_cast<const float2*>(A + o)), b = __ldcs(reinterpret_cast<const float2*>(V + o));
            const float2 c = __ldcs(reinterpret_cast<const float2*>(Hb + o)), d = __ldcs(reinterpret_cast<const float2*>(D + o));
            const float s0 = a.x + c.x, t0 = b.x + d.x, u0 = a.x - c.x, v0 = b.x - d.x;
            const float s1 = a.y + c.y, t1 = b.y + d.y, u1 = a.y - c.y, v1 = b.y - d.y;
            *reinterpret_cast<float4*>(o0 + 4 * p) = make_float4(0.5f * (s0 + t0), 0.5f * (s0 - t0), 0.5f * (s1 + t1), 0.5f * (s1 - t1));
            if (row1) *reinterpret_cast<float4*>(o1 + 4 * p) = make_float4(0.5f * (u0 + v0), 0.5f * (u0 - v0), 0.5f * (u1 + v1), 0.5f * (u1 - v1));
        } else {
            const long long o = ib + (long long)j * nc + p;
            const float a = A[o], b = V[o], c = Hb[o], d = D[o];
            const bool col1 = 2 * p + 1 < Nco;
            o0[2 * p] = 0.5f * ((a + c) + (b + d));
            if (col1) o0[2 * p + 1] = 0.5f * ((a + c) - (b + d));
            if (row1) {
                o1[2 * p] = 0.5f * ((a - c) + (b - d));
                if (col1) o1[2 * p + 1] = 0.5f * ((a - c) - (b - d));
            }
        }
    }
}
inline unsigned grid_for(long long items) {
    long long g = (items + 255) / 256;
    const long long cap = (long long)pwt_sm_count() * 32;
    return (unsigned)(g < 1 ? 1 : (g > cap ? cap : g));
}
inline bool al(const void* a, const void* b, const void* c, const void* d, const void* e, unsigned mask) {
    return ((((uintptr_t)a) | ((uintptr_t)b) | ((uintptr_t)c) | ((uintptr_t)d) | ((uintptr_t)e)) & mask) == 0;
}
}  // namespace

int pwt_haar2d_fwd_flat(const float* in, float* A, float* Hb, float* V, float* D, int batch, int Nr, int Nc, long long in_bs,
                        long long out_bs, cudaStream_t st) {
    if (batch < 1 || batch > 65535 || Nr < 1 || Nc < 1) return 0;
    const int Nr2 = (Nr + 1) >> 1, Nc2 = (Nc + 1) >> 1;
    const bool vec = (Nc & 3) == 0 && (((uintptr_t)in) & 15) == 0 && al(A, Hb, V, D, A, 7) && (in_bs & 3) == 0 && (out_bs & 1) == 0;
    const dim3 grid(grid_for((long long)Nr2 * (vec ? Nc2 / 2 : Nc2)), 1, batch);
    if (vec) pwt_launch_pdl(k_haar2d_fwd<true>, grid, 256, 0, st, in, A, Hb, V, D, Nr, Nc, in_bs, out_bs);
    else pwt_launch_pdl(k_haar2d_fwd<false>, grid, 256, 0, st, in, A, Hb, V, D, Nr, Nc, in_bs, out_bs);
    return 1;
}
int pwt_haar2d_inv_flat(const float* A, const float* Hb, const float* V, const float* D, float* out, int batch, int nr, int nc,
                        int Nro, int Nco, long long in_bs, long long out_bs, cudaStream_t st) {
    if (batch < 1 || batch > 65535 || nr < 1 || nc < 1) return 0;
    const bool vec = (nc & 1) == 0 && Nco == 2 * nc && (((uintptr_t)out) & 15) == 0 && al(A, Hb, V, D, A, 7) && (in_bs & 1) == 0 &&
                     (out_bs & 3) == 0;
    const dim3 grid(grid_for((long long)nr * (vec ? nc / 2 : nc)), 1, batch);
    if (vec) pwt_launch_pdl(k_haar2d_inv<true>, grid, 256, 0, st, A, Hb, V, D, out, nr, nc, Nro, Nco, in_bs, out_bs);
    else pwt_launch_pdl(k_haar2d_inv<false>, grid, 256, 0, st, A, Hb, V, D, out, nr, nc, Nro, Nco, in_bs, out_bs);
    return 1;
}
