// Memory-pattern micro-benchmark: what HBM bandwidth does a "column strip" walk achieve on B200,
// as a function of the strip width and the number of rows per CTA, compared with a linear copy?
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o membench membench.cu
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("%s: %s\n",#x,cudaGetErrorString(e)); return 1;}}while(0)

__global__ void k_linear(const float4* __restrict__ in, float4* __restrict__ out, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x, st = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += st) out[i] = __ldg(in + i);
}

// strip walk: CTA (bx, by) handles columns [bx*4*NT, +4*NT) and rows [by*TH, +TH); reads the image,
// writes 4 quarter-size planes (like a DWT level): out plane p gets (row/2, col/2) from (row, col) parity.
// U = rows loaded back-to-back before any store (memory-level parallelism).
template <int NT, int U>
__global__ void __launch_bounds__(NT) k_strip(const float* __restrict__ in, float* __restrict__ o0, float* __restrict__ o1,
        float* __restrict__ o2, float* __restrict__ o3, int Nr, int Nc, int TH, int split) {
    const int col = blockIdx.x * 4 * NT + 4 * threadIdx.x;
    if (col >= Nc) return;
    const int r0 = blockIdx.y * TH, r1 = min(r0 + TH, Nr);
    const int Nc2 = Nc / 2;
    for (int r = r0; r < r1; r += U) {
        float4 v[U];
#pragma unroll
        for (int u = 0; u < U; u++) v[u] = __ldg(reinterpret_cast<const float4*>(in + (size_t)(r + u) * Nc + col));
#pragma unroll
        for (int u = 0; u < U; u += 2) {
            if (split) {
                const size_t o = (size_t)((r + u) / 2) * Nc2 + col / 2;
                *reinterpret_cast<float2*>(o0 + o) = make_float2(v[u].x, v[u].z);
                *reinterpret_cast<float2*>(o1 + o) = make_float2(v[u].y, v[u].w);
                *reinterpret_cast<float2*>(o2 + o) = make_float2(v[u + 1].x, v[u + 1].z);
                *reinterpret_cast<float2*>(o3 + o) = make_float2(v[u + 1].y, v[u + 1].w);
            } else {
                *reinterpret_cast<float4*>(o0 + (size_t)(r + u) * Nc + col) = v[u];
                *reinterpret_cast<float4*>(o0 + (size_t)(r + u + 1) * Nc + col) = v[u + 1];
            }
        }
    }
}

int main() {
    const int N = 8192;
    const size_t n = (size_t)N * N;
    float *in, *out;
    CK(cudaMalloc(&in, n * 4)); CK(cudaMalloc(&out, n * 4));
    CK(cudaMemset(in, 1, n * 4)); CK(cudaMemset(out, 0, n * 4));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    auto report = [&](const char* name, float ms) { printf("%-44s %8.1f us  %7.1f GB/s\n", name, ms * 1e3, 2.0 * n * 4 / (ms * 1e-3) / 1e9); };
    float ms;
#define TIME(name, launch) do { for (int i = 0; i < 3; i++) { launch; } cudaEventRecord(e0); for (int i = 0; i < 10; i++) { launch; } cudaEventRecord(e1); CK(cudaEventSynchronize(e1)); cudaEventElapsedTime(&ms, e0, e1); report(name, ms / 10); } while (0)
    TIME("cudaMemcpy D2D", cudaMemcpyAsync(out, in, n * 4, cudaMemcpyDeviceToDevice));
    TIME("linear float4 copy, 148*8 x 256", (k_linear<<<148 * 8, 256>>>((const float4*)in, (float4*)out, n / 4)));
    TIME("linear float4 copy, 148*16 x 512", (k_linear<<<148 * 16, 512>>>((const float4*)in, (float4*)out, n / 4)));
    float* o1 = out + n / 4; float* o2 = out + n / 2; float* o3 = out + 3 * n / 4;
    char name[128];
#define STRIP(NT, U) for (int split = 0; split < 2; split++) for (int TH : {32, 64, 128, 256}) { \
        dim3 g((N + 4 * NT - 1) / (4 * NT), (N + TH - 1) / TH); \
        snprintf(name, sizeof name, "strip NT=%d U=%d TH=%d split=%d grid=%dx%d", NT, U, TH, split, g.x, g.y); \
        TIME(name, (k_strip<NT, U><<<g, NT>>>(in, out, o1, o2, o3, N, N, TH, split))); }
    STRIP(64, 8) STRIP(128, 8) STRIP(256, 8) STRIP(512, 8) STRIP(128, 16) STRIP(256, 4) STRIP(256, 2)
    return 0;
}
