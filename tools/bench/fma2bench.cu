// Microbenchmark: issue rate of FFMA (scalar fp32 FMA) vs FFMA2 (packed 2 x fp32, sm_100) on one GPU.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fma2bench fma2bench.cu
#include <cstdio>
#include <cuda_runtime.h>
constexpr int ITER = 4096, ACC = 16;
__global__ void k_ffma(float* out, float a, float b) {
    float acc[ACC];
    for (int i = 0; i < ACC; i++) acc[i] = threadIdx.x + i;
    for (int it = 0; it < ITER; it++)
#pragma unroll
        for (int i = 0; i < ACC; i++) acc[i] = fmaf(acc[i], a, b);
    float s = 0;
    for (int i = 0; i < ACC; i++) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_ffma2(float* out, float a, float b) {
    float2 acc[ACC];
    for (int i = 0; i < ACC; i++) acc[i] = make_float2(threadIdx.x + i, threadIdx.x - i);
    const float2 a2 = make_float2(a, a + 1e-7f), b2 = make_float2(b, b);
    for (int it = 0; it < ITER; it++)
#pragma unroll
        for (int i = 0; i < ACC; i++) acc[i] = __ffma2_rn(acc[i], a2, b2);
    float s = 0;
    for (int i = 0; i < ACC; i++) s += acc[i].x + acc[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const int blocks = sms * 4, threads = 512;
    float* out;
    cudaMalloc(&out, sizeof(float) * blocks * threads);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int which = 0; which < 2; which++) {
        float best = 1e9;
        for (int rep = 0; rep < 5; rep++) {
            cudaEventRecord(e0);
            if (which == 0) k_ffma<<<blocks, threads>>>(out, 1.0001f, 0.5f);
            else k_ffma2<<<blocks, threads>>>(out, 1.0001f, 0.5f);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            if (ms < best) best = ms;
        }
        const double fmas = (double)blocks * threads * ITER * ACC * (which ? 2 : 1);
        printf("%s: %.3f ms, %.2f TFMA/s (%.1f TFLOP/s), %.2f warp-instr/clk/SM at 1.9 GHz\n", which ? "FFMA2" : "FFMA ", best,
               fmas / best / 1e9, 2 * fmas / best / 1e9, (double)blocks * threads / 32 * ITER * ACC / (best * 1e-3) / 1.9e9 / sms);
    }
    return 0;
}
