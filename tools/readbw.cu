// Pure-read HBM bandwidth probe (tuning aid): grid-stride float4 loads over a buffer larger than L2.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void rd(const float4* __restrict__ p, size_t n, float* out) {
    float4 acc = make_float4(0, 0, 0, 0);
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x, st = (size_t)gridDim.x * blockDim.x;
    for (; i + 3 * st < n; i += 4 * st) {
        float4 a = __ldg(p + i), b = __ldg(p + i + st), c = __ldg(p + i + 2 * st), d = __ldg(p + i + 3 * st);
        acc.x += a.x + b.x + c.x + d.x; acc.y += a.y + b.y + c.y + d.y;
        acc.z += a.z + b.z + c.z + d.z; acc.w += a.w + b.w + c.w + d.w;
    }
    if (acc.x + acc.y + acc.z + acc.w == 123.456f) *out = acc.x;
}
// 592 concurrent sequential streams of `chunk` bytes each step (like one CTA per utterance)
__global__ void rd_streams(const float4* __restrict__ p, size_t per_cta_f4, float* out) {
    const float4* q = p + (size_t)blockIdx.x * per_cta_f4;
    float4 acc = make_float4(0, 0, 0, 0);
    for (size_t i = threadIdx.x; i + 3 * blockDim.x < per_cta_f4; i += 4 * blockDim.x) {
        float4 a = __ldg(q + i), b = __ldg(q + i + blockDim.x), c = __ldg(q + i + 2 * blockDim.x), d = __ldg(q + i + 3 * blockDim.x);
        acc.x += a.x + b.x + c.x + d.x; acc.y += a.y + b.y + c.y + d.y;
        acc.z += a.z + b.z + c.z + d.z; acc.w += a.w + b.w + c.w + d.w;
    }
    if (acc.x + acc.y + acc.z + acc.w == 123.456f) *out = acc.x;
}
int main() {
    const size_t bytes = (size_t)2 << 30, n = bytes / 16;
    float4* p; float* o;
    cudaMalloc(&p, bytes); cudaMalloc(&o, 4); cudaMemset(p, 0, bytes);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int cfg = 0; cfg < 4; ++cfg) {
        float best = 1e9;
        for (int it = 0; it < 5; ++it) {
            cudaEventRecord(e0);
            if (cfg == 0) rd<<<148 * 8, 256>>>(p, n, o);
            else if (cfg == 1) rd<<<148 * 16, 512>>>(p, n, o);
            else if (cfg == 2) rd_streams<<<592, 256>>>(p, n / 592, o);
            else rd_streams<<<512, 128>>>(p, (size_t)348 * 1024 * 1024 / 16 / 512, o);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
        }
        const double b = cfg == 3 ? 348.0 * 1024 * 1024 : (double)bytes;
        printf("cfg %d: %.3f ms  %.0f GB/s\n", cfg, best, b / best / 1e6);
    }
    return 0;
}
