// fp32_peak.cu — FMA-saturation micro-benchmark: the measured FP32 (CUDA-core) roofline
// denominator for bench.py.  MEASURED_PEAKS.json carries HBM and bf16-tensor peaks only; the
// path-tracing loop is bound by neither, so its roofline is the FP32 FMA pipe.
#include <cuda_runtime.h>
#include <stdint.h>

template <int ILP>
__global__ void __launch_bounds__(256) k_fma(float* out, int iters, float a, float b) {
    float x[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) x[i] = (float)(threadIdx.x + i) * 1e-3f;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) x[i] = fmaf(x[i], a, b);
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += x[i];
    if (s == 123.456f) out[0] = s;   // never true; keeps the chain live
}

extern "C" int fp32_peak(int device, double* tflops_out, double* ms_out, int* sm_count_out) {
    if (cudaSetDevice(device) != cudaSuccess) return -1;
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    float* out = nullptr;
    cudaMalloc(&out, 4);
    constexpr int ILP = 16;
    const int iters = 1 << 15, blocks = sms * 8, threads = 256;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    double best = 0, best_ms = 0;
    for (int rep = 0; rep < 6; ++rep) {
        cudaEventRecord(e0);
        k_fma<ILP><<<blocks, threads>>>(out, iters, 1.0000001f, 1e-7f);
        cudaEventRecord(e1);
        if (cudaEventSynchronize(e1) != cudaSuccess) return -2;
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        double flops = 2.0 * ILP * (double)iters * blocks * threads;
        double tf = flops / (ms * 1e-3) / 1e12;
        if (rep > 0 && tf > best) { best = tf; best_ms = ms; }
    }
    cudaFree(out);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    if (tflops_out) *tflops_out = best;
    if (ms_out) *ms_out = best_ms;
    if (sm_count_out) *sm_count_out = sms;
    return 0;
}
