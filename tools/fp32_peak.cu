// fp32_peak.cu — FMA-saturation micro-benchmark: the measured FP32 (CUDA-core) roofline
// denominator for bench.py.  MEASURED_PEAKS.json carries HBM and bf16-tensor peaks only; the
// path-tracing loop is bound by neither, so its roofline is the FP32 FMA pipe.
#include <cuda_runtime.h>
#include <stdint.h>

template <int ILP, class T>
__global__ void __launch_bounds__(256) k_fma(T* out, int iters, T a, T b) {
    T x[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) x[i] = (T)(threadIdx.x + i) * (T)1e-3;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) x[i] = fma(x[i], a, b);
    }
    T s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += x[i];
    if (s == (T)123.456) out[0] = s;   // never true; keeps the chain live
}

template <class T> static int fma_peak(int device, int iters, double* tflops_out, double* ms_out, int* sm_count_out) {
    if (cudaSetDevice(device) != cudaSuccess) return -1;
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    T* out = nullptr;
    cudaMalloc(&out, sizeof(T));
    constexpr int ILP = 16;
    const int blocks = sms * 8, threads = 256;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    double best = 0, best_ms = 0;
    for (int rep = 0; rep < 6; ++rep) {
        cudaEventRecord(e0);
        k_fma<ILP, T><<<blocks, threads>>>(out, iters, (T)1.0000001, (T)1e-7);
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

extern "C" int fp32_peak(int device, double* tflops_out, double* ms_out, int* sm_count_out) { return fma_peak<float>(device, 1 << 15, tflops_out, ms_out, sm_count_out); }
// the FP64 pipe, for the f64 instantiation of `F` (lib.rs:5-6)
extern "C" int fp64_peak(int device, double* tflops_out, double* ms_out, int* sm_count_out) { return fma_peak<double>(device, 1 << 12, tflops_out, ms_out, sm_count_out); }
