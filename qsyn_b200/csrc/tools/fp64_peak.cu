// fp64_peak.cu -- measures what bounds the compute side of the sweep kernels on this GPU:
//   * FP64 FMA pipe (DFMA) throughput
//   * FP64 tensor pipe (mma.sync.aligned.m8n8k4 f64 -> SASS DMMA.8x8x4) throughput
// MEASURED_PEAKS.json has no FP64 figure, so the dense-block roofline
// denominator comes from here.  Prints one JSON line.
#include <cuda_runtime.h>

#include <cstdio>

__device__ __forceinline__ void dmma(double (&c)[2], double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c[0]), "+d"(c[1])
                 : "d"(a), "d"(b));
}

template <int CHAINS>
__global__ void __launch_bounds__(256) dmma_kernel(double* out, int iters, double a0, double b0) {
    double c[CHAINS][2];
#pragma unroll
    for (int k = 0; k < CHAINS; ++k) c[k][0] = c[k][1] = 0.0;
    double const a = a0 + threadIdx.x * 1e-9, b = b0;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < CHAINS; ++k) dmma(c[k], a, b);
    }
    double s = 0;
#pragma unroll
    for (int k = 0; k < CHAINS; ++k) s += c[k][0] + c[k][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int CHAINS>
__global__ void __launch_bounds__(256) dfma_kernel(double* out, int iters, double a0, double b0) {
    double c[CHAINS];
#pragma unroll
    for (int k = 0; k < CHAINS; ++k) c[k] = k;
    double const a = a0 + threadIdx.x * 1e-9, b = b0;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < CHAINS; ++k) c[k] = fma(c[k], a, b);
    }
    double s = 0;
#pragma unroll
    for (int k = 0; k < CHAINS; ++k) s += c[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <class F>
float time_ms(F launch) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    launch();
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 5; ++r) {
        cudaEventRecord(e0);
        launch();
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        best = ms < best ? ms : best;
    }
    return best;
}

int main() {
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    int const blocks = sms * 8, threads = 256, iters = 4096;
    double* out;
    cudaMalloc(&out, sizeof(double) * blocks * threads);
    constexpr int CH = 8;
    float const t_mma = time_ms([&] { dmma_kernel<CH><<<blocks, threads>>>(out, iters, 1.0000001, 0.9999999); });
    float const t_fma = time_ms([&] { dfma_kernel<CH><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); });
    double const warps    = (double)blocks * threads / 32;
    double const mma_flop = warps * iters * CH * 2.0 * 8 * 8 * 4;  // 2*m*n*k per warp instruction
    double const fma_flop = (double)blocks * threads * iters * CH * 2.0;
    std::printf("{\"sms\": %d, \"dmma_m8n8k4_tflops\": %.2f, \"dfma_tflops\": %.2f, \"dmma_ms\": %.3f, \"dfma_ms\": %.3f}\n", sms,
                mma_flop / (t_mma * 1e-3) / 1e12, fma_flop / (t_fma * 1e-3) / 1e12, t_mma, t_fma);
    cudaFree(out);
    return cudaGetLastError() != cudaSuccess;
}
