// bcast_probe.cu -- how should a table that every thread reads (same address) be fetched?
// Compares, per fetch of one complex128 (16 B) used in 4 FP64 instructions:
//   (a) LDS.128 from shared memory (broadcast)  (b) LDC from a __grid_constant__ kernel parameter
//   with a warp-uniform runtime index.  Prints one JSON line (ns per fetch per thread-iteration).
#include <cuda_runtime.h>

#include <cstdio>

struct Params {
    double2 tab[1024];  // 16 KB
};

__global__ void __launch_bounds__(256, 2) k_smem(const double2* __restrict__ gtab, double2* out, int iters, int stride) {
    __shared__ double2 tab[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) tab[i] = gtab[i];
    __syncthreads();
    double2 e[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) e[k] = make_double2(1.0 + k, 0.5 * threadIdx.x);
    int idx = blockIdx.x & 7;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            double2 const w = tab[(idx + k) & 1023];
            e[k]            = make_double2(fma(e[k].x, w.x, -(e[k].y * w.y)), fma(e[k].x, w.y, e[k].y * w.x));
        }
        idx = (idx + stride) & 1023;
    }
    double2 s = make_double2(0, 0);
#pragma unroll
    for (int k = 0; k < 8; ++k) s.x += e[k].x, s.y += e[k].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void __launch_bounds__(256, 2) k_const(const __grid_constant__ Params p, double2* out, int iters, int stride) {
    double2 e[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) e[k] = make_double2(1.0 + k, 0.5 * threadIdx.x);
    int idx = blockIdx.x & 7;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            double2 const w = p.tab[(idx + k) & 1023];
            e[k]            = make_double2(fma(e[k].x, w.x, -(e[k].y * w.y)), fma(e[k].x, w.y, e[k].y * w.x));
        }
        idx = (idx + stride) & 1023;
    }
    double2 s = make_double2(0, 0);
#pragma unroll
    for (int k = 0; k < 8; ++k) s.x += e[k].x, s.y += e[k].y;
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
    int const blocks = 148 * 2, threads = 256, iters = 2048;
    Params* hp = new Params;
    for (int i = 0; i < 1024; ++i) hp->tab[i] = make_double2(0.999 + 1e-6 * i, 1e-4);
    double2 *gtab, *out;
    cudaMalloc(&gtab, sizeof(hp->tab));
    cudaMemcpy(gtab, hp->tab, sizeof(hp->tab), cudaMemcpyHostToDevice);
    cudaMalloc(&out, sizeof(double2) * blocks * threads);
    for (int stride : {8, 136}) {  // 8: walks the whole 16 KB table ; 136: jumps around (cache-unfriendly)
        float const ts = time_ms([&] { k_smem<<<blocks, threads>>>(gtab, out, iters, stride); });
        float const tc = time_ms([&] { k_const<<<blocks, threads>>>(*hp, out, iters, stride); });
        double const fetches = (double)iters * 8;  // per thread
        std::printf("{\"stride\": %d, \"smem_ms\": %.3f, \"const_ms\": %.3f, \"fp64_floor_ms\": %.3f, \"smem_ns_per_fetch\": %.3f, \"const_ns_per_fetch\": %.3f}\n",
                    stride, ts, tc, 1e3 * fetches * 4 * blocks * threads / (64.0 * 148 * 1.9e9), ts * 1e6 / fetches, tc * 1e6 / fetches);
    }
    return cudaGetLastError() != cudaSuccess;
}
