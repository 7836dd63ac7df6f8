// jit.hpp -- specialised sweep kernels: code generation (jit_codegen.cpp), run-time compilation with NVRTC and the
// kernel cache (jit.cpp).  Internal.
#pragma once
#include <cuda_runtime.h>

#include <cstddef>
#include <cstdint>
#include <string>

namespace qsx::jit {

// CUDA C++ source of one sweep program (serialized DevSweep blob) as a kernel
//   extern "C" __global__ void <kernel_name>(double2* u, int log2_ncols, int log2_ncb, u32 flags, u32 col0, u32 prefetch_distance)
// with the launch geometry of sweep_kernel<R> (same grid, same CTA size, same dynamic shared memory for the tile).
// persistent: one wave of CTAs looping over the tiles, next tile's first-stage loads prefetched with cp.async (the last
// kernel argument is then the tile count, the grid is the resident wave, the shared memory always holds a tile);
// in: requested, out: whether the generated kernel is (a cluster pair is not).
bool generate_sweep_source(const unsigned char* blob, size_t size, int R, std::string const& kernel_name, int threads, int min_ctas,
                           bool& persistent, std::string& source, std::string& error);

struct Kernel;  // one compiled sweep (cache entry)

// Is run-time compilation usable on this host (libnvrtc + libcuda loadable)?  `why` receives the reason when not.
bool available(std::string* why);

// Looks the program up in the cache (key: content hash + R); on a miss, generates the source and queues the
// compilation on the worker pool.  Never blocks.  Returns nullptr when JIT is unavailable or the program cannot be
// specialised.
Kernel* request(const unsigned char* blob, size_t size, int R, int threads);
// compiled and loadable?  (false while the compilation is in flight, and forever after a failed one)
bool ready(Kernel* k);
// CTA size the kernel was generated for (half of the tile's thread count for a cluster pair)
unsigned cta_threads(Kernel* k);
bool failed(Kernel* k);
// blocks until the compilation of k has finished; returns ready(k)
bool wait(Kernel* k);
// launch on the CURRENT device (the runtime has made the shard's device current)
cudaError_t launch(Kernel* k, double2* u, int log2_ncols, int log2_ncb, uint32_t flags, uint32_t col0, unsigned grid, unsigned threads,
                   size_t smem, cudaStream_t stream);

// process exit: drops the queued compilations and waits for the ones in flight (registered with atexit by the pool)
void shutdown();

struct Stats {
    uint64_t requested = 0, compiled = 0, failed = 0, hits = 0, launches = 0, disk_hits = 0;
    double compile_seconds = 0;  // summed over the worker threads
    int workers            = 0;
};
Stats stats();
std::string last_compile_error();

}  // namespace qsx::jit
