// kernels.hpp -- launchers of the sm_100a kernels (kernels.cu).  Internal.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

#include "program.hpp"

namespace qsx {

// U[r, c] = (r == col0 + c) ? 1 : 0 over a 2^n x ncols shard (K2 of SURVEY 2.2)
cudaError_t launch_set_identity(double2* u, int n, uint64_t col0, uint64_t ncols, cudaStream_t s);

// may the sweep be launched with from_identity (the identity synthesised instead of read)?
bool sweep_can_start_from_identity(int R, const DevSweep& hs);
// one sweep (K1 a-e): `prog` is the device copy of a serialized DevSweep blob, `hs` its host copy
// `u` points at the first column of the block of `grid_ncols` columns to process (a power of two, at most
// `ncols` = the row stride of the shard); col0 is the global column index of that first column
cudaError_t launch_sweep(double2* u, int n, uint64_t ncols, uint64_t grid_ncols, int R, const DevSweep& hs,
                         const unsigned char* prog, size_t prog_bytes, bool reverse, bool from_identity, uint64_t col0,
                         cudaStream_t s);

// launch geometry of an interpreted OR specialised sweep (same grid, same CTA, the tile in dynamic shared memory)
struct SweepGeometry {
    unsigned grid = 0, threads = 0;
    size_t tile_bytes = 0;
    int log2_ncols = 0, log2_ncb = 0;
};
bool sweep_geometry(uint64_t ncols, uint64_t grid_ncols, int R, const DevSweep& hs, SweepGeometry& g);

// fused equivalence reduction (K4): 8 compensated sums of a shard pair.
// `partials` receives 16 doubles per block used (*n_blocks <= equiv_partials_doubles() / 16); the finish kernel merges
// the partials of any number of shard pairs (laid out back to back) into out8 (device memory) in a fixed order.
size_t equiv_partials_doubles();
cudaError_t launch_equiv_partials(const double2* a, const double2* b, uint64_t n_elems, double* partials, int* n_blocks,
                                  cudaStream_t s);
cudaError_t launch_equiv_finish(const double* partials, int n_blocks, double* out8, cudaStream_t s);

// in-place conjugate transpose of a full 2^n x 2^n matrix (K5)
// in-place conjugate transpose of a dim x dim row-major matrix
cudaError_t launch_adjoint_inplace(double2* u, uint64_t dim, cudaStream_t s);
// a <- conj(b^T), b <- conj(a^T) for two c x c row-major blocks (b may live on a peer device).
// part 0: the whole pair; parts 1 and 2 split it between the two owning devices: part 1 is
// launched by a's device, part 2 by b's device with the arguments swapped.
cudaError_t launch_adjoint_block_pair(double2* a, double2* b, uint64_t c, int part, cudaStream_t s);

}  // namespace qsx
