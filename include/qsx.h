/* qsx.h -- C ABI of the B200 tensor-verification engine (libqsx.so).
 *
 * This is the drop-in boundary for qsyn's `convert qcir tensor` / `tensor equiv`
 * hot path.  The reference (DVLab-NTU/qsyn v0.8.1) has no FFI on this path: its
 * seam is the header trio src/tensor/{tensor_util,tensor,qtensor}.hpp plus
 * src/convert/qcir_to_tensor.{hpp,cpp}, the only files that name an xtensor
 * type.  Each entry point below cites the reference interface it replaces
 * (paths relative to the reference root).  Plain C types only: no torch, no
 * C++ types, no CUDA types in any signature.
 *
 * Conventions
 *   - A `qsx_unitary` is ONE column shard of the 2^n x 2^n matrix U[out,in]
 *     that `qc2ts` stores (conversion_cmd.cpp:89): rows = all 2^n output basis
 *     states, columns = [col0, col0+ncols) of the input basis states, row-major,
 *     complex128 (re,im doubles), resident in the HBM of one device.
 *     Qubit 0 is the most significant bit of both indices
 *     (qcir_to_tensor.cpp:201-208, tensor.hpp:266-274).
 *   - Every function returns a qsx_status; qsx_last_error() gives the text
 *     (thread-local).  There is NO CPU fallback: compute entry points return
 *     QSX_ERR_CUDA when no device is usable.
 *   - Work is enqueued on one library-owned stream per device
 *     (qsx_device_stream) and, unless stated otherwise, completed before return.
 */
#ifndef QSX_H_
#define QSX_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define QSX_ABI_VERSION 2
#define QSX_MAX_GATE_QUBITS 16

typedef enum qsx_status {
    QSX_OK              = 0,
    QSX_ERR_BAD_ARG     = 1,
    QSX_ERR_OOM         = 2, /* -> "Memory allocation failed!!" (qcir_to_tensor.cpp:211-214) */
    QSX_ERR_INTERRUPTED = 3, /* -> "Conversion interrupted."    (qcir_to_tensor.cpp:159-162,174-177,196-199) */
    QSX_ERR_CUDA        = 4,
    QSX_ERR_NCCL        = 5,
    QSX_ERR_SHAPE       = 6, /* -> std::invalid_argument / "Shape Mismatch" (tensor.hpp:395, tensor_cmd.cpp:171) */
    QSX_ERR_UNSUPPORTED = 7  /* -> "Conversion of Gate ... is not supported yet!!" (qcir_to_tensor.cpp:180-183) */
} qsx_status;

typedef struct qsx_unitary qsx_unitary; /* opaque: one column shard on one device */
typedef struct qsx_plan    qsx_plan;    /* opaque: a scheduled gate stream        */

/* ---- runtime -------------------------------------------------------------- */
int         qsx_abi_version(void);
const char* qsx_last_error(void);
/* Optional eager start-up: creates the per-device streams of devices 0..n_gpus-1 (n_gpus <= 0: all),
 * enables peer access between them and, for n_gpus > 1, the NCCL communicators of that device set (ncclCommInitAll).
 * Everything it does is otherwise done lazily on first use.  The reference has no counterpart: it is the start-up
 * cost of keeping Tensor storage (tensor.hpp:144) on devices.  qsx_shutdown destroys the communicators. */
int qsx_init(int n_gpus);
/* Call before the process exits (the CLI does at `quit`, the Python binding through atexit): waits for the compiler
 * threads of the specialised sweeps and destroys the NCCL communicators while the CUDA runtime is still alive.  The
 * library also registers exit hooks of its own as a fallback.  Everything can be used again afterwards. */
int qsx_shutdown(void);
/* number of usable CUDA devices (0 on a CPU-only host; never an error) */
int qsx_device_count(int* count);
/* the cudaStream_t (as void*) all work for `device` is enqueued on */
int qsx_device_stream(int device, void** stream);
int qsx_device_synchronize(int device);
/* Shard buffers of destroyed unitaries are parked per device for reuse (up to QSX_POOL_FRACTION, default 0.45, of the
 * device memory; an allocation that fails releases them and retries).  This returns them to the driver now. */
int qsx_device_trim(int device);

/* ---- storage: replaces xt::xarray<complex<double>> inside Tensor<DT> (tensor.hpp:38,144) -------- */
typedef struct qsx_unitary_info {
    int32_t  n_qubits;
    int32_t  device;
    uint64_t col0;   /* first column of this shard          */
    uint64_t ncols;  /* number of columns (a power of two)  */
    uint64_t nrows;  /* 2^n_qubits                          */
    uint64_t bytes;  /* nrows * ncols * 16                  */
    void*    data;   /* device pointer, row-major c128      */
} qsx_unitary_info;

/* uninitialised shard */
int qsx_unitary_create(int device, int n_qubits, uint64_t col0, uint64_t ncols, qsx_unitary** out);
/* identity: replaces the n-fold tensordot(tensor, identity(1)) of qcir_to_tensor.cpp:154-164
 * and QTensor::identity (qtensor.hpp:121-131): one write pass, U[r, r] = 1. */
int qsx_unitary_set_identity(qsx_unitary* u);
/* deep copy on the same device: Tensor copy-ctor used by TensorMgr (data_structure_manager.hpp:138-150) */
int qsx_unitary_clone(const qsx_unitary* u, qsx_unitary** out);
int qsx_unitary_destroy(qsx_unitary* u);
int qsx_unitary_get_info(const qsx_unitary* u, qsx_unitary_info* info);
/* host <-> device; host layout = rows [row0,row0+nrows) x ncols, row-major c128.
 * Replaces element access / xt::adapt (tensor.hpp:156-175,380) for whole blocks. */
int qsx_unitary_upload(qsx_unitary* u, uint64_t row0, uint64_t nrows, const double* host);
int qsx_unitary_download(const qsx_unitary* u, uint64_t row0, uint64_t nrows, double* host);
/* rows [row0,row0+nrows) x shard-local columns [c0,c0+nc) <-> a host block whose rows are host_ld complex128 apart
 * (0: nc, i.e. a packed nrows x nc array; dim to address a column block of the full row-major matrix): one strided
 * copy, no host-side transposition.  Element access `tensor(i, j)` for blocks (tensor.hpp:156-175; decomposer.hpp:64-69
 * reads single elements this way) and the scatter / gather of a column-sharded tensor. */
int qsx_unitary_download_block(const qsx_unitary* u, uint64_t row0, uint64_t nrows, uint64_t c0, uint64_t nc, double* host,
                               uint64_t host_ld);
int qsx_unitary_upload_block(qsx_unitary* u, uint64_t row0, uint64_t nrows, uint64_t c0, uint64_t nc, const double* host,
                             uint64_t host_ld);

/* ---- construction: replaces the per-gate tensordot loop (qcir_to_tensor.cpp:173-194 -> tensor.hpp:413-432) */
typedef struct qsx_gate {
    int32_t n_ctrls;   /* the first n_ctrls entries of qubits[] are controls (ControlGate, basic_gate_type.hpp:233-256) */
    int32_t n_targets; /* then n_targets targets; matrix is 2^t x 2^t, first target = MSB                              */
    int32_t qubits[QSX_MAX_GATE_QUBITS]; /* reference qubit ids: 0 = MSB of the row index                             */
    const double* matrix; /* host, row-major complex128 (re,im), built with the reference formulas (qtensor.hpp:142-277) */
} qsx_gate;

typedef struct qsx_plan_options {
    int32_t struct_size;   /* = sizeof(qsx_plan_options); 0-initialise the rest for defaults                    */
    int32_t fuse;          /* 0: one sweep per gate (no scheduling); 1 (default): blocked multi-gate sweeps     */
    double  snap_tol;      /* matrix entries within snap_tol of 0 / +-1 / +-i are snapped (default 1e-15);
                              < 0 disables snapping: the host-built matrices are applied as given               */
    int32_t reg_qubits;    /* qubits held in registers per thread (1..5, default 4; 5 was measured slower)      */
    int32_t tile_qubits;   /* max row bits staged per shared-memory tile (default 9 with specialised sweeps, else 8) */
    int32_t log2_tile_cols;/* log2 of tile width in columns (default 3 -> 128 B contiguous per row)             */
    int32_t max_ops_per_sweep; /* cap on ops per sweep (default 96)                                             */
    int32_t lookahead;     /* scheduler scan window in gates (default 4096)                                     */
    int32_t dense_block;   /* 0 (default): never; k>=3: fold runs of gates on <=k qubits into dense k-blocks    */
    int32_t no_perm_folding; /* 0 (default): fold X / CX / SWAP into the addressing of stage boundaries when they can
                                be hoisted there, else swap registers; 1: always swap registers; 2: never swap
                                registers (an op that cannot be hoisted waits for the next boundary)            */
    int32_t min_tail_ops;  /* a sweep's trailing stage with fewer ops than this is left to the next sweep
                              (default 8; < 0: keep every stage)                                                */
    int32_t search;        /* rounds of hill climbing over each sweep's tile bits (default 0: 3 rounds when
                              n >= 13, none below; < 0: never, the tile grows first come, first served)         */
    int32_t no_diag_conjugation; /* 1: apply pending diagonal tables before an in-register X / CX / SWAP instead of
                                    carrying them through it (default 0)                                        */
    int32_t no_1q_fusion;  /* 1: never multiply runs of uncontrolled single-qubit gates on one qubit together
                              (default 0: fused on the host when one micro-op is cheaper than the separate ones) */
    int32_t stage_search;  /* 1: choose each stage's register bits by trying every R-subset of the tile bits;
                              < 0: first come, first served; 0 (default): on when n >= 13 (costs ~2 ms of host
                              time per 2000 gates, pays once a sweep takes milliseconds)                        */
    int32_t l2_block_mib;  /* column blocking through the L2: every sweep of the plan is run on a block of columns of
                              about this many MiB before the next block is touched, so only the first sweep reads a
                              block from HBM and only the last write has to reach it.  0 (default): 64 MiB when the
                              cost model says the plan is memory-bound (a few micro-ops per sweep at most);
                              < 0: never; > 0: always, with this block size                                    */
    int32_t jit;           /* specialised sweeps: the program of a sweep is compiled at run time (NVRTC, all host cores, cached
                              by content) into straight-line code and replaces the interpreter once it is ready.
                              0 (default): on for n >= 12 (never blocks: a sweep whose kernel is still compiling is
                              interpreted); 1: on for any n; 2: on and WAIT for the compiler
                              (first use of a program costs ~0.1 s per sweep / host cores); < 0: always interpret      */
} qsx_plan_options;

typedef struct qsx_plan_stats {
    uint64_t n_gates;
    uint64_t n_ops;        /* lowered ops (identity gates dropped)                */
    uint64_t n_sweeps;     /* kernel launches per run                             */
    uint64_t n_stages;     /* register-blocking stages over all sweeps            */
    uint64_t n_micro_ops;  /* register-level micro-ops after merging diagonal ops */
    uint64_t n_folded_perms; /* X / CX / SWAP ops executed purely as address permutations */
    uint64_t n_dense_blocks;
    double   alg_bytes_per_run;   /* sum over sweeps of 32 B x elements touched (this shard geometry) */
    double   gate_bytes_per_run;  /* sum over GATES of f*32*rows*ncols, f per SURVEY 8(d)              */
    double   plan_host_ms; /* time spent scheduling on the host                  */
} qsx_plan_stats;

/* Schedules a gate stream for shards of 2^n_qubits rows x ncols columns.  Pure host work
 * (runs without a GPU).  Gates are applied left to right: U <- G_k ... G_1 G_0 U. */
int qsx_plan_create(int n_qubits, uint64_t ncols, const qsx_gate* gates, size_t n_gates,
                    const qsx_plan_options* opt, qsx_plan** out);
int qsx_plan_destroy(qsx_plan* plan);
int qsx_plan_get_stats(const qsx_plan* plan, qsx_plan_stats* stats);
/* Human/machine readable dump of the schedule (JSON) into buf (NUL-terminated);
 * *needed receives the size required including the NUL. */
int qsx_plan_dump(const qsx_plan* plan, char* buf, size_t buf_size, size_t* needed);

/* Requests the specialised kernels of every sweep of the plan (see qsx_plan_options.jit); wait != 0 blocks until the
 * compiler threads are done with them.  Returns QSX_OK also when run-time compilation is unavailable on this host
 * (no libnvrtc): the plan then simply keeps running in the interpreter. */
int qsx_plan_compile(const qsx_plan* plan, int wait);

typedef struct qsx_jit_stats {
    uint64_t requested;       /* sweep programs looked up                          */
    uint64_t cache_hits;      /* ... found in the cache (compiled or in flight)    */
    uint64_t compiled;        /* kernels compiled so far                           */
    uint64_t failed;          /* programs that could not be compiled or loaded     */
    uint64_t launches;        /* launches of specialised kernels                   */
    double   compile_seconds; /* compiler time summed over the worker threads      */
    int32_t  workers;         /* compiler threads                                  */
    int32_t  available;       /* 0: libnvrtc / driver entry points missing         */
    uint64_t disk_hits;       /* kernels loaded from the on-disk cache (QSX_JIT_CACHE) instead of being compiled */
} qsx_jit_stats;
int qsx_jit_get_stats(qsx_jit_stats* stats);

typedef int (*qsx_stop_fn)(void* arg); /* mirrors `extern bool stop_requested()` (qcir_to_tensor.cpp:20) */

#define QSX_RUN_ASYNC 1u /* do not synchronise at the end (for event timing on qsx_device_stream) */
/* Do not synchronise at the end when the enqueued sweeps are SHORT (less than 64 GB of HBM traffic, about 12 ms): there is
 * nothing worth interrupting, and the caller's next host work (parsing and lowering the other operand of a check)
 * overlaps the device.  A long run is synchronous as without the flag (the stop callback is honoured promptly). */
#define QSX_RUN_ASYNC_IF_SHORT 2u

typedef struct qsx_run_stats {
    uint64_t launches;   /* kernels launched by this call        */
    double   device_ms;  /* CUDA-event time of the launches (0 when QSX_RUN_ASYNC) */
    uint64_t jit_launches; /* ... of which specialised (run-time compiled) sweeps */
    int32_t  plan_cache_hit; /* qsx_apply_gates*: the schedule of this exact gate stream was reused */
    int32_t  reserved;
} qsx_run_stats;

/* U <- plan(U) in place.  should_stop (nullable) is polled between sweeps; when it
 * returns non-zero the call drains the stream and returns QSX_ERR_INTERRUPTED. */
int qsx_plan_run(const qsx_plan* plan, qsx_unitary* u, qsx_stop_fn should_stop, void* stop_arg,
                 uint32_t flags, qsx_run_stats* stats);
/* One-shot: schedule the stream, upload its programs and run it.  This is what to_tensor(QCir) calls
 * (qcir_to_tensor.cpp:143-214).  The last few schedules are kept (keyed by geometry, options and every gate incl. its
 * matrix; QSX_PLAN_CACHE entries, default 8): converting the same circuit again skips the scheduler and finds its
 * programs on the device and its specialised kernels compiled. */
int qsx_apply_gates(qsx_unitary* u, const qsx_gate* gates, size_t n_gates, const qsx_plan_options* opt,
                    qsx_stop_fn should_stop, void* stop_arg, qsx_run_stats* stats);
/* Same, but returns as soon as the sweeps are enqueued on the device's stream: the host can schedule
 * the next operand while this one is being built.  Errors of the sweeps surface at the next
 * synchronising call on the device (download, equiv_partial, qsx_device_synchronize). */
int qsx_apply_gates_async(qsx_unitary* u, const qsx_gate* gates, size_t n_gates, const qsx_plan_options* opt,
                          qsx_run_stats* stats);
/* The general form: qsx_apply_gates with run flags (QSX_RUN_ASYNC, QSX_RUN_ASYNC_IF_SHORT).  The host layer's
 * to_tensor(QCir) calls it with QSX_RUN_ASYNC_IF_SHORT and its stop callback. */
int qsx_apply_gates_ex(qsx_unitary* u, const qsx_gate* gates, size_t n_gates, const qsx_plan_options* opt,
                       qsx_stop_fn should_stop, void* stop_arg, uint32_t flags, qsx_run_stats* stats);

/* ---- verification: replaces inner_product / cosine_similarity (tensor.hpp:393-403) and
 *      global_scalar_factor / global_norm / global_phase / is_equivalent (qtensor.hpp:383-417) ---- */
/* sums[8] = { Re,Im sum conj(a)*b ; sum |a|^2 ; sum |b|^2 ; Re,Im sum a ; Re,Im sum b } over THIS shard pair,
 * one fused read-only pass with compensated summation.  a and b must have identical geometry and device
 * (QSX_ERR_SHAPE otherwise).  Shards of one unitary are combined by adding the 8 doubles
 * (ncclAllReduce / torch.distributed.all_reduce, SUM). */
int qsx_equiv_partial(const qsx_unitary* a, const qsx_unitary* b, double sums[8]);

/* The whole check for a column-sharded pair in ONE call: the reduction is launched on every device concurrently, the
 * per-device sums are combined by a single ncclAllReduce(8 x double, SUM) over NVLink, and 64 bytes are read back.
 * Replaces the three reductions of tensor_cmd.cpp:152-154 (cosine_similarity, global_norm, global_phase).
 *   - one process, several devices: a[k], b[k] = shard k of each operand (any device placement; communicators are
 *     created with ncclCommInitAll on first use or by qsx_init);
 *   - one process per GPU (torchrun): after qsx_comm_init_rank, n_shards = 1 and the all-reduce spans the ranks;
 *   - one device, no communicator: the local sums (no NCCL call is made, libnccl is not even loaded). */
int qsx_equiv_reduce(const qsx_unitary* const* a, const qsx_unitary* const* b, int n_shards, double sums[8]);

/* One process per GPU: rank 0 creates an id (ncclGetUniqueId), the launcher hands its QSX_COMM_ID_BYTES bytes to every
 * rank (torchrun's store, MPI, a file), every rank joins (ncclCommInitRank) with the device that holds its shards. */
#define QSX_COMM_ID_BYTES 128
int qsx_comm_get_unique_id(void* id, size_t id_size);
int qsx_comm_init_rank(int device, int world, int rank, const void* id, size_t id_size);
/* ranks the all-reduce of qsx_equiv_reduce spans beyond this process (1 without qsx_comm_init_rank); NCCL version code */
int qsx_comm_info(int* world, int* nccl_version);

typedef struct qsx_equiv_result {
    int32_t equivalent;      /* verdict incl. --strict rule (tensor_cmd.cpp:152-160)                      */
    int32_t phase_defined;   /* 0 when sum(a) == 0 (reference: NaN -> UB, SURVEY Appendix C trap 2)       */
    double  cosine;          /* |<a,b>| / sqrt(<a,a><b,b>)              (tensor.hpp:401-403)              */
    double  norm;            /* |sum b / sum a|                         (qtensor.hpp:396-398)             */
    double  phase_rad;       /* arg(sum b / sum a)                      (qtensor.hpp:409-411)             */
    int32_t phase_num;       /* dvlab::Phase(phase_rad): rational multiple of pi, Stern-Brocot eps 1e-4   */
    int32_t phase_den;       /*                                         (phase.hpp:38-40, rational.hpp:97-128) */
} qsx_equiv_result;

/* host finish from the (all-reduced) sums with the reference formulas */
int qsx_equiv_finish(const double sums[8], double eps, int strict, qsx_equiv_result* out);

/* ---- whole-tensor ops on one shard ----------------------------------------------- */
/* conjugate transpose in place; only valid for a full (unsharded) unitary: replaces Tensor::adjoint_inplace (tensor.hpp:303-306) */
int qsx_unitary_adjoint_inplace(qsx_unitary* u);
/* conjugate transpose of a column-sharded unitary, in place: shards[k] must hold columns [k c, (k+1) c) and may live
 * on different devices of this process (peer access required).  Diagonal blocks transpose locally; each off-diagonal
 * block pair is exchanged by one kernel that reads and writes the partner block in peer memory. */
int qsx_unitary_adjoint_sharded(qsx_unitary* const* shards, int n_shards);

#ifdef __cplusplus
}
#endif
#endif /* QSX_H_ */
