/* qsx_host.h -- C ABI of the C++ HOST layer (libqsxhost.so): the reference-shaped front end
 * of the tensor-verification path (QCir + QASM reader + to_tensor + tensor equiv) for callers
 * that are not C++ (tests, bench.py).  The device boundary itself is include/qsx.h; this header
 * only wraps qsyn_b200/host/, which mirrors these reference interfaces:
 *   qsxh_circuit_*   <- qsyn::qcir::QCir, from_qasm (src/qcir/qcir.hpp, qcir_reader.cpp:47-120),
 *                       str_to_operation (gate_type.cpp:16-70), Phase::from_string (phase.hpp:167-233)
 *   qsxh_circuit_gate_stream <- the per-gate to_tensor(Operation) loop of qcir_to_tensor.cpp:173-194
 *   qsxh_qc2ts       <- `convert qcir tensor` (conversion_cmd.cpp:84-98)
 *   qsxh_tensor_equiv<- `tensor equiv` (tensor_cmd.cpp:120-177), text included
 */
#ifndef QSX_HOST_H_
#define QSX_HOST_H_

#include "qsx.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct qsxh_circuit qsxh_circuit; /* a qsyn::qcir::QCir + its lowered gate stream */
typedef struct qsxh_tensor  qsxh_tensor;  /* a qsyn::tensor::QTensor<double> (device-resident for qc2ts results) */

const char* qsxh_last_error(void);

/* One process per GPU (torchrun / mpirun): every process runs the same commands; tensors of >= 64 x world columns are
 * column-sharded over the processes (this one holds block `rank` on `device`), smaller ones are held in full by every
 * process.  `tensor equiv` all-reduces over the ranks once qsx_comm_init_rank was called.  world = 1 restores the
 * default (one process, QSX_NUM_GPUS devices). */
int qsxh_set_process_shard(int rank, int world, int device);

int qsxh_circuit_new(size_t n_qubits, qsxh_circuit** out);
int qsxh_circuit_from_qasm_text(const char* text, qsxh_circuit** out);
int qsxh_circuit_from_qasm_file(const char* path, qsxh_circuit** out);
/* `qcir gate add <gate> [-ph <phase>] <qubits...>`: gate name as in gate_type.cpp, phase string as in phase.hpp (nullable) */
int qsxh_circuit_append(qsxh_circuit* c, const char* gate, const char* phase, const size_t* qubits, size_t n_qubits);
/* appends circuit `sub` as ONE gate on `qubits` (the first n_ctrls of them are controls): a nested QCir is an Operation
 * in the reference (qcir.hpp:181-183); its to_tensor is the circuit's unitary (qcir_to_tensor.cpp:147) */
int qsxh_circuit_append_circuit(qsxh_circuit* c, const qsxh_circuit* sub, size_t n_ctrls, const size_t* qubits, size_t n_qubits);
int qsxh_circuit_info(const qsxh_circuit* c, size_t* n_qubits, size_t* n_gates);
/* repr of the i-th gate in topological order, e.g. "rz(277π/376)" (UTF-8) */
int qsxh_circuit_gate_repr(qsxh_circuit* c, size_t i, char* buf, size_t buf_size, size_t* gate_id);
/* the ordered gate stream in the layout qsx_plan_create / qsx_apply_gates take; owned by the circuit,
 * valid until the circuit is modified or destroyed */
int qsxh_circuit_gate_stream(qsxh_circuit* c, const qsx_gate** gates, size_t* n_gates);
/* `qcir adjoint`: every gate replaced by its adjoint, order reversed (qcir_action.cpp:136-153) */
int qsxh_circuit_adjoint_inplace(qsxh_circuit* c);
/* `qcir equiv` (qcir_cmd.cpp:603-650 -> qcir_equiv.cpp:19-66): *equivalent = 1 iff c1^-1 . c2 is the identity up to a
 * global phase.  The reference's tableau pre-check is not part of this path; the tensor half runs on the device. */
int qsxh_circuit_equiv(const qsxh_circuit* c1, const qsxh_circuit* c2, int* equivalent);
int qsxh_circuit_destroy(qsxh_circuit* c);

/* convert qcir tensor: nullopt cases of the reference return QSX_ERR_BAD_ARG (empty circuit),
 * QSX_ERR_UNSUPPORTED, QSX_ERR_INTERRUPTED, QSX_ERR_OOM with the reference's log line as the error text */
int qsxh_qc2ts(const qsxh_circuit* c, qsxh_tensor** out);

/* the ZX operand: qsyn::zx::ZXGraph from `.zx` text (src/zx/zx_io.cpp:83-436) and
 * `convert zx tensor` (src/convert/zxgraph_to_tensor.cpp:95-135).  The tensor is built on the host;
 * qsxh_tensor_equiv lifts it to the device when it is an n-qubit operator. */
typedef struct qsxh_zxgraph qsxh_zxgraph;
int qsxh_zxgraph_from_zx_text(const char* text, int keep_id, qsxh_zxgraph** out);
int qsxh_zxgraph_from_zx_file(const char* path, int keep_id, qsxh_zxgraph** out);
int qsxh_zxgraph_info(const qsxh_zxgraph* g, size_t* n_inputs, size_t* n_outputs, size_t* n_vertices, size_t* n_edges);
int qsxh_zx2ts(const qsxh_zxgraph* g, qsxh_tensor** out);
int qsxh_zxgraph_destroy(qsxh_zxgraph* g);
int qsxh_tensor_shape(const qsxh_tensor* t, size_t* dims, size_t max_dims, size_t* n_dims);
/* host copy of the matrix, row-major complex128; n_doubles = 2 * rows * cols */
int qsxh_tensor_download(const qsxh_tensor* t, double* out, size_t n_doubles);
/* one element of a matrix-shaped tensor: `tensor(i, j)` (tensor.hpp:156-175); a device-resident tensor reads 16 bytes */
int qsxh_tensor_element(const qsxh_tensor* t, size_t i, size_t j, double re_im[2]);
/* `tensor write` / `tensor read` (tensor.hpp:326-384): CSV of "re+imj" cells, one matrix row per line */
int qsxh_tensor_write(const qsxh_tensor* t, const char* path);
int qsxh_tensor_read(const char* path, qsxh_tensor** out);
/* `tensor print`: the text qsyn prints (xtensor's format, sk.log:19-20); buf may be NULL to query *needed */
int qsxh_tensor_print(const qsxh_tensor* t, char* buf, size_t buf_size, size_t* needed);
int qsxh_tensor_adjoint_inplace(qsxh_tensor* t);
/* tensor equiv: result numbers + the exact text qsyn prints ("Equivalent\n- Global Norm : 1\n...") */
int qsxh_tensor_equiv(const qsxh_tensor* t1, const qsxh_tensor* t2, double eps, int strict, qsx_equiv_result* result,
                      char* text, size_t text_size);
int qsxh_tensor_destroy(qsxh_tensor* t);

#ifdef __cplusplus
}
#endif
#endif /* QSX_HOST_H_ */
