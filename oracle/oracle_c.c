/* oracle_c.c -- plain-C restatement of the reference's CPU algorithm for the
 * tensor-verification hot path.  TEST INFRASTRUCTURE ONLY (checker / reported CPU baseline):
 * nothing under qsyn_b200/ links or calls this file.
 *
 *   orc_apply_gate   U <- (G on row bits) . U for one gate -- what one
 *                    tensordot(gate, main) + pin update of qcir_to_tensor.cpp:184-193 computes
 *                    (Appendix A.1 of SURVEY.md), written as the straightforward gather /
 *                    small mat-vec / scatter loop, single thread.
 *   orc_equiv_reference_passes
 *                    the equivalence quantities exactly as the reference evaluates them
 *                    (tensor_cmd.cpp:152-154 -> qtensor.hpp:383-417 -> tensor.hpp:393-403):
 *                    separate single-thread left-to-right loops, one per xt::sum expression:
 *                    ip(t1,t2), ip(t1,t1), ip(t2,t2), then sum(t2)/sum(t1) twice (norm, phase).
 *
 * Parity pinning: validated against oracle.py (itself pinned on the reference's golden logs)
 * by tests/test_oracle_golden.py::test_c_restatement_matches_numpy.
 */
#include <complex.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef double _Complex c128;

/* u: 2^n x ncols row-major.  qubits[0..n_ctrls) controls, then n_targets targets
 * (reference ids, 0 = MSB).  mat: 2^t x 2^t target matrix, first target = MSB. */
int orc_apply_gate(double* u_, int n, uint64_t ncols, int n_ctrls, int n_targets, const int* qubits, const double* mat_) {
    c128* u         = (c128*)u_;
    const c128* mat = (const c128*)mat_;
    if (n_targets < 1 || n_targets > 6) return 1;
    int const dt        = 1 << n_targets;
    uint64_t const rows = 1ull << n;
    uint64_t cmask = 0, tmask = 0;
    uint64_t tbit[6];
    for (int i = 0; i < n_ctrls; ++i) cmask |= 1ull << (n - 1 - qubits[i]);
    for (int i = 0; i < n_targets; ++i) {
        tbit[i] = 1ull << (n - 1 - qubits[n_ctrls + i]);
        tmask |= tbit[i];
    }
    c128* in = (c128*)malloc(sizeof(c128) * dt);
    if (!in) return 2;
    uint64_t idx[64];
    for (uint64_t r = 0; r < rows; ++r) {
        if ((r & tmask) != 0 || (r & cmask) != cmask) continue;
        for (int x = 0; x < dt; ++x) {
            uint64_t rr = r;
            for (int i = 0; i < n_targets; ++i)
                if ((x >> (n_targets - 1 - i)) & 1) rr |= tbit[i];
            idx[x] = rr * ncols;
        }
        for (uint64_t c = 0; c < ncols; ++c) {
            for (int x = 0; x < dt; ++x) in[x] = u[idx[x] + c];
            for (int y = 0; y < dt; ++y) {
                c128 acc = 0;
                for (int x = 0; x < dt; ++x) acc += mat[y * dt + x] * in[x];
                u[idx[y] + c] = acc;
            }
        }
    }
    free(in);
    return 0;
}

static c128 sum_conj_mul(const c128* a, const c128* b, uint64_t n) {
    c128 s = 0;
    for (uint64_t i = 0; i < n; ++i) s += conj(a[i]) * b[i];
    return s;
}
static c128 sum_all(const c128* a, uint64_t n) {
    c128 s = 0;
    for (uint64_t i = 0; i < n; ++i) s += a[i];
    return s;
}

/* out = { cosine, norm, phase_rad }.  8 single-tensor passes like the reference. */
void orc_equiv_reference_passes(const double* a_, const double* b_, uint64_t n_elems, double out[3]) {
    const c128 *a = (const c128*)a_, *b = (const c128*)b_;
    double const ip_ab = cabs(sum_conj_mul(a, b, n_elems));
    double const ip_aa = cabs(sum_conj_mul(a, a, n_elems));
    double const ip_bb = cabs(sum_conj_mul(b, b, n_elems));
    out[0]             = ip_ab / sqrt(ip_aa * ip_bb);
    c128 const s_norm  = sum_all(b, n_elems) / sum_all(a, n_elems); /* global_norm  */
    c128 const s_phase = sum_all(b, n_elems) / sum_all(a, n_elems); /* global_phase */
    out[1]             = cabs(s_norm);
    out[2]             = carg(s_phase);
}
