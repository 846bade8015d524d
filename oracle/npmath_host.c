/* TEST INFRASTRUCTURE ONLY -- host build of the product's numpy-equivalent tanh / arctanh
 * (nonlocal-monte-carlo_b200/csrc/nlmc_npmath.h compiles as plain C).  The oracle for these two functions is
 * numpy itself (np.tanh / np.arctanh, the calls at NMC/nmc.py:87,205,216,252); this file only lets the CPU suite
 * check, without a GPU, that the restated algorithm is bit-equal to numpy and to the committed digests
 * (tests/test_npmath.py).  Parity status: PINNED against live numpy on an AVX-512 host and against
 * tests/golden/npmath_digests.json everywhere. */
#include "../nonlocal-monte-carlo_b200/csrc/nlmc_npmath.h"

void nlmc_oracle_np_tanh(const double *x, double *out, long n) {
    for (long i = 0; i < n; ++i) out[i] = nlmc_np_tanh(x[i]);
}

void nlmc_oracle_np_arctanh(const double *x, double *out, long n) {
    for (long i = 0; i < n; ++i) out[i] = nlmc_np_arctanh(x[i]);
}
