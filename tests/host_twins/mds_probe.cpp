// Test probe (compiled by tests/test_mds_exact.py with g++): the HOST instantiation of mds_freq_half -- the one
// __host__ __device__ function the device MDS layers run (stark-verifier_b200/csrc/poseidon_g.cuh) -- on raw binary64 bit
// patterns, so that the test can feed it the subnormal operands the kernel uses (a u32 in the low word) and the
// un-recombined half-sums of the double-layer partial rounds.
#include "../../stark-verifier_b200/csrc/poseidon_g.cuh"
#include <cstring>

extern "C" void mds_freq_half_bits(const uint64_t x_bits[12], const uint64_t rc_bits[12], int rc_mode, uint64_t y_bits[12]) {
    double x[12], rc[12], y[12];
    memcpy(x, x_bits, sizeof x);
    memcpy(rc, rc_bits, sizeof rc);
    if (rc_mode == 12) svb::mds_freq_half<12>(x, rc, y);
    else if (rc_mode == 1) svb::mds_freq_half<1>(x, rc, y);
    else svb::mds_freq_half<0>(x, rc, y);
    memcpy(y_bits, y, sizeof y);
}
