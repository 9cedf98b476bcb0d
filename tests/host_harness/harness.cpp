// Test-only host build of gbp_b200/csrc/gbp_math.cuh (the per-edge arithmetic the CUDA
// kernels run).  Compiled with g++ by tests/test_math_host.py so the formulas can be checked
// against the oracle in the CPU-only container.  Never loaded by the product package.
#include "../../gbp_b200/csrc/gbp_math.cuh"

using namespace gbp;

extern "C" {

void hh_linearise(const double* x, long n, const double* K4, double* J, double* h) {
    Intrinsics K{K4[0], K4[1], K4[2], K4[3]};
    for (long i = 0; i < n; ++i) linearise(K, x + 9 * i, J + 18 * i, h + 2 * i);
}

// One edge of compute_messages: beliefs / old messages in the packed ABI layout
// (cam: eta6|lam21, lmk: eta3|lam6).  Outputs the new messages.
void hh_messages(const double* x0, const double* z, long n, const double* K4, double var, const double* damping,
                 const double* bel_c, const double* bel_l, const double* msg_c, const double* msg_l,
                 double* out_c, double* out_l) {
    Intrinsics K{K4[0], K4[1], K4[2], K4[3]};
    for (long i = 0; i < n; ++i) {
        double J[18], h0[2], b[2];
        linearise(K, x0 + 9 * i, J, h0);
        factor_rhs(J, x0 + 9 * i, z + 2 * i, h0, b);
        const double *bc = bel_c + 27 * i, *bl = bel_l + 9 * i, *mc = msg_c + 27 * i, *ml = msg_l + 9 * i;
        double P[21], e[6];
        for (int k = 0; k < 6; ++k) e[k] = bc[k] - mc[k];
        for (int k = 0; k < 21; ++k) P[k] = bc[6 + k] - mc[6 + k];
        message<3, 6>(J + 6, J, b, var, P, e, damping[i], ml, out_l + 9 * i, out_l + 9 * i + 3);
        double P3[6], e3[3];
        for (int k = 0; k < 3; ++k) e3[k] = bl[k] - ml[k];
        for (int k = 0; k < 6; ++k) P3[k] = bl[3 + k] - ml[3 + k];
        message<6, 3>(J, J + 6, b, var, P3, e3, damping[i], mc, out_c + 27 * i, out_c + 27 * i + 6);
    }
}

void hh_solve6(const double* P, const double* r, long n, double* x) {
    for (long i = 0; i < n; ++i) spd_solve<6>(P + 21 * i, r + 6 * i, x + 6 * i);
}
void hh_solve3(const double* P, const double* r, long n, double* x) {
    for (long i = 0; i < n; ++i) spd_solve<3>(P + 6 * i, r + 3 * i, x + 3 * i);
}
double hh_robust_variance(int loss, double var0, double nstds, double r0, double r1, int* flag) {
    bool f;
    const double v = robust_variance(loss, var0, nstds, r0, r1, &f);
    *flag = f ? 1 : 0;
    return v;
}
}
