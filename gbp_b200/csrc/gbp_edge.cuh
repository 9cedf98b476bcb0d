// gbp_edge.cuh -- the per-edge step of the synchronous sweep: robustify -> relinearise -> both
// factor-to-variable messages of ONE reprojection edge, plus the plain structs the kernels pass
// around.  Like gbp_math.cuh everything here is __host__ __device__ and free of CUDA intrinsics, so
// the test suite can compile it with g++ and run whole sweeps of the product's arithmetic against the
// reference fixtures without a GPU (tests/host_harness); the product only runs it on the device.
//
// Reference: FactorGraph.synchronous_iteration gbp/gbp.py:86-92 = robustify_loss :296-332,
// relinearise_factors :64-80, compute_all_messages :46-54, Factor.compute_messages :334-373.
#pragma once
#include <stdint.h>

#include "gbp_math.cuh"

namespace gbp {

constexpr int CAM_B = 33, LMK_B = 12, CAM_M = 27, LMK_M = 9;
constexpr int CAM_MF = 18;   // keyframe message with factored precision: eta[6] | W[2][6], Lambda = W^T W (the streaming build, kernel_variant 2)

enum : int { ST_ROBUSTIFY = 1, ST_RELIN = 2, ST_MESSAGES = 4, ST_BELIEFS = 8, ST_LOCAL_DAMPING = 16 };

struct Tile {
    int cam;    // keyframe of every edge in the tile
    int count;  // valid edges (<= T)
};

struct SweepParams {
    const Tile* tiles;
    const int* lmk_idx;
    const double* z;
    double* linpoint;
    double* msg_cam;
    double* msg_lmk;
    int* iters;
    int* flags;
    double* sigma2a;
    const double* cam_belief;
    const double* cam_chol;    // [C][CHOL6] packed Cholesky factor of every keyframe belief's precision (written with the belief)
    const double* lmk_belief;
    double* tile_partial;
    const int* tile_pos;       // [tiles] where a tile's partial sum goes: its position in the keyframe-major tile list
    Intrinsics K;
    double var0, eta_damping, beta, nstds;
    int num_undamped, min_linear, loss, stages;
    int n_tiles;
    int pf_dist;   // > 0: a CTA also prefetches the streams of tile (its tile + pf_dist) into L2 (early-issue kernels)
};

// per-edge register inputs fetched straight from global memory
struct EdgeRegs {
    int it, fl;
    double var;         // adaptive variance (robust losses only), prefetched with the other scalars
    double z[2];
    double bl[LMK_B];   // landmark belief row (gathered)
};

// robustify -> relinearise -> messages for ONE edge.  my_* are the edge's rows in shared memory
// (read, then overwritten in place with the new messages / linearisation point); s_cb is the
// keyframe belief row shared by the whole tile.  Returns true when the edge relinearised.
// FACTORED: my_mc is an 18-double row eta[6] | W[2][6] (Lambda = W^T W) and my_full receives the message in the full
// 27-double form eta | Lambda for the keyframe-side sum (whenever the stages include the belief sums); s_ch is the packed
// Cholesky factor of the keyframe belief's precision (cholesky6_packed, shared by the tile): the message to the landmark
// down-dates it by the old rank-2 message instead of factoring the cavity per edge (message_downdated).
// STAGES: the stages to run as a compile-time constant (the full synchronous iteration: no stage tests in the instruction
// stream), or 0 = take them from p.stages.
// The non-robust path is written without data-dependent branches (the rare relinearisation is a handful of selects and
// predicated stores): together with the branch-free rsqrt / reciprocal the whole edge is then ONE basic block that ptxas can
// schedule across -- the two independent messages interleave -- which matters at 1-3 warps per scheduler.
template <bool ROBUST, bool FACTORED = false, int STAGES = 0>
GBP_HD bool edge_sweep(const SweepParams& p, long long e, EdgeRegs& r, const double* s_cb, double* my_lp, double* my_mc,
                       double* my_ml, double* my_full = nullptr, const double* s_ch = nullptr) {
    const int stages = STAGES ? STAGES : p.stages;
    const double* z = r.z;
    const double* bl = r.bl;
    int it = r.it, fl = r.fl;
    bool relin = false;
    double x0[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) x0[k] = my_lp[k];

    // --- relinearisation test (gbp/gbp.py:72-75): |linpoint - [mu_cam, mu_lmk]| > beta
    if (stages & ST_RELIN) {
        double d2 = 0.0;
#pragma unroll
        for (int k = 0; k < 6; ++k) {
            const double d = x0[k] - s_cb[27 + k];
            d2 += d * d;
        }
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const double d = x0[6 + k] - bl[9 + k];
            d2 += d * d;
        }
        relin = (d2 > p.beta * p.beta) && (it >= p.min_linear);   // |.| > beta without the square root
    }

    double var = p.var0;
    double J[18], h0[2];
    bool lin_done = false;
    if (ROBUST) {
        var = r.var;
        if (stages & ST_ROBUSTIFY) {
            // robustify_loss uses h at the STORED linearisation point (gbp/gbp.py:309-312)
            double r0, r1;
            if (relin) {
                double hold[2];
                meas_fn(p.K, x0, hold);
                r0 = z[0] - hold[0];
                r1 = z[1] - hold[1];
            } else {
                linearise(p.K, x0, J, h0);
                lin_done = true;
                r0 = z[0] - h0[0];
                r1 = z[1] - h0[1];
            }
            bool rf;
            var = robust_variance(p.loss, p.var0, p.nstds, r0, r1, &rf);
            fl = rf ? (fl | 2) : (fl & ~2);
            p.sigma2a[e] = var;
        }
    }

    if (stages & ST_RELIN) {
        if (ROBUST) {
            if (relin) {   // gbp/gbp.py:76-78
#pragma unroll
                for (int k = 0; k < 6; ++k) x0[k] = s_cb[27 + k];
#pragma unroll
                for (int k = 0; k < 3; ++k) x0[6 + k] = bl[9 + k];
#pragma unroll
                for (int k = 0; k < 9; ++k) my_lp[k] = x0[k];
                it = 0;
                fl &= ~1;
                lin_done = false;
            } else {
                it += 1;   // gbp/gbp.py:80
            }
        } else {
            // the same, as selects (gbp/gbp.py:76-80)
#pragma unroll
            for (int k = 0; k < 6; ++k) x0[k] = relin ? s_cb[27 + k] : x0[k];
#pragma unroll
            for (int k = 0; k < 3; ++k) x0[6 + k] = relin ? bl[9 + k] : x0[6 + k];
            if (relin) {
#pragma unroll
                for (int k = 0; k < 9; ++k) my_lp[k] = x0[k];      // predicated stores
            }
            it = relin ? 0 : it + 1;
            fl = relin ? (fl & ~1) : fl;
        }
    }

    if (stages & ST_MESSAGES) {
        if (!lin_done) linearise(p.K, x0, J, h0);
        double b[2];
        factor_rhs(J, x0, z, h0, b);
        double damping = p.eta_damping;
        if (stages & ST_LOCAL_DAMPING) {   // gbp/gbp.py:49-52
            fl = (it == p.num_undamped) ? (fl | 1) : fl;
            damping = (fl & 1) ? p.eta_damping : 0.0;
        }
        // message to the landmark: marginalise the keyframe (6x6 Cholesky; FACTORED: down-date of the tile's shared factor)
        double nl_eta[3], nl_lam[6];
        {
            double ev[6];
#pragma unroll
            for (int k = 0; k < 6; ++k) ev[k] = s_cb[k] - my_mc[k];
            if (FACTORED) {
                message_downdated<3>(J + 6, J, b, var, s_ch, my_mc + 6, ev, damping, my_ml, nl_eta, nl_lam);
            } else {
                double P[21];
#pragma unroll
                for (int k = 0; k < 21; ++k) P[k] = s_cb[6 + k] - my_mc[6 + k];
                message<3, 6>(J + 6, J, b, var, P, ev, damping, my_ml, nl_eta, nl_lam);
            }
        }
        // message to the keyframe: marginalise the landmark (3x3 Cholesky); written in place
        {
            double P[6], ev[3];
#pragma unroll
            for (int k = 0; k < 3; ++k) ev[k] = bl[k] - my_ml[k];
#pragma unroll
            for (int k = 0; k < 6; ++k) P[k] = bl[3 + k] - my_ml[3 + k];
            if (FACTORED) message_factored<6, 3>(J, J + 6, b, var, P, ev, damping, my_mc, my_mc, my_mc + 6);
            else message<6, 3>(J, J + 6, b, var, P, ev, damping, my_mc, my_mc, my_mc + 6);
        }
#pragma unroll
        for (int k = 0; k < 3; ++k) my_ml[k] = nl_eta[k];
#pragma unroll
        for (int k = 0; k < 6; ++k) my_ml[3 + k] = nl_lam[k];
    }
    if (FACTORED && (stages & ST_BELIEFS)) {   // full form of the (new or stored) message for the keyframe-side sum
#pragma unroll
        for (int k = 0; k < 6; ++k) my_full[k] = my_mc[k];
        expand_factored6(my_mc + 6, my_full + 6);
    }
    p.iters[e] = it;
    p.flags[e] = fl;
    return relin;
}

}  // namespace gbp
