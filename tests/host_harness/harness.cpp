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

// The landmark message two ways: message<3, 6> on the explicit cavity P = Lam_b - W0^T W0 (factored per edge), and
// message_downdated<3> on the shared factor of Lam_b.  lam_b packed 21, W0 2x6, e 6, J 18, b 2 per case.
void hh_message_downdated(const double* J, const double* b, const double* var, const double* lam_b, const double* W0, const double* e,
                          const double* damping, const double* old_eta, long n, double* out_explicit, double* out_downdated) {
    for (long i = 0; i < n; ++i) {
        const double *Ji = J + 18 * i, *Wi = W0 + 12 * i;
        double old_lam[21], P[21], ch[CHOL6];
        expand_factored6(Wi, old_lam);
        for (int k = 0; k < 21; ++k) P[k] = lam_b[21 * i + k] - old_lam[k];
        message<3, 6>(Ji + 6, Ji, b + 2 * i, var[i], P, e + 6 * i, damping[i], old_eta + 3 * i, out_explicit + 9 * i, out_explicit + 9 * i + 3);
        cholesky6_packed(lam_b + 21 * i, ch);
        message_downdated<3>(Ji + 6, Ji, b + 2 * i, var[i], ch, Wi, e + 6 * i, damping[i], old_eta + 3 * i, out_downdated + 9 * i,
                             out_downdated + 9 * i + 3);
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

// =====================================================================================================
// HostSweep: the product's per-edge step (gbp_edge.cuh: edge_sweep) and belief arithmetic driven by plain
// host loops over a whole BA graph, in the packed row layout of the ABI.  TEST-ONLY: it lets the CPU suite
// follow complete sweep trajectories of the device arithmetic (relinearisation schedule, damping flags,
// robust losses) against the reference fixtures.  What it does not cover is the kernels' plumbing (tiles,
// bulk copies, shuffle reductions) -- that is what the -m gpu tests are for.
// =====================================================================================================
#include "../../gbp_b200/csrc/gbp_edge.cuh"

#include <algorithm>
#include <vector>

namespace {

struct HostSweep {
    int C = 0, L = 0;
    long F = 0;
    SweepParams prm{};
    bool robust = false;
    bool factored = false;   // keyframe messages stored as eta[6] | W[2][6] (the streaming build); msg_full = their 27-wide form
    std::vector<int> cam, lmk, iters, flags;
    std::vector<double> z, linpoint, msg_cam, msg_lmk, sigma2a, cam_belief, lmk_belief, cam_prior, lmk_prior, msg_full;

    void sweep(int stages) {
        SweepParams p = prm;
        p.stages = stages;
        p.linpoint = linpoint.data(); p.msg_cam = msg_cam.data(); p.msg_lmk = msg_lmk.data();
        p.iters = iters.data(); p.flags = flags.data(); p.sigma2a = sigma2a.data();
        for (long e = 0; e < F; ++e) {
            EdgeRegs r;
            r.it = iters[e]; r.fl = flags[e];
            r.var = robust ? sigma2a[e] : p.var0;
            r.z[0] = z[2 * e]; r.z[1] = z[2 * e + 1];
            for (int k = 0; k < LMK_B; ++k) r.bl[k] = lmk_belief[(size_t)lmk[e] * LMK_B + k];
            const double* cb = &cam_belief[(size_t)cam[e] * CAM_B];
            // beliefs are only read during a sweep, so updating the edge's rows in place is the kernel's
            // "stage in shared memory, overwrite, store back"
            double *lp = &linpoint[9 * e], *ml = &msg_lmk[(size_t)LMK_M * e];
            if (factored) {
                double *mc = &msg_cam[(size_t)CAM_MF * e], *full = &msg_full[(size_t)CAM_M * e];
                double ch[CHOL6];   // the kernels get it once per keyframe from the belief update
                cholesky6_packed(cb + 6, ch);
                if (robust) edge_sweep<true, true>(p, e, r, cb, lp, mc, ml, full, ch);
                else edge_sweep<false, true>(p, e, r, cb, lp, mc, ml, full, ch);
            } else {
                double* mc = &msg_cam[(size_t)CAM_M * e];
                if (robust) edge_sweep<true>(p, e, r, cb, lp, mc, ml);
                else edge_sweep<false>(p, e, r, cb, lp, mc, ml);
            }
        }
    }

    // VariableNode.update_belief (gbp/gbp.py:176-198): messages summed in factor order, then the prior
    void beliefs() {
        std::vector<double> ca((size_t)C * CAM_M, 0.0), la((size_t)L * LMK_M, 0.0);
        const std::vector<double>& mc = factored ? msg_full : msg_cam;   // the kernel sums the tile's 27-wide rows
        for (long e = 0; e < F; ++e) {
            for (int k = 0; k < CAM_M; ++k) ca[(size_t)cam[e] * CAM_M + k] += mc[(size_t)CAM_M * e + k];
            for (int k = 0; k < LMK_M; ++k) la[(size_t)lmk[e] * LMK_M + k] += msg_lmk[(size_t)LMK_M * e + k];
        }
        for (int c = 0; c < C; ++c) {
            double* row = &cam_belief[(size_t)c * CAM_B];
            for (int k = 0; k < CAM_M; ++k) row[k] = ca[(size_t)c * CAM_M + k] + cam_prior[(size_t)c * CAM_M + k];
            spd_solve<6>(row + 6, row, row + 27);
        }
        for (int l = 0; l < L; ++l) {
            double* row = &lmk_belief[(size_t)l * LMK_B];
            for (int k = 0; k < LMK_M; ++k) row[k] = la[(size_t)l * LMK_M + k] + lmk_prior[(size_t)l * LMK_M + k];
            spd_solve<3>(row + 3, row, row + 9);
        }
    }
};

int iteration_stages(int robustify, int local_relin) {
    int st = ST_MESSAGES | ST_BELIEFS;
    if (robustify) st |= ST_ROBUSTIFY;
    if (local_relin) st |= ST_RELIN | ST_LOCAL_DAMPING;
    return st;
}

}  // namespace

extern "C" {

// cam_id / lmk_id / z in FACTOR order (camera-major, file order inside a camera: gbp/gbp_ba.py:128-143)
void* hs_create(double gauss_noise_std, double eta_damping, double beta, double nstds, int num_undamped, int min_linear,
                int loss, int C, int L, long F, const int* cam_id, const int* lmk_id, const double* z,
                const double* cam0, const double* lmk0, const double* K4) {
    HostSweep* h = new HostSweep();
    h->C = C; h->L = L; h->F = F;
    h->robust = loss != 0;
    SweepParams& p = h->prm;
    p.K = Intrinsics{K4[0], K4[1], K4[2], K4[3]};
    p.var0 = gauss_noise_std * gauss_noise_std;
    p.eta_damping = eta_damping; p.beta = beta; p.nstds = nstds;
    p.num_undamped = num_undamped; p.min_linear = min_linear; p.loss = loss;
    h->cam.assign(cam_id, cam_id + F); h->lmk.assign(lmk_id, lmk_id + F); h->z.assign(z, z + 2 * F);
    h->iters.assign(F, 1); h->flags.assign(F, 0); h->sigma2a.assign(F, p.var0);     // gbp/gbp.py:248-249
    h->msg_cam.assign((size_t)F * CAM_M, 0.0); h->msg_lmk.assign((size_t)F * LMK_M, 0.0);
    h->cam_belief.assign((size_t)C * CAM_B, 0.0); h->lmk_belief.assign((size_t)L * LMK_B, 0.0);
    h->cam_prior.assign((size_t)C * CAM_M, 0.0); h->lmk_prior.assign((size_t)L * LMK_M, 0.0);
    for (int c = 0; c < C; ++c) for (int k = 0; k < 6; ++k) h->cam_belief[(size_t)c * CAM_B + 27 + k] = cam0[6 * c + k];
    for (int l = 0; l < L; ++l) for (int k = 0; k < 3; ++k) h->lmk_belief[(size_t)l * LMK_B + 9 + k] = lmk0[3 * l + k];
    h->linpoint.assign((size_t)F * 9, 0.0);                                            // gbp/gbp_ba.py:136-137
    for (long e = 0; e < F; ++e) {
        for (int k = 0; k < 6; ++k) h->linpoint[9 * e + k] = cam0[6 * cam_id[e] + k];
        for (int k = 0; k < 3; ++k) h->linpoint[9 * e + 6 + k] = lmk0[3 * lmk_id[e] + k];
    }
    return h;
}

void hs_destroy(void* hv) { delete static_cast<HostSweep*>(hv); }

// BAFactorGraph.generate_priors_var (gbp/gbp_ba.py:20-34) the way edge_lammax_kernel / *_prior_kernel do it
void hs_generate_priors(void* hv, double weaker) {
    HostSweep* h = static_cast<HostSweep*>(hv);
    std::vector<double> cmax(h->C, 0.0), lmax(h->L, 0.0);
    for (long e = 0; e < h->F; ++e) {
        double J[18], h0[2], m = 0.0;
        linearise(h->prm.K, &h->linpoint[9 * e], J, h0);
        const double var = h->robust ? h->sigma2a[e] : h->prm.var0;
        for (int k = 0; k < 9; ++k) m = std::max(m, (J[k] * J[k] + J[9 + k] * J[9 + k]) / var);
        cmax[h->cam[e]] = std::max(cmax[h->cam[e]], m);
        lmax[h->lmk[e]] = std::max(lmax[h->lmk[e]], m);
    }
    std::fill(h->cam_prior.begin(), h->cam_prior.end(), 0.0);
    std::fill(h->lmk_prior.begin(), h->lmk_prior.end(), 0.0);
    for (int c = 0; c < h->C; ++c) {
        const double lam = cmax[c] / (weaker * weaker);
        for (int i = 0; i < 6; ++i) {
            h->cam_prior[(size_t)c * CAM_M + 6 + sidx<6>(i, i)] = lam;
            h->cam_prior[(size_t)c * CAM_M + i] = lam * h->cam_belief[(size_t)c * CAM_B + 27 + i];
        }
    }
    for (int l = 0; l < h->L; ++l) {
        const double lam = lmax[l] / (weaker * weaker);
        for (int i = 0; i < 3; ++i) {
            h->lmk_prior[(size_t)l * LMK_M + 3 + sidx<3>(i, i)] = lam;
            h->lmk_prior[(size_t)l * LMK_M + i] = lam * h->lmk_belief[(size_t)l * LMK_B + 9 + i];
        }
    }
}

void hs_scale_priors(void* hv, double f) {
    HostSweep* h = static_cast<HostSweep*>(hv);
    for (double& v : h->cam_prior) v *= f;
    for (double& v : h->lmk_prior) v *= f;
}

// gbp_ba_update_beliefs: a sweep that only forms the keyframe-side sums of the STORED messages, then the beliefs
void hs_update_beliefs(void* hv) {
    HostSweep* h = static_cast<HostSweep*>(hv);
    h->sweep(ST_BELIEFS);
    h->beliefs();
}

// switch to the factored keyframe-message layout (before the first sweep: messages are still zero)
void hs_set_factored(void* hv, int on) {
    HostSweep* h = static_cast<HostSweep*>(hv);
    h->factored = on != 0;
    h->msg_cam.assign((size_t)h->F * (h->factored ? CAM_MF : CAM_M), 0.0);
    h->msg_full.assign(h->factored ? (size_t)h->F * CAM_M : 0, 0.0);
}

// round trip of a client-written keyframe message through the factored layout (gbp_ba_write / gbp_ba_read)
void hh_factor_roundtrip(const double* lam21, long n, double* W12, double* lam21_out) {
    for (long i = 0; i < n; ++i) {
        factor_rank2_6(lam21 + 21 * i, W12 + 12 * i);
        expand_factored6(W12 + 12 * i, lam21_out + 21 * i);
    }
}

void hs_iterate(void* hv, int n, int robustify, int local_relin) {
    HostSweep* h = static_cast<HostSweep*>(hv);
    const int st = iteration_stages(robustify, local_relin);
    for (int i = 0; i < n; ++i) {
        h->sweep(st);
        h->beliefs();
    }
}

// one stage set at a time (gbp_ba_sweep_local): GBP_STAGE_* bits; the belief stage = keyframe sums + beliefs
void hs_sweep(void* hv, int stages) {
    HostSweep* h = static_cast<HostSweep*>(hv);
    h->sweep(stages);
    if (stages & ST_BELIEFS) h->beliefs();
}

void hs_fill_iters(void* hv, int v) {
    HostSweep* h = static_cast<HostSweep*>(hv);
    std::fill(h->iters.begin(), h->iters.end(), v);
}

// sum |r|, energy, #factors with iters_since_relin == 0  (metric_kernel: residual at the belief means)
void hs_metrics(void* hv, double out[3]) {
    HostSweep* h = static_cast<HostSweep*>(hv);
    out[0] = out[1] = out[2] = 0.0;
    for (long e = 0; e < h->F; ++e) {
        const double* cm = &h->cam_belief[(size_t)h->cam[e] * CAM_B + 27];
        const double* lm = &h->lmk_belief[(size_t)h->lmk[e] * LMK_B + 9];
        double R[9], hp[2], pp[3];
        so3exp(cm + 3, R);
        project(h->prm.K, R, cm, lm, hp, pp);
        const double r0 = hp[0] - h->z[2 * e], r1 = hp[1] - h->z[2 * e + 1];
        const double a = sqrt(r0 * r0 + r1 * r1);
        out[0] += a;
        out[1] += 0.5 * (a * a) / (h->robust ? h->sigma2a[e] : h->prm.var0);
        out[2] += h->iters[e] == 0 ? 1.0 : 0.0;
    }
}

// field: 0 cam_belief[C][33], 1 lmk_belief[L][12], 2 cam_prior, 3 lmk_prior, 4 msg_cam[F][27], 5 msg_lmk[F][9],
//        6 linpoint[F][9], 9 adaptive variance[F]  (same numbers as gbp_field);  ints: 7 iters, 8 flags
void hs_read(void* hv, int field, double* out) {
    HostSweep* h = static_cast<HostSweep*>(hv);
    const std::vector<double>* v = nullptr;
    switch (field) {
        case 0: v = &h->cam_belief; break;
        case 1: v = &h->lmk_belief; break;
        case 2: v = &h->cam_prior; break;
        case 3: v = &h->lmk_prior; break;
        case 4:
            if (h->factored) {   // what gbp_ba_read returns: the full 27-wide form
                for (long e = 0; e < h->F; ++e) {
                    for (int k = 0; k < 6; ++k) out[(size_t)CAM_M * e + k] = h->msg_cam[(size_t)CAM_MF * e + k];
                    expand_factored6(&h->msg_cam[(size_t)CAM_MF * e + 6], out + (size_t)CAM_M * e + 6);
                }
                return;
            }
            v = &h->msg_cam;
            break;
        case 5: v = &h->msg_lmk; break;
        case 6: v = &h->linpoint; break;
        case 9: v = &h->sigma2a; break;
        default: return;
    }
    std::copy(v->begin(), v->end(), out);
}
void hs_read_int(void* hv, int field, int* out) {
    HostSweep* h = static_cast<HostSweep*>(hv);
    const std::vector<int>& v = field == 7 ? h->iters : h->flags;
    std::copy(v.begin(), v.end(), out);
}
}
