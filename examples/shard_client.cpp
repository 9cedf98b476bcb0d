// shard_client.cpp -- a C++ client of libgbp_b200.so that solves ONE bundle-adjustment problem on N GPUs, one process
// per GPU, with nothing but the C ABI of include/gbp_b200.h (no Python, no torch, no NCCL header).
//
//   g++ -O2 -std=c++17 examples/shard_client.cpp -Iinclude -Lgbp_b200/lib -lgbp_b200 -Wl,-rpath,$PWD/gbp_b200/lib -o shard_client
//   for r in 0 1; do ./shard_client --rank $r --nranks 2 --id-file /tmp/gbp.id --bal problem.txt --iters 20 & done; wait
//
// What ba.py does around its loop (ba.py:40-105), sharded: read the BAL file, cut the graph by LANDMARK (rank r owns a
// contiguous landmark range and the measurements of those landmarks; the keyframes are replicated), create the rank's
// graph as its share of one global landmark chunking, attach the communicator, generate_priors_var, update_all_beliefs,
// n x synchronous_iteration(robustify=True, local_relin=True), then ARE / energy over the whole graph.
// The NCCL unique id travels through a file here; any transport of 128 bytes will do.
#include <unistd.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "gbp_b200.h"

#define CK(expr)                                                                        \
    do {                                                                                \
        int _rc = (expr);                                                               \
        if (_rc != GBP_OK) {                                                            \
            fprintf(stderr, "rank %d: %s -> %d: %s\n", rank, #expr, _rc, gbp_last_error()); \
            return 1;                                                                   \
        }                                                                               \
    } while (0)

int main(int argc, char** argv) {
    int rank = 0, nranks = 1, iters = 20, device = -1;
    std::string id_file = "/tmp/gbp_b200.id", bal;
    for (int i = 1; i + 1 < argc; i += 2) {
        const std::string k = argv[i];
        if (k == "--rank") rank = atoi(argv[i + 1]);
        else if (k == "--nranks") nranks = atoi(argv[i + 1]);
        else if (k == "--iters") iters = atoi(argv[i + 1]);
        else if (k == "--device") device = atoi(argv[i + 1]);
        else if (k == "--id-file") id_file = argv[i + 1];
        else if (k == "--bal") bal = argv[i + 1];
    }
    if (bal.empty() || rank < 0 || rank >= nranks) { fprintf(stderr, "usage: --rank r --nranks n --id-file path --bal file [--iters k] [--device d]\n"); return 2; }
    if (device < 0) device = rank;

    // ---- the whole problem (utils/read_balfile.py:4-37)
    gbp_bal_file* f = nullptr;
    CK(gbp_bal_open(bal.c_str(), &f));
    int64_t sz[3];
    CK(gbp_bal_sizes(f, sz));
    const int64_t C = sz[0], L = sz[1], F = sz[2];
    std::vector<int32_t> cam_id(F), lmk_id(F);
    std::vector<double> z(2 * F), cam_mu(6 * C), lmk_mu(3 * L);
    double K4[4];
    CK(gbp_bal_copy(f, cam_id.data(), lmk_id.data(), z.data(), cam_mu.data(), lmk_mu.data(), K4));
    gbp_bal_close(f);

    // ---- this rank's share: landmarks [l0, l1) = chunk `rank` of `nranks` chunks, file order preserved
    const int64_t l0 = L * rank / nranks, l1 = L * (rank + 1) / nranks;
    std::vector<int32_t> cam_loc, lmk_loc;
    std::vector<double> z_loc;
    for (int64_t i = 0; i < F; ++i)
        if (lmk_id[i] >= l0 && lmk_id[i] < l1) {
            cam_loc.push_back(cam_id[i]);
            lmk_loc.push_back((int32_t)(lmk_id[i] - l0));
            z_loc.push_back(z[2 * i]);
            z_loc.push_back(z[2 * i + 1]);
        }

    gbp_config cfg;
    memset(&cfg, 0, sizeof(cfg));      // ba.py:51-60 defaults
    cfg.gauss_noise_std = 2; cfg.eta_damping = 0.4; cfg.beta = 0.01; cfg.Nstds = 3.0;
    cfg.num_undamped_iters = 6; cfg.min_linear_iters = 8; cfg.loss = GBP_LOSS_NONE;
    cfg.lmk_chunks = 1; cfg.lmk_chunk_first = rank; cfg.lmk_chunks_total = nranks; cfg.lmk_first = l0; cfg.lmk_total = L;

    gbp_handle h = nullptr;
    CK(gbp_ba_create(&cfg, (int32_t)C, (int32_t)(l1 - l0), (int64_t)cam_loc.size(), cam_loc.data(), lmk_loc.data(), z_loc.data(),
                     cam_mu.data(), lmk_mu.data() + 3 * l0, K4, device, nullptr, &h));

    // ---- communicator: rank 0 makes the id, the others wait for the file
    unsigned char id[GBP_COMM_ID_BYTES];
    if (rank == 0) {
        CK(gbp_comm_unique_id(id));
        const std::string tmp = id_file + ".tmp";
        FILE* o = fopen(tmp.c_str(), "wb");
        if (!o || fwrite(id, 1, sizeof(id), o) != sizeof(id)) { fprintf(stderr, "cannot write %s\n", tmp.c_str()); return 1; }
        fclose(o);
        rename(tmp.c_str(), id_file.c_str());
    } else {
        FILE* in = nullptr;
        for (int tries = 0; tries < 600 && !(in = fopen(id_file.c_str(), "rb")); ++tries) usleep(100000);
        if (!in || fread(id, 1, sizeof(id), in) != sizeof(id)) { fprintf(stderr, "rank %d: no id in %s\n", rank, id_file.c_str()); return 1; }
        fclose(in);
    }
    gbp_comm comm = nullptr;
    CK(gbp_comm_create(id, rank, nranks, device, &comm));
    CK(gbp_ba_attach_comm(h, comm));

    // ---- ba.py:62-105 without the viewer
    CK(gbp_ba_generate_priors(h, 50.0, nullptr));       // keyframe maxima over all ranks
    CK(gbp_ba_update_beliefs(h));
    double m[3];
    CK(gbp_ba_metrics(h, m));
    if (rank == 0) printf("initial   ARE %.12f  energy %.9f\n", m[0] / (double)F, m[1]);
    int done = 0;
    for (int mark : {3, 8, iters}) {                     // ba.py:91-93: iters_since_relin = 1 before sweeps 3 and 8
        const int target = mark < iters ? mark : iters;
        if (target > done) {
            CK(gbp_ba_iterate(h, target - done, /*robustify=*/1, /*local_relin=*/1));
            done = target;
        }
        if (done == mark && mark != iters) CK(gbp_ba_fill_iters(h, 1));
    }
    CK(gbp_ba_metrics(h, m));
    if (rank == 0) printf("after %3d ARE %.12f  energy %.9f  relinearising %d  (%d ranks, NCCL %d)\n", iters, m[0] / (double)F, m[1], (int)(m[2] + 0.5), nranks, gbp_comm_version());
    std::vector<double> mu(6 * C);
    CK(gbp_ba_read(h, GBP_F_CAM_MU, mu.data(), mu.size() * sizeof(double)));
    if (rank == 0) printf("keyframe 0 mean %.12f %.12f %.12f %.12f %.12f %.12f\n", mu[0], mu[1], mu[2], mu[3], mu[4], mu[5]);

    CK(gbp_ba_destroy(h));
    CK(gbp_comm_destroy(comm));
    if (rank == 0) remove(id_file.c_str());
    return 0;
}
