// gbp_resident.cuh -- gbp_ba_iterate(n) of an L2-resident graph in ONE launch (no reference counterpart: this is how
// the loop of ba.py:84-105 runs on the device when nobody reads between the sweeps).
//
// A graph like fr1desk (13 k factors, 10 MB of state) is latency-bound, not bandwidth-bound: the two-kernel iteration
// (sweep_kernel, belief_kernel) costs ~9 us, almost all of it launch gaps and chains of dependent L2 round trips
// (tile descriptor -> landmark index -> belief row; CSR pointer -> slot list -> message rows).  This kernel keeps the
// whole call on the SMs:
//   * persistent: one warp per tile for the whole call (cooperative launch, every CTA co-resident).  The tile's message
//     rows, linearisation points, measurement, counters and the linearisation itself (J, h(x0), b -- a pure function of
//     the linearisation point, recomputed only when the factor relinearises) stay in shared memory / registers across
//     iterations;
//   * DATAFLOW instead of barriers.  An iteration has the two phases of synchronous_iteration (gbp/gbp.py:86-92):
//     A  every tile sweeps its edges (messages) and PUBLISHES the 9-double message row of every edge to its landmark and
//        the 27 sums of the tile's messages to its keyframe;
//     B  every landmark / keyframe has ONE owner warp (dealt by the host, balanced by rows) that gathers those rows, adds
//        the prior, solves for the mean and PUBLISHES the belief row, which phase A of the next iteration gathers.
//     Nobody waits for the whole grid: a published row carries its iteration number in every 8-byte word -- a double is
//     stored as two words {low half | epoch}, {high half | epoch}; aligned 8-byte stores are single-copy atomic, so a
//     reader that sees the epoch sees the data (the scheme of NCCL's LL protocol) -- and a consumer polls the LAST word
//     of the rows it needs and then loads them.  No fence, no atomic, no grid barrier (measured on B200: a grid barrier
//     costs ~3.5 k cycles, a kernel boundary inside a CUDA graph about the same, one tagged hand-over between two SMs 600).
//     Every row is read exactly once per iteration; a producer cannot overwrite a row before its only consumer has used it,
//     because it needs that consumer's output to get there (so one buffer suffices);
//   * CONSUMER-ORDERED layouts.  One warp alone needs 1870 cycles to pull 32 scattered rows of 12 tagged words from L2 (the
//     load/store unit handles a warp's distinct sectors one after the other, ~4.5 cycles each; measured, scripts/micro), but
//     ~450 when the words are contiguous.  Loads are on the critical path, stores are not, so everything is published where
//     its reader wants it: message rows at the landmark-CSR position (an owner's rows are one contiguous block), tile sums in
//     keyframe-CSR order, and the landmark belief once per EDGE, in slot order (a tile's 32 belief rows are one block).
//     Blocks are loaded word-striped (lane i takes words i, i + 32, ...: every instruction is one 512 B run) and transposed
//     through shared memory.
// Summation orders are fixed (run-to-run deterministic); the two-kernel path agrees to rounding, not bitwise (the inlined
// linearisation contracts to different FMAs in different kernels).  Iteration 0 reads the belief ARRAYS like sweep_kernel
// does (they are what the client may have written); the belief arrays, means and keyframe partial sums are refreshed once
// at the end of the call by belief_kernel from the final messages.  Every wait has a timeout (~1 s) that raises the error
// flag instead of hanging the GPU.
#pragma once
#include "gbp_kernels.cuh"

namespace gbp {

constexpr int RES_CHUNK = 8;          // rows per chunk of a landmark's message list (one lane sums one chunk)
constexpr int RES_MAX_WARPS = 8;      // tiles (warps) per CTA
constexpr int RES_MAX_B_ROWS = 192;   // message rows a warp may own in phase B
constexpr int RES_MAX_CAM_TILES = 64; // tiles of a keyframe
constexpr int RES_BL_PITCH = 13;      // doubles per staged landmark belief row (12 + 1: conflict-free row-per-lane reads)

typedef ulonglong2 LLWord;            // one published double: {low 32 bits | epoch << 32, high 32 bits | epoch << 32}

struct ResidentTile {                 // per warp = per tile
    int smem_off;                     // byte offset of the warp's block inside its CTA's dynamic shared memory
    int b_q0, b_rows;                 // phase B: the CSR range [b_q0, b_q0 + b_rows) of lmk_slots whose message rows this warp owns
    int b_cam;                        // phase B: the keyframe this warp owns, -1 = none
    int b_nchunks;                    // lanes with a chunk in phase B
    int cam_pos;                      // position of this tile in cam_tiles (where its sums are published)
};

struct ResidentParams {
    SweepParams sweep;                // canonical arrays: msg_cam, msg_lmk, linpoint, iters, flags, sigma2a, tile_partial, beliefs
    LLWord* pub_mq;                   // [F][9]       messages to the landmarks, at the edge's position in the landmark CSR
    LLWord* pub_tq;                   // [tiles][27]  tile sums of the messages to the keyframe, in cam_tiles order
    LLWord* pub_be;                   // [slots][12]  landmark beliefs eta | Lambda | mu, one copy per edge, slot order
    LLWord* pub_cb;                   // [C][33]      keyframe beliefs
    const double* lmk_prior;
    const double* cam_prior;
    const int* lmk_slots;             // CSR by landmark over slots (belief_kernel's table)
    const int* csr_pos;               // [slots]      its inverse: position of every edge in that CSR
    const int* cam_tile_ptr;          // CSR by keyframe over tiles
    const ResidentTile* tile_info;    // [tiles]
    const int* b_chunks;              // [tiles][32][2]  phase B role of every lane: {first row << 16 | rows << 12 | chunk index << 6 |
                                      //                 chunks of the landmark, landmark}; rows = 0: the lane has no chunk
    const int* cta_tiles;             // [grid][warps]   tile of every warp of every CTA, -1 = none
    double* error_flag;               // != 0: a wait timed out (the state is invalid)
    long long* dbg;                   // profiling builds: [tiles][32] time stamps (global timer)
    unsigned int epoch0;              // epoch of the first iteration of this call (epochs never repeat on a graph)
    int n_iters;
};

// bytes of the warp's block of shared memory; every array inside starts 16 B aligned
__host__ __device__ inline size_t resident_tile_smem(int b_rows, int b_cam_tiles) {
    size_t d = (size_t)32 * (CAM_M + LMK_M + 9) + 34 + 32 * LMK_M + 32 * RES_BL_PITCH + 32 * LMK_B
             + (size_t)b_rows * LMK_M + (size_t)b_cam_tiles * CAM_M;
    d = (d + 1) & ~size_t(1);                                                                  // doubles (even count)
    size_t i = (size_t)2 * ((b_rows + 3) & ~3) + LMK_M * 32 + (size_t)2 * ((b_rows * LMK_B + 3) & ~3);   // ints
    return d * 8 + i * 4;
}

__device__ __forceinline__ LLWord ll_pack(double d, unsigned int tag) {
    const unsigned long long bits = (unsigned long long)__double_as_longlong(d), t = (unsigned long long)tag << 32;
    return make_ulonglong2((bits & 0xffffffffull) | t, (bits >> 32) | t);
}
__device__ __forceinline__ bool ll_ok(const LLWord& v, unsigned int tag) {
    return (unsigned int)(v.x >> 32) == tag && (unsigned int)(v.y >> 32) == tag;
}
__device__ __forceinline__ double ll_value(const LLWord& v) {
    return __longlong_as_double((long long)((v.x & 0xffffffffull) | (v.y << 32)));
}
__device__ __forceinline__ void ll_store(LLWord* dst, double d, unsigned int tag) { __stcg(dst, ll_pack(d, tag)); }
__device__ __forceinline__ LLWord ll_load_cg(const LLWord* p) {    // bulk loads: L2 only (2.5 instead of 4 cycles per request); tags verify
    LLWord v;
    asm volatile("ld.global.cg.v2.u64 {%0, %1}, [%2];" : "=l"(v.x), "=l"(v.y) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ LLWord ll_load(const LLWord* p) {       // from L2, never L1: the word is written by another SM
    LLWord v;
    asm volatile("ld.relaxed.gpu.global.v2.u64 {%0, %1}, [%2];" : "=l"(v.x), "=l"(v.y) : "l"(p) : "memory");
    return v;
}

// Spin bookkeeping of the waits: ~1 s without progress raises the error flag (a producer is missing: do not hang the GPU), and
// everybody leaves as soon as the flag is up.  Warp-collective; returns true when the warp has to give up.
struct SpinGuard {
    long long t0;
    int round;
    __device__ __forceinline__ SpinGuard() : t0(0), round(0) {}
    __device__ __forceinline__ bool expired(double* error_flag) {
        if (round == 0) t0 = clock64();
        if ((++round & 255) != 0) return false;
        bool give_up = *reinterpret_cast<volatile double*>(error_flag) != 0.0;
        if (!give_up && clock64() - t0 > 2000000000LL) {
            *reinterpret_cast<volatile double*>(error_flag) = 1.0;
            __threadfence();
            give_up = true;
        }
        return __any_sync(0xffffffffu, give_up) != 0;
    }
};

// A contiguous block of n_rows tagged rows of U words -> plain doubles in shared memory (row pitch PITCH doubles).
//   wait:  lane r polls the LAST word of row r0 + r (rounds of 32 rows) until it carries `tag`;
//   load:  word-striped (lane i takes words i, i + 32, ...: every instruction reads one contiguous 512 B run), every tag
//          verified in registers, repeated as a whole in the rare case that a word of a row was not visible yet.
// Warp-collective; false = gave up.  MAXW = upper bound of the block's words / 32 (compile-time unrolling).
template <int U, int PITCH>
__device__ __forceinline__ bool ll_read_block(const LLWord* block, int n_rows, unsigned int tag, double* s_out, int lane, double* error_flag) {
    SpinGuard guard;
    // The wait proper costs ONE request per round: lane 0 polls the last word of the block's last row.  (The SM's request path
    // takes ~4 cycles per distinct line: three warps polling 32 rows each would keep it busy and delay everybody's real loads.)
    for (;;) {
        bool ready = true;
        if (lane == 0 && n_rows > 0) ready = ll_ok(ll_load(block + (size_t)(n_rows - 1) * U + (U - 1)), tag);
        if (__shfl_sync(0xffffffffu, (int)ready, 0)) break;
        if (guard.expired(error_flag)) return false;
    }
    // Then the whole block, word-striped; the rows come from different producers, so some may still be on their way: every tag
    // is verified in registers and the block is simply read again until all of it carries the epoch.
    const int n_words = n_rows * U;
    for (;;) {
        bool good = true;
#pragma unroll 4
        for (int f = lane; f < n_words; f += 32) {
            const LLWord w = ll_load_cg(block + f);
            good &= ll_ok(w, tag);
            const int row = f / U;
            s_out[row * PITCH + (f - row * U)] = ll_value(w);
        }
        if (__all_sync(0xffffffffu, good)) break;
        if (guard.expired(error_flag)) return false;
    }
    __syncwarp();
    return true;
}

// The per-edge step with the linearisation cached in registers.  Same arithmetic, in the same order, as edge_sweep<ROBUST,
// false> (gbp_edge.cuh): J, h0 = linearise(linpoint) and b = J x0 + z - h0 are pure functions of the stored linearisation
// point, so they are recomputed only when the factor relinearises.  it / fl / var are updated in r (written back to
// global memory once, at the end of the call).
template <bool ROBUST>
__device__ __forceinline__ bool edge_step_cached(const SweepParams& p, EdgeRegs& r, const double* s_cb, double* my_lp, double* my_mc,
                                                 double* my_ml, double* J, double* h0, double* b) {
    const double* z = r.z;
    const double* bl = r.bl;
    int it = r.it, fl = r.fl;
    bool relin = false;
    double x0[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) x0[k] = my_lp[k];
    if (p.stages & ST_RELIN) {
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
        relin = (d2 > p.beta * p.beta) && (it >= p.min_linear);
    }
    double var = p.var0;
    if (ROBUST) {
        var = r.var;
        if (p.stages & ST_ROBUSTIFY) {       // h at the STORED linearisation point (gbp/gbp.py:309-312) = the cached h0
            bool rf;
            var = robust_variance(p.loss, p.var0, p.nstds, z[0] - h0[0], z[1] - h0[1], &rf);
            fl = rf ? (fl | 2) : (fl & ~2);
            r.var = var;
        }
    }
    if (p.stages & ST_RELIN) {
        if (relin) {
#pragma unroll
            for (int k = 0; k < 6; ++k) x0[k] = s_cb[27 + k];
#pragma unroll
            for (int k = 0; k < 3; ++k) x0[6 + k] = bl[9 + k];
#pragma unroll
            for (int k = 0; k < 9; ++k) my_lp[k] = x0[k];
            it = 0;
            fl &= ~1;
            linearise(p.K, x0, J, h0);
            factor_rhs(J, x0, z, h0, b);
        } else {
            it += 1;
        }
    }
    if (p.stages & ST_MESSAGES) {
        double damping = p.eta_damping;
        if (p.stages & ST_LOCAL_DAMPING) {
            if (it == p.num_undamped) fl |= 1;
            damping = (fl & 1) ? p.eta_damping : 0.0;
        }
        double nl_eta[3], nl_lam[6];
        {
            double P[21], ev[6];
#pragma unroll
            for (int k = 0; k < 6; ++k) ev[k] = s_cb[k] - my_mc[k];
#pragma unroll
            for (int k = 0; k < 21; ++k) P[k] = s_cb[6 + k] - my_mc[6 + k];
            message<3, 6>(J + 6, J, b, var, P, ev, damping, my_ml, nl_eta, nl_lam);
        }
        {
            double P[6], ev[3];
#pragma unroll
            for (int k = 0; k < 3; ++k) ev[k] = bl[k] - my_ml[k];
#pragma unroll
            for (int k = 0; k < 6; ++k) P[k] = bl[3 + k] - my_ml[3 + k];
            message<6, 3>(J, J + 6, b, var, P, ev, damping, my_mc, my_mc, my_mc + 6);
        }
#pragma unroll
        for (int k = 0; k < 3; ++k) my_ml[k] = nl_eta[k];
#pragma unroll
        for (int k = 0; k < 6; ++k) my_ml[3 + k] = nl_lam[k];
    }
    r.it = it;
    r.fl = fl;
    return relin;
}

// column `col` of the tile's messages to its keyframe, summed over the rows: four interleaved partial sums (rows q = 0, 1, 2, 3 mod 4;
// the 32-deep dependent chain of shared-memory loads and adds is the longest part of publishing), combined (a0 + a1) + (a2 + a3)
__device__ __forceinline__ double tile_sum(const double* s_mc, int n, int col) {
    double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
    int q = 0;
    for (; q + 4 <= n; q += 4) {
        a0 += s_mc[q * CAM_M + col];
        a1 += s_mc[(q + 1) * CAM_M + col];
        a2 += s_mc[(q + 2) * CAM_M + col];
        a3 += s_mc[(q + 3) * CAM_M + col];
    }
    if (q < n) a0 += s_mc[q * CAM_M + col];
    if (q + 1 < n) a1 += s_mc[(q + 1) * CAM_M + col];
    if (q + 2 < n) a2 += s_mc[(q + 2) * CAM_M + col];
    return (a0 + a1) + (a2 + a3);
}

#ifdef GBP_RESIDENT_PROFILE
__device__ __forceinline__ long long gtimer() { long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
#define RES_STAMP(j) do { if (lane == 0 && k >= 50 && k < 54) rp.dbg[(long long)tile * 32 + (k - 50) * 8 + (j)] = gtimer(); } while (0)
#define RES_T(i) do { const long long _t = clock64(); if (k > 0) prof[i] += _t - tprev; tprev = _t; } while (0)
#else
#define RES_T(i) do { } while (0)
#define RES_STAMP(j) do { } while (0)
#endif

template <bool ROBUST>
__global__ void __launch_bounds__(32 * RES_MAX_WARPS, 1) resident_kernel(const ResidentParams rp) {
#ifdef GBP_RESIDENT_PROFILE
    long long prof[6] = {0, 0, 0, 0, 0, 0}, tprev = clock64();
#endif
    extern __shared__ __align__(16) unsigned char rsm[];
    const SweepParams& p = rp.sweep;
    const int W = (int)(blockDim.x >> 5);
    const int warp = (int)(threadIdx.x >> 5), lane = (int)(threadIdx.x & 31);
    const int tile = rp.cta_tiles[(int)blockIdx.x * W + warp];
    if (tile < 0) return;                                   // no CTA-wide synchronisation anywhere below

    const Tile tl = p.tiles[tile];
    const ResidentTile ti = rp.tile_info[tile];
    const int n = tl.count;
    const int bcam = ti.b_cam;
    const int bct0 = bcam >= 0 ? rp.cam_tile_ptr[bcam] : 0;
    const int bncl = bcam >= 0 ? rp.cam_tile_ptr[bcam + 1] - bct0 : 0;
    const int brows = ti.b_rows;

    unsigned char* mine = rsm + ti.smem_off;
    double* s_mc = reinterpret_cast<double*>(mine);       // [32][27]  messages to the keyframe
    double* s_ml = s_mc + 32 * CAM_M;                     // [32][9]   messages to the landmarks
    double* s_lp = s_ml + 32 * LMK_M;                     // [32][9]   linearisation points
    double* s_cb = s_lp + 32 * 9;                         // [34]      keyframe belief row
    double* s_bsum = s_cb + 34;                           // [32][9]   phase B: chunk sums
    double* s_abl = s_bsum + 32 * LMK_M;                  // [32][13]  phase A: landmark belief rows of the tile's edges
    double* s_bel = s_abl + 32 * RES_BL_PITCH;            // [32][12]  phase B: belief rows of the owned landmarks (by leader lane)
    double* s_brow = s_bel + 32 * LMK_B;                  // [brows][9]  phase B: message rows of the owned landmarks
    double* s_ctp = s_brow + (size_t)brows * LMK_M;       // [bncl][27]  phase B: tile sums of the owned keyframe
    int* s_bslot = reinterpret_cast<int*>(mine + ((((size_t)32 * (CAM_M + LMK_M + 9) + 34 + 32 * LMK_M + 32 * RES_BL_PITCH + 32 * LMK_B
                                                     + (size_t)brows * LMK_M + (size_t)bncl * CAM_M + 1) & ~size_t(1)) * 8));   // [brows] slot of every owned row
    int* s_rowlm = s_bslot + ((brows + 3) & ~3);          // [brows]   leader lane of the landmark of every owned row
    int* s_pubdst = s_rowlm + ((brows + 3) & ~3);         // [9][32]   phase A: where word f = j * 32 + lane of the tile's message rows goes
    int* s_bedst = s_pubdst + LMK_M * 32;                 // [brows * 12, padded] phase B: destination word of belief-copy word f ...
    int* s_besrc = s_bedst + (((brows * LMK_B) + 3) & ~3);   //               ... and its source in s_bel

    const long long base = (long long)tile * 32;
    const long long e = base + lane;
    EdgeRegs r;
    r.it = -1; r.fl = 0; r.var = p.var0; r.z[0] = r.z[1] = 0.0;
    double J[18], h0[2], b[2];
    int lmk = 0, my_q = 0;
    bool ever_relin = false;

    // ---- set-up: the tile's state, the phase-B tables and priors
    for (int i = lane; i < n * CAM_M; i += 32) s_mc[i] = p.msg_cam[base * CAM_M + i];
    for (int i = lane; i < n * LMK_M; i += 32) {
        s_ml[i] = p.msg_lmk[base * LMK_M + i];
        s_lp[i] = p.linpoint[base * 9 + i];
    }
    for (int i = lane; i < brows; i += 32) s_bslot[i] = rp.lmk_slots[ti.b_q0 + i];
    if (lane < n) {
        lmk = load_edge_scalars(p, e, r);
        my_q = rp.csr_pos[e];
    }
    for (int row = 0; row < n; ++row) {                     // destination of every word of the tile's message rows (static)
        const int q = __shfl_sync(0xffffffffu, my_q, row);
        if (lane < LMK_M) s_pubdst[row * LMK_M + lane] = q * LMK_M + lane;
    }
    // phase B role of this lane: a chunk of <= 8 consecutive rows of one owned landmark (b_row0 = first row in s_brow, b_cnt rows;
    // the lane of a landmark's chunk 0 is its leader: it holds the prior and finishes the belief)
    const int bdesc = rp.b_chunks[((size_t)tile * 32 + lane) * 2], b_lmk = rp.b_chunks[((size_t)tile * 32 + lane) * 2 + 1];
    const int b_row0 = bdesc >> 16, b_cnt = (bdesc >> 12) & 15, b_nch = bdesc & 63;
    const bool b_leader = b_cnt > 0 && ((bdesc >> 6) & 63) == 0;
    if (b_cnt > 0) {
        const int leader = lane - ((bdesc >> 6) & 63);
        for (int h = 0; h < b_cnt; ++h) s_rowlm[b_row0 + h] = leader;
    }
    __syncwarp();
    for (int f = lane; f < brows * LMK_B; f += 32) {        // belief copies: one per owned edge row, word-striped (static)
        const int row = f / LMK_B, word = f - row * LMK_B;
        s_bedst[f] = s_bslot[row] * LMK_B + word;
        s_besrc[f] = s_rowlm[row] * LMK_B + word;
    }
    double bprior[LMK_M];
#pragma unroll
    for (int j = 0; j < LMK_M; ++j) bprior[j] = b_leader ? rp.lmk_prior[(long long)b_lmk * LMK_M + j] : 0.0;
    const double cprior = (bcam >= 0 && lane < CAM_M) ? rp.cam_prior[(long long)bcam * CAM_M + lane] : 0.0;
    __syncwarp();
    if (lane < n) {
        double x0[9];
#pragma unroll
        for (int k = 0; k < 9; ++k) x0[k] = s_lp[lane * 9 + k];
        linearise(p.K, x0, J, h0);
        factor_rhs(J, x0, r.z, h0, b);
    }

    bool ok = true;
    for (int k = 0; k < rp.n_iters && ok; ++k) {
        const unsigned int epoch = rp.epoch0 + (unsigned int)k;
        RES_T(5);
        // ================= phase A: beliefs of iteration k-1 -> messages of iteration k
        if (k == 0) {
            // beliefs as stored (what sweep_kernel reads): they may have been written by the client
            for (int i = lane; i < CAM_B; i += 32) s_cb[i] = p.cam_belief[(long long)tl.cam * CAM_B + i];
            if (lane < n) {
                const double2* src = reinterpret_cast<const double2*>(p.lmk_belief + (long long)lmk * LMK_B);
#pragma unroll
                for (int q = 0; q < LMK_B / 2; ++q) {
                    const double2 v = src[q];
                    r.bl[2 * q] = v.x;
                    r.bl[2 * q + 1] = v.y;
                }
            }
        } else {
            // the keyframe's belief row: lane 0 polls its last word (mu[5], written last), then lane i reads word i
            const LLWord* cbw = rp.pub_cb + (long long)tl.cam * CAM_B;
            SpinGuard guard;
            for (;;) {
                bool ready = true;
                if (lane == 0) ready = ll_ok(ll_load(cbw + 32), epoch - 1u);
                if (__shfl_sync(0xffffffffu, (int)ready, 0)) break;
                if (guard.expired(rp.error_flag)) { ok = false; break; }
            }
            while (ok) {
                const LLWord w = ll_load(cbw + lane), w32 = ll_load(cbw + 32);
                if (__all_sync(0xffffffffu, ll_ok(w, epoch - 1u) && ll_ok(w32, epoch - 1u))) {
                    s_cb[lane] = ll_value(w);
                    if (lane == 0) s_cb[32] = ll_value(w32);
                    break;
                }
                if (guard.expired(rp.error_flag)) ok = false;
            }
            // the belief rows of the tile's landmarks: one contiguous block (their owners publish a copy per edge, slot order)
            if (ok) ok = ll_read_block<LMK_B, RES_BL_PITCH>(rp.pub_be + base * LMK_B, n, epoch - 1u, s_abl, lane, rp.error_flag);
            if (!ok) break;
            if (lane < n) {
#pragma unroll
                for (int q = 0; q < LMK_B; ++q) r.bl[q] = s_abl[lane * RES_BL_PITCH + q];
            }
        }
        __syncwarp();       // s_cb complete
        RES_T(0); RES_STAMP(0);
        bool relin = false;
        if (lane < n) relin = edge_step_cached<ROBUST>(p, r, s_cb, s_lp + lane * 9, s_mc + lane * CAM_M, s_ml + lane * LMK_M, J, h0, b);
        ever_relin |= relin;
        __syncwarp();       // new rows in shared memory
        RES_T(1); RES_STAMP(1);
        if (k + 1 == rp.n_iters) break;       // the final beliefs are belief_kernel's job (from the canonical arrays written below)
        // ---- publish: the tile's sums first (the keyframe's path is the longer one), then the message row of every edge at its
        //      place in the landmark CSR (scattered stores: nobody waits for a store)
        if (lane < CAM_M) ll_store(rp.pub_tq + (long long)ti.cam_pos * CAM_M + lane, tile_sum(s_mc, n, lane), epoch);
        // (word-striped: 9 consecutive lanes write one 144 B row -- ~5 lines per instruction instead of 32)
#pragma unroll
        for (int j = 0; j < LMK_M; ++j) {
            const int f = j * 32 + lane;
            if (f < n * LMK_M) ll_store(rp.pub_mq + s_pubdst[f], s_ml[f], epoch);
        }
        RES_T(2); RES_STAMP(2);
        // ================= phase B: messages of iteration k -> beliefs of iteration k (owned landmarks / keyframe)
        if (bcam >= 0) {                        // owned keyframe: tile sums in cam_tiles order, + prior, 6x6 solve on lane 0
            ok = ll_read_block<CAM_M, CAM_M>(rp.pub_tq + (long long)bct0 * CAM_M, bncl, epoch, s_ctp, lane, rp.error_flag);
            if (!ok) break;
            RES_STAMP(4);
            double acc = 0.0;
            if (lane < CAM_M) {
                for (int q = 0; q < bncl; ++q) acc += s_ctp[q * CAM_M + lane];
                acc += cprior;
                ll_store(rp.pub_cb + (long long)bcam * CAM_B + lane, acc, epoch);
            }
            double v[CAM_M];
#pragma unroll
            for (int j = 0; j < CAM_M; ++j) v[j] = __shfl_sync(0xffffffffu, acc, j);
            if (lane == 0) {
                double mu[6];
                spd_solve<6>(v + 6, v, mu);
#pragma unroll
                for (int j = 0; j < 6; ++j) ll_store(rp.pub_cb + (long long)bcam * CAM_B + 27 + j, mu[j], epoch);
            }
        }
        if (brows > 0) {
            ok = ll_read_block<LMK_M, LMK_M>(rp.pub_mq + (long long)ti.b_q0 * LMK_M, brows, epoch, s_brow, lane, rp.error_flag);
            if (!ok) break;
            RES_T(3); RES_STAMP(3);
            if (b_cnt > 0) {                    // chunk sum, rows left to right
                double sum[LMK_M];
#pragma unroll
                for (int j = 0; j < LMK_M; ++j) sum[j] = 0.0;
#pragma unroll
                for (int h = 0; h < RES_CHUNK; ++h) {
                    if (h < b_cnt) {
#pragma unroll
                        for (int j = 0; j < LMK_M; ++j) sum[j] += s_brow[(size_t)(b_row0 + h) * LMK_M + j];
                    }
                }
#pragma unroll
                for (int j = 0; j < LMK_M; ++j) s_bsum[lane * LMK_M + j] = sum[j];
            }
            __syncwarp();
            if (b_leader) {                     // chunk sums left to right, + prior, 3x3 solve
                double acc[LMK_M];
#pragma unroll
                for (int j = 0; j < LMK_M; ++j) acc[j] = s_bsum[lane * LMK_M + j];
                for (int c = 1; c < b_nch; ++c) {
#pragma unroll
                    for (int j = 0; j < LMK_M; ++j) acc[j] += s_bsum[(lane + c) * LMK_M + j];
                }
#pragma unroll
                for (int j = 0; j < LMK_M; ++j) acc[j] += bprior[j];
                double mu[3];
                spd_solve<3>(acc + 3, acc, mu);
#pragma unroll
                for (int j = 0; j < LMK_M; ++j) s_bel[lane * LMK_B + j] = acc[j];
                s_bel[lane * LMK_B + 9] = mu[0]; s_bel[lane * LMK_B + 10] = mu[1]; s_bel[lane * LMK_B + 11] = mu[2];
            }
            __syncwarp();
            // publish: one copy of the belief row per EDGE of the landmark, at the edge's slot (12 consecutive lanes write one
            // 192 B row; the last word of a row, mu[2], is stored by the highest lane of its run = issued last)
            const int n_words = brows * LMK_B;
#pragma unroll 4
            for (int f = lane; f < n_words; f += 32) ll_store(rp.pub_be + s_bedst[f], s_bel[s_besrc[f]], epoch);
        }
        RES_T(4); RES_STAMP(5);
    }
#ifdef GBP_RESIDENT_PROFILE
    if (lane == 0 && (tile % 64 == 0 || tile == p.n_tiles - 1 || bcam == 0))
        printf("resident profile tile %d (n %d, owns %d rows, keyframe %d with %d tiles) cycles/iter: A wait+gather %lld  edge %lld  publish %lld  B wait+gather %lld  B sums+solves+publish %lld  other %lld\n",
               tile, n, brows, bcam, bncl, prof[0] / (rp.n_iters - 1), prof[1] / (rp.n_iters - 1), prof[2] / (rp.n_iters - 1), prof[3] / (rp.n_iters - 1),
               prof[4] / (rp.n_iters - 1), prof[5] / (rp.n_iters - 1));
#endif

    // ---- the tile's private state back to the canonical arrays (nobody reads them during the call)
    if (ok) {
        for (int i = lane; i < n * CAM_M; i += 32) p.msg_cam[base * CAM_M + i] = s_mc[i];
        for (int i = lane; i < n * LMK_M; i += 32) p.msg_lmk[base * LMK_M + i] = s_ml[i];
        if (__any_sync(0xffffffffu, ever_relin))
            for (int i = lane; i < n * 9; i += 32) p.linpoint[base * 9 + i] = s_lp[i];
        if (lane < CAM_M) p.tile_partial[(long long)tile * CAM_M + lane] = tile_sum(s_mc, n, lane);
        if (lane < n) {
            p.iters[e] = r.it;
            p.flags[e] = r.fl;
            if (ROBUST) p.sigma2a[e] = r.var;
        }
    }
}

}  // namespace gbp
