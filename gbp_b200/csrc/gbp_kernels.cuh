// gbp_kernels.cuh -- sm_100a kernels of the GBP bundle-adjustment sweep.
//
// Data layout in HBM (all float64 unless noted; "slot" = position of an edge in the
// engine's storage order, tile t owns slots [t*T, t*T + count)):
//   msg_cam   [slots][27]  factor->keyframe message  eta[6] | Lambda packed[21]
//             [slots][18]  ... or with the rank-2 precision factored: eta[6] | W[2][6], Lambda = W^T W (the streaming build, kernel_variant 2)
//   msg_lmk   [slots][9]   factor->landmark message  eta[3] | Lambda packed[6]
//   linpoint  [slots][9]   linearisation point [t, w, y]
//   z         [slots][2]   measurement
//   lmk_idx   [slots] i32, iters [slots] i32, flags [slots] i32, sigma2a [slots]
//   cam_belief[C][33]      eta[6] | Lambda[21] | mu[6]      (264 B rows)
//   lmk_belief[L][12]      eta[3] | Lambda[6]  | mu[3]      (96 B rows = 3 sectors, 32 B aligned)
//   cam_prior [C][27], lmk_prior[L][9]
//   tile_partial[tiles][27] per-tile sum of new factor->keyframe messages, stored KEYFRAME-MAJOR (row tile_pos[tile]): the tiles
//             of a keyframe (chunk by chunk) are contiguous, so the belief update reads them without an index indirection
// Every tile belongs to exactly ONE keyframe, so the keyframe belief is a CTA-uniform
// broadcast and the keyframe-side sum is a plain per-tile column sum (no atomics,
// deterministic).  Tiles are ordered landmark-block-major so that the 96 B landmark
// belief rows gathered by a wave of CTAs stay L2-resident.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "gbp_edge.cuh"
#include "gbp_math.cuh"

namespace gbp {

// cooperative, coalesced copy of n doubles (16 B vectors where possible; both pointers 16 B aligned)
template <int T>
__device__ __forceinline__ void coop_copy(double* __restrict__ dst, const double* __restrict__ src, int n) {
    const int n2 = n >> 1;
    const double2* s2 = reinterpret_cast<const double2*>(src);
    double2* d2 = reinterpret_cast<double2*>(dst);
    for (int i = threadIdx.x; i < n2; i += T) d2[i] = s2[i];
    if ((n & 1) && threadIdx.x == 0) dst[n - 1] = src[n - 1];
}

// ----------------------------------------------------------------------------------------
// TMA (bulk async copy) + mbarrier primitives, sm_90+ PTX.  The tile's contiguous row blocks
// are moved global <-> shared by the copy engine (SASS: UBLKCP), no registers, no LSU issue.
// ----------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred P1;\n\t"
        "GBP_WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
        "@P1 bra GBP_DONE_%=;\n\t"
        "bra GBP_WAIT_%=;\n\t"
        "GBP_DONE_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void bulk_s2g(void* gmem_dst, const void* smem_src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gmem_dst), "r"(smem_u32(smem_src)), "r"(bytes) : "memory");
}
// L2 eviction policies: the message / linearisation-point rows are touched once per iteration
// (evict_first), the landmark belief rows are re-read by other tiles of the same landmark block
// (evict_last), so 7 GB of streaming rows do not push the 24 MB of hot belief rows out of L2.
__device__ __forceinline__ uint64_t policy_evict_first() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ uint64_t policy_evict_last() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ void bulk_g2s_hint(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar, uint64_t pol) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar)), "l"(pol)
                 : "memory");
}
__device__ __forceinline__ void bulk_s2g_hint(void* gmem_dst, const void* smem_src, uint32_t bytes, uint64_t pol) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;" ::"l"(gmem_dst), "r"(smem_u32(smem_src)),
                 "r"(bytes), "l"(pol)
                 : "memory");
}
__device__ __forceinline__ double2 ldg_hint(const double2* ptr, uint64_t pol) {
    double2 v;
    asm volatile("ld.global.nc.L2::cache_hint.v2.f64 {%0, %1}, [%2], %3;" : "=d"(v.x), "=d"(v.y) : "l"(ptr), "l"(pol));
    return v;
}
// L2 prefetch of a contiguous block (16 B aligned, multiple of 16 B): no shared memory, no completion to wait for
__device__ __forceinline__ void bulk_prefetch_l2(const void* gmem_src, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(gmem_src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }



// scalars of the edge (none of them is written by belief_kernel); returns the landmark index for the gather
__device__ __forceinline__ int load_edge_scalars(const SweepParams& p, long long e, EdgeRegs& r) {
    const int lmk = p.lmk_idx[e];
    r.it = p.iters[e];
    r.fl = p.flags[e];
    r.var = (p.loss != 0) ? p.sigma2a[e] : p.var0;
    const double2 zz = reinterpret_cast<const double2*>(p.z)[e];
    r.z[0] = zz.x;
    r.z[1] = zz.y;
    return lmk;
}

// the dependent gather: 96 B landmark belief row (six 16 B loads)
__device__ __forceinline__ void gather_lmk_belief(const SweepParams& p, int lmk, EdgeRegs& r) {
    const double2* src = reinterpret_cast<const double2*>(p.lmk_belief + (long long)lmk * LMK_B);
    const uint64_t pol = policy_evict_last();
#pragma unroll
    for (int k = 0; k < LMK_B / 2; ++k) {
        const double2 v = ldg_hint(src + k, pol);
        r.bl[2 * k] = v.x;
        r.bl[2 * k + 1] = v.y;
    }
}

__device__ __forceinline__ void load_edge_regs(const SweepParams& p, long long e, EdgeRegs& r) {
    const int lmk = load_edge_scalars(p, e, r);
    gather_lmk_belief(p, lmk, r);
}



// column sums of the tile's (new) messages to its keyframe -> tile_partial[tpos][27]
template <int T>
__device__ __forceinline__ void tile_column_sums(const SweepParams& p, int tpos, int n, const double* s_mc, double* s_red) {
    constexpr int G = T / 32;
    const int tid = threadIdx.x;
    if (tid < CAM_M * G) {
        const int col = tid % CAM_M, g = tid / CAM_M;
        const int r1 = min(n, (g + 1) * 32);
        double acc = 0.0;
        for (int r = g * 32; r < r1; ++r) acc += s_mc[r * CAM_M + col];
        s_red[g * CAM_M + col] = acc;
    }
    __syncthreads();
    if (tid < CAM_M) {
        double acc = s_red[tid];
#pragma unroll
        for (int g = 1; g < G; ++g) acc += s_red[g * CAM_M + tid];
        p.tile_partial[(long long)tpos * CAM_M + tid] = acc;
    }
}


// ----------------------------------------------------------------------------------------
// K1-K3 (+ the keyframe half of K4): robustify -> relinearise -> factor-to-variable messages
// -> per-tile sum of the messages to the keyframe.   One CTA per tile, one thread per edge.
// Replaces gbp/gbp.py:82-84,296-332 / 64-80,267-294 / 46-54,334-373 for reprojection factors.
//
// The tile's row blocks are staged by TMA bulk copies (one elected thread, mbarrier completion); every
// thread issues its gathers (landmark index -> 96 B belief row) BEFORE waiting, so all of a tile's DRAM
// round trips overlap; new rows leave by bulk stores.  Compiled for 384 resident threads per SM
// (<= 170 registers, no spills; a 128-register build for 16 warps per SM was measured slower).
//
// STREAM = false  graphs that live in L2 (up to 8192 tiles): full 27-double keyframe message rows, the bulk
//                 copies are sized by the tile descriptor.
// STREAM = true   HBM-bound graphs: (i) the keyframe messages live in HBM with their rank-2 precision FACTORED
//                 (18 doubles per row instead of 27: 144 B less traffic per edge and sweep), the threads expand
//                 them into s_full for the keyframe-side sum; (ii) EARLY ISSUE: a tile owns T slots (padding
//                 slots hold landmark 0, zero rows, iters = -1), so the bulk loads fetch all T rows and every
//                 thread loads its scalars and gathers its landmark row unconditionally at once; the descriptor
//                 (count, keyframe) arrives in parallel and is only needed for the keyframe row and the stores
//                 (dependent DRAM round trips before the edge code: 3 -> 2); (iii) FAR-AHEAD L2 PREFETCH of the
//                 streams of tile + pf_dist.  Measured on the 10 M-factor graph: 1.257 -> 0.995 ms per launch.
// Experiments that lost against this kernel (persistent double-buffered CTAs, a warp-specialised producer /
// consumer ring, cooperative LDG staging, 128-register builds, register-level column sums, programmatic
// dependent launches, a completion-counter one-kernel iteration) are in the git history and profiles/README.md.
// ----------------------------------------------------------------------------------------
// STAGES  the stages of the pass as a compile-time constant (ST_FULL: the complete synchronous iteration, what gbp_ba_iterate
//         replays; no stage tests in the instruction stream) or 0 = take them from p.stages (staged calls).
constexpr int ST_FULL = ST_ROBUSTIFY | ST_RELIN | ST_MESSAGES | ST_BELIEFS | ST_LOCAL_DAMPING;
template <int T, bool ROBUST, bool STREAM, int STAGES = 0>
__global__ void __launch_bounds__(T, 384 / T) sweep_kernel(const SweepParams p) {
    constexpr bool EARLY = STREAM;   // the early-issue prologue was re-measured on fr1desk in round 2 (same stream, same box): 8.45 vs 8.45 us
    const int stages = STAGES ? STAGES : p.stages;
    extern __shared__ __align__(128) double smem[];
    constexpr int CW = STREAM ? CAM_MF : CAM_M;
    double* s_mc = smem;                 // [T][27]  (or [T][18] factored)
    double* s_ml = s_mc + T * CW;        // [T][9]
    double* s_lp = s_ml + T * LMK_M;     // [T][9]
    double* s_cb = s_lp + T * 9;         // [33] keyframe belief (+pad to 34)
    double* s_red = s_cb + 34;           // [T/32][27]
    uint64_t* bar = reinterpret_cast<uint64_t*>(s_red + (T / 32) * CAM_M);
    double* s_full = s_red + (T / 32) * CAM_M + 2;   // [T][27] full-form messages of the tile (STREAM only)
    double* s_ch = s_full + T * CAM_M;               // [21] packed Cholesky factor of the keyframe belief's precision (STREAM only)

    const int tile = blockIdx.x;
    const int tid = threadIdx.x;
    const long long base = (long long)tile * T;
    EdgeRegs r;
    if (EARLY) {
        if (tid == 0) {
            mbar_init(bar, 1);          // only this thread touches the barrier before the __syncthreads below
            mbar_expect_tx(bar, (uint32_t)T * (CW + LMK_M + 9) * 8);
            const uint64_t pol = policy_evict_first();
            bulk_g2s_hint(s_mc, p.msg_cam + base * CW, (uint32_t)T * CW * 8, bar, pol);
            bulk_g2s_hint(s_ml, p.msg_lmk + base * LMK_M, (uint32_t)T * LMK_M * 8, bar, pol);
            bulk_g2s_hint(s_lp, p.linpoint + base * 9, (uint32_t)T * 72, bar, pol);
        }
        const int lmk = load_edge_scalars(p, base + tid, r);
        gather_lmk_belief(p, lmk, r);
        // Far-ahead L2 prefetch: CTAs start roughly in tile order, so the streams of tile + pf_dist are fetched from HBM
        // now and are L2 hits when that tile's CTA asks for them (its loads then cost an L2 round trip, not a loaded-HBM one)
        if (p.pf_dist > 0 && tid >= 1 && tid <= 8) {
            const long long tp = (long long)tile + p.pf_dist;
            if (tp < p.n_tiles) {
                const long long b = tp * T;
                switch (tid) {
                    case 1: bulk_prefetch_l2(p.msg_cam + b * CW, (uint32_t)T * CW * 8); break;
                    case 2: bulk_prefetch_l2(p.msg_lmk + b * LMK_M, (uint32_t)T * LMK_M * 8); break;
                    case 3: bulk_prefetch_l2(p.linpoint + b * 9, (uint32_t)T * 72); break;
                    case 4: bulk_prefetch_l2(p.z + b * 2, (uint32_t)T * 16); break;
                    case 5: bulk_prefetch_l2(p.lmk_idx + b, (uint32_t)T * 4); break;
                    case 6: bulk_prefetch_l2(p.iters + b, (uint32_t)T * 4); break;
                    case 7: bulk_prefetch_l2(p.flags + b, (uint32_t)T * 4); break;
                    default: if (ROBUST) bulk_prefetch_l2(p.sigma2a + b, (uint32_t)T * 8); break;
                }
            }
        }
    }
    const Tile tl = p.tiles[tile];
    const int tpos = p.tile_pos[tile];
    const int n = tl.count;
    const int n_even = (n + 1) & ~1;     // bulk copies move multiples of 16 B; the extra row is tile padding

    if (!EARLY) {
        if (tid == 0) mbar_init(bar, 1);
        __syncthreads();
        if (tid == 0) {
            mbar_expect_tx(bar, (uint32_t)n_even * (CW + LMK_M + 9) * 8);
            const uint64_t pol = policy_evict_first();
            bulk_g2s_hint(s_mc, p.msg_cam + base * CW, (uint32_t)n_even * CW * 8, bar, pol);
            bulk_g2s_hint(s_ml, p.msg_lmk + base * LMK_M, (uint32_t)n_even * LMK_M * 8, bar, pol);
            bulk_g2s_hint(s_lp, p.linpoint + base * 9, (uint32_t)n_even * 72, bar, pol);
        }
    }
    for (int i = tid; i < CAM_B; i += T) s_cb[i] = p.cam_belief[(long long)tl.cam * CAM_B + i];   // T may be 32 < 33
    if (STREAM && tid >= T - CHOL6) s_ch[tid - (T - CHOL6)] = p.cam_chol[(long long)tl.cam * CHOL6 + tid - (T - CHOL6)];   // the other end of the CTA
    if (!EARLY && tid < n) load_edge_regs(p, base + tid, r);
    __syncthreads();          // s_cb visible
    mbar_wait(bar, 0);        // bulk loads landed

    bool relin = false;
    if (tid < n)
        relin = edge_sweep<ROBUST, STREAM, STAGES>(p, base + tid, r, s_cb, s_lp + tid * 9, s_mc + tid * CW, s_ml + tid * LMK_M,
                                           STREAM ? s_full + tid * CAM_M : nullptr, STREAM ? s_ch : nullptr);
    fence_async_smem();       // generic-proxy writes -> visible to the bulk-copy engine
    const int any_relin = __syncthreads_or(relin ? 1 : 0);

    if (tid == 0) {
        // the message rows of the landmark are gathered again by belief_kernel: only the big keyframe rows
        // and the linearisation points are marked evict_first
        const uint64_t pol = policy_evict_first();
        if (stages & ST_MESSAGES) {
            bulk_s2g_hint(p.msg_cam + base * CW, s_mc, (uint32_t)n_even * CW * 8, pol);
            bulk_s2g(p.msg_lmk + base * LMK_M, s_ml, (uint32_t)n_even * LMK_M * 8);
        }
        if (any_relin) bulk_s2g_hint(p.linpoint + base * 9, s_lp, (uint32_t)n_even * 72, pol);
        bulk_commit();
    }
    if (stages & ST_BELIEFS) tile_column_sums<T>(p, tpos, n, STREAM ? s_full : s_mc, s_red);
    if (tid == 0) bulk_wait_read0();   // shared memory must outlive the engine's reads
}

template <int T, bool STREAM>
constexpr size_t sweep_smem_bytes() {
    return sizeof(double) * (size_t)(T * ((STREAM ? CAM_MF + CAM_M : CAM_M) + LMK_M + 9) + 34 + (T / 32) * CAM_M + 2 /* mbarrier */ +
                                     (STREAM ? CHOL6 + 1 : 0));
}

// ----------------------------------------------------------------------------------------
// K4: VariableNode.update_belief (gbp/gbp.py:176-198).
//   blocks [0, lmk_blocks): one thread per landmark, gathers its 72 B message rows through the
//                           CSR-by-landmark slot list, adds the prior, solves mu = Lambda^-1 eta.
//   blocks [lmk_blocks, ..): one warp per keyframe, lane k sums component k of the per-tile
//                           partial sums (fixed order), writes the local partial for multi-GPU
//                           exchange and, when `finalise`, the belief row.
// ----------------------------------------------------------------------------------------
struct BeliefParams {
    const double* msg_lmk;
    const double* lmk_prior;
    double* lmk_belief;
    const int* lmk_ptr;      // [L+1]
    const int* lmk_slots;    // [F]
    const double* tile_partial;
    const int* cam_tile_ptr; // [C+1]
    const int* cam_tiles;    // [n_tiles]
    const double* cam_prior;
    double* cam_belief;
    double* cam_chol;        // [C][CHOL6] packed Cholesky factor of the keyframe precisions (read by the streaming sweep)
    double* cam_partial;     // [K][C][27]  sums of the factor->keyframe messages per landmark chunk
    const int* cam_chunk_ptr;   // [C][K + 1] positions in the keyframe-major tile list where the chunks of a keyframe start
    int K;                   // landmark chunks
    double* cam_mu;          // [C][6]  compact copy of the means (snapshot region)
    double* lmk_mu;          // [L][3]
    int L, C, finalise;
    int parts;               // bit0: keyframe CTAs, bit1: landmark CTAs (multi-GPU runs them as two launches)
};

__device__ __forceinline__ void cam_finalise_row(const double acc /*lane k<27*/, int lane, double* row, double* mu_out, double* chol_out) {
    // lanes 0..26 hold eta[6] | Lambda[21]; gather to lane 0, factor, solve, write the 33-double row and the packed factor
    double v[CAM_M];
#pragma unroll
    for (int k = 0; k < CAM_M; ++k) v[k] = __shfl_sync(0xffffffffu, acc, k);
    if (lane < CAM_M) row[lane] = acc;
    if (lane == 0) {
        double L[36], invd[6], y[6], mu[6];
        cholesky<6>(v + 6, L, invd);
        forward<6>(L, invd, v, y);
#pragma unroll
        for (int ii = 0; ii < 6; ++ii) {       // back substitution, as in spd_solve
            const int i = 5 - ii;
            double t = y[i];
#pragma unroll
            for (int k = 0; k < 6; ++k)
                if (k > i) t -= L[k * 6 + i] * mu[k];
            mu[i] = t * invd[i];
        }
#pragma unroll
        for (int k = 0; k < 6; ++k) {
            row[27 + k] = mu[k];
            mu_out[k] = mu[k];
        }
        int q = 0;
#pragma unroll
        for (int i = 1; i < 6; ++i)
#pragma unroll
            for (int k = 0; k < 6; ++k)
                if (k < i) chol_out[q++] = L[i * 6 + k];
#pragma unroll
        for (int i = 0; i < 6; ++i) chol_out[15 + i] = invd[i];
    }
}

// 72 B message row (8 B aligned) with 16 B loads: rows of even slots start 16 B aligned, odd ones 8 B later
__device__ __forceinline__ void load_row9(const double* __restrict__ row, bool even, double v[9]) {
    if (even) {
        const double2* r2 = reinterpret_cast<const double2*>(row);
        const double2 a = __ldg(r2), b = __ldg(r2 + 1), c = __ldg(r2 + 2), d = __ldg(r2 + 3);
        v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y; v[4] = c.x; v[5] = c.y; v[6] = d.x; v[7] = d.y;
        v[8] = __ldg(row + 8);
    } else {
        v[0] = __ldg(row);
        const double2* r2 = reinterpret_cast<const double2*>(row + 1);
        const double2 a = __ldg(r2), b = __ldg(r2 + 1), c = __ldg(r2 + 2), d = __ldg(r2 + 3);
        v[1] = a.x; v[2] = a.y; v[3] = b.x; v[4] = b.y; v[5] = c.x; v[6] = c.y; v[7] = d.x; v[8] = d.y;
    }
}

// LMK_LANES lanes cooperate on one landmark: 32 / 8 for small / medium graphs (latency: a landmark of degree 46
// is gathered in 2 / 6 dependent rounds instead of 46), 1-2 for large ones (throughput: full lanes in the 3x3 solve).
template <int LMK_LANES>
__global__ void __launch_bounds__(128) belief_kernel(const BeliefParams p) {
    constexpr int LMK_PER_CTA = 128 / LMK_LANES;
    const int cam_blocks = (p.parts & 1) ? (p.C + 3) / 4 : 0;
    if ((int)blockIdx.x < cam_blocks) {
        // ---- keyframes first in the grid: their serial tile loops overlap the landmark CTAs
        const int c = (int)blockIdx.x * 4 + (threadIdx.x >> 5);
        const int lane = threadIdx.x & 31;
        if (c >= p.C) return;
        // per landmark chunk: tile sums in tile order; then the chunk sums in chunk order (the same association on every number of
        // GPUs: a rank holds whole chunks), then the prior
        double acc = 0.0;
        if (lane < CAM_M) {
            const double prior = p.cam_prior[(long long)c * CAM_M + lane];   // issued with the first loads, not behind the sums
            for (int k = 0; k < p.K; ++k) {
                const int t0 = p.cam_chunk_ptr[c * (p.K + 1) + k], t1 = p.cam_chunk_ptr[c * (p.K + 1) + k + 1];
                double part = 0.0;
                int q = t0;
                for (; q + 8 <= t1; q += 8) {          // 8 independent loads in flight, added in tile order
                    double v[8];
#pragma unroll
                    for (int u = 0; u < 8; ++u) v[u] = p.tile_partial[(long long)(q + u) * CAM_M + lane];
#pragma unroll
                    for (int u = 0; u < 8; ++u) part += v[u];
                }
                for (; q < t1; ++q) part += p.tile_partial[(long long)q * CAM_M + lane];
                p.cam_partial[((long long)k * p.C + c) * CAM_M + lane] = part;
                acc = k == 0 ? part : acc + part;
            }
            acc += prior;
        }
        if (p.finalise) cam_finalise_row(acc, lane, p.cam_belief + (long long)c * CAM_B, p.cam_mu + (long long)c * 6, p.cam_chol + (long long)c * CHOL6);
        return;
    }
    // ---- landmarks: LMK_LANES lanes gather one landmark's message rows in parallel, then a
    //      fixed-shape shuffle tree (deterministic) combines them
    const int sub = threadIdx.x & (LMK_LANES - 1);
    const int l = ((int)blockIdx.x - cam_blocks) * LMK_PER_CTA + (threadIdx.x / LMK_LANES);
    const bool valid = l < p.L;
    double acc[LMK_M];
    double prior_k = 0.0;
#pragma unroll
    for (int k = 0; k < LMK_M; ++k) acc[k] = 0.0;
    if (LMK_LANES == 32) {
        // small graphs (latency-bound): the prior row is fetched with the first loads, not behind the gather and the tree.  Lane k
        // of the landmark's warp loads component k (one register per lane); lane 0 collects them by shuffle after the tree.
        if (valid && sub < LMK_M) prior_k = p.lmk_prior[(long long)l * LMK_M + sub];
    }
    if (valid) {
        const int p0 = p.lmk_ptr[l], p1 = p.lmk_ptr[l + 1];
        int q = p0 + sub;
        // four rows in flight per thread (slot loads, then row loads, then the adds in edge order)
        for (; q + 3 * LMK_LANES < p1; q += 4 * LMK_LANES) {
            int slot[4];
            double v[4][9];
#pragma unroll
            for (int u = 0; u < 4; ++u) slot[u] = p.lmk_slots[q + u * LMK_LANES];
#pragma unroll
            for (int u = 0; u < 4; ++u) load_row9(p.msg_lmk + (long long)slot[u] * LMK_M, (slot[u] & 1) == 0, v[u]);
#pragma unroll
            for (int u = 0; u < 4; ++u)
#pragma unroll
                for (int k = 0; k < LMK_M; ++k) acc[k] += v[u][k];
        }
        for (; q < p1; q += LMK_LANES) {
            const int slot = p.lmk_slots[q];
            double v[9];
            load_row9(p.msg_lmk + (long long)slot * LMK_M, (slot & 1) == 0, v);
#pragma unroll
            for (int k = 0; k < LMK_M; ++k) acc[k] += v[k];
        }
    }
#pragma unroll
    for (int o = LMK_LANES / 2; o > 0; o >>= 1)
#pragma unroll
        for (int k = 0; k < LMK_M; ++k) acc[k] += __shfl_xor_sync(0xffffffffu, acc[k], o);
    if (LMK_LANES == 32) {
#pragma unroll
        for (int k = 0; k < LMK_M; ++k) acc[k] += __shfl_sync(0xffffffffu, prior_k, k);   // only lane 0's sum is used
    }
    if (valid && sub == 0) {
        if (LMK_LANES != 32) {
#pragma unroll
            for (int k = 0; k < LMK_M; ++k) acc[k] += p.lmk_prior[(long long)l * LMK_M + k];
        }
        double mu[3];
        spd_solve<3>(acc + 3, acc, mu);
        double2* dst = reinterpret_cast<double2*>(p.lmk_belief + (long long)l * LMK_B);
        dst[0] = make_double2(acc[0], acc[1]);
        dst[1] = make_double2(acc[2], acc[3]);
        dst[2] = make_double2(acc[4], acc[5]);
        dst[3] = make_double2(acc[6], acc[7]);
        dst[4] = make_double2(acc[8], mu[0]);
        dst[5] = make_double2(mu[1], mu[2]);
        double* m = p.lmk_mu + (long long)l * 3;
        m[0] = mu[0]; m[1] = mu[1]; m[2] = mu[2];
    }
}

// keyframe beliefs from the gathered chunk sums (multi-GPU): chunk sums in chunk order, then the prior -- the association of belief_kernel
__global__ void __launch_bounds__(128) cam_update_kernel(const double* __restrict__ partials, int nranks, int C,
                                                         const double* __restrict__ cam_prior,
                                                         double* __restrict__ cam_belief, double* __restrict__ cam_mu,
                                                         double* __restrict__ cam_chol) {
    const int c = blockIdx.x * 4 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (c >= C) return;
    double acc = 0.0;
    if (lane < CAM_M) {
        for (int r = 0; r < nranks; ++r) {
            const double part = partials[((long long)r * C + c) * CAM_M + lane];
            acc = r == 0 ? part : acc + part;
        }
        acc += cam_prior[(long long)c * CAM_M + lane];
    }
    cam_finalise_row(acc, lane, cam_belief + (long long)c * CAM_B, cam_mu + (long long)c * 6, cam_chol + (long long)c * CHOL6);
}


// ----------------------------------------------------------------------------------------
// K5: BAFactorGraph.are / FactorGraph.energy / relinearisation count
//     (gbp/gbp_ba.py:61-69, gbp/gbp.py:36-44,251-259, ba.py:97-100)
// ----------------------------------------------------------------------------------------
struct MetricParams {
    const Tile* tiles;
    const int* lmk_idx;
    const double* z;
    const int* iters;
    const double* sigma2a;
    const double* cam_belief;
    const double* lmk_belief;
    double* tile_metric;  // [n_tiles][3]
    Intrinsics K;
    double var0;
    int robust;
};

template <int T>
__global__ void __launch_bounds__(T) metric_kernel(const MetricParams p) {
    __shared__ double s_R[9], s_t[3];
    __shared__ double s_w[3][T / 32];
    const int tile = blockIdx.x, tid = threadIdx.x;
    const Tile tl = p.tiles[tile];
    const long long base = (long long)tile * T;
    if (tid == 0) {
        const double* row = p.cam_belief + (long long)tl.cam * CAM_B + 27;
        double w[3] = {row[3], row[4], row[5]};
        double R[9];
        so3exp(w, R);
#pragma unroll
        for (int k = 0; k < 9; ++k) s_R[k] = R[k];
        s_t[0] = row[0]; s_t[1] = row[1]; s_t[2] = row[2];
    }
    __syncthreads();
    double a = 0.0, en = 0.0, cnt = 0.0;
    if (tid < tl.count) {
        const long long e = base + tid;
        const double* lrow = p.lmk_belief + (long long)p.lmk_idx[e] * LMK_B + 9;
        const double y[3] = {lrow[0], lrow[1], lrow[2]};
        double R[9], t[3], h[2], pp[3];
#pragma unroll
        for (int k = 0; k < 9; ++k) R[k] = s_R[k];
        t[0] = s_t[0]; t[1] = s_t[1]; t[2] = s_t[2];
        project(p.K, R, t, y, h, pp);
        const double r0 = h[0] - p.z[2 * e], r1 = h[1] - p.z[2 * e + 1];
        a = sqrt(r0 * r0 + r1 * r1);
        const double var = p.robust ? p.sigma2a[e] : p.var0;
        en = 0.5 * (a * a) / var;
        cnt = (p.iters[e] == 0) ? 1.0 : 0.0;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        a += __shfl_down_sync(0xffffffffu, a, o);
        en += __shfl_down_sync(0xffffffffu, en, o);
        cnt += __shfl_down_sync(0xffffffffu, cnt, o);
    }
    if ((tid & 31) == 0) {
        s_w[0][tid >> 5] = a; s_w[1][tid >> 5] = en; s_w[2][tid >> 5] = cnt;
    }
    __syncthreads();
    if (tid < 3) {
        double v = 0.0;
#pragma unroll
        for (int g = 0; g < T / 32; ++g) v += s_w[tid][g];
        p.tile_metric[(long long)tile * 3 + tid] = v;
    }
}

// deterministic final reduction of [n][W] rows into out[W]   (one block of 256 threads)
template <int W>
__global__ void __launch_bounds__(256) reduce_rows_kernel(const double* __restrict__ rows, int n, double* __restrict__ out) {
    __shared__ double s[W][256];
    double acc[W];
#pragma unroll
    for (int k = 0; k < W; ++k) acc[k] = 0.0;
    for (int i = threadIdx.x; i < n; i += 256)
#pragma unroll
        for (int k = 0; k < W; ++k) acc[k] += rows[(long long)i * W + k];
#pragma unroll
    for (int k = 0; k < W; ++k) s[k][threadIdx.x] = acc[k];
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o)
#pragma unroll
            for (int k = 0; k < W; ++k) s[k][threadIdx.x] += s[k][threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x < W) out[threadIdx.x] = s[threadIdx.x][0];
}

// ----------------------------------------------------------------------------------------
// K6: BAFactorGraph.generate_priors_var (gbp/gbp_ba.py:20-34)
// ----------------------------------------------------------------------------------------
// per edge: largest entry of Lambda_f = J^T J / var (for a PSD matrix: its largest diagonal entry)
template <int T>
__global__ void __launch_bounds__(T) edge_lammax_kernel(const Tile* tiles, const double* linpoint, const double* sigma2a,
                                                        int robust, double var0, Intrinsics K, double* edge_max,
                                                        double* tile_max) {
    __shared__ double s_w[T / 32];
    const int tile = blockIdx.x, tid = threadIdx.x;
    const Tile tl = tiles[tile];
    const long long e = (long long)tile * T + tid;
    double m = 0.0;
    if (tid < tl.count) {
        double x0[9], J[18], h0[2];
#pragma unroll
        for (int k = 0; k < 9; ++k) x0[k] = linpoint[e * 9 + k];
        linearise(K, x0, J, h0);
        const double var = robust ? sigma2a[e] : var0;
#pragma unroll
        for (int k = 0; k < 9; ++k) m = fmax(m, (J[k] * J[k] + J[9 + k] * J[9 + k]) / var);
        edge_max[e] = m;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_down_sync(0xffffffffu, m, o));
    if ((tid & 31) == 0) s_w[tid >> 5] = m;
    __syncthreads();
    if (tid == 0) {
#pragma unroll
        for (int g = 1; g < T / 32; ++g) m = fmax(m, s_w[g]);
        tile_max[tile] = m;
    }
}

__global__ void cam_max_kernel(const double* tile_max, const int* cam_tile_ptr, const int* cam_tiles, int C, double* cam_max) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    double m = 0.0;
    for (int q = cam_tile_ptr[c]; q < cam_tile_ptr[c + 1]; ++q) m = fmax(m, tile_max[cam_tiles[q]]);
    cam_max[c] = m;
}

// prior rows: Lambda = I * max / weaker^2, eta = Lambda mu   (mu = current mean stored in the belief row)
__global__ void cam_prior_kernel(const double* cam_max, double weaker, int C, const double* cam_belief, double* cam_prior) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    const double lam = cam_max[c] / (weaker * weaker);
    double* row = cam_prior + (long long)c * CAM_M;
    for (int k = 0; k < CAM_M; ++k) row[k] = 0.0;
    for (int i = 0; i < 6; ++i) {
        row[6 + sidx<6>(i, i)] = lam;
        row[i] = lam * cam_belief[(long long)c * CAM_B + 27 + i];
    }
}

__global__ void lmk_prior_kernel(const double* edge_max, const int* lmk_ptr, const int* lmk_slots, double weaker, int L,
                                 const double* lmk_belief, double* lmk_prior) {
    const int l = blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= L) return;
    double m = 0.0;
    for (int q = lmk_ptr[l]; q < lmk_ptr[l + 1]; ++q) m = fmax(m, edge_max[lmk_slots[q]]);
    const double lam = m / (weaker * weaker);
    double* row = lmk_prior + (long long)l * LMK_M;
    for (int k = 0; k < LMK_M; ++k) row[k] = 0.0;
    for (int i = 0; i < 3; ++i) {
        row[3 + sidx<3>(i, i)] = lam;
        row[i] = lam * lmk_belief[(long long)l * LMK_B + 9 + i];
    }
}

// set_priors_var (gbp/gbp_ba.py:44-52): eta = Lambda mu for given packed Lambda
template <int N>
__global__ void prior_eta_kernel(int V, const double* belief, int brow, int mu_off, double* prior) {
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= V) return;
    constexpr int W = N + N * (N + 1) / 2;
    double* row = prior + (long long)v * W;
    const double* mu = belief + (long long)v * brow + mu_off;
    for (int i = 0; i < N; ++i) {
        double s = 0.0;
        for (int j = 0; j < N; ++j) s += row[N + sym<N>(i, j)] * mu[j];
        row[i] = s;
    }
}

__global__ void scale_kernel(double* x, long long n, double f) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) x[i] *= f;
}

// ----------------------------------------------------------------------------------------
// initialisation and field access
// ----------------------------------------------------------------------------------------
// linpoint = [mu0_cam, mu0_lmk] (gbp/gbp_ba.py:136-137); iters_since_relin = 1, flags = 0 (gbp/gbp.py:248-249)
template <int T>
__global__ void __launch_bounds__(T) init_edges_kernel(const Tile* tiles, const int* lmk_idx, const double* cam_belief,
                                                       const double* lmk_belief, double var0, double* linpoint,
                                                       int* iters, int* flags, double* sigma2a) {
    const int tile = blockIdx.x, tid = threadIdx.x;
    const Tile tl = tiles[tile];
    const long long e = (long long)tile * T + tid;
    if (tid < tl.count) {
        const double* cm = cam_belief + (long long)tl.cam * CAM_B + 27;
        const double* lm = lmk_belief + (long long)lmk_idx[e] * LMK_B + 9;
        for (int k = 0; k < 6; ++k) linpoint[e * 9 + k] = cm[k];
        for (int k = 0; k < 3; ++k) linpoint[e * 9 + 6 + k] = lm[k];
        iters[e] = 1;
    } else {
        for (int k = 0; k < 9; ++k) linpoint[e * 9 + k] = 0.0;
        iters[e] = -1;
    }
    flags[e] = 0;
    sigma2a[e] = var0;
}

// belief rows from initial means: eta = 0, Lambda = 0, mu = mu0 (gbp/gbp.py:165-168, gbp/gbp_ba.py:116,123)
__global__ void init_belief_kernel(const double* mu0, int V, int N, int brow, double* belief, double* mu_compact) {
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= V) return;
    double* row = belief + (long long)v * brow;
    for (int k = 0; k < brow - N; ++k) row[k] = 0.0;
    for (int k = 0; k < N; ++k) {
        row[brow - N + k] = mu0[(long long)v * N + k];
        mu_compact[(long long)v * N + k] = mu0[(long long)v * N + k];
    }
}

// after the client wrote a belief table: the mean of a variable with a positive-definite precision is Lambda^-1 eta, as every
// consumer of the reference computes it (gbp/gbp.py:71, 192-193) -- a written `mu` only stands where Lambda is still zero (the
// initial state); then the compact means are refreshed
template <int N>
__global__ void refresh_mu_kernel(double* belief, int V, int brow, double* mu_compact, double* chol /* N == 6: [V][CHOL6], else nullptr */) {
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= V) return;
    double* row = belief + (long long)v * brow;
    bool pd = true;
    for (int i = 0; i < N; ++i) pd = pd && row[N + sidx<N>(i, i)] > 0.0;
    if (pd) {
        double lam[N * (N + 1) / 2], eta[N], mu[N];
        for (int k = 0; k < N * (N + 1) / 2; ++k) lam[k] = row[N + k];
        for (int k = 0; k < N; ++k) eta[k] = row[k];
        spd_solve<N>(lam, eta, mu);
        bool ok = true;
        for (int k = 0; k < N; ++k) ok = ok && (mu[k] == mu[k]);      // not positive definite after all: keep the written mean
        if (ok)
            for (int k = 0; k < N; ++k) row[brow - N + k] = mu[k];
        if (N == 6 && chol) cholesky6_packed(lam, chol + (long long)v * CHOL6);
    }
    for (int k = 0; k < N; ++k) mu_compact[(long long)v * N + k] = row[brow - N + k];
}

// dst[f][w] = src[slot_of_factor[f]][w]  (rows of W 4-byte words)
__global__ void gather_rows_kernel(uint32_t* __restrict__ dst, const uint32_t* __restrict__ src,
                                   const int* __restrict__ slot_of_factor, long long F, int W) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= F * W) return;
    const long long f = i / W;
    const int w = (int)(i - f * W);
    dst[i] = src[(long long)slot_of_factor[f] * W + w];
}
__global__ void scatter_rows_kernel(uint32_t* __restrict__ dst, const uint32_t* __restrict__ src,
                                    const int* __restrict__ slot_of_factor, long long F, int W) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= F * W) return;
    const long long f = i / W;
    const int w = (int)(i - f * W);
    dst[(long long)slot_of_factor[f] * W + w] = src[i];
}
// GBP_F_MSG_CAM with the factored layout: the client always sees eta[6] | Lambda[21] in factor order
__global__ void export_msg_cam_factored_kernel(double* __restrict__ dst, const double* __restrict__ src,
                                               const int* __restrict__ slot_of_factor, long long F) {
    const long long f = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= F) return;
    const double* row = src + (long long)slot_of_factor[f] * CAM_MF;
    double W[12], lam[21];
    for (int k = 0; k < 12; ++k) W[k] = row[6 + k];
    expand_factored6(W, lam);
    for (int k = 0; k < 6; ++k) dst[f * CAM_M + k] = row[k];
    for (int k = 0; k < 21; ++k) dst[f * CAM_M + 6 + k] = lam[k];
}
__global__ void import_msg_cam_factored_kernel(double* __restrict__ dst, const double* __restrict__ src,
                                               const int* __restrict__ slot_of_factor, long long F) {
    const long long f = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= F) return;
    double* row = dst + (long long)slot_of_factor[f] * CAM_MF;
    double lam[21], W[12];
    for (int k = 0; k < 21; ++k) lam[k] = src[f * CAM_M + 6 + k];
    factor_rank2_6(lam, W);
    for (int k = 0; k < 6; ++k) row[k] = src[f * CAM_M + k];
    for (int k = 0; k < 12; ++k) row[6 + k] = W[k];
}
__global__ void fill_iters_kernel(int* iters, long long n_slots, int value) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_slots && iters[i] >= 0) iters[i] = value;
}

// J[2x9] | b[2] of every factor at its linearisation point, factor order (GBP_F_JACOBIAN_B)
__global__ void export_jb_kernel(const int* slot_of_factor, long long F, const double* linpoint, const double* z,
                                 Intrinsics K, double* out) {
    const long long f = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= F) return;
    const long long e = slot_of_factor[f];
    double x0[9], J[18], h0[2], b[2];
    for (int k = 0; k < 9; ++k) x0[k] = linpoint[e * 9 + k];
    linearise(K, x0, J, h0);
    const double zz[2] = {z[2 * e], z[2 * e + 1]};
    factor_rhs(J, x0, zz, h0, b);
    for (int k = 0; k < 18; ++k) out[f * 20 + k] = J[k];
    out[f * 20 + 18] = b[0];
    out[f * 20 + 19] = b[1];
}

// standalone reprojection model (parity tests of meas_fn / jac_fn)
__global__ void reprojection_eval_kernel(const double* x, long long n, Intrinsics K, double* out_h, double* out_J) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double x0[9], J[18], h0[2];
    for (int k = 0; k < 9; ++k) x0[k] = x[i * 9 + k];
    linearise(K, x0, J, h0);
    out_h[2 * i] = h0[0];
    out_h[2 * i + 1] = h0[1];
    for (int k = 0; k < 18; ++k) out_J[i * 18 + k] = J[k];
}

}  // namespace gbp
