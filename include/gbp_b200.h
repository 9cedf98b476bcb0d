/*
 * gbp_b200.h  --  C ABI of libgbp_b200.so: the B200-native Gaussian Belief Propagation
 * sweep for bundle adjustment (reprojection factors, 6-dof keyframes, 3-dof landmarks).
 *
 * The reference (joeaortiz/gbp) is pure Python and has NO FFI/plugin interface; its
 * boundary for this path is the Python class API of gbp/gbp.py and gbp/gbp_ba.py.  Each
 * entry point below therefore cites the reference METHOD it replaces (file:line under
 * the reference tree).  The host-side Python mirror of those classes lives in
 * gbp_b200/compat/ and calls only these functions (ctypes); see INTEGRATION.md.
 *
 * Conventions
 *   - every function returns 0 on success, a gbp_status otherwise; gbp_last_error()
 *     returns a thread-local description.  CUDA errors are captured, never abort().
 *   - all host pointers are caller-owned, copied before return, never retained.
 *   - all floating point is IEEE float64, indices are int32.
 *   - one handle = one CUDA device + one stream (its own or the library's); calls on a handle are not re-entrant.
 *   - "factor order" is the reference's: camera-major, file order within a camera
 *     (gbp/gbp_ba.py:128-143), i.e. a stable sort of the measurement list by camera id.
 *   - symmetric matrices cross the ABI PACKED, upper triangle row-major:
 *     6x6 -> 21 doubles, 3x3 -> 6 doubles.
 */
#ifndef GBP_B200_H
#define GBP_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GBP_B200_ABI_VERSION 3

typedef struct gbp_ba_graph* gbp_handle;

typedef enum gbp_status {
    GBP_OK = 0,
    GBP_ERR_INVALID = 1,   /* bad argument / inconsistent graph                      */
    GBP_ERR_CUDA = 2,      /* CUDA runtime error (message in gbp_last_error)         */
    GBP_ERR_NO_DEVICE = 3, /* no CUDA device: there is NO CPU fallback               */
    GBP_ERR_STATE = 4,     /* call order violated (e.g. sweep before priors)         */
    GBP_ERR_COMM = 5       /* NCCL missing or a collective failed                    */
} gbp_status;

typedef enum gbp_loss { GBP_LOSS_NONE = 0, GBP_LOSS_HUBER = 1, GBP_LOSS_CONSTANT = 2 } gbp_loss;

/* The `configs` dict of ba.py:51-60 consumed by create_ba_graph (gbp/gbp_ba.py:104-107,135)
 * and the FactorGraph constructor (gbp/gbp.py:12-34). */
typedef struct gbp_config {
    double gauss_noise_std;      /* sigma of the measurement model (gbp/gbp.py:236)          */
    double eta_damping;          /* graph-level damping (gbp/gbp.py:28)                        */
    double beta;                 /* relinearisation threshold (gbp/gbp.py:32)                  */
    double Nstds;                /* mahalanobis_threshold of the robust loss (gbp/gbp.py:244)  */
    int32_t num_undamped_iters;  /* gbp/gbp.py:33                                              */
    int32_t min_linear_iters;    /* gbp/gbp.py:34                                              */
    int32_t loss;                /* gbp_loss (gbp/gbp.py:243)                                  */
    int32_t tile_edges;          /* 0 = auto; else 32/64/128 edges per tile (engine tuning)    */
    int32_t lmk_block;           /* 0 = auto; landmarks per L2 block of the edge schedule      */
    int32_t kernel_variant;      /* build of the sweep kernel: 0 = automatic (graphs of up to 8192 tiles live in L2 and get build 1,
                                    larger ones stream from HBM and get build 2);
                                    1 = full 27-double factor->keyframe message rows, bulk copies sized by the tile descriptor;
                                    2 = streaming build: factor->keyframe messages stored with their rank-2 precision factored
                                        (eta[6] | W[2][6], Lambda = W^T W: 144 B less traffic per edge and sweep), nothing in the
                                        prologue waits for the tile descriptor, far-ahead L2 prefetch on graphs of more than 8192
                                        tiles; tiles of 32 / 64.  GBP_F_MSG_CAM reads and writes keep the full eta[6] | Lambda[21]
                                        form.  Both builds run the same per-edge arithmetic (gbp_edge.cuh). */
    /* Landmark CHUNKS (no reference counterpart).  The sum of the factor->keyframe messages of a keyframe is formed per chunk of
     * consecutive landmarks (tiles never straddle a chunk) and the chunk sums are added in chunk order.  A multi-GPU run gives every
     * rank whole chunks of the SAME global chunking, so the keyframe beliefs -- and with them the whole trajectory -- are
     * bit-identical for 1, 2, 4 and 8 GPUs.  Chunk k of lmk_chunks_total covers the global landmarks
     * [lmk_total * k / lmk_chunks_total, lmk_total * (k + 1) / lmk_chunks_total); this graph holds lmk_chunks of them starting at
     * chunk lmk_chunk_first, its landmark 0 being global landmark lmk_first.  lmk_chunks = 0: automatic (the whole graph in the
     * largest power of two <= 8 of chunks that leaves a chunk at least 125000 landmarks: 8 from 1 M landmarks on). */
    int32_t lmk_chunks, lmk_chunk_first, lmk_chunks_total, reserved0;
    int64_t lmk_first, lmk_total;
} gbp_config;

/* Stages of FactorGraph.synchronous_iteration (gbp/gbp.py:86-92), OR-able. */
enum {
    GBP_STAGE_ROBUSTIFY = 1, /* robustify_all_factors   gbp/gbp.py:82-84, 296-332 */
    GBP_STAGE_RELIN = 2,     /* relinearise_factors     gbp/gbp.py:64-80          */
    GBP_STAGE_MESSAGES = 4,  /* compute_all_messages    gbp/gbp.py:46-54, 334-373 */
    GBP_STAGE_BELIEFS = 8,   /* update_all_beliefs      gbp/gbp.py:56-58, 176-198 */
    GBP_STAGE_LOCAL_DAMPING = 16, /* local_relin=True: per-factor damping (gbp/gbp.py:49-52) */
    GBP_STAGE_DEFER_LANDMARKS = 32 /* with BELIEFS in gbp_ba_sweep_local: only reduce the keyframe partial sums; the
                                      landmark beliefs are updated by gbp_ba_landmark_update (overlaps the exchange) */
};

/* Fields readable / writable through gbp_ba_read / gbp_ba_write.  Row widths in doubles
 * unless noted; factor-indexed fields are in factor order. */
typedef enum gbp_field {
    GBP_F_CAM_BELIEF = 0, /* C x 33 : eta[6] | Lambda packed[21] | mu[6]   (VariableNode.belief/.mu, gbp/gbp.py:165-168) */
    GBP_F_LMK_BELIEF = 1, /* L x 12 : eta[3] | Lambda packed[6]  | mu[3]                                                  */
    GBP_F_CAM_PRIOR = 2,  /* C x 27 : eta[6] | Lambda packed[21]           (VariableNode.prior, gbp/gbp.py:170)           */
    GBP_F_LMK_PRIOR = 3,  /* L x 9                                                                                        */
    GBP_F_MSG_CAM = 4,    /* F x 27 : factor->keyframe message              (Factor.messages[0], gbp/gbp.py:228)           */
    GBP_F_MSG_LMK = 5,    /* F x 9  : factor->landmark message              (Factor.messages[1])                           */
    GBP_F_LINPOINT = 6,   /* F x 9  : Factor.linpoint (gbp/gbp.py:231)                                                     */
    GBP_F_ITERS_SINCE_RELIN = 7, /* F x int32 : Factor.iters_since_relin (gbp/gbp.py:249; written by ba.py:91-93)         */
    GBP_F_FLAGS = 8,      /* F x int32 : bit0 = per-factor damping on (Factor.eta_damping != 0), bit1 = robust_flag        */
    GBP_F_ADAPTIVE_VAR = 9, /* F x 1 : Factor.adaptive_gauss_noise_var (gbp/gbp.py:242)                                    */
    GBP_F_MEASUREMENT = 10, /* F x 2 : Factor.measurement (read only)                                                      */
    GBP_F_JACOBIAN_B = 11,  /* F x 20: J[2x9] row-major | b[2] = J x0 + z - h(x0), recomputed at linpoint (read only);
                               Factor.factor = (J^T b / var, J^T J / var)  (gbp/gbp.py:287-289)                            */
    GBP_F_ADJ = 12,         /* F x 2 int32 : (camera id, landmark id) = Factor.adj_vIDs (landmark NOT offset by C; read only) */
    GBP_F_FILE_INDEX = 13,  /* F x int32 : position of each factor in the measurement list passed to create (read only)    */
    GBP_F_CAM_PARTIAL = 14, /* chunks x C x 27 : this graph's sums of factor->keyframe messages per landmark chunk (the multi-GPU
                               exchange buffer; read only) */
    GBP_F_CAM_MU = 15,      /* C x 6 : compact copy of the keyframe means (VariableNode.mu, gbp/gbp.py:193; read only)     */
    GBP_F_LMK_MU = 16,      /* L x 3 : compact copy of the landmark means (read only)                                      */
    GBP_F__COUNT
} gbp_field;

const char* gbp_last_error(void);
int gbp_abi_version(void);
/* Number of CUDA devices visible (0 on a CPU-only host; never an error). */
int gbp_device_count(void);

/* create_ba_graph (gbp/gbp_ba.py:97-150) after read_balfile: builds the device-resident graph from
 * the measurement list (file order), linearises every factor at the initial means
 * (gbp/gbp_ba.py:136-137 -> gbp/gbp.py:267-294), zero messages, iters_since_relin = 1, damping 0.
 * `stream` is a cudaStream_t (NULL = the library's stream of that device, shared by all handles created with NULL: CUDA graphs
 * replayed on more than one stream of a process were measured ~5 % slower, on every stream.  Handles that share the library's
 * stream must not be driven from different threads at the same time -- give each its own stream for that).  n_cam_total/n_lmk are the sizes
 * of cam_mu0 / lmk_mu0; every camera id < C, landmark id < L. */
int gbp_ba_create(const gbp_config* cfg, int32_t C, int32_t L, int64_t F,
                  const int32_t* cam_id, const int32_t* lmk_id, const double* z /* F x 2 */,
                  const double* cam_mu0 /* C x 6 */, const double* lmk_mu0 /* L x 3 */,
                  const double K[4] /* fx fy cx cy */, int device, void* stream, gbp_handle* out);
int gbp_ba_destroy(gbp_handle h);
/* Shell cache (no reference counterpart).  gbp_ba_destroy does not return the graph's device arena, its page-locked
 * staging block and its instantiated CUDA graphs to the driver: the next gbp_ba_create on the same device reuses the
 * arena when it is large enough, and the CUDA graphs too when the new graph has the same shape and parameters (ba.py
 * run file after file: no cudaMalloc / cudaFree / cudaGraphInstantiate per problem).  gbp_cache_configure bounds it
 * (defaults: 4 shells, arenas of at most 1 GiB; 0 shells = off, frees everything cached now); gbp_cache_stats:
 * out = {creates, arena reuses, graph reuses, shells evicted, shells cached now, bytes cached now}. */
int gbp_cache_configure(int32_t max_shells, int64_t max_arena_bytes);
int gbp_cache_stats(int64_t out[6]);
/* Back to the state right after gbp_ba_create (zero messages and priors, initial means and
 * linearisation points, iters_since_relin = 1): re-run a solve without rebuilding the graph. */
int gbp_ba_reset(gbp_handle h);

/* Sizes: C, L, F, number of edge tiles, edges per tile, padded edge slots. */
int gbp_ba_sizes(gbp_handle h, int64_t out[6]);

/* The host graph compiler on its own (pure host code, works without a GPU): the storage order gbp_ba_create builds
 * from a measurement list -- factor order = stable sort by camera (gbp/gbp_ba.py:128-143), tiles of <= T edges of ONE
 * keyframe sorted by (landmark block, keyframe), CSR by landmark over slots, CSR by keyframe over tiles.
 * gbp_plan_sizes: C, L, F, tiles, edges per tile, slots.  gbp_plan_copy: any output may be NULL; tiles = (keyframe, count)
 * pairs; adj = (keyframe, landmark) per factor; lmk_idx per slot (0 in padding slots). */
typedef struct gbp_plan_s* gbp_plan;
/* chunking: NULL = automatic, else {lmk_chunks, lmk_chunk_first, lmk_chunks_total, lmk_first, lmk_total} as in gbp_config */
int gbp_plan_create(int32_t tile_edges, int32_t lmk_block, const int64_t* chunking, int32_t C, int32_t L, int64_t F, const int32_t* cam_id,
                    const int32_t* lmk_id, gbp_plan* out);
/* chunk of every tile [tiles]; per keyframe the positions in cam_tiles where its chunks start, C x (chunks + 1); returns the chunk count */
int gbp_plan_chunks(gbp_plan p, int32_t* tile_chunk, int32_t* cam_chunk_ptr);
int gbp_plan_sizes(gbp_plan p, int64_t out[6]);
int gbp_plan_copy(gbp_plan p, int32_t* tiles, int32_t* slot_of_factor, int32_t* file_of_factor, int32_t* adj, int32_t* lmk_idx,
                  int32_t* lmk_ptr, int32_t* lmk_slots, int32_t* cam_tile_ptr, int32_t* cam_tiles);
void gbp_plan_destroy(gbp_plan p);

/* Engine layout chosen for this graph (no reference counterpart; used by bench.py to count the bytes a sweep moves):
 * out[0] doubles per stored factor->keyframe message (27 full, 18 factored), out[1] L2 prefetch distance in tiles,
 * out[2] sweep kernel build in use (gbp_config.kernel_variant after the automatic choice: 1 or 2), out[3] landmark chunks of the
 * keyframe-side sums held by this graph (gbp_config.lmk_chunks after the automatic choice). */
int gbp_ba_layout(gbp_handle h, int64_t out[4]);

/* BAFactorGraph.generate_priors_var (gbp/gbp_ba.py:20-34).  The per-keyframe maxima must cover the factors of ALL ranks: with a
 * communicator attached gbp_ba_prior_scan reduces them itself; a client with its own transport calls gbp_ba_prior_scan,
 * max-reduces the C doubles it returns over the ranks and passes them to gbp_ba_generate_priors as cam_max (NULL = scan here). */
int gbp_ba_prior_scan(gbp_handle h, double* cam_max /* C, host */);
int gbp_ba_generate_priors(gbp_handle h, double weaker_factor, const double* cam_max /* C or NULL */);
/* BAFactorGraph.set_priors_var (gbp/gbp_ba.py:44-52): prior Lambda = given packed precision,
 * eta = Lambda mu (current means). */
int gbp_ba_set_priors(gbp_handle h, const double* cam_lam /* C x 21 */, const double* lmk_lam /* L x 6 */);
/* BAFactorGraph.weaken_priors (gbp/gbp_ba.py:36-42): prior eta, Lambda *= factor. */
int gbp_ba_scale_priors(gbp_handle h, double factor);

/* One pass of the selected stages of synchronous_iteration (gbp/gbp.py:86-92) over the LOCAL edges.
 * Without GBP_STAGE_BELIEFS nothing is reduced.  With it, landmark beliefs are updated and the local
 * sum of factor->keyframe messages is left in GBP_F_CAM_PARTIAL; keyframe beliefs are finalised by
 * gbp_ba_cam_update.  Everything is enqueued on the handle's stream; no host synchronisation. */
int gbp_ba_sweep_local(gbp_handle h, int stages);
/* The landmark half of update_all_beliefs (gbp/gbp.py:176-198) after a sweep with GBP_STAGE_DEFER_LANDMARKS: needs
 * no communication, so a multi-GPU caller runs it while the keyframe partial sums are being exchanged. */
int gbp_ba_landmark_update(gbp_handle h);
/* Finish VariableNode.update_belief (gbp/gbp.py:176-198) for the keyframes: belief = prior + the chunk sums in chunk order.
 * `partials` is a DEVICE pointer to n_partials x C x 27 doubles (the output of an all-gather of every rank's
 * GBP_F_CAM_PARTIAL: ranks in rank order hold the chunks in chunk order); NULL = this handle's own chunk sums (single GPU). */
int gbp_ba_cam_update(gbp_handle h, const double* partials_dev, int n_partials);

/* ---- multi-GPU: one process per GPU, NCCL (no reference counterpart: the reference is one Python process) ----------------
 * The graph is partitioned by LANDMARK: every rank creates its handle from all keyframes, a contiguous landmark range
 * and the measurements of those landmarks, with gbp_config.lmk_chunks / lmk_chunk_first / lmk_chunks_total / lmk_first /
 * lmk_total naming its share of ONE global landmark chunking (lmk_chunks_total / nranks chunks per rank).  Landmarks are
 * interior to a rank; the keyframes are the boundary and are replicated.  gbp_comm_unique_id (one rank) -> the id reaches the
 * other ranks by the client's own means -> gbp_comm_create on every rank (collective; one communicator per process, it outlives
 * the graphs: creating one costs seconds) -> gbp_ba_attach_comm for every graph (collective).  From then on
 *   gbp_ba_iterate / gbp_ba_update_beliefs / gbp_ba_iterate_snapshot  exchange the keyframe chunk sums with ONE ncclAllGather
 *       per iteration on a high-priority side stream while the landmark beliefs are updated on the handle's stream, and every
 *       rank adds the chunk sums in chunk order: keyframe beliefs are bit-identical on every rank AND to the single-GPU run
 *       with the same chunking; an iteration stays one CUDA-graph replay per rank;
 *   gbp_ba_prior_scan / gbp_ba_generate_priors(cam_max = NULL)  take the per-keyframe maximum over all ranks (ncclAllReduce max);
 *   gbp_ba_metrics / snapshots  return sums over the WHOLE graph (ncclAllReduce sum of the three numbers).
 * Every rank must make the same sequence of these calls.  libnccl.so.2 is loaded with dlopen on first use (GBP_ERR_COMM when
 * it is missing); a single-GPU client never needs it. */
#define GBP_COMM_ID_BYTES 128
int gbp_comm_unique_id(void* id /* GBP_COMM_ID_BYTES, out */);
int gbp_comm_version(void);                                   /* NCCL version code, 0 when libnccl cannot be loaded */
typedef struct gbp_comm_s* gbp_comm;
int gbp_comm_create(const void* id /* GBP_COMM_ID_BYTES */, int rank, int nranks, int device, gbp_comm* out);
int gbp_comm_destroy(gbp_comm c);                             /* GBP_ERR_STATE while graphs are attached */
int gbp_ba_attach_comm(gbp_handle h, gbp_comm c);             /* the graph must be detached or destroyed before the communicator */
int gbp_ba_detach_comm(gbp_handle h);
int gbp_ba_comm_info(gbp_handle h, int32_t out[2] /* rank, nranks (0, 1 without a communicator) */);
/* For host-driven schedules and phase timing: after gbp_ba_sweep_local(... | GBP_STAGE_BELIEFS | GBP_STAGE_DEFER_LANDMARKS),
 * all-gather the chunk sums and finalise the keyframe beliefs on the handle's stream (gbp_ba_landmark_update does the rest). */
int gbp_ba_exchange(gbp_handle h);

/* n x synchronous_iteration(robustify, local_relin) (gbp/gbp.py:86-92; the loop of ba.py:84-105 without the client's
 * per-iteration reads): sweep + belief update (with a communicator: + the keyframe exchange), replayed from a CUDA graph. */
int gbp_ba_iterate(gbp_handle h, int n_iters, int robustify, int local_relin);
/* FactorGraph.update_all_beliefs alone (gbp/gbp.py:56-58). */
int gbp_ba_update_beliefs(gbp_handle h);

/* BAFactorGraph.are (gbp/gbp_ba.py:61-69), FactorGraph.energy (gbp/gbp.py:36-44) and the count of
 * factors with iters_since_relin == 0 (ba.py:97-100) over the local edges.  out[0] = sum of |r|
 * (NOT yet divided by F), out[1] = energy, out[2] = count.  Synchronises the stream. */
int gbp_ba_metrics(gbp_handle h, double out[3]);

/* Snapshot = what the client of ba.py looks at between two sweeps (ba.py:95-103): the three numbers of
 * gbp_ba_metrics and the means of all variables (what the viewer draws).  They are contiguous on the
 * device, so a snapshot is ONE small device->host copy into a caller-owned PAGE-LOCKED region
 * (gbp_host_alloc) of out[0] bytes with the metrics at byte offset out[1], the keyframe means
 * (GBP_F_CAM_MU, C x 6) at out[2] and the landmark means (GBP_F_LMK_MU, L x 3) at out[3]. */
int gbp_ba_snapshot_layout(gbp_handle h, uint64_t out[4]);
/* Enqueue metrics + copy behind the work already on the stream; no host synchronisation.  The region must
 * stay alive until gbp_ba_snapshot_wait returns. */
int gbp_ba_snapshot_async(gbp_handle h, void* region);
int gbp_ba_snapshot_wait(gbp_handle h);
/* One synchronous_iteration (gbp/gbp.py:86-92) immediately followed by a snapshot, replayed as ONE CUDA
 * graph (sweep, beliefs, metrics, copy): the whole loop body of ba.py:95-105 in a single launch. */
int gbp_ba_iterate_snapshot(gbp_handle h, int robustify, int local_relin, void* region);
/* Page-locked host memory for the destination buffers of reads (NULL when no device / out of memory:
 * callers then use ordinary memory). */
void* gbp_host_alloc(size_t bytes);
void gbp_host_free(void* p);

/* Field access (synchronises the stream).  `bytes` must equal rows x row bytes of the field. */
int gbp_ba_read(gbp_handle h, int field, void* host_dst, size_t bytes);
int gbp_ba_write(gbp_handle h, int field, const void* host_src, size_t bytes);
/* Fill every factor's iters_since_relin with `value` on the device (the loop of ba.py:91-93). */
int gbp_ba_fill_iters(gbp_handle h, int32_t value);
/* Raw device pointer of a variable-indexed field or GBP_F_CAM_PARTIAL (for zero-copy NCCL / torch). */
int gbp_ba_device_ptr(gbp_handle h, int field, void** dev_ptr, size_t* bytes);
/* Change the relinearisation / damping parameters of the FactorGraph object (gbp/gbp.py:28-34). */
int gbp_ba_set_params(gbp_handle h, double eta_damping, double beta, int32_t num_undamped_iters,
                      int32_t min_linear_iters);

int gbp_ba_synchronize(gbp_handle h);
/* Engine tuning knobs (no reference counterpart; measurement scripts): GBP_TUNE_PREFETCH_TILES = L2 prefetch distance of the
 * streaming build in tiles (0 = off; the automatic choice is ~38 k edges ahead); GBP_TUNE_BELIEF_LANES = lanes per landmark in the
 * belief kernel (1, 8, 32; 0 = chosen by the number of landmarks). */
enum { GBP_TUNE_PREFETCH_TILES = 3, GBP_TUNE_BELIEF_LANES = 5 };
int gbp_ba_tune(gbp_handle h, int knob, int64_t value);
/* Timing helper for benchmarks: runs n_iters iterations bracketed by CUDA events on the handle's
 * stream; *ms_total = elapsed device time, *ms_msg_kernel = summed time of the message kernel alone
 * (measured with per-launch events when per_kernel != 0, else 0). */
int gbp_ba_time_iterations(gbp_handle h, int n_iters, int robustify, int local_relin, int per_kernel,
                           float* ms_total, float* ms_msg_kernel);
/* Number of kernel launches issued by this handle since creation (bench "gpu_launches"). */
int64_t gbp_ba_launch_count(gbp_handle h);

/* Standalone evaluation of the reprojection factor model on the device, for parity tests of
 * reprojection.meas_fn / jac_fn (gbp/factors/reprojection.py:12-44): x is n x 9, out_h n x 2,
 * out_J n x 18 (host pointers). */
int gbp_reprojection_eval(const double* x, int64_t n, const double K[4], int device, double* out_h,
                          double* out_J);

/* ---- BAL-style problem files (host only; no device needed) -------------------------------------------
 * Native replacement of read_balfile (utils/read_balfile.py:4-37): same acceptance rules (leading blank /
 * "# ..." lines skipped; first four tokens of a measurement line; first token of a parameter line).
 * gbp_bal_open parses the whole file; gbp_bal_copy fills caller-owned arrays (any pointer may be NULL). */
typedef struct gbp_bal_file gbp_bal_file;
int gbp_bal_open(const char* path, gbp_bal_file** out);
int gbp_bal_sizes(const gbp_bal_file* bal, int64_t out[3] /* keyframes, landmarks, measurements */);
int gbp_bal_copy(const gbp_bal_file* bal, int32_t* cam_id, int32_t* lmk_id, double* z /* F x 2 */,
                 double* cam_means /* C x 6 */, double* lmk_means /* L x 3 */, double K4[4]);
void gbp_bal_close(gbp_bal_file* bal);

/* ---- graphs of pairwise LINEAR factors between variables of equal dimension (the path of ndim_posegraph.py) ---------------
 * Device counterpart of the generic gbp.FactorGraph(nonlinear_factors=False) (gbp/gbp.py:11-153) for factors that join two
 * variables of `dim` <= 6 dofs each.  A linear factor never relinearises: Factor.compute_factor (gbp/gbp.py:267-294) is evaluated
 * once by the caller (the Python callables meas_fn / jac_fn, e.g. gbp/factors/linear_displacement.py:8-14) and passed as
 *   J   F x dim x 2 dim row-major: jac_fn(x0) = [J_i | J_j], rows beyond the measurement dimension zero (dim(z) <= dim),
 *   b   F x dim: J x0 + z - meas_fn(x0) (gbp/gbp.py:289), zero-padded,   var  F: gauss_noise_std^2.
 * Matrices of THIS interface are full dim x dim row-major (the clients hold NumPy matrices).  adj_ptr / adj_msg: for every variable
 * the incoming messages in VariableNode.adj_factors order (gbp/gbp.py:182-188), an entry = 2 * factor + (0: the variable is the
 * factor's first, 1: its second).  Messages start at zero (gbp/gbp.py:228); gbp_lin_set_messages uploads the client's. */
typedef struct gbp_lin_graph* gbp_lin_handle;
enum { GBP_LIN_BELIEFS = 0, GBP_LIN_MESSAGES = 1 };
int gbp_lin_create(int32_t dim, int32_t V, int64_t F, const int32_t* var_i, const int32_t* var_j, const double* J, const double* b,
                   const double* var, const double* prior_eta /* V x dim */, const double* prior_lam /* V x dim x dim */,
                   const int32_t* adj_ptr /* V + 1 */, const int32_t* adj_msg /* 2 F */, double eta_damping, int device, void* stream,
                   gbp_lin_handle* out);
int gbp_lin_destroy(gbp_lin_handle h);
/* Factor.messages of every factor (gbp/gbp.py:228): eta 2F x dim, lam 2F x dim x dim, message 2 f + side. */
int gbp_lin_set_messages(gbp_lin_handle h, const double* eta, const double* lam);
/* FactorGraph.update_all_beliefs (gbp/gbp.py:56-58, 176-198). */
int gbp_lin_update_beliefs(gbp_lin_handle h);
/* n x FactorGraph.synchronous_iteration() of a linear graph (gbp/gbp.py:86-92): messages with graph-level eta damping, beliefs. */
int gbp_lin_iterate(gbp_lin_handle h, int n_iters);
/* FactorGraph.energy (gbp/gbp.py:36-44). */
int gbp_lin_energy(gbp_lin_handle h, double* out);
/* GBP_LIN_BELIEFS: eta V x dim, lam V x dim x dim, mu V x dim (any may be NULL); GBP_LIN_MESSAGES: eta, lam as in set_messages. */
int gbp_lin_read(gbp_lin_handle h, int what, double* eta, double* lam, double* mu);
/* FactorGraph.joint_distribution_inf / _cov (gbp/gbp.py:94-144): the joint precision and information vector assembled on the
 * device (eta_out n, lam_out n x n, n = V dim; may be NULL), mu = Lambda^-1 eta by a dense Cholesky solve on the device and, when
 * sigma != NULL, sigma = Lambda^-1 (n x n). */
int gbp_lin_joint_solve(gbp_lin_handle h, double* mu, double* sigma, double* eta_out, double* lam_out);
int64_t gbp_lin_launch_count(gbp_lin_handle h);

#ifdef __cplusplus
}
#endif
#endif /* GBP_B200_H */
