/*
 * nrs_b200.h -- C ABI of the B200-native hot paths of navtech-radar-slam.
 *
 * This is the drop-in boundary.  The reference has no FFI layer; its "operator API" is the
 * set of C++ signatures its two ROS nodes call.  The C++ adapters in
 * navtech-radar-slam_b200/cpp/ keep those signatures and forward to the entry points below;
 * INTEGRATION.md shows the binding a maintainer adds on the reference side.
 *
 * Conventions
 *   - plain pointers and sizes only; caller owns every buffer; the library never frees a
 *     caller buffer and never keeps a caller pointer after the call returns.
 *   - every function returns 0 on success or a negative NRS_E* code; no exceptions cross the
 *     ABI; nrs_last_error() returns a thread-local message for the last failure.
 *   - "host" pointers are ordinary CPU memory (pinned is faster, not required);
 *     "_dev" entry points take device pointers, enqueue on the context stream and do NOT
 *     synchronise (call nrs_ctx_sync).
 *   - matrices use the reference's in-memory layout: Eigen::MatrixXd is column-major, so a
 *     20x60 descriptor is 60 columns of 20 doubles (ring index fastest), 9600 bytes; a 3xK
 *     point matrix is K columns of (x,y,z).  cv::Mat images are row-major.
 *   - there is no CPU fallback: every compute entry point fails with NRS_ENODEVICE when no
 *     CUDA device / sm_100 kernel image is available.
 *
 * Paths are relative to /root/reference; ORORA/ = outlier-robust-radar-odometry/,
 * SC/ = pgo/SC-A-LOAM/include/scancontext/.
 */
#ifndef NRS_B200_H
#define NRS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NRS_OK          0
#define NRS_EINVAL     -1   /* bad argument (the reference throws std::invalid_argument) */
#define NRS_ENODEVICE  -2   /* no CUDA device / kernel image: there is no CPU fallback     */
#define NRS_ECUDA      -3   /* CUDA runtime error, see nrs_last_error()                    */
#define NRS_ENOMEM     -4
#define NRS_ECAPACITY  -5   /* caller's output buffer too small                            */
#define NRS_ENCCL      -6

#define NRS_SC_RINGS    20  /* SC/Scancontext.h:85 */
#define NRS_SC_SECTORS  60  /* SC/Scancontext.h:86 */
#define NRS_SC_DESC     (NRS_SC_RINGS * NRS_SC_SECTORS)
#define NRS_SC_MAX_BATCH 1024 /* queries per launch of the batched search (larger batches are split) */

typedef struct nrs_ctx   nrs_ctx;    /* one CUDA device + stream + scratch            */
typedef struct nrs_sc_db nrs_sc_db;  /* device-resident keyframe descriptor database  */

/* Result of one database search: the (min_dist, nn_idx, nn_align) triple of
 * SCManager::detectLoopClosureID (SC/Scancontext.cpp:362-396). 16 bytes. */
typedef struct nrs_sc_result {
    double  dist;    /* 1e7 when nothing beat the reference's initial value */
    int32_t idx;     /* keyframe index in [first,last) (or the candidate id); 0 when none */
    int32_t shift;   /* column shift (sector bins) of the best alignment */
} nrs_sc_result;

/* ---- library / context -------------------------------------------------------------- */
int         nrs_version(void);
const char *nrs_last_error(void);
/* number of this library's kernels launched by the calling process so far (bench.py's
 * gpu_launches claim is a difference of two reads). */
uint64_t    nrs_kernel_launches(void);

int   nrs_ctx_create(int device, nrs_ctx **out);
int   nrs_ctx_destroy(nrs_ctx *ctx);
int   nrs_ctx_sync(nrs_ctx *ctx);
void *nrs_ctx_stream(nrs_ctx *ctx);          /* the cudaStream_t all work is enqueued on */
int   nrs_ctx_device(nrs_ctx *ctx);
int   nrs_ctx_sm_count(nrs_ctx *ctx);

/* ---- ScanContext loop detector (SC/Scancontext.{h,cpp}) ------------------------------- */

/* SCManager::makeScancontext (Scancontext.cpp:151-195).  xyz = P points (x,y,z) float32,
 * host.  desc_out = 1200 doubles column-major, host.  lidar_height/max_radius are
 * LIDAR_HEIGHT (2.0) and PC_MAX_RADIUS (80.0) of Scancontext.h:83,87. */
int nrs_sc_make(nrs_ctx *ctx, const float *xyz_Px3, int P, double lidar_height,
                double max_radius, double *desc_out);

/* makeRingkeyFromScancontext / makeSectorkeyFromScancontext (Scancontext.cpp:198-227).
 * Either output may be NULL. */
int nrs_sc_keys(nrs_ctx *ctx, const double *desc, double *ringkey20, double *sectorkey60);

/* SCManager::fastAlignUsingVkey (Scancontext.cpp:93-113): argmin over 60 column shifts of
 * ||vkey1 - shift(vkey2, s)||, strict '<' (lowest shift on ties). */
int nrs_sc_fast_align(nrs_ctx *ctx, const double *vkey1_60, const double *vkey2_60, int32_t *shift);

/* SCManager::distDirectSC (Scancontext.cpp:69-90): 1 - mean column cosine similarity over the
 * columns where neither side is all-zero (NaN when there is none, like the reference). */
int nrs_sc_dist_direct(nrs_ctx *ctx, const double *sc1, const double *sc2, double *dist);

/* SCManager::distanceBtnScanContext(sc1, sc2) (Scancontext.cpp:116-148): sc1 is the query,
 * sc2 the candidate whose columns are shifted. */
int nrs_sc_pair(nrs_ctx *ctx, const double *sc1, const double *sc2, double *dist, int32_t *shift);

/* The database mirrors SCManager::polarcontexts_ (Scancontext.h:111) on the device:
 * keyframe i occupies bytes [9600*i, 9600*(i+1)) of the descriptor array, verbatim.  Beside it
 * the database keeps, per keyframe, the f64 sector key and column norms (960 B) and the
 * f32 filter record the search streams (nrs_sc_filter_record_bytes() bytes), both derived
 * on the device at append time.  Append is thread-safe against query;
 * a query sees a consistent prefix (the reference's node appends and queries from two
 * threads without a lock, SC/../src/laserPosegraphOptimization.cpp:492 vs :561). */
int nrs_sc_db_create(nrs_ctx *ctx, int capacity_hint, nrs_sc_db **out);
int nrs_sc_db_destroy(nrs_sc_db *db);
int nrs_sc_db_size(nrs_sc_db *db);
int nrs_sc_db_clear(nrs_sc_db *db);
/* saveScancontextAndKeys (Scancontext.cpp:236-246): append `count` descriptors from host. */
int nrs_sc_db_append(nrs_sc_db *db, const double *desc_host, int count);
/* makeAndSaveScancontextAndKeys (Scancontext.cpp:249-260): build the descriptor from a
 * cloud on the device and append it; desc_out (nullable, host) receives a copy so the
 * adapter can keep the reference's public polarcontexts_ vector populated. */
int nrs_sc_db_append_cloud(nrs_sc_db *db, const float *xyz_Px3, int P, double lidar_height,
                           double max_radius, double *desc_out);
/* copy descriptors [first, first+count) back to host */
int nrs_sc_db_read(nrs_sc_db *db, int first, int count, double *desc_host);
/* device base of the descriptor array (9600 B per keyframe, the reference's layout) */
const double *nrs_sc_db_device_ptr(nrs_sc_db *db);
/* The candidate step of detectLoopClosureID (Scancontext.cpp:361-377): the k ring keys of
 * keyframes [first, last) nearest to `ringkey20` in nanoflann's L2_Adaptor arithmetic
 * (nanoflann.hpp:375-412, float, groups of four) -- what the reference's kd-tree returns,
 * without a tree: ascending distance, equal distances by ascending index, NaN distances never
 * listed; idx_out[j] = -1 where fewer than k exist.  The database keeps every keyframe's ring
 * key (polarcontext_invkeys_mat_, Scancontext.h:114) since append.  1 <= k <= 16. */
int nrs_sc_knn(nrs_sc_db *db, const float *ringkey20_host, int first, int last, int k,
               int32_t *idx_out, float *d2_out);
/* detectLoopClosureID exactly as the reference runs it (Scancontext.cpp:331-399): ring key of the
 * query, its k nearest ring keys among keyframes [first, last) (the set the kd-tree holds), then
 * distanceBtnScanContext over those k in tree order, strict '<'.  Everything stays on the
 * device; cand_out (nullable, k ids) reports the candidates. */
int nrs_sc_detect_knn(nrs_sc_db *db, const double *query_host, int first, int last, int k,
                      int32_t *cand_out, nrs_sc_result *result_host);
/* The same for Q queries in one call (Q descriptors in, Q results and, nullable, Q x k candidate
 * ids out): the k-NN kernel takes the query as its second grid dimension, the candidate search
 * runs over every query's own k ids -- two launches for the whole batch.  Results equal Q calls
 * of nrs_sc_detect_knn over the same [first, last) (offline / between-session use; the online
 * node asks one query per keyframe). */
int nrs_sc_detect_knn_batch(nrs_sc_db *db, const double *queries_host, int Q, int first, int last,
                            int k, int32_t *cand_out, nrs_sc_result *results_host);
/* bytes per keyframe that an exhaustive query streams from HBM: the f32 filter record of the
 * general-valued search, and the 256-byte bit-mask record of the {0,c} search (below) */
int nrs_sc_filter_record_bytes(void);
int nrs_sc_bin_record_bytes(void);
/* 1 when a 20x60 descriptor (host memory, no GPU involved) is {0, c} with a c the bit-mask kernel accepts (finite,
 * positive, exact in float, moderate magnitude) -- the host-side test that routes a query to the bit-mask or the
 * general-valued kernel; +-0.0 count as empty bins, NaN as a value. */
int nrs_sc_descriptor_is_binary(const double *desc1200);

/* {0,c} descriptors.  In this pipeline every occupied bin holds the same value (ORORA publishes
 * z = 1.0, ORORA/src/odometry.cpp:305-316; makeScancontext adds LIDAR_HEIGHT, Scancontext.cpp:168),
 * so distanceBtnScanContext is a function of 60 20-bit column masks.  The database classifies every
 * keyframe at append time; while ALL stored keyframes are {0,c} (c > 0 with <= 24 significant bits)
 * exhaustive and candidate searches run on the bit-mask kernel (sc_search_bin.cu) -- same results,
 * bit for bit (the f64 certification reproduces the reference's sums exactly from the masks) -- else
 * on the general-valued f32-filter kernel.  nrs_sc_db_is_binary: 1 when the bit-mask search applies.
 * nrs_sc_db_set_search_mode(db, 1) forces the general-valued kernel (tests, benchmarks); 0 = auto. */
int nrs_sc_db_is_binary(nrs_sc_db *db);
int nrs_sc_db_set_search_mode(nrs_sc_db *db, int mode);

/* Multi-GPU (one process per GPU, keyframe g stored on rank g % world as local slot g / world):
 * result exchange through peer memory instead of a collective.  Each rank creates its slot array
 * and exports a CUDA IPC handle (NRS_SC_XCHG_HANDLE_BYTES bytes); the host program gathers the
 * world's handles any way it likes (MPI, torch.distributed, a file) and hands them to _open in
 * rank order.  nrs_sc_query_dev_sharded then searches local slots [first, last) and the search
 * kernel's last CTA writes this rank's record into every peer's HBM over NVLink, waits for the
 * world's records and merges them: *result_dev holds the GLOBAL (dist, idx, shift) on every
 * rank, ties to the lowest global index like the sequential scan (Scancontext.cpp:380-396).
 * All ranks must issue the same sequence of sharded queries on a database.  A peer that does
 * not answer within ~3 s yields idx == -2 (dist 1e7) instead of a hang. */
#define NRS_SC_XCHG_HANDLE_BYTES 64
int nrs_sc_xchg_create(nrs_sc_db *db, int world, int rank, unsigned char *handle_out);
int nrs_sc_xchg_open(nrs_sc_db *db, const unsigned char *handles_world_x_64);
int nrs_sc_query_dev_sharded(nrs_sc_db *db, const double *query_dev, int first, int last,
                             nrs_sc_result *result_dev);
/* Q queries, ONE exchange: every rank's per-query records travel while later queries are still being
 * scored and the world's records are merged once per launch ({0,c} databases; otherwise Q exchanges). */
int nrs_sc_query_batch_dev_sharded(nrs_sc_db *db, const double *queries_dev, int Q, int first, int last,
                                   nrs_sc_result *results_dev);
/* how long the merge waits for a peer's records before reporting idx == -2 (default ~3000 ms) */
int nrs_sc_xchg_set_timeout_ms(nrs_sc_db *db, int ms);

/* The collective form (north_star: "one NCCL all-reduce of the per-shard min distance over NVLink",
 * SURVEY 8e): local search on every rank, ncclAllReduce(min) on the f64 distances, then ncclAllReduce(min)
 * on int64 (global idx << 8 | shift) where only ranks whose distance equals the global minimum contribute
 * (the others send INT64_MAX) -- ties resolve to the lowest global index like the sequential scan
 * (Scancontext.cpp:380-396).  libnccl.so.2 is opened at run time (dlopen); NRS_ENCCL when it is missing or
 * a call fails.  nrs_nccl_get_unique_id: rank 0 creates the 128-byte id, the host distributes it (any
 * channel) and every rank calls nrs_sc_nccl_init (collective, blocks until all ranks joined);
 * nrs_sc_nccl_attach uses a communicator the host already owns (ncclComm_t, not destroyed here). */
#define NRS_NCCL_UNIQUE_ID_BYTES 128
int nrs_nccl_available(void);
int nrs_nccl_get_unique_id(unsigned char *id_out_128);
int nrs_sc_nccl_init(nrs_sc_db *db, int world, int rank, const unsigned char *unique_id_128);
int nrs_sc_nccl_attach(nrs_sc_db *db, void *nccl_comm, int world, int rank);
int nrs_sc_query_batch_dev_nccl(nrs_sc_db *db, const double *queries_dev, int Q, int first, int last,
                                nrs_sc_result *results_dev);
/* Host queries in, host records out over local slots [first,last): the sharded search as a C++ host calls
 * it (detectLoopClosureID on a keyframe-sharded SCManager).  Uses the peer-memory exchange when
 * nrs_sc_xchg_open succeeded, else the NCCL reduction above.  Identical records on every rank. */
int nrs_sc_query_batch_sharded(nrs_sc_db *db, const double *queries_host, int Q, int first, int last,
                               nrs_sc_result *results_host);
/* diagnostics: %globaltimer (ns) stamps of the database's most recent search, taken by its last
 * CTA: [0] tail entered, [1] local result ready, [2] record + flag written to the peers
 * (sharded queries only), [3] result written.  Synchronises the stream. */
int nrs_sc_db_tail_times(nrs_sc_db *db, uint64_t *t_ns4);

/* The search loop of detectLoopClosureID (Scancontext.cpp:380-396).
 *   candidate_idx == NULL : exhaustive over keyframes [first, last)   (north_star)
 *   candidate_idx != NULL : exactly those n_cand ids, in that order   (reference: 3 kd-tree ids)
 * Strict '<' everywhere: lowest shift, then first keyframe wins ties; a keyframe whose
 * distance is NaN (no effective column) never wins; nothing found -> {1e7, 0, 0}. */
int nrs_sc_query(nrs_sc_db *db, const double *query_host, int first, int last,
                 const int32_t *candidate_idx, int n_cand, nrs_sc_result *result_host);
/* Per-keyframe distanceBtnScanContext(query, db[i]) for i in [first,last):
 * dist_out/shift_out have last-first entries (host). */
int nrs_sc_query_all(nrs_sc_db *db, const double *query_host, int first, int last,
                     double *dist_out, int32_t *shift_out, nrs_sc_result *result_host);
/* Device-resident, asynchronous variant: query_dev (1200 doubles) and result_dev live in
 * device memory; enqueued on the context stream, no host synchronisation.  idx_offset is
 * added to the reported index (a shard's global offset). */
int nrs_sc_query_dev(nrs_sc_db *db, const double *query_dev, int first, int last,
                     int32_t idx_offset, nrs_sc_result *result_dev);
/* Batched search: Q queries (Q x 1200 doubles, query-major) against keyframes [first,last) -- the
 * detectLoopClosureID search loop (Scancontext.cpp:380-396) for Q query descriptors at once (multi-
 * session relocalisation, offline loop mining, Scancontext.cpp:267-328).  results[q] is exactly what
 * nrs_sc_query returns for query q.  On a {0,c} database this is ONE kernel launch whose CTAs share
 * the keyframe records between queries; otherwise Q launches.  _dev: device pointers, enqueued on
 * the context stream, no synchronisation. */
int nrs_sc_query_batch(nrs_sc_db *db, const double *queries_host, int Q, int first, int last,
                       nrs_sc_result *results_host);
int nrs_sc_query_batch_dev(nrs_sc_db *db, const double *queries_dev, int Q, int first, int last,
                           int32_t idx_offset, nrs_sc_result *results_dev);

/* ---- ORORA solver (ORORA/include/ORORA.hpp, ORORA/include/solver/nonminimal_solver.hpp) ---- */

/* NonMinimalSolver::Params (nonminimal_solver.hpp:104-162) + the ORORA ctor arguments
 * (ORORA.hpp:15-28); defaults = ORORA/config/orora_params.yaml as loaded by odometry.cpp:79-135. */
typedef struct nrs_solver_params {
    double  noise_bound;                  /* 0.75 */
    double  cbar2;                        /* noise_bound_coeff, 1.0 */
    int32_t rotation_max_iterations;      /* 100 */
    double  rotation_gnc_factor;          /* 1.4 */
    double  rotation_cost_threshold;      /* 1e-6 */
    double  noise_bound_radial;           /* 0.1 m */
    double  noise_bound_tangential;       /* radians (9 deg) */
    int32_t use_doppler_compensation;     /* 1 */
    double  doppler_beta;                 /* 0.049, odometry.cpp:109 */
    int32_t check_stop_motion;            /* 1 */
    int32_t num_feat_thr_for_stop_motion; /* 700 */
    int64_t clique_node_limit;            /* 0 = default (2e6 branch-and-bound nodes) */
} nrs_solver_params;

typedef struct nrs_solver_info {
    int32_t valid;              /* solution_.valid (nonminimal_solver.hpp:718,820) */
    int32_t stop_motion;        /* 1: skipped by the stop-motion gate (ORORA.hpp:74-79) -> Identity */
    int32_t n_edges;            /* scale-consistent TIMs = edges of the inlier graph */
    int32_t M;                  /* getNumMaxCliqueInliers() */
    int32_t gnc_iters;
    int32_t gnc_early_return;   /* GNC-TLS returned at initialisation (:348-352) */
    int32_t n_rot_inliers;      /* getNumRotaionInliers() */
    int32_t n_final_inliers;    /* getFinalInliersIndices().size() */
    int32_t clique_exact;       /* 0 when the node limit stopped the exact search early */
    int32_t pad_;
    int64_t clique_nodes;
    double  R[4];               /* 2x2 rotation, column-major */
    double  t[2];
    double  cost;
} nrs_solver_info;

/* diagnostics: %globaltimer (ns) stamps of the most recent nrs_orora_solve in this process --
 * clique kernel [0] start [1] adjacency in shared memory [2] greedy bound [3] core pruning
 * [4] exact search done; tail kernel [5] start [6] chain TIMs [7] GNC-TLS [8] COTE x [9] COTE y */
int nrs_orora_last_timings(uint64_t *t_ns12);
void nrs_solver_default_params(nrs_solver_params *p);

/* ORORA::calcTransform (ORORA.hpp:37-93) minus the velocity bookkeeping the caller keeps:
 * Doppler compensation of both point sets with |vel_x| (:56-61,139-151), stop-motion gate
 * (:74-79), then NonMinimalSolver::computeTransformation with source = curr, target = prev
 * (:83-85; nonminimal_solver.hpp:678-825): TIMs + scale consistency -> inlier graph ->
 * exact maximum clique -> chain TIMs -> GNC-TLS rotation -> anisotropic COTE translation.
 *   prev_2xN, curr_2xN : matched points, column-major 2xN doubles (host)
 *   vel_x              : curr_vel_se3_(0) of the previous call = T(0,3)/delta_t
 *   T_out              : 4x4 column-major; Identity when the solver fails or the gate fires
 *   clique_out, final_inliers_out : nullable, room for N int32 (sorted correspondence ids)
 * N is limited to 1024 correspondences (NRS_ECAPACITY above). */
int nrs_orora_solve(nrs_ctx *ctx, const double *prev_2xN, const double *curr_2xN, int N,
                    const nrs_solver_params *p, double vel_x, double T_out[16],
                    int32_t *clique_out, int32_t *final_inliers_out, nrs_solver_info *info_out);

/* ---- ORORA front-end (ORORA/include/radar_loader.hpp, src/features.cpp, src/radar_utils.cpp) ---- */

typedef struct nrs_frame nrs_frame;   /* device-resident RadarFeature (radar_loader.hpp:16-33) */

/* the reference's own values (`#define CTS350 0`, `#define CIR204 1`, radar_utils.hpp:25-26) */
#define NRS_NAVTECH_CTS350 0
#define NRS_NAVTECH_CIR204 1

/* RadarLoader's dataset constants (radar_loader.hpp:48-62) and extractor knobs (:283-297). */
typedef struct nrs_extract_params {
    int32_t navtech_version;        /* what polar_to_cartesian_radar_img is given (radar_utils.cpp:292-296): NRS_NAVTECH_CTS350 =
                                     * uniform azimuth step, NRS_NAVTECH_CIR204 = get_azimuth_index.  RadarLoader::extractFeature
                                     * passes none, i.e. CTS350, for BOTH datasets (radar_loader.hpp:105-107) -- the default here */
    int32_t range_bins;             /* load_radar's navtech_version: 3768 (CTS350, oxford) / 3360 (CIR204, mulran) */
    float   radar_resolution;       /* m / bin: 0.05952 (mulran), 0.0432 (oxford) */
    float   cart_resolution;        /* 0.2592 m / px */
    int32_t cart_pixel_width;       /* 964 */
    int32_t min_range;              /* 58 bins */
    int32_t max_points;             /* 10000 */
    int32_t patch_size;             /* 21 */
    int32_t use_voxelization;       /* 1 */
    int32_t interpolate_crossover;  /* 1 */
    double  voxel_size;             /* 0.6 m */
    float   nndr;                   /* 0.8 */
    int32_t keypoint_extraction;    /* 0 = "cen2019" (default), 1 = "cen2018" (radar_loader.hpp:95-99) */
    float   zq;                     /* cen2018: 3.0 (radar_loader.hpp:284) */
    int32_t sigma_gauss;            /* cen2018: 17, odd (radar_loader.hpp:285) */
} nrs_extract_params;

/* counts and scalars of one extraction, for callers and stage-parity tests */
typedef struct nrs_frame_info {
    int32_t n_candidates;   /* pixels with h > mean_h (features.cpp:217-223) */
    int32_t j_consumed;     /* candidates consumed by the marking loop (:231-249) */
    int32_t l_regions;      /* regions counted, <= max_points */
    int32_t K_polar;        /* cen2019 keypoints = feat_in_polar_.cols() */
    int32_t K_voxel;        /* after voxelisation */
    int32_t K;              /* after the BEV border filter = feat_in_cart_.cols() = t_.size() */
    int32_t K_desc;         /* kp_.size() = desc_.rows after cv::ORB::compute's own border filter (<= K) */
    int32_t n_sorted;       /* candidates the marking loop had sorted before it: the top-candidate subset's size,
                             * or n_candidates when the full-frame sort ran (diagnostic; never changes a result) */
    float   maxg, mean, mean_h, cart_max;
    int64_t t_criteria;
} nrs_frame_info;

/* dataset = "mulran" | "oxford" (anything else -> NRS_EINVAL, the reference throws
 * std::invalid_argument, radar_loader.hpp:62) */
int nrs_extract_default_params(const char *dataset, nrs_extract_params *p);

/* cen2019features (features.hpp:36, features.cpp:197-299) as a stand-alone operator.
 * fft = R x C float32 row-major (host).  targets_3xKcap receives K columns (azimuth, range, 1),
 * column-major doubles; *K is always set; NRS_ECAPACITY if K > Kcap. */
int nrs_cen2019(nrs_ctx *ctx, const float *fft_RxC, int R, int C, int max_points, int min_range,
                double *targets_3xKcap, int Kcap, int *K);

/* cen2018features (features.hpp:27, features.cpp:44-131) as a stand-alone operator -- the extractor the
 * shipped launch file selects (run_orora.launch:11).  Same conventions as nrs_cen2019.  y_RxC (nullable,
 * host, R x C floats) receives the per-pixel response for stage-wise comparison (0 left of min_range). */
int nrs_cen2018(nrs_ctx *ctx, const float *fft_RxC, int R, int C, float zq, int sigma_gauss, int min_range,
                double *targets_3xKcap, int Kcap, int *K, float *y_RxC);

/* RadarLoader::extractFeature (radar_loader.hpp:74-141) on the decoded 8-bit rows of one
 * "oxford-form" polar PNG (what cv::imread hands load_radar, radar_utils.cpp:104-129):
 * row i = [int64 stamp | uint16 encoder | valid | R bins of power ...], row_pitch bytes apart.
 * load_radar -> cen2019features | cen2018features -> polar_to_cartesian_radar_img -> polar_to_cartesian_points ->
 * voxelize -> convert_to_bev -> ORB compute.  t_criteria = the stamp in the file name. */
int nrs_frame_extract(nrs_ctx *ctx, const uint8_t *raw_rows, int R, int row_pitch,
                      const nrs_extract_params *p, int64_t t_criteria, nrs_frame **out);
int nrs_frame_destroy(nrs_frame *f);
int nrs_frame_info_get(const nrs_frame *f, nrs_frame_info *info);
/* diagnostics: %globaltimer (ns) stamps of the most recent extraction's tail kernel on this context:
 * [0] entered, [1] row offsets scanned, [2] points made, [3] voxel keys made, [4] sorted, [5] centroids + BEV done;
 * [6], [7] are not stamps: the radix passes the candidate-subset sort and the voxel sort took */
int nrs_frontend_tail_times(nrs_ctx *ctx, uint64_t *t_ns8);
/* copy products to host; any pointer may be NULL.  Sizes: polar 3 x K_polar doubles (col-major),
 * feat_in_cart 3 x K doubles (row 2 = 0 with voxelisation, 1 without), t K int64, t_norm K doubles,
 * kp K_desc x 2 float (u, v), desc K_desc x 32 bytes, img W x W bytes.  K_desc < K happens when
 * cv::ORB drops a keypoint whose rounded coordinate hits the image border band; the reference then
 * indexes feat_in_cart_ with descriptor rows (radar_loader.hpp:181-182) and so does nrs_match. */
int nrs_frame_get(nrs_frame *f, double *polar_3xK0, double *feat_in_cart_3xK, float *kp_Kx2,
                  int64_t *t_K, double *t_norm_K, uint8_t *desc_Kx32, uint8_t *img_WxW);
/* stage outputs for parity tests: fft (R x C float), R mask (R x C bytes), pre-voxel Cartesian
 * points (2 x K_polar doubles); any pointer may be NULL.  Valid until the next extraction on
 * the same context. */
int nrs_frame_get_stages(nrs_frame *f, float *fft_RxC, uint8_t *marked_RxC, double *cart_pre_2xK0,
                         float *cart_f32_WxW);

/* RadarLoader::matchFeatures (radar_loader.hpp:143-205): brute-force Hamming kNN(2) of
 * prev (query) against curr (train), ratio test d0 < nndr * d1, ties -> lowest train index.
 * Outputs (host, capacity cap): query_idx/train_idx of the good matches in query order and the
 * gathered 2 x N matched points (prev_matched.feat_in_cart_, curr_matched.feat_in_cart_);
 * stamps prev_t/curr_t (nullable).  NRS_ECAPACITY if N > cap (N is still reported). */
int nrs_match(nrs_ctx *ctx, const nrs_frame *prev, const nrs_frame *curr, float nndr, int cap,
              int32_t *query_idx, int32_t *train_idx, double *prev_2xN, double *curr_2xN,
              int64_t *prev_t, int64_t *curr_t, int *N);

/* ---- loop verification behind the ScanContext detector (SC/../src/laserPosegraphOptimization.cpp:355-406) ---- */

/* pcl::IterativeClosestPoint as doICPVirtualRelative configures it (:376-381) and its acceptance gate (:390-391) */
typedef struct nrs_icp_params {
    double  max_correspondence_distance;   /* 150 m (:377) */
    int32_t max_iterations;                /* 100 (:378) */
    double  transformation_epsilon;        /* 1e-6 (:379) */
    double  euclidean_fitness_epsilon;     /* 1e-6 (:380) */
    double  fitness_threshold;             /* loopFitnessScoreThreshold 0.3 (:390) */
    int32_t min_correspondences;           /* 3 (pcl::Registration) */
} nrs_icp_params;
typedef struct nrs_icp_info {
    int32_t converged;          /* icp.hasConverged() */
    int32_t iterations;
    int32_t state;              /* 1 iterations, 2 transform, 3 abs mse, 4 rel mse, 5 no correspondences */
    int32_t n_correspondences;  /* of the last iteration */
    double  fitness;            /* icp.getFitnessScore(): mean squared nearest-neighbour distance after alignment */
    int32_t accepted;           /* converged && fitness <= fitness_threshold: the SC loop is kept */
    int32_t pad_;
} nrs_icp_info;
void nrs_icp_default_params(nrs_icp_params *p);
/* icp.setInputSource(src); icp.setInputTarget(tgt); icp.align(): point-to-point ICP of the current keyframe's
 * cloud against the submap around the loop candidate (both in the loop keyframe's frame, float x,y,z per point,
 * host).  T_out = icp.getFinalTransformation() as a column-major 4x4 (source -> target).  PCL is not vendored in
 * the reference: this restates its published algorithm (icp_kernels.cu header), double-precision sums. */
int nrs_icp_align(nrs_ctx *ctx, const float *src_xyz_Nsx3, int Ns, const float *tgt_xyz_Ntx3, int Nt,
                  const nrs_icp_params *p, double T_out[16], nrs_icp_info *info);

/* One iteration of the node's per-frame loop (odometry.cpp:153-284) in ONE call:
 *   curr = Loader.extractFeature(new frame); (prev_m, curr_m) = Loader.matchFeatures(prev, curr);
 *   T = orora_estimator.calcTransform(params, dt, prev_m, curr_m)
 * with the same results as nrs_frame_extract + nrs_match + nrs_orora_solve, but the matcher is enqueued behind the
 * extraction without waiting for it and the matched points reach the solver without a round trip through host
 * memory, and the solver is enqueued speculatively with the match count still on the device (ONE host
 * synchronisation per scan pair instead of four; the stop-motion gate and the capacity rule are applied by the host
 * once it sees the count, counts above the speculative launch's 512 are solved again with the host-sized launch).
 * prev: the previous frame's handle (same context); *curr_out: the new frame's handle (caller destroys); vel_x as in nrs_orora_solve;
 * *n_matches (nullable): matches fed to the solver. */
int nrs_orora_pair(nrs_ctx *ctx, const nrs_frame *prev, const uint8_t *raw_rows, int R, int row_pitch,
                   const nrs_extract_params *p, int64_t t_criteria, const nrs_solver_params *sp, double vel_x,
                   nrs_frame **curr_out, double T_out[16], int *n_matches, nrs_solver_info *info_out);

#ifdef __cplusplus
}
#endif
#endif /* NRS_B200_H */
