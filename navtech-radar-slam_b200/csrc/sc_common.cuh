// sc_common.cuh -- definitions shared by the ScanContext kernel files (sc_kernels.cu: general-valued
// f32-filter search, descriptor construction, keys, k-NN, host entry points; sc_search_bin.cu: the
// bit-mask search for {0,c} descriptors).  See sc_kernels.cu for the arithmetic contract.
#pragma once
#include <stddef.h>
#include <mutex>

#include "common.cuh"

namespace nrs_sc {

constexpr int R = NRS_SC_RINGS;      // 20
constexpr int S = NRS_SC_SECTORS;    // 60
constexpr int DESC = R * S;          // 1200 doubles = 9600 B (the reference's MatrixXd)
constexpr int NSHIFT = 7;            // 2 * round(0.5 * SEARCH_RATIO * 60) + 1   (Scancontext.cpp:123)
constexpr int RADIUS = 3;
constexpr double SENTINEL = 10000000.0;   // Scancontext.cpp:134,362

struct BlockResult {
    double dist;
    int    pos;
    int    idx;
    int    shift;
    int    pad;
};

// Eigen redux order over 20 values held in registers.
static __device__ __forceinline__ double redux20(const double (&v)[R])
{
    double a0 = v[0], a1 = v[1], b0 = v[2], b1 = v[3];
#pragma unroll
    for (int i = 4; i < R; i += 4) {
        a0 += v[i];
        a1 += v[i + 1];
        b0 += v[i + 2];
        b1 += v[i + 3];
    }
    a0 += b0;
    a1 += b1;
    return a0 + a1;
}

static __device__ __forceinline__ double redux20_dot(const double (&a)[R], const double (&b)[R])
{
    double a0 = a[0] * b[0], a1 = a[1] * b[1], b0 = a[2] * b[2], b1 = a[3] * b[3];
#pragma unroll
    for (int i = 4; i < R; i += 4) {
        a0 += a[i] * b[i];
        a1 += a[i + 1] * b[i + 1];
        b0 += a[i + 2] * b[i + 2];
        b1 += a[i + 3] * b[i + 3];
    }
    a0 += b0;
    a1 += b1;
    return a0 + a1;
}

// strict-'<' sequential scan == lexicographic min over (value, order key); NaN never wins.
static __device__ __forceinline__ bool better(double d1, int k1, double d2, int k2)
{
    return (d1 < d2) || (d1 == d2 && k1 < k2);
}

constexpr int KEYS_D = 2 * S;               // f64 side record: vk[60], kn[60]
constexpr unsigned long long MASK60 = (1ull << S) - 1ull;
constexpr unsigned long long HAZ_B = 1ull << 62, HAZ_C = 1ull << 63;
constexpr float EPS_B = 3.0517578125e-5f;   // 2^-15
constexpr float EPS_C = 7.62939453125e-6f;  // 2^-17
constexpr int BIN_WORDS = 64;               // 256-byte bit-mask record of a {0,c} keyframe (sc_search_bin.cu)
constexpr int XCHG_QCAP = NRS_SC_MAX_BATCH; // queries per launch the exchange slots are sized for

struct __align__(16) Survivor {          // 16 bytes, one per keyframe the filter could not rule out
    int   item;                          // position in the searched range / candidate list
    int   code;                          // s0 | candC << 8 | certain << 16
    float lo;                            // f32 minimum of the 7 distances
    int   pad;
};

// ... and the f64 side for the (rare) certification: keys and norms here, the columns straight from the query
struct QCold {
    const double *q64;                   // the query itself (global memory, the reference's layout)
    double qn[S], qv[S];
};
// search state shared by the launches of one database (device memory, 32 bytes)
struct SearchState {
    unsigned int ticket;                 // last-block-done counter (zero between launches)
    unsigned int gub;                    // smallest f32 upper bound of any keyframe's distance (order-preserving bits)
    unsigned int n_surv;                 // survivor list length
    unsigned int pad;
    unsigned long long t_ns[4];          // %globaltimer of the last launch: tail entered, local result ready,
                                         // record + flag written to the peers, world's records merged
};
static __device__ __forceinline__ unsigned long long globaltimer_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// Result exchange between the ranks of a keyframe-sharded database (one process per GPU): every
// rank owns a slot array in its own HBM, slot[parity][source rank], mapped into all peers (CUDA IPC over
// NVLink).  The last CTA of the search writes the rank's 16-byte record straight into every peer's
// array (payload, fence.sys, epoch flag), waits until the world's records of this epoch have arrived
// in its own array and merges them -- no collective call, no host round trip.  Slots are reused every
// second query: a rank can only start writing epoch e+2 after it has seen every peer's e+1 flag, and a
// peer raises that flag only after it has finished reading epoch e.
struct XchgSlot {
    double dist;
    int    idx;
    int    shift;
    unsigned int epoch;
    unsigned int pad[3];
};
struct XchgView {
    XchgSlot *const *peers;              // device array of `world` base pointers (entry `rank` = local array)
    int world, rank;
    unsigned int epoch;
    long long timeout_cycles;            // a peer that never answers must not hang the GPU
    int q_slot;                          // general-valued kernel: which of the launch group's query slots this launch fills
};
constexpr int XCHG_MAX_WORLD = 32;       // one lane per peer
constexpr long long XCHG_TIMEOUT_CYCLES = 6000000000ll;   // default ~3 s (nrs_sc_xchg_set_timeout_ms)

// warp-wide f32 minimum in one instruction (sm_100a: CREDUX.MIN.F32)
static __device__ __forceinline__ float warp_min_f(float v)
{
    float r;
    asm volatile("redux.sync.min.f32 %0, %1, 0xffffffff;" : "=f"(r) : "f"(v));
    return r;
}
// order-preserving f32 -> u32 (atomicMin on the result == min on the floats)
static __device__ __forceinline__ unsigned int ord_f32(float f)
{
    const unsigned int b = __float_as_uint(f);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
static __device__ __forceinline__ float unord_f32(unsigned int u)
{
    return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}
static __device__ __forceinline__ void st_release_sys(unsigned int *p, unsigned int v)
{
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
static __device__ __forceinline__ unsigned int ld_acquire_sys(const unsigned int *p)
{
    unsigned int v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
static __device__ __forceinline__ int wrap60(int s)          // s in [-60, 120) -> [0, 60)
{
    s = s < 0 ? s + S : s;
    return s >= S ? s - S : s;
}

// magnitude window in which the f32 filter's error bounds hold (see the K10 header)
static __device__ __forceinline__ bool hazard_col(double mx) { return !(mx <= 1e150) || (mx != 0.0 && mx < 1e-140); }
static __device__ __forceinline__ bool hazard_key(double mx) { return !(mx <= 1e15) || (mx != 0.0 && mx < 1e-15); }

// ---- exact f64 evaluation, the reference's order -----------------------------------------------------
// fastAlignUsingVkey over the candidate shifts (bit l of even <=> shift 2l, of odd <=> shift 2l + 1),
// ascending; one warp.
static __device__ __noinline__ int certify_shift(const QCold *__restrict__ qc, double *scr, const double *__restrict__ vk,
                                          unsigned int even, unsigned int odd, int lane)
{
    double fa_n = SENTINEL;
    int s0 = 0;
    while (even | odd) {
        const int se = even ? 2 * (__ffs(even) - 1) : 2 * S, so = odd ? 2 * (__ffs(odd) - 1) + 1 : 2 * S;
        const int s = min(se, so);
        if (se < so) even &= even - 1; else odd &= odd - 1;
        __syncwarp();
        for (int i = lane; i < S; i += 32) {
            const double d = qc->qv[i] - vk[wrap60(i - s)];
            scr[i] = d * d;
        }
        __syncwarp();
        double acc = 0.0;
        if (lane < 4) {
            acc = scr[lane];
#pragma unroll
            for (int i = 4; i < S; i += 4) acc += scr[lane + i];
        }
        const double a2 = __shfl_sync(0xffffffffu, acc, 2), a3 = __shfl_sync(0xffffffffu, acc, 3);
        const double a1 = __shfl_sync(0xffffffffu, acc, 1), a0 = __shfl_sync(0xffffffffu, acc, 0);
        const double nrm = sqrt((a0 + a2) + (a1 + a3));
        if (nrm < fa_n) { fa_n = nrm; s0 = s; }          // ascending s + strict '<' == reference scan
    }
    __syncwarp();
    return s0;
}

// distDirectSC for the shift slots in candC (bit d <=> shift s0 + d - 3); one warp; returns on all lanes
// the lexicographic (dist, shift) minimum, (SENTINEL, 0) when no slot has an effective column.
static __device__ __noinline__ void certify_dist(const QCold *__restrict__ qc, unsigned long long qzero, double *scr,
                                          const double *__restrict__ desc, const double *__restrict__ kn,
                                          unsigned long long kzero, int s0, unsigned int candC, int lane,
                                          double &dd, int &ds)
{
    dd = SENTINEL;
    ds = 0;
    while (candC) {
        const int d = __ffs(candC) - 1;
        candC &= candC - 1;
        const int s = wrap60(s0 + d - RADIUS);
        __syncwarp();
        // key column paired with query column c is (c - s) mod 60: rotate the key mask up by s;
        // effective columns (both norms non-zero), in column order, packed densely into scr[]
        const unsigned long long krot = s ? (((kzero << s) | (kzero >> (S - s))) & MASK60) : kzero;
        const unsigned long long eff = ~(qzero | krot) & MASK60;
#pragma unroll
        for (int h = 0; h < 2; ++h) {                 // both columns' loads in flight together
            const int c = lane + 32 * h;
            if (c >= S || !((eff >> c) & 1ull)) continue;
            const int kc = wrap60(c - s);
            // redux20_dot's order (four running sums over i % 4, combined (s0 + s2) + (s1 + s3)), streamed from
            // memory in 16-byte pairs so that no 20-element register arrays are needed
            const double2 *qa = reinterpret_cast<const double2 *>(qc->q64 + c * R);
            const double2 *kb = reinterpret_cast<const double2 *>(desc + kc * R);
            double a0, a1, b0, b1;
            {
                const double2 x0 = qa[0], y0 = kb[0], x1 = qa[1], y1 = kb[1];
                a0 = x0.x * y0.x; a1 = x0.y * y0.y; b0 = x1.x * y1.x; b1 = x1.y * y1.y;
            }
#pragma unroll
            for (int r = 2; r < R / 2; r += 2) {
                const double2 x0 = qa[r], y0 = kb[r], x1 = qa[r + 1], y1 = kb[r + 1];
                a0 += x0.x * y0.x; a1 += x0.y * y0.y; b0 += x1.x * y1.x; b1 += x1.y * y1.y;
            }
            a0 += b0;
            a1 += b1;
            const double dot = a0 + a1;
            const int pos = __popcll(eff & ((1ull << c) - 1ull));
            scr[pos] = dot / (qc->qn[c] * kn[kc]);
        }
        __syncwarp();
        double dist = SENTINEL;
        if (lane == 0) {
            const int cnt = __popcll(eff);
            double sum = 0.0;                       // distDirectSC's left-to-right sum (:83)
            for (int i = 0; i < cnt; ++i) sum = sum + scr[i];
            dist = 1.0 - sum / (double)cnt;
        }
        dist = __shfl_sync(0xffffffffu, dist, 0);
        if (better(dist, s, dd, ds)) { dd = dist; ds = s; }   // lexicographic (dist, shift) == ascending-shift scan
    }
    __syncwarp();
}

}  // namespace nrs_sc

struct nrs_sc_db {
    nrs_ctx     *ctx = nullptr;
    std::mutex   mu;
    double      *d_desc = nullptr;       // capacity x 1200 doubles: polarcontexts_ verbatim (Scancontext.h:111)
    double      *d_keys = nullptr;       // capacity x 120 doubles: sector key, column norms (f64)
    float       *d_frec = nullptr;       // capacity x FREC_F floats: filter records
    float       *d_rkeys = nullptr;      // capacity x 20 floats: ring keys (polarcontext_invkeys_mat_)
    unsigned int *d_brec = nullptr;      // capacity x 64 words: bit-mask records ({0,c} keyframes), sc_search_bin.cu
    unsigned int *d_ngen = nullptr;      // device counter of appended keyframes that are NOT {0,c}
    unsigned int *h_ngen = nullptr;      // pinned copy
    int          n_general = 0;          // its host value: 0 <=> the bit-mask search applies to the whole database
    int          search_mode = 0;        // 0 auto, 1 always the general-valued kernel (nrs_sc_db_set_search_mode)
    float       *d_knn_pd = nullptr;     // per-block partial k-NN lists (distance, index)
    int32_t     *d_knn_pi = nullptr;
    float       *d_knn_out = nullptr;    // [0,16) d2, [16,32) idx, [32,48) candidate ids, [48,68) query ring key, [68] ticket
    nrs_sc::Survivor    *d_surv = nullptr;       // capacity entries
    int          size = 0;
    int          capacity = 0;
    // per-db search scratch
    nrs_sc::BlockResult   *d_block = nullptr;
    nrs_sc::SearchState   *d_state = nullptr;
    nrs_sc_result *d_result = nullptr;
    double        *d_query = nullptr;
    nrs_sc_result *h_result = nullptr;   // pinned
    double        *h_query = nullptr;    // pinned
    int            max_blocks = 0;
    // bit-mask search scratch: per-query state, survivor lists, per-tile results (fallback), launch ticket
    nrs_sc::SearchState *d_bstate = nullptr;      // XCHG_QCAP entries
    unsigned int  *d_gticket = nullptr;
    nrs::DevBuf    bin_surv, bin_block;
    nrs::DevBuf    knn_rec, knn_part;   // batched k-NN: one 72-float record per query, per-block partial lists
    nrs::PinBuf    knn_host;            // ... and the records' pinned landing zone
    double        *d_queries = nullptr;  // batch staging: XCHG_QCAP x 1200 doubles (device) ...
    double        *h_queries = nullptr;  // ... and pinned host
    nrs_sc_result *d_results = nullptr;  // XCHG_QCAP records
    nrs_sc_result *h_results = nullptr;  // pinned
    // peer-memory result exchange (keyframe-sharded databases)
    nrs_sc::XchgSlot      *x_local = nullptr;    // 2 x world slots in this GPU's HBM, exported by CUDA IPC
    nrs_sc::XchgSlot     **x_peers_dev = nullptr;   // device array of world base pointers
    void          *x_opened[nrs_sc::XCHG_MAX_WORLD] = {};   // peer mappings to close
    int            x_world = 0, x_rank = 0;
    unsigned int   x_epoch = 0;
    bool           x_ready = false;
    long long      x_timeout_cycles = nrs_sc::XCHG_TIMEOUT_CYCLES;
    // collective exchange (sc_nccl.cu): communicator, staging for the two min all-reduces
    void          *n_comm = nullptr;
    bool           n_owned = false;
    int            n_world = 1, n_rank = 0;
    double        *n_dist = nullptr;     // [0,QCAP) local, [QCAP,2 QCAP) reduced
    long long     *n_key = nullptr;
    nrs_sc_result *d_results2 = nullptr; // host-pointer sharded entry point's device records
};

namespace nrs_sc {
// enqueue the bit-mask search of Q queries (device, Q x 1200 doubles) over keyframes [first, first + n_items)
// (or the candidate list) on the database's stream; every keyframe in range must be {0,c} (db->n_general == 0)
int bin_launch(nrs_sc_db *db, const double *d_queries, int Q, const int32_t *d_cand, int first, int n_items,
               int idx_scale, int idx_offset, nrs_sc_result *d_results, XchgView xv, int cand_stride = 0);
// Q searches enqueued on the context stream (db->mu and ctx->mu held by the caller): reported index =
// local slot * idx_scale + idx_offset; sharded = exchange the per-query records over peer memory in-kernel
int batch_dev_locked(nrs_sc_db *db, const double *queries_dev, int Q, int first, int last, int idx_scale,
                     int32_t idx_offset, nrs_sc_result *results_dev, bool sharded);
void nccl_release(nrs_sc_db *db);
}
