// sc_search_bin.cu -- exhaustive ScanContext search over {0, c} descriptors on sm_100a (K10b).
//
// In this pipeline every occupied ScanContext bin holds the same value: ORORA publishes its feature
// cloud with z = 1.0 (ORORA/src/odometry.cpp:305-316) and makeScancontext stores z + LIDAR_HEIGHT
// (SC/Scancontext.cpp:168,182) -- the 20x60 f64 descriptor is {0, 3.0}.  For such descriptors
// everything distanceBtnScanContext (Scancontext.cpp:116-148) computes is a function of bit masks:
//   column c           -> 20-bit occupancy mask m_c, count n_c = popc(m_c)
//   sector key         -> c * n_c / 20                               (Scancontext.cpp:223)
//   column norm        -> sqrt(n_c * c^2)                            (:78)
//   column dot product -> c_q * c_k * popc(mq & mk)                  (:77)
// and, when c has at most 24 significant bits (3.0 has 2; any float z + 2.0 has at most 25), every
// partial sum of those reductions is exact in f64, so the reference's values can be REPRODUCED
// bit-for-bit from the masks (certify_*_bin below) -- no 9600-byte descriptor is ever read.
//
// Storage: a 256-byte BIN RECORD per keyframe, written at append time by sc_pack_kernel beside the
// canonical descriptor: words 0..59 = column masks, 60..61 = 60-bit mask of empty columns, 62 = c as
// f32 (exact by the 24-bit condition), 63 = flags (bit 0: keyframe is binary).  A 10k-keyframe
// database is 2.56 MB of records instead of 96 MB of f64 (or 55 MB of f32 filter records).
//
// Kernel sc_search_bin_kernel, ONE launch for Q >= 1 queries.  Work is cut into tiles
// (query, chunk of keyframes); a 12-warp CTA (BT = 384, two CTAs per SM) prepares the tile's query in
// shared memory (masks, 1/sqrt(n) per column, the count vector as four byte-shifted copies of the
// byte-packed, doubled window) and its warps take one keyframe each:
//   fastAlignUsingVkey (:93-113): argmin_s |vq - shift(vk, s)| == argmax_s sum_j n_k[j] n_q[(j+s)%60];
//        lane l scores shifts l and l + 32 (the latter for l < 28) from ONE 15-word register window of
//        the query counts -- the two shifts read the same words, rotated by 8 -- with 2 x 15 DP4A
//        against the keyframe's byte-packed counts (broadcast from shared memory).  Integer, exact;
//        shifts that tie on the integer score are re-evaluated in the reference's f64 order (certify_shift_bin), because
//        rounding decides between them there.
//   distDirectSC x 7 shifts (:69-90,123-146): lane l owns key columns 2l, 2l+1; per shift slot and
//        column one AND + POPC + two multiplies by the tabulated 1/sqrt(n); the 7 partial sums are
//        reduced by REDUX.SUM on 2^-24 fixed point (NRS_SC_BIN_REDUX=0: a shuffle tree).  f32, error <= 12 * 2^-24 per distance (terms are <= 1, see bound C in
//        sc_kernels.cu; the same 2^-17 margins are used).
//   arg-min over keyframes: pruned exactly like the general kernel -- a keyframe whose lower bound
//        exceeds the smallest upper bound seen (CTA-wide in shared memory, per query in global memory)
//        is dropped, the rest go on the query's survivor list.
// The CTA that finishes the LAST tile of a query certifies that query's survivors in f64 from the
// masks (bit-identical to the reference's arithmetic), writes the (dist, idx, shift) record, and -- for
// a keyframe-sharded database -- stores it into every peer's HBM; the CTA that finishes the last
// query of the launch then merges the world's records (one wait per LAUNCH, not per query).
// A query that is not {0, c} still gets the right answer: its tiles evaluate every keyframe in the
// reference's f64 order from the canonical descriptor (slow path; host entry points that see the query
// route such queries to the general kernel instead).
//
// Bound: with 256 B per keyframe the search is no longer HBM-bound; it is issue-bound (DP4A + POPC)
// and, for a single query, launch/latency-bound.  Roofline figures are reported against both the
// canonical 9600 B/keyframe and the 256 B actually streamed.
#include <stdlib.h>
#include <algorithm>

#include "sc_common.cuh"

#ifndef NRS_SC_BIN_REDUX_DEFAULT
#define NRS_SC_BIN_REDUX_DEFAULT 1      // 1: REDUX.SUM fixed-point column sums (measured 5 % faster), 0: shuffle tree; NRS_SC_BIN_REDUX overrides
#endif

using namespace nrs_sc;

namespace nrs_sc {

constexpr int BT = 384;                      // threads per CTA (2 CTAs per SM: 85 registers per thread, no spills in the hot loop)
constexpr int BWARPS = BT / 32;
constexpr int WIN_STRIDE = 36;               // words per byte-shifted copy of the doubled count vector
constexpr int WORK_CAP = 192;                // certification work list (shared memory); longer lists are certified in place

struct BinWarp {
    unsigned int nkb[16];                    // key column counts, one byte each (bytes 60..63 zero)
    unsigned int krec[BIN_WORDS];            // certification: the keyframe's record
    double       vkd[S + 4];                 // certification: f64 sector key of the keyframe
    double       scr[S + 4];
    Survivor     first;
};

struct BinSmem {
    uint2         qcol[128];                 // (mask, bits of 1/sqrt(count) f32), entry e <-> column e % 60
    unsigned int  win[4][WIN_STRIDE];        // win[a] word w = counts of columns (4w + a + k) % 60, k = 0..3
    QCold         qc;                        // f64 sector key (and, for a non-binary query, norms)
    float         rtab[24];                  // 1 / sqrt(n), n = 0..20 (0 -> 0)
    float         rcnt[64];                  // 1 / cnt, cnt = 0..60 (0 -> 0)
    Survivor      work[WORK_CAP];            // certification: the entries the final bound could not rule out
    unsigned int  n_work;
    unsigned int  qmask[64];
    unsigned char nqb[64];
    unsigned long long qzero;
    unsigned long long gen_zero;
    double        cq;
    unsigned int  cls[4];                    // AND / OR of the bit patterns of the non-zero entries
    unsigned int  ub;
    int           q_binary, is_last, is_glast;
    double        red_d[BWARPS];
    int           red_pos[BWARPS], red_idx[BWARPS], red_s[BWARPS];
    BinWarp       w[BWARPS];
};

static __device__ __forceinline__ float u2f(unsigned int p)     // I2FP.F32.U32: one full-rate instruction on sm_100
{
    return __uint2float_rn(p);
}
static __device__ __forceinline__ bool c_ok(double c)            // see the header: <= 24 significant bits, sane range
{
    return (c > 0.0) && ((double)(float)c == c) && (c >= 7.888609052210118e-31) && (c <= 1.2676506002282294e30);
}
static __device__ __forceinline__ uint2 ldg_u2(const unsigned int *p)
{
    uint2 v;
    asm volatile("ld.global.nc.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p));
    return v;
}
// ---- query preparation (once per tile) ----------------------------------------------------------
static __device__ __forceinline__ void prep_query(BinSmem &sm, const double *__restrict__ query, const SearchState *stq)
{
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    __syncthreads();                               // the previous tile is done with the tables
    if (t < 4) sm.cls[t] = (t < 2) ? 0xffffffffu : 0u;
    if (t == 32) { sm.ub = __ldcg(&stq->gub); sm.gen_zero = 0ull; sm.qc.q64 = query; }
    __syncthreads();
    {
        unsigned int bits = 0u, alo = 0xffffffffu, ahi = 0xffffffffu, olo = 0u, ohi = 0u;
        const bool ld = t < 4 * S;
        const int c = t >> 2, k = t & 3;
        if (ld) {
            const double *p = query + c * R + 5 * k;
#pragma unroll
            for (int r = 0; r < 5; ++r) {
                const double v = p[r];
                if (!(v == 0.0)) {                 // NaN counts as an entry (and fails c_ok below)
                    bits |= 1u << (5 * k + r);
                    const unsigned long long b = (unsigned long long)__double_as_longlong(v);
                    alo &= (unsigned int)b; ahi &= (unsigned int)(b >> 32);
                    olo |= (unsigned int)b; ohi |= (unsigned int)(b >> 32);
                }
            }
        }
        alo = __reduce_and_sync(0xffffffffu, alo); ahi = __reduce_and_sync(0xffffffffu, ahi);
        olo = __reduce_or_sync(0xffffffffu, olo);  ohi = __reduce_or_sync(0xffffffffu, ohi);
        if (lane == 0) {
            atomicAnd(&sm.cls[0], alo); atomicAnd(&sm.cls[1], ahi);
            atomicOr(&sm.cls[2], olo);  atomicOr(&sm.cls[3], ohi);
        }
        bits |= __shfl_xor_sync(0xffffffffu, bits, 1);
        bits |= __shfl_xor_sync(0xffffffffu, bits, 2);
        if (ld && k == 0) { sm.qmask[c] = bits; sm.nqb[c] = (unsigned char)__popc(bits); }
        if (t >= 4 * S && t < 4 * S + 4) { sm.qmask[S + t - 4 * S] = 0u; sm.nqb[S + t - 4 * S] = 0; }
    }
    __syncthreads();
    const unsigned long long pa = sm.cls[0] | ((unsigned long long)sm.cls[1] << 32);
    const unsigned long long po = sm.cls[2] | ((unsigned long long)sm.cls[3] << 32);
    const bool none = (po == 0ull);
    const double cq = none ? 1.0 : __longlong_as_double((long long)po);
    const bool qbin = (none || pa == po) && c_ok(cq);
    if (t < 128) {
        const int cc = t % S;
        sm.qcol[t] = make_uint2(sm.qmask[cc], __float_as_uint(sm.rtab[sm.nqb[cc]]));
    } else if (t < 128 + 4 * WIN_STRIDE) {
        const int i = t - 128, a = i / WIN_STRIDE, w = i - a * WIN_STRIDE;
        unsigned int word = 0u;
#pragma unroll
        for (int k = 0; k < 4; ++k) word |= (unsigned int)sm.nqb[(4 * w + a + k) % S] << (8 * k);
        sm.win[a][w] = word;
    } else if (t < 128 + 4 * WIN_STRIDE + S) {
        const int cc = t - (128 + 4 * WIN_STRIDE);
        sm.qc.qv[cc] = ((double)sm.nqb[cc] * cq) / (double)R;          // == redux20(column) / 20, sums exact
    } else if (warp == BWARPS - 1) {
        const unsigned int z0 = __ballot_sync(0xffffffffu, sm.qmask[lane] == 0u);
        const unsigned int z1 = __ballot_sync(0xffffffffu, lane + 32 < S && sm.qmask[lane + 32] == 0u);
        if (lane == 0) {
            sm.qzero = (z0 | ((unsigned long long)z1 << 32)) & MASK60;
            sm.cq = cq;
            sm.q_binary = qbin ? 1 : 0;
        }
    }
    __syncthreads();
    if (!qbin) {
        // general-valued query: the reference's key / norm arithmetic in Eigen's order (Scancontext.cpp:223, :78)
        if (t < S) {
            // redux20's order, streamed (four running sums over i % 4, combined (s0 + s2) + (s1 + s3))
            const double *col = query + t * R;
            double a0 = col[0], a1 = col[1], b0 = col[2], b1 = col[3];
            double p0 = a0 * a0, p1 = a1 * a1, r0 = b0 * b0, r1 = b1 * b1;
#pragma unroll
            for (int i = 4; i < R; i += 4) {
                const double x0 = col[i], x1 = col[i + 1], x2 = col[i + 2], x3 = col[i + 3];
                a0 += x0; a1 += x1; b0 += x2; b1 += x3;
                p0 += x0 * x0; p1 += x1 * x1; r0 += x2 * x2; r1 += x3 * x3;
            }
            a0 += b0; a1 += b1;
            p0 += r0; p1 += r1;
            const double mean = (a0 + a1) / (double)R, nrm = sqrt(p0 + p1);
            sm.qc.qv[t] = mean;
            sm.qc.qn[t] = nrm;
            if (nrm == 0) atomicOr(&sm.gen_zero, 1ull << t);
        }
        __syncthreads();
        if (t == 0) sm.qzero = sm.gen_zero;
        __syncthreads();
    }
}

// ---- exact f64 re-evaluation from the masks ---------------------------------------------------------
// fastAlignUsingVkey over the candidate shifts (lo: shifts 0..31, hi: 32..59), ascending, strict '<'; vk[j] = (n_k[j] * c_k) / 20 is the
// keyframe's sector key exactly as sc_pack_kernel / the reference compute it (the column sum is exact).
static __device__ __noinline__ int certify_shift_bin(const double *__restrict__ qv, const unsigned int *nkb_w, double ck,
                                                     double *vkd, double *scr, unsigned int lo, unsigned int hi, int lane)
{
    const unsigned char *nkb = reinterpret_cast<const unsigned char *>(nkb_w);
    __syncwarp();
    for (int i = lane; i < S; i += 32) vkd[i] = ((double)nkb[i] * ck) / (double)R;
    double fa_n = SENTINEL;
    int s0 = 0;
    while (lo | hi) {                              // bit l of lo <=> shift l, of hi <=> shift 32 + l: ascending
        int s;
        if (lo) { s = __ffs(lo) - 1; lo &= lo - 1; }
        else { s = 31 + __ffs(hi); hi &= hi - 1; }
        __syncwarp();
        for (int i = lane; i < S; i += 32) {
            const double d = qv[i] - vkd[wrap60(i - s)];
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
        if (nrm < fa_n) { fa_n = nrm; s0 = s; }
    }
    __syncwarp();
    return s0;
}

// distDirectSC for the shift slots in candC (bit d <=> shift s0 + d - 3): dot = popc * (cq * ck), norms
// sqrt(n * c^2) -- all exact products, so each term equals the reference's dot / (|a| |b|) bit for bit; the
// column sum runs left to right like distDirectSC's loop.
static __device__ __noinline__ void certify_dist_bin(const uint2 *__restrict__ qcol, double cq, const unsigned int *__restrict__ krec,
                                                     double ck, unsigned long long qzero, unsigned long long kzero, double *scr,
                                                     int s0, unsigned int candC, int lane, double &dd, int &ds)
{
    dd = SENTINEL;
    ds = 0;
    const double pqk = cq * ck, cq2 = cq * cq, ck2 = ck * ck;
    while (candC) {
        const int d = __ffs(candC) - 1;
        candC &= candC - 1;
        const int s = wrap60(s0 + d - RADIUS);
        __syncwarp();
        const unsigned long long krot = s ? (((kzero << s) | (kzero >> (S - s))) & MASK60) : kzero;
        const unsigned long long eff = ~(qzero | krot) & MASK60;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int c = lane + 32 * h;
            if (c >= S || !((eff >> c) & 1ull)) continue;
            const int kc = wrap60(c - s);
            const unsigned int mq = qcol[c].x, mk = krec[kc];
            const double dot = (double)__popc(mq & mk) * pqk;
            const double nq = sqrt((double)__popc(mq) * cq2), nk = sqrt((double)__popc(mk) * ck2);
            const int pos = __popcll(eff & ((1ull << c) - 1ull));
            scr[pos] = dot / (nq * nk);
        }
        __syncwarp();
        double dist = SENTINEL;
        if (lane == 0) {
            const int cnt = __popcll(eff);
            double sum = 0.0;
            for (int i = 0; i < cnt; ++i) sum = sum + scr[i];
            dist = 1.0 - sum / (double)cnt;
        }
        dist = __shfl_sync(0xffffffffu, dist, 0);
        if (better(dist, s, dd, ds)) { dd = dist; ds = s; }
    }
    __syncwarp();
}

template <bool kRedux, bool kCand>
__global__ void __launch_bounds__(BT, 2)
sc_search_bin_kernel(const unsigned int *__restrict__ brec, const double *__restrict__ desc, const double *__restrict__ keys,
                     const double *__restrict__ queries, int Q, const int32_t *__restrict__ cand_base, int cand_stride,
                     int first, int n_items, int n_chunks, int per, int idx_scale, int idx_offset, Survivor *__restrict__ surv, size_t surv_stride,
                     SearchState *__restrict__ st, BlockResult *__restrict__ block_results, unsigned int *__restrict__ gticket,
                     nrs_sc_result *__restrict__ results, XchgView xv)
{
    __shared__ __align__(16) BinSmem sm;
    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
    BinWarp &ws = sm.w[warp];
    if (t < 21) sm.rtab[t] = t ? (float)(1.0 / sqrt((double)t)) : 0.f;

    if (t < 64) sm.rcnt[t] = t ? 1.0f / (float)t : 0.f;
    const int n_tiles = Q * n_chunks;
    const bool own = lane < S / 2;
    const int lb = own ? lane : 0;

    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int q = (Q == 1) ? 0 : tile / n_chunks, chunk = tile - q * n_chunks;
        const double *query = queries + (size_t)q * DESC;
        const int32_t *cand = kCand ? cand_base + (size_t)q * cand_stride : nullptr;         // this query's own candidate list (compiled out of the exhaustive search)
        SearchState *stq = st + q;
        Survivor *survq = surv + (size_t)q * surv_stride;
        const int lo = chunk * per, hi = min(n_items, lo + per);

        // this warp's first record is in flight while the CTA prepares the query
        uint2 cur = make_uint2(0u, 0u);
        if (lo + warp < hi) {
            const int kf = kCand ? cand[lo + warp] : first + lo + warp;
            cur = ldg_u2(brec + (size_t)kf * BIN_WORDS + 2 * lane);
        }
        prep_query(sm, query, stq);
        const bool qbin = sm.q_binary != 0;
        const unsigned long long qzero = sm.qzero;

        if (qbin) {
            // sliding window of the query's count vector for this lane's shift l: word i = counts of columns
            // (4i + l + k) % 60, k = 0..3.  The window of shift l + 32 is the same 15 words rotated by 8 (32 bytes),
            // so lane l scores the shifts l and l + 32 (< 60: lanes 0..27) from ONE register window.
            unsigned int W[15];
#pragma unroll
            for (int i = 0; i < 15; ++i) W[i] = sm.win[lane & 3][i + (lane >> 2)];
            const bool hasB = lane < S - 32;
            const unsigned int qz_any = (unsigned int)(qzero != 0ull);
            const float cqf = (float)sm.cq;
            float my_ub = INFINITY;                    // warp-uniform
            bool parked = false;                       // warp-uniform
            for (int item = lo + warp; item < hi; item += BWARPS) {
                const uint2 km = cur;
                if (item + BWARPS < hi) {
                    const int kf2 = kCand ? cand[item + BWARPS] : first + item + BWARPS;
                    cur = ldg_u2(brec + (size_t)kf2 * BIN_WORDS + 2 * lane);
                }
                const unsigned int kz_lo = __shfl_sync(0xffffffffu, km.x, 30), kz_hi = __shfl_sync(0xffffffffu, km.y, 30);
                const float ckf = __uint_as_float(__shfl_sync(0xffffffffu, km.x, 31));
                const unsigned int m0 = own ? km.x : 0u, m1 = own ? km.y : 0u;
                const int n0 = __popc(m0), n1 = __popc(m1);

                // ---- fastAlignUsingVkey: integer correlation of the count vectors, 60 shifts ----------------
                __syncwarp();
                reinterpret_cast<unsigned short *>(ws.nkb)[lane] = (unsigned short)(n0 | (n1 << 8));
                __syncwarp();
                unsigned int cA = 0u, cB = 0u;
                {
                    const uint4 *kb = reinterpret_cast<const uint4 *>(ws.nkb);
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const uint4 k4 = kb[i];
                        const unsigned int kk[4] = {k4.x, k4.y, k4.z, k4.w};
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const int w = 4 * i + j;
                            if (w >= 15) continue;                      // bytes 60..63 of the counts are zero
                            cA = __dp4a(kk[j], W[w], cA);
                            cB = __dp4a(kk[j], W[(w + 8) % 15], cB);
                        }
                    }
                }
                if (!hasB) cB = 0u;
                const unsigned int cmx = __reduce_max_sync(0xffffffffu, max(cA, cB));
                // the integer score orders the f64 norms only while the two scales are comparable
                const bool scale_ok = (ckf <= cqf * 1024.0f) && (cqf <= ckf * 1024.0f);
                const unsigned int candLo = __ballot_sync(0xffffffffu, !scale_ok || cA == cmx);
                const unsigned int candHi = __ballot_sync(0xffffffffu, hasB && (!scale_ok || cB == cmx));
                int s0;
                if (__popc(candLo) + __popc(candHi) == 1) s0 = candLo ? __ffs(candLo) - 1 : 31 + __ffs(candHi);
                else s0 = certify_shift_bin(sm.qc.qv, ws.nkb, (double)ckf, ws.vkd, ws.scr, candLo, candHi, lane);

                // ---- 7 shifts x 60 columns of cosine similarity from the masks --------------------------------
                float part[NSHIFT];
                {
                    const int e0 = 2 * lb + wrap60(s0 - RADIUS);        // <= 58 + 59; + 7 < 128
                    const float rk0 = sm.rtab[n0], rk1 = sm.rtab[n1];
                    // (eight 8-byte loads at a 16-byte lane stride: twice the minimal shared-memory wavefronts, yet measured
                    // 2 % faster than four aligned 16-byte loads from two staggered copies of the list -- the kernel is
                    // issue-bound, not LSU-bound, and the wide loads lengthen the dependent chain)
                    uint2 qe[NSHIFT + 1];
#pragma unroll
                    for (int i = 0; i < NSHIFT + 1; ++i) qe[i] = sm.qcol[e0 + i];
#pragma unroll
                    for (int d = 0; d < NSHIFT; ++d) {
                        const float a = u2f(__popc(m0 & qe[d].x)) * __uint_as_float(qe[d].y);
                        const float b = u2f(__popc(m1 & qe[d + 1].x)) * __uint_as_float(qe[d + 1].y);
                        part[d] = fmaf(a, rk0, b * rk1);
                    }
                }
                float dC = INFINITY;
                int cntC = 0;
                {
                    float v;                                       // lane d (< 7): column sum of shift slot d
                    if (kRedux) {
                        // fixed point, 2^-24: a lane's partial is < 2 (two cosines), 30 lanes sum below 2^30; REDUX.SUM leaves
                        // every slot's total in a uniform register (error <= 2^-25 per non-zero lane, cf. bound C)
                        int tot[NSHIFT];
#pragma unroll
                        for (int d = 0; d < NSHIFT; ++d) tot[d] = __reduce_add_sync(0xffffffffu, __float2int_rn(part[d] * 16777216.0f));
                        int mine = tot[0];
#pragma unroll
                        for (int d = 1; d < NSHIFT; ++d) mine = (lane == d) ? tot[d] : mine;
                        v = (float)mine * 5.9604644775390625e-08f;
                    } else {
                        float v4[4], v2[2], v1;
                        const bool up16 = lane & 16, up8 = lane & 8, up4 = lane & 4;
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const float hi_ = (i == 3) ? 0.f : part[4 + i];
                            const float send = up16 ? part[i] : hi_, keep = up16 ? hi_ : part[i];
                            v4[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
                        }
#pragma unroll
                        for (int i = 0; i < 2; ++i) {
                            const float send = up8 ? v4[i] : v4[2 + i], keep = up8 ? v4[2 + i] : v4[i];
                            v2[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
                        }
                        {
                            const float send = up4 ? v2[0] : v2[1], keep = up4 ? v2[1] : v2[0];
                            v1 = keep + __shfl_xor_sync(0xffffffffu, send, 4);
                        }
                        v1 += __shfl_xor_sync(0xffffffffu, v1, 2);
                        v1 += __shfl_xor_sync(0xffffffffu, v1, 1);
                        const int srcl = (lane < 8) ? (((lane >> 2) & 1) << 4 | ((lane >> 1) & 1) << 3 | (lane & 1) << 2) : 0;
                        v = __shfl_sync(0xffffffffu, v1, srcl);
                    }
                    if ((kz_lo | kz_hi | qz_any) == 0u) {          // no empty column on either side (warp-uniform): 60 effective columns
                        if (lane < NSHIFT) { cntC = S; dC = 1.0f - v * 0.016666668f; }
                    } else if (lane < NSHIFT) {
                        const int s = wrap60(s0 + lane - RADIUS);
                        const unsigned long long kzero = (((unsigned long long)kz_hi << 32) | kz_lo) & MASK60;
                        const unsigned long long krot = s ? (((kzero << s) | (kzero >> (S - s))) & MASK60) : kzero;
                        cntC = S - __popcll((qzero | krot) & MASK60);
                        dC = cntC ? 1.0f - v * sm.rcnt[cntC] : INFINITY;     // the reciprocal's rounding is inside bound C
                    }
                }
                const float mC = warp_min_f(dC);
                const unsigned int candC = __ballot_sync(0xffffffffu, lane < NSHIFT && cntC > 0 && dC <= mC + 2.0f * EPS_C);

                // every exact distance of this keyframe is >= mC - e_C and its best one is <= mC + e_C: it survives unless
                // that lower bound already exceeds an upper bound seen by this CTA; all of it warp-uniform, lane 0 acts
                const float bound = fminf(my_ub, unord_f32(*reinterpret_cast<volatile unsigned int *>(&sm.ub)));
                const bool push = candC && (mC - EPS_C <= bound);
                const bool improve = candC && (mC + EPS_C < bound);
                if (improve) my_ub = mC + EPS_C;
                if (push || improve || !parked) {
                    if (lane == 0) {
                        Survivor sv;
                        sv.item = item;
                        sv.code = push ? (s0 | (int)(candC << 8) | (1 << 16)) : (1 << 16);
                        sv.lo = push ? mC : INFINITY;                          // INFINITY: "ruled out"
                        sv.pad = __float_as_int(ckf);
                        if (improve) {
                            atomicMin(&sm.ub, ord_f32(my_ub));
                            atomicMin(&stq->gub, ord_f32(my_ub));
                        }
                        // the first keyframe of a warp may meet no bound yet: park it and judge it after the loop
                        if (!parked) ws.first = sv;
                        else if (push) survq[atomicAdd(&stq->n_surv, 1u)] = sv;
                    }
                    parked = true;
                }
            }
            // the parked first keyframes are judged against the CTA's final bound (one entry per CTA survives, typically)
            __syncthreads();
            if (lane == 0 && parked) {
                const Survivor sv = ws.first;
                const float bound = fminf(my_ub, unord_f32(*reinterpret_cast<volatile unsigned int *>(&sm.ub)));
                if (sv.lo - EPS_C <= bound) survq[atomicAdd(&stq->n_surv, 1u)] = sv;
            }
        } else {
            // ---- general-valued query: every keyframe in the reference's f64 order from the canonical data ----
            double bd = SENTINEL;
            int bp = 0x7fffffff, bi = 0, bs = 0;
            for (int item = lo + warp; item < hi; item += BWARPS) {
                const int kf = kCand ? cand[item] : first + item;
                const uint2 kzw = *reinterpret_cast<const uint2 *>(brec + (size_t)kf * BIN_WORDS + 60);
                const unsigned long long kzero = (((unsigned long long)kzw.y << 32) | kzw.x) & MASK60;
                const int s0 = certify_shift(&sm.qc, ws.scr, keys + (size_t)kf * KEYS_D, 0x3fffffffu, 0x3fffffffu, lane);
                double dd;
                int ds;
                certify_dist(&sm.qc, qzero, ws.scr, desc + (size_t)kf * DESC, keys + (size_t)kf * KEYS_D + S, kzero, s0, 0x7fu,
                             lane, dd, ds);
                if (dd < bd) { bd = dd; bp = item; bi = kf; bs = ds; }
            }
            if (lane == 0) { sm.red_d[warp] = bd; sm.red_pos[warp] = bp; sm.red_idx[warp] = bi; sm.red_s[warp] = bs; }
            __syncthreads();
            if (t == 0) {
                BlockResult br;
                br.dist = sm.red_d[0]; br.pos = sm.red_pos[0]; br.idx = sm.red_idx[0]; br.shift = sm.red_s[0]; br.pad = 0;
                for (int w = 1; w < BWARPS; ++w)
                    if (better(sm.red_d[w], sm.red_pos[w], br.dist, br.pos)) {
                        br.dist = sm.red_d[w]; br.pos = sm.red_pos[w]; br.idx = sm.red_idx[w]; br.shift = sm.red_s[w];
                    }
                block_results[tile] = br;
            }
        }

        // ---- last tile of this query done?  then this CTA certifies the query ------------------------------------
        __syncthreads();
        if (t == 0) {
            __threadfence();
            const unsigned int tk = atomicAdd(&stq->ticket, 1u);
            sm.is_last = (tk == (unsigned int)n_chunks - 1u);
            sm.is_glast = 0;
        }
        __syncthreads();
        if (!sm.is_last) continue;
        __threadfence();

        double d = SENTINEL;
        int p = 0x7fffffff, ix = 0, sh = 0;
        if (qbin) {
            const unsigned int n_list = *reinterpret_cast<volatile unsigned int *>(&stq->n_surv);
            const float bound = unord_f32(*reinterpret_cast<volatile unsigned int *>(&stq->gub));
            const double cq = sm.cq;
            // Phase 1, all warps in parallel: the entries the FINAL bound cannot rule out (typically a handful of the few
            // hundred listed) are compacted into a shared work list.  Phase 2: work items = (entry, one of its candidate
            // shifts), dealt round-robin to the warps.  A list longer than the work list's capacity (adversarial inputs:
            // hundreds of near-ties) is certified in place, every warp taking the entries it filtered itself.
            // Order of the reference's scans: (dist, keyframe position, shift) lexicographic.
            auto certify_one = [&](int it, int cd, float ckf, unsigned int bits) {
                const int kf = kCand ? cand[it] : first + it;
                __syncwarp();
                *reinterpret_cast<uint2 *>(&ws.krec[2 * lane]) = __ldcg(reinterpret_cast<const uint2 *>(brec + (size_t)kf * BIN_WORDS + 2 * lane));
                __syncwarp();
                const unsigned long long kzero = (((unsigned long long)ws.krec[61] << 32) | ws.krec[60]) & MASK60;
                double dd;
                int ds;
                certify_dist_bin(sm.qcol, cq, ws.krec, (double)ckf, qzero, kzero, ws.scr, cd & 0xff, bits, lane, dd, ds);
                if (dd < d || (dd == d && (it < p || (it == p && ds < sh)))) { d = dd; p = it; ix = kf; sh = ds; }
            };
            if (t == 0) sm.n_work = 0u;
            __syncthreads();
            const bool in_place = n_list > 8u * WORK_CAP;          // even the filtered list may not fit: skip the work list
            for (unsigned int base = warp * 32; base < n_list; base += BWARPS * 32) {
                const unsigned int i = base + lane;
                const int4 e = (i < n_list) ? __ldcg(reinterpret_cast<const int4 *>(survq + i)) : make_int4(0, 0, 0x7f800000, 0);
                const bool need = i < n_list && !(__int_as_float(e.z) - EPS_C > bound);
                unsigned int slot = 0xffffffffu;
                if (need && !in_place) slot = atomicAdd(&sm.n_work, 1u);
                if (need && slot < WORK_CAP) { Survivor sv; sv.item = e.x; sv.code = e.y; sv.lo = __int_as_float(e.z); sv.pad = e.w; sm.work[slot] = sv; }
                unsigned int todo = __ballot_sync(0xffffffffu, need && slot >= WORK_CAP);      // overflow / in place: certify here
                while (todo) {
                    const int src = __ffs(todo) - 1;
                    todo &= todo - 1;
                    const int it = __shfl_sync(0xffffffffu, e.x, src), cd = __shfl_sync(0xffffffffu, e.y, src);
                    const float ckf = __int_as_float(__shfl_sync(0xffffffffu, e.w, src));
                    certify_one(it, cd, ckf, (unsigned int)(cd >> 8) & 0x7fu);
                }
            }
            __syncthreads();
            {
                const int nw = (int)min(sm.n_work, (unsigned int)WORK_CAP);
                int wcount = 0;
                for (int w = 0; w < nw; ++w) {
                    const Survivor sv = sm.work[w];
                    unsigned int bits = (unsigned int)(sv.code >> 8) & 0x7fu;
                    while (bits) {
                        const unsigned int one = bits & (0u - bits);
                        bits &= bits - 1;
                        if ((wcount++) % BWARPS == warp) certify_one(sv.item, sv.code, __int_as_float(sv.pad), one);
                    }
                }
            }
        } else if (warp == 0) {
            for (int b = lane; b < n_chunks; b += 32) {
                const volatile BlockResult *br = block_results + (size_t)q * n_chunks + b;
                const double bd = br->dist;
                const int bp = br->pos;
                if (better(bd, bp, d, p)) { d = bd; p = bp; ix = br->idx; sh = br->shift; }
            }
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {            // lanes hold different tiles: combine them
                const double od = __shfl_xor_sync(0xffffffffu, d, off);
                const int op = __shfl_xor_sync(0xffffffffu, p, off);
                const int oi = __shfl_xor_sync(0xffffffffu, ix, off);
                const int os = __shfl_xor_sync(0xffffffffu, sh, off);
                if (better(od, op, d, p)) { d = od; p = op; ix = oi; sh = os; }
            }
        }
        if (lane == 0) { sm.red_d[warp] = d; sm.red_pos[warp] = p; sm.red_idx[warp] = ix; sm.red_s[warp] = sh; }
        __syncthreads();
        if (warp == 0) {
            d = SENTINEL; p = 0x7fffffff; ix = 0; sh = 0;
            if (lane < BWARPS) { d = sm.red_d[lane]; p = sm.red_pos[lane]; ix = sm.red_idx[lane]; sh = sm.red_s[lane]; }
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
                const double od = __shfl_xor_sync(0xffffffffu, d, off);
                const int op = __shfl_xor_sync(0xffffffffu, p, off);
                const int oi = __shfl_xor_sync(0xffffffffu, ix, off);
                const int os = __shfl_xor_sync(0xffffffffu, sh, off);
                if (od < d || (od == d && (op < p || (op == p && os < sh)))) { d = od; p = op; ix = oi; sh = os; }
            }
            ix = (d < SENTINEL) ? ix * idx_scale + idx_offset : 0;
            if (!(d < SENTINEL)) sh = 0;
            if (xv.world > 1) {
                // this rank's record of query q into every rank's slot array (own included), payload then flag
                if (lane < xv.world) {
                    XchgSlot *dst = xv.peers[lane] + ((size_t)(xv.epoch & 1u) * XCHG_QCAP + q) * xv.world + xv.rank;
                    dst->dist = d; dst->idx = ix; dst->shift = sh;
                    st_release_sys(&dst->epoch, xv.epoch);
                }
            } else if (lane == 0) {
                nrs_sc_result res;
                res.dist = d; res.idx = ix; res.shift = sh;
                results[q] = res;
            }
            if (lane == 0) {
                stq->n_surv = 0u;                       // leave the query's state ready for the next launch
                stq->gub = 0xff800000u;                 // ord_f32(+inf)
                stq->ticket = 0u;
                __threadfence();
                const unsigned int g = atomicAdd(gticket, 1u);
                if (g == (unsigned int)Q - 1u) { *gticket = 0u; sm.is_glast = 1; }
            }
        }
        __syncthreads();
        if (sm.is_glast && xv.world > 1) {
            // ---- the launch's last query is certified: merge the world's records of every query ------------------
            for (int qq = t; qq < Q; qq += BT) {
                double md = SENTINEL;
                int mi = 0, ms = 0;
                bool ok = true, found = false;
                for (int r = 0; r < xv.world; ++r) {
                    const XchgSlot *src = xv.peers[xv.rank] + ((size_t)(xv.epoch & 1u) * XCHG_QCAP + qq) * xv.world + r;
                    const long long t0 = clock64();
                    while (ld_acquire_sys(&src->epoch) != xv.epoch)
                        if (clock64() - t0 > xv.timeout_cycles) { ok = false; break; }
                    if (!ok) break;
                    const double rd = src->dist;
                    const int ri = src->idx, rs = src->shift;
                    if (rd < SENTINEL && (!found || rd < md || (rd == md && ri < mi))) { md = rd; mi = ri; ms = rs; found = true; }
                }
                nrs_sc_result res;
                res.dist = md; res.idx = mi; res.shift = ms;
                if (!ok) { res.dist = SENTINEL; res.idx = -2; res.shift = 0; }    // a peer did not answer: reported, never silent
                results[qq] = res;
            }
        }
    }
}

}  // namespace nrs_sc

// -------------------------------------------------------------------------------------------------
// host side: launch on the context stream (no synchronisation)
// -------------------------------------------------------------------------------------------------
int nrs_sc::bin_launch(nrs_sc_db *db, const double *d_queries, int Q, const int32_t *d_cand, int first, int n_items,
                       int idx_scale, int idx_offset, nrs_sc_result *d_results, XchgView xv, int cand_stride)
{
    nrs_ctx *ctx = db->ctx;
    if (Q < 1 || Q > XCHG_QCAP) { nrs::set_error("sc batch: 1 <= Q <= %d", XCHG_QCAP); return NRS_EINVAL; }
    const int max_ctas = 2 * ctx->sm_count;
    int n_chunks;
    if (Q == 1) n_chunks = std::max(1, std::min(max_ctas, (n_items + BWARPS - 1) / BWARPS));
    else {
        // tiles are dealt round-robin to a persistent grid: the launch lasts ceil(tiles / grid) rounds, so among the
        // chunk counts from "~4 tiles per CTA" up to twice that, take the one that wastes the least of the last round
        // (256 queries on 296 CTAs: 8 chunks = 2048 tiles = 6.92 rounds instead of 5 chunks = 4.32 rounds paid as 5)
        const int cmax = std::max(1, (n_items + 8 * BWARPS - 1) / (8 * BWARPS));
        const int c0 = std::max(1, std::min((4 * max_ctas + Q - 1) / Q, cmax));
        double best = -1.0;
        n_chunks = c0;
        for (int c = c0; c <= std::min(cmax, 2 * c0 + 2); ++c) {
            const long long tiles = (long long)Q * c, g = std::min<long long>(max_ctas, tiles);
            const double eff = (double)tiles / (double)(((tiles + g - 1) / g) * g);
            if (eff > best + 1e-9) { best = eff; n_chunks = c; }
        }
    }
    const size_t stride = (size_t)n_items + 64;
    int rc = db->bin_surv.reserve((size_t)Q * stride * sizeof(Survivor));
    if (rc) return rc;
    rc = db->bin_block.reserve((size_t)Q * n_chunks * sizeof(BlockResult));
    if (rc) return rc;
    const int grid = std::min(max_ctas, Q * n_chunks);
    const int per = (n_items + n_chunks - 1) / n_chunks;
    static const int use_redux = []() { const char *e = getenv("NRS_SC_BIN_REDUX"); return e ? atoi(e) : NRS_SC_BIN_REDUX_DEFAULT; }();
#define NRS_BIN_LAUNCH(R_, C_)                                                                                               \
    sc_search_bin_kernel<R_, C_><<<grid, BT, 0, ctx->stream>>>(db->d_brec, db->d_desc, db->d_keys, d_queries, Q, d_cand, cand_stride, \
                                                               first, n_items, n_chunks, per, idx_scale, idx_offset,          \
                                                               db->bin_surv.as<Survivor>(), stride, db->d_bstate,              \
                                                               db->bin_block.as<BlockResult>(), db->d_gticket, d_results, xv)
    if (use_redux) { if (d_cand) NRS_BIN_LAUNCH(true, true); else NRS_BIN_LAUNCH(true, false); }
    else           { if (d_cand) NRS_BIN_LAUNCH(false, true); else NRS_BIN_LAUNCH(false, false); }
#undef NRS_BIN_LAUNCH
    NRS_LAUNCHED();
    return NRS_OK;
}
