// common.cuh -- context, error plumbing and small device helpers shared by every kernel file.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <atomic>
#include <mutex>
#include <string>
#include <utility>
#include <vector>

#include "../../include/nrs_b200.h"

namespace nrs {

void set_error(const char *fmt, ...);
extern std::atomic<uint64_t> g_launches;

#define NRS_CUDA(expr)                                                                      \
    do {                                                                                    \
        cudaError_t e__ = (expr);                                                           \
        if (e__ != cudaSuccess) {                                                           \
            ::nrs::set_error("%s:%d %s -> %s", __FILE__, __LINE__, #expr,                   \
                             cudaGetErrorString(e__));                                      \
            return (e__ == cudaErrorMemoryAllocation) ? NRS_ENOMEM                          \
                   : (e__ == cudaErrorNoDevice || e__ == cudaErrorInsufficientDriver ||     \
                      e__ == cudaErrorNoKernelImageForDevice)                               \
                       ? NRS_ENODEVICE                                                      \
                       : NRS_ECUDA;                                                         \
        }                                                                                   \
    } while (0)

#define NRS_CHECK_ARG(cond, msg)                                                            \
    do {                                                                                    \
        if (!(cond)) {                                                                      \
            ::nrs::set_error("%s: %s", __func__, msg);                                      \
            return NRS_EINVAL;                                                              \
        }                                                                                   \
    } while (0)

// count one launch of one of OUR kernels and check it was accepted
#define NRS_LAUNCHED()                                                                      \
    do {                                                                                    \
        ::nrs::g_launches.fetch_add(1, std::memory_order_relaxed);                          \
        NRS_CUDA(cudaGetLastError());                                                       \
    } while (0)

// A grow-only device buffer (stream-ordered frees are not needed: buffers live with the ctx).
struct DevBuf {
    void  *ptr = nullptr;
    size_t cap = 0;
    int reserve(size_t bytes);
    void release();
    template <class T> T *as() { return reinterpret_cast<T *>(ptr); }
};

// Pinned host staging buffer.
struct PinBuf {
    void  *ptr = nullptr;
    size_t cap = 0;
    int reserve(size_t bytes);
    void release();
    template <class T> T *as() { return reinterpret_cast<T *>(ptr); }
};

// solver entry for callers that already hold ctx->mu (the fused pair call): points may live on the device
int orora_solve_locked(nrs_ctx *ctx, const double *prev_2xN, const double *curr_2xN, bool points_on_device, int N,
                       const nrs_solver_params *p, double vel_x, double T_out[16], int32_t *clique_out,
                       int32_t *final_inliers_out, nrs_solver_info *info_out);
// the same solve enqueued behind the matcher with the match count still on the device (d_n); no synchronisation.
// orora_pair_finish reads the result after the caller's synchronisation, given the count it read back.
int orora_pair_enqueue(nrs_ctx *ctx, const double *d_prev, const double *d_curr, const int *d_n,
                       const nrs_solver_params *p, double vel_x);
int orora_pair_finish(nrs_ctx *ctx, const double *d_prev, const double *d_curr, int N, const nrs_solver_params *p,
                      double vel_x, double T_out[16], nrs_solver_info *info_out);

}  // namespace nrs

struct nrs_ctx {
    int          device   = -1;
    int          sm_count = 0;
    cudaStream_t stream   = nullptr;
    std::mutex   mu;                   // serialises host-side use of the scratch below
    nrs::DevBuf  scratch[8];           // per-subsystem device scratch (indexed by users)
    nrs::PinBuf  pin[4];               // pinned staging
    void        *frontend = nullptr;   // lazily created front-end static state (maps, tables)
    std::vector<std::pair<void *, size_t>> frame_pool;   // recycled nrs_frame device blocks (cudaMalloc/cudaFree cost ms)
    void       (*frontend_free)(void *) = nullptr;
};

// ---------------------------------------------------------------------------------------
// device helpers
// ---------------------------------------------------------------------------------------
#ifdef __CUDACC__
namespace nrs {

__device__ __forceinline__ uint32_t smem_u32(const void *p)
{
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// ---- mbarrier + 1-D TMA bulk copy (cp.async.bulk, SASS: UBLKCP) ------------------------
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init()
{
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async()
{
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// global -> shared bulk copy of `bytes` (multiple of 16, both addresses 16-byte aligned),
// completion signalled on `bar` by transaction bytes.
__device__ __forceinline__ void tma_load_1d(void *smem_dst, const void *gmem_src, uint32_t bytes,
                                            uint64_t *bar)
{
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
        ::"r"(smem_u32(smem_dst)),
        "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}

}  // namespace nrs
#endif
