// ctx.cu -- library / context entry points of include/nrs_b200.h.
#include <stdarg.h>

#include "common.cuh"

namespace nrs {

static thread_local char tl_error[1024] = "";
std::atomic<uint64_t> g_launches{0};

void set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(tl_error, sizeof(tl_error), fmt, ap);
    va_end(ap);
}

int DevBuf::reserve(size_t bytes)
{
    if (bytes <= cap) return NRS_OK;
    size_t want = cap ? cap : 4096;
    while (want < bytes) want *= 2;
    void *np = nullptr;
    NRS_CUDA(cudaMalloc(&np, want));
    if (ptr) cudaFree(ptr);
    ptr = np;
    cap = want;
    return NRS_OK;
}
void DevBuf::release()
{
    if (ptr) cudaFree(ptr);
    ptr = nullptr;
    cap = 0;
}
int PinBuf::reserve(size_t bytes)
{
    if (bytes <= cap) return NRS_OK;
    size_t want = cap ? cap : 4096;
    while (want < bytes) want *= 2;
    void *np = nullptr;
    NRS_CUDA(cudaMallocHost(&np, want));
    if (ptr) cudaFreeHost(ptr);
    ptr = np;
    cap = want;
    return NRS_OK;
}
void PinBuf::release()
{
    if (ptr) cudaFreeHost(ptr);
    ptr = nullptr;
    cap = 0;
}

}  // namespace nrs

extern "C" {

int nrs_version(void) { return 100; }  // 0.1.0

const char *nrs_last_error(void) { return nrs::tl_error; }

uint64_t nrs_kernel_launches(void) { return nrs::g_launches.load(); }

int nrs_ctx_create(int device, nrs_ctx **out)
{
    NRS_CHECK_ARG(out != nullptr, "out is NULL");
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        nrs::set_error("no CUDA device visible (%s); this library has no CPU fallback",
                       e != cudaSuccess ? cudaGetErrorString(e) : "device count 0");
        return NRS_ENODEVICE;
    }
    NRS_CHECK_ARG(device >= 0 && device < n, "device index out of range");
    cudaDeviceProp prop;
    NRS_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) {
        nrs::set_error("device %d is sm_%d%d; this library ships sm_100a code only", device,
                       prop.major, prop.minor);
        return NRS_ENODEVICE;
    }
    NRS_CUDA(cudaSetDevice(device));
    nrs_ctx *c = new nrs_ctx();
    c->device = device;
    c->sm_count = prop.multiProcessorCount;
    cudaError_t se = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
    if (se != cudaSuccess) {
        delete c;
        nrs::set_error("cudaStreamCreate: %s", cudaGetErrorString(se));
        return NRS_ECUDA;
    }
    *out = c;
    return NRS_OK;
}

int nrs_ctx_destroy(nrs_ctx *ctx)
{
    if (!ctx) return NRS_OK;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    if (ctx->frontend && ctx->frontend_free) ctx->frontend_free(ctx->frontend);
    for (auto &fb : ctx->frame_pool) cudaFree(fb.first);
    ctx->frame_pool.clear();
    for (auto &b : ctx->scratch) b.release();
    for (auto &b : ctx->pin) b.release();
    cudaStreamDestroy(ctx->stream);
    delete ctx;
    return NRS_OK;
}

int nrs_ctx_sync(nrs_ctx *ctx)
{
    NRS_CHECK_ARG(ctx != nullptr, "ctx is NULL");
    NRS_CUDA(cudaStreamSynchronize(ctx->stream));
    return NRS_OK;
}

void *nrs_ctx_stream(nrs_ctx *ctx) { return ctx ? (void *)ctx->stream : nullptr; }
int nrs_ctx_device(nrs_ctx *ctx) { return ctx ? ctx->device : -1; }
int nrs_ctx_sm_count(nrs_ctx *ctx) { return ctx ? ctx->sm_count : 0; }

}  // extern "C"
