// hl_comm.cu — the one collective of the path: samples-per-pixel sharding keeps a per-GPU SUM image
// (HL_ACCUM_SUM) and combines the images once per finished picture.  The reference has no multi-GPU path; the blend
// it runs on one GPU is path_trace_rgen.glsl:219-247, and sum / count is that running mean up to fp32 rounding.
//
// Two transports behind the same entry points:
//   * NCCL (one process per GPU under torchrun, or all GPUs of one process): ncclAllReduce / ncclReduce of W*H*4 floats
//     on the context's stream.  libnccl.so.2 is opened with dlopen at the first hl_comm_* call — the library has no
//     link-time dependency on it, and inside a torch process the already-mapped copy is the one that answers.
//   * peer memory (all contexts in ONE process: helios_headless --gpus N): k_peer_reduce — every GPU sums its 1/n slice
//     of the image over all n accumulation images with plain loads over NVLink (cudaDeviceEnablePeerAccess), applies
//     1/count + exposure + tone map + gamma and stores the RGBA8 pixels (and optionally the fp32 sum) straight into the
//     ROOT GPU's image: reduce-scatter + tone map + gather in one kernel per GPU, ordered by CUDA events only.  The sum is
//     taken in rank order, so the result does not depend on timing or topology.
#include "hl_film.h"
#include "hl_internal.h"
#include <dlfcn.h>
#include <cstring>
#include <mutex>

namespace hl
{
// ---- NCCL through dlopen ----------------------------------------------------------------------------------------
// The handful of declarations below are NCCL's stable public ABI (nccl.h: ncclUniqueId is 128 opaque bytes,
// ncclFloat32 = 7, ncclSum = 0, ncclSuccess = 0).
typedef struct ncclComm* ncclComm_t;
struct ncclUniqueId
{
    char internal[128];
};
static_assert(sizeof(ncclUniqueId) == HL_COMM_ID_BYTES, "ncclUniqueId is 128 bytes");
struct NcclApi
{
    void* handle = nullptr;
    int (*GetVersion)(int*)                                                                         = nullptr;
    int (*GetUniqueId)(ncclUniqueId*)                                                               = nullptr;
    int (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int)                                        = nullptr;
    int (*CommInitAll)(ncclComm_t*, int, const int*)                                                = nullptr;
    int (*CommDestroy)(ncclComm_t)                                                                  = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t)                = nullptr;
    int (*Reduce)(const void*, void*, size_t, int, int, int, ncclComm_t, cudaStream_t)              = nullptr;
    int (*GroupStart)()                                                                             = nullptr;
    int (*GroupEnd)()                                                                               = nullptr;
    const char* (*GetErrorString)(int)                                                              = nullptr;
    std::string error;
};
static NcclApi& nccl()
{
    static NcclApi        api;
    static std::once_flag once;
    std::call_once(once, [] {
        const char* names[] = { "libnccl.so.2", "libnccl.so" };
        for (const char* n : names)
            if ((api.handle = dlopen(n, RTLD_NOW | RTLD_LOCAL))) break;
        if (!api.handle)
        {
            api.error = std::string("libnccl.so.2 cannot be loaded: ") + (dlerror() ? dlerror() : "?");
            return;
        }
        auto sym = [&](const char* n) -> void* {
            void* p = dlsym(api.handle, n);
            if (!p && api.error.empty()) api.error = std::string("libnccl.so.2 lacks ") + n;
            return p;
        };
        *(void**)&api.GetVersion     = sym("ncclGetVersion");
        *(void**)&api.GetUniqueId    = sym("ncclGetUniqueId");
        *(void**)&api.CommInitRank   = sym("ncclCommInitRank");
        *(void**)&api.CommInitAll    = sym("ncclCommInitAll");
        *(void**)&api.CommDestroy    = sym("ncclCommDestroy");
        *(void**)&api.AllReduce      = sym("ncclAllReduce");
        *(void**)&api.Reduce         = sym("ncclReduce");
        *(void**)&api.GroupStart     = sym("ncclGroupStart");
        *(void**)&api.GroupEnd       = sym("ncclGroupEnd");
        *(void**)&api.GetErrorString = sym("ncclGetErrorString");
    });
    return api;
}
static void nccl_require()
{
    NcclApi& a = nccl();
    if (!a.error.empty()) throw CudaError(HL_ERR_STATE, a.error);
}
#define HL_NCCL(call)                                                                                                               \
    do                                                                                                                              \
    {                                                                                                                               \
        const int r_ = (call);                                                                                                      \
        if (r_ != 0) throw hl::CudaError(HL_ERR_CUDA, std::string(#call) + ": " + hl::nccl().GetErrorString(r_) + " (" + __FILE__ + ":" + std::to_string(__LINE__) + ")"); \
    } while (0)
enum
{
    kNcclFloat32 = 7,
    kNcclSum     = 0
};

// ---- the peer-memory kernel -------------------------------------------------------------------------------------
#define HL_MAX_COMM_RANKS 16
struct PeerImages
{
    const float4* accum[HL_MAX_COMM_RANKS];
    int           n;
};
// pixels [begin, end) of the accumulation images (row 0 = bottom, as path_trace_rgen.glsl:219-247 stores them): rank-ordered
// sum over the n images, then optionally the fp32 sum into dst_accum and / or tone_map.frag:35-51 of sum * scale into
// dst_rgba8 (row 0 = top, the reference's negative-height viewport, renderer.cpp:369-428).  dst_* may live on another GPU.
__global__ void __launch_bounds__(256) k_peer_reduce(PeerImages src, float4* dst_accum, uint32_t* dst_rgba8, uint32_t begin, uint32_t end, uint32_t W, uint32_t H, float exposure, int op,
                                                     float scale)
{
    for (uint32_t i = begin + blockIdx.x * blockDim.x + threadIdx.x; i < end; i += gridDim.x * blockDim.x)
    {
        float4 a = src.accum[0][i];
        for (int r = 1; r < src.n; r++)
        {
            const float4 b = src.accum[r][i];
            a.x += b.x, a.y += b.y, a.z += b.z;
        }
        a.w = 1.0f;
        if (dst_accum) dst_accum[i] = a;
        if (dst_rgba8)
        {
            const uint32_t py = i / W, px = i - py * W;
            dst_rgba8[(size_t)(H - 1 - py) * W + px] = tone_map_rgba8(mk3(a.x * scale, a.y * scale, a.z * scale), exposure, op);
        }
    }
}

static bool peers_reachable(hl_context_t* const* ctxs, int n)
{
    for (int i = 0; i < n; i++)
        for (int j = 0; j < n; j++)
        {
            if (ctxs[i]->device == ctxs[j]->device) continue;
            int ok = 0;
            if (cudaDeviceCanAccessPeer(&ok, ctxs[i]->device, ctxs[j]->device) != cudaSuccess || !ok)
            {
                cudaGetLastError();
                return false;
            }
        }
    for (int i = 0; i < n; i++)
    {
        HL_CUDA(cudaSetDevice(ctxs[i]->device));
        for (int j = 0; j < n; j++)
        {
            if (ctxs[i]->device == ctxs[j]->device) continue;
            const cudaError_t e = cudaDeviceEnablePeerAccess(ctxs[j]->device, 0);
            if (e == cudaErrorPeerAccessAlreadyEnabled)
                cudaGetLastError();
            else
                HL_CUDA(e);
        }
    }
    return true;
}

static void check_group(hl_context_t* const* ctxs, int n, int root, const char* who)
{
    if (!ctxs || n < 1 || n > HL_MAX_COMM_RANKS || root < 0 || root >= n) throw CudaError(HL_ERR_INVALID_ARGUMENT, std::string(who) + ": invalid argument (1..16 contexts, root inside)");
    for (int i = 0; i < n; i++)
    {
        if (!ctxs[i]) throw CudaError(HL_ERR_INVALID_ARGUMENT, std::string(who) + ": null context");
        if (ctxs[i]->W != ctxs[0]->W || ctxs[i]->H != ctxs[0]->H) throw CudaError(HL_ERR_INVALID_ARGUMENT, std::string(who) + ": contexts differ in extent");
        for (int j = 0; j < i; j++)
            if (ctxs[i] == ctxs[j]) throw CudaError(HL_ERR_INVALID_ARGUMENT, std::string(who) + ": the same context twice");
    }
}

// reduce-scatter + (tone map) + gather into the root over peer memory; asynchronous: the root's main stream carries the result
static void peer_reduce(hl_context_t* const* ctxs, int n, int root, bool want_accum, bool want_rgba8, float exposure, int op, float scale)
{
    hl_context_t* R = ctxs[root];
    PeerImages    src;
    src.n = n;
    for (int i = 0; i < n; i++) src.accum[i] = ctxs[i]->accum.as<float4>();
    // every rank's frames in flight -> its main stream -> an event the others wait on
    for (int i = 0; i < n; i++)
    {
        hl_context_t* c = ctxs[i];
        HL_CUDA(cudaSetDevice(c->device));
        wavefront_join(c);
        if (!c->comm_ready_ev) HL_CUDA(cudaEventCreateWithFlags(&c->comm_ready_ev, cudaEventDisableTiming));
        if (!c->comm_done_ev) HL_CUDA(cudaEventCreateWithFlags(&c->comm_done_ev, cudaEventDisableTiming));
        HL_CUDA(cudaEventRecord(c->comm_ready_ev, c->stream));
    }
    const uint32_t total = R->W * R->H;
    // the sum is written in place into the root's image: its slice kernel must not start overwriting pixels other
    // ranks still read — slices are disjoint, and each pixel is read and written by the same thread, so only the
    // root's OWN input matters, which it reads before it writes.  Other ranks read root pixels of THEIR slice only.
    for (int i = 0; i < n; i++)
    {
        hl_context_t* c = ctxs[i];
        HL_CUDA(cudaSetDevice(c->device));
        for (int j = 0; j < n; j++)
            if (j != i) HL_CUDA(cudaStreamWaitEvent(c->stream, ctxs[j]->comm_ready_ev, 0));
        const uint32_t begin = (uint32_t)((uint64_t)total * i / n), end = (uint32_t)((uint64_t)total * (i + 1) / n);
        if (end > begin)
        {
            const uint32_t blocks = std::min<uint32_t>((end - begin + 255) / 256, (uint32_t)c->sm_count * 8u);
            k_peer_reduce<<<blocks, 256, 0, c->stream>>>(src, want_accum ? R->accum.as<float4>() : nullptr, want_rgba8 ? R->rgba8.as<uint32_t>() : nullptr, begin, end, R->W, R->H, exposure, op,
                                                         scale);
            c->launches++;
        }
        HL_CUDA(cudaEventRecord(c->comm_done_ev, c->stream));
    }
    // the root's stream carries the result; every other rank's stream also waits for all slices, so whatever it
    // enqueues next (a clear, new frames) cannot overwrite an image a peer is still reading
    for (int i = 0; i < n; i++)
    {
        HL_CUDA(cudaSetDevice(ctxs[i]->device));
        for (int j = 0; j < n; j++)
            if (j != i) HL_CUDA(cudaStreamWaitEvent(ctxs[i]->stream, ctxs[j]->comm_done_ev, 0));
    }
    HL_CUDA(cudaSetDevice(R->device));
    if (want_rgba8) R->rgba8_cur = R->rgba8.p;
    HL_CUDA(cudaGetLastError());
}

static void nccl_group_reduce(hl_context_t* const* ctxs, int n, int root, bool all)
{
    nccl_require();
    for (int i = 0; i < n; i++)
        if (!ctxs[i]->comm || ctxs[i]->comm_nranks != n || ctxs[i]->comm_rank != i) throw CudaError(HL_ERR_STATE, "hl_multi_gpu_reduce: the contexts were not bound together by hl_comm_init_all in this order");
    for (int i = 0; i < n; i++)
    {
        HL_CUDA(cudaSetDevice(ctxs[i]->device));
        wavefront_join(ctxs[i]);
    }
    const size_t count = (size_t)ctxs[0]->W * ctxs[0]->H * 4;
    HL_NCCL(nccl().GroupStart());
    for (int i = 0; i < n; i++)
    {
        hl_context_t* c = ctxs[i];
        if (all)
            HL_NCCL(nccl().AllReduce(c->accum.p, c->accum.p, count, kNcclFloat32, kNcclSum, (ncclComm_t)c->comm, c->stream));
        else
            HL_NCCL(nccl().Reduce(c->accum.p, c->accum.p, count, kNcclFloat32, kNcclSum, root, (ncclComm_t)c->comm, c->stream));
    }
    HL_NCCL(nccl().GroupEnd());
}

void comm_release(hl_context_t* c)
{
    if (c->comm && nccl().CommDestroy) nccl().CommDestroy((ncclComm_t)c->comm);
    c->comm = nullptr, c->comm_nranks = 1, c->comm_rank = 0;
    if (c->comm_ready_ev) cudaEventDestroy(c->comm_ready_ev), c->comm_ready_ev = nullptr;
    if (c->comm_done_ev) cudaEventDestroy(c->comm_done_ev), c->comm_done_ev = nullptr;
}
} // namespace hl

using namespace hl;

static thread_local std::string g_comm_error;

// group entry points report through every context of the group (and hl_last_error(NULL))
template <class F>
static hl_status group_call(hl_context* ctxs, int n, F&& f)
{
    int dev = 0;
    cudaGetDevice(&dev);
    hl_status st = HL_OK;
    try
    {
        f();
    }
    catch (const hl::CudaError& e)
    {
        g_comm_error = e.what(), st = e.status;
    }
    catch (const std::exception& e)
    {
        g_comm_error = e.what(), st = HL_ERR_CUDA;
    }
    if (st != HL_OK && ctxs)
        for (int i = 0; i < n && i < HL_MAX_COMM_RANKS; i++)
            if (ctxs[i]) ctxs[i]->err = g_comm_error;
    cudaSetDevice(dev);
    return st;
}

extern "C" {

const char* hl_comm_last_error(void) { return g_comm_error.c_str(); }

hl_status hl_comm_unique_id(uint8_t* id)
{
    return group_call(nullptr, 0, [&] {
        if (!id) throw CudaError(HL_ERR_INVALID_ARGUMENT, "hl_comm_unique_id: null pointer");
        nccl_require();
        ncclUniqueId u;
        HL_NCCL(nccl().GetUniqueId(&u));
        memcpy(id, &u, sizeof(u));
    });
}

hl_status hl_comm_init_rank(hl_context ctx, const uint8_t* id, int n_ranks, int rank)
{
    return group_call(&ctx, ctx ? 1 : 0, [&] {
        if (!ctx || !id || n_ranks < 1 || rank < 0 || rank >= n_ranks) throw CudaError(HL_ERR_INVALID_ARGUMENT, "hl_comm_init_rank: invalid argument");
        nccl_require();
        HL_CUDA(cudaSetDevice(ctx->device));
        comm_release(ctx);
        ncclUniqueId u;
        memcpy(&u, id, sizeof(u));
        ncclComm_t comm = nullptr;
        HL_NCCL(nccl().CommInitRank(&comm, n_ranks, u, rank));
        ctx->comm = comm, ctx->comm_nranks = n_ranks, ctx->comm_rank = rank;
    });
}

hl_status hl_comm_init_all(hl_context* ctxs, int n)
{
    return group_call(ctxs, n, [&] {
        check_group(ctxs, n, 0, "hl_comm_init_all");
        bool distinct = true;
        for (int i = 0; i < n; i++)
            for (int j = 0; j < i; j++)
                if (ctxs[i]->device == ctxs[j]->device) distinct = false;
        for (int i = 0; i < n; i++) comm_release(ctxs[i]);
        const bool p2p = peers_reachable(ctxs, n);
        if (!distinct && !p2p) throw CudaError(HL_ERR_STATE, "hl_comm_init_all: contexts share a device and peer access is unavailable");
        for (int i = 0; i < n; i++) ctxs[i]->comm_nranks = n, ctxs[i]->comm_rank = i, ctxs[i]->comm_p2p = p2p;
        if (distinct && n > 1 && nccl().error.empty())
        {
            // NCCL communicator as well: hl_accum_all_reduce / hl_accum_reduce work in this mode too, and it is the
            // fallback transport of hl_multi_gpu_reduce when some pair of GPUs has no peer access
            std::vector<int>        devs(n);
            std::vector<ncclComm_t> comms(n);
            for (int i = 0; i < n; i++) devs[i] = ctxs[i]->device;
            HL_NCCL(nccl().CommInitAll(comms.data(), n, devs.data()));
            for (int i = 0; i < n; i++) ctxs[i]->comm = comms[i];
        }
        else if (!p2p)
            nccl_require();
    });
}

hl_status hl_comm_destroy(hl_context ctx)
{
    return group_call(&ctx, ctx ? 1 : 0, [&] {
        if (!ctx) throw CudaError(HL_ERR_INVALID_ARGUMENT, "hl_comm_destroy: null context");
        HL_CUDA(cudaSetDevice(ctx->device));
        HL_CUDA(cudaStreamSynchronize(ctx->stream));
        comm_release(ctx);
    });
}

static hl_status rank_reduce(hl_context ctx, int root, bool all, const char* who)
{
    return group_call(&ctx, ctx ? 1 : 0, [&] {
        if (!ctx) throw CudaError(HL_ERR_INVALID_ARGUMENT, std::string(who) + ": null context");
        if (!ctx->comm) throw CudaError(HL_ERR_STATE, std::string(who) + ": no communicator (hl_comm_init_rank / hl_comm_init_all first)");
        if (!all && (root < 0 || root >= ctx->comm_nranks)) throw CudaError(HL_ERR_INVALID_ARGUMENT, std::string(who) + ": root out of range");
        HL_CUDA(cudaSetDevice(ctx->device));
        wavefront_join(ctx);
        const size_t count = (size_t)ctx->W * ctx->H * 4;
        if (all)
            HL_NCCL(nccl().AllReduce(ctx->accum.p, ctx->accum.p, count, kNcclFloat32, kNcclSum, (ncclComm_t)ctx->comm, ctx->stream));
        else
            HL_NCCL(nccl().Reduce(ctx->accum.p, ctx->accum.p, count, kNcclFloat32, kNcclSum, root, (ncclComm_t)ctx->comm, ctx->stream));
    });
}
hl_status hl_accum_all_reduce(hl_context ctx) { return rank_reduce(ctx, 0, true, "hl_accum_all_reduce"); }
hl_status hl_accum_reduce(hl_context ctx, int root) { return rank_reduce(ctx, root, false, "hl_accum_reduce"); }

hl_status hl_multi_gpu_reduce(hl_context* ctxs, int n, int root)
{
    return group_call(ctxs, n, [&] {
        check_group(ctxs, n, root, "hl_multi_gpu_reduce");
        bool p2p = true;
        for (int i = 0; i < n; i++) p2p = p2p && ctxs[i]->comm_p2p && ctxs[i]->comm_nranks == n && ctxs[i]->comm_rank == i;
        if (p2p)
            peer_reduce(ctxs, n, root, true, false, 1.0f, 0, 1.0f);
        else
            nccl_group_reduce(ctxs, n, root, false);
    });
}

hl_status hl_multi_gpu_resolve(hl_context* ctxs, int n, int root, float exposure, int op, float sample_scale, uint8_t* rgba8_host)
{
    return group_call(ctxs, n, [&] {
        check_group(ctxs, n, root, "hl_multi_gpu_resolve");
        if (op != HL_TONE_MAP_ACES && op != HL_TONE_MAP_REINHARD) throw CudaError(HL_ERR_INVALID_ARGUMENT, "hl_multi_gpu_resolve: unknown tone map operator");
        bool p2p = true;
        for (int i = 0; i < n; i++) p2p = p2p && ctxs[i]->comm_p2p && ctxs[i]->comm_nranks == n && ctxs[i]->comm_rank == i;
        hl_context_t* R = ctxs[root];
        if (p2p)
            peer_reduce(ctxs, n, root, true, true, exposure, op, sample_scale);
        else
        {
            nccl_group_reduce(ctxs, n, root, false);
            HL_CUDA(cudaSetDevice(R->device));
            film_tonemap(R, exposure, op, sample_scale);
        }
        if (rgba8_host)
        {
            HL_CUDA(cudaSetDevice(R->device));
            HL_CUDA(cudaMemcpyAsync(rgba8_host, R->rgba8.p, (size_t)R->W * R->H * 4, cudaMemcpyDeviceToHost, R->stream));
            HL_CUDA(cudaStreamSynchronize(R->stream));
        }
    });
}

} // extern "C"
