// context.cu — contexts, error text, NCCL binding (run-time), timing.
#include <dlfcn.h>
#include <nccl.h>
#include <stdarg.h>
#include <string.h>

#include <time.h>

#include "common.cuh"

static thread_local char g_err[1024] = "";

void rid_set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

extern "C" const char *rid_last_error(void) { return g_err; }
extern "C" int rid_version(void) { return 100; }

struct NcclApi {
    void *handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t,
                              cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Broadcast)(const void *, void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
};

static NcclApi *nccl_api()
{
    static NcclApi api;
    static bool tried = false;
    if (!tried) {
        tried = true;
        // If torch (or anything else) already mapped libnccl.so.2, RTLD_NOLOAD-style reuse happens through the soname.
        void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (h) {
            api.handle = h;
            api.GetUniqueId = (decltype(api.GetUniqueId))dlsym(h, "ncclGetUniqueId");
            api.CommInitRank = (decltype(api.CommInitRank))dlsym(h, "ncclCommInitRank");
            api.CommDestroy = (decltype(api.CommDestroy))dlsym(h, "ncclCommDestroy");
            api.AllReduce = (decltype(api.AllReduce))dlsym(h, "ncclAllReduce");
            api.AllGather = (decltype(api.AllGather))dlsym(h, "ncclAllGather");
            api.Broadcast = (decltype(api.Broadcast))dlsym(h, "ncclBroadcast");
            api.GroupStart = (decltype(api.GroupStart))dlsym(h, "ncclGroupStart");
            api.GroupEnd = (decltype(api.GroupEnd))dlsym(h, "ncclGroupEnd");
            api.GetErrorString = (decltype(api.GetErrorString))dlsym(h, "ncclGetErrorString");
        }
    }
    if (!api.handle || !api.GetUniqueId || !api.CommInitRank || !api.AllReduce) return nullptr;
    return &api;
}

#define RID_NCCL(api, call)                                                                            \
    do {                                                                                               \
        ncclResult_t r__ = (call);                                                                     \
        if (r__ != ncclSuccess) {                                                                      \
            rid_set_error("%s:%d: %s failed: %s", __FILE__, __LINE__, #call,                           \
                          (api)->GetErrorString ? (api)->GetErrorString(r__) : "nccl error");          \
            return RID_ERR_CUDA;                                                                       \
        }                                                                                              \
    } while (0)

extern "C" int rid_nccl_unique_id(void *uid128)
{
    NcclApi *api = nccl_api();
    if (!api) {
        rid_set_error("libnccl.so.2 could not be loaded: %s", dlerror());
        return RID_ERR_CUDA;
    }
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is expected to be 128 bytes");
    ncclUniqueId id;
    RID_NCCL(api, api->GetUniqueId(&id));
    memcpy(uid128, &id, sizeof(id));
    return RID_OK;
}

extern "C" int rid_ctx_create(int device, int rank, int world, const void *nccl_uid, rid_ctx **out)
{
    if (!out || world < 1 || rank < 0 || rank >= world) {
        rid_set_error("rid_ctx_create: bad arguments (rank=%d world=%d)", rank, world);
        return RID_ERR_ARG;
    }
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        rid_set_error("rid_ctx_create: no CUDA device is usable (%s); libraceid_b200 has no CPU fallback",
                      e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
        return RID_ERR_CUDA;
    }
    if (device < 0 || device >= ndev) {
        rid_set_error("rid_ctx_create: device %d out of range (0..%d)", device, ndev - 1);
        return RID_ERR_ARG;
    }
    RID_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    RID_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) {
        rid_set_error("rid_ctx_create: device %d is sm_%d%d; this library is built for sm_100a only", device,
                      prop.major, prop.minor);
        return RID_ERR_CUDA;
    }
    rid_ctx *ctx = new rid_ctx();
    ctx->device = device;
    ctx->rank = rank;
    ctx->world = world;
    ctx->sm_count = prop.multiProcessorCount;
    RID_CUDA(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
    RID_CUDA(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
    RID_CUDA(cudaEventCreate(&ctx->ev0));
    RID_CUDA(cudaEventCreate(&ctx->ev1));
    if (world > 1) {
        if (!nccl_uid) {
            rid_set_error("rid_ctx_create: world>1 needs the NCCL unique id");
            delete ctx;
            return RID_ERR_ARG;
        }
        NcclApi *api = nccl_api();
        if (!api) {
            rid_set_error("libnccl.so.2 could not be loaded");
            delete ctx;
            return RID_ERR_CUDA;
        }
        ncclUniqueId id;
        memcpy(&id, nccl_uid, sizeof(id));
        ncclComm_t comm;
        RID_NCCL(api, api->CommInitRank(&comm, world, id, rank));
        ctx->comm = comm;
        ctx->nccl = api;
    }
    *out = ctx;
    return RID_OK;
}

extern "C" int rid_ctx_destroy(rid_ctx *ctx)
{
    if (!ctx) return RID_OK;
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    if (ctx->copy_stream) cudaStreamSynchronize(ctx->copy_stream);
    // (No collective here: a rank whose peers are already gone must still be able to leave.  Callers that destroy
    // contexts while the job goes on call rid_ctx_trim first, which closes the mappings everywhere before anything is
    // freed.)
    for (auto &m : ctx->ipc) cudaIpcCloseMemHandle(m.ptr);
    ctx->ipc.clear();
    ctx->exported.clear();
    rid_pool_trim(ctx);
    if (ctx->scratch) cudaFree(ctx->scratch);
    if (ctx->pin) cudaFreeHost(ctx->pin);
    if (ctx->prof_dev) cudaFree(ctx->prof_dev);
    if (ctx->i8_tiles) cudaFree(ctx->i8_tiles);
    if (ctx->comm && ctx->nccl && ctx->nccl->CommDestroy) ctx->nccl->CommDestroy((ncclComm_t)ctx->comm);
    for (int i = 0; i < 4; ++i) if (ctx->phase_ev[i]) cudaEventDestroy(ctx->phase_ev[i]);
    if (ctx->ev0) cudaEventDestroy(ctx->ev0);
    if (ctx->ev1) cudaEventDestroy(ctx->ev1);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
    delete ctx;
    return RID_OK;
}

extern "C" int rid_ctx_sync(rid_ctx *ctx)
{
    if (!ctx) return RID_ERR_ARG;
    RID_CUDA(cudaSetDevice(ctx->device));
    RID_CUDA(cudaStreamSynchronize(ctx->stream));
    return RID_OK;
}

extern "C" double rid_ctx_last_ms(const rid_ctx *ctx) { return ctx ? ctx->last_ms : 0.0; }
extern "C" long long rid_ctx_launches(const rid_ctx *ctx) { return ctx ? ctx->launches : 0; }
extern "C" long long rid_ctx_swaps(const rid_ctx *ctx) { return ctx ? ctx->swaps : 0; }
extern "C" double rid_ctx_stat(const rid_ctx *ctx, int which)
{
    if (!ctx || which < 0 || which >= 12) return 0.0;
    return which == 0 ? (double)ctx->swaps : ctx->stat[which];
}

// Collective when world > 1: closes the peers' mappings, then gives every cached device buffer back to the driver.
extern "C" int rid_ctx_trim(rid_ctx *ctx)
{
    if (!ctx) return RID_ERR_ARG;
    RID_CUDA(cudaSetDevice(ctx->device));
    RID_CUDA(cudaStreamSynchronize(ctx->stream));
    if (ctx->world > 1 && ctx->live_sharded == 0) RID_TRY(rid_ipc_release_all(ctx));
    rid_pool_trim(ctx);
    return RID_OK;
}

int rid_time_begin(rid_ctx *ctx)
{
    RID_CUDA(cudaSetDevice(ctx->device));
    RID_CUDA(cudaEventRecord(ctx->ev0, ctx->stream));
    return RID_OK;
}

// every recorded bracket has completed (the stream was just synchronised): add them to the per-kind totals and recycle
// their events
static void prof_fold(rid_ctx *ctx)
{
    for (auto &r : ctx->prof) {
        float msf = 0.f;
        if (cudaEventElapsedTime(&msf, r.a, r.b) == cudaSuccess && r.kind >= 0 && r.kind < 16) {
            auto &a = ctx->prof_acc[r.kind];
            a.n += 1; a.ms += msf; a.work += r.work; a.traffic += r.traffic;
        } else {
            cudaGetLastError();
        }
        ctx->prof_pool.push_back(r.a);
        ctx->prof_pool.push_back(r.b);
    }
    ctx->prof.clear();
}

int rid_time_end(rid_ctx *ctx)
{
    RID_CUDA(cudaEventRecord(ctx->ev1, ctx->stream));
    RID_CUDA(cudaEventSynchronize(ctx->ev1));
    float ms = 0.f;
    RID_CUDA(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
    ctx->last_ms = ms;
    // (only when the pool of pre-created events runs low: reading back a step's ~5 000 brackets takes a few ms of host time)
    if (ctx->prof_on && ctx->prof_pool.size() < 12288 && ctx->prof.size() > 2048) {
        RID_CUDA(cudaStreamSynchronize(ctx->stream));
        prof_fold(ctx);
    }
    return RID_OK;
}

// ---- device buffer reuse pool ---------------------------------------------------------------------------
// RID_TRACE=1: every pool miss / eviction (a real cudaMalloc / cudaFree) is reported on stderr with its duration
static bool rid_tracing()
{
    static int on = -1;
    if (on < 0) on = getenv("RID_TRACE") ? 1 : 0;
    return on == 1;
}
static double now_ms()
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
}

static bool is_exported(const rid_ctx *ctx, const void *p)
{
    for (void *q : ctx->exported)
        if (q == p) return true;
    return false;
}

cudaError_t rid_pool_alloc(rid_ctx *ctx, void **p, size_t bytes)
{
    for (size_t i = 0; i < ctx->pool.size(); ++i)
        if (ctx->pool[i].bytes == bytes) {
            *p = ctx->pool[i].p;
            ctx->pool.erase(ctx->pool.begin() + i);
            return cudaSuccess;
        }
    const double t0 = rid_tracing() ? now_ms() : 0.0;
    cudaError_t e = cudaMalloc(p, bytes);
    if (e != cudaSuccess && !ctx->pool.empty()) {
        cudaGetLastError();
        rid_pool_trim(ctx); // give cached buffers back and retry
        e = cudaMalloc(p, bytes);
    }
    if (rid_tracing()) fprintf(stderr, "[rid] pool miss: cudaMalloc %.1f MB took %.2f ms (%zu cached)\n", bytes / 1e6, now_ms() - t0, ctx->pool.size());
    return e;
}

void rid_pool_free(rid_ctx *ctx, void *p, size_t bytes)
{
    if (!p) return;
    const double t0 = rid_tracing() ? now_ms() : 0.0;
    if (bytes < ((size_t)256 << 10)) {
        cudaFree(p);
        if (rid_tracing()) fprintf(stderr, "[rid] small cudaFree %.3f MB took %.2f ms\n", bytes / 1e6, now_ms() - t0);
        return;
    }
    if (ctx->pool.size() >= 64) { // full: drop the buffer that has waited longest, the steady-state sizes stay cached
        for (size_t i = 0; i < ctx->pool.size(); ++i) {
            if (is_exported(ctx, ctx->pool[i].p)) continue; // a peer may have it mapped (CUDA IPC)
            cudaFree(ctx->pool[i].p);
            if (rid_tracing()) fprintf(stderr, "[rid] pool full: cudaFree %.1f MB took %.2f ms\n", ctx->pool[i].bytes / 1e6, now_ms() - t0);
            ctx->pool.erase(ctx->pool.begin() + i);
            break;
        }
    }
    ctx->pool.push_back({p, bytes});
}

void *rid_pinned(rid_ctx *ctx, size_t bytes)
{
    if (bytes <= ctx->pin_bytes) return ctx->pin;
    if (ctx->pin) { cudaStreamSynchronize(ctx->stream); cudaFreeHost(ctx->pin); ctx->pin = nullptr; ctx->pin_bytes = 0; }
    bytes = (bytes + ((size_t)1 << 20) - 1) & ~(((size_t)1 << 20) - 1);
    if (cudaHostAlloc(&ctx->pin, bytes, cudaHostAllocDefault) != cudaSuccess) { cudaGetLastError(); ctx->pin = nullptr; return nullptr; }
    ctx->pin_bytes = bytes;
    return ctx->pin;
}

void rid_pool_trim(rid_ctx *ctx)
{
    // buffers whose IPC handle went out stay (freeing memory a peer still has mapped is undefined behaviour)
    std::vector<rid_ctx::PoolBuf> keep;
    for (auto &b : ctx->pool) {
        if (is_exported(ctx, b.p)) keep.push_back(b);
        else cudaFree(b.p);
    }
    ctx->pool.swap(keep);
}

// ---- per-kernel profiling -------------------------------------------------------------------------------
static cudaEvent_t prof_event(rid_ctx *ctx)
{
    if (!ctx->prof_pool.empty()) {
        cudaEvent_t e = ctx->prof_pool.back();
        ctx->prof_pool.pop_back();
        return e;
    }
    cudaEvent_t e = nullptr;
    cudaEventCreate(&e);
    return e;
}

cudaEvent_t rid_prof_begin(rid_ctx *ctx, int kind)
{
    if (!((ctx->prof_on >> kind) & 1)) return nullptr;
    cudaEvent_t a = prof_event(ctx);
    cudaEventRecord(a, ctx->stream);
    return a;
}

void rid_prof_end(rid_ctx *ctx, cudaEvent_t a, int kind, double work, double traffic)
{
    if (!a) return;
    cudaEvent_t b = prof_event(ctx);
    cudaEventRecord(b, ctx->stream);
    ctx->prof.push_back({kind, a, b, work, traffic});
}

// enable: bit mask of kernel kinds to record (-1 = all; clears what was collected); 0: stop
extern "C" int rid_ctx_profile(rid_ctx *ctx, int enable)
{
    if (!ctx) return RID_ERR_ARG;
    RID_CUDA(cudaSetDevice(ctx->device));
    RID_CUDA(cudaStreamSynchronize(ctx->stream));
    for (auto &r : ctx->prof) { ctx->prof_pool.push_back(r.a); ctx->prof_pool.push_back(r.b); }
    ctx->prof.clear();
    for (auto &a : ctx->prof_acc) a = rid_ctx::ProfAcc();
    if (enable && !ctx->prof_dev) RID_CUDA(cudaMalloc(&ctx->prof_dev, 4 * sizeof(u64)));
    if (ctx->prof_dev) RID_CUDA(cudaMemsetAsync(ctx->prof_dev, 0, 4 * sizeof(u64), ctx->stream));
    ctx->prof_on = enable; // bit k set = record kernels of kind k (enable = -1: every kind)
    // create the events up front: cudaEventCreate inside the timed region costs more than the kernels it brackets
    if (enable) {
        const bool many = enable != -1 && (enable & ((1 << RID_PROF_SCAN_BUILD) | (1 << RID_PROF_SCAN_DELTA)));
        const size_t want = enable == -1 ? 40960 : (many ? 24576 : 2048);
        while (ctx->prof_pool.size() < want) {
            cudaEvent_t e = nullptr;
            if (cudaEventCreateWithFlags(&e, cudaEventDefault) != cudaSuccess) break;
            ctx->prof_pool.push_back(e);
        }
    } else {
        // tens of thousands of live events slow every later synchronisation down: give them back
        for (cudaEvent_t e : ctx->prof_pool) cudaEventDestroy(e);
        ctx->prof_pool.clear();
    }
    return RID_OK;
}

// totals for one kernel kind since profiling was enabled: launches, summed device ms, summed algorithmic work
// (bytes for the scans, flops for the Gram kernel) and summed bytes actually requested from memory
extern "C" int rid_ctx_profile_get(rid_ctx *ctx, int kind, long long *launches, double *ms, double *work,
                                   double *traffic)
{
    if (!ctx || kind < 0 || kind >= RID_PROF_NKINDS) return RID_ERR_ARG;
    RID_CUDA(cudaSetDevice(ctx->device));
    RID_CUDA(cudaStreamSynchronize(ctx->stream));
    prof_fold(ctx);
    long long n = ctx->prof_acc[kind].n;
    double t = ctx->prof_acc[kind].ms, w = ctx->prof_acc[kind].work, tr = ctx->prof_acc[kind].traffic;
    for (auto &r : ctx->prof)
        if (r.kind == kind) {
            float msf = 0.f;
            RID_CUDA(cudaEventElapsedTime(&msf, r.a, r.b));
            ++n; t += msf; w += r.work; tr += r.traffic;
        }
    if (ctx->prof_dev && (kind == RID_PROF_SCAN_BUILD || kind == RID_PROF_SCAN_DELTA)) {
        // passes sized on the device: their work was counted there (cells -> bytes)
        u64 h[4];
        RID_CUDA(cudaMemcpyAsync(h, ctx->prof_dev, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream));
        RID_CUDA(cudaStreamSynchronize(ctx->stream));
        const int o = kind == RID_PROF_SCAN_BUILD ? 0 : 2;
        w += 4.0 * (double)h[o];
        tr += 4.0 * (double)h[o + 1];
    }
    if (launches) *launches = n;
    if (ms) *ms = t;
    if (work) *work = w;
    if (traffic) *traffic = tr;
    return RID_OK;
}

// user events on the library's stream (bench.py brackets its timed region with them)
extern "C" int rid_ctx_event_record(rid_ctx *ctx, int slot)
{
    if (!ctx || slot < 0 || slot >= 8) return RID_ERR_ARG;
    RID_CUDA(cudaSetDevice(ctx->device));
    if (!ctx->user_ev[slot]) RID_CUDA(cudaEventCreate(&ctx->user_ev[slot]));
    RID_CUDA(cudaEventRecord(ctx->user_ev[slot], ctx->stream));
    return RID_OK;
}

extern "C" int rid_ctx_event_elapsed_ms(rid_ctx *ctx, int slot_a, int slot_b, double *ms)
{
    if (!ctx || !ms || slot_a < 0 || slot_a >= 8 || slot_b < 0 || slot_b >= 8 || !ctx->user_ev[slot_a] ||
        !ctx->user_ev[slot_b])
        return RID_ERR_ARG;
    RID_CUDA(cudaSetDevice(ctx->device));
    RID_CUDA(cudaEventSynchronize(ctx->user_ev[slot_b]));
    float f = 0.f;
    RID_CUDA(cudaEventElapsedTime(&f, ctx->user_ev[slot_a], ctx->user_ev[slot_b]));
    *ms = f;
    return RID_OK;
}

static int allreduce(rid_ctx *ctx, void *buf, size_t count, ncclDataType_t dt)
{
    if (ctx->world == 1) return RID_OK;
    RID_NCCL(ctx->nccl, ctx->nccl->AllReduce(buf, buf, count, dt, ncclSum, (ncclComm_t)ctx->comm, ctx->stream));
    return RID_OK;
}
int rid_allreduce_u64(rid_ctx *ctx, u64 *buf, size_t count) { return allreduce(ctx, buf, count, ncclUint64); }
int rid_allreduce_i32(rid_ctx *ctx, int *buf, size_t count) { return allreduce(ctx, buf, count, ncclInt32); }
int rid_allreduce_f64(rid_ctx *ctx, double *buf, size_t count) { return allreduce(ctx, buf, count, ncclFloat64); }

// ---- peer memory (CUDA IPC) ---------------------------------------------------------------------------------
static int ensure_scratch(rid_ctx *ctx)
{
    if (!ctx->scratch) {
        RID_CUDA(cudaMalloc(&ctx->scratch, 4096));
        // (on the library's stream: it is not ordered against the legacy default stream)
        RID_CUDA(cudaMemsetAsync(ctx->scratch, 0, 4096, ctx->stream));
    }
    return RID_OK;
}

int rid_allgather(rid_ctx *ctx, const void *send, void *recv, size_t bytes_per_rank)
{
    if (ctx->world == 1) {
        if (send != recv) RID_CUDA(cudaMemcpyAsync(recv, send, bytes_per_rank, cudaMemcpyDeviceToDevice, ctx->stream));
        return RID_OK;
    }
    if (!ctx->nccl->AllGather) { rid_set_error("ncclAllGather is not available"); return RID_ERR_CUDA; }
    RID_NCCL(ctx->nccl, ctx->nccl->AllGather(send, recv, bytes_per_rank, ncclUint8, (ncclComm_t)ctx->comm, ctx->stream));
    return RID_OK;
}

// Every rank holds the byte range [off[r], off[r+1]) of `buf` that rank r produced (uploaded); afterwards every rank holds
// all of them: one in-place broadcast per rank, grouped.  Returns RID_ERR_UNSUPPORTED if the NCCL entry points are missing.
int rid_bcast_ranges(rid_ctx *ctx, void *buf, const size_t *off)
{
    if (ctx->world == 1) return RID_OK;
    if (!ctx->nccl->Broadcast || !ctx->nccl->GroupStart || !ctx->nccl->GroupEnd) return RID_ERR_UNSUPPORTED;
    RID_NCCL(ctx->nccl, ctx->nccl->GroupStart());
    for (int r = 0; r < ctx->world; ++r) {
        const size_t n = off[r + 1] - off[r];
        if (!n) continue;
        char *p = (char *)buf + off[r];
        RID_NCCL(ctx->nccl, ctx->nccl->Broadcast(p, p, n, ncclUint8, r, (ncclComm_t)ctx->comm, ctx->stream));
    }
    RID_NCCL(ctx->nccl, ctx->nccl->GroupEnd());
    return RID_OK;
}

int rid_barrier(rid_ctx *ctx)
{
    if (ctx->world == 1) return RID_OK;
    RID_TRY(ensure_scratch(ctx));
    return rid_allreduce_i32(ctx, (int *)ctx->scratch + 512, 1);
}

// base of the allocation a device pointer lies in (cuMemGetAddressRange through the runtime's driver entry points)
static void *alloc_base(void *p)
{
    typedef int (*PFN_range)(unsigned long long *, size_t *, unsigned long long);
    static PFN_range fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void *f = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuMemGetAddressRange", &f, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (PFN_range)f;
        else
            cudaGetLastError();
    }
    if (!fn) return p;
    unsigned long long base = 0;
    size_t size = 0;
    if (fn(&base, &size, (unsigned long long)(uintptr_t)p) != 0) return p;
    return (void *)(uintptr_t)base;
}

int rid_peer_pointers(rid_ctx *ctx, void *mine, void **peers, int *all_ok)
{
    *all_ok = 0;
    for (int r = 0; r < ctx->world; ++r) peers[r] = nullptr;
    peers[ctx->rank] = mine;
    if (ctx->world == 1) { *all_ok = 1; return RID_OK; }
    if (ctx->world > 16 || !ctx->nccl->AllGather) return RID_OK;
    RID_TRY(ensure_scratch(ctx));
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is expected to be 64 bytes");
    // record: 64-byte handle of the ALLOCATION + 8-byte offset of `mine` inside it (cudaIpcOpenMemHandle returns the base)
    const size_t REC = 72;
    unsigned char rec[72];
    void *base = alloc_base(mine);
    cudaIpcMemHandle_t h;
    int ok = cudaIpcGetMemHandle(&h, base) == cudaSuccess ? 1 : 0;
    if (!ok) { cudaGetLastError(); memset(&h, 0, sizeof(h)); }
    const unsigned long long off = (unsigned long long)((char *)mine - (char *)base);
    memcpy(rec, &h, 64);
    memcpy(rec + 64, &off, 8);
    if (ok && !is_exported(ctx, base)) ctx->exported.push_back(base);
    unsigned char *dev = (unsigned char *)ctx->scratch; // world * 72 bytes <= 1152 (the barrier / flag words sit at 2048+)
    RID_CUDA(cudaMemcpyAsync(dev + REC * ctx->rank, rec, REC, cudaMemcpyHostToDevice, ctx->stream));
    RID_NCCL(ctx->nccl, ctx->nccl->AllGather(dev + REC * ctx->rank, dev, REC, ncclUint8, (ncclComm_t)ctx->comm, ctx->stream));
    unsigned char all[16 * 72];
    RID_CUDA(cudaMemcpyAsync(all, dev, REC * (size_t)ctx->world, cudaMemcpyDeviceToHost, ctx->stream));
    RID_CUDA(cudaStreamSynchronize(ctx->stream));
    for (int r = 0; r < ctx->world && ok; ++r) {
        if (r == ctx->rank) continue;
        void *ptr = nullptr;
        for (auto &m : ctx->ipc)
            if (memcmp(m.handle, all + REC * r, 64) == 0) { ptr = m.ptr; break; }
        if (!ptr) {
            cudaIpcMemHandle_t ph;
            memcpy(&ph, all + REC * r, 64);
            if (cudaIpcOpenMemHandle(&ptr, ph, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
                cudaGetLastError();
                ok = 0;
                break;
            }
            rid_ctx::IpcMap m;
            memcpy(m.handle, all + REC * r, 64);
            m.ptr = ptr;
            ctx->ipc.push_back(m);
        }
        unsigned long long poff = 0;
        memcpy(&poff, all + REC * r + 64, 8);
        peers[r] = (char *)ptr + poff;
    }
    // every rank must take the same path
    int *flag = (int *)ctx->scratch + 600;
    RID_CUDA(cudaMemcpyAsync(flag, &ok, sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    RID_TRY(rid_allreduce_i32(ctx, flag, 1));
    int sum = 0;
    RID_CUDA(cudaMemcpyAsync(&sum, flag, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    RID_CUDA(cudaStreamSynchronize(ctx->stream));
    *all_ok = sum == ctx->world ? 1 : 0;
    return RID_OK;
}

int rid_agree_all(rid_ctx *ctx, int ok, int *all_ok)
{
    *all_ok = ok ? 1 : 0;
    if (ctx->world == 1) return RID_OK;
    RID_TRY(ensure_scratch(ctx));
    int *flag = (int *)ctx->scratch + 680;
    int v = ok ? 1 : 0;
    RID_CUDA(cudaMemcpyAsync(flag, &v, sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    RID_TRY(rid_allreduce_i32(ctx, flag, 1));
    RID_CUDA(cudaMemcpyAsync(&v, flag, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    RID_CUDA(cudaStreamSynchronize(ctx->stream));
    *all_ok = v == ctx->world ? 1 : 0;
    return RID_OK;
}

int rid_ipc_release_all(rid_ctx *ctx)
{
    if (ctx->world == 1) return RID_OK;
    RID_CUDA(cudaStreamSynchronize(ctx->stream));
    if (ctx->copy_stream) RID_CUDA(cudaStreamSynchronize(ctx->copy_stream));
    for (auto &m : ctx->ipc) cudaIpcCloseMemHandle(m.ptr);
    ctx->ipc.clear();
    RID_TRY(rid_barrier(ctx)); // nobody holds a mapping of anybody's buffer any more
    RID_CUDA(cudaStreamSynchronize(ctx->stream));
    ctx->exported.clear();
    return RID_OK;
}

// ---- cancellation ---------------------------------------------------------------------------------------------
extern "C" int rid_ctx_request_abort(rid_ctx *ctx)
{
    if (!ctx) return RID_ERR_ARG;
    ctx->abort_flag = 1;
    return RID_OK;
}
extern "C" int rid_ctx_clear_abort(rid_ctx *ctx)
{
    if (!ctx) return RID_ERR_ARG;
    ctx->abort_flag = 0;
    return RID_OK;
}
extern "C" int rid_ctx_set_interrupt_poll(rid_ctx *ctx, int (*poll)(void *), void *arg)
{
    if (!ctx) return RID_ERR_ARG;
    ctx->poll_fn = poll;
    ctx->poll_arg = arg;
    return RID_OK;
}

int rid_check_abort(rid_ctx *ctx, bool collective)
{
    int want = ctx->abort_flag ? 1 : 0;
    if (!want && ctx->poll_fn && ctx->poll_fn(ctx->poll_arg)) want = 1;
    if (ctx->world > 1) {
        if (!collective) return RID_OK; // a rank must never leave a collective sequence on its own
        RID_TRY(ensure_scratch(ctx));
        int *flag = (int *)ctx->scratch + 640;
        RID_CUDA(cudaMemcpyAsync(flag, &want, sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
        RID_TRY(rid_allreduce_i32(ctx, flag, 1));
        RID_CUDA(cudaMemcpyAsync(&want, flag, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        RID_CUDA(cudaStreamSynchronize(ctx->stream));
    }
    if (!want) return RID_OK;
    ctx->abort_flag = 0;
    cudaStreamSynchronize(ctx->stream);
    rid_set_error("interrupted: the caller asked for cancellation");
    return RID_ERR_INTERRUPTED;
}
