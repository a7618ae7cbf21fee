// bcfgpu.cu -- host side of libbcfgpu.so: the C-ABI declared in include/bcfgpu.h.
// No CPU fallback: every compute entry point needs a CUDA device and fails with BCFGPU_ERR_CUDA otherwise.
#include "../../include/bcfgpu.h"
#include "bcf_kernels.cuh"
#include "bcf_persistent.cuh"
#include "bcf_batched.cuh"
#include "bcf_pipe.cuh"
#include "bcf_hist_mma.cuh"
#include "bcf_host.h"

#include <dlfcn.h>
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <chrono>
#include <cstring>
#include <atomic>
#include <condition_variable>
#include <map>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

using namespace bcf;

static_assert(MAXL == BCFGPU_MAX_LEAVES, "header / device constant mismatch");
static_assert(MAX_DEPTH == BCFGPU_MAX_DEPTH, "header / device constant mismatch");

// ------------------------------------------------------------------------------------------------
// errors
// ------------------------------------------------------------------------------------------------
static thread_local std::string g_err;
static int fail(int code, const std::string& msg) { g_err = msg; return code; }
#define CK(call)                                                                                       \
    do {                                                                                               \
        cudaError_t e_ = (call);                                                                       \
        if (e_ != cudaSuccess)                                                                         \
            return fail(e_ == cudaErrorMemoryAllocation ? BCFGPU_ERR_OOM : BCFGPU_ERR_CUDA,            \
                        std::string(#call) + ": " + cudaGetErrorString(e_));                           \
    } while (0)

// ------------------------------------------------------------------------------------------------
// NCCL through dlopen (torch ships libnccl.so.2; only the sharded mode needs it)
// ------------------------------------------------------------------------------------------------
struct NcclId { char b[128]; };
struct NcclApi {
    void* lib = nullptr;
    int (*GetUniqueId)(NcclId*) = nullptr;
    int (*CommInitRank)(void**, int, NcclId, int) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
    int (*AllGather)(const void*, void*, size_t, int, void*, cudaStream_t) = nullptr;
    int (*CommDestroy)(void*) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
};
static NcclApi g_nccl;
static int nccl_load()
{
    if (g_nccl.lib) return BCFGPU_OK;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    void* lib = nullptr;
    for (const char* nm : names) { lib = dlopen(nm, RTLD_NOW | RTLD_GLOBAL); if (lib) break; }
    if (!lib) return fail(BCFGPU_ERR_NCCL, "cannot dlopen libnccl.so.2 (put torch's nvidia/nccl/lib on LD_LIBRARY_PATH or import torch first)");
    g_nccl.GetUniqueId = (int (*)(NcclId*))dlsym(lib, "ncclGetUniqueId");
    g_nccl.CommInitRank = (int (*)(void**, int, NcclId, int))dlsym(lib, "ncclCommInitRank");
    g_nccl.AllReduce = (int (*)(const void*, void*, size_t, int, int, void*, cudaStream_t))dlsym(lib, "ncclAllReduce");
    g_nccl.AllGather = (int (*)(const void*, void*, size_t, int, void*, cudaStream_t))dlsym(lib, "ncclAllGather");
    g_nccl.CommDestroy = (int (*)(void*))dlsym(lib, "ncclCommDestroy");
    g_nccl.GetErrorString = (const char* (*)(int))dlsym(lib, "ncclGetErrorString");
    if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.AllReduce || !g_nccl.CommDestroy)
        return fail(BCFGPU_ERR_NCCL, "libnccl is missing expected symbols");
    g_nccl.lib = lib;
    return BCFGPU_OK;
}
constexpr int NCCL_INT8 = 0, NCCL_INT64 = 4, NCCL_SUM = 0, NCCL_MIN = 3;
#define NK(call)                                                                                       \
    do {                                                                                               \
        int r_ = (call);                                                                               \
        if (r_ != 0) return fail(BCFGPU_ERR_NCCL, std::string(#call) + ": " +                          \
                                 (g_nccl.GetErrorString ? g_nccl.GetErrorString(r_) : "nccl error"));   \
    } while (0)

// ------------------------------------------------------------------------------------------------
// handle
// ------------------------------------------------------------------------------------------------
struct Block { void* p; size_t bytes; };   // a block of the caching allocator (below)
struct LocalGroup;
struct ForestHost {
    int p = 0, p8 = 0, p16 = 0;
    std::vector<int32_t> off, ncut, col_index;
    std::vector<uint8_t> is16;
    std::vector<double> cuts;
    double* d_cuts = nullptr; int32_t* d_off = nullptr; int32_t* d_ncut = nullptr; int32_t* d_col_index = nullptr;
    uint8_t* d_is16 = nullptr; uint8_t* d_bins8 = nullptr; uint16_t* d_bins16 = nullptr;
};

struct bcfgpu_handle {
    bcfgpu_config cfg;
    int device = 0, num_sms = 148, max_optin = 0;
    cudaStream_t stream = nullptr, copy_stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    cudaEvent_t ev_copied[2] = {nullptr, nullptr}, ev_binned[2] = {nullptr, nullptr};   // set_data: chunks of X in flight
    bool have_pos = false;            // h->pos mirrors d_pos
    int rank = 0, nranks = 1;
    void* comm = nullptr;
    Mailbox* d_mail = nullptr;        // [nranks] this rank's mailboxes (peer-memory exchange of the sharded mode)
    Mailbox* peer_mail[MAX_RANKS] = {nullptr};   // the peers' mailbox arrays, mapped with cudaIpc
    LocalGroup* local = nullptr;      // shards inside one process (bcfgpu_set_shard_local)
    bool no_pipe = false;             // ... three or more of them on this device: the batched kernel only
    bool peer_ok = false;             // every rank mapped every peer: the sweep kernel exchanges through NVLink itself
    unsigned long long launch_id = 0;
    int64_t pipe_sweeps = 0, fallback_sweeps = 0;   // sweeps the pipelined kernel ran / handed to the batched kernel (a tree wider than 8 keys)
    unsigned long long pipe_batches = 0;   // batches the pipelined kernel has exchanged with the peers so far (tags of Mailbox::ll)
    bool shard_decided = false, shard_persistent = false;
    int last_kernel = 0;
    int64_t n = 0, n_pad = 0, n_global = 0;
    int ntrt = 0, T = 0;
    bool has_data = false;
    std::vector<Block> allocs;        // data-lifetime allocations (blocks of the caching allocator)
    std::vector<Block> run_allocs;    // run-lifetime allocations (posterior buffers)
    ForestHost fh[2];
    std::vector<int32_t> pos;
    int32_t *d_pos = nullptr, *d_orig = nullptr;
    double* d_tmp = nullptr;          // n_pad doubles scratch
    double* d_tmp2 = nullptr;
    Dev dev;
    Dev graph_dev;                    // Dev the graph was captured with
    cudaGraph_t graph = nullptr;
    cudaGraphExec_t gexec = nullptr;
    int grid_obs = 0, grid_fin = 0;
    int engine = BCFGPU_ENGINE_GRAPH;
    int coop_grid = 0;
    int64_t h2d_bytes = 0;                    // copied by the last set_data
    int res_ntl = 0;                  // tiles per CTA of the shared-memory-resident kernels (0 = state in HBM)
    int kernel = 0;                   // persistent kernel in use: 1 k_persistent, 2 k_resident, 3 k_batched, 4 k_pipe (+ k_batched)
    int fallback_kernel = 3;          // what k_pipe hands a sweep with a wide tree to: 3 (k_batched resident) or 5 (HBM)
    size_t pipe_smem = 0;             // dynamic shared memory of k_pipe (kernel 4; res_smem is then k_batched's)
    int pipe_ntl = 0;                 // tiles per CTA of k_pipe (res_ntl is then k_batched's: 0 when that one keeps its state in HBM)
    int* d_done = nullptr;            // sweeps completed by the last k_pipe launch
    size_t res_smem = 0;
    unsigned int* d_bar = nullptr;
    int64_t launches = 0;
    double last_ms = 0.0;
    int32_t nsaved = 0, post_cap = 0;
    int keep = 0;
    // saved forests (keep_draws bit 3): node pool of the last run, per (draw, tree) offsets, per-draw scales
    SavedNode* d_pool = nullptr; long long pool_cap = 0; unsigned long long* d_pool_used = nullptr;
    long long* d_tree_off = nullptr; int32_t* d_tree_cnt = nullptr; SavedDraw* d_draw_scal = nullptr;
    int32_t forests_saved = 0;
    int* dbg_host = nullptr;          // mapped pinned words the kernels write before they trap (Dev::dbg)
    uint8_t* d_ybin = nullptr;        // probit BART: the observed 0/1 response, internal order
    bool latent_drawn = false;        // probit BART: the latent of the next sweep exists (k_pipe draws it in its finishing pass)
    bool fuse_latent = false;         // probit BART: k_pipe<true> is in use (BCFGPU_NO_FUSED_LATENT=1: k_probit_latent before every sweep)
    int* d_cflist = nullptr;          // probit BART: trees that split on the treatment (k_cf_collect)
    std::atomic<int> stop{0};         // bcfgpu_request_stop: the run in progress ends at its next launch boundary
};

// ------------------------------------------------------------------------------------------------
// Caching allocator.  cudaMalloc / cudaFree of the sampler's buffers (1.2 GB at n = 1e6, p = 100) cost 5 - 35 ms per
// fit and vary from call to call; bcf_iv() fits twice per call and bcf(n_chains) many times, always with the same sizes.
// Freed blocks are therefore parked per device (exact size classes: multiples of 2 MiB above 1 MiB, of 512 B below) and
// handed out again; the cache is bounded (BCFGPU_CACHE_MB, default 4096 per device; 0 = off) and
// bcfgpu_release_cached_memory() empties it.  The same for the pinned host staging of pageable inputs.
// ------------------------------------------------------------------------------------------------
struct MemCache {
    std::mutex mu;
    std::multimap<size_t, void*> free_;
    size_t bytes = 0;
};
constexpr int MAX_DEVICES = 64;
static MemCache g_dev_cache[MAX_DEVICES];
static MemCache g_pin_cache;
static size_t cache_cap_bytes()
{
    static const size_t cap = [] { const char* e = getenv("BCFGPU_CACHE_MB"); return (size_t)(e ? std::max(0, atoi(e)) : 4096) << 20; }();
    return cap;
}
static size_t size_class(size_t bytes)
{
    bytes = std::max<size_t>(bytes, 16);
    const size_t g = bytes >= ((size_t)1 << 20) ? ((size_t)2 << 20) : 512;
    return (bytes + g - 1) / g * g;
}
static void* cache_take(MemCache& c, size_t cls)
{
    std::lock_guard<std::mutex> lk(c.mu);
    auto it = c.free_.find(cls);
    if (it == c.free_.end()) return nullptr;
    void* p = it->second;
    c.free_.erase(it);
    c.bytes -= cls;
    return p;
}
static bool cache_put(MemCache& c, void* p, size_t cls)
{
    std::lock_guard<std::mutex> lk(c.mu);
    if (c.bytes + cls > cache_cap_bytes()) return false;
    c.free_.emplace(cls, p);
    c.bytes += cls;
    return true;
}
static void cache_drain(MemCache& c, bool pinned)
{
    std::lock_guard<std::mutex> lk(c.mu);
    for (auto& kv : c.free_) { if (pinned) cudaFreeHost(kv.second); else cudaFree(kv.second); }
    c.free_.clear();
    c.bytes = 0;
}
// device memory of the CURRENT device
static cudaError_t dev_alloc(int device, size_t bytes, Block* out)
{
    const size_t cls = size_class(bytes);
    void* p = (device >= 0 && device < MAX_DEVICES) ? cache_take(g_dev_cache[device], cls) : nullptr;
    if (!p) {
        cudaError_t e = cudaMalloc(&p, cls);
        if (e != cudaSuccess && device >= 0 && device < MAX_DEVICES) {   // out of memory: give the parked blocks back and retry
            cudaGetLastError();
            cache_drain(g_dev_cache[device], false);
            e = cudaMalloc(&p, cls);
        }
        if (e != cudaSuccess) { out->p = nullptr; out->bytes = 0; return e; }
    }
    out->p = p; out->bytes = cls;
    return cudaSuccess;
}
static void dev_free(int device, Block b)
{
    if (!b.p) return;
    if (device < 0 || device >= MAX_DEVICES || !cache_put(g_dev_cache[device], b.p, b.bytes)) cudaFree(b.p);
}
static cudaError_t pin_alloc(size_t bytes, Block* out)
{
    const size_t cls = size_class(bytes);
    void* p = cache_take(g_pin_cache, cls);
    if (!p) {
        cudaError_t e = cudaHostAlloc(&p, cls, cudaHostAllocPortable);
        if (e != cudaSuccess) { out->p = nullptr; out->bytes = 0; return e; }
    }
    out->p = p; out->bytes = cls;
    return cudaSuccess;
}
static void pin_free(Block b)
{
    if (b.p && !cache_put(g_pin_cache, b.p, b.bytes)) cudaFreeHost(b.p);
}

template <typename Tp>
static int dmalloc(bcfgpu_handle* h, Tp** p, size_t count, bool run_scope = false)
{
    Block b;
    const size_t bytes = std::max<size_t>(count * sizeof(Tp), 16);
    cudaError_t e = dev_alloc(h->device, bytes, &b);
    if (e != cudaSuccess) {
        *p = nullptr;
        cudaGetLastError();
        return fail(BCFGPU_ERR_OOM, std::string("cudaMalloc of ") + std::to_string(bytes) + " bytes: " + cudaGetErrorString(e));
    }
    (run_scope ? h->run_allocs : h->allocs).push_back(b);
    *p = (Tp*)b.p;
    return BCFGPU_OK;
}
#define DM(...) do { int r_ = dmalloc(__VA_ARGS__); if (r_) return r_; } while (0)

// (the owner has synchronised its stream: nothing in flight touches these blocks)
static void free_list(bcfgpu_handle* h, std::vector<Block>& v) { for (Block b : v) dev_free(h->device, b); v.clear(); }
static void drop_graph(bcfgpu_handle* h)
{
    if (h->gexec) { cudaGraphExecDestroy(h->gexec); h->gexec = nullptr; }
    if (h->graph) { cudaGraphDestroy(h->graph); h->graph = nullptr; }
}

static int use_device(bcfgpu_handle* h)
{
    CK(cudaSetDevice(h->device));
    return BCFGPU_OK;
}
#define USE(h) do { if (!(h)) return fail(BCFGPU_ERR_BAD_ARG, "null handle"); int r_ = use_device(h); if (r_) return r_; } while (0)
#define NEED_DATA(h) do { if (!(h)->has_data) return fail(BCFGPU_ERR_STATE, "set_data has not been called"); } while (0)

// ------------------------------------------------------------------------------------------------
// lifecycle
// ------------------------------------------------------------------------------------------------
extern "C" {

int bcfgpu_abi_version(void) { return BCFGPU_ABI_VERSION; }

int bcfgpu_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

const char* bcfgpu_last_error(void) { return g_err.c_str(); }

void bcfgpu_config_init(bcfgpu_config* c)
{
    if (!c) return;
    memset(c, 0, sizeof(*c));
    c->struct_size = (int32_t)sizeof(bcfgpu_config);
    c->ntree_con = 200; c->ntree_mod = 50;
    c->alpha_con = 0.95; c->beta_con = 2.0; c->alpha_mod = 0.25; c->beta_mod = 3.0;
    c->con_sd = 2.0; c->mod_sd = 1.0 / 0.674;
    c->nu = 3.0; c->lambda = 1.0; c->sigma_init = 1.0;
    c->use_muscale = 1; c->use_tauscale = 1;
    c->min_leaf = 5; c->max_leaves = 0;
    c->seed = 42; c->chain = 0; c->device = -1; c->engine = BCFGPU_ENGINE_AUTO; c->keep_draws = 0;
    c->model = BCFGPU_MODEL_BCF; c->treatment_col = -1; c->probit_offset = 0.0;
}

int bcfgpu_release_cached_memory(void)
{
    int cur = 0;
    const bool have = cudaGetDevice(&cur) == cudaSuccess;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess) { cudaGetLastError(); ndev = 0; }
    for (int dv = 0; dv < std::min(ndev, MAX_DEVICES); ++dv) {
        { std::lock_guard<std::mutex> lk(g_dev_cache[dv].mu); if (g_dev_cache[dv].free_.empty()) continue; }
        cudaSetDevice(dv);
        cache_drain(g_dev_cache[dv], false);
    }
    cache_drain(g_pin_cache, true);
    if (have) cudaSetDevice(cur);
    return BCFGPU_OK;
}
int bcfgpu_cache_stats(int64_t out2[2])
{
    if (!out2) return fail(BCFGPU_ERR_BAD_ARG, "null output");
    out2[0] = out2[1] = 0;
    for (int dv = 0; dv < MAX_DEVICES; ++dv) { std::lock_guard<std::mutex> lk(g_dev_cache[dv].mu); out2[0] += (int64_t)g_dev_cache[dv].bytes; }
    { std::lock_guard<std::mutex> lk(g_pin_cache.mu); out2[1] = (int64_t)g_pin_cache.bytes; }
    return BCFGPU_OK;
}

// The 64 mapped pinned bytes a handle's kernels leave their give-up note in: pinned allocations cost about a millisecond
// (a twentieth of a 20-sweep fit), so handles that are destroyed hand theirs to the next one (a few words per process).
static std::mutex g_dbg_mu;
static std::vector<int*> g_dbg_free;
static cudaError_t dbg_words_take(int** out)
{
    {
        std::lock_guard<std::mutex> lk(g_dbg_mu);
        if (!g_dbg_free.empty()) { *out = g_dbg_free.back(); g_dbg_free.pop_back(); return cudaSuccess; }
    }
    return cudaHostAlloc((void**)out, 64, cudaHostAllocMapped | cudaHostAllocPortable);
}
static void dbg_words_give(int* p)
{
    std::lock_guard<std::mutex> lk(g_dbg_mu);
    if (g_dbg_free.size() < 64) g_dbg_free.push_back(p); else cudaFreeHost(p);
}

int bcfgpu_create(const bcfgpu_config* cfg, bcfgpu_handle** out)
{
    if (!cfg || !out) return fail(BCFGPU_ERR_BAD_ARG, "null argument");
    *out = nullptr;
    if (cfg->struct_size != (int32_t)sizeof(bcfgpu_config)) return fail(BCFGPU_ERR_BAD_ARG, "bcfgpu_config.struct_size mismatch (ABI)");
    if (cfg->ntree_con < 1 || cfg->ntree_mod < 0) return fail(BCFGPU_ERR_BAD_ARG, "ntree_con must be >= 1 and ntree_mod >= 0");
    if (cfg->ntree_con + cfg->ntree_mod > 65535) return fail(BCFGPU_ERR_BAD_ARG, "too many trees");
    if (!(cfg->alpha_con > 0 && cfg->alpha_con < 1 && cfg->alpha_mod > 0 && cfg->alpha_mod < 1))
        return fail(BCFGPU_ERR_BAD_ARG, "base_control / base_moderate must be in (0,1)");
    if (!(cfg->con_sd > 0 && cfg->mod_sd > 0 && cfg->nu > 0 && cfg->lambda > 0 && cfg->sigma_init > 0))
        return fail(BCFGPU_ERR_BAD_ARG, "con_sd, mod_sd, nu, lambda, sigma_init must be positive");
    if (cfg->max_leaves < 0 || cfg->max_leaves > MAXL || cfg->max_leaves == 1) return fail(BCFGPU_ERR_BAD_ARG, "max_leaves must be 0 or in [2,128]");
    if (cfg->min_leaf < 1) return fail(BCFGPU_ERR_BAD_ARG, "min_leaf must be >= 1");
    if (cfg->model != BCFGPU_MODEL_BCF && cfg->model != BCFGPU_MODEL_PROBIT_BART) return fail(BCFGPU_ERR_BAD_ARG, "unknown model");
    if (cfg->model == BCFGPU_MODEL_PROBIT_BART && (cfg->ntree_mod != 0 || cfg->use_muscale != 0))
        return fail(BCFGPU_ERR_BAD_ARG, "probit BART is one forest without scales: set ntree_mod = 0 and use_muscale = 0");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return fail(BCFGPU_ERR_CUDA, "no CUDA device available (libbcfgpu has no CPU fallback)");
    }
    int dev = cfg->device;
    if (dev < 0) CK(cudaGetDevice(&dev));
    if (dev >= ndev) return fail(BCFGPU_ERR_BAD_ARG, "device ordinal out of range");
    bcfgpu_handle* h = new bcfgpu_handle();
    h->cfg = *cfg;
    if (h->cfg.max_leaves == 0) h->cfg.max_leaves = MAXL;
    h->device = dev;
    h->T = cfg->ntree_con + cfg->ntree_mod;
    memset(&h->dev, 0, sizeof(Dev));
    memset(&h->graph_dev, 0, sizeof(Dev));
    cudaError_t e2 = cudaSetDevice(dev);
    if (e2 == cudaSuccess) e2 = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking);
    if (e2 == cudaSuccess) e2 = cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking);
    for (int b = 0; b < 2; ++b) {
        if (e2 == cudaSuccess) e2 = cudaEventCreateWithFlags(&h->ev_copied[b], cudaEventDisableTiming);
        if (e2 == cudaSuccess) e2 = cudaEventCreateWithFlags(&h->ev_binned[b], cudaEventDisableTiming);
    }
    if (e2 == cudaSuccess) e2 = dbg_words_take(&h->dbg_host);
    if (e2 == cudaSuccess) memset(h->dbg_host, 0, 64);
    if (e2 == cudaSuccess) e2 = cudaEventCreate(&h->ev0);
    if (e2 == cudaSuccess) e2 = cudaEventCreate(&h->ev1);
    if (e2 == cudaSuccess) e2 = cudaDeviceGetAttribute(&h->num_sms, cudaDevAttrMultiProcessorCount, dev);
    if (e2 != cudaSuccess) { delete h; return fail(BCFGPU_ERR_CUDA, std::string("device setup: ") + cudaGetErrorString(e2)); }
    if (e2 == cudaSuccess) e2 = cudaFuncSetAttribute(k_propose_all, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)STEP_SMEM_BYTES);
    if (e2 == cudaSuccess) e2 = cudaFuncSetAttribute(k_tree_step, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)STEP_SMEM_BYTES);
    if (e2 == cudaSuccess) e2 = cudaFuncSetAttribute(k_sweep_end, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)STEP_SMEM_BYTES);
    if (e2 == cudaSuccess) e2 = cudaFuncSetAttribute(k_persistent, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)STEP_SMEM_BYTES);
    if (e2 == cudaSuccess) e2 = cudaFuncSetAttribute(k_batched<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)BATCH_SMEM_BYTES);
    // The kernels that keep an SM's observations in shared memory: the attribute is per function and per DEVICE, not per
    // handle, so it is raised once to the device's opt-in maximum (handles with different n on one GPU would otherwise
    // lower each other's limit between select_kernel() and the launch); select_kernel() only checks its own need.
    if (e2 == cudaSuccess) e2 = cudaDeviceGetAttribute(&h->max_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    for (const void* fn : {(const void*)k_batched<true>, (const void*)k_pipe<false>, (const void*)k_pipe<true>, (const void*)k_resident, (const void*)k_hist_mma}) {
        cudaFuncAttributes fa;
        if (e2 == cudaSuccess) e2 = cudaFuncGetAttributes(&fa, fn);
        if (e2 == cudaSuccess) e2 = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, h->max_optin - (int)fa.sharedSizeBytes);
    }
    if (e2 != cudaSuccess) { delete h; return fail(BCFGPU_ERR_CUDA, std::string("kernel attribute setup: ") + cudaGetErrorString(e2)); }
    if (cfg->engine < BCFGPU_ENGINE_AUTO || cfg->engine > BCFGPU_ENGINE_PIPELINED) { delete h; return fail(BCFGPU_ERR_BAD_ARG, "unknown engine"); }
    h->engine = cfg->engine == BCFGPU_ENGINE_AUTO ? BCFGPU_ENGINE_PIPELINED : cfg->engine;
    *out = h;
    return BCFGPU_OK;
}

static void release_local_group(LocalGroup* g);
int bcfgpu_destroy(bcfgpu_handle* h)
{
    if (!h) return BCFGPU_OK;
    cudaSetDevice(h->device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    drop_graph(h);
    free_list(h, h->allocs);
    free_list(h, h->run_allocs);
    if (!h->local) for (int p = 0; p < MAX_RANKS; ++p) if (h->peer_mail[p]) cudaIpcCloseMemHandle(h->peer_mail[p]);
    if (h->d_mail) cudaFree(h->d_mail);
    if (h->local) release_local_group(h->local);
    if (h->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(h->comm);
    for (int b = 0; b < 2; ++b) { if (h->ev_copied[b]) cudaEventDestroy(h->ev_copied[b]); if (h->ev_binned[b]) cudaEventDestroy(h->ev_binned[b]); }
    if (h->copy_stream) cudaStreamDestroy(h->copy_stream);
    if (h->dbg_host) dbg_words_give(h->dbg_host);
    if (h->ev0) cudaEventDestroy(h->ev0);
    if (h->ev1) cudaEventDestroy(h->ev1);
    if (h->stream) cudaStreamDestroy(h->stream);
    delete h;
    return BCFGPU_OK;
}

// ------------------------------------------------------------------------------------------------
// sharding
// ------------------------------------------------------------------------------------------------
// Map every peer's mailbox array into this process (cudaIpc over NVLink / NVSwitch).  Collective over the ranks; if any
// rank cannot map any peer, all ranks agree to keep the NCCL all-reduce path.
static int setup_peers(bcfgpu_handle* h)
{
    h->peer_ok = false;
    if (h->nranks > MAX_RANKS || !g_nccl.AllGather || getenv("BCFGPU_NO_PEER")) return BCFGPU_OK;
    CK(cudaMalloc((void**)&h->d_mail, sizeof(Mailbox) * h->nranks));
    CK(cudaMemsetAsync(h->d_mail, 0, sizeof(Mailbox) * h->nranks, h->stream));   // (ordered on the handle's stream, see set_shard_local)
    CK(cudaStreamSynchronize(h->stream));
    cudaIpcMemHandle_t mine;
    CK(cudaIpcGetMemHandle(&mine, h->d_mail));
    char *d_send = nullptr, *d_recv = nullptr;
    long long* d_flag = nullptr;
    CK(cudaMalloc((void**)&d_send, sizeof(mine)));
    CK(cudaMalloc((void**)&d_recv, sizeof(mine) * h->nranks));
    CK(cudaMalloc((void**)&d_flag, 8));
    CK(cudaMemcpy(d_send, &mine, sizeof(mine), cudaMemcpyHostToDevice));
    NK(g_nccl.AllGather(d_send, d_recv, sizeof(mine), NCCL_INT8, h->comm, h->stream));
    std::vector<cudaIpcMemHandle_t> all((size_t)h->nranks);
    CK(cudaMemcpyAsync(all.data(), d_recv, sizeof(mine) * h->nranks, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    long long ok = 1;
    for (int p = 0; p < h->nranks && ok; ++p) {
        if (p == h->rank) continue;
        void* q = nullptr;
        if (cudaIpcOpenMemHandle(&q, all[p], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); ok = 0; }
        else h->peer_mail[p] = (Mailbox*)q;
    }
    CK(cudaMemcpy(d_flag, &ok, 8, cudaMemcpyHostToDevice));
    NK(g_nccl.AllReduce(d_flag, d_flag, 1, NCCL_INT64, NCCL_MIN, h->comm, h->stream));
    CK(cudaMemcpyAsync(&ok, d_flag, 8, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    cudaFree(d_send); cudaFree(d_recv); cudaFree(d_flag);
    h->peer_ok = ok != 0;
    return BCFGPU_OK;
}

int bcfgpu_nccl_unique_id(char id_out[128])
{
    if (!id_out) return fail(BCFGPU_ERR_BAD_ARG, "null argument");
    int r = nccl_load();
    if (r) return r;
    NcclId id;
    NK(g_nccl.GetUniqueId(&id));
    memcpy(id_out, id.b, 128);
    return BCFGPU_OK;
}

int bcfgpu_set_shard(bcfgpu_handle* h, int rank, int nranks, const char nccl_unique_id[128])
{
    USE(h);
    if (h->has_data) return fail(BCFGPU_ERR_STATE, "set_shard must be called before set_data");
    if (nranks < 1 || rank < 0 || rank >= nranks) return fail(BCFGPU_ERR_BAD_ARG, "bad rank / nranks");
    h->rank = rank; h->nranks = nranks;
    if (nranks == 1) return BCFGPU_OK;
    if (!nccl_unique_id) return fail(BCFGPU_ERR_BAD_ARG, "null nccl_unique_id");
    int r = nccl_load();
    if (r) return r;
    NcclId id;
    memcpy(id.b, nccl_unique_id, 128);
    NK(g_nccl.CommInitRank(&h->comm, nranks, id, rank));
    return setup_peers(h);
}

}  // extern "C"

// Shards that live in ONE process (bcfgpu_set_shard_local: one handle per shard, each driven by its own host thread;
// the shards may even share a GPU, each on a share of its SMs): the few host-side reductions go through this object
// instead of NCCL, the in-kernel exchange through mailboxes the handles reach directly (same address space).
struct LocalGroup {
    std::mutex mu;
    std::condition_variable cv;
    int n = 0, arrived = 0, refs = 0;
    long long gen = 0;
    std::vector<long long> acc, res;
};
static int local_allreduce(bcfgpu_handle* h, long long* dbuf, size_t count, bool take_min)
{
    LocalGroup* g = h->local;
    std::vector<long long> v(count);
    CK(cudaMemcpyAsync(v.data(), dbuf, count * 8, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    {
        std::unique_lock<std::mutex> lk(g->mu);
        if (g->arrived == 0) g->acc = v;
        else for (size_t i = 0; i < count; ++i) g->acc[i] = take_min ? std::min(g->acc[i], v[i]) : g->acc[i] + v[i];
        if (++g->arrived == g->n) { g->res = g->acc; g->arrived = 0; ++g->gen; g->cv.notify_all(); }
        else { const long long my = g->gen; g->cv.wait(lk, [&] { return g->gen != my; }); }
        v = g->res;
    }
    CK(cudaMemcpyAsync(dbuf, v.data(), count * 8, cudaMemcpyHostToDevice, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return BCFGPU_OK;
}
static void release_local_group(LocalGroup* g)
{
    bool last;
    { std::lock_guard<std::mutex> lk(g->mu); last = --g->refs == 0; }
    if (last) delete g;
}
// all-reduce of int64 words across shards (sum, or min); no-op on one rank
static int shard_allreduce(bcfgpu_handle* h, long long* buf, size_t count, bool take_min = false)
{
    if (h->nranks <= 1) return BCFGPU_OK;
    if (h->local) return local_allreduce(h, buf, count, take_min);
    NK(g_nccl.AllReduce(buf, buf, count, NCCL_INT64, take_min ? NCCL_MIN : NCCL_SUM, h->comm, h->stream));
    return BCFGPU_OK;
}

extern "C" int bcfgpu_set_shard_local(bcfgpu_handle** hs, int nranks)
{
    if (!hs || nranks < 1 || nranks > MAX_RANKS) return fail(BCFGPU_ERR_BAD_ARG, "need 1 .. 8 handles");
    for (int r = 0; r < nranks; ++r) {
        if (!hs[r]) return fail(BCFGPU_ERR_BAD_ARG, "null handle");
        if (hs[r]->has_data || hs[r]->nranks > 1) return fail(BCFGPU_ERR_STATE, "set_shard_local must come before set_data, once");
        if (hs[r]->cfg.engine == BCFGPU_ENGINE_GRAPH) return fail(BCFGPU_ERR_BAD_ARG, "in-process shards need a sweep kernel with the in-kernel exchange (not ENGINE_GRAPH)");
    }
    if (nranks == 1) { hs[0]->rank = 0; hs[0]->nranks = 1; return BCFGPU_OK; }
    LocalGroup* g = new LocalGroup();
    g->n = nranks; g->refs = nranks;
    int cur = 0;
    cudaGetDevice(&cur);
    for (int r = 0; r < nranks; ++r) {
        bcfgpu_handle* h = hs[r];
        // Shards that share a device run side by side, each spinning on the others' words: together they must leave a few
        // SMs free, or the small kernels (memsets, permutations) of a shard that is not in its sweep kernel yet would wait
        // for an SM that never comes.
        int same = 0;
        for (int q = 0; q < nranks; ++q) same += hs[q]->device == h->device ? 1 : 0;
        if (same > 1) {
            const int share = std::max(1, (h->num_sms - 4) / same);
            h->cfg.max_ctas = h->cfg.max_ctas > 0 ? std::min(h->cfg.max_ctas, share) : share;
        }
        // Observed on B200: a third pipelined sweep kernel (1024-thread CTAs, the whole register file and ~200 KB of shared
        // memory per CTA) is not scheduled while two others spin on the same device, whatever their grid sizes; four batched
        // kernels are.  Three or more shards of one device therefore run the batched kernel.
        h->no_pipe = same > 2;
        CK(cudaSetDevice(h->device));
        CK(cudaMalloc((void**)&h->d_mail, sizeof(Mailbox) * nranks));
        // (on the handle's own stream and waited for: a memset on the legacy stream is not ordered before the sweep kernels
        // of the non-blocking streams, and one that ran late would wipe words a peer has already pushed)
        CK(cudaMemsetAsync(h->d_mail, 0, sizeof(Mailbox) * nranks, h->stream));
        CK(cudaStreamSynchronize(h->stream));
        CK(cudaDeviceSynchronize());
        h->rank = r; h->nranks = nranks; h->local = g;
    }
    for (int r = 0; r < nranks; ++r) {
        CK(cudaSetDevice(hs[r]->device));
        for (int p = 0; p < nranks; ++p) {
            if (p == r) continue;
            if (hs[p]->device != hs[r]->device) {
                int can = 0;
                CK(cudaDeviceCanAccessPeer(&can, hs[r]->device, hs[p]->device));
                if (!can) return fail(BCFGPU_ERR_CUDA, "no peer access between the shards' devices");
                cudaError_t e = cudaDeviceEnablePeerAccess(hs[p]->device, 0);
                if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) return fail(BCFGPU_ERR_CUDA, cudaGetErrorString(e));
                cudaGetLastError();
            }
            hs[r]->peer_mail[p] = hs[p]->d_mail;
        }
        hs[r]->peer_ok = true;
    }
    cudaSetDevice(cur);
    return BCFGPU_OK;
}

// ------------------------------------------------------------------------------------------------
// data
// ------------------------------------------------------------------------------------------------
// share != nullptr: this forest's covariates are the leading columns of that forest's -- same binned columns
static int setup_forest(bcfgpu_handle* h, int f, int p, const double* cuts, const int32_t* off, const ForestHost* share = nullptr)
{
    ForestHost& F = h->fh[f];
    F = ForestHost();
    F.p = p;
    if (p == 0) return BCFGPU_OK;
    if (off[0] != 0) return fail(BCFGPU_ERR_BAD_ARG, "cut_off[0] must be 0");
    F.off.assign(off, off + p + 1);
    F.ncut.resize(p); F.col_index.resize(p); F.is16.resize(p);
    for (int v = 0; v < p; ++v) {
        const int nc = off[v + 1] - off[v];
        if (nc < 1) return fail(BCFGPU_ERR_BAD_ARG, "every covariate needs at least one cutpoint");
        if (nc > 65535) return fail(BCFGPU_ERR_BAD_ARG, "at most 65535 cutpoints per covariate");
        for (int c = 1; c < nc; ++c)
            if (!(cuts[off[v] + c] >= cuts[off[v] + c - 1])) return fail(BCFGPU_ERR_BAD_ARG, "cutpoints must be sorted and finite");
        F.ncut[v] = nc;
        F.is16[v] = nc > 255;
        F.col_index[v] = F.is16[v] ? F.p16++ : F.p8++;
    }
    F.cuts.assign(cuts, cuts + off[p]);
    DM(h, &F.d_cuts, F.cuts.size());
    DM(h, &F.d_off, F.off.size());
    DM(h, &F.d_ncut, (size_t)p);
    DM(h, &F.d_col_index, (size_t)p);
    DM(h, &F.d_is16, (size_t)p);
    if (share) { F.d_bins8 = share->d_bins8; F.d_bins16 = share->d_bins16; }
    else {
        DM(h, &F.d_bins8, (size_t)std::max(F.p8, 1) * h->n_pad);
        DM(h, &F.d_bins16, (size_t)std::max(F.p16, 1) * h->n_pad);
    }
    CK(cudaMemcpyAsync(F.d_cuts, F.cuts.data(), F.cuts.size() * 8, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(F.d_off, F.off.data(), F.off.size() * 4, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(F.d_ncut, F.ncut.data(), (size_t)p * 4, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(F.d_col_index, F.col_index.data(), (size_t)p * 4, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(F.d_is16, F.is16.data(), (size_t)p, cudaMemcpyHostToDevice, h->stream));
    h->h2d_bytes += (int64_t)(F.cuts.size() * 8 + F.off.size() * 4 + (size_t)p * 9);
    if (!share) {
        CK(cudaMemsetAsync(F.d_bins8, 0, (size_t)std::max(F.p8, 1) * h->n_pad, h->stream));
        CK(cudaMemsetAsync(F.d_bins16, 0, (size_t)std::max(F.p16, 1) * h->n_pad * 2, h->stream));
    }
    return BCFGPU_OK;
}

// ------------------------------------------------------------------------------------------------
// host-side helpers of set_data: a few threads over row ranges (the R process calling us is single-threaded)
// ------------------------------------------------------------------------------------------------
static int host_threads()
{
    static const int nt = [] {
        if (const char* e = getenv("BCFGPU_HOST_THREADS")) return std::max(1, std::min(64, atoi(e)));
        return (int)std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
    }();
    return nt;
}
// f(part, lo, hi) over [0, n) cut into `parts` ranges whose bounds are multiples of `align`
template <class F>
static void parallel_ranges(int64_t n, int64_t align, int parts, F f)
{
    const int64_t units = (n + align - 1) / align;
    parts = (int)std::max<int64_t>(1, std::min<int64_t>(parts, units));
    auto bound = [&](int t) { return std::min<int64_t>(n, units * t / parts * align); };
    std::vector<std::thread> th;
    for (int t = 1; t < parts; ++t) th.emplace_back([&, t] { f(t, bound(t), bound(t + 1)); });
    f(0, bound(0), bound(1));
    for (auto& x : th) x.join();
}

// ------------------------------------------------------------------------------------------------
// host pre-pass of bcf's R wrapper over the covariates (no device needed)
// ------------------------------------------------------------------------------------------------
extern "C" int bcfgpu_scan_columns(const double* X, int64_t n, int32_t p, int64_t ld, double* lo, double* hi, uint8_t* flags)
{
    if (!X || !lo || !hi || !flags || n < 1 || p < 1 || ld < n) return fail(BCFGPU_ERR_BAD_ARG, "bad argument");
    std::atomic<int> next{0};
    auto work = [&] { for (int v; (v = next.fetch_add(1)) < p;) scan_column(X + (int64_t)v * ld, n, lo + v, hi + v, flags + v); };
    const int nth = (int)std::max<int64_t>(1, std::min<int64_t>({(int64_t)host_threads(), (int64_t)p, n * (int64_t)p / 65536 + 1}));
    std::vector<std::thread> th;
    for (int t = 1; t < nth; ++t) th.emplace_back(work);
    work();
    for (auto& t : th) t.join();
    return BCFGPU_OK;
}

constexpr int PLACE_CHUNK = 4096;                 // rows per CTA of k_place (the host hands over one prefix count per chunk)
constexpr size_t STAGE_BYTES = (size_t)64 << 20;  // fp64 staging buffer of K0, two of them: copy of chunk k+1 under the binning of chunk k

// K0 over a column-major host matrix: column chunks cross the bus on the copy stream (straight from the caller's buffer
// when it is pinned; through two pinned staging buffers filled by host threads when it is pageable, as R's memory is)
// while the previous chunk is binned on the compute stream.  No n x p fp64 copy ever exists on the device.
// Runs of ONE-cutpoint columns (the reference's two-valued covariates) cross as BYTES: the host threads that fill the pinned
// staging buffer write (cut <= x) instead of copying 8 bytes -- the same pass over the caller's memory, an eighth of the
// traffic over PCIe (which is what bounds set_data: 820 MB at 50 GB/s at n = 1e6, p = 100) -- and k_bin only places them.
// BCFGPU_NO_HOST_PACK=1 sends every column as doubles (the tests run both ways); BCFGPU_HOST_PACK=1 skips the probe below.
// o[i] = (cut <= x[i]) for i in [i0, i1): the staging pass of a one-cutpoint column (memory-bound; the AVX2 clone keeps a
// single thread near its streaming bandwidth)
__attribute__((target_clones("avx2", "default")))
static void pack_column(const double* __restrict__ xs, double cut, int64_t i0, int64_t i1, uint8_t* __restrict__ o)
{
    for (int64_t i = i0; i < i1; ++i) o[i] = (uint8_t)(cut <= xs[i]);
}
static int bin_columns(bcfgpu_handle* h, ForestHost& F, const double* x, int v0, int v1, bool as_bytes, bool pinned, double* pack_s = nullptr)
{
    const int64_t n = h->n;
    const size_t es = as_bytes ? 1 : 8, colb = (size_t)n * es;
    const int ncol = v1 - v0;
    // (bytes: smaller chunks -- the pinned staging buffers are allocated on a process's first fit, and the host pass of the next
    // chunk runs under the copy and the placement of this one)
    const int chunk = (int)std::max<size_t>(1, std::min<size_t>((size_t)ncol, (as_bytes ? STAGE_BYTES / 2 : STAGE_BYTES) / colb));
    const size_t stage_bytes = (size_t)chunk * colb;
    const bool staged = as_bytes || !pinned;
    const int nbuf = (ncol > chunk) ? 2 : 1;
    Block stage[2] = {{nullptr, 0}, {nullptr, 0}}, pin[2] = {{nullptr, 0}, {nullptr, 0}};
    int rc = BCFGPU_OK;
    for (int b = 0; b < nbuf && rc == BCFGPU_OK; ++b) {
        if (dev_alloc(h->device, stage_bytes, &stage[b]) != cudaSuccess) rc = fail(BCFGPU_ERR_OOM, "staging buffer of the binning kernel");
        else if (staged && pin_alloc(stage_bytes, &pin[b]) != cudaSuccess) rc = fail(BCFGPU_ERR_OOM, "pinned staging buffer");
    }
    cudaError_t e = cudaSuccess;
    for (int c0 = v0, k = 0; c0 < v1 && rc == BCFGPU_OK && e == cudaSuccess; c0 += chunk, ++k) {
        const int b = k & (nbuf - 1);
        const int nc = std::min(chunk, v1 - c0);
        const size_t bytes = (size_t)nc * colb;
        const double* src = x + (size_t)c0 * n;
        const void* from = src;
        if (k >= nbuf) e = cudaStreamWaitEvent(h->copy_stream, h->ev_binned[b], 0);       // the staging buffer is free again
        if (staged && e == cudaSuccess) {
            if (k >= nbuf) e = cudaEventSynchronize(h->ev_copied[b]);                     // ... and so is the pinned one
            if (as_bytes) {
                uint8_t* dst = (uint8_t*)pin[b].p;
                const double* cuts = F.cuts.data();
                const int32_t* off = F.off.data();
                const auto tp0 = std::chrono::steady_clock::now();
                parallel_ranges((int64_t)nc * n, 1 << 14, host_threads(), [&](int, int64_t lo, int64_t hi) {
                    while (lo < hi) {                                                     // (a range may span columns)
                        const int64_t c = lo / n, i0 = lo - c * n, i1 = std::min<int64_t>(n, i0 + (hi - lo));
                        const double cut = cuts[off[c0 + c]];
                        pack_column(src + c * n, cut, i0, i1, dst + c * n);
                        lo += i1 - i0;
                    }
                });
                if (pack_s) *pack_s += std::chrono::duration<double>(std::chrono::steady_clock::now() - tp0).count();
            } else {
                char* dst = (char*)pin[b].p;
                parallel_ranges((int64_t)bytes, 1 << 16, host_threads(), [&](int, int64_t lo, int64_t hi) { memcpy(dst + lo, (const char*)src + lo, (size_t)(hi - lo)); });
            }
            from = pin[b].p;
        }
        if (e == cudaSuccess) e = cudaMemcpyAsync(stage[b].p, from, bytes, cudaMemcpyHostToDevice, h->copy_stream);
        if (e == cudaSuccess) e = cudaEventRecord(h->ev_copied[b], h->copy_stream);
        if (e == cudaSuccess) e = cudaStreamWaitEvent(h->stream, h->ev_copied[b], 0);
        if (e != cudaSuccess) break;
        h->h2d_bytes += (int64_t)bytes;
        dim3 grid((unsigned)((n + BIN_ROWS - 1) / BIN_ROWS), (unsigned)((nc + BIN_COLS - 1) / BIN_COLS));
        const int trt_pad = (int)((int64_t)h->dev.trt_tiles * TILE);
        if (as_bytes)
            k_bin<uint8_t><<<grid, 256, 0, h->stream>>>((const uint8_t*)stage[b].p, n, nc, c0, F.d_cuts, F.d_off, F.d_col_index, F.d_is16, h->d_pos,
                                                         h->n_pad, trt_pad, F.d_bins8, F.d_bins16);
        else
            k_bin<double><<<grid, 256, 0, h->stream>>>((const double*)stage[b].p, n, nc, c0, F.d_cuts, F.d_off, F.d_col_index, F.d_is16, h->d_pos,
                                                        h->n_pad, trt_pad, F.d_bins8, F.d_bins16);
        h->launches++;
        e = cudaGetLastError();
        if (e == cudaSuccess) e = cudaEventRecord(h->ev_binned[b], h->stream);
    }
    if (rc == BCFGPU_OK && e != cudaSuccess) rc = fail(BCFGPU_ERR_CUDA, std::string("H2D / binning of X: ") + cudaGetErrorString(e));
    // the buffers go back to the cache only when nothing in flight uses them (and the caller's X may be freed on return)
    cudaStreamSynchronize(h->copy_stream);
    e = cudaStreamSynchronize(h->stream);
    if (rc == BCFGPU_OK && e != cudaSuccess) rc = fail(BCFGPU_ERR_CUDA, std::string("k_bin: ") + cudaGetErrorString(e));
    for (int b = 0; b < 2; ++b) { dev_free(h->device, stage[b]); pin_free(pin[b]); }
    return rc;
}
constexpr double HOST_PACK_MIN_RATE = 36e9;   // bytes / s of the probe below which pinned one-cutpoint columns cross by DMA as doubles
static int bin_forest(bcfgpu_handle* h, int f, const double* x)
{
    ForestHost& F = h->fh[f];
    if (F.p == 0) return BCFGPU_OK;
    cudaPointerAttributes pa;
    bool pinned = cudaPointerGetAttributes(&pa, x) == cudaSuccess && (pa.type == cudaMemoryTypeHost || pa.type == cudaMemoryTypeManaged);
    cudaGetLastError();
    if (getenv("BCFGPU_FORCE_PAGEABLE")) pinned = false;
    const bool pack = !getenv("BCFGPU_NO_HOST_PACK");
    auto byteable = [&](int v) { return pack && F.ncut[v] == 1 && !F.is16[v]; };
    for (int v0 = 0; v0 < F.p;) {                      // maximal runs of columns of one kind
        const bool bts = byteable(v0);
        int v1 = v0 + 1;
        while (v1 < F.p && byteable(v1) == bts) ++v1;
        int rc;
        if (bts && pinned && !getenv("BCFGPU_HOST_PACK") && v1 - v0 > 8) {
            // Pinned input could also cross by DMA without a host pass (50 GB/s over PCIe gen5): the byte path only pays when
            // the host threads read faster than that -- they do on an idle box (85 GB/s on 16 cores), they may not when
            // several ranks share few cores.  The first columns are the probe; it reads low (thread wake-up inside 64 MB:
            // a box whose probe said 47-52 GB/s moved X in 11.8 ms as bytes, 16.5 ms as doubles), hence the margin.
            double pack_s = 0.0;
            rc = bin_columns(h, F, x, v0, v0 + 8, true, pinned, &pack_s);
            const double rate = 8.0 * (double)h->n * 8.0 / std::max(pack_s, 1e-9);
            if (getenv("BCFGPU_SETDATA_PROFILE")) fprintf(stderr, "[bcfgpu set_data] host pass of one-cutpoint columns: %.1f GB/s -> %s\n", rate * 1e-9, rate >= HOST_PACK_MIN_RATE ? "bytes" : "doubles by DMA");
            if (!rc) rc = bin_columns(h, F, x, v0 + 8, v1, rate >= HOST_PACK_MIN_RATE, pinned);
        } else
            rc = bin_columns(h, F, x, v0, v1, bts, pinned);
        if (rc) return rc;
        v0 = v1;
    }
    return BCFGPU_OK;
}

static void fill_forest_dev(const bcfgpu_handle* h, int f, ForestDev& D)
{
    const ForestHost& F = h->fh[f];
    D.p = F.p;
    D.ntree = f == 0 ? h->cfg.ntree_con : h->cfg.ntree_mod;
    D.tree0 = f == 0 ? 0 : h->cfg.ntree_con;
    D.is_con = f == 0;
    D.bins8 = F.d_bins8; D.bins16 = F.d_bins16; D.col_index = F.d_col_index; D.col_is16 = F.d_is16; D.ncut = F.d_ncut;
    const double a = f == 0 ? h->cfg.alpha_con : h->cfg.alpha_mod, b = f == 0 ? h->cfg.beta_con : h->cfg.beta_mod;
    for (int dpt = 0; dpt < 32; ++dpt) D.pg[dpt] = a / std::pow(1.0 + (double)dpt, b);
}

static void root_tree(TreeRec& tr, double mu, int32_t ntrt = 0, int32_t nctl = 0)
{
    memset(&tr, 0, sizeof(tr));
    tr.nnodes = 1; tr.nleaves = 1;
    for (int i = 0; i < MAXN; ++i) { tr.parent[i] = tr.left[i] = tr.right[i] = -1; tr.var[i] = tr.cut[i] = 0xFFFF; }
    tr.node_slot[0] = 0; tr.depth[0] = 0; tr.heap[0] = 1; tr.slot_node[0] = 0; tr.mu[0] = mu;
    tr.cnt1[0] = ntrt; tr.cnt0[0] = nctl;
}

static int build_graph(bcfgpu_handle* h);

extern "C" int bcfgpu_set_data(bcfgpu_handle* h, int64_t n, int32_t p_con, int32_t p_mod, const double* y,
                               const int32_t* z, const double* xcon, const double* xmod, const double* cuts_con,
                               const int32_t* cut_off_con, const double* cuts_mod, const int32_t* cut_off_mod)
{
    USE(h);
    const bcfgpu_config& c = h->cfg;
    const bool sdprof = getenv("BCFGPU_SETDATA_PROFILE") != nullptr;
    auto sd_t0 = std::chrono::steady_clock::now();
    auto sd_mark = [&](const char* what, bool sync = false) {
        if (!sdprof) return;
        if (sync) { cudaStreamSynchronize(h->copy_stream); cudaStreamSynchronize(h->stream); }
        const auto now = std::chrono::steady_clock::now();
        fprintf(stderr, "[bcfgpu set_data] %-34s %.2f ms\n", what, std::chrono::duration<double, std::milli>(now - sd_t0).count());
        sd_t0 = now;
    };
    if (n < 1 || n > 2000000000ll) return fail(BCFGPU_ERR_BAD_ARG, "n must be in [1, 2e9]");
    if (!y || !z || !xcon || !cuts_con || !cut_off_con) return fail(BCFGPU_ERR_BAD_ARG, "null data pointer");
    if (p_con < 1 || p_con > 65535) return fail(BCFGPU_ERR_BAD_ARG, "p_con must be in [1, 65535]");
    if (c.ntree_mod > 0 && (p_mod < 1 || p_mod > 65535 || !xmod || !cuts_mod || !cut_off_mod))
        return fail(BCFGPU_ERR_BAD_ARG, "moderate forest needs x_moderate and its cutpoints");
    if (c.ntree_mod == 0) p_mod = 0;
    // ---- one threaded pass over (y, z): validation, treated rows per chunk of PLACE_CHUNK rows (what the device needs
    // to apply bcf's perm = order(z, decreasing = TRUE) itself), and ybar's order-independent fixed-point sum, so that any
    // sharding starts from identical bits
    const int64_t nchunk = (n + PLACE_CHUNK - 1) / PLACE_CHUNK;
    std::vector<int32_t> chunk_trt((size_t)nchunk + 1, 0);
    const int nth = n >= 65536 ? host_threads() : 1;
    const bool probit01 = c.model == BCFGPU_MODEL_PROBIT_BART;
    struct Part { long long s[4] = {0, 0, 0, 0}; int bad = 0; char pad_[28]; };
    std::vector<Part> part((size_t)nth);
    parallel_ranges(n, PLACE_CHUNK, nth, [&](int t, int64_t lo, int64_t hi) {
        Part P;
        for (int64_t c0 = lo; c0 < hi; c0 += PLACE_CHUNK) {
            const int64_t c1 = std::min(hi, c0 + PLACE_CHUNK);
            int32_t k = 0;
            for (int64_t i = c0; i < c1; ++i) {
                const int32_t zi = z[i];
                const double yi = y[i];
                if (zi != 0 && zi != 1) P.bad |= 1;
                if (probit01 && yi != 0.0 && yi != 1.0) P.bad |= 8;
                if (!(std::fabs(yi) < 32768.0)) { P.bad |= std::isfinite(yi) ? 4 : 2; continue; }
                k += zi;
                const long long v = std::llrint(yi * 17592186044416.0);
                P.s[1] += v & 0x1FFFFF; P.s[2] += (v >> LIMB_BITS) & 0x1FFFFF; P.s[3] += v >> (2 * LIMB_BITS);
            }
            chunk_trt[(size_t)(c0 / PLACE_CHUNK) + 1] = k;
        }
        part[(size_t)t] = P;
    });
    long long ysum[5] = {(long long)n, 0, 0, 0, 0};   // count, limb0, limb1, limb2, treated count
    int badbits = 0;
    for (const Part& P : part) { badbits |= P.bad; ysum[1] += P.s[1]; ysum[2] += P.s[2]; ysum[3] += P.s[3]; }
    if (badbits & 1) return fail(BCFGPU_ERR_BAD_ARG, "z must be 0/1");
    if (badbits & 2) return fail(BCFGPU_ERR_BAD_ARG, "y must be finite");
    if (badbits & 4) return fail(BCFGPU_ERR_BAD_ARG, "|y| too large: pass the standardised outcome");
    if (badbits & 8) return fail(BCFGPU_ERR_BAD_ARG, "probit BART needs a 0/1 response");
    for (int64_t k = 0; k < nchunk; ++k) chunk_trt[(size_t)k + 1] += chunk_trt[(size_t)k];     // exclusive prefix: treated rows before chunk k
    const int64_t ntrt = chunk_trt[(size_t)nchunk];
    ysum[4] = ntrt;
    sd_mark("checks, counts, ybar (host threads)");
    CK(cudaStreamSynchronize(h->stream));
    drop_graph(h);
    free_list(h, h->allocs);
    free_list(h, h->run_allocs);
    h->has_data = false;
    h->h2d_bytes = 0;
    h->coop_grid = 0;
    h->shard_decided = false;
    h->have_pos = false;
    h->n = n;
    const int64_t trt_pad = (ntrt + TILE - 1) / TILE * TILE;
    const int64_t ctl_pad = (n - ntrt + TILE - 1) / TILE * TILE;
    h->ntrt = (int)ntrt;
    h->n_pad = trt_pad + ctl_pad;
    const int64_t n_pad = h->n_pad;
    Dev& d = h->dev;
    memset(&d, 0, sizeof(Dev));
    d.pr.rank = h->rank; d.pr.nranks = h->peer_ok ? h->nranks : 1; d.pr.seq0 = 0; d.pr.mine = h->d_mail;
    for (int p = 0; p < MAX_RANKS; ++p) d.pr.peer[p] = h->peer_mail[p];
    d.dbg = h->dbg_host;       // (unified addressing: the mapped host pointer is valid on the device)
    d.n_pad = n_pad; d.n_tiles = (int32_t)(n_pad / TILE); d.trt_tiles = (int32_t)(trt_pad / TILE);
    d.T = h->T; d.max_leaves = c.max_leaves; d.min_leaf = c.min_leaf;
    d.use_muscale = c.use_muscale; d.use_tauscale = c.use_tauscale;
    const bool probit = c.model == BCFGPU_MODEL_PROBIT_BART;
    d.fix_sigma = probit ? 1 : 0;
    if (probit) {
        if (h->nranks > 1) return fail(BCFGPU_ERR_BAD_ARG, "probit BART does not run sharded");
        if (c.treatment_col >= p_con) return fail(BCFGPU_ERR_BAD_ARG, "treatment_col is not a column of x_control");
        if (c.treatment_col >= 0 && cut_off_con[c.treatment_col + 1] - cut_off_con[c.treatment_col] != 1)
            return fail(BCFGPU_ERR_BAD_ARG, "the treatment column must be 0/1 with one cutpoint");
        ysum[1] = ysum[2] = ysum[3] = 0;       // the working response starts at 0: every leaf starts at 0
    }
    {
        const uint64_t s = c.seed + (uint64_t)c.chain * 0x9E3779B97F4A7C15ull;
        d.key0 = (uint32_t)s; d.key1 = (uint32_t)(s >> 32);
    }
    d.nu = c.nu; d.lambda = c.lambda; d.con_sd = c.con_sd; d.mod_sd = c.mod_sd;

    DM(h, &h->d_pos, (size_t)n);
    DM(h, &h->d_orig, (size_t)n_pad);
    double* dy = nullptr;
    DM(h, &dy, (size_t)n_pad);
    DM(h, &d.rfx, (size_t)n_pad); DM(h, &d.Gc, (size_t)n_pad); DM(h, &d.Gm, (size_t)n_pad);
    DM(h, &d.tau_sum, (size_t)n_pad); DM(h, &d.tau_sq, (size_t)n_pad); DM(h, &d.mu_sum, (size_t)n_pad); DM(h, &d.yhat_sum, (size_t)n_pad);
    DM(h, &h->d_tmp, (size_t)n_pad); DM(h, &h->d_tmp2, (size_t)n_pad);
    DM(h, &d.slots, (size_t)h->T * n_pad);
    DM(h, &d.trees[0], (size_t)h->T); DM(h, &d.trees[1], (size_t)h->T);
    DM(h, &d.proprec, (size_t)h->T);
    DM(h, &d.totals, (size_t)2 * h->T * KEYS * 8);
    DM(h, &d.cross, (size_t)2 * h->T * PIPE_XSTRIDE);
    DM(h, &d.mom, (size_t)2 * 2 * NMOM * 3);
    DM(h, &d.sc, 1);
    DM(h, &d.err, 4);
    DM(h, &d.trace, (size_t)h->T * 5);
    DM(h, &d.trace_logr, (size_t)h->T);
    DM(h, &d.counters, 4);
    DM(h, &h->d_bar, 64);
    DM(h, &h->d_done, 4);
    int32_t* d_chunk = nullptr;
    DM(h, &d_chunk, (size_t)nchunk + 1);
    if (getenv("BCFGPU_PROFILE")) { DM(h, &d.prof, 1024 + 512 * 8); CK(cudaMemsetAsync(d.prof, 0, (1024 + 512 * 8) * 8, h->stream)); }
    d.y = dy;
    sd_mark("device allocations");
    // ---- y, z and the chunk counts cross first (12 B per row), then the device permutes: pos, orig, y in internal order
    double* d_yraw = h->d_tmp;                         // scratch until the first fetch
    int32_t* d_zraw = (int32_t*)h->d_tmp2;
    CK(cudaMemcpyAsync(d_yraw, y, (size_t)n * 8, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(d_zraw, z, (size_t)n * 4, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(d_chunk, chunk_trt.data(), ((size_t)nchunk + 1) * 4, cudaMemcpyHostToDevice, h->stream));
    h->h2d_bytes += (int64_t)n * 12 + (nchunk + 1) * 4;
    CK(cudaMemsetAsync(h->d_orig, 0xFF, (size_t)n_pad * 4, h->stream));       // padding rows: orig = -1, y = 0
    CK(cudaMemsetAsync(dy, 0, (size_t)n_pad * 8, h->stream));
    k_place<<<(unsigned)nchunk, 256, 0, h->stream>>>(d_yraw, d_zraw, d_chunk, n, trt_pad, h->d_pos, h->d_orig, dy);
    h->launches++;
    CK(cudaGetLastError());
    if (probit) {
        DM(h, &h->d_ybin, (size_t)n_pad);
        DM(h, &h->d_cflist, (size_t)h->T + 1);
        k_probit_prepare<<<(unsigned)std::min<int64_t>((n_pad + 255) / 256, 4096), 256, 0, h->stream>>>(dy, h->d_ybin, n_pad);
        d.ybin = h->d_ybin; d.orig = h->d_orig; d.probit_offset = c.probit_offset; d.probit = 1;
        h->fuse_latent = !getenv("BCFGPU_NO_FUSED_LATENT");
        h->launches++;
        CK(cudaGetLastError());
    }
    CK(cudaMemsetAsync(d.totals, 0, (size_t)2 * h->T * KEYS * 8 * 8, h->stream));
    CK(cudaMemsetAsync(d.cross, 0, (size_t)2 * h->T * PIPE_XSTRIDE * 8, h->stream));
    CK(cudaMemsetAsync(d.mom, 0, 2 * 2 * NMOM * 3 * 8, h->stream));
    CK(cudaMemsetAsync(d.err, 0, 16, h->stream));
    CK(cudaMemsetAsync(d.trace, 0, (size_t)h->T * 5 * 4, h->stream));
    CK(cudaMemsetAsync(d.trace_logr, 0, (size_t)h->T * 8, h->stream));
    CK(cudaMemsetAsync(d.counters, 0, 32, h->stream));
    CK(cudaMemsetAsync(h->d_bar, 0, 64 * 4, h->stream));
    int rc;
    // x_moderate = the leading columns of x_control (same buffer, column-major, same cutpoints)?  Then bin once.
    bool shared_x = p_mod > 0 && xmod == xcon && p_mod <= p_con && cut_off_mod[0] == 0 && cut_off_con[0] == 0;
    for (int v = 0; shared_x && v < p_mod; ++v) {
        const int nc = cut_off_con[v + 1] - cut_off_con[v];
        shared_x = cut_off_mod[v] == cut_off_con[v] && cut_off_mod[v + 1] - cut_off_mod[v] == nc &&
                   memcmp(cuts_mod + cut_off_mod[v], cuts_con + cut_off_con[v], (size_t)std::max(nc, 0) * 8) == 0;
    }
    if ((rc = setup_forest(h, 0, p_con, cuts_con, cut_off_con))) return rc;
    if ((rc = setup_forest(h, 1, p_mod, cuts_mod, cut_off_mod, shared_x ? &h->fh[0] : nullptr))) return rc;
    sd_mark("y, z, cutpoints queued");
    if ((rc = bin_forest(h, 0, xcon))) return rc;
    if (!shared_x && (rc = bin_forest(h, 1, xmod))) return rc;
    sd_mark("X: H2D + binning (overlapped)");
    fill_forest_dev(h, 0, d.f[0]);
    fill_forest_dev(h, 1, d.f[1]);

    if (h->nranks > 1) {
        long long* dq = (long long*)h->d_tmp;
        CK(cudaMemcpyAsync(dq, ysum, 40, cudaMemcpyHostToDevice, h->stream));
        if ((rc = shard_allreduce(h, dq, 5))) return rc;
        CK(cudaMemcpyAsync(ysum, dq, 40, cudaMemcpyDeviceToHost, h->stream));
        CK(cudaStreamSynchronize(h->stream));
    }
    h->n_global = ysum[0];
    d.n_global = h->n_global;
    if (h->n_global > 2000000000ll) return fail(BCFGPU_ERR_BAD_ARG, "more than 2e9 observations over all shards");
    const int32_t ntrt_g = (int32_t)ysum[4], nctl_g = (int32_t)(ysum[0] - ysum[4]);
    const double ybar = limbs_to_double(ysum[1], ysum[2], ysum[3]) / (double)h->n_global;

    // trees, scalars [SURVEY A.2]
    {
        const double trt_init = 1.0;
        k_init_trees<<<h->T, 256, 0, h->stream>>>(d.trees[0], d.trees[1], c.ntree_con, ybar / c.ntree_con,
                                                  c.ntree_mod > 0 ? trt_init / c.ntree_mod : 0.0, ntrt_g, nctl_g);
        h->launches++;
        CK(cudaGetLastError());
        Scalars sc;
        memset(&sc, 0, sizeof(sc));
        sc.sigma = c.sigma_init; sc.mscale = 1.0; sc.bscale1 = 0.5; sc.bscale0 = -0.5;
        sc.delta_con = sc.delta_mod = 1.0;
        sc.tau_con = c.con_sd / std::sqrt((double)c.ntree_con);
        sc.tau_mod = c.ntree_mod > 0 ? c.mod_sd / std::sqrt((double)c.ntree_mod) : 1.0;
        sc.sweep = 0; sc.parity = 0; sc.run_it = 0; sc.nburn = 0; sc.nthin = 1; sc.nsim = 0; sc.save_ctr = 0; sc.save_row = -1;
        CK(cudaMemcpyAsync(d.sc, &sc, sizeof(sc), cudaMemcpyHostToDevice, h->stream));
        CK(cudaStreamSynchronize(h->stream));
    }
    h->grid_obs = (int)std::max<int64_t>(1, std::min<int64_t>((d.n_tiles + (THREADS / 32) - 1) / (THREADS / 32), (int64_t)h->num_sms * 4));
    h->grid_fin = (int)std::max<int64_t>(1, std::min<int64_t>((n_pad + 255) / 256, (int64_t)h->num_sms * 8));
    k_init_state<<<h->grid_fin, 256, 0, h->stream>>>(d, h->d_orig, ybar, c.ntree_mod > 0 ? 1.0 : 0.0);
    h->launches++;
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(h->stream));
    sd_mark("trees + initial state");
    h->has_data = true;
    h->nsaved = 0;
    return BCFGPU_OK;
}

// the host copy of the row permutation (only the stand-alone operators need it)
static int ensure_host_pos(bcfgpu_handle* h)
{
    if (h->have_pos) return BCFGPU_OK;
    h->pos.resize((size_t)h->n);
    CK(cudaMemcpyAsync(h->pos.data(), h->d_pos, (size_t)h->n * 4, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    h->have_pos = true;
    return BCFGPU_OK;
}

// ------------------------------------------------------------------------------------------------
// sweep launch: GRAPH engine
// ------------------------------------------------------------------------------------------------
static int enqueue_sweep(bcfgpu_handle* h)
{
    const Dev& d = h->dev;
    int rc;
    k_propose_all<<<h->T, THREADS, STEP_SMEM_BYTES, h->stream>>>(d);
    for (int j = 0; j < h->T; ++j) {
        k_tree_step<<<h->grid_obs, THREADS, STEP_SMEM_BYTES, h->stream>>>(d, j - 1, j);
        if ((rc = shard_allreduce(h, d.totals + (size_t)j * KEYS * 8, KEYS * 8))) return rc;
    }
    k_sweep_end<<<h->grid_obs, THREADS, STEP_SMEM_BYTES, h->stream>>>(d, h->T - 1);
    // both parity buffers (the sweep's own and the all-zero one of the next sweep): the host does not track the parity
    if ((rc = shard_allreduce(h, d.mom, 2 * 2 * NMOM * 3))) return rc;
    k_scalars<<<1, THREADS, 0, h->stream>>>(d);
    k_finish<<<h->grid_fin, 256, 0, h->stream>>>(d);
    CK(cudaGetLastError());
    return BCFGPU_OK;
}

static int build_graph(bcfgpu_handle* h)
{
    if (h->gexec && memcmp(&h->graph_dev, &h->dev, sizeof(Dev)) == 0) return BCFGPU_OK;
    drop_graph(h);
    CK(cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal));
    int rc = enqueue_sweep(h);
    cudaError_t e = cudaStreamEndCapture(h->stream, &h->graph);
    if (rc) { if (h->graph) { cudaGraphDestroy(h->graph); h->graph = nullptr; } return rc; }
    if (e != cudaSuccess) return fail(BCFGPU_ERR_CUDA, std::string("graph capture: ") + cudaGetErrorString(e));
    CK(cudaGraphInstantiate(&h->gexec, h->graph, 0));
    h->graph_dev = h->dev;
    return BCFGPU_OK;
}

static int run_persistent(bcfgpu_handle* h, int total, bool no_sync = false);
static int select_kernel(bcfgpu_handle* h);

static int check_device_err(bcfgpu_handle* h)
{
    struct { int32_t err; } sc;
    CK(cudaMemcpy(&sc.err, h->dev.err, 4, cudaMemcpyDeviceToHost));
    if (sc.err) cudaMemset(h->dev.err, 0, 4);
    if (sc.err & ERR_BARRIER_TIMEOUT) return fail(BCFGPU_ERR_CUDA, "grid barrier timed out inside the persistent sweep kernel");
    if (sc.err & ERR_NAN) return fail(BCFGPU_ERR_NUMERIC, "NaN in a leaf / scale / sigma draw");
    if (sc.err & ERR_FX_RANGE) return fail(BCFGPU_ERR_NUMERIC, "residual outside the fixed-point range (|r| >= 2^15) or NaN");
    return BCFGPU_OK;
}

static int run_impl(bcfgpu_handle* h, int32_t nburn, int32_t nsim, int32_t nthin);
extern "C" int bcfgpu_run(bcfgpu_handle* h, int32_t nburn, int32_t nsim, int32_t nthin)
{
    USE(h); NEED_DATA(h);
    const int rc = run_impl(h, nburn, nsim, nthin);
    h->stop.store(0);                 // a stop request ends ONE run
    return rc;
}
extern "C" int bcfgpu_request_stop(bcfgpu_handle* h)
{
    if (!h) return fail(BCFGPU_ERR_BAD_ARG, "null handle");
    h->stop.store(1);                 // no CUDA call, no lock: safe from any thread
    return BCFGPU_OK;
}
static int run_impl(bcfgpu_handle* h, int32_t nburn, int32_t nsim, int32_t nthin)
{
    if (nburn < 0 || nsim < 0 || nthin < 1) return fail(BCFGPU_ERR_BAD_ARG, "need nburn >= 0, nsim >= 0, nthin >= 1");
    const int64_t total64 = (int64_t)nburn + (int64_t)nsim * nthin;
    if (total64 > 2000000000ll) return fail(BCFGPU_ERR_BAD_ARG, "too many sweeps");
    const int total = (int)total64;
    Dev& d = h->dev;
    CK(cudaStreamSynchronize(h->stream));
    // posterior buffers of this run
    free_list(h, h->run_allocs);
    d.draws_tau = d.draws_mu = d.draws_yhat = nullptr;
    const int cap = std::max(nsim, 1);
    DM(h, &d.sigma_post, (size_t)cap, true); DM(h, &d.msd_post, (size_t)cap, true); DM(h, &d.bsd_post, (size_t)cap, true);
    d.post_cap = cap;
    h->keep = h->cfg.keep_draws;
    if (nsim > 0) {
        if (h->keep & 1) DM(h, &d.draws_tau, (size_t)nsim * d.n_pad, true);
        if (h->keep & 2) DM(h, &d.draws_mu, (size_t)nsim * d.n_pad, true);
        if (h->keep & 4) DM(h, &d.draws_yhat, (size_t)nsim * d.n_pad, true);
    }
    h->d_pool = nullptr; h->forests_saved = 0; h->pool_cap = 0;
    const bool keep_trees = (h->keep & 8) && nsim > 0 && h->cfg.model == BCFGPU_MODEL_BCF;
    if (keep_trees) {
        if (h->nranks > 1) return fail(BCFGPU_ERR_BAD_ARG, "saved forests (keep_draws bit 3) are not available in the sharded mode");
        // trees average 2 - 3 nodes at stationarity: room for 12 per tree and draw (+ 512 per draw: forests of a few wide trees), at most 4 GiB
        h->pool_cap = std::min<long long>((long long)nsim * ((long long)h->T * 12 + 512), ((long long)4 << 30) / (long long)sizeof(SavedNode));
        DM(h, &h->d_pool, (size_t)h->pool_cap, true);
        DM(h, &h->d_pool_used, 2, true);
        DM(h, &h->d_tree_off, (size_t)nsim * h->T, true);
        DM(h, &h->d_tree_cnt, (size_t)nsim * h->T, true);
        DM(h, &h->d_draw_scal, (size_t)nsim, true);
        CK(cudaMemsetAsync(h->d_pool_used, 0, 16, h->stream));
    }
    CK(cudaMemsetAsync(d.tau_sum, 0, (size_t)d.n_pad * 8, h->stream));
    CK(cudaMemsetAsync(d.tau_sq, 0, (size_t)d.n_pad * 8, h->stream));
    CK(cudaMemsetAsync(d.mu_sum, 0, (size_t)d.n_pad * 8, h->stream));
    CK(cudaMemsetAsync(d.yhat_sum, 0, (size_t)d.n_pad * 8, h->stream));
    Scalars sc;
    CK(cudaMemcpyAsync(&sc, d.sc, sizeof(sc), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    sc.run_it = 0; sc.nburn = nburn; sc.nthin = nthin; sc.nsim = nsim; sc.save_ctr = 0; sc.save_row = -1;
    if (h->cfg.model == BCFGPU_MODEL_PROBIT_BART) { sc.nburn = total; sc.nsim = 0; }      // (saved by the counterfactual pass instead)
    CK(cudaMemcpyAsync(d.sc, &sc, sizeof(sc), cudaMemcpyHostToDevice, h->stream));
    h->nsaved = 0;
    int rc = BCFGPU_OK;
    bool persistent = h->engine != BCFGPU_ENGINE_GRAPH && h->nranks == 1;
    if (h->nranks > 1 && h->peer_ok && h->engine != BCFGPU_ENGINE_GRAPH) {
        // observation shards with the in-kernel peer exchange: only the batched kernel has it, and every rank must
        // take the same path (the shards differ by a row, so "does it fit in shared memory" is agreed on, not assumed)
        if (!h->shard_decided) {
            if ((rc = select_kernel(h))) return rc;
            // level 2: the pipelined kernel (batches of 4, four mailbox slots); level 1: the batched kernel (batches of 8,
            // two slots; its shared-memory and HBM variants speak the same protocol); 0: neither.  Every rank runs the
            // lowest level any rank has.
            long long ok = h->kernel == 4 ? 2 : (h->kernel == 3 || h->kernel == 5) ? 1 : 0;
            long long* dq = (long long*)h->d_tmp;
            CK(cudaMemcpyAsync(dq, &ok, 8, cudaMemcpyHostToDevice, h->stream));
            if ((rc = shard_allreduce(h, dq, 1, true))) return rc;
            CK(cudaMemcpyAsync(&ok, dq, 8, cudaMemcpyDeviceToHost, h->stream));
            CK(cudaStreamSynchronize(h->stream));
            if (ok == 1 && h->kernel == 4) h->kernel = h->fallback_kernel;     // a peer's shard does not fit the pipelined kernel
            h->shard_persistent = ok != 0;
            h->shard_decided = true;
            if (!h->shard_persistent) d.pr.nranks = 1;
        }
        persistent = h->shard_persistent;
    }
    if (!persistent && h->nranks == 1) { if ((rc = build_graph(h))) return rc; }
    CK(cudaEventRecord(h->ev0, h->stream));
    h->last_kernel = 0;
    auto launch_sweeps = [&](int k, bool no_sync = false) -> int {
        if (persistent) {
            int r = run_persistent(h, k, no_sync);
            h->last_kernel = h->kernel;
            return r;
        }
        if (h->nranks == 1) {
            for (int it = 0; it < k; ++it) {
                if ((it & 63) == 0 && h->stop.load(std::memory_order_relaxed)) { cudaStreamSynchronize(h->stream); return fail(BCFGPU_ERR_INTERRUPTED, "run stopped by bcfgpu_request_stop"); }
                CK(cudaGraphLaunch(h->gexec, h->stream));
            }
            h->launches += (int64_t)k * (h->T + 4);
            return BCFGPU_OK;
        }
        int r = BCFGPU_OK;
        for (int it = 0; it < k && r == BCFGPU_OK; ++it) r = enqueue_sweep(h);
        h->launches += (int64_t)k * (h->T + 4);
        return r;
    };
    int probit_saved = 0;
    // Sweep-by-sweep running (probit BART: a latent draw before and a counterfactual pass after the sweep; saved forests: a
    // copy of the trees after it).  pre(it, stall) / post(it, stall) queue the kernels around sweep `it`.  With the
    // pipelined kernel the sweeps of a group are queued WITHOUT asking after each one whether it ran: k_pipe refuses a sweep
    // that holds a tree wider than 8 keys and raises a device flag that makes everything queued behind it return at once;
    // the host looks at the group's count, runs the refused sweep with the batched kernel and goes on behind it.
    auto stepwise = [&](int it0, int it1, auto pre, auto post) -> int {
        int r = BCFGPU_OK;
        const bool spec = persistent && h->nranks == 1 && (select_kernel(h) == BCFGPU_OK) && h->kernel == 4;
        int* stall = h->d_done + 2;
        for (int it = it0; it < it1 && r == BCFGPU_OK;) {
            if (h->stop.load(std::memory_order_relaxed)) return fail(BCFGPU_ERR_INTERRUPTED, "run stopped by bcfgpu_request_stop");
            if (!spec) {
                pre(it, nullptr);
                r = launch_sweeps(1);
                if (!persistent) h->latent_drawn = false;
                if (r == BCFGPU_OK) post(it, nullptr);
                ++it;
                continue;
            }
            const int m = std::min(32, it1 - it);
            CK(cudaMemsetAsync(h->d_done, 0, 12, h->stream));
            for (int j = 0; j < m && r == BCFGPU_OK; ++j) {
                pre(it + j, stall);
                r = run_persistent(h, 1, true);      // (probit BART: the sweeps queued behind it find their latent drawn by this one)
                post(it + j, stall);
            }
            if (r) break;
            int dn_host[3] = {0, 0, 0};
            CK(cudaMemcpyAsync(dn_host, h->d_done, 12, cudaMemcpyDeviceToHost, h->stream));
            CK(cudaStreamSynchronize(h->stream));
            h->last_kernel = 4;
            h->pipe_sweeps += dn_host[0];
            if (dn_host[0] >= m) { it += m; continue; }
            // sweep it + c was refused (its pre() had run): the batched kernel takes it
            const int c = dn_host[0];
            h->fallback_sweeps++;
            CK(cudaMemsetAsync(h->d_done, 0, 12, h->stream));
            const int keep_kernel = h->kernel;
            h->kernel = h->fallback_kernel;
            r = run_persistent(h, 1);                // (the batched kernel draws no latent: latent_drawn is false behind it)
            h->kernel = keep_kernel;
            if (r == BCFGPU_OK) post(it + c, nullptr);
            it += c + 1;
        }
        return r;
    };
    auto save_index = [&](int it) -> int {   // bcf: sweep it is saved when it >= nburn and it % nthin == 0; which saved sweep is it?
        if (it < nburn || (it % nthin) != 0) return -1;
        const int first = (nburn + nthin - 1) / nthin * nthin;
        const int k = (it - first) / nthin;
        return k < nsim ? k : -1;
    };
    if (h->cfg.model == BCFGPU_MODEL_PROBIT_BART) {
        // A latent draw opens every sweep, the sweep kernel runs it as a burn-in sweep (the engines' own posterior accumulation
        // stays off), and a saved sweep ends with the counterfactual pass.  The pipelined kernel makes the next sweep's draw in
        // its own finishing pass (h->latent_drawn: run_persistent keeps it), so burn-in runs many sweeps per launch and a saved
        // sweep costs one launch + the counterfactual pass; the other engines get the draw from k_probit_latent before every
        // sweep, one sweep per launch (every engine parks the chain state in HBM between launches).
        const int grid = h->grid_fin;
        const int cfgrid = (int)((d.n_pad + 256 * CF_OBS_PER_THREAD - 1) / (256 * CF_OBS_PER_THREAD));
        h->latent_drawn = false;                       // (a run's first sweep: drawn here; redrawing an existing draw gives the same numbers)
        auto draw = [&](int, const int* stall) {
            if (h->latent_drawn) return;
            k_probit_latent<<<grid, 256, 0, h->stream>>>(d, h->d_ybin, h->d_orig, 0, h->cfg.probit_offset, const_cast<double*>(d.y), stall);
            h->launches++;
        };
        int it0 = 0;
        if (h->fuse_latent && persistent && h->nranks == 1 && select_kernel(h) == BCFGPU_OK && h->kernel == 4) {
            int first_saved = total;
            for (int it = 0; it < total; ++it) if (save_index(it) >= 0) { first_saved = it; break; }
            if (first_saved > 0) {
                draw(0, nullptr);
                rc = launch_sweeps(first_saved);
                it0 = first_saved;
            }
        }
        if (rc == BCFGPU_OK) rc = stepwise(it0, total, draw,
            [&](int it, const int* stall) {
                if (save_index(it) < 0) return;
                const int par = (sc.parity + it + 1) & 1;     // the sweep flipped the tree parity
                if (h->cfg.treatment_col >= 0) {
                    k_cf_collect<<<h->T, 64, 0, h->stream>>>(d, d.trees[par], h->cfg.treatment_col, h->d_cflist, stall);
                }
                k_cf_accumulate<<<cfgrid, 256, 0, h->stream>>>(d, d.trees[par], h->d_cflist, h->cfg.treatment_col,
                                                               h->cfg.probit_offset, h->d_orig, stall);
                h->launches += 2;
            });
        if (rc == BCFGPU_OK) CK(cudaGetLastError());
        for (int it = 0; it < total; ++it) probit_saved += save_index(it) >= 0 ? 1 : 0;
    } else if (keep_trees) {
        // the sampling part runs one sweep per launch so that the trees of every saved sweep can be copied out between
        // launches (the engines keep only the current forest)
        rc = launch_sweeps(nburn);
        if (rc == BCFGPU_OK)
            rc = stepwise(nburn, total, [&](int, const int*) {},
                [&](int it, const int* stall) {
                    const int k = save_index(it);
                    if (k < 0) return;
                    const int par = (sc.parity + it + 1) & 1;
                    k_save_trees<<<h->T, 64, 0, h->stream>>>(d, d.trees[par], k, h->d_pool, h->pool_cap, h->d_pool_used,
                                                             h->d_tree_off, h->d_tree_cnt, h->d_draw_scal, stall);
                    h->launches++;
                });
        if (rc == BCFGPU_OK) CK(cudaGetLastError());
        int saved = 0;
        for (int it = nburn; it < total; ++it) saved += save_index(it) >= 0 ? 1 : 0;
        h->forests_saved = saved;
    } else
        rc = launch_sweeps(total);
    if (rc) {
        cudaStreamSynchronize(h->stream);
        if (h->dbg_host && !h->dbg_host[0] && h->nranks > 1) {
            char note[160];
            snprintf(note, sizeof(note), " [rank %d: k_pipe launches started %d, past their first barrier %d]", h->rank, h->dbg_host[8], h->dbg_host[9]);
            g_err += note;
        }
        if (h->dbg_host && h->dbg_host[0]) {
            char note[320];
            snprintf(note, sizeof(note), " [kernel gave up: code %d (20: peer words, 30: utility warp's grid barrier; 5x: k_hist_mma), CTA %d, %d, %d; rank %d; k_pipe launches started %d, past their first barrier %d]",
                     h->dbg_host[1], h->dbg_host[2], h->dbg_host[3], h->dbg_host[4], h->rank, h->dbg_host[8], h->dbg_host[9]);
            g_err += note;
        }
        return rc;
    }
    CK(cudaEventRecord(h->ev1, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, h->ev0, h->ev1));
    h->last_ms = ms;
    if ((rc = check_device_err(h))) return rc;
    CK(cudaMemcpy(&sc, d.sc, sizeof(sc), cudaMemcpyDeviceToHost));
    h->nsaved = h->cfg.model == BCFGPU_MODEL_PROBIT_BART ? probit_saved : sc.save_ctr;
#ifndef BCF_PROF
    if (d.prof) {
        static bool told = false;
        if (!told) fprintf(stderr, "[bcfgpu profile] the phase timers of the sweep kernels are compiled out (they cost ~6 %% of a sweep): "
                                   "rebuild with BCFGPU_NVCC_EXTRA=-DBCF_PROF python -m bcf_iv_b200.build\n");
        told = true;
    } else
#endif
    if (d.prof) {   // BCFGPU_PROFILE=1: phase cycle counters of CTA 0 (persistent engines)
        long long pc[16];
        CK(cudaMemcpy(pc, d.prof, 128, cudaMemcpyDeviceToHost));
        CK(cudaMemset(d.prof, 0, 128));
        const char* nm[8] = {"tiles", "flush", "barrier", "decide", "store", "load+propose", "epilogue", "-"};   // k_pipe: stats, flush, wait H2, apply, -, plan
        long long tot = 0; for (int k = 0; k < 7; ++k) tot += pc[k];
        fprintf(stderr, "[bcfgpu profile] %d sweeps:", total);
        for (int k = 0; k < 8; ++k) fprintf(stderr, " %s=%.1f%%(%.0f cyc/tree)", nm[k], 100.0 * pc[k] / (double)std::max(tot, 1ll), (double)pc[k] / ((double)total * h->T));
        fprintf(stderr, "\n");
        bool any = false; for (int k = 8; k < 16; ++k) any |= pc[k] != 0;
        if (any) {
            fprintf(stderr, "[bcfgpu profile] resolve detail (cyc/tree):");
            for (int k = 8; k < 16; ++k) fprintf(stderr, " r%d=%.0f", k - 8, (double)pc[k] / ((double)total * h->T));
            fprintf(stderr, "\n");
        }
        if (getenv("BCFGPU_TRACE")) {   // k_pipe: per-CTA phase totals (cycles per batch): 0 stats, 2 wait H2, 3 apply, 5 wait plan | 4 wait H1, 6 wait grid barrier, 7 resolve
            std::vector<long long> pcx(512 * 8);
            CK(cudaMemcpy(pcx.data(), d.prof + 1024, pcx.size() * 8, cudaMemcpyDeviceToHost));
            CK(cudaMemset(d.prof + 1024, 0, pcx.size() * 8));
            const double nbat = (double)total * ((h->cfg.ntree_con + 3) / 4 + (h->cfg.ntree_mod + 3) / 4);
            for (int k = 0; k < 16; ++k) {   // 8..10: utility warp (wait for the flush, plan, stage); 11..15: sweep epilogue (moments, barrier, scalars, proposals, finishing pass + barrier)
                std::vector<double> v;
                for (int c = 0; c < h->coop_grid && c < 256; ++c) v.push_back((double)pcx[(k < 8 ? c * 8 + k : 2048 + c * 8 + k - 8)] / nbat);
                if (v.empty()) break;
                std::vector<double> sv = v; std::sort(sv.begin(), sv.end());
                if (k == 7 || k == 6 || k == 3 || k == 0) { fprintf(stderr, "[bcfgpu trace] phase %d all:", k); for (double x : v) fprintf(stderr, " %.0f", x / 100.0); fprintf(stderr, "\n"); }
                fprintf(stderr, "[bcfgpu trace] per-CTA phase %d: cta0=%.0f min=%.0f med=%.0f max=%.0f (cta of max %d)\n", k, v[0], sv.front(), sv[sv.size() / 2], sv.back(),
                        (int)(std::max_element(v.begin(), v.end()) - v.begin()));
            }
        }
        if (getenv("BCFGPU_TRACE")) {   // k_pipe: CTA 0's batch timeline in the launch's first sweep (cycles since batch 8's plan)
            long long tr[64 * 8];
            CK(cudaMemcpy(tr, d.prof + 16, sizeof(tr), cudaMemcpyDeviceToHost));
            const long long t0 = tr[8 * 8];
            const char* ev[8] = {"plan", "stats", "flushed", "H2got", "applied", "H1got", "gridbar", "resolved"};
            for (int b = 8; b < 16 && t0 != 0; ++b) {
                fprintf(stderr, "[bcfgpu trace] batch %2d:", b);
                for (int k = 0; k < 8; ++k) fprintf(stderr, " %s=%lld", ev[k], tr[b * 8 + k] - t0);
                fprintf(stderr, "\n");
            }
            long long ch[256];
            CK(cudaMemcpy(ch, d.prof + 16 + 512, sizeof(ch), cudaMemcpyDeviceToHost));
            for (int b = 8; b < 12 && t0 != 0; ++b) {   // decision chain: per warp enter, prev folded, turn, arrived, booked, left
                fprintf(stderr, "[bcfgpu trace] chain %2d:", b);
                for (int w = 0; w < 4; ++w) {
                    fprintf(stderr, " |w%d", w);
                    for (int e = 0; e < 6; ++e) fprintf(stderr, " %lld", ch[(b - 8) * 32 + w * 8 + e] - t0);
                }
                const long long* u = ch + (b - 8) * 32;
                fprintf(stderr, " |util S=%lld flush=%lld rel=%lld plan=%lld H2=%lld staged=%lld\n", u[6] - t0, u[7] - t0, u[14] - t0, u[15] - t0, u[22] - t0, u[23] - t0);
            }
        }
    }
    return BCFGPU_OK;
}

extern "C" {

int bcfgpu_last_run_ms(bcfgpu_handle* h, double* ms)
{
    if (!h || !ms) return fail(BCFGPU_ERR_BAD_ARG, "null argument");
    *ms = h->last_ms;
    return BCFGPU_OK;
}
int bcfgpu_h2d_bytes(bcfgpu_handle* h, int64_t* bytes)
{
    if (!h || !bytes) return fail(BCFGPU_ERR_BAD_ARG, "null argument");
    *bytes = h->h2d_bytes;
    return BCFGPU_OK;
}
int bcfgpu_kernel_launches(bcfgpu_handle* h, int64_t* count)
{
    if (!h || !count) return fail(BCFGPU_ERR_BAD_ARG, "null argument");
    *count = h->launches;
    return BCFGPU_OK;
}
int bcfgpu_engine_in_use(bcfgpu_handle* h, int32_t out2[2])
{
    if (!h || !out2) return fail(BCFGPU_ERR_BAD_ARG, "null argument");
    out2[0] = h->last_kernel;
    out2[1] = h->nranks <= 1 ? 0 : (h->last_kernel == 0 ? 1 : 2);
    return BCFGPU_OK;
}
int bcfgpu_engine_stats(bcfgpu_handle* h, int64_t out2[2])
{
    if (!h || !out2) return fail(BCFGPU_ERR_BAD_ARG, "null argument");
    out2[0] = h->pipe_sweeps; out2[1] = h->fallback_sweeps;
    return BCFGPU_OK;
}
int bcfgpu_num_saved(bcfgpu_handle* h, int32_t* ns)
{
    if (!h || !ns) return fail(BCFGPU_ERR_BAD_ARG, "null argument");
    *ns = h->nsaved;
    return BCFGPU_OK;
}

}  // extern "C"

// ------------------------------------------------------------------------------------------------
// results
// ------------------------------------------------------------------------------------------------
static int fetch_vec(bcfgpu_handle* h, const double* a, const double* b, int which, double st, double sc_, double* out)
{
    const Dev& d = h->dev;
    const int grid = (int)std::min<int64_t>((h->n + 255) / 256, 4096);
    k_unpermute<<<grid, 256, 0, h->stream>>>(a, b, h->d_pos, h->n, (double)h->nsaved, which, st, sc_,
                                              (int64_t)d.trt_tiles * TILE, h->d_tmp);
    h->launches++;
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(out, h->d_tmp, (size_t)h->n * 8, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return BCFGPU_OK;
}
#define NEED_SAVED(h) do { if ((h)->nsaved < 1) return fail(BCFGPU_ERR_STATE, "no saved draws: call run with nsim >= 1"); } while (0)

extern "C" {

int bcfgpu_get_tau_mean(bcfgpu_handle* h, double* out)
{
    USE(h); NEED_DATA(h); NEED_SAVED(h);
    if (!out) return fail(BCFGPU_ERR_BAD_ARG, "null output");
    return fetch_vec(h, h->dev.tau_sum, nullptr, 1, 1, 1, out);
}
int bcfgpu_get_tau_var(bcfgpu_handle* h, double* out)
{
    USE(h); NEED_DATA(h); NEED_SAVED(h);
    if (!out) return fail(BCFGPU_ERR_BAD_ARG, "null output");
    return fetch_vec(h, h->dev.tau_sum, h->dev.tau_sq, 2, 1, 1, out);
}
int bcfgpu_get_mu_mean(bcfgpu_handle* h, double* out)
{
    USE(h); NEED_DATA(h); NEED_SAVED(h);
    if (!out) return fail(BCFGPU_ERR_BAD_ARG, "null output");
    return fetch_vec(h, h->dev.mu_sum, nullptr, 1, 1, 1, out);
}
int bcfgpu_get_yhat_mean(bcfgpu_handle* h, double* out)
{
    USE(h); NEED_DATA(h); NEED_SAVED(h);
    if (!out) return fail(BCFGPU_ERR_BAD_ARG, "null output");
    return fetch_vec(h, h->dev.yhat_sum, nullptr, 1, 1, 1, out);
}
int bcfgpu_get_sigma(bcfgpu_handle* h, double* out)
{
    USE(h); NEED_DATA(h);
    if (!out) return fail(BCFGPU_ERR_BAD_ARG, "null output");
    if (h->nsaved > 0) CK(cudaMemcpy(out, h->dev.sigma_post, (size_t)h->nsaved * 8, cudaMemcpyDeviceToHost));
    return BCFGPU_OK;
}
int bcfgpu_get_scales(bcfgpu_handle* h, double* mu_scale, double* tau_scale)
{
    USE(h); NEED_DATA(h);
    if (h->nsaved > 0 && mu_scale) CK(cudaMemcpy(mu_scale, h->dev.msd_post, (size_t)h->nsaved * 8, cudaMemcpyDeviceToHost));
    if (h->nsaved > 0 && tau_scale) CK(cudaMemcpy(tau_scale, h->dev.bsd_post, (size_t)h->nsaved * 8, cudaMemcpyDeviceToHost));
    return BCFGPU_OK;
}
int bcfgpu_get_draws(bcfgpu_handle* h, int32_t which, double* out)
{
    USE(h); NEED_DATA(h);
    if (!out) return fail(BCFGPU_ERR_BAD_ARG, "null output");
    const double* src = which == 0 ? h->dev.draws_tau : which == 1 ? h->dev.draws_mu : which == 2 ? h->dev.draws_yhat : nullptr;
    if (!src) return fail(BCFGPU_ERR_STATE, "these draws were not kept (bcfgpu_config.keep_draws)");
    for (int s = 0; s < h->nsaved; ++s) {
        int rc = fetch_vec(h, src + (size_t)s * h->dev.n_pad, nullptr, 0, 1.0, 1.0, out + (size_t)s * h->n);
        if (rc) return rc;
    }
    return BCFGPU_OK;
}

int bcfgpu_get_scalars(bcfgpu_handle* h, double out[8])
{
    USE(h); NEED_DATA(h);
    CK(cudaStreamSynchronize(h->stream));
    Scalars sc;
    CK(cudaMemcpy(&sc, h->dev.sc, sizeof(sc), cudaMemcpyDeviceToHost));
    out[0] = sc.sigma; out[1] = sc.mscale; out[2] = sc.bscale1; out[3] = sc.bscale0; out[4] = sc.delta_con;
    out[5] = sc.tau_con; out[6] = sc.tau_mod; out[7] = (double)sc.sweep;
    return BCFGPU_OK;
}
int bcfgpu_get_fits(bcfgpu_handle* h, double* ac, double* am)
{
    USE(h); NEED_DATA(h);
    CK(cudaStreamSynchronize(h->stream));
    Scalars sc;
    CK(cudaMemcpy(&sc, h->dev.sc, sizeof(sc), cudaMemcpyDeviceToHost));
    int rc;
    if (ac && (rc = fetch_vec(h, h->dev.Gc, nullptr, 0, sc.mscale, sc.mscale, ac))) return rc;
    if (am && (rc = fetch_vec(h, h->dev.Gm, nullptr, 0, sc.bscale1, sc.bscale0, am))) return rc;
    return BCFGPU_OK;
}
int bcfgpu_get_counters(bcfgpu_handle* h, int64_t out[4])
{
    USE(h); NEED_DATA(h);
    CK(cudaStreamSynchronize(h->stream));
    CK(cudaMemcpy(out, h->dev.counters, 32, cudaMemcpyDeviceToHost));
    return BCFGPU_OK;
}
int bcfgpu_get_trace(bcfgpu_handle* h, int32_t* trace5, double* logr)
{
    USE(h); NEED_DATA(h);
    CK(cudaStreamSynchronize(h->stream));
    if (trace5) CK(cudaMemcpy(trace5, h->dev.trace, (size_t)h->T * 5 * 4, cudaMemcpyDeviceToHost));
    if (logr) CK(cudaMemcpy(logr, h->dev.trace_logr, (size_t)h->T * 8, cudaMemcpyDeviceToHost));
    return BCFGPU_OK;
}

}  // extern "C"

static int current_parity(bcfgpu_handle* h, int* par)
{
    Scalars sc;
    CK(cudaStreamSynchronize(h->stream));
    CK(cudaMemcpy(&sc, h->dev.sc, sizeof(sc), cudaMemcpyDeviceToHost));
    *par = sc.parity;
    return BCFGPU_OK;
}
static int fetch_tree(bcfgpu_handle* h, int tree, TreeRec* tr)
{
    if (tree < 0 || tree >= h->T) return fail(BCFGPU_ERR_BAD_ARG, "tree index out of range");
    int par, rc;
    if ((rc = current_parity(h, &par))) return rc;
    CK(cudaMemcpy(tr, h->dev.trees[par] + tree, sizeof(TreeRec), cudaMemcpyDeviceToHost));
    return BCFGPU_OK;
}
static void preorder(const TreeRec& tr, int node, uint32_t* heap, int32_t* var, int32_t* cut, double* mu, int& k)
{
    const bool leaf = tr.left[node] < 0;
    heap[k] = tr.heap[node]; var[k] = leaf ? -1 : (int32_t)tr.var[node]; cut[k] = leaf ? -1 : (int32_t)tr.cut[node];
    mu[k] = leaf ? tr.mu[tr.node_slot[node]] : 0.0;
    ++k;
    if (!leaf) { preorder(tr, tr.left[node], heap, var, cut, mu, k); preorder(tr, tr.right[node], heap, var, cut, mu, k); }
}

extern "C" {

int bcfgpu_tree_size(bcfgpu_handle* h, int32_t tree, int32_t* nnodes)
{
    USE(h); NEED_DATA(h);
    TreeRec tr; int rc;
    if ((rc = fetch_tree(h, tree, &tr))) return rc;
    *nnodes = tr.nnodes;
    return BCFGPU_OK;
}
int bcfgpu_get_tree(bcfgpu_handle* h, int32_t tree, uint32_t* heap, int32_t* var, int32_t* cut, double* mu)
{
    USE(h); NEED_DATA(h);
    TreeRec tr; int rc;
    if ((rc = fetch_tree(h, tree, &tr))) return rc;
    int k = 0;
    preorder(tr, 0, heap, var, cut, mu, k);
    return BCFGPU_OK;
}

int bcfgpu_set_tree(bcfgpu_handle* h, int32_t tree, int32_t nnodes, const uint32_t* heap, const int32_t* var,
                    const int32_t* cut, const double* mu)
{
    USE(h); NEED_DATA(h);
    if (tree < 0 || tree >= h->T) return fail(BCFGPU_ERR_BAD_ARG, "tree index out of range");
    if (nnodes < 1 || nnodes > 2 * h->cfg.max_leaves - 1 || (nnodes % 2) == 0) return fail(BCFGPU_ERR_BAD_ARG, "bad node count");
    const int f = tree >= h->cfg.ntree_con ? 1 : 0;
    const ForestHost& F = h->fh[f];
    TreeRec tr;
    root_tree(tr, 0.0);
    tr.nnodes = nnodes; tr.nleaves = 0;
    // nodes keep the caller's order; link them through their heap ids
    std::vector<std::pair<uint32_t, int>> byheap;
    for (int k = 0; k < nnodes; ++k) byheap.push_back({heap[k], k});
    std::sort(byheap.begin(), byheap.end());
    auto find = [&](uint32_t id) -> int {
        auto it = std::lower_bound(byheap.begin(), byheap.end(), std::make_pair(id, -1));
        return (it != byheap.end() && it->first == id) ? it->second : -1;
    };
    const int root = find(1u);
    if (root < 0) return fail(BCFGPU_ERR_BAD_ARG, "serialised tree has no root (heap id 1)");
    // node 0 must be the root: swap indices
    std::vector<int> idx(nnodes);
    for (int k = 0; k < nnodes; ++k) idx[k] = k;
    std::swap(idx[root], idx[0]);   // caller node k -> TreeRec node idx[k]
    for (int k = 0; k < nnodes; ++k) {
        const int i = idx[k];
        const uint32_t id = heap[k];
        if (id == 0) return fail(BCFGPU_ERR_BAD_ARG, "heap id 0");
        int depth = 0;
        for (uint32_t t = id; t > 1; t >>= 1) ++depth;
        if (depth > MAX_DEPTH) return fail(BCFGPU_ERR_BAD_ARG, "tree deeper than BCFGPU_MAX_DEPTH");
        tr.heap[i] = id; tr.depth[i] = (uint8_t)depth;
        const int par = id == 1 ? -1 : find(id >> 1);
        if (id != 1 && par < 0) return fail(BCFGPU_ERR_BAD_ARG, "node without parent");
        tr.parent[i] = (int16_t)(par < 0 ? -1 : idx[par]);
        const int l = find(2 * id), r = find(2 * id + 1);
        if ((l < 0) != (r < 0)) return fail(BCFGPU_ERR_BAD_ARG, "node with one child");
        if (l >= 0) {
            if (var[k] < 0 || var[k] >= F.p || cut[k] < 0 || cut[k] >= F.ncut[var[k]]) return fail(BCFGPU_ERR_BAD_ARG, "split rule out of range");
            tr.left[i] = (int16_t)idx[l]; tr.right[i] = (int16_t)idx[r];
            tr.var[i] = (uint16_t)var[k]; tr.cut[i] = (uint16_t)cut[k];
        } else {
            const int s = tr.nleaves++;
            if (s >= h->cfg.max_leaves) return fail(BCFGPU_ERR_BAD_ARG, "too many leaves");
            tr.node_slot[i] = (uint8_t)s; tr.slot_node[s] = (int16_t)i; tr.mu[s] = mu[k];
        }
    }
    if (2 * tr.nleaves - 1 != nnodes) return fail(BCFGPU_ERR_BAD_ARG, "not a full binary tree");
    int par, rc;
    if ((rc = current_parity(h, &par))) return rc;
    CK(cudaMemcpy(h->dev.trees[par] + tree, &tr, sizeof(TreeRec), cudaMemcpyHostToDevice));
    k_assign<<<h->grid_fin, 256, 0, h->stream>>>(h->dev, tree, f, h->dev.trees[par], 0, nullptr, h->d_orig);
    // leaf counts (cached in the tree): count the fresh assignment, over all shards
    long long* dcnt = (long long*)h->d_tmp2;
    CK(cudaMemsetAsync(dcnt, 0, (size_t)MAXL * 2 * 8, h->stream));
    k_count_leaves<<<h->grid_fin, 256, 0, h->stream>>>(h->dev, tree, dcnt);
    if ((rc = shard_allreduce(h, dcnt, (size_t)MAXL * 2))) return rc;
    std::vector<long long> cnt((size_t)MAXL * 2);
    CK(cudaMemcpyAsync(cnt.data(), dcnt, cnt.size() * 8, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    for (int s = 0; s < tr.nleaves; ++s) { tr.cnt0[s] = (int32_t)cnt[2 * s]; tr.cnt1[s] = (int32_t)cnt[2 * s + 1]; }
    CK(cudaMemcpy(h->dev.trees[0] + tree, &tr, sizeof(TreeRec), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(h->dev.trees[1] + tree, &tr, sizeof(TreeRec), cudaMemcpyHostToDevice));
    k_refit_all<<<h->grid_fin, 256, 0, h->stream>>>(h->dev, h->dev.trees[par]);
    h->launches += 3;
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(h->stream));
    return check_device_err(h);
}

int bcfgpu_get_leaf_ids(bcfgpu_handle* h, int32_t tree, uint32_t* out)
{
    USE(h); NEED_DATA(h);
    if (tree < 0 || tree >= h->T || !out) return fail(BCFGPU_ERR_BAD_ARG, "bad argument");
    int par, rc;
    if ((rc = current_parity(h, &par))) return rc;
    uint32_t* tmp = (uint32_t*)h->d_tmp;
    k_leaf_heap<<<(int)std::min<int64_t>((h->n + 255) / 256, 4096), 256, 0, h->stream>>>(h->dev, tree, h->dev.trees[par], h->d_pos, h->n, tmp);
    h->launches++;
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(out, tmp, (size_t)h->n * 4, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return BCFGPU_OK;
}

int bcfgpu_verify_leaf_cache(bcfgpu_handle* h, int64_t* mismatches)
{
    USE(h); NEED_DATA(h);
    if (!mismatches) return fail(BCFGPU_ERR_BAD_ARG, "null output");
    int par, rc;
    if ((rc = current_parity(h, &par))) return rc;
    unsigned long long* cnt = (unsigned long long*)h->d_tmp2;
    CK(cudaMemsetAsync(cnt, 0, 8, h->stream));
    for (int t = 0; t < h->T; ++t)
        k_assign<<<h->grid_fin, 256, 0, h->stream>>>(h->dev, t, t >= h->cfg.ntree_con ? 1 : 0, h->dev.trees[par], 1, cnt, h->d_orig);
    h->launches += h->T;
    CK(cudaGetLastError());
    unsigned long long v = 0;
    CK(cudaMemcpyAsync(&v, cnt, 8, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    *mismatches = (int64_t)v;
    return BCFGPU_OK;
}

// ------------------------------------------------------------------------------------------------
// stand-alone operators
// ------------------------------------------------------------------------------------------------
int bcfgpu_op_bin_column(const double* x, int64_t n, const double* cuts, int32_t ncut, uint16_t* out)
{
    if (!x || !cuts || !out || n < 1 || ncut < 1 || ncut > 65535) return fail(BCFGPU_ERR_BAD_ARG, "bad argument");
    if (bcfgpu_device_count() == 0) return fail(BCFGPU_ERR_CUDA, "no CUDA device available (libbcfgpu has no CPU fallback)");
    double *dx = nullptr, *dc = nullptr; uint16_t* dout = nullptr;
    CK(cudaMalloc((void**)&dx, (size_t)n * 8));
    CK(cudaMalloc((void**)&dc, (size_t)ncut * 8));
    CK(cudaMalloc((void**)&dout, (size_t)n * 2));
    CK(cudaMemcpy(dx, x, (size_t)n * 8, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dc, cuts, (size_t)ncut * 8, cudaMemcpyHostToDevice));
    k_bin_simple<<<(int)std::min<int64_t>((n + 255) / 256, 4096), 256>>>(dx, n, dc, ncut, dout);
    CK(cudaGetLastError());
    CK(cudaMemcpy(out, dout, (size_t)n * 2, cudaMemcpyDeviceToHost));
    cudaFree(dx); cudaFree(dc); cudaFree(dout);
    return BCFGPU_OK;
}

int bcfgpu_op_suffstats(bcfgpu_handle* h, int32_t tree, const double* r, uint32_t birth_leaf_heap, int32_t var,
                        int32_t cut, int32_t* nleaves, uint32_t* heap, int64_t* cnt_trt, int64_t* cnt_ctl,
                        double* sum_trt, double* sum_ctl, double out_birth[8])
{
    USE(h); NEED_DATA(h);
    if (!r || !nleaves || !heap || !cnt_trt || !cnt_ctl || !sum_trt || !sum_ctl) return fail(BCFGPU_ERR_BAD_ARG, "null argument");
    TreeRec tr; int rc;
    if ((rc = fetch_tree(h, tree, &tr))) return rc;
    const int f = tree >= h->cfg.ntree_con ? 1 : 0;
    const int L = tr.nleaves;
    int sx = -1;
    if (birth_leaf_heap) {
        for (int s = 0; s < L; ++s) if (tr.heap[tr.slot_node[s]] == birth_leaf_heap) sx = s;
        if (sx < 0) return fail(BCFGPU_ERR_BAD_ARG, "no leaf with that heap id");
        if (var < 0 || var >= h->fh[f].p || cut < 0 || cut >= h->fh[f].ncut[var]) return fail(BCFGPU_ERR_BAD_ARG, "split rule out of range");
        if (!out_birth) return fail(BCFGPU_ERR_BAD_ARG, "null out_birth");
    }
    const int K = L + (sx >= 0 ? 1 : 0);
    // r in the caller's order -> internal order (padding rows 0)
    CK(cudaMemsetAsync(h->d_tmp2, 0, (size_t)h->n_pad * 8, h->stream));
    CK(cudaMemcpyAsync(h->d_tmp, r, (size_t)h->n * 8, cudaMemcpyHostToDevice, h->stream));
    k_permute_in<<<(int)std::min<int64_t>((h->n + 255) / 256, 4096), 256, 0, h->stream>>>(h->d_tmp, h->d_pos, h->n, h->d_tmp2);
    long long* dout = nullptr;
    CK(cudaMalloc((void**)&dout, (size_t)KEYS * 8 * 8));
    CK(cudaMemsetAsync(dout, 0, (size_t)KEYS * 8 * 8, h->stream));
    k_op_stats<<<h->grid_obs, THREADS, 0, h->stream>>>(h->dev, tree, f, h->d_tmp2, K, sx >= 0, sx, var, cut, dout);
    h->launches += 2;
    std::vector<long long> tot((size_t)KEYS * 8);
    cudaError_t e = cudaMemcpyAsync(tot.data(), dout, tot.size() * 8, cudaMemcpyDeviceToHost, h->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
    cudaFree(dout);
    if (e != cudaSuccess) return fail(BCFGPU_ERR_CUDA, std::string("op_suffstats: ") + cudaGetErrorString(e));
    if ((rc = check_device_err(h))) return rc;
    // counts: cached in the tree (kept exact by the sampler: birth children are counted, everything else follows);
    // only the would-be right child of the birth split (key L) is counted by this pass
    auto cntv = [&](int k, int g) -> long long {
        if (k == L) return tot[(size_t)k * 8 + g * 4];
        const long long whole = g ? tr.cnt1[k] : tr.cnt0[k];
        return (k == sx) ? whole - tot[(size_t)L * 8 + g * 4] : whole;
    };
    auto sumv = [&](int k, int g) { const long long* q = &tot[(size_t)k * 8 + g * 4]; return limbs_to_double(q[1], q[2], q[3]); };
    auto sum_exact = [&](int k1, int k2, int g) {   // exact merge of two keys before rounding
        const long long* a = &tot[(size_t)k1 * 8 + g * 4]; const long long* b = &tot[(size_t)k2 * 8 + g * 4];
        return limbs_to_double(a[1] + b[1], a[2] + b[2], a[3] + b[3]);
    };
    // leaves in DFS (left-to-right) order
    std::vector<std::pair<uint32_t, int>> order;
    for (int s = 0; s < L; ++s) { const int nd = tr.slot_node[s]; order.push_back({tr.heap[nd] << (MAX_DEPTH - tr.depth[nd]), s}); }
    std::sort(order.begin(), order.end());
    *nleaves = L;
    for (int q = 0; q < L; ++q) {
        const int s = order[q].second;
        heap[q] = tr.heap[tr.slot_node[s]];
        if (s == sx) {
            cnt_trt[q] = cntv(s, 1) + cntv(L, 1); cnt_ctl[q] = cntv(s, 0) + cntv(L, 0);
            sum_trt[q] = sum_exact(s, L, 1); sum_ctl[q] = sum_exact(s, L, 0);
        } else {
            cnt_trt[q] = cntv(s, 1); cnt_ctl[q] = cntv(s, 0); sum_trt[q] = sumv(s, 1); sum_ctl[q] = sumv(s, 0);
        }
    }
    if (sx >= 0) {
        out_birth[0] = (double)cntv(sx, 1); out_birth[1] = (double)cntv(sx, 0); out_birth[2] = sumv(sx, 1); out_birth[3] = sumv(sx, 0);
        out_birth[4] = (double)cntv(L, 1); out_birth[5] = (double)cntv(L, 0); out_birth[6] = sumv(L, 1); out_birth[7] = sumv(L, 0);
    }
    return BCFGPU_OK;
}

// cuTensorMapEncodeTiled through the runtime (the library links cudart only): the TMA descriptor of K2''s B operand
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn tensor_map_encoder()
{
    static EncodeTiledFn fn = []() -> EncodeTiledFn {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) {
            cudaGetLastError();
            return nullptr;
        }
        return (EncodeTiledFn)p;
    }();
    return fn;
}

int bcfgpu_op_hist_all(bcfgpu_handle* h, int32_t forest, const int32_t* slot, int32_t nleaf, const double* r,
                       int64_t* cnt, double* sum, double* device_ms)
{
    USE(h); NEED_DATA(h);
    if (forest < 0 || forest > 1 || !slot || !r || !cnt || !sum || nleaf < 1 || nleaf > 254) return fail(BCFGPU_ERR_BAD_ARG, "bad argument");
    const ForestHost& F = h->fh[forest];
    if (F.p == 0) return fail(BCFGPU_ERR_BAD_ARG, "forest has no covariates");
    const int64_t nbt = (int64_t)F.off[F.p] + F.p;
    // labels -> internal order, uint8, padding = 255
    { int rcp = ensure_host_pos(h); if (rcp) return rcp; }
    std::vector<uint8_t> lab((size_t)h->n_pad, 255);
    for (int64_t i = 0; i < h->n; ++i) lab[h->pos[i]] = (slot[i] >= 0 && slot[i] < nleaf) ? (uint8_t)slot[i] : 255;
    uint8_t* dlab = nullptr; long long* dout = nullptr;
    CK(cudaMalloc((void**)&dlab, (size_t)h->n_pad));
    cudaError_t e = cudaMalloc((void**)&dout, (size_t)nleaf * nbt * 32);
    if (e != cudaSuccess) { cudaFree(dlab); return fail(BCFGPU_ERR_OOM, "op_hist_all: output too large"); }
    CK(cudaMemcpyAsync(dlab, lab.data(), (size_t)h->n_pad, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemsetAsync(h->d_tmp2, 0, (size_t)h->n_pad * 8, h->stream));
    CK(cudaMemcpyAsync(h->d_tmp, r, (size_t)h->n * 8, cudaMemcpyHostToDevice, h->stream));
    k_permute_in<<<(int)std::min<int64_t>((h->n + 255) / 256, 4096), 256, 0, h->stream>>>(h->d_tmp, h->d_pos, h->n, h->d_tmp2);
    CK(cudaMemsetAsync(dout, 0, (size_t)nleaf * nbt * 32, h->stream));
    // two-valued columns (one cutpoint) with at most 8 leaves: k_hist_bits (register accumulators, one predicated fp64 add
    // per observation and column); other columns with at most HIST2_MAX_BINS bins go through the one-pass kernel, the rest
    // through the per-column one.
    // Few (leaf, bin) cells per column: one masked warp sum per cell (bin 0 by subtraction); many: one shared atomic
    // per word and observation (every bin has its cell).
    const bool per_obs = false;   // measured on B200 (n = 1e6, 100 binary columns): the masked sums win up to 16 leaves
    const size_t smem_cap = 160 * 1024;
    std::vector<int32_t> small, large, cellbase, twoval;
    int ncells = 0;
    const int nl2 = nleaf <= 1 ? 1 : nleaf <= 2 ? 2 : nleaf <= 4 ? 4 : 8;     // k_hist_bits is instantiated for 1, 2, 4, 8 leaves
    // two-valued columns, up to 12 leaves: the int8 tensor-core contraction (k_hist_mma: exact, HBM-bound, independent of the
    // number of leaves); BCFGPU_NO_HIST_MMA=1: the CUDA-core kernel of round 1 (k_hist_bits, up to 8 leaves)
    const bool use_mma = nleaf <= HM_MAX_LEAF && F.p8 > 0 && !getenv("BCFGPU_NO_HIST_MMA") && tensor_map_encoder() != nullptr;
    const bool use_bits = !use_mma && nleaf <= 8 && !getenv("BCFGPU_NO_HIST_BITS");
    for (int v = 0; v < F.p; ++v) {
        if ((use_mma || use_bits) && !F.is16[v] && F.ncut[v] == 1) { twoval.push_back(v); continue; }   // two-valued: register / tensor-memory accumulators
        const int add = per_obs ? nleaf * (F.ncut[v] + 1) : nleaf * F.ncut[v];
        const size_t need = per_obs ? (size_t)(ncells + add) * 32 : (size_t)(ncells + add + nleaf) * 16;
        if (!F.is16[v] && F.ncut[v] + 1 <= HIST2_MAX_BINS && nleaf <= HIST2_MAX_LEAF && need <= smem_cap) {
            small.push_back(v); cellbase.push_back(ncells); ncells += add;
        } else large.push_back(v);
    }
    int32_t* dcols = nullptr;
    constexpr int LEAF_TOT = 16;                                              // leaf totals kept for up to 16 leaves x 4 words
    const size_t ncolw = small.size() * 2 + large.size() + twoval.size() + (size_t)F.p8;
    CK(cudaMalloc((void**)&dcols, (ncolw + 4) * 4 + LEAF_TOT * 32));
    int32_t* dsmall = dcols; int32_t* dbase = dcols + small.size(); int32_t* dlarge = dcols + 2 * small.size();
    int32_t* dtwo = dlarge + large.size();
    int32_t* dcolvar = dtwo + twoval.size();                                  // [p8] variable of a uint8 column if it is two-valued, else -1
    long long* dleaf = reinterpret_cast<long long*>(dcols + ((ncolw + 2) & ~(size_t)1));
    if (!twoval.empty()) {
        CK(cudaMemcpyAsync(dtwo, twoval.data(), twoval.size() * 4, cudaMemcpyHostToDevice, h->stream));
        CK(cudaMemsetAsync(dleaf, 0, LEAF_TOT * 32, h->stream));
    }
    std::vector<int32_t> colvar((size_t)std::max(F.p8, 1), -1);
    if (use_mma && !twoval.empty()) {
        for (int v : twoval) colvar[(size_t)F.col_index[v]] = v;
        CK(cudaMemcpyAsync(dcolvar, colvar.data(), (size_t)F.p8 * 4, cudaMemcpyHostToDevice, h->stream));
    }
    if (!small.empty()) {
        CK(cudaMemcpyAsync(dsmall, small.data(), small.size() * 4, cudaMemcpyHostToDevice, h->stream));
        CK(cudaMemcpyAsync(dbase, cellbase.data(), cellbase.size() * 4, cudaMemcpyHostToDevice, h->stream));
    }
    if (!large.empty()) CK(cudaMemcpyAsync(dlarge, large.data(), large.size() * 4, cudaMemcpyHostToDevice, h->stream));
    CK(cudaEventRecord(h->ev0, h->stream));
    if (!twoval.empty() && use_mma) {
        // B operand = the forest's uint8 bins as they lie in HBM: [p8 columns][n_pad observations], one byte each
        const int ncb = std::min(F.p8, HM_MAX_COLS), gy = (F.p8 + ncb - 1) / ncb;
        const int ktiles = (int)(h->n_pad / HM_TILE);             // (n_pad is a multiple of 256)
        const int gx0 = std::max(1, std::min(ktiles, std::max(1, h->num_sms / gy)));
        const int tpc = (ktiles + gx0 - 1) / gx0;
        CUtensorMap tm;
        const cuuint64_t gdim[2] = {(cuuint64_t)h->n_pad, (cuuint64_t)F.p8}, gstr[1] = {(cuuint64_t)h->n_pad};
        const cuuint32_t box[2] = {(cuuint32_t)HM_KT, (cuuint32_t)ncb}, estr[2] = {1u, 1u};
        const CUresult cr = tensor_map_encoder()(&tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, (void*)F.d_bins8, gdim, gstr, box, estr,
                                                 CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                                 CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (cr != CUDA_SUCCESS) { cudaFree(dlab); cudaFree(dout); cudaFree(dcols); return fail(BCFGPU_ERR_CUDA, "op_hist_all: cuTensorMapEncodeTiled failed (" + std::to_string((int)cr) + ")"); }
        // three rings (bcf_hist_mma.cuh): 4 stages for A whenever two B stages still fit beside them (measured at 8 and 12
        // leaves: the A ring matters more than the B ring), and as many B stages as the rest of shared memory holds (6 at 100
        // columns and 1-4 leaves).  EVEN numbers: a stage is then always served by the same producer and the same MMA issuer
        // -- the tiles' parity --, which the parity waits on its barriers rely on.
        int na = hm_smem_bytes(ncb, nleaf, HM_MAX_A_STAGES, 2) <= (size_t)h->max_optin ? HM_MAX_A_STAGES : 2;
        int nb = HM_MAX_B_STAGES;
        while (nb > 2 && hm_smem_bytes(ncb, nleaf, na, nb) > (size_t)h->max_optin) nb -= 2;
        if (const char* e = getenv("BCFGPU_HIST_RINGS")) {   // experiments: "na,nb"
            int a_ = 0, b_ = 0;
            if (sscanf(e, "%d,%d", &a_, &b_) == 2 && (a_ == 2 || a_ == 4) && b_ >= 2 && b_ <= HM_MAX_B_STAGES && b_ % 2 == 0 &&
                hm_smem_bytes(ncb, nleaf, a_, b_) <= (size_t)h->max_optin) { na = a_; nb = b_; }
        }
        const size_t smem = hm_smem_bytes(ncb, nleaf, na, nb);
        if (smem > (size_t)h->max_optin) { cudaFree(dlab); cudaFree(dout); cudaFree(dcols); return fail(BCFGPU_ERR_CUDA, "op_hist_all: k_hist_mma does not fit in shared memory"); }
        dim3 grid((unsigned)((ktiles + tpc - 1) / tpc), (unsigned)gy);
        const int probe = getenv("BCFGPU_HIST_PROBE") ? atoi(getenv("BCFGPU_HIST_PROBE")) : 0;   // timing experiments only
        long long* dtrace = nullptr;
        if (getenv("BCFGPU_HIST_TRACE")) { CK(cudaMalloc((void**)&dtrace, 6 * 64 * 8)); CK(cudaMemsetAsync(dtrace, 0, 6 * 64 * 8, h->stream)); }
        k_hist_mma<<<grid, HM_THREADS, smem, h->stream>>>(tm, dlab, h->d_tmp2, nleaf, ncb, F.p8, ktiles, tpc, na, nb, dcolvar, F.d_off,
                                                          (long long)nbt, dout, dleaf, h->dev.err, h->dev.dbg, probe, dtrace);
        if (dtrace) {   // the timeline of CTA 0, in cycles since its start (a debugging aid: synchronises)
            long long tr[6 * 64];
            CK(cudaMemcpyAsync(tr, dtrace, sizeof(tr), cudaMemcpyDeviceToHost, h->stream));
            CK(cudaStreamSynchronize(h->stream));
            cudaFree(dtrace);
            const long long t0 = tr[5 * 64];
            fprintf(stderr, "[k_hist_mma trace] A stages %d, B stages %d, K-tiles of 256 observations per CTA %d: set-up %lld, accumulator complete %lld, in smem %lld, end %lld\n", na, nb, tpc,
                    tr[5 * 64 + 1] - t0, tr[5 * 64 + 2] - t0, tr[5 * 64 + 3] - t0, tr[5 * 64 + 4] - t0);
            for (int i = 0; i < std::min(tpc, 64); ++i)
                fprintf(stderr, "  tile %2d: bins issued %7lld inputs landed %7lld built %7lld mma-ready %7lld committed %7lld\n", i, tr[i] - t0,
                        tr[64 + i] - t0, tr[128 + i] - t0, tr[192 + i] - t0, tr[256 + i] - t0);
        }
        h->launches++;
    } else if (!twoval.empty()) {
        const int cpc = HISTB_CELLS / nl2, gy = ((int)twoval.size() + cpc - 1) / cpc;
        const int tiles = h->dev.n_tiles;
        const int waves = getenv("BCFGPU_HIST_WAVES") ? std::max(1, atoi(getenv("BCFGPU_HIST_WAVES"))) : 4;
        const int gx = std::max(1, std::min(tiles, (h->num_sms * 2 * waves + gy - 1) / gy));   // four waves of two CTAs per SM (measured best of 1, 2, 4)
        const int tpc = (tiles + gx - 1) / gx;
        dim3 grid((unsigned)((tiles + tpc - 1) / tpc), (unsigned)gy);
        switch (nl2) {
        case 1: k_hist_bits<1><<<grid, 256, 0, h->stream>>>(h->dev, forest, dlab, h->d_tmp2, F.d_off, dtwo, (int)twoval.size(), tpc, dout, dleaf); break;
        case 2: k_hist_bits<2><<<grid, 256, 0, h->stream>>>(h->dev, forest, dlab, h->d_tmp2, F.d_off, dtwo, (int)twoval.size(), tpc, dout, dleaf); break;
        case 4: k_hist_bits<4><<<grid, 256, 0, h->stream>>>(h->dev, forest, dlab, h->d_tmp2, F.d_off, dtwo, (int)twoval.size(), tpc, dout, dleaf); break;
        default: k_hist_bits<8><<<grid, 256, 0, h->stream>>>(h->dev, forest, dlab, h->d_tmp2, F.d_off, dtwo, (int)twoval.size(), tpc, dout, dleaf); break;
        }
        h->launches++;
    }
    if (!small.empty()) {
        // at most 48 tiles per CTA between flushes (32-bit table words), at least ~2 CTAs per SM
        const int tiles = h->dev.n_tiles;
        int tpc = std::max(1, std::min(48, (tiles + h->num_sms * 2 - 1) / (h->num_sms * 2)));
        const int gridx = (tiles + tpc - 1) / tpc;
        const size_t smem = per_obs ? (size_t)ncells * 32 : (size_t)(ncells + nleaf) * 16;
        CK(cudaFuncSetAttribute(k_hist_small, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_cap));   // one value for every handle (the attribute is per device)
        k_hist_small<<<gridx, 256, smem, h->stream>>>(h->dev, forest, dlab, h->d_tmp2, nleaf, F.d_off, dsmall, (int)small.size(), dbase,
                                                       ncells, tpc, dout, per_obs ? 1 : 0);
    }
    if (!large.empty()) {
        const int64_t nchunk = h->n_pad / 16;
        dim3 grid((unsigned)std::max<int64_t>(1, std::min<int64_t>((nchunk + 255) / 256, (int64_t)h->num_sms * 2)), (unsigned)large.size());
        k_hist<<<grid, 256, 0, h->stream>>>(h->dev, forest, dlab, h->d_tmp2, nleaf, F.d_off, dout, dlarge);
    }
    CK(cudaEventRecord(h->ev1, h->stream));
    h->launches += 2;
    std::vector<long long> tot((size_t)nleaf * nbt * 4);
    long long leaf_tot[LEAF_TOT * 4] = {0};
    e = cudaMemcpyAsync(tot.data(), dout, tot.size() * 8, cudaMemcpyDeviceToHost, h->stream);
    if (e == cudaSuccess && !twoval.empty()) e = cudaMemcpyAsync(leaf_tot, dleaf, sizeof(leaf_tot), cudaMemcpyDeviceToHost, h->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
    for (int v : twoval)               // bin 0 of a two-valued column: the leaf's total minus bin 1 (word by word: the representation is linear)
        for (int l = 0; l < nleaf; ++l) {
            long long* o0 = &tot[((size_t)l * nbt + (size_t)F.off[v] + v) * 4];
            for (int q = 0; q < 4; ++q) o0[q] = leaf_tot[l * 4 + q] - o0[4 + q];
        }
    float ms = 0.f;
    if (e == cudaSuccess) cudaEventElapsedTime(&ms, h->ev0, h->ev1);
    cudaFree(dlab); cudaFree(dout); cudaFree(dcols);
    if (e != cudaSuccess) return fail(BCFGPU_ERR_CUDA, std::string("op_hist_all: ") + cudaGetErrorString(e));
    if (device_ms) *device_ms = ms;
    for (size_t k = 0; k < (size_t)nleaf * nbt; ++k) {
        cnt[k] = tot[k * 4];
        sum[k] = limbs_to_double(tot[k * 4 + 1], tot[k * 4 + 2], tot[k * 4 + 3]);
    }
    return check_device_err(h);
}

int bcfgpu_op_lil(const double* sum_phi, const double* sum_phi_r, const double* tau, int32_t m, double* out)
{
    if (!sum_phi || !sum_phi_r || !tau || !out || m < 1) return fail(BCFGPU_ERR_BAD_ARG, "bad argument");
    if (bcfgpu_device_count() == 0) return fail(BCFGPU_ERR_CUDA, "no CUDA device available (libbcfgpu has no CPU fallback)");
    double* d = nullptr;
    CK(cudaMalloc((void**)&d, (size_t)m * 8 * 4));
    CK(cudaMemcpy(d, sum_phi, (size_t)m * 8, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d + m, sum_phi_r, (size_t)m * 8, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d + 2 * m, tau, (size_t)m * 8, cudaMemcpyHostToDevice));
    k_lil<<<(m + 127) / 128, 128>>>(d, d + m, d + 2 * m, m, d + 3 * m);
    CK(cudaGetLastError());
    CK(cudaMemcpy(out, d + 3 * m, (size_t)m * 8, cudaMemcpyDeviceToHost));
    cudaFree(d);
    return BCFGPU_OK;
}

int bcfgpu_op_probit_latent(const double* m, const int32_t* y, const double* u, int32_t n, double* out)
{
    if (!m || !y || !u || !out || n < 1) return fail(BCFGPU_ERR_BAD_ARG, "bad argument");
    if (bcfgpu_device_count() == 0) return fail(BCFGPU_ERR_CUDA, "no CUDA device available (libbcfgpu has no CPU fallback)");
    double* dm = nullptr; int* dyv = nullptr;
    CK(cudaMalloc((void**)&dm, (size_t)n * 8 * 3));
    CK(cudaMalloc((void**)&dyv, (size_t)n * 4));
    CK(cudaMemcpy(dm, m, (size_t)n * 8, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dm + n, u, (size_t)n * 8, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dyv, y, (size_t)n * 4, cudaMemcpyHostToDevice));
    k_probit_probe<<<(n + 127) / 128, 128>>>(dm, dyv, dm + n, n, dm + 2 * (size_t)n);
    CK(cudaGetLastError());
    CK(cudaMemcpy(out, dm + 2 * (size_t)n, (size_t)n * 8, cudaMemcpyDeviceToHost));
    cudaFree(dm); cudaFree(dyv);
    return BCFGPU_OK;
}

int bcfgpu_op_rng_probe(uint64_t seed, uint32_t chain, uint32_t sweep, uint32_t tree, uint32_t purpose, uint32_t idx,
                        double out3[3])
{
    if (!out3) return fail(BCFGPU_ERR_BAD_ARG, "null output");
    if (bcfgpu_device_count() == 0) return fail(BCFGPU_ERR_CUDA, "no CUDA device available (libbcfgpu has no CPU fallback)");
    const uint64_t s = seed + (uint64_t)chain * 0x9E3779B97F4A7C15ull;
    double* d = nullptr;
    CK(cudaMalloc((void**)&d, 24));
    k_rng_probe<<<1, 1>>>((uint32_t)s, (uint32_t)(s >> 32), sweep, tree, purpose, idx, d);
    CK(cudaGetLastError());
    CK(cudaMemcpy(out3, d, 24, cudaMemcpyDeviceToHost));
    cudaFree(d);
    return BCFGPU_OK;
}

}  // extern "C"

// ------------------------------------------------------------------------------------------------
// PERSISTENT engine launch
// ------------------------------------------------------------------------------------------------
static int select_kernel(bcfgpu_handle* h)
{
    if (h->coop_grid == 0) {
        int coop = 0;
        CK(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, h->device));
        if (!coop) return fail(BCFGPU_ERR_CUDA, "device does not support cooperative launch");
        const int max_optin = h->max_optin;
        // One CTA per SM whatever n: a CTA without observation tiles still replays the decisions and takes its share of the
        // per-tree work that is dealt by CTA (proposals, tree write-back).  Measured (k_pipe): n = 500: 930 -> 1530
        // sweeps/s, n = 5e4: 1275 -> 1543 against a grid of one CTA per tile.
        int grid = h->num_sms * PERSIST_CTAS_PER_SM;
        if (h->cfg.max_ctas > 0) grid = std::min<int>(grid, h->cfg.max_ctas);      // a share of the GPU (fits running side by side)
        if (getenv("BCFGPU_GRID_BY_TILES")) grid = std::min<int>(grid, std::max(1, h->dev.n_tiles));
        if (const char* mc = getenv("BCFGPU_MAX_CTAS")) grid = std::max(1, std::min(grid, atoi(mc)));   // tests: many tiles per CTA at small n
        const int ntl_max = (h->dev.n_tiles + grid - 1) / grid;
        h->res_ntl = 0;
        h->kernel = 0;
        const bool want_batched = h->engine == BCFGPU_ENGINE_BATCHED || h->engine == BCFGPU_ENGINE_PIPELINED || h->cfg.engine == BCFGPU_ENGINE_AUTO;
        const bool may_reside = h->engine != BCFGPU_ENGINE_PERSISTENT_HBM && !getenv("BCFGPU_NO_RESIDENT");
        // does this CTA's share of the observation state fit in shared memory next to the kernel's scratch?
        auto fits = [&](const void* fn, size_t need, int threads = PTHREADS) -> int {
            cudaFuncAttributes fa;
            if (cudaFuncGetAttributes(&fa, fn) != cudaSuccess) return 0;
            if (need + fa.sharedSizeBytes > (size_t)max_optin) return 0;   // (the attribute was raised to the maximum in bcfgpu_create)
            int per_sm = 0;
            if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, threads, need) != cudaSuccess) return 0;
            return per_sm >= 1;
        };
        if (want_batched) {
            const size_t need = BATCH_SMEM_BYTES + (size_t)ntl_max * BAT_BYTES_PER_TILE;
            if (may_reside && fits((const void*)k_batched<true>, need)) { h->kernel = 3; h->res_ntl = ntl_max; h->res_smem = need; }
            else if (fits((const void*)k_batched<false>, BATCH_SMEM_BYTES)) {
                // the SM's observations do not fit in shared memory: same kernel with the state in HBM / L2
                h->kernel = 5; h->res_ntl = 0; h->res_smem = BATCH_SMEM_BYTES;
                if (!h->dev.keys) {
                    DM(h, &h->dev.keys, (size_t)h->dev.n_pad * NB);
                }
            }
            // the pipelined kernel on top of it (the batched one then serves the sweeps that hold a wide tree)
            if ((h->kernel == 3 || h->kernel == 5) && may_reside && h->engine == BCFGPU_ENGINE_PIPELINED && h->T <= 4096 && (h->nranks == 1 || h->peer_ok) && !h->no_pipe) {
                h->fallback_kernel = h->kernel;
                const size_t needp = PIPE_SMEM_BYTES + (size_t)ntl_max * PIPE_BYTES_PER_TILE + (size_t)h->T * sizeof(TMeta) + 16;
                if (fits(h->fuse_latent ? (const void*)k_pipe<true> : (const void*)k_pipe<false>, needp, PIPE_THREADS)) { h->kernel = 4; h->pipe_smem = needp; h->pipe_ntl = ntl_max; }
                else if (h->cfg.engine == BCFGPU_ENGINE_PIPELINED)
                    return fail(BCFGPU_ERR_BAD_ARG, "BCFGPU_ENGINE_PIPELINED: the observations of one SM do not fit in shared memory at this n");
            }
        }
        if (h->kernel == 0 && may_reside) {
            const size_t need = STEP_SMEM_BYTES + (size_t)ntl_max * RES_BYTES_PER_TILE;
            if (fits((const void*)k_resident, need)) { h->kernel = 2; h->res_ntl = ntl_max; h->res_smem = need; }
        }
        if (h->kernel == 0) {
            int per_sm = 0;
            CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_persistent, PTHREADS, STEP_SMEM_BYTES));
            if (per_sm < 1) return fail(BCFGPU_ERR_CUDA, "persistent kernel does not fit on an SM");
            h->kernel = 1;
        }
        cudaGetLastError();
        h->coop_grid = grid;
    }
    return BCFGPU_OK;
}

static int run_persistent(bcfgpu_handle* h, int total, bool no_sync)
{
    if (total <= 0) return BCFGPU_OK;
    { int rc = select_kernel(h); if (rc) return rc; }
    // keep launches short enough to stay interruptible (R_CheckUserInterrupt between batches)
    const int batch = 256;
    for (int done = 0; done < total;) {
        if (h->stop.load(std::memory_order_relaxed)) {
            cudaStreamSynchronize(h->stream);
            return fail(BCFGPU_ERR_INTERRUPTED, "run stopped by bcfgpu_request_stop");
        }
        int nsw = std::min(batch, total - done);
        Dev d = h->dev;
        d.pr.seq0 = (++h->launch_id) << 32;
        unsigned int* bar = h->d_bar;
        CK(cudaMemsetAsync(bar, 0, 64, h->stream));
        int ntl = h->res_ntl;
        if (h->kernel == 4) {
            // pipelined kernel; it returns early at a sweep that holds a tree wider than KB keys: that one sweep is
            // run by the batched kernel, then the pipelined one resumes
            int* dn = h->d_done;
            if (!no_sync) CK(cudaMemsetAsync(dn, 0, 12, h->stream));
            d.pr.batch0 = h->pipe_batches;
            int pntl = h->pipe_ntl;
            void* args[] = {(void*)&d, (void*)&nsw, (void*)&bar, (void*)&pntl, (void*)&dn};
            CK(cudaLaunchCooperativeKernel(h->fuse_latent ? (const void*)k_pipe<true> : (const void*)k_pipe<false>, dim3(h->coop_grid), dim3(PIPE_THREADS), args, h->pipe_smem, h->stream));
            h->launches++;
            h->latent_drawn = h->fuse_latent;         // probit BART: k_pipe<true>'s finishing pass draws the next sweep's latent
            if (no_sync) {   // speculative (stepwise()): the caller zeroed d_done, checks it later and reruns what was refused
                done += nsw;
                continue;
            }
            int dn_host[2] = {0, 0};
            CK(cudaMemcpyAsync(dn_host, dn, 8, cudaMemcpyDeviceToHost, h->stream));
            CK(cudaStreamSynchronize(h->stream));
            const int ndone = dn_host[0];
            h->pipe_batches += (unsigned long long)dn_host[1];
            h->pipe_sweeps += ndone;
            done += ndone;
            if (ndone < nsw) {
                h->fallback_sweeps++;
                int one = 1;
                d.pr.seq0 = (++h->launch_id) << 32;                  // (its own launch: the mailbox sequence numbers restart)
                CK(cudaMemsetAsync(bar, 0, 16, h->stream));
                const int* none = nullptr;
                void* a2[] = {(void*)&d, (void*)&one, (void*)&bar, (void*)&ntl, (void*)&none};
                CK(cudaLaunchCooperativeKernel(h->fallback_kernel == 3 ? (const void*)k_batched<true> : (const void*)k_batched<false>, dim3(h->coop_grid),
                                               dim3(PTHREADS), a2, h->res_smem, h->stream));
                h->launches++;
                done += 1;
                if (h->fuse_latent) {   // (the batched kernel does not draw: the sweep behind it gets its latent from the stand-alone kernel)
                    k_probit_latent<<<h->grid_fin, 256, 0, h->stream>>>(d, h->d_ybin, h->d_orig, 0, h->cfg.probit_offset, const_cast<double*>(d.y), nullptr);
                    h->launches++;
                }
            }
            continue;
        }
        h->latent_drawn = false;                      // (the other sweep kernels leave the draw to k_probit_latent)
        const int* none = nullptr;
        if (h->kernel == 3) {
            void* args[] = {(void*)&d, (void*)&nsw, (void*)&bar, (void*)&ntl, (void*)&none};
            CK(cudaLaunchCooperativeKernel((const void*)k_batched<true>, dim3(h->coop_grid), dim3(PTHREADS), args, h->res_smem, h->stream));
        } else if (h->kernel == 5) {
            void* args[] = {(void*)&d, (void*)&nsw, (void*)&bar, (void*)&ntl, (void*)&none};
            CK(cudaLaunchCooperativeKernel((const void*)k_batched<false>, dim3(h->coop_grid), dim3(PTHREADS), args, h->res_smem, h->stream));
        } else if (h->kernel == 2) {
            void* args[] = {(void*)&d, (void*)&nsw, (void*)&bar, (void*)&ntl};
            CK(cudaLaunchCooperativeKernel((const void*)k_resident, dim3(h->coop_grid), dim3(PTHREADS), args, h->res_smem, h->stream));
        } else {
            void* args[] = {(void*)&d, (void*)&nsw, (void*)&bar};
            CK(cudaLaunchCooperativeKernel((const void*)k_persistent, dim3(h->coop_grid), dim3(PTHREADS), args, STEP_SMEM_BYTES, h->stream));
        }
        h->launches++;
        done += nsw;
    }
    return BCFGPU_OK;
}

// ------------------------------------------------------------------------------------------------
// out-of-sample evaluation of the saved forests (bcf 2.x's predict(): tau(x), mu(x) at new covariate rows)
// ------------------------------------------------------------------------------------------------
static int bin_new_rows(bcfgpu_handle* h, const ForestHost& F, const double* x, int64_t n, uint16_t* d_out)
{
    const size_t colb = (size_t)n * 8;
    const int chunk = (int)std::max<size_t>(1, std::min<size_t>((size_t)F.p, STAGE_BYTES / colb));
    Block stage;
    if (dev_alloc(h->device, (size_t)chunk * colb, &stage) != cudaSuccess) return fail(BCFGPU_ERR_OOM, "staging buffer of predict");
    cudaError_t e = cudaSuccess;
    for (int c0 = 0; c0 < F.p && e == cudaSuccess; c0 += chunk) {
        const int nc = std::min(chunk, F.p - c0);
        e = cudaMemcpyAsync(stage.p, x + (size_t)c0 * n, (size_t)nc * colb, cudaMemcpyHostToDevice, h->stream);
        if (e != cudaSuccess) break;
        dim3 grid((unsigned)std::max<int64_t>(1, std::min<int64_t>((n + 255) / 256, 4096)), (unsigned)nc);
        k_bin_plain<<<grid, 256, 0, h->stream>>>((const double*)stage.p, n, nc, F.d_cuts, F.d_off + c0, d_out + (size_t)c0 * n);
        h->launches++;
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
    dev_free(h->device, stage);
    if (e != cudaSuccess) return fail(BCFGPU_ERR_CUDA, std::string("predict: binning the new rows: ") + cudaGetErrorString(e));
    return BCFGPU_OK;
}

extern "C" int bcfgpu_num_saved_forests(bcfgpu_handle* h, int32_t* ndraws, int64_t* nnodes)
{
    USE(h); NEED_DATA(h);
    if (ndraws) *ndraws = h->forests_saved;
    if (nnodes) {
        unsigned long long used = 0;
        if (h->d_pool_used) CK(cudaMemcpy(&used, h->d_pool_used, 8, cudaMemcpyDeviceToHost));
        *nnodes = (int64_t)used;
    }
    return BCFGPU_OK;
}

extern "C" int bcfgpu_predict(bcfgpu_handle* h, int64_t n_new, const double* xcon_new, const double* xmod_new, double* tau_mean,
                              double* tau_var, double* mu_mean, double* tau_draws, double* mu_draws)
{
    USE(h); NEED_DATA(h);
    if (n_new < 1 || !xcon_new) return fail(BCFGPU_ERR_BAD_ARG, "bad argument");
    if (h->forests_saved < 1 || !h->d_pool) return fail(BCFGPU_ERR_STATE, "no saved forests: run with keep_draws bit 3 (8) set and nsim >= 1");
    const ForestHost &Fc = h->fh[0], &Fm = h->fh[1];
    if (Fm.p > 0 && !xmod_new) return fail(BCFGPU_ERR_BAD_ARG, "the moderate forest needs x_moderate rows");
    CK(cudaStreamSynchronize(h->stream));
    unsigned long long used = 0;
    CK(cudaMemcpy(&used, h->d_pool_used, 8, cudaMemcpyDeviceToHost));
    if ((long long)used > h->pool_cap) return fail(BCFGPU_ERR_OOM, "the saved forests outgrew their node pool (trees much larger than expected)");
    const int ns = h->forests_saved;
    Block bc{nullptr, 0}, bm{nullptr, 0}, acc{nullptr, 0}, dr{nullptr, 0};
    int rc = BCFGPU_OK;
    auto cleanup = [&]() { dev_free(h->device, bc); dev_free(h->device, bm); dev_free(h->device, acc); dev_free(h->device, dr); };
    if (dev_alloc(h->device, (size_t)Fc.p * n_new * 2, &bc) != cudaSuccess || dev_alloc(h->device, (size_t)std::max(Fm.p, 1) * n_new * 2, &bm) != cudaSuccess ||
        dev_alloc(h->device, (size_t)n_new * 8 * 3, &acc) != cudaSuccess) { cleanup(); return fail(BCFGPU_ERR_OOM, "predict: device buffers"); }
    if ((rc = bin_new_rows(h, Fc, xcon_new, n_new, (uint16_t*)bc.p)) || (Fm.p > 0 && (rc = bin_new_rows(h, Fm, xmod_new, n_new, (uint16_t*)bm.p)))) { cleanup(); return rc; }
    double* a = (double*)acc.p;
    cudaError_t e = cudaMemsetAsync(a, 0, (size_t)n_new * 8 * 3, h->stream);
    // draws leave in chunks, so that the device never holds more than ~512 MB of them
    const bool want_draws = tau_draws || mu_draws;
    const int per = want_draws ? (int)std::max<int64_t>(1, std::min<int64_t>(ns, ((int64_t)256 << 20) / (n_new * 8))) : ns;
    if (want_draws && dev_alloc(h->device, (size_t)per * n_new * 8 * 2, &dr) != cudaSuccess) { cleanup(); return fail(BCFGPU_ERR_OOM, "predict: draw buffers"); }
    double* dt = (double*)dr.p;
    double* dm = dt ? dt + (size_t)per * n_new : nullptr;
    const unsigned grid = (unsigned)((n_new + 127) / 128);
    for (int d0 = 0; d0 < ns && e == cudaSuccess; d0 += per) {
        const int d1 = std::min(ns, d0 + per);
        k_predict<<<grid, 128, 0, h->stream>>>(n_new, h->T, h->cfg.ntree_con, d0, d1, h->d_pool, h->d_tree_off, h->d_draw_scal,
                                               (const uint16_t*)bc.p, (const uint16_t*)bm.p, a, a + n_new, a + 2 * n_new,
                                               tau_draws ? dt - (size_t)d0 * n_new : nullptr, mu_draws ? dm - (size_t)d0 * n_new : nullptr);
        h->launches++;
        e = cudaGetLastError();
        if (e == cudaSuccess && tau_draws) e = cudaMemcpyAsync(tau_draws + (size_t)d0 * n_new, dt, (size_t)(d1 - d0) * n_new * 8, cudaMemcpyDeviceToHost, h->stream);
        if (e == cudaSuccess && mu_draws) e = cudaMemcpyAsync(mu_draws + (size_t)d0 * n_new, dm, (size_t)(d1 - d0) * n_new * 8, cudaMemcpyDeviceToHost, h->stream);
        if (e == cudaSuccess && want_draws) e = cudaStreamSynchronize(h->stream);
    }
    std::vector<double> host((size_t)n_new * 3);
    if (e == cudaSuccess) e = cudaMemcpyAsync(host.data(), a, (size_t)n_new * 8 * 3, cudaMemcpyDeviceToHost, h->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
    cleanup();
    if (e != cudaSuccess) return fail(BCFGPU_ERR_CUDA, std::string("predict: ") + cudaGetErrorString(e));
    for (int64_t i = 0; i < n_new; ++i) {
        const double m = host[(size_t)i] / ns;
        if (tau_mean) tau_mean[i] = m;
        if (tau_var) tau_var[i] = ns > 1 ? std::max(0.0, (host[(size_t)n_new + i] - ns * m * m) / (ns - 1)) : 0.0;
        if (mu_mean) mu_mean[i] = host[(size_t)2 * n_new + i] / ns;
    }
    return BCFGPU_OK;
}

// ------------------------------------------------------------------------------------------------
// several chains in one call (bcf 2.x's n_chains): one host thread per chain, chain c on devices[c % n_devices]
// ------------------------------------------------------------------------------------------------
extern "C" int bcfgpu_fit_chains(const bcfgpu_config* cfg, int32_t n_chains, const int32_t* devices, int32_t n_devices,
                                 int64_t n, int32_t p_con, int32_t p_mod, const double* y, const int32_t* z,
                                 const double* xcon, const double* xmod, const double* cuts_con, const int32_t* cut_off_con,
                                 const double* cuts_mod, const int32_t* cut_off_mod, int32_t nburn, int32_t nsim, int32_t nthin,
                                 int (*poll)(void*), void* poll_ctx, bcfgpu_handle** handles_out)
{
    if (!cfg || !handles_out || n_chains < 1 || n_chains > 4096) return fail(BCFGPU_ERR_BAD_ARG, "bad argument");
    if (n_devices < 0 || (n_devices > 0 && !devices)) return fail(BCFGPU_ERR_BAD_ARG, "bad device list");
    for (int c = 0; c < n_chains; ++c) handles_out[c] = nullptr;
    std::vector<int> rc((size_t)n_chains, BCFGPU_OK);
    std::vector<std::string> msg((size_t)n_chains);
    std::vector<std::thread> th;
    std::atomic<int> running{n_chains};
    // chains that share a device run one after the other on it (each launch fills the GPU) unless the caller gave them a
    // share of the SMs (cfg.max_ctas): one worker per device walks its chains in order
    const int nd = n_devices > 0 ? n_devices : 1;
    const bool side_by_side = cfg->max_ctas > 0;
    auto one_chain = [&](int c) {
        bcfgpu_config cc = *cfg;
        cc.chain = cfg->chain + (uint32_t)c;
        if (n_devices > 0) cc.device = devices[c % n_devices];
        int r = bcfgpu_create(&cc, &handles_out[c]);
        if (!r) r = bcfgpu_set_data(handles_out[c], n, p_con, p_mod, y, z, xcon, xmod, cuts_con, cut_off_con, cuts_mod, cut_off_mod);
        if (!r) r = bcfgpu_run(handles_out[c], nburn, nsim, nthin);
        rc[c] = r;
        if (r) msg[c] = g_err;
        running.fetch_sub(1);
    };
    if (side_by_side) for (int c = 0; c < n_chains; ++c) th.emplace_back(one_chain, c);
    else for (int dv = 0; dv < std::min(nd, n_chains); ++dv) th.emplace_back([&, dv]() { for (int c = dv; c < n_chains; c += nd) one_chain(c); });
    bool stopped = false;
    while (running.load() > 0) {      // the calling thread polls the caller's interrupt hook (R_CheckUserInterrupt, ...)
        std::this_thread::sleep_for(std::chrono::milliseconds(poll ? 20 : 50));
        if (poll && !stopped && poll(poll_ctx)) {
            stopped = true;
            for (int c = 0; c < n_chains; ++c) if (handles_out[c]) handles_out[c]->stop.store(1);
        }
        if (stopped) for (int c = 0; c < n_chains; ++c) if (handles_out[c]) handles_out[c]->stop.store(1);   // chains created since
    }
    for (auto& t : th) t.join();
    int first = -1;
    for (int c = 0; c < n_chains; ++c) if (rc[c] && first < 0) first = c;
    if (first >= 0 || stopped) {
        for (int c = 0; c < n_chains; ++c) { if (handles_out[c]) bcfgpu_destroy(handles_out[c]); handles_out[c] = nullptr; }
        if (stopped) return fail(BCFGPU_ERR_INTERRUPTED, "interrupted");
        return fail(rc[first], "chain " + std::to_string(first) + ": " + msg[first]);
    }
    return BCFGPU_OK;
}

// ------------------------------------------------------------------------------------------------
// dense pre-steps of bcf_iv() / bcf()
// ------------------------------------------------------------------------------------------------
#include "bcfgpu_glm.inl"
