/*
 * engine.cu -- lifecycle and sweep orchestration of the B200 Gibbs engine (layer 2 of
 * include/pht_b200.h).  One engine = one GPU = one shard of the observations.
 *
 * A sweep is the body of the reference's iteration loop (src/PHT_MCMC_Aslett.c:268-405):
 *     k_assemble -> [k_spectral] -> path kernel of the selected method -> [all-reduce of
 *     the statistics block over NCCL] -> k_update
 * enqueued on one stream with no host synchronisation in between, and captured once into
 * a CUDA graph that is replayed per sweep (the sweep index lives in device memory).
 */
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cstdarg>
#include <cmath>
#include <vector>
#include <thread>
#include <dlfcn.h>
#include <unistd.h>
#include <time.h>
#include "engine_internal.h"

/* ---------------------------------------------------------------- errors */
static thread_local char g_err[512] = "";
static int fail(const char *fmt, ...) {
    va_list ap; va_start(ap, fmt); vsnprintf(g_err, sizeof(g_err), fmt, ap); va_end(ap);
    return -1;
}
extern "C" const char *pht_last_error(void) { return g_err; }
#define CU(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return fail("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); } while (0)

/* ---------------------------------------------------------------- NCCL (loaded lazily) */
typedef struct ncclComm *ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
struct NcclApi {
    void *h = nullptr;
    int (*GetUniqueId)(ncclUniqueId *) = nullptr;
    int (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    int (*CommDestroy)(ncclComm_t) = nullptr;
    int (*AllReduce)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
};
static NcclApi g_nccl;
static int nccl_load() {
    if (g_nccl.h) return 0;
    const char *names[] = { "libnccl.so.2", "libnccl.so" };
    for (const char *nm : names) { g_nccl.h = dlopen(nm, RTLD_NOW | RTLD_GLOBAL); if (g_nccl.h) break; }
    if (!g_nccl.h) return fail("NCCL not found: %s", dlerror());
    g_nccl.GetUniqueId = (int (*)(ncclUniqueId *))dlsym(g_nccl.h, "ncclGetUniqueId");
    g_nccl.CommInitRank = (int (*)(ncclComm_t *, int, ncclUniqueId, int))dlsym(g_nccl.h, "ncclCommInitRank");
    g_nccl.CommDestroy = (int (*)(ncclComm_t))dlsym(g_nccl.h, "ncclCommDestroy");
    g_nccl.AllReduce = (int (*)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t))dlsym(g_nccl.h, "ncclAllReduce");
    g_nccl.GetErrorString = (const char *(*)(int))dlsym(g_nccl.h, "ncclGetErrorString");
    if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.CommDestroy || !g_nccl.AllReduce)
        return fail("NCCL symbols missing");
    return 0;
}
static const int NCCL_INT64 = 4, NCCL_SUM = 0;     /* ncclInt64, ncclSum (stable enum values of nccl.h) */

/* ---------------------------------------------------------------- engine */
struct pht_engine {
    pht_config cfg;
    std::vector<int> T; std::vector<double> C, nu, zeta;
    long l_local = 0;
    cudaStream_t stream = nullptr;
    /* device buffers */
    double *d_y = nullptr; uint8_t *d_cens = nullptr;
    double *d_ys = nullptr; uint8_t *d_cs = nullptr; uint32_t *d_perm = nullptr;   /* MHRS: the same observations by decreasing y */
    uint32_t *d_glist = nullptr; XchgWindow *d_xw = nullptr;                        /* MHRS global tail */
    XchgWindow *xpeer[PHT_MAX_WORLD] = {}; std::vector<void *> ipc_opened; bool peers_attached = false;
    bool peer_reduce = false;          /* statistics all-reduced by k_allreduce_peer through the windows (else NCCL, if a communicator exists) */
    uint32_t k_switch = 1u << 15;
    double *d_model = nullptr; long long *d_stats = nullptr; DevState *d_state = nullptr;
    int *d_T = nullptr; double *d_C = nullptr, *d_nu = nullptr, *d_zeta = nullptr;
    int *d_var_ptr = nullptr, *d_cell_i = nullptr, *d_cell_j = nullptr;
    TailItem *d_items = nullptr; uint32_t *d_pend0 = nullptr, *d_pend1 = nullptr, *d_done = nullptr;
    unsigned long long *d_found = nullptr; uint32_t item_cap = 0;
    double *d_res = nullptr; int res_rows = 0;
    double *d_beta = nullptr, *d_pires = nullptr;      /* Dirichlet prior of pi (n) and the pi draws of the last run (res_rows x n) */
    uint32_t *d_idx_exact = nullptr, *d_idx_cens = nullptr; unsigned long long n_exact = 0, n_cens = 0;   /* ECS launch lists */
    std::vector<uint8_t> h_cens;       /* host copy of the flags (ECS parity ranges) */
    double *d_inject = nullptr;        /* host-supplied evals | Q | Qinv (parity hook), else nullptr */
    void *d_flush = nullptr; size_t flush_bytes = 0;   /* L2 flush scratch (measurement aid) */
    ModelLayout L;
    int grid_blocks = 0, tail_blocks = 0, replay_blocks = 0;
    uint4 *d_recs = nullptr;           /* MHRS search -> replay records, one per observation position */
    /* graph */
    cudaGraphExec_t graph_exec = nullptr; int graph_res_rows = -1; double *graph_res = nullptr;
    unsigned long long graph_launches = 0;      /* kernels one replay of the captured sweep launches */
    /* nccl */
    ncclComm_t comm = nullptr;
    /* timing */
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    std::vector<cudaEvent_t> kev; int kev_used = 0;
    unsigned long long launches = 0;
};

extern "C" int pht_device_count(void) {
    int c = 0;
    if (cudaGetDeviceCount(&c) != cudaSuccess) { cudaGetLastError(); return 0; }
    return c;
}

extern "C" int pht_choose_zbits(double sum_y) {
    if (!(sum_y > 0.0) || !std::isfinite(sum_y)) return 30;
    int b = 62 - (int)std::ceil(std::log2(16.0 * sum_y + 1.0));
    if (b > 52) b = 52;
    if (b < 0) b = 0;
    return b;
}

static SweepParams sweep_params(pht_engine *e) {
    SweepParams p; memset(&p, 0, sizeof(p));
    p.y = e->d_y; p.cens = e->d_cens; p.l_local = e->l_local;
    p.obs_rank = (uint32_t)e->cfg.rank; p.obs_world = (uint32_t)e->cfg.world;
    p.model = e->d_model; p.stats = e->d_stats; p.state = e->d_state;
    p.n = e->cfg.n; p.m = e->cfg.m; p.mhit = e->cfg.mhit; p.zbits = e->cfg.zbits;
    p.k0 = (uint32_t)e->cfg.seed; p.k1 = (uint32_t)(e->cfg.seed >> 32);
    pht_roundkeys_init(&p.rk, p.k0, p.k1);
    p.items = e->d_items; p.pend0 = e->d_pend0; p.pend1 = e->d_pend1; p.done = e->d_done; p.found = e->d_found;
    p.item_cap = e->item_cap; p.mhrs_cap = e->cfg.mhrs_cap;
    if (e->d_ys) { p.y = e->d_ys; p.cens = e->d_cs; p.perm = e->d_perm; }      /* production layout: decreasing y */
    p.glist = e->d_glist; p.k_switch = e->k_switch; p.recs = e->d_recs;
    if (e->peers_attached) { p.xw = e->d_xw; for (int r = 0; r < e->cfg.world && r < PHT_MAX_WORLD; r++) p.xpeer[r] = e->xpeer[r]; }
    return p;
}
static UpdateParams update_params(pht_engine *e, double *res, int res_rows) {
    UpdateParams u; memset(&u, 0, sizeof(u));
    u.model = e->d_model; u.stats = e->d_stats; u.state = e->d_state; u.res = res; u.res_rows = res_rows;
    u.T = e->d_T; u.C = e->d_C; u.nu = e->d_nu; u.zeta = e->d_zeta;
    u.var_ptr = e->d_var_ptr; u.cell_i = e->d_cell_i; u.cell_j = e->d_cell_j;
    u.n = e->cfg.n; u.m = e->cfg.m; u.zbits = e->cfg.zbits;
    u.k0 = (uint32_t)e->cfg.seed; u.k1 = (uint32_t)(e->cfg.seed >> 32);
    u.beta = e->d_beta; u.pires = (e->d_beta && res) ? e->d_pires : nullptr;
    return u;
}

/* ---- device memory for the large per-call buffers (observations, sorted layout, records, tail lists): stream-ordered
 * allocations from the device's default pool, whose release threshold is raised so that the memory of a finished call
 * is reused by the next one.  Two reasons: a call allocates ~0.5 GB, and once peer access is on (several GPUs in one
 * process) every plain cudaMalloc/cudaFree also maps/unmaps the block on the peers -- 100..700 ms per call measured on
 * 2 GPUs.  Pool memory is not peer-mapped; the exchange window, which must be, stays a plain allocation.
 * PHT_B200_NO_POOL=1 goes back to cudaMalloc/cudaFree; pht_release_device_memory() hands the cached memory back. */
static bool pool_off() { static const bool off = getenv("PHT_B200_NO_POOL") != nullptr; return off; }
cudaError_t pht_dev_alloc(void **p, size_t bytes, cudaStream_t st) {
    if (pool_off()) return cudaMalloc(p, bytes);
    return cudaMallocAsync(p, bytes, st);
}
cudaError_t pht_dev_free(void *p, cudaStream_t st) {
    if (!p) return cudaSuccess;
    if (pool_off()) return cudaFree(p);
    return cudaFreeAsync(p, st);
}
static void pool_keep(int dev) {
    cudaMemPool_t pool;
    if (!pool_off() && cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
        unsigned long long thr = ~0ull;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
    }
    cudaGetLastError();
}
extern "C" int pht_release_device_memory(void) {
    int ndev = 0, cur = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess) { cudaGetLastError(); return 0; }
    cudaGetDevice(&cur);
    for (int d = 0; d < ndev; d++) {
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, d) == cudaSuccess) { cudaSetDevice(d); cudaDeviceSynchronize(); cudaMemPoolTrimTo(pool, 0); }
    }
    cudaSetDevice(cur); cudaGetLastError();
    return 0;
}

static int method_of(const pht_config &c) {       /* dispatch priority of src/PHT_MCMC_Aslett.c:325-337 */
    if (c.method & PHT_METHOD_MHRS) return PHT_METHOD_MHRS;
    if (c.method & PHT_METHOD_DCS) return PHT_METHOD_DCS;
    if (c.method & PHT_METHOD_ECS) return PHT_METHOD_ECS;
    if (c.method & PHT_METHOD_MHS_HOBOLTH) return PHT_METHOD_MHS_HOBOLTH;      /* the engine's extension (pht_b200.h) */
    if (c.method & PHT_METHOD_MHS_ASLETT) return PHT_METHOD_MHS_ASLETT;
    return 0;
}

/* enqueue the path kernel of the configured method */
static int enqueue_paths(pht_engine *e, const SweepParams &p, const uint32_t *idx_exact = nullptr, unsigned long long n_exact = 0,
                         const uint32_t *idx_cens = nullptr, unsigned long long n_cens = 0, bool lists_given = false) {
    switch (method_of(e->cfg)) {
    case PHT_METHOD_ECS:
        if (lists_given) CU(pht_launch_ecs(p, e->grid_blocks, idx_exact, n_exact, idx_cens, n_cens, e->stream));
        else CU(pht_launch_ecs(p, e->grid_blocks, e->d_idx_exact, e->n_exact, e->d_idx_cens, e->n_cens, e->stream));
        e->launches += (lists_given ? (n_exact != 0) + (n_cens != 0) : (e->n_exact != 0) + (e->n_cens != 0)) - 1;
        break;
    case PHT_METHOD_MHRS: CU(pht_launch_mhrs(p, e->grid_blocks, e->tail_blocks, e->replay_blocks, e->stream)); e->launches += 2; break;
    case PHT_METHOD_DCS: CU(pht_launch_dcs(p, e->grid_blocks, e->stream)); e->launches++; break;           /* unit machine + complex-spectrum kernel */
    case PHT_METHOD_MHS_HOBOLTH: CU(pht_launch_dcs(p, e->grid_blocks, e->stream, true)); e->launches++; break;
    case PHT_METHOD_MHS_ASLETT:
        /* every observation through one kernel: the window list in parity mode, else the whole shard (identity) */
        if (lists_given) CU(pht_launch_mhs_aslett(p, e->grid_blocks, idx_exact, n_exact, e->stream));
        else CU(pht_launch_mhs_aslett(p, e->grid_blocks, nullptr, (unsigned long long)e->l_local, e->stream));
        break;
    default: return fail("sampling method %d has no kernel in this build", e->cfg.method);
    }
    e->launches++;
    return 0;
}

/* k_assemble (+ spectral data for ECS / DCS): everything the path kernels read from the model block */
static int enqueue_model(pht_engine *e, const UpdateParams &u) {
    CU(pht_launch_assemble(u, e->stream)); e->launches++;
    if (method_of(e->cfg) != PHT_METHOD_MHRS) {
        CU(pht_launch_spectral(u, e->d_inject, e->stream)); e->launches++;
    }
    return 0;
}

static int enqueue_sweep(pht_engine *e, double *res, int res_rows, bool time_kernel) {
    UpdateParams u = update_params(e, res, res_rows);
    if (enqueue_model(e, u)) return -1;
    SweepParams p = sweep_params(e);
    if (time_kernel && e->kev_used + 2 <= (int)e->kev.size()) CU(cudaEventRecord(e->kev[e->kev_used++], e->stream));
    if (enqueue_paths(e, p)) return -1;
    if (time_kernel && (e->kev_used & 1)) CU(cudaEventRecord(e->kev[e->kev_used++], e->stream));
    if (e->peer_reduce) {
        ReduceParams r; memset(&r, 0, sizeof(r));
        r.stats = e->d_stats; r.len = stats_len(e->cfg.n); r.state = e->d_state; r.xw = e->d_xw;
        r.rank = (uint32_t)e->cfg.rank; r.world = (uint32_t)e->cfg.world;
        for (int k = 0; k < e->cfg.world && k < PHT_MAX_WORLD; k++) r.xpeer[k] = e->xpeer[k];
        CU(pht_launch_peer_allreduce(r, e->stream)); e->launches++;
    } else if (e->comm) {
        CU(pht_launch_pack_error(u, e->stream)); e->launches++;
        int rc = g_nccl.AllReduce(e->d_stats, e->d_stats, (size_t)stats_len(e->cfg.n), NCCL_INT64, NCCL_SUM, e->comm, e->stream);
        if (rc != 0) return fail("ncclAllReduce failed: %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "?");
    }
    CU(pht_launch_update(u, e->stream)); e->launches++;
    return 0;
}

extern "C" void pht_engine_destroy(pht_engine *e) {
    if (!e) return;
    const bool timing = getenv("PHT_B200_TIMING") != nullptr;
    struct timespec ts0; clock_gettime(CLOCK_MONOTONIC, &ts0);
    auto stage = [&](const char *what) {
        if (!timing) return;
        struct timespec t; clock_gettime(CLOCK_MONOTONIC, &t);
        fprintf(stderr, "[pht_engine_destroy] %-27s %8.3f ms\n", what, (t.tv_sec - ts0.tv_sec) * 1e3 + (t.tv_nsec - ts0.tv_nsec) * 1e-6);
        ts0 = t;
    };
    cudaSetDevice(e->cfg.device);
    if (e->stream) cudaStreamSynchronize(e->stream);
    /* the graph holds captured NCCL work: it must go before the communicator it refers to */
    if (e->graph_exec) { cudaGraphExecDestroy(e->graph_exec); e->graph_exec = nullptr; }
    stage("sync, graph");
    if (e->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(e->comm);
    for (cudaEvent_t ev : e->kev) cudaEventDestroy(ev);
    if (e->ev0) cudaEventDestroy(e->ev0);
    if (e->ev1) cudaEventDestroy(e->ev1);
    for (void *q : e->ipc_opened) cudaIpcCloseMemHandle(q);
    stage("events, ipc");
    void *big[] = { e->d_recs, e->d_ys, e->d_cs, e->d_perm, e->d_y, e->d_cens, e->d_items /* arena: found, pend0, pend1, done live inside */, e->d_idx_exact, e->d_idx_cens,
                    e->d_glist, e->d_model, e->d_stats, e->d_state, e->d_T, e->d_C, e->d_nu, e->d_zeta, e->d_var_ptr, e->d_cell_i, e->d_cell_j, e->d_pires, e->d_res };
    for (void *b : big) pht_dev_free(b, e->stream);
    if (e->stream) cudaStreamSynchronize(e->stream);
    stage("pool frees");
    void *bufs[] = { e->d_xw /* mapped by the peers: a plain allocation */, e->d_beta, e->d_inject, e->d_flush };
    for (void *b : bufs) if (b) cudaFree(b);
    stage("plain frees");
    if (e->stream) cudaStreamDestroy(e->stream);
    stage("stream");
    delete e;
}

extern "C" int pht_engine_create(pht_engine **out, const pht_config *cfg, const double *y_local,
                                 const int *cens_local, long l_local) {
    if (!out || !cfg) return fail("null argument");
    *out = nullptr;
    const int n = cfg->n, m = cfg->m, n1 = n + 1;
    if (n < 1 || n > PHT_MAX_PHASES) return fail("n = %d outside 1..%d", n, PHT_MAX_PHASES);
    if (m < 1 || m > n * n1) return fail("m = %d outside 1..n(n+1)", m);
    if (method_of(*cfg) == 0) return fail("unknown sampling method (code = %d)", cfg->method);
    if (cfg->mhit < 0 || cfg->mhit > 65535) return fail("mhit = %d outside 0..65535", cfg->mhit);
    if (cfg->world < 1 || cfg->rank < 0 || cfg->rank >= cfg->world) return fail("bad shard (rank %d of %d)", cfg->rank, cfg->world);
    if (cfg->zbits < 0 || cfg->zbits > 52) return fail("zbits = %d outside 0..52", cfg->zbits);
    if (l_local < 0 || l_local > 0x7fffffffL) return fail("l_local out of range");
    if ((double)l_local * cfg->world > 4.0e9) return fail("global observation index exceeds 32 bits");
    for (int i = 0; i < n1 * n1; i++) if (cfg->T[i] < 0 || cfg->T[i] > m) return fail("T[%d] = %d outside 0..m", i, cfg->T[i]);
    /* the absorbing state has no exits: a parameter in row n would make the update read beyond N and z */
    for (int j = 0; j < n1; j++) if (cfg->T[n + j * n1] != 0) return fail("T[%d, %d] = %d: row n+1 (the absorbing state) must be all zero", n, j, cfg->T[n + j * n1]);
    if (cfg->world > PHT_MAX_WORLD) return fail("world = %d exceeds %d", cfg->world, PHT_MAX_WORLD);

    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { cudaGetLastError(); return fail("no CUDA device available (this library has no CPU path)"); }
    if (cfg->device < 0 || cfg->device >= ndev) return fail("device %d not present (%d devices)", cfg->device, ndev);
    CU(cudaSetDevice(cfg->device));
    /* (cudaGetDeviceProperties takes 2-13 ms per call here; two attributes take microseconds) */
    int cc_major = 0, cc_minor = 0;
    CU(cudaDeviceGetAttribute(&cc_major, cudaDevAttrComputeCapabilityMajor, cfg->device));
    CU(cudaDeviceGetAttribute(&cc_minor, cudaDevAttrComputeCapabilityMinor, cfg->device));
    if (cc_major < 10) return fail("device %d is sm_%d%d; this library is built for sm_100a only", cfg->device, cc_major, cc_minor);

    /* PHT_B200_TIMING=1: stage times of the set-up on stderr (tools/e2e_breakdown.py) */
    const bool timing = getenv("PHT_B200_TIMING") != nullptr;
    struct timespec ts0; clock_gettime(CLOCK_MONOTONIC, &ts0);
    auto stage = [&](const char *what) {
        if (!timing) return;
        struct timespec t; clock_gettime(CLOCK_MONOTONIC, &t);
        fprintf(stderr, "[pht_engine_create] %-28s %8.3f ms\n", what, (t.tv_sec - ts0.tv_sec) * 1e3 + (t.tv_nsec - ts0.tv_nsec) * 1e-6);
        ts0 = t;
    };
    pht_engine *e = new pht_engine();
    e->cfg = *cfg;
    e->T.assign(cfg->T, cfg->T + n1 * n1); e->C.assign(cfg->C, cfg->C + n1 * n1);
    e->nu.assign(cfg->nu, cfg->nu + m); e->zeta.assign(cfg->zeta, cfg->zeta + m);
    e->cfg.T = e->T.data(); e->cfg.C = e->C.data(); e->cfg.nu = e->nu.data(); e->cfg.zeta = e->zeta.data();
    const bool auto_cap = e->cfg.mhrs_cap <= 0;
    if (auto_cap) e->cfg.mhrs_cap = 256;
    e->l_local = l_local;
    e->L = ModelLayout::make(n, m);
#define CUE(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { fail("%s failed: %s", #call, cudaGetErrorString(e_)); pht_engine_destroy(e); return -1; } } while (0)
    CUE(cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking));
    CUE(cudaEventCreate(&e->ev0)); CUE(cudaEventCreate(&e->ev1));

    /* observations: y as is, censoring flags repacked int32 -> uint8 (9 B per path in HBM) */
    const size_t ln = (size_t)(l_local > 0 ? l_local : 1);
    pool_keep(cfg->device);
    CUE(pht_dev_alloc((void **)&e->d_y, ln * sizeof(double), e->stream));
    CUE(pht_dev_alloc((void **)&e->d_cens, ln, e->stream));
    stage("stream, events, 2 mallocs");
    if (l_local > 0) {
        /* the int32 flags are repacked to bytes by two helper threads while this one feeds y to the copy engine (a copy
         * from pageable memory occupies the calling thread): 7 ms each at 1e7 observations, now side by side.
         * (a pageable staging vector: pinning 10 MB per call cost more than the copy it would speed up) */
        std::vector<uint8_t> hc((size_t)l_local);
        uint8_t *hcp = hc.data();
        const long half = l_local / 2;
        std::thread r0([=] { for (long i = 0; i < half; i++) hcp[i] = cens_local[i] != 0; });
        std::thread r1([=] { for (long i = half; i < l_local; i++) hcp[i] = cens_local[i] != 0; });
        const cudaError_t ey = cudaMemcpyAsync(e->d_y, y_local, (size_t)l_local * sizeof(double), cudaMemcpyHostToDevice, e->stream);
        r0.join(); r1.join();
        CUE(ey);
        stage("y upload | flag repack");
        CUE(cudaMemcpyAsync(e->d_cens, hc.data(), (size_t)l_local, cudaMemcpyHostToDevice, e->stream));
        CUE(cudaStreamSynchronize(e->stream));
        stage("flag upload + sync");
        if (method_of(e->cfg) == PHT_METHOD_ECS) e->h_cens.swap(hc);
    }

    /* parameter -> cells CSR in the reference's insertion order (row-major walk of T, src/PHT_MCMC_Aslett.c:210-228) */
    std::vector<int> var_ptr(m + 1, 0), cell_i, cell_j;
    for (int i = 0; i < n1; i++) for (int j = 0; j < n1; j++) { int v = e->T[i + j * n1]; if (v) var_ptr[v]++; }
    for (int v = 0; v < m; v++) var_ptr[v + 1] += var_ptr[v];
    cell_i.resize(var_ptr[m] > 0 ? var_ptr[m] : 1); cell_j.resize(cell_i.size());
    { std::vector<int> fill(var_ptr.begin(), var_ptr.end() - 1);
      for (int i = 0; i < n1; i++) for (int j = 0; j < n1; j++) { int v = e->T[i + j * n1]; if (v) { int c = fill[v - 1]++; cell_i[c] = i; cell_j[c] = j; } } }

    /* the small buffers too come from the pool (a plain cudaFree per buffer cost up to 66 ms per call inside a process that
     * holds other allocations: PHT_B200_TIMING); the host vectors they are filled from live until the synchronize below */
#define SMALL(ptr, bytes) CUE(pht_dev_alloc((void **)&(ptr), (bytes), e->stream))
    SMALL(e->d_model, sizeof(double) * e->L.total); CUE(cudaMemsetAsync(e->d_model, 0, sizeof(double) * e->L.total, e->stream));
    SMALL(e->d_stats, sizeof(long long) * stats_len(n)); CUE(cudaMemsetAsync(e->d_stats, 0, sizeof(long long) * stats_len(n), e->stream));
    SMALL(e->d_state, sizeof(DevState)); CUE(cudaMemsetAsync(e->d_state, 0, sizeof(DevState), e->stream));
    SMALL(e->d_T, sizeof(int) * n1 * n1); CUE(cudaMemcpyAsync(e->d_T, e->T.data(), sizeof(int) * n1 * n1, cudaMemcpyHostToDevice, e->stream));
    SMALL(e->d_C, sizeof(double) * n1 * n1); CUE(cudaMemcpyAsync(e->d_C, e->C.data(), sizeof(double) * n1 * n1, cudaMemcpyHostToDevice, e->stream));
    SMALL(e->d_nu, sizeof(double) * m); CUE(cudaMemcpyAsync(e->d_nu, e->nu.data(), sizeof(double) * m, cudaMemcpyHostToDevice, e->stream));
    SMALL(e->d_zeta, sizeof(double) * m); CUE(cudaMemcpyAsync(e->d_zeta, e->zeta.data(), sizeof(double) * m, cudaMemcpyHostToDevice, e->stream));
    SMALL(e->d_var_ptr, sizeof(int) * (m + 1)); CUE(cudaMemcpyAsync(e->d_var_ptr, var_ptr.data(), sizeof(int) * (m + 1), cudaMemcpyHostToDevice, e->stream));
    SMALL(e->d_cell_i, sizeof(int) * cell_i.size()); CUE(cudaMemcpyAsync(e->d_cell_i, cell_i.data(), sizeof(int) * cell_i.size(), cudaMemcpyHostToDevice, e->stream));
    SMALL(e->d_cell_j, sizeof(int) * cell_j.size()); CUE(cudaMemcpyAsync(e->d_cell_j, cell_j.data(), sizeof(int) * cell_j.size(), cudaMemcpyHostToDevice, e->stream));
#undef SMALL
    CUE(cudaStreamSynchronize(e->stream));
    stage("model buffers");
    if (cfg->world > 1) {
        /* this rank's exchange window: all-reduce slots (every method) and the global MHRS tail (flags 0, found words NONE) */
        CUE(cudaMalloc(&e->d_xw, sizeof(XchgWindow)));
        CUE(cudaMemsetAsync(e->d_xw, 0, sizeof(XchgWindow), e->stream));
        CUE(cudaMemsetAsync(e->d_xw->gfound, 0xFF, sizeof(e->d_xw->gfound), e->stream));
        CUE(cudaStreamSynchronize(e->stream));
    }
    if (method_of(e->cfg) == PHT_METHOD_MHRS) {
        /* Tail work lists: one arena with room for a quarter of the shard (at least 2^20 observations, at most all of
         * them).  In practice ~1 % of the observations plus one per resident lane are handed over; if the lists ever
         * fill up, lanes simply keep their observation (k_mhrs.cu), so the size is a speed matter, not a limit. */
        size_t slots = ln / 4 > ((size_t)1 << 20) ? ln / 4 : ((size_t)1 << 20);
        if (slots > ln) slots = ln;
        if (const char *ev = getenv("PHT_B200_TAIL_SLOTS")) { const long v = atol(ev); if (v >= 1 && (size_t)v < slots) slots = (size_t)v; }
        e->item_cap = (uint32_t)slots;
        unsigned char *arena = nullptr;
        CUE(pht_dev_alloc((void **)&arena, slots * (sizeof(TailItem) + sizeof(unsigned long long) + 3 * sizeof(uint32_t)), e->stream));
        e->d_items = reinterpret_cast<TailItem *>(arena); arena += slots * sizeof(TailItem);
        e->d_found = reinterpret_cast<unsigned long long *>(arena); arena += slots * sizeof(unsigned long long);
        e->d_pend0 = reinterpret_cast<uint32_t *>(arena); arena += slots * sizeof(uint32_t);
        e->d_pend1 = reinterpret_cast<uint32_t *>(arena); arena += slots * sizeof(uint32_t);
        e->d_done = reinterpret_cast<uint32_t *>(arena);
        stage("tail arena");
        /* the production layout: observations by decreasing y (k_sort.cu); parity hooks keep using the upload order */
        if (l_local > 1 && !getenv("PHT_B200_NO_SORT")) {
            CUE(pht_dev_alloc((void **)&e->d_ys, ln * sizeof(double), e->stream)); CUE(pht_dev_alloc((void **)&e->d_cs, ln, e->stream));
            CUE(pht_dev_alloc((void **)&e->d_perm, ln * sizeof(uint32_t), e->stream));
            stage("sorted-layout mallocs");
            CUE(pht_sort_by_y_desc(e->d_y, e->d_cens, l_local, e->d_ys, e->d_cs, e->d_perm, e->stream));
            stage("sort by y");
        }
        /* global tail: the canonical list */
        if (const char *ev = getenv("PHT_B200_KSWITCH")) { const long v = atol(ev); if (v >= 512 && v <= (1l << 24)) e->k_switch = (uint32_t)v; }
        if (cfg->world > 1) CUE(pht_dev_alloc((void **)&e->d_glist, sizeof(uint32_t) * 2 * PHT_MAX_WORLD * PHT_GCAP, e->stream));      /* double buffered */
        CUE(pht_dev_alloc((void **)&e->d_recs, ln * sizeof(uint4), e->stream));
        stage("record list");
        if (pht_mhrs_grid_blocks(cfg->device, n, &e->grid_blocks, &e->tail_blocks, &e->replay_blocks) != 0) e->grid_blocks = 0;
        if (e->grid_blocks <= 0) { fail("MHRS kernel does not fit on the device: %s", cudaGetErrorString(cudaGetLastError())); pht_engine_destroy(e); return -1; }
        if (auto_cap) {
            /* Attempts a lane spends on one observation before it hands it to the tail.  A lane is a slow worker (one
             * attempt is ~9 dependent jump-steps of ~0.9 us each with the SM full), so an observation that runs to the cap
             * occupies its lane for cap x 8 us, and the lane phase cannot end sooner: 2 ms at 256 -- nothing against the
             * 4.3 ms the lane phase of 10^7 observations takes anyway, twice the whole lane phase of a 1.25 x 10^6 shard
             * (one GPU's share of the 8-GPU run).  But the tail is the slower searcher per attempt (24 warps per SM, whole
             * warps per observation), so handing over early only pays at small shards.  Measured (tools/gpu_r2_capsweep.sh,
             * ms per sweep at cap 256 / 64 / 32): 1.25e6 observations 3.57 / 3.07 / 3.20; 2.5e6: 5.05 / 5.07 / 5.43; 5e6:
             * 7.45 / 8.16 / 9.02; 1e7: 13.9 / 15.5 / 17.3.  Rule: cap = 8 x observations per resident lane, as a power of
             * two between 32 and 256.  Results do not depend on it. */
            const double per_lane = (double)l_local / ((double)e->grid_blocks * 256.0);
            int cap = 32; while (cap < 1024 && (double)(cap * 2) <= 8.0 * per_lane) cap *= 2;
            e->cfg.mhrs_cap = cap;
        }
    }
    if (method_of(e->cfg) == PHT_METHOD_ECS) {
        std::vector<uint32_t> ie, ic;
        for (long i = 0; i < l_local; i++) (e->h_cens[i] ? ic : ie).push_back((uint32_t)i);
        e->n_exact = ie.size(); e->n_cens = ic.size();
        CUE(pht_dev_alloc((void **)&e->d_idx_exact, sizeof(uint32_t) * (ie.size() + 1), e->stream)); CUE(pht_dev_alloc((void **)&e->d_idx_cens, sizeof(uint32_t) * (ic.size() + 1), e->stream));
        if (!ie.empty()) CUE(cudaMemcpyAsync(e->d_idx_exact, ie.data(), sizeof(uint32_t) * ie.size(), cudaMemcpyHostToDevice, e->stream));
        if (!ic.empty()) CUE(cudaMemcpyAsync(e->d_idx_cens, ic.data(), sizeof(uint32_t) * ic.size(), cudaMemcpyHostToDevice, e->stream));
        CUE(cudaStreamSynchronize(e->stream));
        e->grid_blocks = pht_ecs_grid_blocks(cfg->device, n);
        if (e->grid_blocks <= 0) { fail("ECS kernels do not fit on the device: %s", cudaGetErrorString(cudaGetLastError())); pht_engine_destroy(e); return -1; }
    }
    if (method_of(e->cfg) == PHT_METHOD_MHS_HOBOLTH || method_of(e->cfg) == PHT_METHOD_MHS_ASLETT) {
        e->grid_blocks = method_of(e->cfg) == PHT_METHOD_MHS_HOBOLTH ? pht_dcs_grid_blocks(cfg->device, n, true) : pht_mhs_aslett_grid_blocks(cfg->device, n);
        if (e->grid_blocks <= 0) { fail("the MH sampler variant does not fit on the device: %s", cudaGetErrorString(cudaGetLastError())); pht_engine_destroy(e); return -1; }
    }
    if (method_of(e->cfg) == PHT_METHOD_DCS) {
        e->grid_blocks = pht_dcs_grid_blocks(cfg->device, n);
        if (e->grid_blocks <= 0) { fail("DCS kernel does not fit on the device: %s", cudaGetErrorString(cudaGetLastError())); pht_engine_destroy(e); return -1; }
    }
    { const double one = 1.0;         /* start distribution e1, as the reference fixes it (src/PHT_MCMC_Aslett.c:190-193) */
      CUE(cudaMemcpyAsync(e->d_model + e->L.pi, &one, sizeof(double), cudaMemcpyHostToDevice, e->stream)); CUE(cudaStreamSynchronize(e->stream)); }
    /* sweep index 1, start-value assembly (src/PHT_MCMC_Aslett.c:268) */
    DevState st; memset(&st, 0, sizeof(st)); st.iter = 1; st.first_assembly = 1;
    CUE(cudaMemcpyAsync(e->d_state, &st, sizeof(st), cudaMemcpyHostToDevice, e->stream));
    CUE(cudaStreamSynchronize(e->stream));
#undef CUE
    stage("occupancy queries, state");
    *out = e;
    return 0;
}

extern "C" int pht_comm_unique_id(void *id128) {
    if (nccl_load()) return -1;
    ncclUniqueId id; int rc = g_nccl.GetUniqueId(&id);
    if (rc != 0) return fail("ncclGetUniqueId failed (%d)", rc);
    memcpy(id128, &id, sizeof(id));
    return 0;
}
extern "C" int pht_engine_comm_init(pht_engine *e, const void *id128) {
    if (!e) return fail("null engine");
    if (e->cfg.world == 1) return 0;
    if (nccl_load()) return -1;
    CU(cudaSetDevice(e->cfg.device));
    ncclUniqueId id; memcpy(&id, id128, sizeof(id));
    int rc = g_nccl.CommInitRank(&e->comm, e->cfg.world, id, e->cfg.rank);
    if (rc != 0) return fail("ncclCommInitRank failed: %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "?");
    /* one collective outside any graph capture: NCCL connects its transports lazily on first use */
    CU(cudaMemsetAsync(e->d_stats, 0, sizeof(long long) * stats_len(e->cfg.n), e->stream));
    rc = g_nccl.AllReduce(e->d_stats, e->d_stats, (size_t)stats_len(e->cfg.n), NCCL_INT64, NCCL_SUM, e->comm, e->stream);
    if (rc != 0) return fail("ncclAllReduce (warm-up) failed: %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "?");
    CU(cudaStreamSynchronize(e->stream));
    return 0;
}

/* ---------------------------------------------------------------- peer exchange windows (MHRS global tail) */
struct PeerHandle { unsigned long long pid; unsigned long long ptr; int device; int valid; cudaIpcMemHandle_t ipc; unsigned long long cfg_hash; };
/* what every rank of a run must agree on (the chain is one chain): key, shape, sampler, fixed-point scale, world size */
static unsigned long long cfg_hash(const pht_config &c) {
    unsigned long long h = 0xcbf29ce484222325ULL;
    const unsigned long long v[] = { c.seed, (unsigned long long)c.n, (unsigned long long)c.m, (unsigned long long)c.method, (unsigned long long)c.mhit,
                                     (unsigned long long)c.zbits, (unsigned long long)c.world };
    for (unsigned long long x : v) for (int b = 0; b < 8; b++) { h ^= (x >> (8 * b)) & 0xffull; h *= 0x100000001b3ULL; }
    return h;
}
static_assert(sizeof(PeerHandle) <= PHT_PEER_HANDLE_BYTES, "peer handle does not fit its ABI slot");

extern "C" int pht_engine_peer_handle(pht_engine *e, void *handle) {
    if (!e || !handle) return fail("null argument");
    memset(handle, 0, PHT_PEER_HANDLE_BYTES);
    PeerHandle h; memset(&h, 0, sizeof(h));
    if (e->d_xw) {
        CU(cudaSetDevice(e->cfg.device));
        h.pid = (unsigned long long)getpid(); h.ptr = (unsigned long long)(uintptr_t)e->d_xw; h.device = e->cfg.device; h.valid = 1;
        h.cfg_hash = cfg_hash(e->cfg);
        CU(cudaIpcGetMemHandle(&h.ipc, e->d_xw));
    }
    memcpy(handle, &h, sizeof(h));
    return 0;
}

extern "C" int pht_engine_peer_attach(pht_engine *e, const void *handles) {
    if (!e || !handles) return fail("null argument");
    if (e->cfg.world == 1 || !e->d_xw) return 0;           /* nothing to share (single rank) */
    CU(cudaSetDevice(e->cfg.device));
    for (int r = 0; r < e->cfg.world; r++) {
        PeerHandle h; memcpy(&h, (const char *)handles + (size_t)r * PHT_PEER_HANDLE_BYTES, sizeof(h));
        if (!h.valid) return fail("rank %d offers no exchange window", r);
        if (h.cfg_hash != cfg_hash(e->cfg)) return fail("rank %d runs a different configuration (seed, n, m, method, mhit, zbits or world differ from rank %d's): the ranks of a run share one chain", r, e->cfg.rank);
        if (r == e->cfg.rank) { e->xpeer[r] = e->d_xw; continue; }
        if (h.pid == (unsigned long long)getpid()) {
            /* same process (one host thread per GPU): the peer's pointer is valid here once peer access is on */
            if (h.device != e->cfg.device) {
                int can = 0; CU(cudaDeviceCanAccessPeer(&can, e->cfg.device, h.device));
                if (!can) return fail("device %d cannot access device %d", e->cfg.device, h.device);
                cudaError_t pe = cudaDeviceEnablePeerAccess(h.device, 0);
                if (pe != cudaSuccess && pe != cudaErrorPeerAccessAlreadyEnabled) return fail("cudaDeviceEnablePeerAccess(%d): %s", h.device, cudaGetErrorString(pe));
                cudaGetLastError();
            }
            e->xpeer[r] = reinterpret_cast<XchgWindow *>((uintptr_t)h.ptr);
        } else {
            void *q = nullptr;
            CU(cudaIpcOpenMemHandle(&q, h.ipc, cudaIpcMemLazyEnablePeerAccess));
            e->ipc_opened.push_back(q);
            e->xpeer[r] = reinterpret_cast<XchgWindow *>(q);
        }
    }
    e->peers_attached = true;
    /* the statistics go through the windows too, unless the launcher asks for the NCCL all-reduce (PHT_B200_NCCL=1) */
    { const char *ev = getenv("PHT_B200_NCCL"); e->peer_reduce = !(ev && *ev && *ev != '0'); }
    if (e->graph_exec) { cudaGraphExecDestroy(e->graph_exec); e->graph_exec = nullptr; }     /* kernel parameters change */
    return 0;
}

extern "C" int pht_engine_set_theta(pht_engine *e, const double *theta, uint32_t next_iter) {
    if (!e || !theta) return fail("null argument");
    CU(cudaSetDevice(e->cfg.device));
    CU(cudaStreamSynchronize(e->stream));
    CU(cudaMemcpy(e->d_model + e->L.theta, theta, sizeof(double) * e->cfg.m, cudaMemcpyHostToDevice));
    DevState st; CU(cudaMemcpy(&st, e->d_state, sizeof(st), cudaMemcpyDeviceToHost));
    st.iter = next_iter; st.first_assembly = 1;
    CU(cudaMemcpy(e->d_state, &st, sizeof(st), cudaMemcpyHostToDevice));
    return 0;
}
extern "C" int pht_engine_get_theta(pht_engine *e, double *theta) {
    if (!e || !theta) return fail("null argument");
    CU(cudaSetDevice(e->cfg.device));
    CU(cudaStreamSynchronize(e->stream));
    CU(cudaMemcpy(theta, e->d_model + e->L.theta, sizeof(double) * e->cfg.m, cudaMemcpyDeviceToHost));
    return 0;
}

extern "C" int pht_engine_set_pi(pht_engine *e, const double *pi, const double *beta) {
    if (!e) return fail("null engine");
    CU(cudaSetDevice(e->cfg.device));
    CU(cudaStreamSynchronize(e->stream));
    const int n = e->cfg.n;
    if (pi) {
        double sum = 0.0;
        for (int i = 0; i < n; i++) { if (!(pi[i] >= 0.0)) return fail("pi[%d] = %g is not a probability", i, pi[i]); sum += pi[i]; }
        if (!(sum > 0.999999 && sum < 1.000001)) return fail("pi sums to %g, not 1", sum);
        CU(cudaMemcpy(e->d_model + e->L.pi, pi, sizeof(double) * n, cudaMemcpyHostToDevice));
    }
    if (e->graph_exec) { cudaGraphExecDestroy(e->graph_exec); e->graph_exec = nullptr; }     /* kernel parameters change */
    if (beta) {
        /* The update needs start states drawn from their conditional law given y.  MHRS does that by construction (a
         * rejection sampler from pi).  The reference's ECS and DCS samplers draw the start state from pi itself
         * (src/Simulate_AbsCTMC_eq_Aslett_ECS.c:231-238, src/Simulate_AbsCTMC_gt_Hobolth_DCS.c:88-95) -- exact only for
         * the degenerate pi the reference always uses -- so B carries no information about pi there. */
        if (method_of(e->cfg) != PHT_METHOD_MHRS) return fail("the start-distribution update needs method MHRS (ECS and DCS draw the start state from pi, not from its conditional law given y)");
        for (int i = 0; i < n; i++) if (!(beta[i] > 0.0)) return fail("beta[%d] = %g: Dirichlet parameters must be positive", i, beta[i]);
        if (!e->d_beta) CU(cudaMalloc(&e->d_beta, sizeof(double) * n));
        CU(cudaMemcpy(e->d_beta, beta, sizeof(double) * n, cudaMemcpyHostToDevice));
    } else if (e->d_beta) { CU(cudaFree(e->d_beta)); e->d_beta = nullptr; }
    return 0;
}
extern "C" int pht_engine_get_pi(pht_engine *e, double *pi) {
    if (!e || !pi) return fail("null argument");
    CU(cudaSetDevice(e->cfg.device));
    CU(cudaStreamSynchronize(e->stream));
    CU(cudaMemcpy(pi, e->d_model + e->L.pi, sizeof(double) * e->cfg.n, cudaMemcpyDeviceToHost));
    return 0;
}
extern "C" int pht_engine_pi_rows(pht_engine *e, int rows, double *out) {
    if (!e || !out || rows < 0) return fail("bad argument");
    if (!e->d_beta || !e->d_pires) return fail("pi is not being inferred (pht_engine_set_pi with a prior first)");
    if (rows > e->res_rows) return fail("%d rows asked, the last run produced at most %d", rows, e->res_rows);
    CU(cudaSetDevice(e->cfg.device));
    CU(cudaStreamSynchronize(e->stream));
    if (rows) CU(cudaMemcpy(out, e->d_pires, sizeof(double) * (size_t)rows * e->cfg.n, cudaMemcpyDeviceToHost));
    return 0;
}

static int ensure_res(pht_engine *e, int rows) {
    if (rows <= e->res_rows) return 0;
    CU(cudaStreamSynchronize(e->stream));
    if (e->d_res) { CU(pht_dev_free(e->d_res, e->stream)); e->d_res = nullptr; }
    CU(pht_dev_alloc((void **)&e->d_res, sizeof(double) * (size_t)rows * e->cfg.m, e->stream));
    if (e->d_pires) { CU(pht_dev_free(e->d_pires, e->stream)); e->d_pires = nullptr; }
    CU(pht_dev_alloc((void **)&e->d_pires, sizeof(double) * (size_t)rows * e->cfg.n, e->stream));
    CU(cudaStreamSynchronize(e->stream));
    e->res_rows = rows;
    return 0;
}

extern "C" int pht_engine_enqueue(pht_engine *e, int nsweeps) {
    if (!e || nsweeps < 0) return fail("bad argument");
    CU(cudaSetDevice(e->cfg.device));
    if (ensure_res(e, nsweeps > 0 ? nsweeps : 1)) return -1;
    CU(cudaMemsetAsync(&e->d_state->res_row, 0, sizeof(uint32_t), e->stream));
    e->kev_used = 0;
    const bool graph = e->cfg.use_graph != 0;
    if (!graph) {
        const int want = 2 * (nsweeps < 64 ? nsweeps : 64);
        while ((int)e->kev.size() < want) { cudaEvent_t ev; CU(cudaEventCreate(&ev)); e->kev.push_back(ev); }
    }
    if (graph && (e->graph_exec == nullptr || e->graph_res != e->d_res || e->graph_res_rows != e->res_rows)) {
        if (e->graph_exec) { cudaGraphExecDestroy(e->graph_exec); e->graph_exec = nullptr; }
        cudaGraph_t g = nullptr;
        CU(cudaStreamBeginCapture(e->stream, cudaStreamCaptureModeThreadLocal));
        const unsigned long long before = e->launches;
        int rc = enqueue_sweep(e, e->d_res, e->res_rows, false);
        e->graph_launches = e->launches - before; e->launches = before;      /* counted per replay below */
        cudaError_t ce = cudaStreamEndCapture(e->stream, &g);
        if (rc != 0) { if (g) cudaGraphDestroy(g); return -1; }
        if (ce != cudaSuccess) return fail("graph capture failed: %s", cudaGetErrorString(ce));
        ce = cudaGraphInstantiate(&e->graph_exec, g, 0);
        cudaGraphDestroy(g);
        if (ce != cudaSuccess) return fail("graph instantiate failed: %s", cudaGetErrorString(ce));
        e->graph_res = e->d_res; e->graph_res_rows = e->res_rows;
    }
    CU(cudaEventRecord(e->ev0, e->stream));
    for (int k = 0; k < nsweeps; k++) {
        if (e->d_flush) CU(cudaMemsetAsync(e->d_flush, k & 0xff, e->flush_bytes, e->stream));
        if (graph) { CU(cudaGraphLaunch(e->graph_exec, e->stream)); e->launches += e->graph_launches; }
        else if (enqueue_sweep(e, e->d_res, e->res_rows, k < 64)) return -1;
    }
    CU(cudaEventRecord(e->ev1, e->stream));
    return 0;
}

static int check_state(pht_engine *e) {
    DevState st; CU(cudaMemcpy(&st, e->d_state, sizeof(st), cudaMemcpyDeviceToHost));
    if (st.error) {
        int err = st.error;
        cudaMemsetAsync(&e->d_state->error, 0, sizeof(int), e->stream);
        cudaStreamSynchronize(e->stream);
        return fail("device error word 0x%x (%s%s%s%s%s%s%s)", err, (err & 2) ? "sojourn total overflows the fixed-point range (lower PHT_B200_ZBITS); " : "",
                    (err & 4) ? "MHRS tail list overflow; " : "",
                    (err & 8) ? "an observation's survival probability is too small for rejection sampling (more than 4e9 attempts); " : "",
                    (err & 16) ? "(unused) " : "",
                    (err & 32) ? "spectral decomposition failed; " : "",
                    (err & 64) ? "a peer GPU did not arrive at a barrier of the global MHRS tail; " : "",
                    (err & 128) ? "another rank of the run raised its error word; " : "");
    }
    return 0;
}

extern "C" int pht_engine_sync(pht_engine *e) {
    if (!e) return fail("null engine");
    CU(cudaSetDevice(e->cfg.device));
    CU(cudaStreamSynchronize(e->stream));
    return check_state(e);
}

extern "C" int pht_engine_set_l2_flush(pht_engine *e, unsigned long long bytes) {
    if (!e) return fail("null engine");
    CU(cudaSetDevice(e->cfg.device));
    CU(cudaStreamSynchronize(e->stream));
    if (e->d_flush) { CU(cudaFree(e->d_flush)); e->d_flush = nullptr; e->flush_bytes = 0; }
    if (bytes) { CU(cudaMalloc(&e->d_flush, (size_t)bytes)); e->flush_bytes = (size_t)bytes; }
    return 0;
}

extern "C" int pht_engine_last_ms(pht_engine *e, float *total_ms, float *path_kernel_ms) {
    if (!e) return fail("null engine");
    if (total_ms) CU(cudaEventElapsedTime(total_ms, e->ev0, e->ev1));
    if (path_kernel_ms) {
        float acc = 0.f; int pairs = e->kev_used / 2;
        for (int i = 0; i < pairs; i++) { float ms; CU(cudaEventElapsedTime(&ms, e->kev[2 * i], e->kev[2 * i + 1])); acc += ms; }
        *path_kernel_ms = pairs ? acc / pairs : -1.f;
    }
    return 0;
}

extern "C" int pht_engine_run(pht_engine *e, int nsweeps, double *out) {
    if (pht_engine_enqueue(e, nsweeps)) return -1;
    if (pht_engine_sync(e)) return -1;
    if (out && nsweeps > 0) CU(cudaMemcpy(out, e->d_res, sizeof(double) * (size_t)nsweeps * e->cfg.m, cudaMemcpyDeviceToHost));
    return 0;
}

extern "C" int pht_engine_sweep_stats(pht_engine *e, long long *N, long long *B, long long *zfix) {
    if (!e) return fail("null engine");
    CU(cudaSetDevice(e->cfg.device));
    const int n = e->cfg.n;
    UpdateParams u = update_params(e, nullptr, 0);
    if (enqueue_model(e, u)) return -1;
    SweepParams p = sweep_params(e);
    if (enqueue_paths(e, p)) return -1;
    if (pht_engine_sync(e)) return -1;
    std::vector<long long> h(stats_len(n));
    CU(cudaMemcpy(h.data(), e->d_stats, sizeof(long long) * h.size(), cudaMemcpyDeviceToHost));
    if (N) memcpy(N, h.data(), sizeof(long long) * n * n);
    if (B) memcpy(B, h.data() + n * n, sizeof(long long) * n);
    if (zfix) for (int i = 0; i < n; i++) {
        const __int128 tot = ((__int128)h[n * n + 2 * n + i] << 32) + (__int128)(unsigned long long)h[n * n + n + i];
        if (tot > (__int128)0x7fffffffffffffffLL || tot < -(__int128)0x7fffffffffffffffLL) return fail("sojourn total of state %d overflows the fixed-point range (zbits = %d)", i, e->cfg.zbits);
        zfix[i] = (long long)tot;
    }
    return 0;
}

extern "C" int pht_engine_paths(pht_engine *e, long first, long count, int *B, int *N, double *z) {
    if (!e || !B || !N || !z) return fail("null argument");
    if (first < 0 || count < 0 || first + count > e->l_local) return fail("range [%ld, %ld) outside the shard", first, first + count);
    if (count == 0) return 0;
    CU(cudaSetDevice(e->cfg.device));
    const int n = e->cfg.n;
    int *dB = nullptr, *dN = nullptr; double *dz = nullptr;
    uint32_t *t_exact = nullptr, *t_cens = nullptr;
    /* every exit goes through `out`, which releases the temporaries */
    int rc = -1;
#define CUP(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { fail("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); goto out; } } while (0)
    {
        CUP(cudaMalloc(&dB, sizeof(int) * count)); CUP(cudaMalloc(&dN, sizeof(int) * count * n * n)); CUP(cudaMalloc(&dz, sizeof(double) * count * n));
        CUP(cudaMemsetAsync(dB, 0, sizeof(int) * count, e->stream)); CUP(cudaMemsetAsync(dN, 0, sizeof(int) * count * n * n, e->stream));
        CUP(cudaMemsetAsync(dz, 0, sizeof(double) * count * n, e->stream));
        UpdateParams u = update_params(e, nullptr, 0);
        if (enqueue_model(e, u)) goto out;
        SweepParams p = sweep_params(e);
        p.y = e->d_y; p.cens = e->d_cens; p.perm = nullptr;         /* the window [first, first + count) is in upload order */
        p.outB = dB; p.outN = dN; p.outz = dz; p.first = first; p.count = count;
        if (method_of(e->cfg) == PHT_METHOD_ECS) {
            std::vector<uint32_t> ie, ic;
            for (long i = first; i < first + count; i++) (e->h_cens[i] ? ic : ie).push_back((uint32_t)i);
            CUP(cudaMalloc(&t_exact, sizeof(uint32_t) * (ie.size() + 1))); CUP(cudaMalloc(&t_cens, sizeof(uint32_t) * (ic.size() + 1)));
            if (!ie.empty()) CUP(cudaMemcpy(t_exact, ie.data(), sizeof(uint32_t) * ie.size(), cudaMemcpyHostToDevice));
            if (!ic.empty()) CUP(cudaMemcpy(t_cens, ic.data(), sizeof(uint32_t) * ic.size(), cudaMemcpyHostToDevice));
            if (enqueue_paths(e, p, t_exact, ie.size(), t_cens, ic.size(), true)) goto out;
        } else if (method_of(e->cfg) == PHT_METHOD_MHS_ASLETT) {
            std::vector<uint32_t> all((size_t)count);
            for (long i = 0; i < count; i++) all[(size_t)i] = (uint32_t)(first + i);
            CUP(cudaMalloc(&t_exact, sizeof(uint32_t) * all.size()));
            CUP(cudaMemcpy(t_exact, all.data(), sizeof(uint32_t) * all.size(), cudaMemcpyHostToDevice));
            if (enqueue_paths(e, p, t_exact, all.size(), nullptr, 0, true)) goto out;
        } else if (enqueue_paths(e, p)) goto out;
        if (pht_engine_sync(e)) goto out;
        CUP(cudaMemcpy(B, dB, sizeof(int) * count, cudaMemcpyDeviceToHost));
        CUP(cudaMemcpy(N, dN, sizeof(int) * count * n * n, cudaMemcpyDeviceToHost));
        CUP(cudaMemcpy(z, dz, sizeof(double) * count * n, cudaMemcpyDeviceToHost));
        rc = 0;
    }
#undef CUP
out:
    cudaFree(dB); cudaFree(dN); cudaFree(dz);
    if (t_exact) cudaFree(t_exact);
    if (t_cens) cudaFree(t_cens);
    return rc;
}

extern "C" int pht_engine_set_spectral(pht_engine *e, const double *evals, const double *Q, const double *Qinv) {
    if (!e) return fail("null engine");
    CU(cudaSetDevice(e->cfg.device));
    CU(cudaStreamSynchronize(e->stream));
    const int n = e->cfg.n;
    if (e->graph_exec) { cudaGraphExecDestroy(e->graph_exec); e->graph_exec = nullptr; }     /* the sweep changes shape */
    if (!evals || !Q || !Qinv) {
        if (e->d_inject) { CU(cudaFree(e->d_inject)); e->d_inject = nullptr; }
        return 0;
    }
    if (!e->d_inject) CU(cudaMalloc(&e->d_inject, sizeof(double) * (n + 2 * n * n)));
    CU(cudaMemcpy(e->d_inject, evals, sizeof(double) * n, cudaMemcpyHostToDevice));
    CU(cudaMemcpy(e->d_inject + n, Q, sizeof(double) * n * n, cudaMemcpyHostToDevice));
    CU(cudaMemcpy(e->d_inject + n + n * n, Qinv, sizeof(double) * n * n, cudaMemcpyHostToDevice));
    return 0;
}

extern "C" int pht_engine_get_model(pht_engine *e, double *S, double *s, double *P, double *Pfull,
                                    double *evals, double *Q, double *Qinv) {
    if (!e) return fail("null engine");
    CU(cudaSetDevice(e->cfg.device));
    CU(cudaStreamSynchronize(e->stream));
    const int n = e->cfg.n;
    std::vector<double> h(e->L.total);
    CU(cudaMemcpy(h.data(), e->d_model, sizeof(double) * h.size(), cudaMemcpyDeviceToHost));
    if (S) memcpy(S, &h[e->L.S], sizeof(double) * n * n);
    if (s) memcpy(s, &h[e->L.s], sizeof(double) * n);
    if (P) memcpy(P, &h[e->L.P], sizeof(double) * n * n);
    if (Pfull) memcpy(Pfull, &h[e->L.Pfull], sizeof(double) * n * (n + 1));
    if (evals) memcpy(evals, &h[e->L.evals], sizeof(double) * n);
    if (Q) memcpy(Q, &h[e->L.Q], sizeof(double) * n * n);
    if (Qinv) memcpy(Qinv, &h[e->L.Qinv], sizeof(double) * n * n);
    return 0;
}

extern "C" int pht_engine_round_trace(pht_engine *e, unsigned long long *out) {
    if (!e || !out) return fail("null argument");
    CU(cudaSetDevice(e->cfg.device));
    CU(cudaStreamSynchronize(e->stream));
    DevState *st = new DevState; cudaError_t ce = cudaMemcpy(st, e->d_state, sizeof(*st), cudaMemcpyDeviceToHost);
    if (ce == cudaSuccess) memcpy(out, st->round_trace, sizeof(st->round_trace));
    delete st;
    CU(ce);
    return 0;
}

extern "C" int pht_engine_counters(pht_engine *e, unsigned long long *out) {
    if (!e || !out) return fail("null argument");
    CU(cudaSetDevice(e->cfg.device));
    CU(cudaStreamSynchronize(e->stream));
    DevState st; CU(cudaMemcpy(&st, e->d_state, sizeof(st), cudaMemcpyDeviceToHost));
    memcpy(out, st.counters, sizeof(st.counters));
    out[PHT_CNT_LAUNCHES] = e->launches;
    return 0;
}
