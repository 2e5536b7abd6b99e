// Engine: the iteration driver and the C ABI of libsqsm_b200 (include/sqsm_b200.h).  sm_100a.
//
// One handle = one SpconvBasedContraction object (core/core_algorithms/spconv_based_contraction.py:34-151):
// state (contracted points, displacements, epoch, Chamfer value) lives on the device between
// the thunks of get_pipeline(); nothing crosses PCIe per batch (the reference copies every batch
// back with .cpu().numpy(), :107-108).
#include "engine.h"
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <limits>

using namespace sqsm;

struct DebugCapture {
    bool valid = false;
    int n_levels = 0;
    int n_batches = 0;
    int batch_size = 16;
    LevelTables lv[kLevels];
    Level0Rows rows;
    ForwardBuffers fb;
};

struct sqsm_engine {
    SqsmConfig cfg;
    Ctx cx;
    std::string err;
    std::map<std::string, std::vector<float>> state;
    NetWeights nw;
    DBuf<float> P, D;
    int64_t n = 0;
    DBuf<float> candP, candD;       // candidate state of the running thunk (this rank's segment, or the assembled cloud)
    int64_t cand_n = 0;
    bool cand_valid = false;
    int64_t step_launches0 = 0;
    int epoch = 0;
    double chamfer_prev = std::numeric_limits<double>::infinity();
    bool early_stopped = false;
    bool capture = false;
    DebugCapture dbg;
    std::map<std::string, ConvProfile> prof;   // resolved kernel-family records
    cudaEvent_t tm_a = nullptr, tm_b = nullptr; // sqsm_timer_start / stop
    SkeletalResult sk;
    bool sk_valid = false;
    double sk_voxel = 0.0;          // voxel_size_for_downsampling the cached sample was made with
    DBuf<int64_t> edges;            // cached radius graph of `sk`
    int64_t n_edges = 0;
    double edges_r = -1.0;
    int64_t max_entries_per_pass = 48ll << 20;
    size_t pool_reserved = 0;
    PrepState prep;
    GraphState graph;
};

static std::string g_create_error;

// The double a config value meant: SqsmConfig carries voxel_size as float32, the reference passes the YAML's double to
// Open3D.  The shortest decimal that round-trips the float (<= 9 significant digits) recovers it exactly for every value
// a config file would hold (0.01, 0.005, 0.0125, ...); a point within ~1e-5 voxel of a cell face would otherwise fall
// into a different Chamfer voxel than in the reference.
static double decimal_double(float v) {
    char buf[64];
    for (int digits = 1; digits <= 9; ++digits) {
        snprintf(buf, sizeof(buf), "%.*g", digits, (double)v);
        if ((float)strtod(buf, nullptr) == v) return strtod(buf, nullptr);
    }
    return (double)v;
}

#define SQSM_API_BEGIN(h)                                                                         \
    if (!(h)) return 1;                                                                           \
    try {                                                                                         \
        SQSM_CUDA(cudaSetDevice((h)->cfg.device));
#define SQSM_API_END(h)                                                                           \
        return 0;                                                                                 \
    } catch (const std::exception& e) {                                                           \
        (h)->err = e.what();                                                                      \
        cudaGetLastError();                                                                       \
        return 2;                                                                                 \
    }

extern "C" int sqsm_abi_version(void) { return 1; }

extern "C" int sqsm_sizeof(int32_t which) {
    switch (which) {
        case 0: return (int)sizeof(SqsmConfig);
        case 1: return (int)sizeof(SqsmIterStats);
        case 2: return (int)sizeof(SqsmPrepConfig);
        case 3: return (int)sizeof(SqsmPrepStats);
        default: return -1;
    }
}

extern "C" const char* sqsm_last_error(sqsm_handle h) { return h ? h->err.c_str() : g_create_error.c_str(); }

extern "C" int sqsm_create(const SqsmConfig* cfg, sqsm_handle* out) {
    if (!cfg || !out) { g_create_error = "null argument"; return 1; }
    *out = nullptr;
    sqsm_engine* e = nullptr;
    try {
        int count = 0;
        cudaError_t ce = cudaGetDeviceCount(&count);
        if (ce != cudaSuccess || count == 0)
            throw CudaError("no CUDA device: libsqsm_b200 has no CPU fallback (needs an sm_100 GPU)");
        SQSM_CHECK(cfg->device >= 0 && cfg->device < count, "device ordinal out of range");
        cudaDeviceProp prop;
        SQSM_CUDA(cudaGetDeviceProperties(&prop, cfg->device));
        if (prop.major != 10) {
            char b[256];
            snprintf(b, sizeof(b), "device %d (%s) is sm_%d%d; libsqsm_b200 is built for sm_100a only", cfg->device,
                     prop.name, prop.major, prop.minor);
            throw CudaError(b);
        }
        SQSM_CHECK(cfg->batch_size > 0 && cfg->voxel_size > 0.f && cfg->cube_size > 0.f && cfg->buffer_size >= 0.f,
                   "bad configuration");
        SQSM_CUDA(cudaSetDevice(cfg->device));
        e = new sqsm_engine();
        e->cfg = *cfg;
        SQSM_CUDA(cudaStreamCreateWithFlags(&e->cx.st, cudaStreamNonBlocking));
        cudaMemPool_t pool;
        SQSM_CUDA(cudaDeviceGetDefaultMemPool(&pool, cfg->device));
        uint64_t thr = UINT64_MAX;                          // keep freed blocks cached in the pool
        SQSM_CUDA(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr));
        if (const char* env = getenv("SQSM_MAX_ENTRIES_PER_PASS")) e->max_entries_per_pass = atoll(env);
        *out = e;
        return 0;
    } catch (const std::exception& ex) {
        g_create_error = ex.what();
        delete e;
        return 2;
    }
}

extern "C" void sqsm_destroy(sqsm_handle h) {
    if (!h) return;
    cudaSetDevice(h->cfg.device);
    cudaStream_t st = h->cx.st;
    cudaStreamSynchronize(st);
    const int dev = h->cfg.device;
    delete h;                                               // DBufs free on the stream
    cudaStreamSynchronize(st);
    cudaStreamDestroy(st);
    // hand the cached blocks of the (process-wide) default pool back to the driver: the reference pipeline and torch
    // keep running in the same process after the engine is gone
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) cudaMemPoolTrimTo(pool, 0);
}

extern "C" int sqsm_set_weight(sqsm_handle h, const char* name, const float* data, int64_t n) {
    SQSM_API_BEGIN(h)
    SQSM_CHECK(name && data && n > 0, "bad weight tensor");
    h->state[name] = std::vector<float>(data, data + n);
    h->nw.ready = false;
    SQSM_API_END(h)
}

extern "C" int sqsm_finalize_weights(sqsm_handle h) {
    SQSM_API_BEGIN(h)
    net_upload(h->cx, h->nw, h->state, h->cfg.down_kernel_flip != 0);
    int mode = h->cfg.conv_mode;
    if (mode == 0) { const char* env = getenv("SQSM_CONV_MODE"); mode = env ? atoi(env) : 4; }
    SQSM_CHECK(mode >= 1 && mode <= 5, "conv_mode must be 0..5");
    h->nw.mode = mode - 1;                                   // 0 FFMA, 1 3xTF32, 2 TF32, 3 3xFP16, 4 3xFP16 + TMA gather
    if (const char* env = getenv("SQSM_TC_MIN_CO")) h->nw.tc_min_co = atoi(env);
    SQSM_API_END(h)
}

// Pre-grow the stream-ordered memory pool so that the temporaries of a step are carved out of memory the
// pool already owns (growing the pool mid-step costs milliseconds per allocation).
static void reserve_pool(sqsm_engine* h, int64_t n_points) {
    size_t want = (size_t)n_points * 2200 + (1ull << 30);        // ~1.5-2 kB of tables and features per point
    if (want <= h->pool_reserved) return;
    size_t free_b = 0, total_b = 0;
    if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess) return;
    want = std::min(want, (size_t)(free_b * 0.8));
    if (want <= h->pool_reserved) return;
    void* p = nullptr;
    if (cudaMallocAsync(&p, want, h->cx.st) == cudaSuccess) {
        cudaFreeAsync(p, h->cx.st);
        cudaStreamSynchronize(h->cx.st);
        h->pool_reserved = want;
    } else {
        cudaGetLastError();
    }
}

static void reset_state(sqsm_engine* h) {
    h->epoch = 0;
    h->chamfer_prev = std::numeric_limits<double>::infinity();
    h->early_stopped = false;
    h->sk_valid = false;
    h->edges_r = -1.0;
    h->dbg.valid = false;
}

__global__ void __launch_bounds__(256) k_f64_to_f32(const double* __restrict__ in, float* __restrict__ out, int64_t n) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = __double2float_rn(in[i]);           // np.copy(points).astype(np.float32) (:82)
}

extern "C" int sqsm_set_points_f64(sqsm_handle h, const double* xyz, int64_t n) {
    SQSM_API_BEGIN(h)
    SQSM_CHECK(xyz && n > 0 && n < (1ll << 31), "bad point array");
    cudaStream_t st = h->cx.st;
    DBuf<double> tmp(n * 3, st);
    SQSM_CUDA(cudaMemcpyAsync(tmp.p, xyz, sizeof(double) * n * 3, cudaMemcpyHostToDevice, st));
    h->P.alloc(n * 3, st);
    h->D.alloc(n * 3, st);
    h->D.zero();                                            // np.zeros_like(points) (:83)
    k_f64_to_f32<<<div_up(n * 3, 256), 256, 0, st>>>(tmp.p, h->P.p, n * 3);
    h->cx.launches++;
    h->n = n;
    reset_state(h);
    SQSM_CUDA(cudaStreamSynchronize(st));
    reserve_pool(h, n);
    SQSM_API_END(h)
}

extern "C" int sqsm_set_points_dev_f32(sqsm_handle h, const float* d_xyz, int64_t n) {
    SQSM_API_BEGIN(h)
    SQSM_CHECK(d_xyz && n > 0 && n < (1ll << 31), "bad point array");
    cudaStream_t st = h->cx.st;
    h->P.alloc(n * 3, st);
    h->D.alloc(n * 3, st);
    h->D.zero();
    SQSM_CUDA(cudaMemcpyAsync(h->P.p, d_xyz, sizeof(float) * n * 3, cudaMemcpyDeviceToDevice, st));
    h->n = n;
    reset_state(h);
    SQSM_CUDA(cudaStreamSynchronize(st));
    reserve_pool(h, n);
    SQSM_API_END(h)
}

// ---- cloud preparation (row f2)
extern "C" int sqsm_prepare_cloud_f64(sqsm_handle h, const double* xyz, int64_t n, const SqsmPrepConfig* cfg, SqsmPrepStats* stats) {
    SQSM_API_BEGIN(h)
    SQSM_CHECK(cfg != nullptr, "null SqsmPrepConfig");
    h->prep.valid = false;
    SqsmPrepStats st_local;
    prepare_cloud(h->cx, xyz, n, *cfg, h->prep, st_local);
    if (stats) *stats = st_local;
    SQSM_API_END(h)
}

extern "C" int sqsm_prepared_size(sqsm_handle h, int64_t* n_kept, int64_t* n_voxel) {
    SQSM_API_BEGIN(h)
    SQSM_CHECK(h->prep.valid, "no prepared cloud: call sqsm_prepare_cloud_f64 first");
    if (n_kept) *n_kept = h->prep.n_kept;
    if (n_voxel) *n_voxel = h->prep.n_voxel;
    SQSM_API_END(h)
}

extern "C" int sqsm_get_prepared(sqsm_handle h, double* xyz) {
    SQSM_API_BEGIN(h)
    SQSM_CHECK(h->prep.valid && xyz, "no prepared cloud / null output");
    SQSM_CUDA(cudaMemcpyAsync(xyz, h->prep.pts.p, sizeof(double) * h->prep.n_kept * 3, cudaMemcpyDeviceToHost, h->cx.st));
    SQSM_CUDA(cudaStreamSynchronize(h->cx.st));
    SQSM_API_END(h)
}

extern "C" int sqsm_set_points_prepared(sqsm_handle h) {
    SQSM_API_BEGIN(h)
    SQSM_CHECK(h->prep.valid, "no prepared cloud: call sqsm_prepare_cloud_f64 first");
    cudaStream_t st = h->cx.st;
    const int64_t n = h->prep.n_kept;
    h->P.alloc(n * 3, st);
    h->D.alloc(n * 3, st);
    h->D.zero();
    prepared_to_f32(h->cx, h->prep, h->P.p);
    h->n = n;
    reset_state(h);
    SQSM_CUDA(cudaStreamSynchronize(st));
    reserve_pool(h, n);
    SQSM_API_END(h)
}

extern "C" int sqsm_debug_prepared(sqsm_handle h, double* vox, double* avg, uint8_t* keep) {
    SQSM_API_BEGIN(h)
    SQSM_CHECK(h->prep.valid, "no prepared cloud");
    cudaStream_t st = h->cx.st;
    const int64_t nv = h->prep.n_voxel;
    if (vox) {
        SQSM_CHECK(h->prep.vox.n == nv * 3, "voxel means were not kept (keep_debug = 0)");
        SQSM_CUDA(cudaMemcpyAsync(vox, h->prep.vox.p, sizeof(double) * nv * 3, cudaMemcpyDeviceToHost, st));
    }
    if (avg) {
        SQSM_CHECK(h->prep.avg.n == nv, "avg distances were not kept (keep_debug = 0 or filter = 0)");
        SQSM_CUDA(cudaMemcpyAsync(avg, h->prep.avg.p, sizeof(double) * nv, cudaMemcpyDeviceToHost, st));
    }
    if (keep) {
        SQSM_CHECK(h->prep.keep.n == nv, "keep flags were not kept (keep_debug = 0 or filter = 0)");
        SQSM_CUDA(cudaMemcpyAsync(keep, h->prep.keep.p, nv, cudaMemcpyDeviceToHost, st));
    }
    SQSM_CUDA(cudaStreamSynchronize(st));
    SQSM_API_END(h)
}

extern "C" int sqsm_set_state_f32(sqsm_handle h, const float* xyz, const float* disp, int64_t n) {
    SQSM_API_BEGIN(h)
    SQSM_CHECK(xyz && disp && n > 0 && n < (1ll << 31), "bad state arrays");
    cudaStream_t st = h->cx.st;
    h->P.alloc(n * 3, st);
    h->D.alloc(n * 3, st);
    SQSM_CUDA(cudaMemcpyAsync(h->P.p, xyz, sizeof(float) * n * 3, cudaMemcpyHostToDevice, st));
    SQSM_CUDA(cudaMemcpyAsync(h->D.p, disp, sizeof(float) * n * 3, cudaMemcpyHostToDevice, st));
    h->n = n;
    reset_state(h);
    SQSM_CUDA(cudaStreamSynchronize(st));
    SQSM_API_END(h)
}

struct StageTimer {
    cudaEvent_t ev[8];
    int n = 0;
    cudaStream_t st;
    explicit StageTimer(cudaStream_t s) : st(s) { for (auto& e : ev) cudaEventCreate(&e); }
    ~StageTimer() { for (auto& e : ev) cudaEventDestroy(e); }
    void mark() { if (n < 8) cudaEventRecord(ev[n++], st); }
    float between(int a, int b) { float ms = 0; cudaEventElapsedTime(&ms, ev[a], ev[b]); return ms; }
};

// ---- one _contract() thunk in three phases (so that a single tree can be sharded over several ranks) --------------
//
//   begin   : tiling of the whole cloud, then voxelisation -> rulebooks -> U-Net -> decode for the cube batches of this
//             shard (contiguous batch range, balanced by entry count); result = a SEGMENT of the new cloud
//   (ranks exchange their segments; the concatenation in rank order is the new cloud in canonical order)
//   chamfer : partial sums of the monitor over this shard's query bricks (all-reduced by the caller)
//   finish  : accept (state <- candidate) or reject
// sqsm_contract_step() is begin(0,1) + chamfer(0,1) + finish on one rank.

static void shard_batch_range(const TilingResult& tl, int batch_size, int shard, int n_shards, int* s_lo, int* s_hi) {
    // contiguous batch ranges with about E / n_shards entries each; identical on every rank
    const int nb = tl.n_batches;
    const int64_t E = tl.n_entries;
    auto batch_start_entry = [&](int b) { return tl.start[std::min(b * batch_size, tl.n_samples)]; };
    auto first_batch_at = [&](int64_t target) {
        int lo = 0, hi = nb;
        while (lo < hi) { int mid = (lo + hi) / 2; if (batch_start_entry(mid) < target) lo = mid + 1; else hi = mid; }
        return lo;
    };
    int b0 = (shard == 0) ? 0 : first_batch_at(E * shard / n_shards);
    int b1 = (shard == n_shards - 1) ? nb : first_batch_at(E * (shard + 1) / n_shards);
    *s_lo = std::min(b0 * batch_size, tl.n_samples);
    *s_hi = std::min(b1 * batch_size, tl.n_samples);
}

static void step_begin(sqsm_engine* h, int shard, int n_shards, SqsmIterStats& s) {
    SQSM_CHECK(h->n > 0, "no points set");
    SQSM_CHECK(h->nw.ready, "weights not finalised");
    SQSM_CHECK(n_shards >= 1 && shard >= 0 && shard < n_shards, "bad shard index");
    Ctx& cx = h->cx;
    cudaStream_t st = cx.st;
    const SqsmConfig& cfg = h->cfg;
    std::memset(&s, 0, sizeof(s));
    s.chamfer = std::numeric_limits<double>::quiet_NaN();
    h->epoch += 1;                                                          // :76
    s.epoch = h->epoch;
    s.n_in = h->n;
    h->cand_valid = false;
    h->step_launches0 = cx.launches;
    if (h->early_stopped) {                                                 // :77-80
        s.skipped = 1;
        return;
    }
    h->sk_valid = false;
    h->edges_r = -1.0;
    h->dbg.valid = false;
    float ms_vox = 0, ms_rule = 0, ms_net = 0;
    StageTimer T0(st);
    T0.mark();
    TilingResult tl;
    {
        Trace tr(cx, "tile_points");
        ProfScope ps(cx, "tiling", (double)h->n * 12.0 * 3 + 0.0, 0.0);
        tile_points(cx, h->P.p, h->n, cfg.cube_size, cfg.buffer_size, cfg.voxel_size, cfg.batch_size, tl);   // :85-91
    }
    T0.mark();
    s.n_entries = tl.n_entries;
    s.n_cubes = tl.n_samples;
    s.n_batches = tl.n_batches;
    int S0 = 0, S1 = tl.n_samples;
    if (n_shards > 1) shard_batch_range(tl, cfg.batch_size, shard, n_shards, &S0, &S1);
    const int64_t seg_entries = tl.start[S1] - tl.start[S0];
    h->candP.alloc(std::max<int64_t>(seg_entries, 1) * 3, st);
    h->candD.alloc(std::max<int64_t>(seg_entries, 1) * 3, st);
    int64_t out_base = 0;
    for (int s0 = S0; s0 < S1;) {
        // a pass = as many whole batches (cfg.batch_size cubes) as fit the entry budget
        int s1 = s0;
        while (s1 < S1) {
            int nxt = std::min(s1 + cfg.batch_size, S1);
            if (s1 > s0 && tl.start[nxt] - tl.start[s0] > h->max_entries_per_pass) break;
            s1 = nxt;
        }
        const int batch0 = s0 / cfg.batch_size;
        const int nb = (s1 - 1) / cfg.batch_size - batch0 + 1;
        StageTimer T(st);
        T.mark();
        LevelTables lv[kLevels];
        Level0Rows rows;
        DBuf<int32_t> ent_row;
        int64_t n_out = 0;
        {
            // B_vox of SURVEY 8(d): per entry 24 B in (point + displacement), 8 B key, 16 B coordinates, 12 B features
            Trace tr(cx, "build_level0");
            ProfScope ps(cx, "voxelise", (double)(tl.start[s1] - tl.start[s0]) * (24.0 + 8.0 + 16.0 + 12.0), 0.0);
            build_level0(cx, h->P.p, h->D.p, tl, s0, s1, cfg.batch_size, cfg.voxel_size, lv[0], rows, out_base, &n_out, ent_row);
        }
        ent_row.release();
        ProfScope* rs = new ProfScope(cx, "rulebook", 0.0, 0.0);
        T.mark();
        int64_t far = 0;
        {
            // level 0: the compact rulebook when every 27-offset layer of the level runs on the sparse-loop kernel; the dense
            // table in addition only for the debug capture (the parity tests compare it with the oracle's)
            Trace tr(cx, "build_subm L0");
            const bool csr = net_level0_uses_csr(h->nw);
            build_subm(cx, lv[0], batch0, nb, cfg.batch_size, &far, !csr || h->capture, csr);
        }
        s.far_face_rows += far;
        int n_levels = 1;
        for (int l = 0; l + 1 < kLevels; ++l) {
            { Trace tr(cx, "build_down"); if (!build_down(cx, lv[l], lv[l + 1], batch0, nb, cfg.batch_size)) break; }
            { Trace tr(cx, "build_subm"); build_subm(cx, lv[l + 1], batch0, nb, cfg.batch_size, nullptr); }
            n_levels = l + 2;
        }
        {   // algorithmic bytes of the rulebook stage: 27 int32 table entries written per row + brick lines read
            double by = 0;
            for (int l = 0; l < n_levels; ++l) by += (double)lv[l].n * (27.0 * 4 + 6.0) + (double)lv[l].map.n_bricks * 128.0 * 2;
            rs->r.bytes = by;
        }
        delete rs;
        T.mark();
        s.n_vox += lv[0].n;
        s.pairs_l0 += lv[0].pairs;
        if (n_levels > 1) { s.n_rows_l1 += lv[1].n; s.pairs_l1 += lv[1].pairs; }
        if (n_levels > 2) { s.n_rows_l2 += lv[2].n; s.pairs_l2 += lv[2].pairs; }
        if (n_levels > 3) { s.n_rows_l3 += lv[3].n; s.pairs_l3 += lv[3].pairs; }
        ForwardBuffers fb;
        { Trace tr(cx, "net_forward"); net_forward(cx, h->nw, lv, rows, n_levels, fb, h->capture, h->candP.p, h->candD.p); }  // :100-108
        T.mark();
        SQSM_CUDA(cudaStreamSynchronize(st));
        ms_vox += T.between(0, 1);
        ms_rule += T.between(1, 2);
        ms_net += T.between(2, 3);
        out_base += n_out;
        if (h->capture) {
            SQSM_CHECK(n_shards == 1 && s0 == 0 && s1 == tl.n_samples,
                       "debug capture needs the whole iteration in one pass (raise SQSM_MAX_ENTRIES_PER_PASS)");
            DebugCapture& d = h->dbg;
            d.n_levels = n_levels;
            d.n_batches = nb;
            d.batch_size = cfg.batch_size;
            for (int l = 0; l < kLevels; ++l) d.lv[l] = std::move(lv[l]);
            d.rows = std::move(rows);
            d.fb = std::move(fb);
            d.valid = true;
        }
        s0 = s1;
    }
    if (h->nw.mode == 3 || h->nw.mode == 4) {
        int flag = 0;
        SQSM_CUDA(cudaMemcpyAsync(&flag, h->nw.range_flag, sizeof(int), cudaMemcpyDeviceToHost, st));
        SQSM_CUDA(cudaStreamSynchronize(st));
        if (flag != 0) {
            SQSM_CUDA(cudaMemsetAsync(h->nw.range_flag, 0, sizeof(int), st));    // the next step starts clean
            h->cand_valid = false;
            SQSM_CHECK(false, "an activation exceeded the FP16 range in conv_mode 4/5 (3xFP16); use conv_mode 2 (3xTF32)");
        }
    }
    s.n_out = out_base;
    if (Trace::enabled()) fprintf(stderr, "[trace] ---- end of passes, n_out %lld\n", (long long)out_base);
    SQSM_CHECK(n_shards > 1 || out_base > 0, "contraction produced no interior voxel");
    h->cand_n = out_base;
    h->cand_valid = true;
    SQSM_CUDA(cudaStreamSynchronize(st));
    s.ms_tiling = T0.between(0, 1); s.ms_voxel = ms_vox; s.ms_rulebook = ms_rule; s.ms_network = ms_net;
    s.gpu_launches = (int32_t)(cx.launches - h->step_launches0);
}

static void step_chamfer(sqsm_engine* h, int shard, int n_shards, double sums[2], int64_t counts[2], float* ms) {
    SQSM_CHECK(h->cand_valid && h->cand_n > 0, "no candidate state (call sqsm_step_begin / sqsm_step_set_candidate first)");
    Ctx& cx = h->cx;
    StageTimer T(cx.st);
    T.mark();
    {
        // B_cd of SURVEY 8(d): both clouds read (12 B/point) + voxel means written and read (2 x 24 B/voxel ~ per point)
        Trace tr(cx, "chamfer");
        ProfScope ps(cx, "chamfer", (double)(h->n + h->cand_n) * (12.0 + 2 * 20.0), 0.0);
        // the reference hands Open3D the Python float (a double) of the config, e.g. 0.01, not float32(0.01) = 0.00999999977...
        chamfer_partial(cx, h->P.p, h->n, h->candP.p, h->cand_n, decimal_double(h->cfg.voxel_size), shard, n_shards, sums, counts);
    }
    T.mark();
    SQSM_CUDA(cudaStreamSynchronize(cx.st));
    if (ms) *ms = T.between(0, 1);
}

// accept/reject exactly like spconv_based_contraction.py:123-131
static bool step_finish(sqsm_engine* h, bool monitored, double cd) {
    SQSM_CHECK(h->cand_valid, "no candidate state");
    bool accept = true;
    if (monitored) {
        if (cd > h->chamfer_prev) {
            accept = false;
            if (h->cfg.latch_early_stop) h->early_stopped = true;
        } else {
            h->chamfer_prev = cd;
        }
    }
    if (accept) {                                                            // :130-131
        std::swap(h->P, h->candP);
        std::swap(h->D, h->candD);
        h->n = h->cand_n;
    }
    h->cand_valid = false;
    SQSM_CUDA(cudaStreamSynchronize(h->cx.st));
    return accept;
}

extern "C" int sqsm_contract_step(sqsm_handle h, SqsmIterStats* stats) {
    SQSM_API_BEGIN(h)
    SqsmIterStats s;
    step_begin(h, 0, 1, s);
    if (!s.skipped) {
        double cd = std::numeric_limits<double>::quiet_NaN();
        if (h->cfg.monitored_by_chamfer_distance) {                          // :112-128
            double sums[2];
            int64_t counts[2];
            step_chamfer(h, 0, 1, sums, counts, &s.ms_chamfer);
            cd = sums[0] / (double)counts[0] + sums[1] / (double)counts[1];  // :122
            s.chamfer = cd;
        }
        s.accepted = step_finish(h, h->cfg.monitored_by_chamfer_distance != 0, cd) ? 1 : 0;
        s.gpu_launches = (int32_t)(h->cx.launches - h->step_launches0);
    }
    if (stats) *stats = s;
    SQSM_API_END(h)
}

extern "C" int sqsm_step_begin(sqsm_handle h, int32_t shard, int32_t n_shards, SqsmIterStats* stats) {
    SQSM_API_BEGIN(h)
    SqsmIterStats s;
    step_begin(h, shard, n_shards, s);
    if (stats) *stats = s;
    SQSM_API_END(h)
}

extern "C" int sqsm_step_segment(sqsm_handle h, int64_t* n, float* d_dst_xyz, float* d_dst_disp) {
    SQSM_API_BEGIN(h)
    SQSM_CHECK(h->cand_valid && n, "no segment (call sqsm_step_begin first)");
    *n = h->cand_n;
    cudaStream_t st = h->cx.st;
    if (d_dst_xyz && h->cand_n > 0)
        SQSM_CUDA(cudaMemcpyAsync(d_dst_xyz, h->candP.p, sizeof(float) * 3 * h->cand_n, cudaMemcpyDeviceToDevice, st));
    if (d_dst_disp && h->cand_n > 0)
        SQSM_CUDA(cudaMemcpyAsync(d_dst_disp, h->candD.p, sizeof(float) * 3 * h->cand_n, cudaMemcpyDeviceToDevice, st));
    SQSM_CUDA(cudaStreamSynchronize(st));
    SQSM_API_END(h)
}

extern "C" int sqsm_step_set_candidate(sqsm_handle h, const float* d_xyz, const float* d_disp, int64_t n) {
    SQSM_API_BEGIN(h)
    SQSM_CHECK(d_xyz && d_disp && n > 0 && n < (1ll << 31), "bad candidate arrays");
    cudaStream_t st = h->cx.st;
    h->candP.alloc(n * 3, st);
    h->candD.alloc(n * 3, st);
    SQSM_CUDA(cudaMemcpyAsync(h->candP.p, d_xyz, sizeof(float) * 3 * n, cudaMemcpyDeviceToDevice, st));
    SQSM_CUDA(cudaMemcpyAsync(h->candD.p, d_disp, sizeof(float) * 3 * n, cudaMemcpyDeviceToDevice, st));
    h->cand_n = n;
    h->cand_valid = true;
    SQSM_CUDA(cudaStreamSynchronize(st));
    SQSM_API_END(h)
}

extern "C" int sqsm_step_chamfer(sqsm_handle h, int32_t shard, int32_t n_shards, double* sums2, int64_t* counts2) {
    SQSM_API_BEGIN(h)
    SQSM_CHECK(sums2 && counts2, "null argument");
    step_chamfer(h, shard, n_shards, sums2, counts2, nullptr);
    SQSM_API_END(h)
}

extern "C" int sqsm_step_finish(sqsm_handle h, int32_t monitored, double chamfer, int32_t* accepted) {
    SQSM_API_BEGIN(h)
    bool acc = step_finish(h, monitored != 0, chamfer);
    if (accepted) *accepted = acc ? 1 : 0;
    SQSM_API_END(h)
}

extern "C" int sqsm_result_size(sqsm_handle h, int64_t* n) {
    SQSM_API_BEGIN(h)
    SQSM_CHECK(n, "null argument");
    *n = h->n;
    SQSM_API_END(h)
}

extern "C" int sqsm_get_result(sqsm_handle h, float* pts, float* disp) {
    SQSM_API_BEGIN(h)
    SQSM_CHECK(h->n > 0, "no state");
    cudaStream_t st = h->cx.st;
    if (pts) SQSM_CUDA(cudaMemcpyAsync(pts, h->P.p, sizeof(float) * h->n * 3, cudaMemcpyDeviceToHost, st));
    if (disp) SQSM_CUDA(cudaMemcpyAsync(disp, h->D.p, sizeof(float) * h->n * 3, cudaMemcpyDeviceToHost, st));
    SQSM_CUDA(cudaStreamSynchronize(st));
    SQSM_API_END(h)
}

extern "C" int sqsm_skeletal_sample(sqsm_handle h, double voxel, int64_t* m, int64_t* ids, float* pts, float* radii) {
    SQSM_API_BEGIN(h)
    SQSM_CHECK(h->n > 0 && m, "no state");
    cudaStream_t st = h->cx.st;
    SQSM_CHECK(voxel > 0.0, "voxel_size_for_downsampling must be positive");
    if (!h->sk_valid || h->sk_voxel != voxel) {              // cached per (state, voxel): the binding asks twice (size, then data)
        ProfScope ps(h->cx, "skeletal", (double)h->n * (24.0 + 4.0), 0.0);
        h->sk_valid = false;
        h->edges_r = -1.0;
        skeletal_sample(h->cx, h->P.p, h->D.p, h->n, voxel, h->sk);
        h->sk_valid = true;
        h->sk_voxel = voxel;
    }
    *m = h->sk.m;
    if (ids) SQSM_CUDA(cudaMemcpyAsync(ids, h->sk.ids.p, sizeof(int64_t) * h->sk.m, cudaMemcpyDeviceToHost, st));
    if (pts) SQSM_CUDA(cudaMemcpyAsync(pts, h->sk.pts.p, sizeof(float) * h->sk.m * 3, cudaMemcpyDeviceToHost, st));
    if (radii) SQSM_CUDA(cudaMemcpyAsync(radii, h->sk.radii.p, sizeof(float) * h->sk.m, cudaMemcpyDeviceToHost, st));
    SQSM_CUDA(cudaStreamSynchronize(st));
    SQSM_API_END(h)
}

static void edges_out(sqsm_engine* h, const float* d_pts, int64_t m, double r, int64_t* n_edges, int64_t* out,
                      int64_t capacity) {
    DBuf<int64_t> edges;
    int64_t E = 0;
    radius_graph(h->cx, d_pts, m, r, edges, &E);
    *n_edges = E;
    if (out && E > 0) {
        SQSM_CHECK(capacity < 0 || E <= capacity, "edge buffer too small");
        SQSM_CUDA(cudaMemcpyAsync(out, edges.p, sizeof(int64_t) * 2 * E, cudaMemcpyDeviceToHost, h->cx.st));
    }
    SQSM_CUDA(cudaStreamSynchronize(h->cx.st));
}

extern "C" int sqsm_radius_graph(sqsm_handle h, double r, int64_t* n_edges, int64_t* out) {
    SQSM_API_BEGIN(h)
    SQSM_CHECK(h->sk_valid && n_edges, "call sqsm_skeletal_sample first");
    if (h->edges_r != r) {                                   // computed once per skeletal sample and radius
        ProfScope ps(h->cx, "radius_graph", (double)h->sk.m * 12.0, 0.0);
        radius_graph(h->cx, h->sk.pts.p, h->sk.m, r, h->edges, &h->n_edges);
        h->edges_r = r;
    }
    *n_edges = h->n_edges;
    if (out && h->n_edges > 0)
        SQSM_CUDA(cudaMemcpyAsync(out, h->edges.p, sizeof(int64_t) * 2 * h->n_edges, cudaMemcpyDeviceToHost, h->cx.st));
    SQSM_CUDA(cudaStreamSynchronize(h->cx.st));
    SQSM_API_END(h)
}

extern "C" int sqsm_radius_graph_range(sqsm_handle h, double r, int64_t node_lo, int64_t node_hi, int64_t* n_edges,
                                      int64_t* out, int64_t capacity) {
    SQSM_API_BEGIN(h)
    SQSM_CHECK(h->sk_valid && n_edges, "call sqsm_skeletal_sample first");
    SQSM_CHECK(node_lo >= 0 && node_lo <= node_hi && node_hi <= h->sk.m, "bad node range");
    DBuf<int64_t> edges;
    int64_t E = 0;
    {
        ProfScope ps(h->cx, "radius_graph", (double)h->sk.m * 12.0, 0.0);
        radius_graph(h->cx, h->sk.pts.p, h->sk.m, r, edges, &E, node_lo, node_hi);
    }
    *n_edges = E;
    if (out && E > 0) {
        SQSM_CHECK(capacity < 0 || E <= capacity, "edge buffer too small");
        SQSM_CUDA(cudaMemcpyAsync(out, edges.p, sizeof(int64_t) * 2 * E, cudaMemcpyDeviceToHost, h->cx.st));
    }
    SQSM_CUDA(cudaStreamSynchronize(h->cx.st));
    SQSM_API_END(h)
}

extern "C" int sqsm_radius_graph_points(sqsm_handle h, const float* pts, int64_t m, double r, int64_t* n_edges,
                                        int64_t* out, int64_t capacity) {
    SQSM_API_BEGIN(h)
    SQSM_CHECK(pts && m > 0 && n_edges, "bad arguments");
    DBuf<float> d(m * 3, h->cx.st);
    SQSM_CUDA(cudaMemcpyAsync(d.p, pts, sizeof(float) * m * 3, cudaMemcpyHostToDevice, h->cx.st));
    edges_out(h, d.p, m, r, n_edges, out, capacity);
    SQSM_API_END(h)
}

// ---- refinement's ball occupancy queries (row f3)
extern "C" int sqsm_ball_lists(sqsm_handle h, const float* pts, int64_t m, const double* radii, int64_t* offsets,
                               int64_t* n_total, int32_t* indices, int64_t capacity) {
    SQSM_API_BEGIN(h)
    SQSM_CHECK(pts && radii && offsets && n_total && m > 0, "bad arguments");
    cudaStream_t st = h->cx.st;
    DBuf<float> d_pts(m * 3, st);
    DBuf<double> d_rad(m, st);
    SQSM_CUDA(cudaMemcpyAsync(d_pts.p, pts, sizeof(float) * 3 * m, cudaMemcpyHostToDevice, st));
    SQSM_CUDA(cudaMemcpyAsync(d_rad.p, radii, sizeof(double) * m, cudaMemcpyHostToDevice, st));
    // finest grid: cell edge = the mean of the finite positive radii; coarsest: an eighth of the largest finite radius
    double finest = 0, largest = 0;
    int64_t nf = 0;
    for (int64_t i = 0; i < m; ++i) {
        const double a = std::fabs(radii[i]);
        if (a > 0 && a < 1e30) { finest += a; ++nf; largest = std::max(largest, a); }
    }
    finest = nf ? finest / (double)nf : 0.0;
    DBuf<int64_t> d_off;
    DBuf<uint32_t> d_idx;
    int64_t T = 0;
    {
        ProfScope ps(h->cx, "ball_lists", (double)m * 20.0, 0.0);
        ball_lists(h->cx, d_pts.p, m, d_rad.p, finest, largest, d_off, d_idx, &T);
    }
    *n_total = T;
    SQSM_CUDA(cudaMemcpyAsync(offsets, d_off.p, sizeof(int64_t) * (m + 1), cudaMemcpyDeviceToHost, st));
    if (indices && T > 0) {
        SQSM_CHECK(T <= capacity, "index buffer too small (n_total holds the required size)");
        SQSM_CUDA(cudaMemcpyAsync(indices, d_idx.p, sizeof(int32_t) * T, cudaMemcpyDeviceToHost, st));
    }
    SQSM_CUDA(cudaStreamSynchronize(st));
    SQSM_API_END(h)
}

// ---- connectivity graph + skeleton extraction (row a16 / f1)
extern "C" int sqsm_graph_construct(sqsm_handle h, const float* pts, int64_t m, const int32_t* simplices, int64_t n_simplices,
                                    double r, const int64_t* preset_edges, int64_t n_preset, int64_t* n_edges) {
    SQSM_API_BEGIN(h)
    SQSM_CHECK(simplices && n_simplices > 0 && n_edges, "bad arguments");
    cudaStream_t st = h->cx.st;
    DBuf<float> d_own;
    const float* d_pts = nullptr;
    if (pts) {                                               // caller-provided skeletal points
        SQSM_CHECK(m > 0, "bad point count");
        d_own.alloc(m * 3, st);
        SQSM_CUDA(cudaMemcpyAsync(d_own.p, pts, sizeof(float) * m * 3, cudaMemcpyHostToDevice, st));
        d_pts = d_own.p;
    } else {                                                 // the engine's last skeletal sample
        SQSM_CHECK(h->sk_valid, "call sqsm_skeletal_sample first (or pass the skeletal points)");
        m = h->sk.m;
        d_pts = h->sk.pts.p;
    }
    DBuf<int64_t> own_edges;
    const int64_t* d_radius = nullptr;
    int64_t n_radius = 0;
    if (r > 0) {
        ProfScope ps(h->cx, "radius_graph", (double)m * 12.0, 0.0);
        if (pts) {
            radius_graph(h->cx, d_pts, m, r, own_edges, &n_radius);
            d_radius = own_edges.p;
        } else {
            if (h->edges_r != r) {
                radius_graph(h->cx, d_pts, m, r, h->edges, &h->n_edges);
                h->edges_r = r;
            }
            d_radius = h->edges.p;
            n_radius = h->n_edges;
        }
    }
    {
        ProfScope ps(h->cx, "graph_construct", (double)n_simplices * 16.0 + (double)n_radius * 16.0, 0.0);
        graph_construct(h->cx, h->graph, d_pts, m, simplices, n_simplices, d_radius, n_radius, preset_edges, n_preset);
    }
    *n_edges = h->graph.n_edges;
    SQSM_API_END(h)
}

extern "C" int sqsm_graph_edges(sqsm_handle h, int64_t* edges) {
    SQSM_API_BEGIN(h)
    SQSM_CHECK(edges, "null output");
    graph_edges(h->cx, h->graph, edges);
    SQSM_API_END(h)
}

extern "C" int sqsm_graph_info(sqsm_handle h, int64_t* info5) {
    SQSM_API_BEGIN(h)
    SQSM_CHECK(h->graph.valid && info5, "no graph");
    info5[0] = h->graph.delaunay_root; info5[1] = h->graph.n_delaunay; info5[2] = h->graph.n_spt; info5[3] = h->graph.n_mst;
    info5[4] = h->graph.n_edges;
    SQSM_API_END(h)
}

extern "C" int sqsm_graph_extract_skeleton(sqsm_handle h, int32_t weight_fn, int64_t* root, int64_t* n_reached, int64_t* pred,
                                           int64_t* order, float* weight) {
    SQSM_API_BEGIN(h)
    ProfScope ps(h->cx, "graph_skeleton", (double)h->graph.n_edges * 16.0, 0.0);
    graph_extract_skeleton(h->cx, h->graph, weight_fn, root, n_reached, pred, order, weight);
    SQSM_API_END(h)
}

// ---------------------------------------------------------------------------------- inspection hooks

extern "C" int sqsm_debug_capture(sqsm_handle h, int32_t enable) {
    SQSM_API_BEGIN(h)
    h->capture = enable != 0;
    if (!enable) h->dbg = DebugCapture();
    SQSM_API_END(h)
}

extern "C" int sqsm_debug_level_size(sqsm_handle h, int32_t level, int64_t* n) {
    SQSM_API_BEGIN(h)
    SQSM_CHECK(h->dbg.valid && n && level >= 0 && level < kLevels, "no capture");
    *n = level < h->dbg.n_levels ? h->dbg.lv[level].n : 0;
    SQSM_API_END(h)
}

static void row_coords_host(sqsm_engine* h, const LevelTables& L, std::vector<int32_t>& coords) {
    auto bricks = L.map.bricks.to_host();
    auto rb = L.row_brick.to_host();
    auto rp = L.row_pos.to_host();
    (void)h;
    coords.resize((size_t)L.n * 4);
    for (int64_t r = 0; r < L.n; ++r) {
        uint64_t key = bricks[rb[r]].key;
        uint32_t pos = rp[r];
        coords[4 * r + 0] = (int32_t)key_s(key);
        coords[4 * r + 1] = (int32_t)key_z(key) * 8 + (int32_t)(pos >> 6);
        coords[4 * r + 2] = (int32_t)key_y(key) * 8 + (int32_t)((pos >> 3) & 7);
        coords[4 * r + 3] = (int32_t)key_x(key) * 8 + (int32_t)(pos & 7);
    }
}

extern "C" int sqsm_debug_voxels(sqsm_handle h, int32_t* indices, int64_t* winner, uint8_t* mask) {
    SQSM_API_BEGIN(h)
    SQSM_CHECK(h->dbg.valid, "no capture");
    const DebugCapture& d = h->dbg;
    std::vector<int32_t> coords;
    row_coords_host(h, d.lv[0], coords);
    auto canon = d.rows.canon.to_host();
    auto pt = d.rows.pt.to_host();
    auto in = d.rows.interior.to_host();
    for (int64_t r = 0; r < d.lv[0].n; ++r) {
        int64_t c = canon[r];
        if (indices) for (int k = 0; k < 4; ++k) indices[4 * c + k] = coords[4 * r + k];
        if (winner) winner[c] = pt[r];
        if (mask) mask[c] = in[r];
    }
    SQSM_API_END(h)
}

extern "C" int sqsm_debug_batch_shapes(sqsm_handle h, int32_t* shapes, int32_t* descend) {
    SQSM_API_BEGIN(h)
    SQSM_CHECK(h->dbg.valid, "no capture");
    const DebugCapture& d = h->dbg;
    if (shapes) {
        auto s = d.lv[0].shape.to_host();
        std::memcpy(shapes, s.data(), s.size() * sizeof(int32_t));
    }
    if (descend) {
        for (int l = 0; l < kLevels - 1; ++l) {
            for (int b = 0; b < d.n_batches; ++b) descend[l * d.n_batches + b] = 0;
            if (l < d.n_levels && d.lv[l].desc.n == d.n_batches) {
                auto f = d.lv[l].desc.to_host();
                for (int b = 0; b < d.n_batches; ++b) descend[l * d.n_batches + b] = f[b];
            }
        }
    }
    SQSM_API_END(h)
}

extern "C" int sqsm_debug_level_tables(sqsm_handle h, int32_t level, int32_t* coords, int32_t* nbr, int32_t* parent,
                                       int32_t* koff) {
    SQSM_API_BEGIN(h)
    SQSM_CHECK(h->dbg.valid && level >= 0 && level < h->dbg.n_levels, "no capture for this level");
    const LevelTables& L = h->dbg.lv[level];
    if (coords) {
        std::vector<int32_t> c;
        row_coords_host(h, L, c);
        std::memcpy(coords, c.data(), c.size() * sizeof(int32_t));
    }
    if (nbr) { auto v = L.nbr.to_host(); std::memcpy(nbr, v.data(), v.size() * sizeof(int32_t)); }
    if (parent) {
        SQSM_CHECK(L.parent.n == L.n, "level has no down-sample tables");
        auto v = L.parent.to_host();
        std::memcpy(parent, v.data(), v.size() * sizeof(int32_t));
    }
    if (koff) {
        SQSM_CHECK(L.koff.n == L.n, "level has no down-sample tables");
        auto v = L.koff.to_host();
        std::memcpy(koff, v.data(), v.size() * sizeof(int32_t));
    }
    SQSM_API_END(h)
}

extern "C" int sqsm_debug_canon_to_internal(sqsm_handle h, int32_t* perm) {
    SQSM_API_BEGIN(h)
    SQSM_CHECK(h->dbg.valid && perm, "no capture");
    auto canon = h->dbg.rows.canon.to_host();
    for (int64_t r = 0; r < (int64_t)canon.size(); ++r) perm[canon[r]] = (int32_t)r;
    SQSM_API_END(h)
}

extern "C" int sqsm_debug_predictions(sqsm_handle h, float* pred, float* offset, float* regression) {
    SQSM_API_BEGIN(h)
    SQSM_CHECK(h->dbg.valid && h->dbg.fb.pred.n > 0, "no capture");
    auto p = h->dbg.fb.pred.to_host();
    auto o = h->dbg.fb.offset.to_host();
    int64_t n0 = h->dbg.lv[0].n;
    for (int64_t r = 0; r < n0; ++r) {
        if (pred) for (int k = 0; k < 3; ++k) pred[3 * r + k] = p[4 * r + k];
        if (offset) for (int k = 0; k < 3; ++k) offset[3 * r + k] = o[4 * r + k];
        if (regression) regression[r] = o[4 * r + 3];
    }
    SQSM_API_END(h)
}

extern "C" int sqsm_debug_features(sqsm_handle h, const char* stage, int32_t* level, int32_t* channels, float* data,
                                   int64_t capacity) {
    SQSM_API_BEGIN(h)
    SQSM_CHECK(h->dbg.valid && stage, "no capture");
    auto it = h->dbg.fb.keep.find(stage);
    SQSM_CHECK(it != h->dbg.fb.keep.end(), "unknown stage");
    auto meta = h->dbg.fb.keep_meta.at(stage);
    if (level) *level = meta.first;
    if (channels) *channels = meta.second;
    if (data) {
        SQSM_CHECK(capacity >= it->second.n, "feature buffer too small");
        auto v = it->second.to_host();
        std::memcpy(data, v.data(), v.size() * sizeof(float));
    }
    SQSM_API_END(h)
}

// ---------------------------------------------------------------------------------- measurement hooks

static void resolve_profile(sqsm_engine* h) {
    Ctx& cx = h->cx;
    if (cx.pending.empty()) return;
    SQSM_CUDA(cudaStreamSynchronize(cx.st));
    for (auto& r : cx.pending) {
        float ms = 0;
        cudaEventElapsedTime(&ms, r.a, r.b);
        ConvProfile& c = h->prof[cx.fam_names[r.fam]];
        c.ms += ms; c.launches += r.launches; c.bytes += r.bytes; c.flops += r.flops;
        cudaEventDestroy(r.a); cudaEventDestroy(r.b);
    }
    cx.pending.clear();
}

extern "C" int sqsm_profile_enable(sqsm_handle h, int32_t enable) {
    SQSM_API_BEGIN(h)
    h->cx.prof_on = enable != 0;
    SQSM_API_END(h)
}

extern "C" int sqsm_profile_reset(sqsm_handle h) {
    SQSM_API_BEGIN(h)
    resolve_profile(h);
    h->prof.clear();
    SQSM_API_END(h)
}

extern "C" int sqsm_profile_get(sqsm_handle h, const char* family, double* ms, int64_t* launches, double* bytes,
                                double* flops) {
    SQSM_API_BEGIN(h)
    SQSM_CHECK(family, "null family");
    resolve_profile(h);
    ConvProfile p;
    auto it = h->prof.find(family);
    if (it != h->prof.end()) p = it->second;
    if (ms) *ms = p.ms;
    if (launches) *launches = p.launches;
    if (bytes) *bytes = p.bytes;
    if (flops) *flops = p.flops;
    SQSM_API_END(h)
}

extern "C" int sqsm_profile_families(sqsm_handle h, char* buf, int64_t capacity) {
    SQSM_API_BEGIN(h)
    resolve_profile(h);
    std::string all;
    for (auto& kv : h->prof) { if (!all.empty()) all += ","; all += kv.first; }
    SQSM_CHECK(buf && capacity > (int64_t)all.size(), "buffer too small");
    std::memcpy(buf, all.c_str(), all.size() + 1);
    SQSM_API_END(h)
}

extern "C" int sqsm_timer_start(sqsm_handle h) {
    SQSM_API_BEGIN(h)
    if (!h->tm_a) { SQSM_CUDA(cudaEventCreate(&h->tm_a)); SQSM_CUDA(cudaEventCreate(&h->tm_b)); }
    SQSM_CUDA(cudaStreamSynchronize(h->cx.st));
    SQSM_CUDA(cudaEventRecord(h->tm_a, h->cx.st));
    SQSM_API_END(h)
}

extern "C" int sqsm_timer_stop(sqsm_handle h, double* ms) {
    SQSM_API_BEGIN(h)
    SQSM_CHECK(h->tm_a && ms, "timer not started");
    SQSM_CUDA(cudaEventRecord(h->tm_b, h->cx.st));
    SQSM_CUDA(cudaEventSynchronize(h->tm_b));
    float f = 0;
    SQSM_CUDA(cudaEventElapsedTime(&f, h->tm_a, h->tm_b));
    *ms = f;
    SQSM_API_END(h)
}

extern "C" int sqsm_get_stream(sqsm_handle h, void** stream) {
    SQSM_API_BEGIN(h)
    SQSM_CHECK(stream, "null argument");
    *stream = (void*)h->cx.st;
    SQSM_API_END(h)
}

extern "C" int sqsm_launch_count(sqsm_handle h, int64_t* n) {
    SQSM_API_BEGIN(h)
    SQSM_CHECK(n, "null argument");
    *n = h->cx.launches;
    SQSM_API_END(h)
}
