// BrickMap: the coordinate structure behind voxelisation, rulebooks, the Chamfer monitor,
// skeletal sampling and the radius graph.
//
// Two levels: an open-addressing hash (linear probing, 64-bit keys, <= 50 % load) maps the
// coordinates of an 8x8x8-voxel brick to a brick id; a brick is one 128-byte line holding a
// 512-bit occupancy bitmap plus popcount prefixes, so "is voxel v active and what is its
// row" is one hash probe per brick and a popcount.  Bricks are numbered in sorted key order
// (sample, bz, by, bx) and rows follow brick order, then (z, y, x) inside the brick, so rows
// that are neighbours in space are neighbours in memory and the gathers of the sparse
// convolutions hit L2.
#pragma once
#include "common.cuh"
#include <string>
#include <stdlib.h>
#include <time.h>
#include <vector>

namespace sqsm {

struct ProfRec { cudaEvent_t a, b; int fam; int64_t launches; double bytes, flops; };

struct Ctx {                       // per-engine launch context
    cudaStream_t st = nullptr;
    int64_t launches = 0;
    bool prof_on = false;          // kernel-family timing (events are resolved lazily, no sync here)
    std::vector<std::string> fam_names;
    std::vector<ProfRec> pending;
    int family(const std::string& name) {
        for (size_t i = 0; i < fam_names.size(); ++i) if (fam_names[i] == name) return (int)i;
        fam_names.push_back(name);
        return (int)fam_names.size() - 1;
    }
};

// RAII: CUDA events around a kernel family on the launching stream.
struct ProfScope {
    Ctx& cx; ProfRec r; bool on;
    ProfScope(Ctx& c, const char* fam, double bytes, double flops) : cx(c), on(c.prof_on) {
        if (!on) return;
        r.fam = cx.family(std::string(fam)); r.bytes = bytes; r.flops = flops; r.launches = cx.launches;
        cudaEventCreate(&r.a); cudaEventCreate(&r.b);
        cudaEventRecord(r.a, cx.st);
    }
    ~ProfScope() {
        if (!on) return;
        cudaEventRecord(r.b, cx.st);
        r.launches = cx.launches - r.launches;
        cx.pending.push_back(r);
    }
};

// SQSM_TRACE=1: host wall-clock per stage, with and without draining the stream (finds host stalls)
struct Trace {
    Ctx& cx; const char* name; double t0; bool on;
    static bool enabled() { static int e = -1; if (e < 0) { const char* v = getenv("SQSM_TRACE"); e = (v && v[0] == '1') ? 1 : 0; } return e == 1; }
    static double now() { timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6; }
    Trace(Ctx& c, const char* n) : cx(c), name(n), on(enabled()) { if (on) { cudaStreamSynchronize(cx.st); t0 = now(); } }
    ~Trace() {
        if (!on) return;
        double t1 = now();
        cudaStreamSynchronize(cx.st);
        double t2 = now();
        fprintf(stderr, "[trace] %-22s host %8.3f ms  +drain %8.3f ms\n", name, t1 - t0, t2 - t1);
    }
};

struct BrickMap {
    DBuf<uint64_t> hkeys;
    DBuf<int32_t>  hvals;
    uint32_t cap = 0;
    DBuf<Brick> bricks;
    DBuf<int32_t> bnbr;            // [n_bricks][32]: ids of the 27 neighbour bricks (-1 = none)
    int n_bricks = 0;
    int64_t n_rows = 0;
#ifdef __CUDACC__
    BrickHash view() const { return BrickHash{hkeys.p, hvals.p, cap - 1}; }
#endif
};

// Build a BrickMap from entries (brick key, position in brick).  Entries whose key is
// kEmptyKey are ignored.  On return ent_brick[e] holds the brick id of every valid entry.
// Synchronises the stream twice (brick count, row count).
void brickmap_build(Ctx& cx, BrickMap& m, const uint64_t* d_ent_key, const uint16_t* d_ent_pos,
                    int64_t n_ent, int32_t* d_ent_brick, int64_t expected_bricks = 0);

// bnbr table (27 neighbour brick ids per brick); needed by the rulebook / NN / graph kernels.
void brickmap_build_neighbours(Ctx& cx, BrickMap& m);

// row -> (brick id, position) expansion; arrays sized n_rows.
void brickmap_expand_rows(Ctx& cx, const BrickMap& m, int32_t* d_row_brick, uint16_t* d_row_pos);

// small utilities shared by the translation units
void exclusive_scan_i32(Ctx& cx, const int32_t* d_in, int32_t* d_out, int64_t n);   // out has n elements
void exclusive_scan_i64(Ctx& cx, const int64_t* d_in, int64_t* d_out, int64_t n);
void sort_keys_u64(Ctx& cx, const uint64_t* d_in, uint64_t* d_out, int64_t n, int begin_bit, int end_bit);
void sort_pairs_u32(Ctx& cx, const uint32_t* d_kin, uint32_t* d_kout, const uint32_t* d_vin, uint32_t* d_vout,
                    int64_t n, int begin_bit, int end_bit);
// ascending sort of the keys inside every segment [d_off[s] - base, d_off[s + 1] - base) (cub::DeviceSegmentedSort)
void segmented_sort_u32(Ctx& cx, const uint32_t* d_in, uint32_t* d_out, int64_t n, int64_t n_segments, const int64_t* d_off,
                        int64_t base = 0);

}  // namespace sqsm
