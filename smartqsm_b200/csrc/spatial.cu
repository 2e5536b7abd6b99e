// Chamfer monitor, skeletal sampling + radii, radius-neighbour edge set.  sm_100a.
//
// Replaces, on the device:
//   Chamfer monitor      core/core_algorithms/spconv_based_contraction.py:112-128
//                        (Open3D voxel_down_sample_and_trace x2 + compute_point_cloud_distance x2)
//   skeletal sampling    core/thinning_based_skeletonization_pipeline.py:68-80
//   radius edges         core/skeletonization_pipeline_base.py:132-144 merged/sorted at :158
//                        (SciPy cKDTree.query_ball_point + Python double loop + np.unique)
// All three run on the BrickMap: an 8-voxel brick is the search cell; exact nearest neighbours
// are found by growing rings of bricks until the best distance is below the ring's lower bound.
#include "engine.h"
#include <cub/cub.cuh>
#include <cmath>
#include <cstring>
#include <limits>
#include <memory>
#include <vector>

namespace sqsm {

// ---------------------------------------------------------------------------------- bounds

__device__ __forceinline__ uint32_t f2ord(float f) {
    uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
static float ord2f(uint32_t o) {
    uint32_t u = (o & 0x80000000u) ? (o & 0x7FFFFFFFu) : ~o;
    float f;
    memcpy(&f, &u, 4);
    return f;
}

// per-axis min / max of an [n][3] float cloud -> 6 order-preserving uint32 (min x,y,z, max x,y,z)
__global__ void __launch_bounds__(256) k_bounds(const float* __restrict__ P, int64_t n, uint32_t* __restrict__ out) {
    uint32_t mn[3] = {0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu}, mx[3] = {0, 0, 0};
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        #pragma unroll
        for (int a = 0; a < 3; ++a) {
            uint32_t o = f2ord(P[3 * i + a]);
            mn[a] = min(mn[a], o);
            mx[a] = max(mx[a], o);
        }
    }
    #pragma unroll
    for (int a = 0; a < 3; ++a) {
        mn[a] = __reduce_min_sync(0xFFFFFFFFu, mn[a]);
        mx[a] = __reduce_max_sync(0xFFFFFFFFu, mx[a]);
    }
    if ((threadIdx.x & 31) == 0) {
        #pragma unroll
        for (int a = 0; a < 3; ++a) {
            atomicMin(&out[a], mn[a]);
            atomicMax(&out[3 + a], mx[a]);
        }
    }
}

static void cloud_bounds(Ctx& cx, const float* d_P, int64_t n, double mn[3], double mx[3]) {
    DBuf<uint32_t> b(6, cx.st);
    uint32_t init[6] = {0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0, 0, 0};
    SQSM_CUDA(cudaMemcpyAsync(b.p, init, sizeof(init), cudaMemcpyHostToDevice, cx.st));
    int grid = std::min<int64_t>(div_up(n, 256), kNumSMs * 8);
    k_bounds<<<grid, 256, 0, cx.st>>>(d_P, n, b.p);
    cx.launches++;
    uint32_t h[6];
    SQSM_CUDA(cudaMemcpyAsync(h, b.p, sizeof(h), cudaMemcpyDeviceToHost, cx.st));
    SQSM_CUDA(cudaStreamSynchronize(cx.st));
    for (int a = 0; a < 3; ++a) { mn[a] = (double)ord2f(h[a]); mx[a] = (double)ord2f(h[3 + a]); }
}

// ---------------------------------------------------------------------------------- global voxel grid

struct GridParams { double mn[3]; double voxel; };

// Open3D voxel_down_sample_and_trace: voxel index = floor((p - min_bound) / voxel) in float64
__device__ __forceinline__ bool grid_voxel(const float* __restrict__ P, int64_t i, const GridParams& g, int v[3]) {
    bool ok = true;
    #pragma unroll
    for (int a = 0; a < 3; ++a) {
        double f = floor(__ddiv_rn(__dsub_rn((double)P[3 * i + a], g.mn[a]), g.voxel));
        ok = ok && (f >= 0.0) && (f < 524288.0);          // 65536 bricks x 8 voxels per axis
        v[a] = (int)f;
    }
    return ok;
}

__global__ void __launch_bounds__(256) k_grid_entries(const float* __restrict__ P, int64_t n, const __grid_constant__ GridParams g,
                                                      uint64_t* __restrict__ ent_key, uint16_t* __restrict__ ent_pos,
                                                      int32_t* __restrict__ err) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int v[3];
    if (!grid_voxel(P, i, g, v)) { ent_key[i] = kEmptyKey; ent_pos[i] = 0; atomicOr(err, 1); return; }
    ent_key[i] = pack_key(0u, (uint32_t)(v[2] >> 3), (uint32_t)(v[1] >> 3), (uint32_t)(v[0] >> 3));
    ent_pos[i] = (uint16_t)pos_of(v[2] & 7, v[1] & 7, v[0] & 7);
}

struct GridMap {
    BrickMap map;
    DBuf<int32_t> pt_row;      // [n] voxel row of every point
    DBuf<int32_t> ent_brick;   // [n] brick of every point (-1: outside the grid)
    DBuf<uint16_t> ent_pos;    // [n] voxel position inside the brick
    GridParams g;
};

__global__ void __launch_bounds__(256) k_point_rows(const Brick* __restrict__ bricks, const int32_t* __restrict__ ent_brick,
                                                    const uint16_t* __restrict__ ent_pos, int64_t n, int32_t* __restrict__ pt_row) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int b = ent_brick[i];
    pt_row[i] = (b < 0) ? -1 : brick_row(&bricks[b], ent_pos[i]);
}

static void grid_build(Ctx& cx, const float* d_P, int64_t n, const double mn[3], double voxel, GridMap& gm) {
    cudaStream_t st = cx.st;
    for (int a = 0; a < 3; ++a) gm.g.mn[a] = mn[a];
    gm.g.voxel = voxel;
    DBuf<uint64_t> ent_key(n, st);
    gm.ent_pos.alloc(n, st);
    gm.ent_brick.alloc(n, st);
    DBuf<int32_t> err(1, st);
    err.zero();
    k_grid_entries<<<div_up(n, 256), 256, 0, st>>>(d_P, n, gm.g, ent_key.p, gm.ent_pos.p, err.p);
    cx.launches++;
    brickmap_build(cx, gm.map, ent_key.p, gm.ent_pos.p, n, gm.ent_brick.p, n / 16);
    int32_t herr = 0;
    SQSM_CUDA(cudaMemcpyAsync(&herr, err.p, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    gm.pt_row.alloc(n, st);
    k_point_rows<<<div_up(n, 256), 256, 0, st>>>(gm.map.bricks.p, gm.ent_brick.p, gm.ent_pos.p, n, gm.pt_row.p);
    cx.launches++;
    SQSM_CUDA(cudaStreamSynchronize(st));
    SQSM_CHECK(herr == 0, "cloud extent exceeds 524288 voxels per axis (or NaN coordinates)");
}

// ---------------------------------------------------------------------------------- Chamfer monitor

// Voxel means (Open3D AccumulatedPoint: sum / count).  Order-independent and therefore bit-reproducible: every point adds
// its offset from the origin of its OWN voxel, in fixed point (2^-58 m: the offset is below one voxel edge, so even 2^20
// points per voxel cannot overflow 63 bits for edges up to 0.5 m), with 64-bit integer atomics; the mean is the origin
// plus the exactly accumulated offset sum over the count, rounded once.  Open3D's sequential float64 sum differs from it
// by its own rounding only (<= ~1e-15 m).
constexpr double kFixScale = 288230376151711744.0;               // 2^58

__device__ __forceinline__ void voxel_origin(const Brick* __restrict__ bricks, int brick, uint32_t pos, const GridParams& g,
                                             double o[3]) {
    const uint64_t key = bricks[brick].key;
    o[0] = g.mn[0] + (double)((int)key_x(key) * 8 + (int)(pos & 7u)) * g.voxel;
    o[1] = g.mn[1] + (double)((int)key_y(key) * 8 + (int)((pos >> 3) & 7u)) * g.voxel;
    o[2] = g.mn[2] + (double)((int)key_z(key) * 8 + (int)(pos >> 6)) * g.voxel;
}

__global__ void __launch_bounds__(256) k_voxel_accumulate(const float* __restrict__ P, const int32_t* __restrict__ pt_row,
                                                          const int32_t* __restrict__ ent_brick, const uint16_t* __restrict__ ent_pos,
                                                          const Brick* __restrict__ bricks, const __grid_constant__ GridParams g,
                                                          int64_t n, long long* __restrict__ sum, int32_t* __restrict__ cnt) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int r = pt_row[i];
    if (r < 0) return;
    double o[3];
    voxel_origin(bricks, ent_brick[i], ent_pos[i], g, o);
    #pragma unroll
    for (int a = 0; a < 3; ++a) {
        const long long q = __double2ll_rn(((double)P[3 * i + a] - o[a]) * kFixScale);
        atomicAdd(reinterpret_cast<unsigned long long*>(&sum[3 * (size_t)r + a]), (unsigned long long)q);
    }
    atomicAdd(&cnt[r], 1);
}

// mean (float64) and the float32 copy relative to the brick origin (screening copy of the NN search) in one pass
__global__ void __launch_bounds__(256) k_voxel_mean(const long long* __restrict__ sum, const int32_t* __restrict__ cnt,
                                                    const int32_t* __restrict__ row_brick, const uint16_t* __restrict__ row_pos,
                                                    const Brick* __restrict__ bricks, const __grid_constant__ GridParams g,
                                                    int64_t rows, double* __restrict__ mean, float4* __restrict__ loc) {
    int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= rows) return;
    const int b = row_brick[r];
    const uint32_t pos = row_pos[r];
    double o[3];
    voxel_origin(bricks, b, pos, g, o);
    const double c = (double)cnt[r] * kFixScale;
    const double edge = 8.0 * g.voxel;
    const uint64_t key = bricks[b].key;
    const double bo[3] = {g.mn[0] + (double)key_x(key) * edge, g.mn[1] + (double)key_y(key) * edge, g.mn[2] + (double)key_z(key) * edge};
    double m[3];
    #pragma unroll
    for (int a = 0; a < 3; ++a) {
        m[a] = o[a] + __ddiv_rn((double)sum[3 * r + a], c);
        mean[3 * r + a] = m[a];
    }
    loc[r] = make_float4((float)(m[0] - bo[0]), (float)(m[1] - bo[1]), (float)(m[2] - bo[2]), 0.f);
}

// For every brick of map A: (first row, count) of the 27 bricks of map B around the same coordinates.
__global__ void __launch_bounds__(256) k_cross_neighbours(const Brick* __restrict__ A, int nA, BrickHash hB,
                                                          const Brick* __restrict__ B, int2* __restrict__ abn,
                                                          int32_t* __restrict__ abid, int64_t row_lo, int64_t row_hi) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    int a = t >> 5, j = t & 31;
    if (a >= nA) return;
    if ((int64_t)A[a].base + A[a].count <= row_lo || (int64_t)A[a].base >= row_hi) return;   // no query row of this rank
    int2 r = make_int2(0, 0);
    int id = -1;
    if (j < 27) {
        uint64_t key = A[a].key;
        int z = (int)key_z(key) + j / 9 - 1, y = (int)key_y(key) + (j / 3) % 3 - 1, x = (int)key_x(key) + j % 3 - 1;
        if (z >= 0 && y >= 0 && x >= 0 && z < 65536 && y < 65536 && x < 65536) {
            id = hash_find(hB, pack_key(0u, (uint32_t)z, (uint32_t)y, (uint32_t)x));
            if (id >= 0) r = make_int2(B[id].base, B[id].count);
        }
    }
    abid[(size_t)a * 32 + j] = id;                                                    // brick ids (cell-ring search)
    const unsigned present = __ballot_sync(0xFFFFFFFFu, r.y > 0);
    if (j == 27) r.x = (int)present;                                                 // entry 27: which of the 27 exist
    abn[(size_t)a * 32 + j] = r;
}

// First stage of the exact 1-NN search: both clouds live on the SAME voxel grid (:116-121), a voxel mean lies inside its
// voxel, so the candidates of a query are enumerated by growing cubes of voxel CELLS around the query's own cell, read
// straight from the occupancy bitmaps of B's bricks (one 8-bit row of a brick line per (z, y) row, a popcount gives the
// row of an occupied cell).  After the (2R+1)^3 cube every unseen mean is farther than the query's distance to the faces
// of that cube; if the best candidate is closer (minus 1 nm of slack for the rounding of means on cell faces) the query
// is settled - about half of them at R = 2 (~20 candidates instead of the ~110 of a brick-level search).  Unsettled
// queries are appended to a list for the next stage.  One thread per query; neighbouring queries share bricks (L1 hits).
template <int R>
__global__ void __launch_bounds__(256) k_nn_cells(const Brick* __restrict__ A, const int32_t* __restrict__ row_brick,
                                                  const uint16_t* __restrict__ row_pos, const int32_t* __restrict__ list,
                                                  int64_t n, const double* __restrict__ meanA, const int32_t* __restrict__ abid,
                                                  const Brick* __restrict__ B, const double* __restrict__ meanB,
                                                  const __grid_constant__ GridParams g, int64_t row_lo, int64_t row_hi,
                                                  double* __restrict__ nn2, int32_t* __restrict__ unsettled,
                                                  int32_t* __restrict__ n_unsettled) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n) return;
    const int64_t r = list ? (int64_t)list[idx] : idx;
    if (r < row_lo || r >= row_hi) return;               // this rank's share of the query rows
    const int a = row_brick[r];
    const uint32_t pos = row_pos[r];
    const int z = (int)(pos >> 6), y = (int)(pos >> 3) & 7, x = (int)pos & 7;
    const double qx = meanA[3 * r], qy = meanA[3 * r + 1], qz = meanA[3 * r + 2];
    const int32_t* __restrict__ ids = abid + (size_t)a * 32;
    double best = 1e300;
    #pragma unroll 1
    for (int dz = -R; dz <= R; ++dz) {
        const int zz = z + dz, bzo = zz < 0 ? 0 : (zz > 7 ? 2 : 1), zl = zz & 7;
        #pragma unroll 1
        for (int dy = -R; dy <= R; ++dy) {
            const int yy = y + dy, byo = yy < 0 ? 0 : (yy > 7 ? 2 : 1), yl = yy & 7;
            const int w = zl * 2 + (yl >> 2), sh = (yl & 3) * 8;
            #pragma unroll
            for (int seg = 0; seg < 3; ++seg) {
                int lo, hi;                                       // x range inside brick column seg (0 left, 1 own, 2 right)
                if (seg == 0) { lo = x - R + 8; hi = 7; }
                else if (seg == 1) { lo = max(0, x - R); hi = min(7, x + R); }
                else { lo = 0; hi = x + R - 8; }
                if (lo > hi) continue;
                const int id = ids[bzo * 9 + byo * 3 + seg];
                if (id < 0) continue;
                const Brick* __restrict__ b = &B[id];
                const uint32_t word = b->mask[w];
                uint32_t bits = (word >> sh) & 0xFFu & (((1u << (hi - lo + 1)) - 1u) << lo);
                if (!bits) continue;
                const int rbase = b->base + (int)b->prefix[w];
                while (bits) {
                    const int bit = sh + __ffs(bits) - 1;
                    bits &= bits - 1;
                    const size_t row = (size_t)(rbase + __popc(word & ((1u << bit) - 1u)));
                    const double ex = qx - meanB[3 * row], ey = qy - meanB[3 * row + 1], ez = qz - meanB[3 * row + 2];
                    best = fmin(best, ex * ex + ey * ey + ez * ez);
                }
            }
        }
    }
    // distance from the query to the outside of the scanned cube of cells
    const uint64_t key = A[a].key;
    const int cx = (int)key_x(key) * 8 + x, cy = (int)key_y(key) * 8 + y, cz = (int)key_z(key) * 8 + z;
    double lb = fmin(qx - (g.mn[0] + (double)(cx - R) * g.voxel), (g.mn[0] + (double)(cx + R + 1) * g.voxel) - qx);
    lb = fmin(lb, fmin(qy - (g.mn[1] + (double)(cy - R) * g.voxel), (g.mn[1] + (double)(cy + R + 1) * g.voxel) - qy));
    lb = fmin(lb, fmin(qz - (g.mn[2] + (double)(cz - R) * g.voxel), (g.mn[2] + (double)(cz + R + 1) * g.voxel) - qz));
    lb -= 1e-9;
    if (lb > 0.0 && best <= lb * lb) nn2[r] = best;
    else unsettled[atomicAdd(n_unsettled, 1)] = (int32_t)r;
}

// Exact 1-NN squared distances from the voxel means of map A to those of map B, summed.
// Bricks of B are visited nearest first (own brick, faces, edges + corners); the query lies inside its brick, so after
// the 27-neighbourhood everything else is at least one brick edge away; a query whose best distance exceeds that bound
// keeps growing rings of bricks (hash probes) until it does not - exact 1-NN like the KD-tree search of Open3D's
// compute_point_cloud_distance.
// The monitor's nearest neighbours are close (median 2 cm, 99 % < 5.5 cm
// against 8 cm bricks), so a query can prune most of the 27 bricks - but only with ITS OWN bound: a warp-wide vote
// almost never prunes because some query of the warp sits next to every face, and a plain
// thread-per-query loop diverges to the same union of bricks (ncu: 8 of 32 lanes active).  Here a warp owns 32
// consecutive queries and works in three rounds - own brick, face bricks, edge+corner bricks.  In every round each lane
// lists the bricks its query still needs (float32 lower bound against its current best) into a per-warp shared list,
// and the (query, brick) pairs of that list are processed 8 at a time by 4-lane groups: float32 screening in
// brick-local coordinates, float64 re-evaluation of anything that could win, shared-memory atomicMin on the owner's
// best.  The ring fallback keeps the search exact.
struct NNWarpShared {
    float4 ql[32];
    unsigned long long best[32];
    int brick[32];
    int row[32];
    uint16_t list[32 * 20];
};
__device__ __forceinline__ float nn_gap2(int nb, const float* gm2, const float* gp2) {
    const int dz = nb / 9 - 1, dy = (nb / 3) % 3 - 1, dx = nb % 3 - 1;
    return ((dx == 0 ? 0.f : (dx < 0 ? gm2[0] : gp2[0])) + (dy == 0 ? 0.f : (dy < 0 ? gm2[1] : gp2[1])) +
            (dz == 0 ? 0.f : (dz < 0 ? gm2[2] : gp2[2]))) * 0.999999f;
}
// exact squared distance in float64, one fixed operation order for every kernel of the search (bit-identical results)
__device__ __forceinline__ double nn_exact_d2(const double* __restrict__ q, const double* __restrict__ c) {
    const double ex = q[0] - c[0], ey = q[1] - c[1], ez = q[2] - c[2];
    return __fma_rn(ez, ez, __fma_rn(ey, ey, __dmul_rn(ex, ex)));
}

__global__ void __launch_bounds__(256) k_nn_sqdist_pairs(const Brick* __restrict__ A, const int32_t* __restrict__ row_brick,
                                                         const int32_t* __restrict__ qlist, int64_t n_rows,
                                                         const double* __restrict__ meanA,
                                                         const float4* __restrict__ locA, const int2* __restrict__ abn,
                                                         BrickHash hB, const Brick* __restrict__ B,
                                                         const double* __restrict__ meanB, const float4* __restrict__ locB,
                                                         const __grid_constant__ GridParams g, int max_ring, int64_t row_lo,
                                                         int64_t row_hi, double* __restrict__ nn2,
                                                         unsigned long long* __restrict__ n_fallback) {
    // queries = the rows listed in qlist (n_rows entries; unsettled by the cell-ring stages) or all rows 0 .. n_rows-1;
    // the exact squared 1-NN distance of query row r goes to nn2[r]
    __shared__ NNWarpShared s_warp[8];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    NNWarpShared& S = s_warp[w];
    const double edge = 8.0 * g.voxel;
    const float edge_f = (float)edge;
    const float slack = 1e-6f * edge_f;
    constexpr unsigned kRoundMask[3] = {1u << 13,
                                        (1u << 4) | (1u << 10) | (1u << 12) | (1u << 14) | (1u << 16) | (1u << 22),
                                        0x07FFFFFFu & ~((1u << 13) | (1u << 4) | (1u << 10) | (1u << 12) | (1u << 14) | (1u << 16) | (1u << 22))};
    unsigned fb = 0;
    // a rank's share of the queries is a contiguous ROW range (full warps; a share by brick id modulo the rank count left
    // 4 of 32 lanes valid at 8 ranks): without a list the loop walks that range only
    const int64_t first = qlist ? 0 : (row_lo & ~(int64_t)31);
    if (!qlist) n_rows = row_hi < n_rows ? row_hi : n_rows;
    for (int64_t base = first + ((int64_t)blockIdx.x * 8 + w) * 32; base < n_rows; base += (int64_t)gridDim.x * 256) {
        bool valid = base + lane < n_rows;
        const int64_t r = valid ? (qlist ? (int64_t)qlist[base + lane] : base + lane) : 0;
        if (r < row_lo || r >= row_hi) valid = false;
        const int a = valid ? row_brick[r] : 0;
        if (!__any_sync(0xFFFFFFFFu, valid)) continue;
        const float4 ql = valid ? locA[r] : make_float4(0.f, 0.f, 0.f, 0.f);
        float gm2[3] = {fmaxf(ql.x - slack, 0.f), fmaxf(ql.y - slack, 0.f), fmaxf(ql.z - slack, 0.f)};
        float gp2[3] = {fmaxf(edge_f - ql.x - slack, 0.f), fmaxf(edge_f - ql.y - slack, 0.f), fmaxf(edge_f - ql.z - slack, 0.f)};
        #pragma unroll
        for (int d = 0; d < 3; ++d) { gm2[d] *= gm2[d]; gp2[d] *= gp2[d]; }
        const unsigned present = valid ? (unsigned)abn[(size_t)a * 32 + 27].x : 0u;
        __syncwarp();
        S.ql[lane] = ql;
        S.brick[lane] = a;
        S.row[lane] = (int)r;
        S.best[lane] = (unsigned long long)__double_as_longlong(1e300);
        double best = 1e300;
        #pragma unroll 1
        for (int round = 0; round < 3; ++round) {
            // the bricks of this round my query still needs
            unsigned m = present & kRoundMask[round];
            if (round > 0 && m) {
                const float blf = best < 1e30 ? (float)best : 3.0e38f;
                unsigned keep = 0;
                for (unsigned mm = m; mm;) {
                    const int nb = __ffs(mm) - 1;
                    mm &= mm - 1;
                    if (nn_gap2(nb, gm2, gp2) < blf) keep |= 1u << nb;
                }
                m = keep;
            }
            const int cnt = __popc(m);
            int incl = cnt;
            #pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int v = __shfl_up_sync(0xFFFFFFFFu, incl, d);
                if (lane >= d) incl += v;
            }
            const int total = __shfl_sync(0xFFFFFFFFu, incl, 31);
            if (total == 0) continue;
            int off = incl - cnt;
            while (m) {
                const int nb = __ffs(m) - 1;
                m &= m - 1;
                S.list[off++] = (uint16_t)((lane << 5) | nb);
            }
            __syncwarp();
            const int grp = lane >> 2, sub = lane & 3;
            for (int k0 = 0; k0 < total; k0 += 8) {
                const int k = k0 + grp;
                const bool on = k < total;
                double bl = 1e300, cur = 1e300;
                int o = 0;
                if (on) {
                    const int e = S.list[k];
                    o = e >> 5;
                    const int nb = e & 31;
                    cur = __longlong_as_double((long long)S.best[o]);
                    bl = cur;
                    const float4 q = S.ql[o];
                    const int2 rc = abn[(size_t)S.brick[o] * 32 + nb];
                    const int dz = nb / 9 - 1, dy = (nb / 3) % 3 - 1, dx = nb % 3 - 1;
                    const float sx = q.x - (float)dx * edge_f, sy = q.y - (float)dy * edge_f, sz = q.z - (float)dz * edge_f;
                    const double* qd = meanA + 3 * (size_t)S.row[o];
                    const double* c = meanB + 3 * (size_t)rc.x;
                    const float4* cl = locB + (size_t)rc.x;
                    float blf = bl < 1e30 ? (float)bl * 1.00004f + 1e-9f : 3.0e38f;
                    for (int i = sub; i < rc.y; i += 4) {
                        const float4 p = __ldg(cl + i);
                        const float fx = sx - p.x, fy = sy - p.y, fz = sz - p.z;
                        const float d2f = fx * fx + fy * fy + fz * fz;
                        if (d2f <= blf) {                       // float32 error ~2e-6 relative: could win -> float64
                            bl = fmin(bl, nn_exact_d2(qd, c + 3 * i));
                            blf = (float)bl * 1.00004f + 1e-9f;
                        }
                    }
                }
                bl = fmin(bl, __shfl_xor_sync(0xFFFFFFFFu, bl, 1));
                bl = fmin(bl, __shfl_xor_sync(0xFFFFFFFFu, bl, 2));
                if (on && sub == 0 && bl < cur) atomicMin(&S.best[o], (unsigned long long)__double_as_longlong(bl));
            }
            __syncwarp();
            best = __longlong_as_double((long long)S.best[lane]);
        }
        // exact ring fallback (a few queries per thousand): the warp serves its unsettled queries one at a time, the 32 lanes
        // sharing the bricks of every ring (hash probes) instead of one lane probing alone while 31 wait
        unsigned need = __ballot_sync(0xFFFFFFFFu, valid && best > edge * edge);
        fb += (need >> lane) & 1u;
        while (need) {
            const int src = __ffs(need) - 1;
            need &= need - 1;
            const int64_t rq = S.row[src];
            const double* qdr = meanA + 3 * rq;
            const uint64_t key = A[__shfl_sync(0xFFFFFFFFu, a, src)].key;
            const int bz = (int)key_z(key), by = (int)key_y(key), bx = (int)key_x(key);
            double bq = __shfl_sync(0xFFFFFFFFu, best, src);
            int ring = 1;
            while (bq > (double)ring * edge * (double)ring * edge && ring < max_ring) {
                ++ring;
                const int side = 2 * ring + 1;
                double mine = bq;
                for (int t = lane; t < side * side; t += 32) {
                    const int dz = t / side - ring, dy = t % side - ring;
                    const bool shell = (abs(dz) == ring) || (abs(dy) == ring);
                    const int z = bz + dz, y = by + dy;
                    if (z < 0 || y < 0 || z > 65535 || y > 65535) continue;
                    for (int dx = -ring; dx <= ring; dx += (shell ? 1 : 2 * ring)) {
                        const int x = bx + dx;
                        if (x < 0 || x > 65535) continue;
                        const int id = hash_find(hB, pack_key(0u, (uint32_t)z, (uint32_t)y, (uint32_t)x));
                        if (id < 0) continue;
                        const int cb = B[id].base, cn = B[id].count;
                        for (int i = 0; i < cn; ++i) mine = fmin(mine, nn_exact_d2(qdr, meanB + 3 * (size_t)(cb + i)));
                    }
                }
                #pragma unroll
                for (int d = 16; d > 0; d >>= 1) mine = fmin(mine, __shfl_xor_sync(0xFFFFFFFFu, mine, d));
                bq = mine;
            }
            if (lane == src) best = bq;
        }
        if (valid) nn2[r] = best;
    }
    fb = __reduce_add_sync(0xFFFFFFFFu, fb);
    if (lane == 0 && fb) atomicAdd(n_fallback, (unsigned long long)fb);
}

struct VoxelMeans {
    GridMap gm;
    DBuf<double> mean;         // [rows][3]
    DBuf<float4> loc;          // [rows] mean relative to its brick origin, float32
    DBuf<int32_t> row_brick;   // [rows] brick of every voxel mean
    DBuf<uint16_t> row_pos;
};

static void voxel_means(Ctx& cx, const float* d_P, int64_t n, const double mn[3], double voxel, VoxelMeans& vm) {
    cudaStream_t st = cx.st;
    { Trace tr(cx, " cd grid_build"); grid_build(cx, d_P, n, mn, voxel, vm.gm); }
    Trace tr2(cx, " cd accumulate");
    SQSM_CHECK(voxel <= 0.5, "Chamfer monitor: voxel_size above 0.5 m is not supported (fixed-point voxel sums)");
    int64_t rows = vm.gm.map.n_rows;
    DBuf<long long> sum(rows * 3, st);
    sum.zero();
    DBuf<int32_t> cnt(rows, st);
    cnt.zero();
    k_voxel_accumulate<<<div_up(n, 256), 256, 0, st>>>(d_P, vm.gm.pt_row.p, vm.gm.ent_brick.p, vm.gm.ent_pos.p, vm.gm.map.bricks.p,
                                                       vm.gm.g, n, sum.p, cnt.p);
    vm.gm.ent_brick.release();
    vm.gm.ent_pos.release();
    vm.row_brick.alloc(rows, st);
    vm.row_pos.alloc(rows, st);
    brickmap_expand_rows(cx, vm.gm.map, vm.row_brick.p, vm.row_pos.p);
    vm.mean.alloc(rows * 3, st);
    vm.loc.alloc(rows, st);
    k_voxel_mean<<<div_up(rows, 256), 256, 0, st>>>(sum.p, cnt.p, vm.row_brick.p, vm.row_pos.p, vm.gm.map.bricks.p, vm.gm.g, rows,
                                                    vm.mean.p, vm.loc.p);
    cx.launches += 2;
}

// Partial Chamfer sums: the 1-NN squared distances of the voxel means of P to those of Q (sums[0]) and of Q to P
// (sums[1]) restricted to this shard's contiguous range of query rows; counts = number of voxel means of P and Q.
// With n_shards = 1 this is the whole monitor; with several ranks the sums are all-reduced by the caller.
void chamfer_partial(Ctx& cx, const float* d_P, int64_t np, const float* d_Q, int64_t nq, double voxel, int shard,
                     int n_shards, double sums[2], int64_t counts[2]) {
    cudaStream_t st = cx.st;
    SQSM_CHECK(np > 0 && nq > 0, "Chamfer monitor needs two non-empty clouds");
    SQSM_CHECK(n_shards >= 1 && shard >= 0 && shard < n_shards, "bad shard");
    double mnP[3], mxP[3], mnQ[3], mxQ[3], mn[3], mx[3];
    cloud_bounds(cx, d_P, np, mnP, mxP);
    cloud_bounds(cx, d_Q, nq, mnQ, mxQ);
    for (int a = 0; a < 3; ++a) { mn[a] = std::min(mnP[a], mnQ[a]); mx[a] = std::max(mxP[a], mxQ[a]); }   // :116-117
    VoxelMeans A, B;
    {
        ProfScope ps(cx, "cd_voxel_means", (double)(np + nq) * (12.0 + 2 * 20.0), 0.0);
        voxel_means(cx, d_P, np, mn, voxel, A);                                                            // :120
        voxel_means(cx, d_Q, nq, mn, voxel, B);                                                            // :121
    }
    Trace trnn(cx, " cd nn");
    ProfScope ps_nn(cx, "cd_nn", (double)(A.gm.map.n_rows + B.gm.map.n_rows) * 24.0 * 2, 0.0);
    double ext = 0;
    for (int a = 0; a < 3; ++a) ext = std::max(ext, mx[a] - mn[a]);
    const double edge = 8.0 * voxel;
    int max_ring = (int)std::ceil(ext / edge) + 2;
    DBuf<unsigned long long> nfb(2, st);
    nfb.zero();
    const int nA = A.gm.map.n_bricks, nB = B.gm.map.n_bricks;
    DBuf<int2> abn((int64_t)nA * 32, st), ban((int64_t)nB * 32, st);
    DBuf<int32_t> abid((int64_t)nA * 32, st), baid((int64_t)nB * 32, st);
    // a rank's share of the queries: the rows [rows * shard / n_shards, rows * (shard + 1) / n_shards) of each cloud
    const int64_t rowsA = A.gm.map.n_rows, rowsB = B.gm.map.n_rows;
    const int64_t a_lo = rowsA * shard / n_shards, a_hi = rowsA * (shard + 1) / n_shards;
    const int64_t b_lo = rowsB * shard / n_shards, b_hi = rowsB * (shard + 1) / n_shards;
    k_cross_neighbours<<<div_up((int64_t)nA * 32, 256), 256, 0, st>>>(A.gm.map.bricks.p, nA, B.gm.map.view(),
                                                                      B.gm.map.bricks.p, abn.p, abid.p, a_lo, a_hi);
    k_cross_neighbours<<<div_up((int64_t)nB * 32, 256), 256, 0, st>>>(B.gm.map.bricks.p, nB, A.gm.map.view(),
                                                                      A.gm.map.bricks.p, ban.p, baid.p, b_lo, b_hi);
    cx.launches += 2;
    // optional cell-ring stages (SQSM_CD_RINGS, e.g. "2,3": cubes of 5^3 then 7^3 voxel cells) in front of the brick-level
    // search, which then only sees what is still unsettled.  Default "0" = brick-level search only: on the C3 tree the ring
    // stages settle 66 % / 41 % of the queries (P->Q / Q->P: the contracted cloud sits ~5 cm inside the old surface) but cost
    // what they save (cd_nn 124 ms -> 125 ms with "2", 144 ms with "2,3"; profiles/r02_chamfer_notes.md)
    static const std::vector<int> rings = [] {
        std::vector<int> r;
        const char* e = getenv("SQSM_CD_RINGS");
        std::string t = e ? e : "0";
        for (size_t i = 0; i < t.size(); ++i)
            if (t[i] >= '1' && t[i] <= '4' && r.size() < 4) r.push_back(t[i] - '0');
        return r;
    }();
    int64_t n_left[2][4] = {{0, 0, 0, 0}, {0, 0, 0, 0}};
    auto direction = [&](VoxelMeans& X, VoxelMeans& Y, const int2* xyn, const int32_t* xyid, int which, double* h_sum) {
        const int64_t rows = X.gm.map.n_rows;
        const int64_t row_lo = which == 0 ? a_lo : b_lo, row_hi = which == 0 ? a_hi : b_hi;
        DBuf<double> nn2(rows, st);
        nn2.zero();                                              // rows of other shards contribute nothing
        DBuf<int32_t> lst[2];
        DBuf<int32_t> cnt(1, st);
        const int32_t* cur = nullptr;                            // nullptr = all rows
        int64_t n_cur = rows;
        int flip = 0;
        for (size_t s = 0; s < rings.size() && n_cur > 0; ++s) {
            lst[flip].alloc(n_cur, st);
            cnt.zero();
            const int grid = div_up(n_cur, 256);
#define SQSM_RING(R_)                                                                                                       \
            k_nn_cells<R_><<<grid, 256, 0, st>>>(X.gm.map.bricks.p, X.row_brick.p, X.row_pos.p, cur, n_cur, X.mean.p, xyid,   \
                                                 Y.gm.map.bricks.p, Y.mean.p, X.gm.g, row_lo, row_hi, nn2.p, lst[flip].p, cnt.p)
            switch (rings[s]) {
                case 1: SQSM_RING(1); break;
                case 2: SQSM_RING(2); break;
                case 3: SQSM_RING(3); break;
                default: SQSM_RING(4); break;
            }
#undef SQSM_RING
            cx.launches++;
            int32_t h = 0;
            SQSM_CUDA(cudaMemcpyAsync(&h, cnt.p, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
            SQSM_CUDA(cudaStreamSynchronize(st));
            cur = lst[flip].p;
            n_cur = h;
            n_left[which][s] = h;
            flip ^= 1;
        }
        if (n_cur > 0) {
            const int64_t span = cur ? n_cur : std::max<int64_t>(row_hi - row_lo, 1) + 32;
            const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(div_up(span, 256), kNumSMs * 16));
            k_nn_sqdist_pairs<<<grid, 256, 0, st>>>(X.gm.map.bricks.p, X.row_brick.p, cur, n_cur, X.mean.p, X.loc.p, xyn,
                                                    Y.gm.map.view(), Y.gm.map.bricks.p, Y.mean.p, Y.loc.p, X.gm.g, max_ring, row_lo,
                                                    row_hi, nn2.p, nfb.p + which);
            cx.launches++;
        }
        // fixed-shape reduction over the rows in row order: the sum is bit-reproducible
        DBuf<double> d_sum(1, st);
        size_t tb = 0;
        SQSM_CUDA(cub::DeviceReduce::Sum(nullptr, tb, nn2.p, d_sum.p, (int)rows, st));
        DBuf<uint8_t> tmp((int64_t)tb + 16, st);
        SQSM_CUDA(cub::DeviceReduce::Sum(tmp.p, tb, nn2.p, d_sum.p, (int)rows, st));
        cx.launches += 2;
        SQSM_CUDA(cudaMemcpyAsync(h_sum, d_sum.p, sizeof(double), cudaMemcpyDeviceToHost, st));
        SQSM_CUDA(cudaStreamSynchronize(st));
    };
    direction(A, B, abn.p, abid.p, 0, &sums[0]);
    direction(B, A, ban.p, baid.p, 1, &sums[1]);
    if (Trace::enabled()) {
        auto f = nfb.to_host();
        fprintf(stderr, "[trace]  cd nn: unsettled after the cell rings %lld/%lld of %lld, %lld/%lld of %lld; brick-ring fallbacks %llu, %llu\n",
                (long long)n_left[0][0], (long long)n_left[0][1], (long long)A.gm.map.n_rows, (long long)n_left[1][0],
                (long long)n_left[1][1], (long long)B.gm.map.n_rows, f[0], f[1]);
    }
    counts[0] = A.gm.map.n_rows;
    counts[1] = B.gm.map.n_rows;
}

double chamfer_distance(Ctx& cx, const float* d_P, int64_t np, const float* d_Q, int64_t nq, double voxel) {
    double sums[2];
    int64_t counts[2];
    chamfer_partial(cx, d_P, np, d_Q, nq, voxel, 0, 1, sums, counts);
    return sums[0] / (double)counts[0] + sums[1] / (double)counts[1];                                      // :122
}

// ---------------------------------------------------------------------------------- skeletal sampling

__global__ void __launch_bounds__(256) k_min_index(const int32_t* __restrict__ pt_row, int64_t n, uint32_t* __restrict__ winner) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int r = pt_row[i];
    if (r >= 0) atomicMin(&winner[r], (uint32_t)i);
}

__global__ void __launch_bounds__(256) k_is_first(const int32_t* __restrict__ pt_row, const uint32_t* __restrict__ winner,
                                                  int64_t n, int32_t* __restrict__ flag) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i > n) return;
    int f = 0;
    if (i < n) { int r = pt_row[i]; f = (r >= 0 && winner[r] == (uint32_t)i) ? 1 : 0; }
    flag[i] = f;
}

// skeletal_points = contracted[ids]; radii = ||displacement[ids]||_2 computed like np.linalg.norm on float32:
// sqrt(add.reduce(x*x)) with float32 multiplies and adds in index order, no FMA.
__global__ void __launch_bounds__(256) k_skeletal_gather(const float* __restrict__ P, const float* __restrict__ D,
                                                         const int32_t* __restrict__ flag, const int32_t* __restrict__ scan,
                                                         int64_t n, int64_t* __restrict__ ids, float* __restrict__ pts,
                                                         float* __restrict__ radii) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || !flag[i]) return;
    int o = scan[i];
    ids[o] = i;
    pts[3 * (size_t)o + 0] = P[3 * i + 0]; pts[3 * (size_t)o + 1] = P[3 * i + 1]; pts[3 * (size_t)o + 2] = P[3 * i + 2];
    float dx = D[3 * i + 0], dy = D[3 * i + 1], dz = D[3 * i + 2];
    float s = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
    radii[o] = __fsqrt_rn(s);
}

void skeletal_sample(Ctx& cx, const float* d_P, const float* d_D, int64_t n, double voxel, SkeletalResult& out) {
    cudaStream_t st = cx.st;
    SQSM_CHECK(n > 0, "no contracted points");
    double mn[3], mx[3];
    cloud_bounds(cx, d_P, n, mn, mx);                                 // cloud.get_min_bound() (thinning...py:76)
    GridMap gm;
    grid_build(cx, d_P, n, mn, voxel, gm);
    int64_t rows = gm.map.n_rows;
    DBuf<uint32_t> winner(rows, st);
    winner.fill_ff();
    k_min_index<<<div_up(n, 256), 256, 0, st>>>(gm.pt_row.p, n, winner.p);
    DBuf<int32_t> flag(n + 1, st), scan(n + 1, st);
    k_is_first<<<div_up(n + 1, 256), 256, 0, st>>>(gm.pt_row.p, winner.p, n, flag.p);
    cx.launches += 2;
    exclusive_scan_i32(cx, flag.p, scan.p, n + 1);
    out.m = rows;
    out.ids.alloc(rows, st);
    out.pts.alloc(rows * 3, st);
    out.radii.alloc(rows, st);
    k_skeletal_gather<<<div_up(n, 256), 256, 0, st>>>(d_P, d_D, flag.p, scan.p, n, out.ids.p, out.pts.p, out.radii.p);
    cx.launches++;
    int32_t tot = 0;
    SQSM_CUDA(cudaMemcpyAsync(&tot, scan.p + n, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    SQSM_CUDA(cudaStreamSynchronize(st));
    SQSM_CHECK(tot == rows, "internal error: skeletal sample count != voxel count");
}

// ---------------------------------------------------------------------------------- radius graph

__global__ void __launch_bounds__(256) k_cell_entries(const float* __restrict__ P, int64_t n, const __grid_constant__ GridParams g,
                                                      uint64_t* __restrict__ ent_key, uint16_t* __restrict__ ent_pos,
                                                      int32_t* __restrict__ err) {
    // cells of edge g.voxel: one brick per cell (pos unused)
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int v[3];
    bool ok = true;
    #pragma unroll
    for (int a = 0; a < 3; ++a) {
        double f = floor(((double)P[3 * i + a] - g.mn[a]) / g.voxel);
        ok = ok && (f >= 0.0) && (f < 65536.0);
        v[a] = (int)f;
    }
    if (!ok) { ent_key[i] = kEmptyKey; ent_pos[i] = 0; atomicOr(err, 1); return; }
    ent_key[i] = pack_key(0u, (uint32_t)v[2], (uint32_t)v[1], (uint32_t)v[0]);
    ent_pos[i] = 0;
}

__global__ void __launch_bounds__(256) k_cell_count(const int32_t* __restrict__ ent_brick, int64_t n, int32_t* __restrict__ cnt) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && ent_brick[i] >= 0) atomicAdd(&cnt[ent_brick[i]], 1);
}

// mode 0: count neighbours j > i within r; mode 1: write those of the nodes i in [i_lo, i_hi) at off[i] - off[i_lo]
// (the edge list of a big cloud is produced in node-range chunks of < 2^31 edges, see radius_graph)
template <int MODE>
__global__ void __launch_bounds__(256) k_radius_pairs(const float* __restrict__ P, const uint32_t* __restrict__ order,
                                                      const uint32_t* __restrict__ cell_of, int64_t n,
                                                      const int32_t* __restrict__ bnbr, const int32_t* __restrict__ cstart,
                                                      double r2, int32_t* __restrict__ cnt, const int64_t* __restrict__ off,
                                                      uint32_t i_lo, uint32_t i_hi, uint32_t* __restrict__ ebuf_i,
                                                      uint32_t* __restrict__ ebuf_j) {
    int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    const uint32_t i = order[s];
    if (i < i_lo || i >= i_hi) {                         // MODE 0: node range of this call (a rank's share), MODE 1: chunk
        if (MODE == 0) cnt[i] = 0;
        return;
    }
    const int cell = (int)cell_of[s];
    const double x = (double)P[3 * (size_t)i], y = (double)P[3 * (size_t)i + 1], z = (double)P[3 * (size_t)i + 2];
    int c = 0;
    int64_t o = (MODE == 1) ? off[i] - off[i_lo] : 0;
    for (int nb = 0; nb < 27; ++nb) {
        int id = bnbr[(size_t)cell * 32 + nb];
        if (id < 0) continue;
        int b = cstart[id], e = cstart[id + 1];
        for (int t = b; t < e; ++t) {
            uint32_t j = order[t];
            if (j <= i) continue;
            double dx = x - (double)P[3 * (size_t)j], dy = y - (double)P[3 * (size_t)j + 1], dz = z - (double)P[3 * (size_t)j + 2];
            double d2 = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
            if (d2 <= r2) {
                if (MODE == 1) { ebuf_i[o + c] = i; ebuf_j[o + c] = j; }
                ++c;
            }
        }
    }
    if (MODE == 0) cnt[i] = c;
}

// first node whose edge offset reaches k * budget, k = 0 .. n_chunks (chunk boundaries of the edge list)
__global__ void k_chunk_bounds(const int64_t* __restrict__ off, int64_t m, int64_t budget, int n_bounds, int64_t* __restrict__ bounds) {
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_bounds) return;
    const int64_t target = (int64_t)k * budget;
    int64_t lo = 0, hi = m;                                   // first i in [0, m] with off[i] >= target
    while (lo < hi) { int64_t mid = (lo + hi) / 2; if (off[mid] < target) lo = mid + 1; else hi = mid; }
    bounds[k] = lo;
}

// (i, sorted j) 32-bit keys -> the int64 [E][2] edge list of the ABI; fully coalesced
__global__ void __launch_bounds__(256) k_edges_expand(const uint32_t* __restrict__ ei, const uint32_t* __restrict__ ej, int64_t n,
                                                      longlong2* __restrict__ edges) {
    int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e < n) edges[e] = make_longlong2((long long)ei[e], (long long)ej[e]);
}

__global__ void k_gather_i64(const int64_t* __restrict__ src, const int64_t* __restrict__ idx, int n, int64_t* __restrict__ dst) {
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n) dst[k] = src[idx[k]];
}

__global__ void __launch_bounds__(256) k_iota_u32(uint32_t* __restrict__ a, int64_t n) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) a[i] = (uint32_t)i;
}

__global__ void __launch_bounds__(256) k_i32_to_i64(const int32_t* __restrict__ a, int64_t* __restrict__ b, int64_t n) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) b[i] = (int64_t)a[i];
}

// node_lo / node_hi (default: all nodes): only the edges (i, j) with node_lo <= i < node_hi are produced - the list of the
// whole cloud is the concatenation of the lists of consecutive node ranges, so ranks can share the work of one cloud
void radius_graph(Ctx& cx, const float* d_pts, int64_t m, double r, DBuf<int64_t>& edges, int64_t* n_edges, int64_t node_lo,
                  int64_t node_hi) {
    if (node_hi < 0 || node_hi > m) node_hi = m;
    node_lo = std::max<int64_t>(0, std::min(node_lo, node_hi));
    cudaStream_t st = cx.st;
    *n_edges = 0;
    if (m <= 1) return;
    SQSM_CHECK(r > 0 && m < (1ll << 31), "radius graph: bad arguments (at most 2^31 - 1 nodes)");
    double mn[3], mx[3];
    cloud_bounds(cx, d_pts, m, mn, mx);
    GridParams g;
    for (int a = 0; a < 3; ++a) g.mn[a] = mn[a];
    g.voxel = r * 1.0009765625;                         // cell edge slightly above r: 27 cells cover the ball
    // points whose cell coordinate would overflow 16 bits: widen the cells
    double ext = 0;
    for (int a = 0; a < 3; ++a) ext = std::max(ext, mx[a] - mn[a]);
    if (ext / g.voxel > 65000.0) g.voxel = ext / 65000.0;
    DBuf<uint64_t> ent_key(m, st);
    DBuf<uint16_t> ent_pos(m, st);
    DBuf<int32_t> ent_brick(m, st);
    DBuf<int32_t> err(1, st);
    err.zero();
    k_cell_entries<<<div_up(m, 256), 256, 0, st>>>(d_pts, m, g, ent_key.p, ent_pos.p, err.p);
    cx.launches++;
    BrickMap map;
    brickmap_build(cx, map, ent_key.p, ent_pos.p, m, ent_brick.p);
    int32_t herr = 0;
    SQSM_CUDA(cudaMemcpyAsync(&herr, err.p, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    SQSM_CUDA(cudaStreamSynchronize(st));
    SQSM_CHECK(herr == 0, "radius graph: non-finite coordinates");
    brickmap_build_neighbours(cx, map);
    const int nc = map.n_bricks;
    // cell lists: points sorted by cell id (stable -> ascending point index inside a cell)
    DBuf<int32_t> ccnt((int64_t)nc + 1, st), cstart((int64_t)nc + 1, st);
    ccnt.zero();
    k_cell_count<<<div_up(m, 256), 256, 0, st>>>(ent_brick.p, m, ccnt.p);
    exclusive_scan_i32(cx, ccnt.p, cstart.p, (int64_t)nc + 1);
    DBuf<uint32_t> iota(m, st), order(m, st), cell_sorted(m, st);
    k_iota_u32<<<div_up(m, 256), 256, 0, st>>>(iota.p, m);
    cx.launches += 2;
    int bits = 1;
    while ((1ll << bits) < nc) ++bits;
    sort_pairs_u32(cx, (const uint32_t*)ent_brick.p, cell_sorted.p, iota.p, order.p, m, 0, bits);
    // count, scan, fill
    DBuf<int32_t> cnt(m + 1, st);
    SQSM_CUDA(cudaMemsetAsync(cnt.p + m, 0, sizeof(int32_t), st));
    const double r2 = r * r;
    k_radius_pairs<0><<<div_up(m, 256), 256, 0, st>>>(d_pts, order.p, cell_sorted.p, m, map.bnbr.p, cstart.p, r2, cnt.p,
                                                      nullptr, (uint32_t)node_lo, (uint32_t)node_hi, nullptr, nullptr);
    DBuf<int64_t> cnt64(m + 1, st), off(m + 1, st);
    k_i32_to_i64<<<div_up(m + 1, 256), 256, 0, st>>>(cnt.p, cnt64.p, m + 1);
    cx.launches += 2;
    exclusive_scan_i64(cx, cnt64.p, off.p, m + 1);
    int64_t E = 0;
    SQSM_CUDA(cudaMemcpyAsync(&E, off.p + m, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    SQSM_CUDA(cudaStreamSynchronize(st));
    *n_edges = E;
    if (E == 0) return;
    cnt.release();
    cnt64.release();
    // fill (i, j) as 32-bit keys, sort j ascending inside every node's segment (the merged edge list is lexicographic,
    // :158) with a segmented sort - no quadratic worst case for dense clusters - then expand to the int64 pairs of the ABI.
    // The list is produced in chunks of whole nodes with at most `budget` edges each (64-bit offsets throughout), so
    // neither the 32-bit item count of the segmented sort nor the scratch buffers limit E: a 200 M-point scan has
    // ~2.5 G edges.  A node with more than `budget` neighbours gets a chunk of its own.
    const char* env_budget = getenv("SQSM_EDGE_CHUNK");
    const int64_t budget = env_budget ? atoll(env_budget) : (1ll << 30);
    SQSM_CHECK(budget > 0, "SQSM_EDGE_CHUNK must be positive");
    edges.alloc(2 * E, st);
    const int n_bounds = (int)(E / budget) + 2;
    std::vector<int64_t> bounds(n_bounds);
    {
        DBuf<int64_t> d_bounds(n_bounds, st);
        k_chunk_bounds<<<div_up(n_bounds, 128), 128, 0, st>>>(off.p, m, budget, n_bounds, d_bounds.p);
        cx.launches++;
        SQSM_CUDA(cudaMemcpyAsync(bounds.data(), d_bounds.p, sizeof(int64_t) * n_bounds, cudaMemcpyDeviceToHost, st));
        SQSM_CUDA(cudaStreamSynchronize(st));
    }
    bounds[n_bounds - 1] = m;
    std::vector<int64_t> cuts;                                   // strictly increasing node boundaries 0 .. m
    cuts.push_back(0);
    for (int k = 1; k < n_bounds; ++k) {
        // off[bounds[k]] >= k * budget: stepping one node back keeps the chunk inside the budget (unless it is a single node)
        int64_t c = std::min<int64_t>(bounds[k], m);
        if (k < n_bounds - 1 && c - 1 > cuts.back()) c -= 1;
        if (c > cuts.back()) cuts.push_back(c);
    }
    if (cuts.back() != m) cuts.push_back(m);
    std::vector<int64_t> cut_off(cuts.size());
    {
        DBuf<int64_t> d_cuts((int64_t)cuts.size(), st), d_co((int64_t)cuts.size(), st);
        SQSM_CUDA(cudaMemcpyAsync(d_cuts.p, cuts.data(), sizeof(int64_t) * cuts.size(), cudaMemcpyHostToDevice, st));
        k_gather_i64<<<div_up((int64_t)cuts.size(), 128), 128, 0, st>>>(off.p, d_cuts.p, (int)cuts.size(), d_co.p);
        cx.launches++;
        SQSM_CUDA(cudaMemcpyAsync(cut_off.data(), d_co.p, sizeof(int64_t) * cuts.size(), cudaMemcpyDeviceToHost, st));
        SQSM_CUDA(cudaStreamSynchronize(st));
    }
    for (size_t c = 0; c + 1 < cuts.size(); ++c) {
        const int64_t i0 = cuts[c], i1 = cuts[c + 1], e0 = cut_off[c], ne = cut_off[c + 1] - cut_off[c];
        if (ne == 0) continue;
        SQSM_CHECK(ne < (1ll << 31), "radius graph: one node range holds more than 2^31 edges (lower SQSM_EDGE_CHUNK)");
        DBuf<uint32_t> ebuf_i(ne, st), ebuf_j(ne, st), ebuf_s(ne, st);
        k_radius_pairs<1><<<div_up(m, 256), 256, 0, st>>>(d_pts, order.p, cell_sorted.p, m, map.bnbr.p, cstart.p, r2, nullptr,
                                                          off.p, (uint32_t)std::max(i0, node_lo), (uint32_t)std::min(i1, node_hi),
                                                          ebuf_i.p, ebuf_j.p);   // (nodes outside the range own no edges)
        segmented_sort_u32(cx, ebuf_j.p, ebuf_s.p, ne, i1 - i0, off.p + i0, e0);
        ebuf_j.release();
        k_edges_expand<<<div_up(ne, 256), 256, 0, st>>>(ebuf_i.p, ebuf_s.p, ne, reinterpret_cast<longlong2*>(edges.p) + e0);
        cx.launches += 2;
    }
}

// ---------------------------------------------------------------------------------- variable-radius ball lists (row f3)
//
// core/refinement.py:448-456 occupies the skeleton nodes around a path with
//   kdtree.query_ball_point(points[path], r = reference_radii[path] * factor)  ->  np.unique(np.concatenate(...))
// inside a sequential loop over the paths.  The radii of ALL nodes are known before that loop, so the device computes the
// ball of every node once - a radius graph with a per-node radius - and the loop only unites precomputed lists.
// Semantics of cKDTree.query_ball_point (p = 2): j is in the ball of i when the float64 squared distance is <= r_i * r_i;
// i itself is always a member (a negative radius acts like its magnitude, NaN gives an empty list).  Ascending node index.
//
// Radii span three orders of magnitude (twigs 2 mm, trunk 0.4 m), so ONE cell size does not work: with cells of the mean
// radius a trunk node would probe 10^6 cells.  The points are binned on up to 8 grids whose cell edge grows by 8x per level,
// the finest at the mean radius; a query uses the finest grid whose cell edge is at least an eighth of its radius, i.e. at
// most 19^3 cells (27 or 125 for the bulk of the nodes; finer grids for everybody - a quarter of the radius, 11^3 cells - were
// measured 7x slower on 400 k nodes: the hash probes dominate).  A WARP serves one query: one hash probe per lane and
// cell, the points of the 32 cells of a batch are dealt evenly to the lanes.  Count pass, scan, fill pass (positions inside
// the segment from warp ballots: no atomics, a fixed order), segmented sort of every segment.
constexpr int kBallLevels = 8;
struct BallLevel { double voxel; BrickHash h; const int32_t* cstart; const uint32_t* order; int k_max; };
struct BallGrids { BallLevel lv[kBallLevels]; int n; double mn[3]; };

template <int MODE>
__global__ void __launch_bounds__(256) k_ball_lists(const float* __restrict__ P, int64_t m, const double* __restrict__ radii,
                                                    const __grid_constant__ BallGrids G, int32_t* __restrict__ cnt,
                                                    const int64_t* __restrict__ off, uint32_t* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (i >= m) return;
    const double x = (double)P[3 * i], y = (double)P[3 * i + 1], z = (double)P[3 * i + 2];
    // cKDTree compares squared distances with r^2: a negative radius acts like its magnitude, NaN gives an empty list
    const double r = fabs(radii[i]);
    const double r2 = r * r;
    if (!(r >= 0.0)) {
        if (MODE == 0 && lane == 0) cnt[i] = 0;
        return;
    }
    int lvl = 0;
    while (lvl + 1 < G.n && 8.0 * G.lv[lvl].voxel < r) ++lvl;
    const BallLevel& L = G.lv[lvl];
    const int cx0 = (int)floor((x - G.mn[0]) / L.voxel), cy0 = (int)floor((y - G.mn[1]) / L.voxel), cz0 = (int)floor((z - G.mn[2]) / L.voxel);
    const int k = (int)fmin(floor(r / L.voxel) + 1.0, (double)L.k_max);     // k_max rings cover the whole grid (r = inf)
    const int side = 2 * k + 1;
    const int64_t n_cells = (int64_t)side * side * side;
    int total = 0;
    const int64_t base = MODE == 1 ? off[i] : 0;
    for (int64_t t0 = 0; t0 < n_cells; t0 += 32) {
        const int64_t t = t0 + lane;
        int b = 0, e = 0;
        if (t < n_cells) {
            const int dz = (int)(t / ((int64_t)side * side)) - k, dy = (int)((t / side) % side) - k, dx = (int)(t % side) - k;
            const int cz = cz0 + dz, cy = cy0 + dy, cx = cx0 + dx;
            if (cz >= 0 && cy >= 0 && cx >= 0 && cz < 65536 && cy < 65536 && cx < 65536) {
                const int id = hash_find(L.h, pack_key(0u, (uint32_t)cz, (uint32_t)cy, (uint32_t)cx));
                if (id >= 0) { b = L.cstart[id]; e = L.cstart[id + 1]; }
            }
        }
        // the candidates of the 32 cells of this batch are dealt evenly to the lanes (a lane walking its own cell would
        // make the warp wait for the fullest cell): inclusive prefix of the cell sizes, then candidate c belongs to the
        // first cell whose prefix exceeds c
        int incl = e - b;
        #pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int v = __shfl_up_sync(0xFFFFFFFFu, incl, d);
            if (lane >= d) incl += v;
        }
        const int n_cand = __shfl_sync(0xFFFFFFFFu, incl, 31);
        for (int c0 = 0; c0 < n_cand; c0 += 32) {
            const int c = c0 + lane;
            int cell = 0;
            #pragma unroll
            for (int st = 16; st > 0; st >>= 1) {
                const int v = __shfl_sync(0xFFFFFFFFu, incl, cell + st - 1);
                if (v <= c) cell += st;
            }
            cell = min(cell, 31);
            const int cb = __shfl_sync(0xFFFFFFFFu, b, cell), ci = __shfl_sync(0xFFFFFFFFu, incl, cell), ce = __shfl_sync(0xFFFFFFFFu, e, cell);
            bool hit = false;
            uint32_t j = 0;
            if (c < n_cand) {
                j = L.order[cb + (c - (ci - (ce - cb)))];
                const double dx = x - (double)P[3 * (size_t)j], dy = y - (double)P[3 * (size_t)j + 1], dz = z - (double)P[3 * (size_t)j + 2];
                const double d2 = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
                hit = d2 <= r2;
            }
            const unsigned w = __ballot_sync(0xFFFFFFFFu, hit);
            if (MODE == 1 && hit) out[base + total + __popc(w & ((1u << lane) - 1u))] = j;
            total += __popc(w);
        }
    }
    if (MODE == 0 && lane == 0) cnt[i] = total;
}

// finest_cell / largest_radius: cell edge of the finest grid and the largest finite radius (from the caller's host copy)
void ball_lists(Ctx& cx, const float* d_pts, int64_t m, const double* d_radii, double finest_cell, double largest_radius,
                DBuf<int64_t>& offsets, DBuf<uint32_t>& indices, int64_t* n_total) {
    cudaStream_t st = cx.st;
    *n_total = 0;
    offsets.alloc(m + 1, st);
    offsets.zero();
    if (m <= 0) return;
    SQSM_CHECK(m < (1ll << 31), "ball lists: at most 2^31 - 1 nodes");
    double mn[3], mx[3];
    cloud_bounds(cx, d_pts, m, mn, mx);
    double ext = 0;
    for (int a = 0; a < 3; ++a) ext = std::max(ext, mx[a] - mn[a]);
    struct LevelData { BrickMap map; DBuf<int32_t> cstart; DBuf<uint32_t> order; };
    std::vector<std::unique_ptr<LevelData>> levels;
    BallGrids G;
    std::memset(&G, 0, sizeof(G));
    for (int a = 0; a < 3; ++a) G.mn[a] = mn[a];
    double cell = std::max(std::max(finest_cell, 1e-6), ext / 65000.0);
    for (int l = 0; l < kBallLevels; ++l) {
        Trace trl(cx, " ball level");
        const bool last = l == kBallLevels - 1 || 8.0 * cell >= largest_radius;
        if (last) cell = std::max(cell, 0.125 * largest_radius);
        GridParams g;
        for (int a = 0; a < 3; ++a) g.mn[a] = mn[a];
        g.voxel = cell;
        auto L = std::make_unique<LevelData>();
        DBuf<uint64_t> ent_key(m, st);
        DBuf<uint16_t> ent_pos(m, st);
        DBuf<int32_t> ent_brick(m, st);
        DBuf<int32_t> err(1, st);
        err.zero();
        k_cell_entries<<<div_up(m, 256), 256, 0, st>>>(d_pts, m, g, ent_key.p, ent_pos.p, err.p);
        cx.launches++;
        brickmap_build(cx, L->map, ent_key.p, ent_pos.p, m, ent_brick.p, m);     // fine cells: up to one cell per point
        int32_t herr = 0;
        SQSM_CUDA(cudaMemcpyAsync(&herr, err.p, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
        SQSM_CUDA(cudaStreamSynchronize(st));
        SQSM_CHECK(herr == 0, "ball lists: non-finite coordinates");
        const int nc = L->map.n_bricks;
        DBuf<int32_t> ccnt((int64_t)nc + 1, st);
        L->cstart.alloc((int64_t)nc + 1, st);
        ccnt.zero();
        k_cell_count<<<div_up(m, 256), 256, 0, st>>>(ent_brick.p, m, ccnt.p);
        exclusive_scan_i32(cx, ccnt.p, L->cstart.p, (int64_t)nc + 1);
        DBuf<uint32_t> iota(m, st), cell_sorted(m, st);
        L->order.alloc(m, st);
        k_iota_u32<<<div_up(m, 256), 256, 0, st>>>(iota.p, m);
        cx.launches += 2;
        int bits = 1;
        while ((1ll << bits) < nc) ++bits;
        sort_pairs_u32(cx, (const uint32_t*)ent_brick.p, cell_sorted.p, iota.p, L->order.p, m, 0, bits);
        G.lv[l].voxel = cell;
        G.lv[l].h = L->map.view();
        G.lv[l].cstart = L->cstart.p;
        G.lv[l].order = L->order.p;
        G.lv[l].k_max = (int)std::min(std::ceil(ext / cell) + 1.0, 65536.0);
        G.n = l + 1;
        levels.push_back(std::move(L));
        if (last) break;
        cell *= 8.0;
    }
    DBuf<int32_t> cnt(m + 1, st);
    SQSM_CUDA(cudaMemsetAsync(cnt.p + m, 0, sizeof(int32_t), st));
    const int grid = div_up(m * 32, 256);
    { Trace tr(cx, " ball count"); k_ball_lists<0><<<grid, 256, 0, st>>>(d_pts, m, d_radii, G, cnt.p, nullptr, nullptr); }
    DBuf<int64_t> cnt64(m + 1, st);
    k_i32_to_i64<<<div_up(m + 1, 256), 256, 0, st>>>(cnt.p, cnt64.p, m + 1);
    cx.launches += 2;
    exclusive_scan_i64(cx, cnt64.p, offsets.p, m + 1);
    int64_t T = 0;
    SQSM_CUDA(cudaMemcpyAsync(&T, offsets.p + m, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    SQSM_CUDA(cudaStreamSynchronize(st));
    *n_total = T;
    if (T == 0) return;
    SQSM_CHECK(T < (1ll << 31), "ball lists: more than 2^31 - 1 list entries (split the nodes over several calls)");
    DBuf<uint32_t> raw(T, st);
    { Trace tr(cx, " ball fill"); k_ball_lists<1><<<grid, 256, 0, st>>>(d_pts, m, d_radii, G, nullptr, offsets.p, raw.p); }
    cx.launches++;
    indices.alloc(T, st);
    { Trace tr(cx, " ball sort"); segmented_sort_u32(cx, raw.p, indices.p, T, m, offsets.p, 0); }
    SQSM_CUDA(cudaStreamSynchronize(st));               // the level tables go out of scope
}

}  // namespace sqsm
