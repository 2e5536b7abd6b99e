// Cube tiling, per-cube voxelisation and rulebook construction.  sm_100a.
//
// Replaces, on the device and for all cubes at once:
//   SingleTreeDataset.__init__      external/SmartTreeXX/datasets/synthetic_tree.py:243-281
//   SingleTreeDataset.__getitem__   :290-333  (spconv PointToVoxel, CPU first-come semantics)
//   collate_synthetic_tree          :176-208  (batch ids, spatial_shape = max index)
//   spconv SubMConv3d / SparseConv3d(k2,s2) index generation
#include "engine.h"
#include <algorithm>
#include <cmath>

namespace sqsm {

// ---------------------------------------------------------------------------------- float32 rules

// torch.div(a, b, rounding_mode="floor") for float32 (c10::div_floor_floating), synthetic_tree.py:246
__device__ __forceinline__ float div_floor_f32(float a, float b) {
    float mod = fmodf(a, b);
    float div = __fdiv_rn(__fsub_rn(a, mod), b);
    if (mod != 0.0f && ((b < 0.0f) != (mod < 0.0f))) div = __fsub_rn(div, 1.0f);
    float fl;
    if (div != 0.0f) {
        fl = floorf(div);
        if (__fsub_rn(div, fl) > 0.5f) fl = __fadd_rn(fl, 1.0f);
    } else {
        fl = copysignf(0.0f, __fdiv_rn(a, b));
    }
    return fl;
}

// cube bounds exactly as synthetic_tree.py:249,259-262 computes them in float32
__host__ __device__ __forceinline__ void cube_bounds(int c, float cs, float half, float buf, float& lo, float& hi,
                                                     float& slo, float& shi) {
#ifdef __CUDA_ARCH__
    float centre = __fadd_rn(__fmul_rn((float)c, cs), half);
    lo = __fsub_rn(centre, half);
    hi = __fadd_rn(centre, half);
    slo = __fsub_rn(lo, buf);
    shi = __fadd_rn(hi, buf);
#else
    volatile float m = (float)c * cs;
    volatile float centre = m + half;
    volatile float l = centre - half, h = centre + half;
    volatile float sl = l - buf, sh = h + buf;
    lo = l; hi = h; slo = sl; shi = sh;
#endif
}

struct TileParams {
    float cs, half, buf;
    int ox, oy, oz;       // origin of the dense cube -> sample table (cube index of cell 0)
    int dx, dy, dz;       // its extents
    const int32_t* lut;   // sample id or -1
};

constexpr int kCubeOff = 32768;
__host__ __device__ __forceinline__ uint64_t cube_key(int cx, int cy, int cz) {   // sorts as (x, y, z)
    return ((uint64_t)(uint32_t)(cx + kCubeOff) << 32) | ((uint64_t)(uint32_t)(cy + kCubeOff) << 16) |
           (uint64_t)(uint32_t)(cz + kCubeOff);
}

// ---------------------------------------------------------------------------------- tiling kernels

__global__ void __launch_bounds__(256) k_cube_set(const float* __restrict__ P, int64_t n, float cs, BrickHash h,
                                                  uint64_t* __restrict__ list, int list_cap, int32_t* __restrict__ counter,
                                                  int32_t* __restrict__ err) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float x = P[3 * i], y = P[3 * i + 1], z = P[3 * i + 2];
    float fx = div_floor_f32(x, cs), fy = div_floor_f32(y, cs), fz = div_floor_f32(z, cs);
    if (!(fabsf(fx) < 30000.f && fabsf(fy) < 30000.f && fabsf(fz) < 30000.f)) { atomicOr(err, 1); return; }
    uint64_t key = cube_key((int)fx, (int)fy, (int)fz);
    unsigned peers = __match_any_sync(__activemask(), key);
    if ((int)(__ffs(peers) - 1) != (int)(threadIdx.x & 31)) return;
    bool ins;
    uint32_t slot = hash_insert(h, key, &ins);
    if (slot == 0xFFFFFFFFu) { atomicOr(err, 2); return; }
    if (ins) {
        const int k = atomicAdd(counter, 1);
        if (k < list_cap) list[k] = key; else atomicOr(err, 2);      // more distinct cubes than the list holds
    }
}

// Enumerate the cube samples whose halo box contains point p (synthetic_tree.py:263-272).
template <typename F>
__device__ __forceinline__ void for_each_membership(float x, float y, float z, const TileParams& tp, F&& f) {
    float pa[3] = {x, y, z};
    int ci[3];
    bool halo[3][3], inner[3][3];
    #pragma unroll
    for (int a = 0; a < 3; ++a) {
        ci[a] = (int)div_floor_f32(pa[a], tp.cs);
        #pragma unroll
        for (int d = 0; d < 3; ++d) {
            float lo, hi, slo, shi;
            cube_bounds(ci[a] + d - 1, tp.cs, tp.half, tp.buf, lo, hi, slo, shi);
            halo[a][d] = (pa[a] >= slo) && (pa[a] <= shi);
            inner[a][d] = (pa[a] >= lo) && (pa[a] <= hi);
        }
    }
    #pragma unroll
    for (int ix = 0; ix < 3; ++ix) {
        if (!halo[0][ix]) continue;
        int cx = ci[0] + ix - 1 - tp.ox;
        if (cx < 0 || cx >= tp.dx) continue;
        #pragma unroll
        for (int iy = 0; iy < 3; ++iy) {
            if (!halo[1][iy]) continue;
            int cy = ci[1] + iy - 1 - tp.oy;
            if (cy < 0 || cy >= tp.dy) continue;
            #pragma unroll
            for (int iz = 0; iz < 3; ++iz) {
                if (!halo[2][iz]) continue;
                int cz = ci[2] + iz - 1 - tp.oz;
                if (cz < 0 || cz >= tp.dz) continue;
                int s = tp.lut[((size_t)cx * tp.dy + cy) * tp.dz + cz];
                if (s < 0) continue;
                f(s, inner[0][ix] && inner[1][iy] && inner[2][iz]);
            }
        }
    }
}

__global__ void __launch_bounds__(256) k_member_count(const float* __restrict__ P, int64_t n, TileParams tp,
                                                      int32_t* __restrict__ cnt) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int c = 0;
    for_each_membership(P[3 * i], P[3 * i + 1], P[3 * i + 2], tp, [&](int, bool) { ++c; });
    cnt[i] = c;
}

__global__ void __launch_bounds__(256) k_member_emit(const float* __restrict__ P, int64_t n, TileParams tp,
                                                     const int32_t* __restrict__ off, uint32_t* __restrict__ ent_sample,
                                                     uint32_t* __restrict__ ent_val) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int o = off[i];
    for_each_membership(P[3 * i], P[3 * i + 1], P[3 * i + 2], tp, [&](int s, bool inner) {
        ent_sample[o] = (uint32_t)s;
        ent_val[o] = (uint32_t)i | (inner ? 0x80000000u : 0u);
        ++o;
    });
}

__global__ void __launch_bounds__(256) k_boundaries(const uint32_t* __restrict__ key, int64_t n,
                                                    int64_t* __restrict__ start) {
    int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n) return;
    uint32_t s = key[e];
    if (e == 0 || key[e - 1] != s) start[s] = e;
}

void tile_points(Ctx& cx, const float* d_P, int64_t n, float cube_size, float buffer_size, float voxel_size,
                 int batch_size, TilingResult& out) {
    cudaStream_t st = cx.st;
    SQSM_CHECK(n > 0 && n < (1ll << 31), "point count must be in [1, 2^31)");
    SQSM_CHECK(buffer_size < cube_size && buffer_size >= 0.f, "buffer_size must be in [0, cube_size)");
    // 1. the set of occupied cubes
    const uint32_t cap = 1u << 17;
    DBuf<uint64_t> hk(cap, st);
    DBuf<int32_t> hv(cap, st);
    DBuf<uint64_t> list(cap / 2, st);
    DBuf<int32_t> ctr(2, st);
    hk.fill_ff();
    ctr.zero();
    BrickHash h{hk.p, hv.p, cap - 1};
    k_cube_set<<<div_up(n, 256), 256, 0, st>>>(d_P, n, cube_size, h, list.p, (int)(cap / 2), ctr.p, ctr.p + 1);
    cx.launches++;
    int32_t hc[2];
    SQSM_CUDA(cudaMemcpyAsync(hc, ctr.p, sizeof(hc), cudaMemcpyDeviceToHost, st));
    SQSM_CUDA(cudaStreamSynchronize(st));
    SQSM_CHECK((hc[1] & 1) == 0, "cube index out of range (point coordinates / cube_size too large or NaN)");
    SQSM_CHECK((hc[1] & 2) == 0, "unsupported number of cubes (more than 65534 occupied cubes)");
    int nc = hc[0];
    SQSM_CHECK(nc > 0 && nc < 65535 && nc <= (int)cap / 2, "unsupported number of cubes");
    std::vector<uint64_t> keys(nc);
    SQSM_CUDA(cudaMemcpyAsync(keys.data(), list.p, sizeof(uint64_t) * nc, cudaMemcpyDeviceToHost, st));
    SQSM_CUDA(cudaStreamSynchronize(st));
    std::sort(keys.begin(), keys.end());                       // torch.unique(dim=0): lexicographic (x, y, z)
    // 2. sample table + dense cube -> sample lookup
    out.n_samples = nc;
    out.n_batches = (nc + batch_size - 1) / batch_size;
    out.samples.assign(nc, SampleInfo{});
    const float half = cube_size / 2.0f;
    int mn[3] = {1 << 30, 1 << 30, 1 << 30}, mx[3] = {-(1 << 30), -(1 << 30), -(1 << 30)};
    for (int s = 0; s < nc; ++s) {
        int c[3] = {(int)((keys[s] >> 32) & 0xFFFF) - kCubeOff, (int)((keys[s] >> 16) & 0xFFFF) - kCubeOff,
                    (int)(keys[s] & 0xFFFF) - kCubeOff};
        SampleInfo& si = out.samples[s];
        for (int a = 0; a < 3; ++a) {
            si.cube[a] = c[a];
            cube_bounds(c[a], cube_size, half, buffer_size, si.lo[a], si.hi[a], si.slo[a], si.shi[a]);
            volatile float ext = si.shi[a] - si.slo[a];
            volatile float q = ext / voxel_size;
            si.grid[a] = (int)std::round((float)q);            // spconv calc_meta_data (App. C)
            SQSM_CHECK(si.grid[a] > 0 && si.grid[a] <= 65535 * 8, "voxel grid per cube out of range");
            mn[a] = std::min(mn[a], c[a]);
            mx[a] = std::max(mx[a], c[a]);
        }
    }
    TileParams tp;
    tp.cs = cube_size; tp.half = half; tp.buf = buffer_size;
    tp.ox = mn[0] - 1; tp.oy = mn[1] - 1; tp.oz = mn[2] - 1;
    tp.dx = mx[0] - mn[0] + 3; tp.dy = mx[1] - mn[1] + 3; tp.dz = mx[2] - mn[2] + 3;
    int64_t cells = (int64_t)tp.dx * tp.dy * tp.dz;
    SQSM_CHECK(cells <= (1ll << 26), "cube bounding box too sparse for the dense cube table");
    std::vector<int32_t> lut((size_t)cells, -1);
    for (int s = 0; s < nc; ++s) {
        const SampleInfo& si = out.samples[s];
        lut[((size_t)(si.cube[0] - tp.ox) * tp.dy + (si.cube[1] - tp.oy)) * tp.dz + (si.cube[2] - tp.oz)] = s;
    }
    DBuf<int32_t> d_lut(cells, st);
    SQSM_CUDA(cudaMemcpyAsync(d_lut.p, lut.data(), sizeof(int32_t) * cells, cudaMemcpyHostToDevice, st));
    out.d_samples.alloc(nc, st);
    SQSM_CUDA(cudaMemcpyAsync(out.d_samples.p, out.samples.data(), sizeof(SampleInfo) * nc, cudaMemcpyHostToDevice, st));
    tp.lut = d_lut.p;
    // 3. membership: count, scan, emit (entries in ascending point order)
    DBuf<int32_t> cnt(n + 1, st), off(n + 1, st);
    SQSM_CUDA(cudaMemsetAsync(cnt.p + n, 0, sizeof(int32_t), st));
    k_member_count<<<div_up(n, 256), 256, 0, st>>>(d_P, n, tp, cnt.p);
    cx.launches++;
    exclusive_scan_i32(cx, cnt.p, off.p, n + 1);
    int32_t E32 = 0;
    SQSM_CUDA(cudaMemcpyAsync(&E32, off.p + n, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    SQSM_CUDA(cudaStreamSynchronize(st));      // also guarantees lut / samples were consumed from live host memory
    int64_t E = E32;
    // a point belongs to at most 8 cube samples; below 2^28 points the 32-bit scan cannot wrap, above it a negative or
    // too small total gives the wrap away (every point has at least one membership: its own cube)
    SQSM_CHECK(E >= n && E > 0, "cube-sample memberships overflow 2^31 entries (shard the cloud: ShardedTreeContraction)");
    out.n_entries = E;
    DBuf<uint32_t> es(E, st), ev(E, st);
    k_member_emit<<<div_up(n, 256), 256, 0, st>>>(d_P, n, tp, off.p, es.p, ev.p);
    cx.launches++;
    // 4. group by sample (stable => ascending point index inside every sample, synthetic_tree.py:268)
    out.ent_sample.alloc(E, st);
    out.ent_val.alloc(E, st);
    int bits = 1;
    while ((1 << bits) < nc) ++bits;
    sort_pairs_u32(cx, es.p, out.ent_sample.p, ev.p, out.ent_val.p, E, 0, bits);
    DBuf<int64_t> d_start((int64_t)nc + 1, st);
    d_start.zero();                            // a sample without entries (buffer_size = 0 corner cases) keeps a defined start
    k_boundaries<<<div_up(E, 256), 256, 0, st>>>(out.ent_sample.p, E, d_start.p);
    cx.launches++;
    out.start.assign((size_t)nc + 1, 0);
    SQSM_CUDA(cudaMemcpyAsync(out.start.data(), d_start.p, sizeof(int64_t) * nc, cudaMemcpyDeviceToHost, st));
    SQSM_CUDA(cudaStreamSynchronize(st));
    out.start[nc] = E;
}

// ---------------------------------------------------------------------------------- level-0 voxelisation

// PointToVoxel (CPU rule): c = floor((p - slo) / vsize) in float32, dropped unless 0 <= c < grid.
__global__ void __launch_bounds__(256) k_entry_voxel(const float* __restrict__ P, const uint32_t* __restrict__ ent_sample,
                                                     const uint32_t* __restrict__ ent_val, int64_t e0, int64_t n,
                                                     const SampleInfo* __restrict__ samples, float vsize,
                                                     uint64_t* __restrict__ ent_key, uint16_t* __restrict__ ent_pos) {
    int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n) return;
    uint32_t s = ent_sample[e0 + e];
    uint32_t pt = ent_val[e0 + e] & 0x7FFFFFFFu;
    const SampleInfo si = samples[s];
    int c[3];
    bool ok = true;
    #pragma unroll
    for (int a = 0; a < 3; ++a) {
        float f = floorf(__fdiv_rn(__fsub_rn(P[3 * (size_t)pt + a], si.slo[a]), vsize));
        ok = ok && (f >= 0.0f) && (f < (float)si.grid[a]);
        c[a] = (int)f;
    }
    if (!ok) { ent_key[e] = kEmptyKey; ent_pos[e] = 0; return; }
    ent_key[e] = pack_key(s, (uint32_t)(c[2] >> 3), (uint32_t)(c[1] >> 3), (uint32_t)(c[0] >> 3));
    ent_pos[e] = (uint16_t)pos_of(c[2] & 7, c[1] & 7, c[0] & 7);
}

// first-come rule: the entry with the lowest index (= lowest point index inside its sample) wins the voxel
__global__ void __launch_bounds__(256) k_winner(const Brick* __restrict__ bricks, const int32_t* __restrict__ ent_brick,
                                                const uint16_t* __restrict__ ent_pos, int64_t n,
                                                int32_t* __restrict__ ent_row, uint32_t* __restrict__ winner) {
    int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n) return;
    int b = ent_brick[e];
    if (b < 0) { ent_row[e] = -1; return; }
    int row = brick_row(&bricks[b], ent_pos[e]);
    ent_row[e] = row;
    atomicMin(&winner[row], (uint32_t)e);
}

// flags for the canonical numbering: low word counts winners, high word counts interior winners
__global__ void __launch_bounds__(256) k_winner_flags(const int32_t* __restrict__ ent_row, const uint32_t* __restrict__ winner,
                                                      const uint32_t* __restrict__ ent_val, int64_t e0, int64_t n,
                                                      int64_t* __restrict__ flags) {
    int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e > n) return;
    int64_t f = 0;
    if (e < n) {
        int row = ent_row[e];
        if (row >= 0 && winner[row] == (uint32_t)e) f = 1 + (((int64_t)(ent_val[e0 + e] >> 31)) << 32);
    }
    flags[e] = f;      // flags[n] = 0 so that the exclusive scan yields the totals at [n]
}

// one thread per ROW (= voxel): the winner entry of the row is looked up, so the six row arrays are written coalesced and only
// the reads (winner entry -> point) are scattered (entry-parallel with scattered 45-byte writes: 1.43 ms per C3 thunk)
__global__ void __launch_bounds__(256) k_rows_fill(const float* __restrict__ P, const float* __restrict__ D,
                                                   const uint32_t* __restrict__ winner,
                                                   const uint32_t* __restrict__ ent_val, const int64_t* __restrict__ scan,
                                                   int64_t e0, int64_t n_rows, int64_t out_base, float4* __restrict__ xyz,
                                                   float4* __restrict__ disp, int32_t* __restrict__ pt_out,
                                                   int32_t* __restrict__ out_pos, int32_t* __restrict__ canon,
                                                   uint8_t* __restrict__ interior) {
    int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= n_rows) return;
    const uint32_t e = winner[row];
    uint32_t v = ent_val[e0 + e];
    uint32_t pt = v & 0x7FFFFFFFu;
    bool inner = (v >> 31) != 0;
    int64_t sc = scan[e];
    xyz[row] = make_float4(P[3 * (size_t)pt], P[3 * (size_t)pt + 1], P[3 * (size_t)pt + 2], 0.0f);
    disp[row] = make_float4(D[3 * (size_t)pt], D[3 * (size_t)pt + 1], D[3 * (size_t)pt + 2], 0.0f);
    pt_out[row] = (int32_t)pt;
    canon[row] = (int32_t)(sc & 0xFFFFFFFFll);
    out_pos[row] = inner ? (int32_t)(out_base + (sc >> 32)) : -1;
    interior[row] = inner ? 1 : 0;
}

// collate: spatial_shape = per-axis max voxel index over the batch (synthetic_tree.py:201, F6)
__global__ void __launch_bounds__(256) k_batch_shape(const Brick* __restrict__ bricks, int n_bricks, int batch_size,
                                                     int batch0, int32_t* __restrict__ shape) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    int batch = -1, mz = 0, my = 0, mxx = 0;
    if (i < n_bricks) {
        const Brick& b = bricks[i];
        batch = (int)key_s(b.key) / batch_size - batch0;
        int zl = 0, yl = 0; uint32_t xb = 0;
        #pragma unroll
        for (int w = 0; w < 16; ++w) {
            uint32_t m = b.mask[w];
            if (m) {
                zl = w >> 1;
                yl = max(yl, ((w & 1) << 2) + ((31 - __clz(m)) >> 3));
                xb |= (m | (m >> 8) | (m >> 16) | (m >> 24)) & 0xFFu;
            }
        }
        mz = (int)key_z(b.key) * 8 + zl;
        my = (int)key_y(b.key) * 8 + yl;
        mxx = (int)key_x(b.key) * 8 + (31 - __clz(xb));
    }
    // bricks are sorted by sample, so a warp almost always holds one batch: reduce before the atomics
    unsigned live = __ballot_sync(0xFFFFFFFFu, batch >= 0);
    if (batch < 0) return;
    unsigned peers = __match_any_sync(live, batch);
    int lane = threadIdx.x & 31;
    int leader = __ffs(peers) - 1;
    for (unsigned rest = peers & ~(1u << leader); rest; rest &= rest - 1) {
        int src = __ffs(rest) - 1;
        int oz = __shfl_sync(peers, mz, src), oy = __shfl_sync(peers, my, src), ox = __shfl_sync(peers, mxx, src);
        if (lane == leader) { mz = max(mz, oz); my = max(my, oy); mxx = max(mxx, ox); }
    }
    if (lane == leader) {
        atomicMax(&shape[batch * 3 + 0], mz);
        atomicMax(&shape[batch * 3 + 1], my);
        atomicMax(&shape[batch * 3 + 2], mxx);
    }
}

void build_level0(Ctx& cx, const float* d_P, const float* d_D, const TilingResult& tl, int s0, int s1, int batch_size,
                  float voxel_size, LevelTables& L0, Level0Rows& rows, int64_t out_base, int64_t* n_out,
                  DBuf<int32_t>& ent_row) {
    cudaStream_t st = cx.st;
    int64_t e0 = tl.start[s0], n = tl.start[s1] - e0;
    for (int s = s0; s < s1; ++s)
        SQSM_CHECK(tl.start[s + 1] - tl.start[s] <= kMaxIdxF32,
                   "a cube sample holds more than 2^24 points: the reference's float32 index column (F7) "
                   "is no longer exact; not supported");
    DBuf<uint64_t> ent_key(n, st);
    DBuf<uint16_t> ent_pos(n, st);
    DBuf<int32_t> ent_brick(n, st);
    int grid = div_up(n, 256);
    k_entry_voxel<<<grid, 256, 0, st>>>(d_P, tl.ent_sample.p, tl.ent_val.p, e0, n, tl.d_samples.p, voxel_size,
                                        ent_key.p, ent_pos.p);
    cx.launches++;
    {
        ProfScope ps(cx, "vx_brickmap", (double)n * 14.0 * 2, 0.0);
        brickmap_build(cx, L0.map, ent_key.p, ent_pos.p, n, ent_brick.p, n / 16);
    }
    L0.n = L0.map.n_rows;
    SQSM_CHECK(L0.n > 0, "no voxel survived voxelisation");
    int64_t n0 = L0.n;
    ent_row.alloc(n, st);
    DBuf<uint32_t> winner(n0, st);
    winner.fill_ff();
    k_winner<<<grid, 256, 0, st>>>(L0.map.bricks.p, ent_brick.p, ent_pos.p, n, ent_row.p, winner.p);
    DBuf<int64_t> flags(n + 1, st), scan(n + 1, st);
    k_winner_flags<<<div_up(n + 1, 256), 256, 0, st>>>(ent_row.p, winner.p, tl.ent_val.p, e0, n, flags.p);
    cx.launches += 2;
    exclusive_scan_i64(cx, flags.p, scan.p, n + 1);
    rows.xyz.alloc(n0, st);
    rows.disp.alloc(n0, st);
    rows.pt.alloc(n0, st);
    rows.out.alloc(n0, st);
    rows.canon.alloc(n0, st);
    rows.interior.alloc(n0, st);
    k_rows_fill<<<div_up(n0, 256), 256, 0, st>>>(d_P, d_D, winner.p, tl.ent_val.p, scan.p, e0, n0, out_base, rows.xyz.p,
                                                 rows.disp.p, rows.pt.p, rows.out.p, rows.canon.p, rows.interior.p);
    cx.launches++;
    int64_t tot = 0;
    SQSM_CUDA(cudaMemcpyAsync(&tot, scan.p + n, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    // batch shapes
    int batch0 = s0 / batch_size;
    int nb = (s1 - 1) / batch_size - batch0 + 1;
    L0.shape.alloc((int64_t)nb * 3, st);
    L0.shape.zero();
    k_batch_shape<<<div_up(L0.map.n_bricks, 256), 256, 0, st>>>(L0.map.bricks.p, L0.map.n_bricks, batch_size, batch0,
                                                               L0.shape.p);
    cx.launches++;
    L0.row_brick.alloc(n0, st);
    L0.row_pos.alloc(n0, st);
    brickmap_expand_rows(cx, L0.map, L0.row_brick.p, L0.row_pos.p);
    SQSM_CUDA(cudaStreamSynchronize(st));
    SQSM_CHECK((tot & 0xFFFFFFFFll) == n0, "internal error: winner count != voxel count");
    *n_out = tot >> 32;
}

// ---------------------------------------------------------------------------------- SubM rulebook

// nbr[k][o] = row of the active voxel at coords(o) + (kz-1, ky-1, kx-1), k = kz*9+ky*3+kx, else -1.
// "Clean bounds" rule: output rows outside [0, spatial_shape) keep the centre only.
// DENSE: the [27][n] table (-1 = absent).  CSR: a 27-bit presence mask, an offset and the present rows only, in ascending
// kernel offset - 30 instead of 108 bytes per row at level 0, where 5 of the 27 neighbours exist and every convolution
// of the level re-reads the table from HBM.  The space of a block's rows is reserved with one atomicAdd per block (a block's
// rows are contiguous in the list; the order of the blocks in it varies from run to run, the content of a row does not).
template <bool DENSE, bool CSR>
__global__ void __launch_bounds__(256) k_subm_nbr(const Brick* __restrict__ bricks, const int32_t* __restrict__ bnbr,
                                                  const int32_t* __restrict__ row_brick, const uint16_t* __restrict__ row_pos,
                                                  int64_t n, const int32_t* __restrict__ shape, int batch_size, int batch0,
                                                  int32_t* __restrict__ nbr, uint32_t* __restrict__ c_mask,
                                                  uint32_t* __restrict__ c_off, int32_t* __restrict__ c_list,
                                                  unsigned long long* __restrict__ stats) {
    int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int pairs = 0, far = 0;
    int rr[CSR ? 27 : 1];
    if (CSR) {
        #pragma unroll
        for (int k = 0; k < 27; ++k) rr[k] = -1;
    }
    uint32_t mask = 0;
    if (o < n) {
        int b = row_brick[o];
        uint32_t pos = row_pos[o];
        int z = (int)(pos >> 6), y = (int)(pos >> 3) & 7, x = (int)pos & 7;
        const Brick* B = &bricks[b];
        uint64_t key = B->key;
        int batch = (int)key_s(key) / batch_size - batch0;
        int gz = (int)key_z(key) * 8 + z, gy = (int)key_y(key) * 8 + y, gx = (int)key_x(key) * 8 + x;
        bool inb = gz < shape[batch * 3 + 0] && gy < shape[batch * 3 + 1] && gx < shape[batch * 3 + 2];
        far = inb ? 0 : 1;
        const int32_t* nb = bnbr + (size_t)b * 32;
        #pragma unroll
        for (int kz = 0; kz < 3; ++kz) {
            int zz = z + kz - 1;
            int bz = (zz < 0) ? 0 : (zz > 7 ? 2 : 1);
            #pragma unroll
            for (int ky = 0; ky < 3; ++ky) {
                int yy = y + ky - 1;
                int by = (yy < 0) ? 0 : (yy > 7 ? 2 : 1);
                #pragma unroll
                for (int kx = 0; kx < 3; ++kx) {
                    int k = kz * 9 + ky * 3 + kx;
                    int r = -1;
                    if (k == 13) r = (int)o;
                    else if (inb) {
                        int xx = x + kx - 1;
                        int bx = (xx < 0) ? 0 : (xx > 7 ? 2 : 1);
                        int nbid = nb[bz * 9 + by * 3 + bx];
                        if (nbid >= 0) r = brick_row(&bricks[nbid], pos_of(zz & 7, yy & 7, xx & 7));
                    }
                    if (DENSE) nbr[(size_t)k * n + o] = r;
                    if (CSR) { rr[k] = r; mask |= (r >= 0 ? 1u : 0u) << k; }
                    pairs += (r >= 0);
                }
            }
        }
    }
    // block-level tally of active pairs / far-face rows (one atomic per block: the counters share a cache line); the CSR
    // form reserves the block's run of the list with the same atomic
    const int lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
    __shared__ int s_wtot[8], s_wfar[8];
    __shared__ unsigned long long s_base;
    int incl = pairs;
    #pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int v = __shfl_up_sync(0xFFFFFFFFu, incl, d);
        if (lane >= d) incl += v;
    }
    const int total = __shfl_sync(0xFFFFFFFFu, incl, 31);
    far = __reduce_add_sync(0xFFFFFFFFu, far);
    if (lane == 0) { s_wtot[wp] = total; s_wfar[wp] = far; }
    __syncthreads();
    if (threadIdx.x == 0) {
        int run = 0, f = 0;
        #pragma unroll
        for (int i = 0; i < 8; ++i) { const int v = s_wtot[i]; s_wtot[i] = run; run += v; f += s_wfar[i]; }
        s_base = run ? atomicAdd(&stats[0], (unsigned long long)run) : 0ull;
        if (f) atomicAdd(&stats[1], (unsigned long long)f);
    }
    __syncthreads();
    if (CSR) {
        const unsigned long long base = s_base + (unsigned long long)s_wtot[wp];
        // the warp's entries are staged in shared memory and leave as one contiguous, coalesced run
        __shared__ int32_t s_list[8][27 * 32];
        int32_t* sl = s_list[wp];
        int w_off = incl - pairs;
        if (o < n) {
            c_mask[o] = mask;
            c_off[o] = (uint32_t)base + (uint32_t)w_off;
        }
        #pragma unroll
        for (int k = 0; k < 27; ++k)
            if (rr[k] >= 0) sl[w_off++] = rr[k];
        __syncwarp();
        for (int j = lane; j < total; j += 32) c_list[base + j] = sl[j];
    }
}

void build_subm(Ctx& cx, LevelTables& L, int batch0, int n_batches, int batch_size, int64_t* far_face, bool want_dense,
                bool want_csr) {
    cudaStream_t st = cx.st;
    (void)n_batches;
    if (L.n == 0) return;
    SQSM_CHECK(want_dense || want_csr, "build_subm: no table requested");
    brickmap_build_neighbours(cx, L.map);
    if (want_dense) L.nbr.alloc(27 * L.n, st);
    if (want_csr) {
        SQSM_CHECK(27 * L.n < (1ll << 32), "level too large for 32-bit rulebook offsets (lower SQSM_MAX_ENTRIES_PER_PASS)");
        L.nbr_mask.alloc(L.n, st);
        L.nbr_off.alloc(L.n, st);
        L.nbr_list.alloc(27 * L.n, st);      // upper bound; only the first `pairs` entries are ever touched
    }
    ProfScope ps(cx, "rb_subm_nbr", (double)L.n * ((want_dense ? 27.0 * 4 : 0.0) + (want_csr ? 8.0 : 0.0) + 6.0), 0.0);
    DBuf<unsigned long long> stats(2, st);
    stats.zero();
    const int grid = div_up(L.n, 256);
#define SQSM_SUBM(D_, C_) k_subm_nbr<D_, C_><<<grid, 256, 0, st>>>(L.map.bricks.p, L.map.bnbr.p, L.row_brick.p, L.row_pos.p, L.n, \
        L.shape.p, batch_size, batch0, L.nbr.p, L.nbr_mask.p, L.nbr_off.p, L.nbr_list.p, stats.p)
    if (want_dense && want_csr) SQSM_SUBM(true, true);
    else if (want_csr) SQSM_SUBM(false, true);
    else SQSM_SUBM(true, false);
#undef SQSM_SUBM
    cx.launches++;
    unsigned long long hs[2];
    SQSM_CUDA(cudaMemcpyAsync(hs, stats.p, sizeof(hs), cudaMemcpyDeviceToHost, st));
    SQSM_CUDA(cudaStreamSynchronize(st));
    L.pairs = (int64_t)hs[0];
    if (want_csr) ps.r.bytes += 4.0 * (double)L.pairs;
    if (far_face) *far_face = (int64_t)hs[1];
}

// ---------------------------------------------------------------------------------- k=2 s=2 rulebook

__device__ __forceinline__ int floordiv2(int a) { return (a >= 0) ? (a >> 1) : -((-a + 1) >> 1); }

// out_shape = (shape - 2)//2 + 1 per axis; a batch may descend only if all(shape > 2) (sparse_module.py:105-109)
__global__ void k_next_shape(const int32_t* __restrict__ shape, int n_batches, int32_t* __restrict__ next,
                             int32_t* __restrict__ geom_ok) {
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= n_batches) return;
    bool ok = true;
    for (int a = 0; a < 3; ++a) {
        int s = shape[b * 3 + a];
        ok = ok && (s > 2);
        next[b * 3 + a] = floordiv2(s - 2) + 1;
    }
    geom_ok[b] = ok ? 1 : 0;
}

__global__ void __launch_bounds__(256) k_down_entries(const Brick* __restrict__ bricks, const int32_t* __restrict__ row_brick,
                                                      const uint16_t* __restrict__ row_pos, int64_t n,
                                                      const int32_t* __restrict__ next_shape, const int32_t* __restrict__ geom_ok,
                                                      int batch_size, int batch0, uint64_t* __restrict__ ent_key,
                                                      uint16_t* __restrict__ ent_pos, int32_t* __restrict__ koff,
                                                      int32_t* __restrict__ desc) {
    int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    uint64_t key = bricks[row_brick[r]].key;
    uint32_t pos = row_pos[r];
    int s = (int)key_s(key);
    int batch = s / batch_size - batch0;
    int gz = (int)key_z(key) * 8 + (int)(pos >> 6), gy = (int)key_y(key) * 8 + ((int)(pos >> 3) & 7),
        gx = (int)key_x(key) * 8 + ((int)pos & 7);
    koff[r] = ((gz & 1) << 2) | ((gy & 1) << 1) | (gx & 1);
    int oz = gz >> 1, oy = gy >> 1, ox = gx >> 1;
    bool ok = geom_ok[batch] && oz < next_shape[batch * 3 + 0] && oy < next_shape[batch * 3 + 1] &&
              ox < next_shape[batch * 3 + 2];
    if (!ok) { ent_key[r] = kEmptyKey; ent_pos[r] = 0; return; }
    desc[batch] = 1;                     // benign race: every writer stores 1
    ent_key[r] = pack_key((uint32_t)s, (uint32_t)(oz >> 3), (uint32_t)(oy >> 3), (uint32_t)(ox >> 3));
    ent_pos[r] = (uint16_t)pos_of(oz & 7, oy & 7, ox & 7);
}

__global__ void __launch_bounds__(256) k_down_tables(const Brick* __restrict__ cbricks, const int32_t* __restrict__ ent_brick,
                                                     const uint16_t* __restrict__ ent_pos, const int32_t* __restrict__ koff,
                                                     int64_t n, int64_t n_next, int32_t* __restrict__ parent,
                                                     int32_t* __restrict__ child) {
    int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    int b = ent_brick[r];
    int p = -1;
    if (b >= 0) p = brick_row(&cbricks[b], ent_pos[r]);
    parent[r] = p;
    if (p >= 0) {
        int k = koff[r];
        child[(size_t)k * n_next + p] = (int32_t)r;
    }
}

__global__ void __launch_bounds__(256) k_row_desc(const Brick* __restrict__ bricks, const int32_t* __restrict__ row_brick,
                                                  int64_t n, const int32_t* __restrict__ desc, int batch_size, int batch0,
                                                  uint8_t* __restrict__ row_desc) {
    int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    int batch = (int)key_s(bricks[row_brick[r]].key) / batch_size - batch0;
    row_desc[r] = desc[batch] ? 1 : 0;
}

// Returns false when no batch of this super-batch descends below level L.
bool build_down(Ctx& cx, LevelTables& L, LevelTables& Ln, int batch0, int n_batches, int batch_size) {
    cudaStream_t st = cx.st;
    int64_t n = L.n;
    Ln.n = 0;
    Ln.shape.alloc((int64_t)n_batches * 3, st);
    DBuf<int32_t> geom_ok(n_batches, st);
    L.desc.alloc(n_batches, st);
    L.desc.zero();
    k_next_shape<<<div_up(n_batches, 128), 128, 0, st>>>(L.shape.p, n_batches, Ln.shape.p, geom_ok.p);
    DBuf<uint64_t> ent_key(n, st);
    DBuf<uint16_t> ent_pos(n, st);
    DBuf<int32_t> ent_brick(n, st);
    L.koff.alloc(n, st);
    k_down_entries<<<div_up(n, 256), 256, 0, st>>>(L.map.bricks.p, L.row_brick.p, L.row_pos.p, n, Ln.shape.p, geom_ok.p,
                                                   batch_size, batch0, ent_key.p, ent_pos.p, L.koff.p, L.desc.p);
    cx.launches += 2;
    L.row_desc.alloc(n, st);
    k_row_desc<<<div_up(n, 256), 256, 0, st>>>(L.map.bricks.p, L.row_brick.p, n, L.desc.p, batch_size, batch0,
                                               L.row_desc.p);
    cx.launches++;
    {
        ProfScope ps(cx, "rb_brickmap", (double)n * 14.0 * 2, 0.0);
        brickmap_build(cx, Ln.map, ent_key.p, ent_pos.p, n, ent_brick.p, std::max<int64_t>(L.map.n_bricks / 4, 1024));
    }
    Ln.n = Ln.map.n_rows;
    L.parent.alloc(n, st);
    if (Ln.n == 0) {
        L.parent.fill_ff();
        return false;
    }
    L.child.alloc(8 * Ln.n, st);
    L.child.fill_ff();
    k_down_tables<<<div_up(n, 256), 256, 0, st>>>(Ln.map.bricks.p, ent_brick.p, ent_pos.p, L.koff.p, n, Ln.n, L.parent.p,
                                                  L.child.p);
    cx.launches++;
    Ln.row_brick.alloc(Ln.n, st);
    Ln.row_pos.alloc(Ln.n, st);
    brickmap_expand_rows(cx, Ln.map, Ln.row_brick.p, Ln.row_pos.p);
    return true;
}

}  // namespace sqsm
