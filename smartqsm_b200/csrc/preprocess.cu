// Cloud preparation in front of the contraction path.  sm_100a.
//
// Replaces, on the device, entrypoints/smartqsm.py:882-913 of the reference:
//   :882-888  bounding box -> global shift -> translate
//   :889      Open3D voxel_down_sample(voxel_down_size): float64 mean of the points of every voxel
//   :895-898  Open3D remove_statistical_outlier(nb_neighbors, std_ratio)
//   :908-913  density adaptation: scale = voxel_down_size / mean distance to the nearest other point (SciPy KDTree)
// Everything is float64 like Open3D / SciPy.  The voxel grid of the down-sampling doubles as the search structure of
// both neighbour queries: a mean lies inside its own voxel, so the BrickMap (8^3-voxel bricks, rows in brick order)
// indexes the down-sampled cloud directly.
//
// Determinism: voxel sums run in input order (stable sort of the points by voxel row + one thread per voxel), global
// sums use a fixed reduction tree - same input, same bits, run to run.  Output order = BrickMap row order (bricks by
// key (z, y, x), then z, y, x inside the brick); Open3D's own order is that of a std::unordered_map and not reproducible.
#include "engine.h"
#include <cmath>
#include <cstdlib>
#include <cstring>

namespace sqsm {

namespace {

__device__ __forceinline__ unsigned long long d2ord(double d) {
    unsigned long long u = (unsigned long long)__double_as_longlong(d);
    return (u & 0x8000000000000000ull) ? ~u : (u | 0x8000000000000000ull);
}
double ord2d(unsigned long long o) {
    unsigned long long u = (o & 0x8000000000000000ull) ? (o & 0x7FFFFFFFFFFFFFFFull) : ~o;
    double d;
    memcpy(&d, &u, 8);
    return d;
}

__global__ void __launch_bounds__(256) k_bounds_f64(const double* __restrict__ P, int64_t n, unsigned long long* __restrict__ out,
                                                    int32_t* __restrict__ bad) {
    unsigned long long mn[3] = {~0ull, ~0ull, ~0ull}, mx[3] = {0, 0, 0};
    bool nonfinite = false;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        #pragma unroll
        for (int a = 0; a < 3; ++a) {
            const double v = P[3 * i + a];
            nonfinite = nonfinite || !isfinite(v);
            const unsigned long long o = d2ord(v);
            mn[a] = min(mn[a], o);
            mx[a] = max(mx[a], o);
        }
    }
    #pragma unroll
    for (int a = 0; a < 3; ++a)
        for (int d = 16; d > 0; d >>= 1) {
            mn[a] = min(mn[a], __shfl_xor_sync(0xFFFFFFFFu, mn[a], d));
            mx[a] = max(mx[a], __shfl_xor_sync(0xFFFFFFFFu, mx[a], d));
        }
    if ((threadIdx.x & 31) == 0) {
        #pragma unroll
        for (int a = 0; a < 3; ++a) {
            atomicMin(&out[a], mn[a]);
            atomicMax(&out[3 + a], mx[a]);
        }
    }
    if (nonfinite) atomicOr(bad, 1);
}

struct Shift3 { double s[3]; };

__global__ void __launch_bounds__(256) k_translate(double* __restrict__ P, int64_t n, const __grid_constant__ Shift3 sh) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    #pragma unroll
    for (int a = 0; a < 3; ++a) P[3 * i + a] = __dadd_rn(P[3 * i + a], sh.s[a]);          // cloud.translate (:888)
}

struct PrepGrid { double vmin[3]; double voxel; };

// Open3D voxel_down_sample: index = floor((p - (min_bound - voxel / 2)) / voxel), float64
__global__ void __launch_bounds__(256) k_prep_entries(const double* __restrict__ P, int64_t n, const __grid_constant__ PrepGrid g,
                                                      uint64_t* __restrict__ ent_key, uint16_t* __restrict__ ent_pos,
                                                      int32_t* __restrict__ err) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int v[3];
    bool ok = true;
    #pragma unroll
    for (int a = 0; a < 3; ++a) {
        const double f = floor(__ddiv_rn(__dsub_rn(P[3 * i + a], g.vmin[a]), g.voxel));
        ok = ok && (f >= 0.0) && (f < 524288.0);
        v[a] = (int)f;
    }
    if (!ok) { ent_key[i] = kEmptyKey; ent_pos[i] = 0; atomicOr(err, 1); return; }
    ent_key[i] = pack_key(0u, (uint32_t)(v[2] >> 3), (uint32_t)(v[1] >> 3), (uint32_t)(v[0] >> 3));
    ent_pos[i] = (uint16_t)pos_of(v[2] & 7, v[1] & 7, v[0] & 7);
}

__global__ void __launch_bounds__(256) k_prep_rows(const Brick* __restrict__ bricks, const int32_t* __restrict__ ent_brick,
                                                   const uint16_t* __restrict__ ent_pos, int64_t n, uint32_t* __restrict__ row,
                                                   uint32_t* __restrict__ idx) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    row[i] = (uint32_t)brick_row(&bricks[ent_brick[i]], ent_pos[i]);
    idx[i] = (uint32_t)i;
}

// sorted (row, point) pairs -> first position of every row (every row owns at least one point)
__global__ void __launch_bounds__(256) k_row_starts(const uint32_t* __restrict__ row_sorted, int64_t n, int32_t* __restrict__ start) {
    int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    if (j == 0 || row_sorted[j] != row_sorted[j - 1]) start[row_sorted[j]] = (int32_t)j;
}

// one thread per voxel: sum its points in input order (the sort is stable), divide by the count
__global__ void __launch_bounds__(256) k_row_means(const double* __restrict__ P, const uint32_t* __restrict__ idx_sorted,
                                                   const int32_t* __restrict__ start, int64_t rows, int64_t n,
                                                   double* __restrict__ V) {
    int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= rows) return;
    const int64_t a = start[r], b = (r + 1 < rows) ? start[r + 1] : n;
    double sx = 0.0, sy = 0.0, sz = 0.0;
    for (int64_t j = a; j < b; ++j) {
        const size_t i = idx_sorted[j];
        sx = __dadd_rn(sx, P[3 * i]);
        sy = __dadd_rn(sy, P[3 * i + 1]);
        sz = __dadd_rn(sz, P[3 * i + 2]);
    }
    const double c = (double)(b - a);
    V[3 * r] = __ddiv_rn(sx, c);
    V[3 * r + 1] = __ddiv_rn(sy, c);
    V[3 * r + 2] = __ddiv_rn(sz, c);
}

constexpr int kKnnMax = 32;
__constant__ int c_knn_order[27] = {13, 4, 10, 12, 14, 16, 22, 1, 3, 5, 7, 9, 11, 15, 17, 19, 21, 23, 25, 0, 2, 6, 8, 18, 20, 24, 26};

struct KnnBest {
    double d[kKnnMax];
    int k;
    __device__ __forceinline__ void init(int kk) {
        k = kk;
        for (int j = 0; j < kk; ++j) d[j] = 1e300;
    }
    __device__ __forceinline__ double worst() const { return d[k - 1]; }
    __device__ __forceinline__ void push(double v) {           // sorted insertion, ascending
        if (!(v < d[k - 1])) return;
        int j = k - 1;
        while (j > 0 && d[j - 1] > v) { d[j] = d[j - 1]; --j; }
        d[j] = v;
    }
};

// Exact k nearest neighbours (the query itself included, like a KD-tree query at a cloud point) over the voxel means.
// Thread per query; bricks nearest first with a float64 lower bound against the current k-th distance; after the 27
// bricks everything else is at least one brick edge away; a query whose k-th distance exceeds that bound (isolated
// points - the very outliers the filter is after) is handed to k_knn_far.  mode 0: mean of the k distances (Open3D avg_distances); mode 1: distance to the nearest OTHER point.
// `alive` (optional) restricts both queries and candidates to a subset of the rows.
__global__ void __launch_bounds__(128) k_knn(const Brick* __restrict__ bricks, const int32_t* __restrict__ bnbr, BrickHash h,
                                             const int32_t* __restrict__ row_brick, const double* __restrict__ V,
                                             const uint8_t* __restrict__ alive, int64_t rows, int k,
                                             const __grid_constant__ PrepGrid g, int mode, double* __restrict__ out,
                                             int32_t* __restrict__ far_list, int32_t* __restrict__ n_far) {
    int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= rows) return;
    if (alive && !alive[r]) { out[r] = -1.0; return; }
    const double qx = V[3 * r], qy = V[3 * r + 1], qz = V[3 * r + 2];
    const int b = row_brick[r];
    const uint64_t key = bricks[b].key;
    const int bz = (int)key_z(key), by = (int)key_y(key), bx = (int)key_x(key);
    const double edge = 8.0 * g.voxel;
    const double ox = g.vmin[0] + (double)bx * edge, oy = g.vmin[1] + (double)by * edge, oz = g.vmin[2] + (double)bz * edge;
    const double slack = 1e-9 * edge;                     // the mean may sit a rounding error outside its voxel
    const double gm[3] = {fmax(qx - ox - slack, 0.0), fmax(qy - oy - slack, 0.0), fmax(qz - oz - slack, 0.0)};
    const double gp[3] = {fmax(ox + edge - qx - slack, 0.0), fmax(oy + edge - qy - slack, 0.0), fmax(oz + edge - qz - slack, 0.0)};
    KnnBest best;
    best.init(k);
    auto scan = [&](int cb, int cn) {
        for (int i = 0; i < cn; ++i) {
            const size_t c = (size_t)(cb + i);
            if (alive && !alive[c]) continue;
            const double ex = qx - V[3 * c], ey = qy - V[3 * c + 1], ez = qz - V[3 * c + 2];
            best.push(ex * ex + ey * ey + ez * ez);
        }
    };
    #pragma unroll 1
    for (int step = 0; step < 27; ++step) {
        const int nb = c_knn_order[step];
        const int id = bnbr[(size_t)b * 32 + nb];
        if (id < 0) continue;
        const int dz = nb / 9 - 1, dy = (nb / 3) % 3 - 1, dx = nb % 3 - 1;
        const double gx = dx == 0 ? 0.0 : (dx < 0 ? gm[0] : gp[0]), gy = dy == 0 ? 0.0 : (dy < 0 ? gm[1] : gp[1]),
                     gz = dz == 0 ? 0.0 : (dz < 0 ? gm[2] : gp[2]);
        if ((gx * gx + gy * gy + gz * gz) * (1.0 - 1e-12) >= best.worst()) continue;
        scan(bricks[id].base, bricks[id].count);
    }
    const double near_edge = edge * (1.0 - 1e-9);
    if (best.worst() > near_edge * near_edge) {           // not settled inside the 27 bricks: a warp finishes it (k_knn_far)
        far_list[atomicAdd(n_far, 1)] = (int32_t)r;
        return;
    }
    if (mode == 0) {
        double s = 0.0;
        for (int j = 0; j < k; ++j) s = __dadd_rn(s, sqrt(best.d[j]));       // ascending, like accumulate over the sorted result
        out[r] = __ddiv_rn(s, (double)k);
    } else {
        out[r] = sqrt(best.d[1]);
    }
}

// Deferred queries, one warp each: rings of bricks (hash probes) are scanned by the 32 lanes together until the k-th
// distance is below the ring bound.  Every lane keeps its own sorted list; the warp-wide k smallest are read off by k
// rounds of "minimum of the list heads", which also yields them in ascending order for the final sum.
__device__ __forceinline__ double warp_min_f64(double v) {
    #pragma unroll
    for (int d = 16; d > 0; d >>= 1) v = fmin(v, __shfl_xor_sync(0xFFFFFFFFu, v, d));
    return v;
}
struct BrickBitmap { const uint32_t* bits; int nx, ny, nz, wx; };      // occupancy of the brick lattice, x-major bit rows

__global__ void __launch_bounds__(256) k_brick_bitmap(const Brick* __restrict__ bricks, int n_bricks, int ny, int wx,
                                                      uint32_t* __restrict__ bits) {
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= n_bricks) return;
    const uint64_t key = bricks[b].key;
    const size_t w = ((size_t)key_z(key) * ny + key_y(key)) * wx + (key_x(key) >> 5);
    atomicOr(&bits[w], 1u << (key_x(key) & 31));
}

__global__ void __launch_bounds__(128) k_knn_far(const Brick* __restrict__ bricks, BrickHash h, const int32_t* __restrict__ row_brick,
                                                 const double* __restrict__ V, const uint8_t* __restrict__ alive, int k,
                                                 const __grid_constant__ PrepGrid g, const __grid_constant__ BrickBitmap bm,
                                                 int max_ring, int mode, const int32_t* __restrict__ far_list,
                                                 const int32_t* __restrict__ n_far, double* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int n = *n_far;
    const double edge = 8.0 * g.voxel, near_edge = edge * (1.0 - 1e-9);
    for (int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; w < n; w += (gridDim.x * blockDim.x) >> 5) {
        const int64_t r = far_list[w];
        const double qx = V[3 * r], qy = V[3 * r + 1], qz = V[3 * r + 2];
        const uint64_t key = bricks[row_brick[r]].key;
        const int bz = (int)key_z(key), by = (int)key_y(key), bx = (int)key_x(key);
        KnnBest best;
        best.init(k);
        double bound = 1e300;                               // warp-wide k-th distance after the previous ring
        auto scan_brick = [&](int z, int y, int x) {
            const int id = hash_find(h, pack_key(0u, (uint32_t)z, (uint32_t)y, (uint32_t)x));
            if (id < 0) return;
            const int cb = bricks[id].base, cn = bricks[id].count;
            for (int i = 0; i < cn; ++i) {
                const size_t c = (size_t)(cb + i);
                if (alive && !alive[c]) continue;
                const double ex = qx - V[3 * c], ey = qy - V[3 * c + 1], ez = qz - V[3 * c + 2];
                const double d2 = ex * ex + ey * ey + ez * ez;
                if (d2 <= bound) best.push(d2);
            }
        };
        for (int ring = 0; ring <= max_ring; ++ring) {
            const int side = 2 * ring + 1;
            for (int t = lane; t < side * side; t += 32) {
                const int dz = t / side - ring, dy = t % side - ring;
                const int z = bz + dz, y = by + dy;
                if (z < 0 || y < 0 || z > 65535 || y > 65535) continue;
                const bool shell = (abs(dz) == ring) || (abs(dy) == ring);
                if (bm.bits != nullptr) {
                    if (z >= bm.nz || y >= bm.ny) continue;
                    const uint32_t* rowbits = bm.bits + ((size_t)z * bm.ny + y) * bm.wx;
                    const int x0 = max(bx - ring, 0), x1 = min(bx + ring, bm.nx - 1);
                    if (shell) {                                            // whole x row of the shell, 32 bricks per load
                        for (int wd = x0 >> 5; wd <= (x1 >> 5); ++wd) {
                            uint32_t m = rowbits[wd];
                            const int lo = wd * 32;
                            if (x0 > lo) m &= ~0u << (x0 - lo);
                            if (x1 < lo + 31) m &= ~0u >> (lo + 31 - x1);
                            while (m) {
                                const int x = lo + __ffs(m) - 1;
                                m &= m - 1;
                                scan_brick(z, y, x);
                            }
                        }
                    } else {                                                // only the two end caps
                        if (bx - ring >= 0 && ((rowbits[(bx - ring) >> 5] >> ((bx - ring) & 31)) & 1u)) scan_brick(z, y, bx - ring);
                        if (bx + ring < bm.nx && ((rowbits[(bx + ring) >> 5] >> ((bx + ring) & 31)) & 1u)) scan_brick(z, y, bx + ring);
                    }
                } else {
                    for (int dx = -ring; dx <= ring; dx += (shell || ring == 0 ? 1 : 2 * ring)) {
                        const int x = bx + dx;
                        if (x < 0 || x > 65535) continue;
                        scan_brick(z, y, x);
                    }
                }
            }
            // warp-wide k-th smallest
            int pos = 0;
            double kth = 1e300;
            for (int j = 0; j < k; ++j) {
                const double head = pos < k ? best.d[pos] : 1e300;
                kth = warp_min_f64(head);
                const unsigned who = __ballot_sync(0xFFFFFFFFu, head == kth);
                if (lane == __ffs(who) - 1) ++pos;
            }
            bound = kth;
            if (ring >= 1 && kth <= (double)ring * near_edge * (double)ring * near_edge) break;
        }
        // ascending read-out
        int pos = 0;
        double sum = 0.0, second = 0.0;
        for (int j = 0; j < k; ++j) {
            const double head = pos < k ? best.d[pos] : 1e300;
            const double m = warp_min_f64(head);
            const unsigned who = __ballot_sync(0xFFFFFFFFu, head == m);
            if (lane == __ffs(who) - 1) ++pos;
            sum = __dadd_rn(sum, sqrt(m));
            if (j == 1) second = sqrt(m);
        }
        if (lane == 0) out[r] = mode == 0 ? __ddiv_rn(sum, (double)k) : second;
    }
}

// Open3D remove_radius_outlier: number of voxel means with squared distance < r2 (the query itself included).  Thread per
// query; the bricks within ceil(r / edge) rings are taken from the neighbour table (ring 1) or probed in the hash, and
// skipped when their box is farther than r.  The count is stored as a double (same array as the k-NN averages).
__global__ void __launch_bounds__(128) k_radius_count(const Brick* __restrict__ bricks, const int32_t* __restrict__ bnbr, BrickHash h,
                                                      const int32_t* __restrict__ row_brick, const double* __restrict__ V, int64_t rows,
                                                      const __grid_constant__ PrepGrid g, double r2, int rings,
                                                      double* __restrict__ out) {
    int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= rows) return;
    const double qx = V[3 * r], qy = V[3 * r + 1], qz = V[3 * r + 2];
    const int b = row_brick[r];
    const uint64_t key = bricks[b].key;
    const int bz = (int)key_z(key), by = (int)key_y(key), bx = (int)key_x(key);
    const double edge = 8.0 * g.voxel, slack = 1e-9 * edge;
    int count = 0;
    for (int dz = -rings; dz <= rings; ++dz)
        for (int dy = -rings; dy <= rings; ++dy)
            for (int dx = -rings; dx <= rings; ++dx) {
                const int z = bz + dz, y = by + dy, x = bx + dx;
                if (z < 0 || y < 0 || x < 0 || z > 65535 || y > 65535 || x > 65535) continue;
                // distance from the query to the brick box
                const double ox = g.vmin[0] + (double)x * edge, oy = g.vmin[1] + (double)y * edge, oz = g.vmin[2] + (double)z * edge;
                const double gx = fmax(fmax(ox - qx, qx - ox - edge) - slack, 0.0), gy = fmax(fmax(oy - qy, qy - oy - edge) - slack, 0.0),
                             gz = fmax(fmax(oz - qz, qz - oz - edge) - slack, 0.0);
                if (gx * gx + gy * gy + gz * gz >= r2) continue;
                int id;
                if (dz >= -1 && dz <= 1 && dy >= -1 && dy <= 1 && dx >= -1 && dx <= 1) id = bnbr[(size_t)b * 32 + (dz + 1) * 9 + (dy + 1) * 3 + (dx + 1)];
                else id = hash_find(h, pack_key(0u, (uint32_t)z, (uint32_t)y, (uint32_t)x));
                if (id < 0) continue;
                const int cb = bricks[id].base, cn = bricks[id].count;
                for (int i = 0; i < cn; ++i) {
                    const size_t c = (size_t)(cb + i);
                    const double ex = V[3 * c] - qx, ey = V[3 * c + 1] - qy, ez = V[3 * c + 2] - qz;
                    const double d2 = __dadd_rn(__dadd_rn(__dmul_rn(ex, ex), __dmul_rn(ey, ey)), __dmul_rn(ez, ez));
                    count += d2 < r2 ? 1 : 0;
                }
            }
    out[r] = (double)count;
}

// fixed-tree sums: block b adds elements b, b + G, ... in order per thread, then a shared-memory tree; the host adds the
// G partials in order.  f selects the summand: 0 = x (x > 0), 1 = (x - c)^2 (x > 0), 2 = x (flag set)
__global__ void __launch_bounds__(256) k_partial_sums(const double* __restrict__ x, const uint8_t* __restrict__ flag, int64_t n, int f,
                                                      double c, double* __restrict__ partial) {
    __shared__ double s[256];
    double acc = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) {
        const double v = x[i];
        if (f == 0) { if (v > 0.0) acc += v; }
        else if (f == 1) { if (v > 0.0) acc += (v - c) * (v - c); }
        else { if (flag[i]) acc += v; }
    }
    s[threadIdx.x] = acc;
    __syncthreads();
    for (int d = 128; d > 0; d >>= 1) {
        if (threadIdx.x < d) s[threadIdx.x] += s[threadIdx.x + d];
        __syncthreads();
    }
    if (threadIdx.x == 0) partial[blockIdx.x] = s[0];
}

double fixed_sum(Ctx& cx, const double* d_x, const uint8_t* d_flag, int64_t n, int f, double c) {
    constexpr int G = 1024;
    DBuf<double> partial(G, cx.st);
    k_partial_sums<<<G, 256, 0, cx.st>>>(d_x, d_flag, n, f, c, partial.p);
    cx.launches++;
    auto hp = partial.to_host();
    double s = 0.0;
    for (int i = 0; i < G; ++i) s += hp[i];
    return s;
}

// filter 1: 0 < avg < thr (statistical); filter 2: count > thr (radius, thr = nb_points)
__global__ void __launch_bounds__(256) k_keep_flags(const double* __restrict__ avg, int64_t n, double thr, int filter,
                                                    uint8_t* __restrict__ keep, int32_t* __restrict__ keep_i) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const bool k = filter == 1 ? (avg[i] > 0.0 && avg[i] < thr) : (avg[i] > thr);
    keep[i] = k ? 1 : 0;
    keep_i[i] = k ? 1 : 0;
}

__global__ void __launch_bounds__(256) k_compact(const double* __restrict__ V, const uint8_t* __restrict__ keep,
                                                 const int32_t* __restrict__ off, int64_t n, double* __restrict__ out) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || !keep[i]) return;
    const size_t o = (size_t)off[i];
    out[3 * o] = V[3 * i];
    out[3 * o + 1] = V[3 * i + 1];
    out[3 * o + 2] = V[3 * i + 2];
}

void run_knn(Ctx& cx, const BrickMap& map, const BrickBitmap& bm, const int32_t* row_brick, const double* V, const uint8_t* alive,
             int64_t rows, int k, const PrepGrid& g, int max_ring, int mode, double* out, int64_t* n_far_out) {
    cudaStream_t st = cx.st;
    DBuf<int32_t> far_list(rows, st), n_far(1, st);
    n_far.zero();
    k_knn<<<div_up(rows, 128), 128, 0, st>>>(map.bricks.p, map.bnbr.p, map.view(), row_brick, V, alive, rows, k, g, mode, out,
                                              far_list.p, n_far.p);
    k_knn_far<<<kNumSMs * 4, 128, 0, st>>>(map.bricks.p, map.view(), row_brick, V, alive, k, g, bm, max_ring, mode, far_list.p, n_far.p,
                                            out);
    cx.launches += 2;
    if (n_far_out) *n_far_out = n_far.to_host()[0];
}

int bits_for(int64_t n) {
    int b = 1;
    while ((1ll << b) < n) ++b;
    return b;
}

}  // namespace

__global__ void __launch_bounds__(256) k_scale_to_f32(const double* __restrict__ in, double scale, float* __restrict__ out, int64_t n) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = __double2float_rn(__dmul_rn(in[i], scale));       // points * scale (:922), then .astype(float32)
}

void prepared_to_f32(Ctx& cx, const PrepState& ps, float* d_out) {
    k_scale_to_f32<<<div_up(ps.n_kept * 3, 256), 256, 0, cx.st>>>(ps.pts.p, ps.scale, d_out, ps.n_kept * 3);
    cx.launches++;
}

void prepare_cloud(Ctx& cx, const double* h_xyz, int64_t n, const SqsmPrepConfig& cfg, PrepState& ps, SqsmPrepStats& out) {
    cudaStream_t st = cx.st;
    SQSM_CHECK(h_xyz && n > 0 && n < (1ll << 31), "bad point array");
    SQSM_CHECK(cfg.voxel_down_size > 0.0, "voxel_down_size must be positive");
    SQSM_CHECK(cfg.filter >= 0 && cfg.filter <= 2, "filter must be 0 (none), 1 (statistical_outlier_removal) or 2 (radius_outlier_removal)");
    SQSM_CHECK(cfg.filter != 1 || (cfg.nb_neighbors >= 1 && cfg.nb_neighbors <= kKnnMax), "nb_neighbors must be 1..32");
    SQSM_CHECK(cfg.filter != 2 || (cfg.radius > 0.0 && cfg.nb_points >= 0), "radius_outlier_removal needs radius > 0 and nb_points >= 0");
    memset(&out, 0, sizeof(out));
    out.n_in = n;
    const int64_t launches0 = cx.launches;
    cudaEvent_t ev[4];
    for (auto& e : ev) SQSM_CUDA(cudaEventCreate(&e));
    SQSM_CUDA(cudaEventRecord(ev[0], st));

    // ---- bounds, shift, translate (:882-888)
    DBuf<double> P(n * 3, st);
    SQSM_CUDA(cudaMemcpyAsync(P.p, h_xyz, sizeof(double) * n * 3, cudaMemcpyHostToDevice, st));
    DBuf<unsigned long long> bnd(6, st);
    DBuf<int32_t> err(2, st);
    err.zero();
    const unsigned long long init[6] = {~0ull, ~0ull, ~0ull, 0, 0, 0};
    SQSM_CUDA(cudaMemcpyAsync(bnd.p, init, sizeof(init), cudaMemcpyHostToDevice, st));
    k_bounds_f64<<<(int)std::min<int64_t>(div_up(n, 256), kNumSMs * 8), 256, 0, st>>>(P.p, n, bnd.p, err.p);
    cx.launches++;
    auto hb = bnd.to_host();
    auto he = err.to_host();
    SQSM_CHECK(he[0] == 0, "point cloud holds NaN or infinite coordinates");
    double mn[3], mx[3];
    for (int a = 0; a < 3; ++a) { mn[a] = ord2d(hb[a]); mx[a] = ord2d(hb[3 + a]); }
    Shift3 sh;
    sh.s[0] = -(mn[0] + mx[0]) / 2.0;
    sh.s[1] = -(mn[1] + mx[1]) / 2.0;
    sh.s[2] = -mn[2];
    for (int a = 0; a < 3; ++a) out.global_shift[a] = sh.s[a];
    k_translate<<<div_up(n, 256), 256, 0, st>>>(P.p, n, sh);
    cx.launches++;

    // ---- voxel_down_sample (:889)
    PrepGrid g;
    for (int a = 0; a < 3; ++a) g.vmin[a] = (mn[a] + sh.s[a]) - cfg.voxel_down_size * 0.5;       // rounding is monotone: min(p + s) = min(p) + s
    g.voxel = cfg.voxel_down_size;
    BrickMap map;
    DBuf<double> V;
    int64_t rows = 0;
    {
        DBuf<uint64_t> ent_key(n, st);
        DBuf<uint16_t> ent_pos(n, st);
        DBuf<int32_t> ent_brick(n, st);
        k_prep_entries<<<div_up(n, 256), 256, 0, st>>>(P.p, n, g, ent_key.p, ent_pos.p, err.p + 1);
        cx.launches++;
        he = err.to_host();
        SQSM_CHECK(he[1] == 0, "cloud extent exceeds 524288 voxels per axis");
        brickmap_build(cx, map, ent_key.p, ent_pos.p, n, ent_brick.p);
        rows = map.n_rows;
        DBuf<uint32_t> row(n, st), idx(n, st), row_s(n, st), idx_s(n, st);
        k_prep_rows<<<div_up(n, 256), 256, 0, st>>>(map.bricks.p, ent_brick.p, ent_pos.p, n, row.p, idx.p);
        sort_pairs_u32(cx, row.p, row_s.p, idx.p, idx_s.p, n, 0, bits_for(rows));
        DBuf<int32_t> start(rows, st);
        k_row_starts<<<div_up(n, 256), 256, 0, st>>>(row_s.p, n, start.p);
        V.alloc(rows * 3, st);
        k_row_means<<<div_up(rows, 256), 256, 0, st>>>(P.p, idx_s.p, start.p, rows, n, V.p);
        cx.launches += 3;
    }
    P.release();
    out.n_voxel = rows;
    SQSM_CUDA(cudaEventRecord(ev[1], st));

    // ---- search structure shared by the two neighbour queries
    const bool need_nn = cfg.filter != 0 || cfg.density_adaptation;
    DBuf<int32_t> row_brick;
    DBuf<uint32_t> bitmap;
    BrickBitmap bm{nullptr, 0, 0, 0, 0};
    int max_ring = 2;
    if (need_nn) {
        brickmap_build_neighbours(cx, map);
        row_brick.alloc(rows, st);
        DBuf<uint16_t> row_pos(rows, st);
        brickmap_expand_rows(cx, map, row_brick.p, row_pos.p);
        double ext = 0;
        for (int a = 0; a < 3; ++a) ext = std::max(ext, mx[a] - mn[a]);
        max_ring = (int)std::ceil(ext / (8.0 * cfg.voxel_down_size)) + 2;
        // brick occupancy bitmap for the far search (skipped when the lattice would need more than 256 MB)
        const double edge = 8.0 * cfg.voxel_down_size;
        int nb[3];
        for (int a = 0; a < 3; ++a) nb[a] = (int)std::min(65536.0, std::floor((mx[a] - mn[a] + cfg.voxel_down_size) / edge) + 2.0);
        const int wx = (nb[0] + 31) / 32;
        const int64_t words = (int64_t)nb[2] * nb[1] * wx;
        const char* no_bm = getenv("SQSM_PREP_NO_BITMAP");                 // test hook: force the hash-probe ring search
        if (words <= (64ll << 20) && !(no_bm && no_bm[0] == '1')) {
            bitmap.alloc(words, st);
            bitmap.zero();
            k_brick_bitmap<<<div_up(map.n_bricks, 256), 256, 0, st>>>(map.bricks.p, map.n_bricks, nb[1], wx, bitmap.p);
            cx.launches++;
            bm = BrickBitmap{bitmap.p, nb[0], nb[1], nb[2], wx};
        }
    }

    // ---- remove_statistical_outlier (:895-898)
    DBuf<uint8_t> keep;
    ps.avg.release();
    ps.keep.release();
    int64_t n_kept = rows;
    if (cfg.filter != 0) {
        double thr;
        ps.avg.alloc(rows, st);
        if (cfg.filter == 1) {
            const int k = (int)std::min<int64_t>(cfg.nb_neighbors, rows);
            run_knn(cx, map, bm, row_brick.p, V.p, nullptr, rows, k, g, max_ring, 0, ps.avg.p, &out.n_far_filter);
            const double cloud_mean = fixed_sum(cx, ps.avg.p, nullptr, rows, 0, 0.0) / (double)rows;
            const double sq = fixed_sum(cx, ps.avg.p, nullptr, rows, 1, cloud_mean);
            const double sd = rows > 1 ? std::sqrt(sq / (double)(rows - 1)) : 0.0;
            thr = cloud_mean + cfg.std_ratio * sd;
            out.cloud_mean = cloud_mean;
            out.std_dev = sd;
            out.threshold = thr;
        } else {                                             // remove_radius_outlier (:899-902)
            const int rings = (int)std::ceil(cfg.radius / (8.0 * cfg.voxel_down_size));
            SQSM_CHECK(rings <= 64, "radius_outlier_removal: radius above 64 brick edges");
            k_radius_count<<<div_up(rows, 128), 128, 0, st>>>(map.bricks.p, map.bnbr.p, map.view(), row_brick.p, V.p, rows, g,
                                                               cfg.radius * cfg.radius, rings, ps.avg.p);
            cx.launches++;
            thr = (double)cfg.nb_points;
            out.threshold = thr;
        }
        keep.alloc(rows, st);
        DBuf<int32_t> keep_i(rows + 1, st), off(rows + 1, st);
        keep_i.zero();
        k_keep_flags<<<div_up(rows, 256), 256, 0, st>>>(ps.avg.p, rows, thr, cfg.filter, keep.p, keep_i.p);
        exclusive_scan_i32(cx, keep_i.p, off.p, rows + 1);
        int32_t total = 0;
        SQSM_CUDA(cudaMemcpyAsync(&total, off.p + rows, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
        SQSM_CUDA(cudaStreamSynchronize(st));
        n_kept = total;
        SQSM_CHECK(n_kept > 0, "the outlier filter removed every point");
        ps.pts.alloc(n_kept * 3, st);
        k_compact<<<div_up(rows, 256), 256, 0, st>>>(V.p, keep.p, off.p, rows, ps.pts.p);
        cx.launches += 2;
    } else {
        ps.pts.alloc(rows * 3, st);
        SQSM_CUDA(cudaMemcpyAsync(ps.pts.p, V.p, sizeof(double) * rows * 3, cudaMemcpyDeviceToDevice, st));
    }
    out.n_kept = n_kept;
    SQSM_CUDA(cudaEventRecord(ev[2], st));

    // ---- density adaptation (:908-913)
    double scale = 1.0;
    if (cfg.density_adaptation) {
        SQSM_CHECK(n_kept >= 2, "density adaptation needs at least two points");
        DBuf<double> nn(rows, st);
        run_knn(cx, map, bm, row_brick.p, V.p, cfg.filter != 0 ? keep.p : nullptr, rows, 2, g, max_ring, 1, nn.p, &out.n_far_scale);
        const double mean_nn = fixed_sum(cx, nn.p, nullptr, rows, 0, 0.0) / (double)n_kept;     // dropped rows hold -1, distances are > 0
        out.mean_nn = mean_nn;
        scale = cfg.voxel_down_size / mean_nn;
    }
    out.scale = scale;
    SQSM_CUDA(cudaEventRecord(ev[3], st));
    SQSM_CUDA(cudaStreamSynchronize(st));
    ps.vox.release();
    if (cfg.keep_debug) {
        ps.vox.alloc(rows * 3, st);
        SQSM_CUDA(cudaMemcpyAsync(ps.vox.p, V.p, sizeof(double) * rows * 3, cudaMemcpyDeviceToDevice, st));
        if (cfg.filter != 0) {
            ps.keep.alloc(rows, st);
            SQSM_CUDA(cudaMemcpyAsync(ps.keep.p, keep.p, rows, cudaMemcpyDeviceToDevice, st));
        }
        SQSM_CUDA(cudaStreamSynchronize(st));
    } else {
        ps.avg.release();
    }
    ps.n_voxel = rows;
    ps.n_kept = n_kept;
    ps.scale = scale;
    ps.valid = true;
    SQSM_CUDA(cudaEventElapsedTime(&out.ms_downsample, ev[0], ev[1]));
    SQSM_CUDA(cudaEventElapsedTime(&out.ms_filter, ev[1], ev[2]));
    SQSM_CUDA(cudaEventElapsedTime(&out.ms_scale, ev[2], ev[3]));
    SQSM_CUDA(cudaEventElapsedTime(&out.ms_total, ev[0], ev[3]));
    for (auto& e : ev) cudaEventDestroy(e);
    out.gpu_launches = cx.launches - launches0;
}

}  // namespace sqsm
