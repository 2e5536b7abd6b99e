// Connectivity graph and skeleton extraction (SURVEY.md section 8 row a16 = f1).  sm_100a + host C++.
//
// Replaces SkeletonizationPipelineBase._construct_graph (core/skeletonization_pipeline_base.py:75-174) and
// ._extract_skeleton with topology_extractor == "spt" (:190-215): networkx graphs built edge by edge in Python, two
// Dijkstra runs, one Kruskal run and an np.unique over Python tuples.  What must be reproduced bit for bit is not only the
// arithmetic (float32 weights, float32 distance accumulation) but the ORDER networkx visits things in, because
// `predecessors[0]` is "whoever reached the final distance first":
//   * nx.Graph keeps nodes and adjacency entries in insertion order = first appearance in the edge stream;
//   * Dijkstra pops (distance, push counter); ties in the float32 distance go to the earlier push;
//   * Kruskal sorts G.edges (node order, then adjacency order, each edge reported by its earlier endpoint) stably by weight.
// Split of the work: everything data parallel runs on the device - expanding the tetrahedra into the edge stream,
// de-duplicating it with "first occurrence wins" (stable radix sort), node ranks, insertion-ordered CSR adjacency, float32
// weights, the Kruskal sort key, the merge / unique of the four edge sources (:158) and the final CSR; the two
// priority-queue loops and the union-find, whose semantics are sequential by definition, run in C++ on the host over
// those arrays.  The Delaunay tetrahedra themselves come from Qhull through SciPy, exactly like in the reference (:77):
// on the 1 cm voxel lattice most point sets are co-spherical and the result depends on Qhull's own choices.
#include "engine.h"
#include <cub/cub.cuh>
#include <algorithm>
#include <cmath>
#include <cstring>
#include <functional>
#include <limits>
#include <queue>

namespace sqsm {

// ---------------------------------------------------------------------------------- device pieces

// the 6 edges of tetrahedron i, in the reference's block order (:80-95): stream index e = block * T + i
__global__ void __launch_bounds__(256) k_tet_edges(const int32_t* __restrict__ simp, int64_t T, uint64_t* __restrict__ key,
                                                   uint32_t* __restrict__ val, uint32_t* __restrict__ first_pos) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= 6 * T) return;
    const int b = (int)(e / T);
    const int64_t i = e % T;
    const int ia[6] = {0, 0, 0, 1, 1, 2}, ib[6] = {1, 2, 3, 2, 3, 3};
    const uint32_t u = (uint32_t)simp[4 * i + ia[b]], v = (uint32_t)simp[4 * i + ib[b]];
    key[e] = ((uint64_t)min(u, v) << 32) | (uint64_t)max(u, v);
    val[e] = (uint32_t)e;
    // node insertion order of nx.Graph.add_weighted_edges_from: u of edge e at position 2e, v at 2e + 1
    atomicMin(&first_pos[u], (uint32_t)(2 * e));
    atomicMin(&first_pos[v], (uint32_t)(2 * e + 1));
}

__global__ void __launch_bounds__(256) k_head_flags(const uint64_t* __restrict__ key, int64_t n, int32_t* __restrict__ flag) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i > n) return;
    flag[i] = (i < n && (i == 0 || key[i] != key[i - 1])) ? 1 : 0;
}

__global__ void __launch_bounds__(256) k_compact_edges(const uint64_t* __restrict__ key, const uint32_t* __restrict__ val,
                                                       const int32_t* __restrict__ flag, const int32_t* __restrict__ scan, int64_t n,
                                                       uint64_t* __restrict__ ukey, uint32_t* __restrict__ uval) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || !flag[i]) return;
    ukey[scan[i]] = key[i];
    if (uval) uval[scan[i]] = val[i];
}

// "L2" weight of the Delaunay graph (:97-100): np.linalg.norm((a - b) ** 2, axis=1) in float32, no FMA
__device__ __forceinline__ float w_delaunay(const float* __restrict__ P, uint32_t u, uint32_t v) {
    float s = 0.f;
    #pragma unroll
    for (int a = 0; a < 3; ++a) {
        const float d = __fsub_rn(P[3 * (size_t)u + a], P[3 * (size_t)v + a]);
        const float q = __fmul_rn(d, d);
        const float q2 = __fmul_rn(q, q);
        s = (a == 0) ? q2 : __fadd_rn(s, q2);
    }
    return __fsqrt_rn(s);
}
// "l1" weight (:62): np.linalg.norm(a - b) of ONE vector = sqrt(sdot): float32 products accumulated in a C double
// (OpenBLAS tail loop), rounded to float32, float32 sqrt.  "l2" (:63) squares it in float32.
__device__ __forceinline__ float w_l1(const float* __restrict__ P, uint32_t u, uint32_t v) {
    double s = 0.0;
    #pragma unroll
    for (int a = 0; a < 3; ++a) {
        const float d = __fsub_rn(P[3 * (size_t)u + a], P[3 * (size_t)v + a]);
        s = __dadd_rn(s, (double)__fmul_rn(d, d));
    }
    return __fsqrt_rn(__double2float_rn(s));
}

// both directions of every unique edge: key = (tail node, order inside the tail's adjacency list), value = edge id
__global__ void __launch_bounds__(256) k_arcs(const uint64_t* __restrict__ ukey, const uint32_t* __restrict__ order, int64_t n,
                                              uint64_t* __restrict__ akey, uint32_t* __restrict__ aval) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t u = (uint32_t)(ukey[i] >> 32), v = (uint32_t)ukey[i];
    const uint32_t o = order ? order[i] : 0u;
    akey[2 * i] = ((uint64_t)u << 32) | (order ? o : v);          // no explicit order: ascending neighbour id
    aval[2 * i] = v;
    akey[2 * i + 1] = ((uint64_t)v << 32) | (order ? o : u);
    aval[2 * i + 1] = u;
}

__global__ void __launch_bounds__(256) k_arc_rows(const uint64_t* __restrict__ akey, int64_t n, int64_t m, int64_t* __restrict__ off) {
    // off[x] = first arc whose tail is >= x (arcs sorted by tail); off[m] = n
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i > n) return;
    const int64_t cur = (i < n) ? (int64_t)(akey[i] >> 32) : m;
    const int64_t prev = (i == 0) ? -1 : (int64_t)(akey[i - 1] >> 32);
    for (int64_t x = prev + 1; x <= cur; ++x) off[x] = i;
}

template <int WHICH>
__global__ void __launch_bounds__(256) k_arc_weights(const float* __restrict__ P, const uint64_t* __restrict__ akey,
                                                     const uint32_t* __restrict__ head, int64_t n, float* __restrict__ w) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t t = (uint32_t)(akey[i] >> 32), h = head[i];
    float x;
    if (WHICH == 0) x = w_delaunay(P, t, h);
    else { x = w_l1(P, t, h); if (WHICH == 2) x = __fmul_rn(x, x); }
    w[i] = x;
}

// Kruskal order of networkx (mst.py kruskal_mst_edges): stable sort by weight of G.edges(), which lists every edge at its
// endpoint that comes first in node order, nodes in insertion order, neighbours in adjacency (first-occurrence) order.
__global__ void __launch_bounds__(256) k_kruskal_keys(const float* __restrict__ P, const uint64_t* __restrict__ ukey,
                                                      const uint32_t* __restrict__ first_pos, int64_t n, uint32_t* __restrict__ wbits,
                                                      uint32_t* __restrict__ owner_pos) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t u = (uint32_t)(ukey[i] >> 32), v = (uint32_t)ukey[i];
    wbits[i] = __float_as_uint(w_delaunay(P, u, v));             // non-negative floats order like their bit patterns
    owner_pos[i] = min(first_pos[u], first_pos[v]);
}

__global__ void __launch_bounds__(256) k_iota32(uint32_t* __restrict__ a, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) a[i] = (uint32_t)i;
}
__global__ void __launch_bounds__(256) k_gather_u32(const uint32_t* __restrict__ src, const uint32_t* __restrict__ idx, int64_t n,
                                                    uint32_t* __restrict__ dst) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = src[idx[i]];
}
__global__ void __launch_bounds__(256) k_gather_u64(const uint64_t* __restrict__ src, const uint32_t* __restrict__ idx, int64_t n,
                                                    uint64_t* __restrict__ dst) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = src[idx[i]];
}
__global__ void __launch_bounds__(256) k_pairs_to_keys(const int64_t* __restrict__ e, int64_t n, uint64_t* __restrict__ key) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint64_t a = (uint64_t)e[2 * i], b = (uint64_t)e[2 * i + 1];
    key[i] = (min(a, b) << 32) | max(a, b);                      // np.sort(axis=1) (:158)
}
__global__ void __launch_bounds__(256) k_keys_to_pairs(const uint64_t* __restrict__ key, int64_t n, int64_t* __restrict__ e) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    e[2 * i] = (int64_t)(key[i] >> 32);
    e[2 * i + 1] = (int64_t)(key[i] & 0xFFFFFFFFull);
}
// node insertion order of the final graph (add_edges_from over the lexicographic list): u at 2e, v at 2e + 1
__global__ void __launch_bounds__(256) k_first_pos(const uint64_t* __restrict__ key, int64_t n, uint32_t* __restrict__ first_pos) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n) return;
    atomicMin(&first_pos[(uint32_t)(key[e] >> 32)], (uint32_t)(2 * e));
    atomicMin(&first_pos[(uint32_t)key[e]], (uint32_t)(2 * e + 1));
}

static void sort_pairs_64(Ctx& cx, DBuf<uint64_t>& k, DBuf<uint32_t>& v, int64_t n, int end_bit = 64) {
    if (n <= 0) return;
    SQSM_CHECK(n < (1ll << 31), "graph: more than 2^31 items in one sort");
    DBuf<uint64_t> k2(n, cx.st);
    DBuf<uint32_t> v2(n, cx.st);
    size_t tb = 0;
    SQSM_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tb, k.p, k2.p, v.p, v2.p, (int)n, 0, end_bit, cx.st));
    DBuf<uint8_t> tmp((int64_t)tb + 16, cx.st);
    SQSM_CUDA(cub::DeviceRadixSort::SortPairs(tmp.p, tb, k.p, k2.p, v.p, v2.p, (int)n, 0, end_bit, cx.st));
    cx.launches += 2 * ((end_bit + 7) / 8) + 1;
    k = std::move(k2);
    v = std::move(v2);
}
static void sort_pairs_32(Ctx& cx, DBuf<uint32_t>& k, DBuf<uint32_t>& v, int64_t n) {
    if (n <= 0) return;
    DBuf<uint32_t> k2(n, cx.st), v2(n, cx.st);
    sort_pairs_u32(cx, k.p, k2.p, v.p, v2.p, n, 0, 32);
    k = std::move(k2);
    v = std::move(v2);
}

// sorted keys (+ values) -> unique keys keeping the FIRST value of every run (stable sort: the smallest stream index)
static int64_t unique_first(Ctx& cx, DBuf<uint64_t>& key, DBuf<uint32_t>* val, int64_t n) {
    if (n <= 0) return 0;
    cudaStream_t st = cx.st;
    DBuf<int32_t> flag(n + 1, st), scan(n + 1, st);
    k_head_flags<<<div_up(n + 1, 256), 256, 0, st>>>(key.p, n, flag.p);
    cx.launches++;
    exclusive_scan_i32(cx, flag.p, scan.p, n + 1);
    int32_t nu = 0;
    SQSM_CUDA(cudaMemcpyAsync(&nu, scan.p + n, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    SQSM_CUDA(cudaStreamSynchronize(st));
    DBuf<uint64_t> uk(nu, st);
    DBuf<uint32_t> uv;
    if (val) uv.alloc(nu, st);
    k_compact_edges<<<div_up(n, 256), 256, 0, st>>>(key.p, val ? val->p : nullptr, flag.p, scan.p, n, uk.p, val ? uv.p : nullptr);
    cx.launches++;
    key = std::move(uk);
    if (val) *val = std::move(uv);
    return nu;
}

// CSR adjacency on the host: arcs sorted by (tail, insertion order)
struct HostCsr {
    std::vector<int64_t> off;
    std::vector<uint32_t> head;
    std::vector<float> w;
};

static void build_csr(Ctx& cx, const float* d_P, int64_t m, const DBuf<uint64_t>& ukey, const uint32_t* d_order, int64_t ne, int which,
                      HostCsr& out) {
    cudaStream_t st = cx.st;
    const int64_t na = 2 * ne;
    DBuf<uint64_t> akey(na, st);
    DBuf<uint32_t> aval(na, st);
    k_arcs<<<div_up(ne, 256), 256, 0, st>>>(ukey.p, d_order, ne, akey.p, aval.p);
    cx.launches++;
    sort_pairs_64(cx, akey, aval, na);
    DBuf<int64_t> off(m + 1, st);
    k_arc_rows<<<div_up(na + 1, 256), 256, 0, st>>>(akey.p, na, m, off.p);
    DBuf<float> w(na, st);
    if (which == 0) k_arc_weights<0><<<div_up(na, 256), 256, 0, st>>>(d_P, akey.p, aval.p, na, w.p);
    else if (which == 1) k_arc_weights<1><<<div_up(na, 256), 256, 0, st>>>(d_P, akey.p, aval.p, na, w.p);
    else k_arc_weights<2><<<div_up(na, 256), 256, 0, st>>>(d_P, akey.p, aval.p, na, w.p);
    cx.launches += 2;
    out.off = off.to_host();
    out.head = aval.to_host();
    out.w = w.to_host();
}

// ---------------------------------------------------------------------------------- host pieces (networkx semantics)

// connected components enumerated like nx.connected_components (`for v in G`, i.e. by node insertion order); returns the
// first component of maximal size (max(..., key=len)) as a node mask, and its root = lowest z, ties to the lowest index
// (np.argmin over list(set): CPython iterates a set of small non-negative ints in ascending order)
static int64_t largest_component_root(const HostCsr& g, const std::vector<uint32_t>& first_pos, const std::vector<float>& P,
                                      int64_t m, std::vector<uint8_t>& in_comp) {
    std::vector<int32_t> comp(m, -1);
    std::vector<int64_t> size;
    std::vector<uint32_t> start_pos;                       // insertion position of the component's earliest node
    std::vector<int64_t> stack;
    for (int64_t s = 0; s < m; ++s) {
        if (comp[s] >= 0 || first_pos[s] == 0xFFFFFFFFu) continue;          // not a node of this graph
        const int32_t c = (int32_t)size.size();
        int64_t cnt = 0;
        uint32_t fp = 0xFFFFFFFFu;
        stack.push_back(s);
        comp[s] = c;
        while (!stack.empty()) {
            const int64_t v = stack.back();
            stack.pop_back();
            ++cnt;
            fp = std::min(fp, first_pos[v]);
            for (int64_t a = g.off[v]; a < g.off[v + 1]; ++a) {
                const uint32_t u = g.head[a];
                if (comp[u] < 0) { comp[u] = c; stack.push_back(u); }
            }
        }
        size.push_back(cnt);
        start_pos.push_back(fp);
    }
    SQSM_CHECK(!size.empty(), "graph has no node");
    int32_t best = 0;
    for (int32_t c = 1; c < (int32_t)size.size(); ++c)
        if (size[c] > size[best] || (size[c] == size[best] && start_pos[c] < start_pos[best])) best = c;
    in_comp.assign(m, 0);
    int64_t root = -1;
    for (int64_t v = 0; v < m; ++v) {
        if (comp[v] != best) continue;
        in_comp[v] = 1;
        if (root < 0 || P[3 * v + 2] < P[3 * root + 2]) root = v;
    }
    return root;
}

// networkx _dijkstra_multisource with one source and a predecessor dict (weighted.py): float32 distances, heap keyed by
// (distance, push counter), pred[u] = [v] whenever a strictly shorter tentative distance is found.  `order` receives the
// nodes in the order their predecessor entry was first created (insertion order of the pred dict).
static void dijkstra_nx(const HostCsr& g, int64_t m, int64_t root, std::vector<int64_t>& pred, std::vector<int64_t>* order) {
    struct Item { float d; uint64_t c; uint32_t v; };
    auto later = [](const Item& a, const Item& b) { return a.d > b.d || (a.d == b.d && a.c > b.c); };
    std::priority_queue<Item, std::vector<Item>, decltype(later)> heap(later);
    std::vector<uint8_t> done(m, 0), seen_set(m, 0);
    std::vector<float> seen(m, 0.f);
    pred.assign(m, -1);
    uint64_t counter = 0;
    seen_set[root] = 1;
    heap.push(Item{0.f, counter++, (uint32_t)root});
    while (!heap.empty()) {
        const Item it = heap.top();
        heap.pop();
        const uint32_t v = it.v;
        if (done[v]) continue;
        done[v] = 1;
        for (int64_t a = g.off[v]; a < g.off[v + 1]; ++a) {
            const uint32_t u = g.head[a];
            volatile float vu = it.d + g.w[a];                  // float32 addition, rounded once (no excess precision)
            if (done[u]) continue;
            if (!seen_set[u] || vu < seen[u]) {
                if (!seen_set[u] && order) order->push_back(u);
                seen_set[u] = 1;
                seen[u] = vu;
                heap.push(Item{vu, counter++, u});
                pred[u] = v;
            }
        }
    }
}

struct UnionFind {
    std::vector<int32_t> p;
    explicit UnionFind(int64_t n) : p(n) { for (int64_t i = 0; i < n; ++i) p[i] = (int32_t)i; }
    int32_t find(int32_t x) { while (p[x] != x) { p[x] = p[p[x]]; x = p[x]; } return x; }
};

}  // namespace sqsm

using namespace sqsm;

// ---------------------------------------------------------------------------------- entry points (called from engine.cu)

namespace sqsm {

// d_radius: sorted unique (i, j) pairs on the device (engine's radius graph) or nullptr
void graph_construct(Ctx& cx, GraphState& gs, const float* d_pts, int64_t m, const int32_t* h_simplices, int64_t T,
                     const int64_t* d_radius, int64_t n_radius, const int64_t* h_preset, int64_t n_preset) {
    cudaStream_t st = cx.st;
    SQSM_CHECK(m > 0 && m < (1ll << 31), "graph: bad node count");
    SQSM_CHECK(T > 0 && 6 * T < (1ll << 31), "graph: needs at least one Delaunay tetrahedron (and fewer than 2^31 / 6)");
    gs.valid = false;
    gs.m = m;
    gs.pts.alloc(m * 3, st);
    SQSM_CUDA(cudaMemcpyAsync(gs.pts.p, d_pts, sizeof(float) * 3 * m, cudaMemcpyDeviceToDevice, st));
    gs.h_pts = gs.pts.to_host();
    for (int64_t i = 0; i < 4 * T; ++i) SQSM_CHECK(h_simplices[i] >= 0 && h_simplices[i] < m, "graph: simplex vertex out of range");
    // ---- Delaunay edge stream -> unique edges with first occurrence, node insertion positions (:77-102)
    const int64_t S = 6 * T;
    DBuf<int32_t> simp(4 * T, st);
    SQSM_CUDA(cudaMemcpyAsync(simp.p, h_simplices, sizeof(int32_t) * 4 * T, cudaMemcpyHostToDevice, st));
    DBuf<uint64_t> key(S, st);
    DBuf<uint32_t> val(S, st), first_pos(m, st);
    first_pos.fill_ff();
    k_tet_edges<<<div_up(S, 256), 256, 0, st>>>(simp.p, T, key.p, val.p, first_pos.p);
    cx.launches++;
    sort_pairs_64(cx, key, val, S);
    const int64_t ne = unique_first(cx, key, &val, S);          // key: unique edges (lexicographic), val: first stream index
    gs.n_delaunay = ne;
    HostCsr g1;
    build_csr(cx, gs.pts.p, m, key, val.p, ne, /*which=*/0, g1);
    std::vector<uint32_t> h_first = first_pos.to_host();
    // ---- largest component, root, shortest-path tree (:104-120)
    std::vector<uint8_t> in_comp;
    const int64_t root = largest_component_root(g1, h_first, gs.h_pts, m, in_comp);
    gs.delaunay_root = root;
    std::vector<int64_t> pred;
    dijkstra_nx(g1, m, root, pred, nullptr);
    std::vector<uint64_t> extra;
    extra.reserve((size_t)(2 * m + n_preset));
    for (int64_t v = 0; v < m; ++v)
        if (pred[v] >= 0) extra.push_back(((uint64_t)std::min<int64_t>(pred[v], v) << 32) | (uint64_t)std::max<int64_t>(pred[v], v));
    gs.n_spt = (int64_t)extra.size();
    // ---- minimum spanning edges, Kruskal in networkx's edge order (:122-130)
    {
        DBuf<uint32_t> wbits(ne, st), opos(ne, st), idx(ne, st), tfirst(ne, st), tmp(ne, st);
        k_kruskal_keys<<<div_up(ne, 256), 256, 0, st>>>(gs.pts.p, key.p, first_pos.p, ne, wbits.p, opos.p);
        k_iota32<<<div_up(ne, 256), 256, 0, st>>>(idx.p, ne);
        cx.launches += 2;
        // stable LSD passes: first occurrence (position in the owner's adjacency list), owner's node position, weight
        SQSM_CUDA(cudaMemcpyAsync(tfirst.p, val.p, sizeof(uint32_t) * ne, cudaMemcpyDeviceToDevice, st));
        sort_pairs_32(cx, tfirst, idx, ne);
        k_gather_u32<<<div_up(ne, 256), 256, 0, st>>>(opos.p, idx.p, ne, tmp.p);
        sort_pairs_32(cx, tmp, idx, ne);
        DBuf<uint32_t> tmp2(ne, st);
        k_gather_u32<<<div_up(ne, 256), 256, 0, st>>>(wbits.p, idx.p, ne, tmp2.p);
        sort_pairs_32(cx, tmp2, idx, ne);
        DBuf<uint64_t> sorted(ne, st);
        k_gather_u64<<<div_up(ne, 256), 256, 0, st>>>(key.p, idx.p, ne, sorted.p);
        cx.launches += 3;
        std::vector<uint64_t> hs = sorted.to_host();
        UnionFind uf(m);
        int64_t n_mst = 0;
        for (int64_t i = 0; i < ne; ++i) {
            const int32_t a = uf.find((int32_t)(hs[i] >> 32)), b = uf.find((int32_t)(hs[i] & 0xFFFFFFFFull));
            if (a != b) { uf.p[a] = b; extra.push_back(hs[i]); ++n_mst; }
        }
        gs.n_mst = n_mst;
    }
    for (int64_t i = 0; i < n_preset; ++i) {
        const int64_t a = h_preset[2 * i], b = h_preset[2 * i + 1];
        SQSM_CHECK(a >= 0 && b >= 0 && a < m && b < m, "graph: preset edge out of range");
        extra.push_back(((uint64_t)std::min(a, b) << 32) | (uint64_t)std::max(a, b));
    }
    // ---- merge with the radius-neighbour edges: np.unique(np.sort(edges, axis=1), axis=0) (:158)
    const int64_t nx = (int64_t)extra.size(), total = nx + n_radius;
    DBuf<uint64_t> all(total, st);
    if (nx) SQSM_CUDA(cudaMemcpyAsync(all.p, extra.data(), sizeof(uint64_t) * nx, cudaMemcpyHostToDevice, st));
    if (n_radius) {
        k_pairs_to_keys<<<div_up(n_radius, 256), 256, 0, st>>>(d_radius, n_radius, all.p + nx);
        cx.launches++;
    }
    {
        SQSM_CHECK(total < (1ll << 31), "graph: more than 2^31 candidate edges");
        DBuf<uint64_t> s2(total, st);
        sort_keys_u64(cx, all.p, s2.p, total, 0, 64);
        all = std::move(s2);
    }
    SQSM_CUDA(cudaStreamSynchronize(st));                        // `extra` was read by the copy above
    gs.n_edges = unique_first(cx, all, nullptr, total);
    gs.merged = std::move(all);
    gs.valid = true;
}

void graph_edges(Ctx& cx, const GraphState& gs, int64_t* h_out) {
    SQSM_CHECK(gs.valid, "no graph: call sqsm_graph_construct first");
    if (gs.n_edges == 0) return;
    DBuf<int64_t> e(2 * gs.n_edges, cx.st);
    k_keys_to_pairs<<<div_up(gs.n_edges, 256), 256, 0, cx.st>>>(gs.merged.p, gs.n_edges, e.p);
    cx.launches++;
    SQSM_CUDA(cudaMemcpyAsync(h_out, e.p, sizeof(int64_t) * 2 * gs.n_edges, cudaMemcpyDeviceToHost, cx.st));
    SQSM_CUDA(cudaStreamSynchronize(cx.st));
}

// weighted graph (:160-174) + shortest-path-tree skeleton (:190-215)
void graph_extract_skeleton(Ctx& cx, const GraphState& gs, int weight_fn, int64_t* root_out, int64_t* n_reached, int64_t* h_pred,
                            int64_t* h_order, float* h_weight) {
    cudaStream_t st = cx.st;
    SQSM_CHECK(gs.valid && gs.n_edges > 0, "no graph: call sqsm_graph_construct first");
    SQSM_CHECK(weight_fn == 0 || weight_fn == 1, "edge_weight_function must be 0 (l1) or 1 (l2)");
    const int64_t m = gs.m;
    HostCsr g2;
    build_csr(cx, gs.pts.p, m, gs.merged, nullptr, gs.n_edges, weight_fn == 0 ? 1 : 2, g2);
    DBuf<uint32_t> first_pos(m, st);
    first_pos.fill_ff();
    k_first_pos<<<div_up(gs.n_edges, 256), 256, 0, st>>>(gs.merged.p, gs.n_edges, first_pos.p);
    cx.launches++;
    std::vector<uint32_t> h_first = first_pos.to_host();
    std::vector<uint8_t> in_comp;
    const int64_t root = largest_component_root(g2, h_first, gs.h_pts, m, in_comp);
    std::vector<int64_t> pred, order;
    order.reserve(m);
    dijkstra_nx(g2, m, root, pred, &order);
    if (root_out) *root_out = root;
    if (n_reached) *n_reached = (int64_t)order.size() + 1;
    if (h_pred) std::memcpy(h_pred, pred.data(), sizeof(int64_t) * m);
    if (h_order) std::memcpy(h_order, order.data(), sizeof(int64_t) * order.size());
    if (h_weight) {
        // weight of the skeleton edge (pred, node) (:209-213): always the plain norm, whatever the graph was weighted with
        for (int64_t v = 0; v < m; ++v) {
            float w = 0.f;
            if (pred[v] >= 0) {
                double s = 0.0;
                for (int a = 0; a < 3; ++a) {
                    volatile float d = gs.h_pts[3 * pred[v] + a] - gs.h_pts[3 * v + a];
                    volatile float q = d * d;
                    s += (double)q;
                }
                volatile float sf = (float)s;
                w = std::sqrt(sf);
            }
            h_weight[v] = w;
        }
    }
}

}  // namespace sqsm
