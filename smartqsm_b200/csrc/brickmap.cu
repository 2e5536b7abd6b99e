// BrickMap construction kernels (see brickmap.h).  sm_100a.
#include "brickmap.h"
#include <cub/cub.cuh>

namespace sqsm {

// ------------------------------------------------------------------ scan / sort wrappers (CUB)

void exclusive_scan_i32(Ctx& cx, const int32_t* d_in, int32_t* d_out, int64_t n) {
    if (n <= 0) return;
    size_t tb = 0;
    SQSM_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tb, d_in, d_out, (int)n, cx.st));
    DBuf<uint8_t> tmp((int64_t)tb + 16, cx.st);
    SQSM_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, tb, d_in, d_out, (int)n, cx.st));
    cx.launches += 2;
}

void exclusive_scan_i64(Ctx& cx, const int64_t* d_in, int64_t* d_out, int64_t n) {
    if (n <= 0) return;
    size_t tb = 0;
    SQSM_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tb, d_in, d_out, (int)n, cx.st));
    DBuf<uint8_t> tmp((int64_t)tb + 16, cx.st);
    SQSM_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, tb, d_in, d_out, (int)n, cx.st));
    cx.launches += 2;
}

void sort_keys_u64(Ctx& cx, const uint64_t* d_in, uint64_t* d_out, int64_t n, int begin_bit, int end_bit) {
    if (n <= 0) return;
    size_t tb = 0;
    SQSM_CUDA(cub::DeviceRadixSort::SortKeys(nullptr, tb, d_in, d_out, (int)n, begin_bit, end_bit, cx.st));
    DBuf<uint8_t> tmp((int64_t)tb + 16, cx.st);
    SQSM_CUDA(cub::DeviceRadixSort::SortKeys(tmp.p, tb, d_in, d_out, (int)n, begin_bit, end_bit, cx.st));
    cx.launches += 2 * ((end_bit - begin_bit + 7) / 8) + 1;
}

void sort_pairs_u32(Ctx& cx, const uint32_t* d_kin, uint32_t* d_kout, const uint32_t* d_vin, uint32_t* d_vout,
                    int64_t n, int begin_bit, int end_bit) {
    if (n <= 0) return;
    size_t tb = 0;
    SQSM_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tb, d_kin, d_kout, d_vin, d_vout, (int)n, begin_bit,
                                              end_bit, cx.st));
    DBuf<uint8_t> tmp((int64_t)tb + 16, cx.st);
    SQSM_CUDA(cub::DeviceRadixSort::SortPairs(tmp.p, tb, d_kin, d_kout, d_vin, d_vout, (int)n, begin_bit,
                                              end_bit, cx.st));
    cx.launches += 2 * ((end_bit - begin_bit + 7) / 8) + 1;
}

struct OffsetMinusBase {
    const int64_t* off; int64_t base;
    __host__ __device__ __forceinline__ int64_t operator()(int64_t i) const { return off[i] - base; }
};

void segmented_sort_u32(Ctx& cx, const uint32_t* d_in, uint32_t* d_out, int64_t n, int64_t n_segments, const int64_t* d_off,
                        int64_t base) {
    if (n <= 0 || n_segments <= 0) return;
    SQSM_CHECK(n < (1ll << 31) && n_segments < (1ll << 31), "segmented sort: more than 2^31 items");
    // segment s of this call = [d_off[s] - base, d_off[s + 1] - base): the offsets stay 64-bit and global, so a caller can
    // sort a list of more than 2^31 items chunk by chunk
    cub::CountingInputIterator<int64_t> cnt(0);
    cub::TransformInputIterator<int64_t, OffsetMinusBase, cub::CountingInputIterator<int64_t>> beg(cnt, OffsetMinusBase{d_off, base});
    size_t tb = 0;
    SQSM_CUDA(cub::DeviceSegmentedSort::SortKeys(nullptr, tb, d_in, d_out, (int)n, (int)n_segments, beg, beg + 1, cx.st));
    DBuf<uint8_t> tmp((int64_t)tb + 16, cx.st);
    SQSM_CUDA(cub::DeviceSegmentedSort::SortKeys(tmp.p, tb, d_in, d_out, (int)n, (int)n_segments, beg, beg + 1, cx.st));
    cx.launches += 3;
}

// ------------------------------------------------------------------ kernels

// Insert every entry's brick key; the thread that claims a slot appends the key to the list.
__global__ void __launch_bounds__(256) k_brick_insert(BrickHash h, const uint64_t* __restrict__ ent_key,
                                                      int64_t n_ent, uint64_t* __restrict__ list, int list_cap,
                                                      int32_t* __restrict__ counter) {
    int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    const uint64_t key = e < n_ent ? ent_key[e] : kEmptyKey;
    bool ins = false;
    if (key != kEmptyKey) {
        // consecutive entries often fall in the same brick: let one lane per distinct key probe (every lane probing for
        // itself: 252 instead of 212 us for 8.5 M entries)
        unsigned peers = __match_any_sync(__activemask(), key);
        if ((int)(__ffs(peers) - 1) == lane) {
            // (the overflow flag is polled inside long probe sequences only: once the table is known to be too small the
            // host retries with a larger one, and walking a FULL table costs a whole-table scan per insert)
            uint32_t slot = hash_insert(h, key, &ins, counter + 1);
            if (slot == 0xFFFFFFFFu) { counter[1] = 1; ins = false; }      // table too small: the host retries
        }
    }
    // one append-cursor atomic per warp (the cursor is a single address: per-key atomics serialise in L2)
    const unsigned won = __ballot_sync(0xFFFFFFFFu, ins);
    if (won == 0) return;
    int base = 0;
    if (lane == __ffs(won) - 1) base = atomicAdd(counter, __popc(won));
    base = __shfl_sync(0xFFFFFFFFu, base, __ffs(won) - 1);
    if (ins) {
        const int idx = base + __popc(won & ((1u << lane) - 1u));
        if (idx < list_cap) list[idx] = key; else counter[1] = 1;
    }
}

// Brick i takes the i-th sorted key; publish the id in the hash.
__global__ void __launch_bounds__(256) k_brick_init(BrickHash h, const uint64_t* __restrict__ sorted,
                                                    int n_bricks, Brick* __restrict__ bricks) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_bricks) return;
    uint64_t key = sorted[i];
    bricks[i].key = key;
    uint32_t slot = (uint32_t)mix64(key) & h.cap_mask;
    while (h.keys[slot] != key) slot = (slot + 1) & h.cap_mask;
    h.vals[slot] = i;
}

__global__ void __launch_bounds__(256) k_brick_setbits(BrickHash h, const uint64_t* __restrict__ ent_key,
                                                       const uint16_t* __restrict__ ent_pos, int64_t n_ent,
                                                       Brick* __restrict__ bricks, int32_t* __restrict__ ent_brick) {
    int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n_ent) return;
    uint64_t key = ent_key[e];
    if (key == kEmptyKey) { ent_brick[e] = -1; return; }
    int b = hash_find(h, key);
    ent_brick[e] = b;
    uint32_t pos = ent_pos[e];
    atomicOr(&bricks[b].mask[pos >> 5], 1u << (pos & 31u));
}

// 16 lanes per brick: popcount per word, exclusive prefix inside the half warp.
__global__ void __launch_bounds__(256) k_brick_prefix(Brick* __restrict__ bricks, int n_bricks,
                                                      int32_t* __restrict__ counts) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    int b = t >> 4, w = t & 15;
    bool live = b < n_bricks;
    uint32_t m = live ? bricks[b].mask[w] : 0u;
    int c = __popc(m);
    int incl = c;
    #pragma unroll
    for (int d = 1; d < 16; d <<= 1) {
        int v = __shfl_up_sync(0xFFFFFFFFu, incl, d, 16);
        if (w >= d) incl += v;
    }
    if (!live) return;
    bricks[b].prefix[w] = (uint16_t)(incl - c);
    if (w == 15) { bricks[b].count = incl; counts[b] = incl; }
}

__global__ void __launch_bounds__(256) k_brick_setbase(Brick* __restrict__ bricks, int n_bricks,
                                                       const int32_t* __restrict__ base) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_bricks) bricks[i].base = base[i];
}

__global__ void __launch_bounds__(256) k_brick_neighbours(BrickHash h, const Brick* __restrict__ bricks,
                                                          int n_bricks, int32_t* __restrict__ bnbr) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    int b = t >> 5, j = t & 31;
    if (b >= n_bricks) return;
    int id = -1;
    if (j < 27) {
        uint64_t key = bricks[b].key;
        int z = (int)key_z(key) + j / 9 - 1, y = (int)key_y(key) + (j / 3) % 3 - 1, x = (int)key_x(key) + j % 3 - 1;
        if (j == 13) id = b;
        else if (z >= 0 && y >= 0 && x >= 0 && z < 65536 && y < 65536 && x < 65536)
            id = hash_find(h, pack_key(key_s(key), (uint32_t)z, (uint32_t)y, (uint32_t)x));
    }
    bnbr[(size_t)b * 32 + j] = id;
}

__global__ void __launch_bounds__(256) k_brick_expand(const Brick* __restrict__ bricks, int n_bricks,
                                                      int32_t* __restrict__ row_brick, uint16_t* __restrict__ row_pos) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    int b = t >> 4, w = t & 15;
    if (b >= n_bricks) return;
    uint32_t m = bricks[b].mask[w];
    int row = bricks[b].base + (int)bricks[b].prefix[w];
    while (m) {
        int bit = __ffs(m) - 1;
        m &= m - 1;
        row_brick[row] = b;
        row_pos[row] = (uint16_t)((w << 5) | bit);
        ++row;
    }
}

// ------------------------------------------------------------------ host side

static uint32_t next_pow2(uint64_t v) {
    uint64_t p = 1024;
    while (p < v) p <<= 1;
    SQSM_CHECK(p <= (1ull << 31), "brick hash too large");
    return (uint32_t)p;
}

void brickmap_build(Ctx& cx, BrickMap& m, const uint64_t* d_ent_key, const uint16_t* d_ent_pos, int64_t n_ent,
                    int32_t* d_ent_brick, int64_t expected_bricks) {
    cudaStream_t st = cx.st;
    Trace tr(cx, "  brickmap_build");
    SQSM_CHECK(n_ent < (1ll << 31), "too many entries for one BrickMap build");
    m.n_bricks = 0;
    m.n_rows = 0;
    if (n_ent == 0) return;
    // Expect far fewer bricks than entries (a brick holds up to 512 voxels): start with a table of four slots per expected
    // brick (callers pass what they know: about n_ent / 10 bricks for a 1 cm grid over a surface scan, a quarter of the finer
    // level's bricks for a down-sampled level; default n_ent / 8) so that keys + values stay L2 resident, and retry with 4x
    // if the table exceeds 50 % load.
    DBuf<uint64_t> list;
    DBuf<int32_t> counter(2, st);
    int grid = div_up(n_ent, 256);
    int32_t nb = 0;
    if (expected_bricks <= 0) expected_bricks = n_ent / 8;
    uint64_t want = std::max<uint64_t>(4 * (uint64_t)std::min<int64_t>(expected_bricks, n_ent), 4096);
    for (int attempt = 0;; ++attempt) {
        m.cap = next_pow2(want);
        const int list_cap = (int)std::min<uint64_t>((uint64_t)m.cap / 2, (uint64_t)n_ent);
        m.hkeys.alloc(m.cap, st);
        m.hvals.alloc(m.cap, st);
        m.hkeys.fill_ff();
        list.alloc(list_cap, st);
        counter.zero();
        k_brick_insert<<<grid, 256, 0, st>>>(m.view(), d_ent_key, n_ent, list.p, list_cap, counter.p);
        cx.launches++;
        int32_t hc[2];
        SQSM_CUDA(cudaMemcpyAsync(hc, counter.p, sizeof(hc), cudaMemcpyDeviceToHost, st));
        SQSM_CUDA(cudaStreamSynchronize(st));
        nb = hc[0];
        if (hc[1] == 0) break;
        SQSM_CHECK(attempt < 4, "brick hash overflow");
        want = (uint64_t)m.cap * 4;
    }
    m.n_bricks = nb;
    if (nb == 0) return;
    DBuf<uint64_t> sorted(nb, st);
    sort_keys_u64(cx, list.p, sorted.p, nb, 0, 64);
    m.bricks.alloc(nb, st);
    m.bricks.zero();
    k_brick_init<<<div_up(nb, 256), 256, 0, st>>>(m.view(), sorted.p, nb, m.bricks.p);
    k_brick_setbits<<<grid, 256, 0, st>>>(m.view(), d_ent_key, d_ent_pos, n_ent, m.bricks.p, d_ent_brick);
    DBuf<int32_t> counts((int64_t)nb + 1, st);
    DBuf<int32_t> base((int64_t)nb + 1, st);
    counts.zero();
    k_brick_prefix<<<div_up((int64_t)nb * 16, 256), 256, 0, st>>>(m.bricks.p, nb, counts.p);
    exclusive_scan_i32(cx, counts.p, base.p, (int64_t)nb + 1);
    k_brick_setbase<<<div_up(nb, 256), 256, 0, st>>>(m.bricks.p, nb, base.p);
    cx.launches += 4;
    int32_t total = 0;
    SQSM_CUDA(cudaMemcpyAsync(&total, base.p + nb, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    SQSM_CUDA(cudaStreamSynchronize(st));
    m.n_rows = total;
}

void brickmap_build_neighbours(Ctx& cx, BrickMap& m) {
    if (m.n_bricks == 0) return;
    m.bnbr.alloc((int64_t)m.n_bricks * 32, cx.st);
    k_brick_neighbours<<<div_up((int64_t)m.n_bricks * 32, 256), 256, 0, cx.st>>>(m.view(), m.bricks.p, m.n_bricks,
                                                                                m.bnbr.p);
    cx.launches++;
}

void brickmap_expand_rows(Ctx& cx, const BrickMap& m, int32_t* d_row_brick, uint16_t* d_row_pos) {
    if (m.n_bricks == 0) return;
    k_brick_expand<<<div_up((int64_t)m.n_bricks * 16, 256), 256, 0, cx.st>>>(m.bricks.p, m.n_bricks, d_row_brick,
                                                                             d_row_pos);
    cx.launches++;
}

}  // namespace sqsm
