// Common device/host helpers for libsqsm_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <stdexcept>
#include <vector>

namespace sqsm {

struct CudaError : std::runtime_error {
    using std::runtime_error::runtime_error;
};

#define SQSM_CUDA(expr)                                                                          \
    do {                                                                                         \
        cudaError_t _e = (expr);                                                                 \
        if (_e != cudaSuccess) {                                                                 \
            char _b[512];                                                                        \
            snprintf(_b, sizeof(_b), "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e),     \
                     __FILE__, __LINE__);                                                        \
            throw ::sqsm::CudaError(_b);                                                         \
        }                                                                                        \
    } while (0)

#define SQSM_CHECK(cond, msg)                                                                    \
    do {                                                                                         \
        if (!(cond)) {                                                                           \
            char _b[512];                                                                        \
            snprintf(_b, sizeof(_b), "%s (%s:%d)", msg, __FILE__, __LINE__);                     \
            throw ::sqsm::CudaError(_b);                                                         \
        }                                                                                        \
    } while (0)

constexpr int kNumSMs = 148;            // B200: 2 dies x 74 SMs

inline int div_up(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

// Stream-ordered device buffer (cudaMallocAsync pool; freed on the same stream).
template <typename T>
struct DBuf {
    T* p = nullptr;
    int64_t n = 0;
    cudaStream_t s = nullptr;
    DBuf() = default;
    DBuf(int64_t n_, cudaStream_t s_) { alloc(n_, s_); }
    DBuf(const DBuf&) = delete;
    DBuf& operator=(const DBuf&) = delete;
    DBuf(DBuf&& o) noexcept : p(o.p), n(o.n), s(o.s) { o.p = nullptr; o.n = 0; }
    DBuf& operator=(DBuf&& o) noexcept {
        if (this != &o) { release(); p = o.p; n = o.n; s = o.s; o.p = nullptr; o.n = 0; }
        return *this;
    }
    ~DBuf() { release(); }
    void alloc(int64_t n_, cudaStream_t s_) {
        release();
        n = n_; s = s_;
        if (n > 0) SQSM_CUDA(cudaMallocAsync((void**)&p, sizeof(T) * (size_t)n, s));
    }
    void release() {
        if (p) cudaFreeAsync(p, s);
        p = nullptr; n = 0;
    }
    void zero() { if (n > 0) SQSM_CUDA(cudaMemsetAsync(p, 0, sizeof(T) * (size_t)n, s)); }
    void fill_ff() { if (n > 0) SQSM_CUDA(cudaMemsetAsync(p, 0xFF, sizeof(T) * (size_t)n, s)); }
    size_t bytes() const { return sizeof(T) * (size_t)n; }
    std::vector<T> to_host() const {
        std::vector<T> h((size_t)n);
        if (n > 0) {
            SQSM_CUDA(cudaMemcpyAsync(h.data(), p, bytes(), cudaMemcpyDeviceToHost, s));
            SQSM_CUDA(cudaStreamSynchronize(s));
        }
        return h;
    }
};

// ---------------------------------------------------------------------------- device helpers
#ifdef __CUDACC__

__device__ __forceinline__ uint64_t mix64(uint64_t k) {      // murmur3 finaliser
    k ^= k >> 33; k *= 0xff51afd7ed558ccdULL;
    k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ULL;
    k ^= k >> 33;
    return k;
}

constexpr uint64_t kEmptyKey = 0xFFFFFFFFFFFFFFFFULL;

// Brick key: 4 x 16-bit fields (sample | bz | by | bx); sorts lexicographically as an integer.
__host__ __device__ __forceinline__ uint64_t pack_key(uint32_t s, uint32_t bz, uint32_t by, uint32_t bx) {
    return ((uint64_t)s << 48) | ((uint64_t)bz << 32) | ((uint64_t)by << 16) | (uint64_t)bx;
}
__host__ __device__ __forceinline__ uint32_t key_s(uint64_t k) { return (uint32_t)(k >> 48); }
__host__ __device__ __forceinline__ uint32_t key_z(uint64_t k) { return (uint32_t)(k >> 32) & 0xFFFFu; }
__host__ __device__ __forceinline__ uint32_t key_y(uint64_t k) { return (uint32_t)(k >> 16) & 0xFFFFu; }
__host__ __device__ __forceinline__ uint32_t key_x(uint64_t k) { return (uint32_t)k & 0xFFFFu; }

// One 8x8x8-voxel brick = one 128-byte line: occupancy bitmap, per-word rank prefix, first row.
struct __align__(128) Brick {
    uint32_t mask[16];        // word = z*2 + (y>>2); bit = (y&3)*8 + x
    uint16_t prefix[16];      // exclusive popcount prefix over words
    int32_t  base;            // row of the first voxel of this brick
    int32_t  count;           // voxels in this brick
    uint64_t key;             // packed (sample, bz, by, bx)
    uint32_t pad[4];
};
static_assert(sizeof(Brick) == 128, "Brick must be one cache line");

__device__ __forceinline__ uint32_t pos_of(int z, int y, int x) { return (uint32_t)((z << 6) | (y << 3) | x); }

// rank of voxel `pos` inside brick `b` if occupied, else -1
__device__ __forceinline__ int brick_row(const Brick* __restrict__ b, uint32_t pos) {
    uint32_t w = pos >> 5, bit = pos & 31u;
    uint32_t m = b->mask[w];
    if (!((m >> bit) & 1u)) return -1;
    return b->base + (int)b->prefix[w] + __popc(m & ((1u << bit) - 1u));
}

// Open-addressing (linear probing) brick hash: 64-bit key -> brick id.
struct BrickHash {
    uint64_t* keys;
    int32_t*  vals;
    uint32_t  cap_mask;       // capacity - 1 (capacity is a power of two)
};

__device__ __forceinline__ int hash_find(const BrickHash& h, uint64_t key) {
    uint32_t slot = (uint32_t)mix64(key) & h.cap_mask;
    #pragma unroll 1
    for (uint32_t probe = 0; probe <= h.cap_mask; ++probe) {
        uint64_t k = h.keys[slot];
        if (k == key) return h.vals[slot];
        if (k == kEmptyKey) return -1;
        slot = (slot + 1) & h.cap_mask;
    }
    return -1;
}

// returns slot; *inserted = true when this thread claimed it
// `abort_flag` (optional): polled every 64 probes - once the caller knows the table is too small (it retries with a larger
// one) a probe sequence must not keep walking a full table
__device__ __forceinline__ uint32_t hash_insert(const BrickHash& h, uint64_t key, bool* inserted,
                                                const int32_t* abort_flag = nullptr) {
    uint32_t slot = (uint32_t)mix64(key) & h.cap_mask;
    *inserted = false;
    #pragma unroll 1
    for (uint32_t probe = 0; probe <= h.cap_mask; ++probe) {
        if (abort_flag && (probe & 63u) == 63u && *reinterpret_cast<const volatile int32_t*>(abort_flag) != 0) return 0xFFFFFFFFu;
        uint64_t k = h.keys[slot];
        if (k == key) return slot;
        if (k == kEmptyKey) {
            unsigned long long prev = atomicCAS((unsigned long long*)&h.keys[slot],
                                                (unsigned long long)kEmptyKey, (unsigned long long)key);
            if (prev == (unsigned long long)kEmptyKey) { *inserted = true; return slot; }
            if (prev == (unsigned long long)key) return slot;
        }
        slot = (slot + 1) & h.cap_mask;
    }
    return 0xFFFFFFFFu;       // table full (caller sized it at <= 50 % load, cannot happen)
}

#endif  // __CUDACC__

}  // namespace sqsm
