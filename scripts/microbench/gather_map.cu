// Micro-benchmark: LSU gather (cp.async 16 B) of 64-byte operand rows [hi 16 ch | lo 16 ch] into shared memory with three
// lane mappings, lean instruction stream, rulebook-like indices with a given present fraction.  Answers: what is the
// LDGSTS gather rate per SM when the producer loop itself is not the limit, and which mapping is best?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 --cudart shared -o /tmp/gather_map gather_map.cu && /tmp/gather_map
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void cp16(uint32_t dst, const void* src, bool on) {
    asm volatile("{.reg .pred p; setp.eq.u32 p, %2, 0; cp.async.cg.shared.global [%0], [%1], 16, p;}" ::"r"(dst), "l"(src), "r"((uint32_t)on) : "memory");
}
__device__ __forceinline__ void commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void wait_group() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// idx[tile][k][128] : source row of slot (tile row r, offset k) or -1
__global__ void k_make_idx(int* idx, int64_t n_tiles, int K, int64_t n_rows, int present_256) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_tiles * K * 128) return;
    int r = (int)(i & 127);
    int64_t t = i >> 7;
    int k = (int)(t % K);
    int64_t tile = t / K;
    uint64_t h = (uint64_t)i * 0x9E3779B97F4A7C15ull; h ^= h >> 29; h *= 0xBF58476D1CE4E5B9ull; h ^= h >> 32;
    int64_t delta = ((k % 3) - 1) + ((k / 3) % 3 - 1) * 24 + ((k / 9) - 1) * 1500;
    int64_t row = tile * 128 + r + delta;
    if (row < 0 || row >= n_rows || (int)(h & 255) >= present_256) row = -1;
    idx[i] = (int)row;
}

// MODE 0: lane = (row r = lane + 32*w..., all 4 pieces by the same lane)      thread-per-slot
// MODE 1: 4 lanes per slot (64 B contiguous), 8 slots (rows r .. r+7 of one offset) per instruction
// MODE 2: 8 lanes per tile row cover 4 offsets x 32 B (hi), second instruction the 4 x 32 B lo halves   (production-like)
template <int MODE, int W, int S>
__global__ void __launch_bounds__(W * 32) k_gather(const uint8_t* __restrict__ table, const int* __restrict__ idx, int64_t n_tiles, int K) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // unit of work per warp: (tile, group of 4 offsets) = 128 rows x 4 slots x 64 B = 32 KB -> too big per warp stage; use 32 rows:
    // a warp handles 32 tile rows x 4 offsets (= one 128-byte hi chunk row + lo) = 8 KB per stage
    uint8_t* ring = smem + (size_t)w * S * 8192;
    const int KG = (K + 3) / 4;
    const int64_t units = n_tiles * KG * 4;                   // (tile, kgroup, quarter of the tile rows)
    int64_t it = 0;
    for (int64_t u = (int64_t)blockIdx.x * W + w; u < units; u += (int64_t)gridDim.x * W, ++it) {
        const int quarter = (int)(u & 3);
        const int64_t tk = u >> 2;
        const int kg = (int)(tk % KG);
        const int64_t tile = tk / KG;
        const uint32_t dst = smem_u32(ring + (size_t)(it % S) * 8192);
        if (it >= S) wait_group<S - 1>();
        const int* ib = idx + (tile * K) * 128 + quarter * 32;
        if (MODE == 0) {
            #pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int k = kg * 4 + q;
                const int n = (k < K) ? __ldg(ib + (int64_t)k * 128 + lane) : -1;
                const uint8_t* src = table + (size_t)(n < 0 ? 0 : n) * 64;
                #pragma unroll
                for (int c = 0; c < 4; ++c) cp16(dst + lane * 256 + q * 64 + c * 16, src + c * 16, n >= 0);
            }
        } else if (MODE == 1) {
            #pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int k = kg * 4 + q;
                #pragma unroll
                for (int m = 0; m < 4; ++m) {
                    const int r = m * 8 + (lane >> 2);
                    const int n = (k < K) ? __ldg(ib + (int64_t)k * 128 + r) : -1;
                    const uint8_t* src = table + (size_t)(n < 0 ? 0 : n) * 64 + (lane & 3) * 16;
                    cp16(dst + r * 256 + q * 64 + (lane & 3) * 16, src, n >= 0);
                }
            }
        } else {
            const int q = (lane & 7) >> 1;
            const int k = kg * 4 + q;
            #pragma unroll
            for (int m = 0; m < 8; ++m) {
                const int r = m * 4 + (lane >> 3);
                const int n = (k < K) ? __ldg(ib + (int64_t)k * 128 + r) : -1;
                const uint8_t* src = table + (size_t)(n < 0 ? 0 : n) * 64 + (lane & 1) * 16;
                cp16(dst + r * 128 + (lane & 7) * 16, src, n >= 0);                  // hi half
                cp16(dst + 4096 + r * 128 + (lane & 7) * 16, src + 32, n >= 0);     // lo half
            }
        }
        commit();
    }
    wait_group<0>();
}

template <int MODE, int W, int S>
void run(const uint8_t* table, int* idx, int64_t n_tiles, int K, int present_256, int64_t n_rows) {
    k_make_idx<<<(unsigned)((n_tiles * K * 128 + 255) / 256), 256>>>(idx, n_tiles, K, n_rows, present_256);
    CK(cudaDeviceSynchronize());
    const size_t smem = (size_t)W * S * 8192;
    auto kern = k_gather<MODE, W, S>;
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int sms = 0;
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    kern<<<sms, W * 32, smem>>>(table, idx, n_tiles, K);
    CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int rep = 0; rep < 3; ++rep) {
        CK(cudaEventRecord(e0));
        kern<<<sms, W * 32, smem>>>(table, idx, n_tiles, K);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        best = ms < best ? ms : best;
    }
    const double slots = (double)n_tiles * K * 128, present = slots * present_256 / 256.0;
    printf("mode=%d warps=%2d stages=%d present=%3d/256 : %8.3f ms  %5.2f ns*SM/slot  %5.2f ns*SM/present slot  %6.1f B/clk/SM (present bytes, 1.92 GHz)  tile(27x128 slots)=%5.2f us\n",
           MODE, W, S, present_256, best, best * 1e6 * sms / slots, best * 1e6 * sms / present, present * 64 / (best * 1e-3) / sms / 1.92e9,
           best * 1e3 * sms / n_tiles);
}

int main() {
    const int64_t n_rows = 16ll << 20;                // 1 GB of 64-byte rows
    uint8_t* table; CK(cudaMalloc(&table, n_rows * 64)); CK(cudaMemset(table, 1, n_rows * 64));
    const int K = 27;
    const int64_t n_tiles = n_rows / 128 / 4;
    int* idx; CK(cudaMalloc(&idx, n_tiles * K * 128 * sizeof(int)));
    for (int present : {97, 256}) {                    // 38 % (level 1) and 100 %
        run<0, 8, 3>(table, idx, n_tiles, K, present, n_rows);
        run<1, 8, 3>(table, idx, n_tiles, K, present, n_rows);
        run<2, 8, 3>(table, idx, n_tiles, K, present, n_rows);
        run<1, 12, 2>(table, idx, n_tiles, K, present, n_rows);
        run<2, 12, 2>(table, idx, n_tiles, K, present, n_rows);
        run<1, 16, 1>(table, idx, n_tiles, K, present, n_rows);
        run<2, 16, 1>(table, idx, n_tiles, K, present, n_rows);
        run<1, 24, 1>(table, idx, n_tiles, K, present, n_rows);
    }
    return 0;
}
