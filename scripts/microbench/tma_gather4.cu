// Micro-benchmark: row-gather throughput into shared memory, TMA tile::gather4 against cp.async (LDGSTS) 16-byte copies.
// Question it answers (profiles/r01_tc_kernel_notes.md): can the tensor-core convolution's rulebook gather move from the
// LSU to the TMA unit, and how many cycles per gathered row does each path cost per SM?
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tma_gather4 tma_gather4.cu && ./tma_gather4
//
// One persistent CTA per SM; W producer warps each own a ring of S stages of 128 rows; chunk = 128 gathered rows.
// Index patterns: 0 = uniform random rows of a 1 GB table, 1 = rulebook-like (rows near the tile), 2 = pattern 1 with
// half of the rows absent (-1: TMA zero-fills out-of-bounds rows, cp.async uses src-size 0).
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, int n) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(n)); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* b, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
    asm volatile(
        "{\n.reg .pred p;\nWAIT:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra DONE;\nbra WAIT;\nDONE:\n}" ::"r"(smem_u32(b)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tma_gather4(const CUtensorMap* tm, uint64_t* bar, void* dst, int col, int r0, int r1, int r2, int r3) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];" ::"r"(
            smem_u32(dst)),
        "l"(tm), "r"(col), "r"(r0), "r"(r1), "r"(r2), "r"(r3), "r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void cp_async16(void* dst, const void* src, int src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(dst)), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_arrive(uint64_t* b) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(b)) : "memory");
}

__global__ void k_make_idx(int* idx, int64_t n_chunks, int K, int64_t n_rows, int pattern) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_chunks * 128) return;
    int64_t chunk = i >> 7;
    int r = (int)(i & 127);
    int64_t tile = chunk / K;
    int k = (int)(chunk % K);
    uint64_t h = (uint64_t)i * 0x9E3779B97F4A7C15ull;
    h ^= h >> 29; h *= 0xBF58476D1CE4E5B9ull; h ^= h >> 32;
    int64_t row;
    if (pattern == 0) row = (int64_t)(h % (uint64_t)n_rows);
    else {
        // neighbours of a voxel row live within a few thousand rows (brick order); consecutive rows have consecutive neighbours
        int64_t delta = ((k % 3) - 1) + ((k / 3) % 3 - 1) * 24 + ((k / 9) - 1) * 1500;
        row = tile * 128 + r + delta;
        if (row < 0 || row >= n_rows) row = -1;
        if (pattern == 2 && (h & 1)) row = -1;
        // pattern 3/4: absent rows (half / 62 %) are redirected to a pool of 4096 in-bounds "zero rows" spread over the table
        if (pattern == 3 && (h & 1)) row = (int64_t)(((h >> 8) & 4095) * 997 % (uint64_t)n_rows);
        if (pattern == 4 && ((h & 255) < 159)) row = (int64_t)(((h >> 8) & 4095) * 997 % (uint64_t)n_rows);
    }
    idx[i] = (int)row;
}

template <int ROW_BYTES, int W, int S, bool TMA>
__global__ void __launch_bounds__(W * 32) k_gather(const __grid_constant__ CUtensorMap tm, const uint8_t* __restrict__ table,
                                                    const int* __restrict__ idx, int64_t n_chunks, uint8_t* dump) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    __shared__ uint64_t full[W * S];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    constexpr int STAGE = 128 * ROW_BYTES;
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0)
        for (int i = 0; i < W * S; ++i) mbar_init(&full[i], TMA ? 1 : 32);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async;" ::: "memory");
    __syncthreads();
    uint8_t* ring = smem + (size_t)w * S * STAGE;
    int64_t it = 0;
    int64_t last_chunk = -1;
    for (int64_t chunk = (int64_t)blockIdx.x * W + w; chunk < n_chunks; chunk += (int64_t)gridDim.x * W, ++it) {
        const int st = (int)(it % S);
        if (it >= S) mbar_wait(&full[w * S + st], (uint32_t)((it / S - 1) & 1));
        uint8_t* dst = ring + (size_t)st * STAGE;
        const int4 rows = *reinterpret_cast<const int4*>(idx + chunk * 128 + lane * 4);
        if (TMA) {
            if (lane == 0) mbar_expect_tx(&full[w * S + st], STAGE);
            __syncwarp();
            tma_gather4(&tm, &full[w * S + st], dst + lane * 4 * ROW_BYTES, 0, rows.x, rows.y, rows.z, rows.w);
        } else {
            // lane copies its 4 rows, ROW_BYTES / 16 chunks each, into the same swizzled layout
            const int rr[4] = {rows.x, rows.y, rows.z, rows.w};
            #pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int row = lane * 4 + j;
                const uint8_t* src = table + (size_t)(rr[j] < 0 ? 0 : rr[j]) * ROW_BYTES;
                #pragma unroll
                for (int c = 0; c < ROW_BYTES / 16; ++c) {
                    const int cs = ROW_BYTES == 128 ? (c ^ (row & 7)) : (c ^ ((row >> 1) & 3));
                    cp_async16(dst + row * ROW_BYTES + cs * 16, src + c * 16, rr[j] < 0 ? 0 : 16);
                }
            }
            cp_async_arrive(&full[w * S + st]);
        }
        last_chunk = chunk;
    }
    // drain
    for (int64_t j = (it > S ? it - S : 0); j < it; ++j) mbar_wait(&full[w * S + (int)(j % S)], (uint32_t)((j / S) & 1));
    if (dump != nullptr && blockIdx.x == 0 && w == 0 && last_chunk >= 0) {
        const int st = (int)((it - 1) % S);
        for (int i = lane; i < STAGE; i += 32) dump[i] = ring[(size_t)st * STAGE + i];
        if (lane == 0) *reinterpret_cast<int64_t*>(dump + STAGE) = last_chunk;
    }
}

typedef CUresult (*EncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

template <int ROW_BYTES, int W, int S, bool TMA>
void run(EncodeTiled enc, uint8_t* table, int64_t n_rows, int* idx, int64_t n_chunks, int K, int pattern, const std::vector<uint8_t>& h_table) {
    CUtensorMap tm;
    cuuint64_t dims[2] = {(cuuint64_t)ROW_BYTES / 2, (cuuint64_t)n_rows};
    cuuint64_t strides[1] = {(cuuint64_t)ROW_BYTES};
    cuuint32_t box[2] = {(cuuint32_t)ROW_BYTES / 2, 1};
    cuuint32_t es[2] = {1, 1};
    CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, table, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     ROW_BYTES == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("cuTensorMapEncodeTiled failed %d\n", (int)r); exit(1); }
    k_make_idx<<<(unsigned)((n_chunks * 128 + 255) / 256), 256>>>(idx, n_chunks, K, n_rows, pattern);
    CK(cudaDeviceSynchronize());
    constexpr int STAGE = 128 * ROW_BYTES;
    const size_t smem = (size_t)W * S * STAGE + 1024;
    auto kern = k_gather<ROW_BYTES, W, S, TMA>;
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    uint8_t* dump;
    CK(cudaMalloc(&dump, STAGE + 8));
    int sms = 0;
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    kern<<<sms, W * 32, smem>>>(tm, table, idx, n_chunks, dump);     // warm-up + layout check
    CK(cudaDeviceSynchronize());
    std::vector<uint8_t> h(STAGE + 8);
    CK(cudaMemcpy(h.data(), dump, STAGE + 8, cudaMemcpyDeviceToHost));
    const int64_t lc = *reinterpret_cast<int64_t*>(h.data() + STAGE);
    std::vector<int> hidx(128);
    CK(cudaMemcpy(hidx.data(), idx + lc * 128, 512, cudaMemcpyDeviceToHost));
    int64_t bad = 0;
    for (int row = 0; row < 128; ++row)
        for (int c = 0; c < ROW_BYTES / 16; ++c) {
            const int cs = ROW_BYTES == 128 ? (c ^ (row & 7)) : (c ^ ((row >> 1) & 3));
            for (int b = 0; b < 16; ++b) {
                const uint8_t want = hidx[row] < 0 ? 0 : h_table[((size_t)hidx[row] * ROW_BYTES + c * 16 + b) % h_table.size()];
                if (h[row * ROW_BYTES + cs * 16 + b] != want) ++bad;
            }
        }
    float best = 1e30f;
    for (int rep = 0; rep < 3; ++rep) {
        CK(cudaEventRecord(e0));
        kern<<<sms, W * 32, smem>>>(tm, table, idx, n_chunks, nullptr);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        best = ms < best ? ms : best;
    }
    const double rows = (double)n_chunks * 128;
    printf("%-6s row=%3dB warps=%d stages=%d pattern=%d : %8.3f ms  %7.1f Grows/s  %6.1f GB/s(all rows)  %5.2f ns*SM/row  layout_mismatch_bytes=%lld\n",
           TMA ? "TMA" : "LDGSTS", ROW_BYTES, W, S, pattern, best, rows / best / 1e6, rows * ROW_BYTES / best / 1e6,
           best * 1e6 * sms / rows, (long long)bad);
    CK(cudaFree(dump));
}

int main(int argc, char**) {
    EncodeTiled enc = nullptr;
    cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&enc, cudaEnableDefault, &q));
    if (!enc) { printf("no cuTensorMapEncodeTiled\n"); return 1; }
    const int64_t table_bytes = 1ll << 30;
    uint8_t* table;
    CK(cudaMalloc(&table, table_bytes));
    // table content: periodic pseudo-random bytes (period 1 MiB + 64 so rows differ)
    std::vector<uint8_t> h_table((1 << 20) + 64);
    uint32_t s = 12345;
    for (auto& b : h_table) { s = s * 1664525u + 1013904223u; b = (uint8_t)(s >> 24); }
    for (int64_t off = 0; off < table_bytes; off += (int64_t)h_table.size()) {
        const int64_t n = std::min<int64_t>((int64_t)h_table.size(), table_bytes - off);
        CK(cudaMemcpy(table + off, h_table.data(), n, cudaMemcpyHostToDevice));
    }
    const int K = 27;
    int* idx;
    {
        const int64_t max_rows = table_bytes / 64;
        CK(cudaMalloc(&idx, (max_rows / 128) * K * 128 * sizeof(int) / 4));       // a quarter of the tiles is plenty
    }
    const bool quick = argc > 1;              // any argument: only the 16-warp / zero-row-pool follow-up
    for (int pattern = quick ? 1 : 0; pattern < (quick ? 5 : 3); ++pattern) {
        {
            const int64_t n_rows = table_bytes / 64, n_chunks = (n_rows / 128 / 4) * K;
            if (!quick) {
                run<64, 1, 8, true>(enc, table, n_rows, idx, n_chunks, K, pattern, h_table);
                run<64, 4, 4, true>(enc, table, n_rows, idx, n_chunks, K, pattern, h_table);
            }
            run<64, 8, 2, true>(enc, table, n_rows, idx, n_chunks, K, pattern, h_table);
            if (quick) {
                run<64, 16, 1, true>(enc, table, n_rows, idx, n_chunks, K, pattern, h_table);
                run<64, 24, 1, true>(enc, table, n_rows, idx, n_chunks, K, pattern, h_table);
            }
            if (!quick) run<64, 4, 4, false>(enc, table, n_rows, idx, n_chunks, K, pattern, h_table);
            run<64, 8, 2, false>(enc, table, n_rows, idx, n_chunks, K, pattern, h_table);
        }
        {
            const int64_t n_rows = table_bytes / 128, n_chunks = (n_rows / 128 / 2) * K;
            if (!quick) run<128, 1, 8, true>(enc, table, n_rows, idx, n_chunks, K, pattern, h_table);
            run<128, 4, 3, true>(enc, table, n_rows, idx, n_chunks, K, pattern, h_table);
            if (quick) {
                run<128, 8, 1, true>(enc, table, n_rows, idx, n_chunks, K, pattern, h_table);
                run<128, 12, 1, true>(enc, table, n_rows, idx, n_chunks, K, pattern, h_table);
            }
            run<128, 4, 3, false>(enc, table, n_rows, idx, n_chunks, K, pattern, h_table);
        }
    }
    return 0;
}
