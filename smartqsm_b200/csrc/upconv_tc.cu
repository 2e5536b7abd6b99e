// SparseInverseConv3d(k = 2) on the tensor cores: a DENSE GEMM over the coarse rows with a scatter epilogue.  sm_100a only.
//
// Replaces spconv's SparseInverseConv3d forward reached from external/SmartTreeXX/models/sparse_module.py:65,113
// (`self.deconv`, indice_key shared with the matching SparseConv3d) and the BatchNorm1d + ReLU in front of it.
//
// Every fine row i has exactly one (coarse row o, kernel offset k) with i = 2o + k, so
//     out[child[k][o]] = W[k]^T . bnrelu(in_c[o])        for the 8 offsets k of every coarse row o
// is ONE dense product  Y[o, (k, c_out)] = A[o, c_in] . Wcat[c_in, 8*C_out]  whose A operand is the contiguous coarse
// feature matrix - no gather at all - followed by a scatter of the 8 row segments of Y[o] to the fine rows listed in the
// down-sample rulebook (child[8][n_coarse], -1 = no such fine voxel).  About half of the products are wasted on absent
// children (a coarse voxel has 3.5 of 8 on a surface), which costs nothing here: the kernel is bound by its stores.
//
//   * persistent CTA per SM; a work item = (tile of 128 coarse rows, block of 64 output columns);
//   * two producer groups copy the pre-split FP16 hi / lo operand rows (written by the epilogue of the convolution that
//     produced them) and the block's weight tile with cp.async into SWIZZLE_128B K-major tiles; completion on mbarriers;
//   * one thread issues tcgen05.mma kind::f16 twice per 16-element K step: A_hi.[W_hi | W_lo] (N = 128) then A_lo.W_hi
//     (N = 64) - the same 3-product FP16 split as conv_tc.cu (fp32-class accuracy), fp32 accumulation in TMEM;
//   * four epilogue warps read the accumulator (two TMEM stages), look the fine rows up and write, per fine row, whatever
//     the consumers need: the float32 row, the raw FP16 hi/lo split (operands of the 1x1 i_branch convolution) and the
//     BatchNorm + ReLU FP16 hi/lo split (operands of the tail block's first 3^3 convolution, second half of the concat).
// Fine rows without a parent (voxel dropped by the down-sample, or batch that did not descend) are filled by
// k_up_orphans: zeros, and relu(beta) for the activated split.
#include "engine.h"
#include "tc_ptx.cuh"
#include <cuda_fp16.h>
#include <cstring>

namespace sqsm {

struct UpArgs {
    const __half* in_hi;      // [n_coarse][CI] activated input, FP16 hi part
    const __half* in_lo;
    const int32_t* child;     // [8][n_coarse] fine row or -1
    const float* wimg;        // per 64-column block: hi tile [64][64] then lo tile, SWIZZLE_128B images of __half
    float* out;               // [n_fine][CO] or nullptr
    __half* raw_hi;           // [n_fine][CO] split of the raw output, or nullptr
    __half* raw_lo;
    __half* act_hi;           // [n_fine][CO] split of relu(out * alpha + beta), or nullptr
    __half* act_lo;
    const float* act_alpha;   // [CO] scale and shift of the consumer's BatchNorm (its channels of the concat)
    const float* act_beta;
    int* range_flag;
    int n_coarse;
    int n_tiles;              // tiles of 128 coarse rows
};

constexpr int kUpStages = 4;
constexpr int kUpProducers = 128;
constexpr int kUpGroups = 2;
constexpr int kUpThreads = kUpGroups * kUpProducers + 128 + 32;
constexpr int kUpNP = 64;                                  // output columns per work item
constexpr int kUpStageBytes = 2 * 128 * 128 + 2 * kUpNP * 128;

__device__ __forceinline__ void split_store8(const float* x, __half* hi, __half* lo, size_t o, int* range_flag) {
    float mx = 0.f;
    uint32_t hw[4], lw[4];
    #pragma unroll
    for (int i = 0; i < 4; ++i) {
        mx = fmaxf(mx, fmaxf(fabsf(x[2 * i]), fabsf(x[2 * i + 1])));
        const __half2 h = __floats2half2_rn(x[2 * i], x[2 * i + 1]);
        const float2 f = __half22float2(h);
        const __half2 l = __floats2half2_rn(x[2 * i] - f.x, x[2 * i + 1] - f.y);
        hw[i] = *reinterpret_cast<const uint32_t*>(&h);
        lw[i] = *reinterpret_cast<const uint32_t*>(&l);
    }
    if (!(mx < 3.0e4f)) *range_flag = 1;
    *reinterpret_cast<uint4*>(hi + o) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
    *reinterpret_cast<uint4*>(lo + o) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
}

template <int CI, int CO>
__global__ void __launch_bounds__(kUpThreads, 1) k_upconv_tc(const __grid_constant__ UpArgs a) {
    constexpr int NB = 8 * CO / kUpNP;                // column blocks per tile
    constexpr int KS = CI / 16;                       // K steps of 16 FP16 elements
    constexpr int PIECES = CI / 8;                    // 16-byte pieces per operand row
    constexpr int A_BYTES = 128 * 128, B_BYTES = kUpNP * 128;
    constexpr uint32_t IDESC = (1u << 4) | ((uint32_t)(kUpNP >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    constexpr uint32_t IDESC2 = (1u << 4) | ((uint32_t)((2 * kUpNP) >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    constexpr uint32_t ACC = 2 * kUpNP;
    static_assert(CI % 16 == 0 && CI <= 64 && (8 * CO) % kUpNP == 0 && kUpNP % CO == 0, "shape");

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    __shared__ __align__(8) uint64_t bars[2 * kUpStages + 4];
    __shared__ uint32_t tmem_slot;
    const int t = threadIdx.x, warp = t >> 5;
    const uint32_t full0 = smem_u32(&bars[0]), empty0 = smem_u32(&bars[kUpStages]);
    const uint32_t tfull0 = smem_u32(&bars[2 * kUpStages]), tempty0 = smem_u32(&bars[2 * kUpStages + 2]);
    constexpr int kProdWarps = kUpGroups * 4;
    if (t == 0) {
        for (int s = 0; s < kUpStages; ++s) { mbar_init(full0 + 8 * s, kUpProducers); mbar_init(empty0 + 8 * s, 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(tfull0 + 8 * s, 1); mbar_init(tempty0 + 8 * s, 128); }
        fence_barrier_init();
    }
    if (warp == kProdWarps + 4) tmem_alloc(smem_u32(&tmem_slot), 256);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_slot;
    const int n_items = a.n_tiles * NB;                       // item v = tile * NB + block, strided over the CTAs
    const int my_items = (n_items - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;

    if (warp < kProdWarps) {
        // ===================== producers: contiguous operand rows + the block's weight tile =====================
        const int grp = warp >> 2, tp = t & (kUpProducers - 1), lane = tp & 31, piece = lane & 7;
        const uint32_t smem_base = smem_u32(smem);
        for (int i = grp; i < my_items; i += kUpGroups) {
            const int v = (int)blockIdx.x + i * (int)gridDim.x;
            const int tile = v / NB, blk = v % NB;
            const uint32_t stage = (uint32_t)i % kUpStages, ph = ((uint32_t)i / kUpStages) & 1u;
            const uint32_t sA_hi = smem_base + stage * kUpStageBytes, sA_lo = sA_hi + A_BYTES, sB = sA_lo + A_BYTES;
            mbar_wait(empty0 + 8 * stage, ph ^ 1u);
            if (piece < PIECES) {
                #pragma unroll
                for (int m = 0; m < 8; ++m) {
                    const int rl = (lane >> 3) + 4 * m;
                    const uint32_t rt = (uint32_t)(tp & ~31) + (uint32_t)rl;
                    const int row = tile * 128 + (int)rt;
                    const bool on = row < a.n_coarse;
                    const size_t eoff = (on ? (size_t)row : 0) * CI + piece * 8;
                    const uint32_t off = rt * 128u + (uint32_t)((piece ^ (int)(rt & 7u)) << 4);
                    cp_async16(sA_hi + off, a.in_hi + eoff, on ? 16u : 0u);
                    cp_async16(sA_lo + off, a.in_lo + eoff, on ? 16u : 0u);
                }
            }
            {
                const float* src = a.wimg + (size_t)blk * (2 * B_BYTES / 4);
                constexpr int NPIECE = 2 * B_BYTES / 16;
                #pragma unroll
                for (int w = 0; w < NPIECE / kUpProducers; ++w) {
                    const int q = tp + w * kUpProducers;
                    cp_async16(sB + 16u * q, src + 4 * q, 16u);
                }
            }
            cp_async_arrive(full0 + 8 * stage);
        }
    } else if (warp < kProdWarps + 4) {
        // ===================== epilogue: TMEM -> registers -> scatter to the fine rows =====================
        const int te = t - kUpGroups * kUpProducers;
        for (int i = 0; i < my_items; ++i) {
            const int v = (int)blockIdx.x + i * (int)gridDim.x;
            const int tile = v / NB, blk = v % NB;
            const uint32_t as = (uint32_t)i & 1u, aph = ((uint32_t)i >> 1) & 1u;
            const int o = tile * 128 + te;
            const bool live = o < a.n_coarse;
            mbar_wait(tfull0 + 8 * as, aph);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + as * ACC;
            #pragma unroll
            for (int c0 = 0; c0 < kUpNP; c0 += 16) {
                float vv[16], ww[16];
                tmem_ld16(taddr + c0, vv);
                tmem_ld16(taddr + kUpNP + c0, ww);
                #pragma unroll
                for (int c = 0; c < 16; ++c) vv[c] += ww[c];
                if (c0 + 16 >= kUpNP) {
                    tc_fence_before();
                    mbar_arrive(tempty0 + 8 * as);
                }
                if (!live) continue;
                // 16 columns = 16 / CO offsets (CO = 8), one offset (CO = 16) or half an offset (CO = 32)
                #pragma unroll
                for (int c = 0; c < 16; c += 8) {
                    const int g = blk * kUpNP + c0 + c;                 // global column = k * CO + channel
                    const int k = g / CO, ch = g % CO;
                    const int f = __ldg(a.child + (size_t)k * a.n_coarse + o);
                    if (f < 0) continue;
                    const size_t fo = (size_t)f * CO + ch;
                    if (a.out) {
                        *reinterpret_cast<float4*>(a.out + fo) = make_float4(vv[c], vv[c + 1], vv[c + 2], vv[c + 3]);
                        *reinterpret_cast<float4*>(a.out + fo + 4) = make_float4(vv[c + 4], vv[c + 5], vv[c + 6], vv[c + 7]);
                    }
                    if (a.raw_hi) split_store8(vv + c, a.raw_hi, a.raw_lo, fo, a.range_flag);
                    if (a.act_hi) {
                        float x[8];
                        #pragma unroll
                        for (int q = 0; q < 8; ++q) x[q] = fmaxf(fmaf(vv[c + q], a.act_alpha[ch + q], a.act_beta[ch + q]), 0.f);
                        split_store8(x, a.act_hi, a.act_lo, fo, a.range_flag);
                    }
                }
            }
        }
    } else if ((t & 31) == 0) {
        // ===================== MMA issuer =====================
        for (int i = 0; i < my_items; ++i) {
            const uint32_t as = (uint32_t)i & 1u, aph = ((uint32_t)i >> 1) & 1u;
            const uint32_t stage = (uint32_t)i % kUpStages, ph = ((uint32_t)i / kUpStages) & 1u;
            mbar_wait(tempty0 + 8 * as, aph ^ 1u);
            tc_fence_after();
            const uint32_t tmem_d = tmem_base + as * ACC;
            mbar_wait(full0 + 8 * stage, ph);
            fence_proxy_async();
            tc_fence_after();
            const uint32_t sA_hi = smem_u32(smem + (size_t)stage * kUpStageBytes);
            const uint32_t sA_lo = sA_hi + A_BYTES, sB_hi = sA_lo + A_BYTES;
            #pragma unroll
            for (int ks = 0; ks < KS; ++ks) {
                const uint64_t ah = make_desc_sw128(sA_hi + ks * 32), al = make_desc_sw128(sA_lo + ks * 32);
                const uint64_t bh = make_desc_sw128(sB_hi + ks * 32);
                umma_f16(tmem_d, ah, bh, IDESC2, ks ? 1u : 0u);
                umma_f16(tmem_d, al, bh, IDESC, 1u);
            }
            umma_commit(empty0 + 8 * stage);
            umma_commit(tfull0 + 8 * as);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == kProdWarps + 4) {
        __syncwarp();
        tmem_dealloc(tmem_base, 256);
    }
}

// fine rows without a parent: zeros ("receive a zero decoder half"), relu(beta) after the consumer's BatchNorm
__global__ void __launch_bounds__(256) k_up_orphans(const int32_t* __restrict__ parent, int64_t n, int co, float* __restrict__ out,
                                                    __half* __restrict__ raw_hi, __half* __restrict__ raw_lo,
                                                    __half* __restrict__ act_hi, __half* __restrict__ act_lo,
                                                    const float* __restrict__ act_beta) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int per_row = co / 8;
    const int64_t r = i / per_row;
    if (r >= n || parent[r] >= 0) return;
    const int ch = (int)(i % per_row) * 8;
    const size_t o = (size_t)r * co + ch;
    if (out) {
        *reinterpret_cast<float4*>(out + o) = make_float4(0.f, 0.f, 0.f, 0.f);
        *reinterpret_cast<float4*>(out + o + 4) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    if (raw_hi) {
        *reinterpret_cast<uint4*>(raw_hi + o) = make_uint4(0, 0, 0, 0);
        *reinterpret_cast<uint4*>(raw_lo + o) = make_uint4(0, 0, 0, 0);
    }
    if (act_hi) {
        uint32_t hw[4], lw[4];
        #pragma unroll
        for (int q = 0; q < 4; ++q) {
            const float x0 = fmaxf(act_beta[ch + 2 * q], 0.f), x1 = fmaxf(act_beta[ch + 2 * q + 1], 0.f);
            const __half2 h = __floats2half2_rn(x0, x1);
            const float2 f = __half22float2(h);
            const __half2 l = __floats2half2_rn(x0 - f.x, x1 - f.y);
            hw[q] = *reinterpret_cast<const uint32_t*>(&h);
            lw[q] = *reinterpret_cast<const uint32_t*>(&l);
        }
        *reinterpret_cast<uint4*>(act_hi + o) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
        *reinterpret_cast<uint4*>(act_lo + o) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
    }
}

// ------------------------------------------------------------------------------------------ host side

std::vector<float> make_tc_weight_image_f16(const float* w_kcico, int kv, int ci, int co);   // conv_tc.cu

// Weight image of the inverse convolution: for every block of 64 output columns g = k * co + c the K-major tile
// B[g][ci] = W[k][ci][c] (K padded to 64), FP16 hi then lo, SWIZZLE_128B - i.e. the image of a 1x1 convolution ci -> 64.
std::vector<float> make_up_weight_image_f16(const float* w_kcico, int ci, int co) {
    const int nb = 8 * co / kUpNP;
    std::vector<float> img;
    std::vector<float> wb((size_t)ci * kUpNP);
    for (int b = 0; b < nb; ++b) {
        for (int c = 0; c < ci; ++c)
            for (int j = 0; j < kUpNP; ++j) {
                const int g = b * kUpNP + j, k = g / co, o = g % co;
                wb[(size_t)c * kUpNP + j] = w_kcico[((size_t)k * ci + c) * co + o];
            }
        std::vector<float> one = make_tc_weight_image_f16(wb.data(), 1, ci, kUpNP);
        img.insert(img.end(), one.begin(), one.end());
    }
    return img;
}

bool upconv_tc(Ctx& cx, int ci, int co, const void* in_hi, const void* in_lo, const int32_t* child, const int32_t* parent,
               const float* wimg, float* out, void* raw_hi, void* raw_lo, void* act_hi, void* act_lo, const float* act_alpha,
               const float* act_beta, int* range_flag, int64_t n_coarse, int64_t n_fine) {
    UpArgs a;
    a.in_hi = (const __half*)in_hi; a.in_lo = (const __half*)in_lo; a.child = child; a.wimg = wimg; a.out = out;
    a.raw_hi = (__half*)raw_hi; a.raw_lo = (__half*)raw_lo; a.act_hi = (__half*)act_hi; a.act_lo = (__half*)act_lo;
    a.act_alpha = act_alpha; a.act_beta = act_beta; a.range_flag = range_flag; a.n_coarse = (int)n_coarse; a.n_tiles = div_up(n_coarse, 128);
    if (n_fine <= 0) return true;
    {
        const int64_t items = n_fine * (co / 8);
        k_up_orphans<<<div_up(items, 256), 256, 0, cx.st>>>(parent, n_fine, co, out, a.raw_hi, a.raw_lo, a.act_hi, a.act_lo, act_beta);
        cx.launches++;
    }
    if (n_coarse <= 0) return true;
    const size_t smem = (size_t)kUpStages * kUpStageBytes + 1024;
#define SQSM_UP_CASE(CI_, CO_)                                                                                   \
    if (ci == CI_ && co == CO_) {                                                                                \
        auto kern = k_upconv_tc<CI_, CO_>;                                                                       \
        SQSM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));          \
        const int items = a.n_tiles * (8 * CO_ / kUpNP);                                                         \
        kern<<<std::min(items, kNumSMs), kUpThreads, smem, cx.st>>>(a);                                          \
        cx.launches++;                                                                                           \
        return true;                                                                                             \
    }
    SQSM_UP_CASE(16, 8) SQSM_UP_CASE(32, 16) SQSM_UP_CASE(64, 32)
#undef SQSM_UP_CASE
    return false;
}

}  // namespace sqsm
