// Sparse convolutions (gather-GEMM, output stationary, fp32 FFMA), the MLP heads and the fused
// decode + contraction update.  sm_100a.
//
// Replaces spconv's SubMConv3d / SparseConv3d(k2,s2) / SparseInverseConv3d forward kernels reached
// from external/SmartTreeXX/models/smart_tree_xx.py:64-78 and sparse_module.py:32-39,101-119, the
// un-fused BatchNorm1d/ReLU/add/cat ATen kernels between them (K7-K9 of SURVEY.md §2.2), the two
// MLP heads with F.normalize / F.softplus (K10) and the contraction update of
// core/core_algorithms/spconv_based_contraction.py:103-108 (K11).
//
// One kernel template serves every convolution: each thread owns RPT output rows x CPT output
// channels in registers, walks the KV kernel offsets of the output-stationary rulebook
// (nbr[k][row] = input row or -1), applies the preceding BatchNorm+ReLU while gathering, and takes
// the weights of the current offsets from shared memory as warp-broadcast float4 reads.  Residual
// add, skip-concat (two input pointers) and the per-batch "did not descend" select are fused in.
#include "engine.h"
#include <cuda_fp16.h>
#include <algorithm>
#include <cstring>

namespace sqsm {

struct ConvArgs {
    const float* inA;        // [n_in][CI] (or first CI/2 channels when CAT)
    const float* inB;        // second CI/2 channels when CAT
    const int32_t* nbr;      // [KV][n_out] or nullptr (identity, KV == 1)
    const float* w;          // [KV][CI][CO]
    const float* bn;         // [2][CI] or nullptr
    const float* res;        // [n_out][CO] residual or nullptr
    const float* sel_src;    // [n_out][CO] value for rows with sel_flag == 0 (or nullptr)
    const uint8_t* sel_flag;
    float* out;              // [n_out][CO]
    const uint8_t* tile_need;// per 128-row tile: 0 = no row of the tile can reach an interior output (decoder layers): the
                             // block writes zeros instead of computing; nullptr = compute everything
    int n_out;
    int kc;                  // kernel offsets per shared-memory weight chunk
};

__device__ __forceinline__ float comp4(const float4& v, int i) {
    return i == 0 ? v.x : (i == 1 ? v.y : (i == 2 ? v.z : v.w));
}

template <int CI, int CO, int KV, int RPT, int CPT, bool BN, bool CAT>
__global__ void __launch_bounds__(256) k_conv(ConvArgs a) {
    constexpr int SLICES = CO / CPT;
    constexpr int RG = 256 / SLICES;         // row groups per block (multiple of 32: slice is warp-uniform)
    constexpr int TR = RG * RPT;             // rows per block
    static_assert(CO % CPT == 0 && CPT % 4 == 0 && CI % 4 == 0 && RG % 32 == 0, "tile shape");
    extern __shared__ __align__(16) float smem[];
    float* sbn = smem;                       // [2][CI]
    float* sw = smem + 2 * CI;               // [kc][CI][CO]

    const int t = threadIdx.x;
    const int slice = t / RG, rg = t % RG;
    const int row0 = blockIdx.x * TR + rg;
    if (a.tile_need) {                                   // skipped block: defined (zero) outputs, no gather, no FMA
        static_assert(TR % 128 == 0, "tile-need flags are per 128 rows");
        bool any = false;
        #pragma unroll
        for (int q = 0; q < TR / 128; ++q) {
            const int tile = blockIdx.x * (TR / 128) + q;
            any |= (tile * 128 < a.n_out) && a.tile_need[tile] != 0;
        }
        if (!any) {
            const int64_t e0 = (int64_t)blockIdx.x * TR * CO, e1 = min((int64_t)a.n_out * CO, e0 + (int64_t)TR * CO);
            for (int64_t e = e0 + 4 * t; e < e1; e += 4 * 256) *reinterpret_cast<float4*>(a.out + e) = make_float4(0.f, 0.f, 0.f, 0.f);
            return;
        }
    }
    if (BN) {
        for (int i = t; i < 2 * CI; i += 256) sbn[i] = a.bn[i];
    }
    // accumulators as float2 pairs: fma.rn.f32x2 (__ffma2_rn, sm_100) issues two exact fp32 FMAs per instruction slot
    float2 acc2[RPT][CPT / 2];
    #pragma unroll
    for (int j = 0; j < RPT; ++j)
        #pragma unroll
        for (int c = 0; c < CPT / 2; ++c) acc2[j][c] = make_float2(0.0f, 0.0f);

    for (int k0 = 0; k0 < KV; k0 += a.kc) {
        const int kn = min(a.kc, KV - k0);
        __syncthreads();                     // previous chunk fully consumed (and sbn visible)
        {
            const float4* src = reinterpret_cast<const float4*>(a.w + (size_t)k0 * CI * CO);
            float4* dst = reinterpret_cast<float4*>(sw);
            for (int i = t; i < kn * CI * CO / 4; i += 256) dst[i] = src[i];
        }
        __syncthreads();
        for (int kk = 0; kk < kn; ++kk) {
            const int k = k0 + kk;
            int nj[RPT];
            bool any = false;
            #pragma unroll
            for (int j = 0; j < RPT; ++j) {
                int row = row0 + j * RG;
                int n = -1;
                if (row < a.n_out) n = a.nbr ? a.nbr[(size_t)k * a.n_out + row] : row;
                nj[j] = n;
                any |= (n >= 0);
            }
            if (!__any_sync(0xFFFFFFFFu, any)) continue;
            const float* wk = sw + (size_t)kk * CI * CO + slice * CPT;
            #pragma unroll
            for (int c0 = 0; c0 < CI; c0 += 4) {
                float4 v[RPT];
                #pragma unroll
                for (int j = 0; j < RPT; ++j) {
                    if (nj[j] >= 0) {
                        const float* p;
                        if (CAT) p = (c0 < CI / 2) ? a.inA + (size_t)nj[j] * (CI / 2) + c0
                                                   : a.inB + (size_t)nj[j] * (CI / 2) + (c0 - CI / 2);
                        else p = a.inA + (size_t)nj[j] * CI + c0;
                        float4 x = __ldg(reinterpret_cast<const float4*>(p));
                        if (BN) {
                            x.x = fmaxf(fmaf(x.x, sbn[c0 + 0], sbn[CI + c0 + 0]), 0.0f);
                            x.y = fmaxf(fmaf(x.y, sbn[c0 + 1], sbn[CI + c0 + 1]), 0.0f);
                            x.z = fmaxf(fmaf(x.z, sbn[c0 + 2], sbn[CI + c0 + 2]), 0.0f);
                            x.w = fmaxf(fmaf(x.w, sbn[c0 + 3], sbn[CI + c0 + 3]), 0.0f);
                        }
                        v[j] = x;
                    } else {
                        v[j] = make_float4(0.f, 0.f, 0.f, 0.f);
                    }
                }
                #pragma unroll
                for (int cc = 0; cc < 4; ++cc) {
                    const float* wrow = wk + (size_t)(c0 + cc) * CO;
                    #pragma unroll
                    for (int q = 0; q < CPT / 4; ++q) {
                        const float4 wv = *reinterpret_cast<const float4*>(wrow + 4 * q);
                        const float2 w01 = make_float2(wv.x, wv.y), w23 = make_float2(wv.z, wv.w);
                        #pragma unroll
                        for (int j = 0; j < RPT; ++j) {
                            const float s = comp4(v[j], cc);
                            const float2 ss = make_float2(s, s);
                            acc2[j][2 * q + 0] = __ffma2_rn(ss, w01, acc2[j][2 * q + 0]);
                            acc2[j][2 * q + 1] = __ffma2_rn(ss, w23, acc2[j][2 * q + 1]);
                        }
                    }
                }
            }
        }
    }
    // epilogue: residual add, "batch did not descend" select, store
    #pragma unroll
    for (int j = 0; j < RPT; ++j) {
        int row = row0 + j * RG;
        if (row >= a.n_out) continue;
        const size_t o = (size_t)row * CO + slice * CPT;
        const bool take_src = a.sel_flag && a.sel_flag[row] == 0;
        #pragma unroll
        for (int q = 0; q < CPT / 4; ++q) {
            float4 r = make_float4(acc2[j][2 * q].x, acc2[j][2 * q].y, acc2[j][2 * q + 1].x, acc2[j][2 * q + 1].y);
            if (a.res) {
                float4 e = *reinterpret_cast<const float4*>(a.res + o + 4 * q);
                r.x += e.x; r.y += e.y; r.z += e.z; r.w += e.w;
            }
            if (take_src) r = *reinterpret_cast<const float4*>(a.sel_src + o + 4 * q);
            *reinterpret_cast<float4*>(a.out + o + 4 * q) = r;
        }
    }
}

// SparseInverseConv3d(k=2): every fine row has exactly one (parent, kernel offset) pair, so
// out[i] = W[koff[i]] . bnrelu(in_c[parent[i]]) (zeros when the voxel did not survive the down-sample).
// A block takes a tile of 1024 rows, buckets them by kernel offset into 8 shared-memory lists and then runs 8
// dense passes: inside a pass every lane uses the SAME weight matrix, so the weights are warp-broadcast
// shared-memory reads (a per-lane offset makes them 8-way conflicting and the kernel shared-memory bound).
template <int CI, int CO>
__global__ void __launch_bounds__(256) k_upconv(const float* __restrict__ in, const int32_t* __restrict__ parent,
                                                const int32_t* __restrict__ koff, const float* __restrict__ w,
                                                const float* __restrict__ bn, float* __restrict__ out, int n) {
    constexpr int T = 1024;
    extern __shared__ __align__(16) float smem[];
    float* sbn = smem;                               // [2][CI]
    float* sw = smem + 2 * CI;                       // [CI][CO] weights of the current offset
    uint16_t* lists = reinterpret_cast<uint16_t*>(sw + CI * CO);   // [8][T] local rows per kernel offset
    __shared__ int cnt[8];
    const int t = threadIdx.x, lane = t & 31;
    const int tile0 = blockIdx.x * T;
    for (int i = t; i < 2 * CI; i += 256) sbn[i] = bn[i];
    if (t < 8) cnt[t] = 0;
    __syncthreads();
    // bucket the tile's rows by kernel offset (one shared-memory atomic per warp and offset)
    #pragma unroll
    for (int j = 0; j < T / 256; ++j) {
        const int lr = j * 256 + t, row = tile0 + lr;
        int k = -1;
        if (row < n) {
            if (parent[row] >= 0) k = koff[row];
            else {
                float4* dst = reinterpret_cast<float4*>(out + (size_t)row * CO);
                #pragma unroll
                for (int c = 0; c < CO / 4; ++c) dst[c] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
        #pragma unroll
        for (int kk = 0; kk < 8; ++kk) {
            const unsigned m = __ballot_sync(0xFFFFFFFFu, k == kk);
            if (m == 0) continue;
            int base = 0;
            if (lane == 0) base = atomicAdd(&cnt[kk], __popc(m));
            base = __shfl_sync(0xFFFFFFFFu, base, 0);
            if (k == kk) lists[kk * T + base + __popc(m & ((1u << lane) - 1u))] = (uint16_t)lr;
        }
    }
    for (int kk = 0; kk < 8; ++kk) {
        __syncthreads();                             // lists complete / previous weights consumed
        for (int i = t; i < CI * CO / 4; i += 256)
            reinterpret_cast<float4*>(sw)[i] = reinterpret_cast<const float4*>(w + (size_t)kk * CI * CO)[i];
        __syncthreads();
        const int P = cnt[kk];
        for (int i = t; i < P; i += 256) {
            const int row = tile0 + (int)lists[kk * T + i];
            const float4* src = reinterpret_cast<const float4*>(in + (size_t)parent[row] * CI);
            float2 acc[CO / 2];                          // float2 pairs: fma.rn.f32x2, two exact fp32 FMAs per issue slot
            #pragma unroll
            for (int c = 0; c < CO / 2; ++c) acc[c] = make_float2(0.0f, 0.0f);
            #pragma unroll
            for (int c0 = 0; c0 < CI; c0 += 4) {
                const float4 x = __ldg(src + c0 / 4);
                const float v[4] = {x.x, x.y, x.z, x.w};
                #pragma unroll
                for (int cc = 0; cc < 4; ++cc) {
                    const float av = fmaxf(fmaf(v[cc], sbn[c0 + cc], sbn[CI + c0 + cc]), 0.0f);
                    const float2 aa = make_float2(av, av);
                    const float* wrow = sw + (c0 + cc) * CO;
                    #pragma unroll
                    for (int q = 0; q < CO / 4; ++q) {
                        const float4 wv = *reinterpret_cast<const float4*>(wrow + 4 * q);
                        acc[2 * q + 0] = __ffma2_rn(aa, make_float2(wv.x, wv.y), acc[2 * q + 0]);
                        acc[2 * q + 1] = __ffma2_rn(aa, make_float2(wv.z, wv.w), acc[2 * q + 1]);
                    }
                }
            }
            float4* dst = reinterpret_cast<float4*>(out + (size_t)row * CO);
            #pragma unroll
            for (int c = 0; c < CO / 4; ++c) dst[c] = make_float4(acc[2 * c].x, acc[2 * c].y, acc[2 * c + 1].x, acc[2 * c + 1].y);
        }
    }
}

template <int CI, int CO>
static void launch_upconv(Ctx& cx, const ConvW& w, const float* in, const int32_t* parent, const int32_t* koff, float* out,
                          int64_t n) {
    if (n <= 0) return;
    const size_t smem = (size_t)(2 * CI + CI * CO) * 4 + (size_t)8 * 1024 * sizeof(uint16_t);
    auto kern = k_upconv<CI, CO>;
    if (smem > 48 * 1024) SQSM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<div_up(n, 1024), 256, smem, cx.st>>>(in, parent, koff, w.w, w.bn, out, (int)n);
    cx.launches++;
}

constexpr int kSmemBudget = 96 * 1024;

template <int CI, int CO, int KV, int RPT, int CPT, bool BN, bool CAT>
static void launch_conv(Ctx& cx, ConvArgs a) {
    constexpr int SLICES = CO / CPT;
    constexpr int TR = (256 / SLICES) * RPT;
    const int per_k = CI * CO * 4;
    int kc_max = (kSmemBudget - 2 * CI * 4) / per_k;
    if (kc_max < 1) kc_max = 1;
    int chunks = (KV + kc_max - 1) / kc_max;
    a.kc = (KV + chunks - 1) / chunks;
    size_t smem = (size_t)2 * CI * 4 + (size_t)a.kc * per_k;
    auto kern = k_conv<CI, CO, KV, RPT, CPT, BN, CAT>;
    if (a.n_out <= 0) return;
    if (smem > 48 * 1024)
        SQSM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBudget));
    kern<<<div_up(a.n_out, TR), 256, smem, cx.st>>>(a);
    cx.launches++;
}

// Sparse-loop variant for the layers whose rulebook is mostly empty (level 0: 18-21 % of the 27 neighbours exist; k = 8 down
// convolution out of level 0: 1.7 of 8 children; inverse convolutions: exactly one (parent, offset) pair per row).
// One thread per output row walks ONLY the row's present offsets.  Phase 1 reads the KV rulebook entries of the row
// (coalesced over the block) and compacts the present ones into a per-thread column of shared memory plus a bit mask;
// phase 2 loops over that list - a per-lane trip count (level 0: mean 5, warp maximum about 10) instead of 27 warp-wide
// iterations of which 94-97 % cannot be skipped.  The price is a per-lane kernel offset: the weight reads are no longer warp
// broadcasts.  The weight image of an offset is padded to an ODD number of 16-byte chunks (CI*CO/4 + 1), so lanes with
// different offsets fall into different bank groups and a float4 weight read costs 1-3 wavefronts instead of 8-10.
// Accumulation order per row = ascending offset, channels inside an offset in order, like k_conv (absent neighbours add
// exact zeros there): the two kernels are bit-identical.  ncu (C3): L1TEX 94-99 % busy - the per-lane weight reads - with 19-21
// of 32 lanes active.  Measured and dropped (profiles/r02_conv_knobs.md): re-dealing a block's rows to its threads by
// neighbour count so that the lanes of a warp finish together (8 -> 8: 48.4 -> 52.5 ms per C3 step), and four consecutive
// rows per thread that share the weight reads of the union of their offsets (163 registers, 12 warps per SM: 73.8 ms).  The epilogue can also write the FP16 hi / lo operands of up to two
// tensor-core consumers (same arithmetic as k_act_split_f16 / the conv_tc epilogue).
struct SpOut { __half* hi; __half* lo; const float* alpha; const float* beta; };
struct SpArgs {
    ConvArgs c;
    const int32_t* up_parent;    // UP: inverse convolution, the single pair of row r is (up_parent[r], up_koff[r])
    const int32_t* up_koff;
    SpOut sp[2];
    int* range_flag;
    const uint32_t* csr_mask;    // CSR: compact rulebook (LevelTables::nbr_mask / nbr_off / nbr_list) instead of c.nbr
    const uint32_t* csr_off;
    const int32_t* csr_list;
};

// Epilogue of 8 output channels [ch, ch + 8) of one row: residual, "did not descend" select, float32 store and the FP16
// hi / lo operands of up to two tensor-core consumers (same arithmetic as k_act_split_f16 / the conv_tc epilogue).
__device__ __forceinline__ void sp_epilogue8(const SpArgs& s, int row, int co, int ch, float (&r)[8]) {
    const ConvArgs& a = s.c;
    const size_t o = (size_t)row * co;
    const bool take_src = a.sel_flag && a.sel_flag[row] == 0;
    if (a.res) {
        const float4 e0 = *reinterpret_cast<const float4*>(a.res + o + ch), e1 = *reinterpret_cast<const float4*>(a.res + o + ch + 4);
        r[0] += e0.x; r[1] += e0.y; r[2] += e0.z; r[3] += e0.w; r[4] += e1.x; r[5] += e1.y; r[6] += e1.z; r[7] += e1.w;
    }
    if (take_src) {
        const float4 e0 = *reinterpret_cast<const float4*>(a.sel_src + o + ch), e1 = *reinterpret_cast<const float4*>(a.sel_src + o + ch + 4);
        r[0] = e0.x; r[1] = e0.y; r[2] = e0.z; r[3] = e0.w; r[4] = e1.x; r[5] = e1.y; r[6] = e1.z; r[7] = e1.w;
    }
    if (a.out) {
        *reinterpret_cast<float4*>(a.out + o + ch) = make_float4(r[0], r[1], r[2], r[3]);
        *reinterpret_cast<float4*>(a.out + o + ch + 4) = make_float4(r[4], r[5], r[6], r[7]);
    }
    #pragma unroll
    for (int u = 0; u < 2; ++u) {                        // the consumers' BN + ReLU + FP16 hi/lo split
        if (!s.sp[u].hi) continue;
        float x[8];
        float mx = 0.f;
        #pragma unroll
        for (int i = 0; i < 8; ++i) {
            x[i] = s.sp[u].alpha ? fmaxf(fmaf(r[i], s.sp[u].alpha[ch + i], s.sp[u].beta[ch + i]), 0.f) : r[i];
            mx = fmaxf(mx, fabsf(x[i]));
        }
        if (!(mx < 3.0e4f)) *s.range_flag = 1;
        uint32_t hw[4], lw[4];
        #pragma unroll
        for (int i = 0; i < 4; ++i) {
            const __half2 h = __floats2half2_rn(x[2 * i], x[2 * i + 1]);
            const float2 f = __half22float2(h);
            const __half2 l = __floats2half2_rn(x[2 * i] - f.x, x[2 * i + 1] - f.y);
            hw[i] = *reinterpret_cast<const uint32_t*>(&h);
            lw[i] = *reinterpret_cast<const uint32_t*>(&l);
        }
        *reinterpret_cast<uint4*>(s.sp[u].hi + o + ch) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
        *reinterpret_cast<uint4*>(s.sp[u].lo + o + ch) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
    }
}

template <int CI, int CO, int KV, bool BN, bool CAT, bool UP, bool CSR>
__global__ void __launch_bounds__(256) k_conv_sp(SpArgs s) {
    const ConvArgs& a = s.c;
    constexpr int WS = CI * CO + 4;                      // padded floats per offset
    constexpr bool PREF = !UP && CI <= 16;               // software prefetch of the next neighbour's features
    constexpr int CA = CAT ? CI / 2 : CI;                // channels per source
    extern __shared__ __align__(16) float smem[];
    float* sbn = smem;                                   // [2][CI]
    float* sw = smem + 2 * 64;                           // [KV][WS]
    int32_t* list = reinterpret_cast<int32_t*>(sw + KV * WS);   // [LK][256] present neighbours of thread t, column t (!CSR)

    const int t = threadIdx.x;
    // persistent blocks: the weight image is staged once per block, then the block walks 256-row groups
    if (BN) { if (t < 2 * CI) sbn[t] = a.bn[t]; }
    for (int i = t; i < KV * CI * CO / 4; i += 256) {
        const int k = i / (CI * CO / 4), r = i % (CI * CO / 4);
        *reinterpret_cast<float4*>(sw + k * WS + 4 * r) = reinterpret_cast<const float4*>(a.w)[i];
    }
    __syncthreads();
    const int n_groups = (a.n_out + 255) / 256;
  for (int grp = blockIdx.x; grp < n_groups; grp += gridDim.x) {
    const int row = grp * 256 + t;
    if (a.tile_need) {
        const int tile = grp * 2;
        const bool any = a.tile_need[tile] != 0 || ((tile + 1) * 128 < a.n_out && a.tile_need[tile + 1] != 0);
        if (!any) {                                      // skipped group: defined (zero) outputs, no gather, no FMA
            const int64_t e0 = (int64_t)grp * 256 * CO, e1 = min((int64_t)a.n_out * CO, e0 + (int64_t)256 * CO);
            if (a.out)
                for (int64_t e = e0 + 4 * t; e < e1; e += 4 * 256) *reinterpret_cast<float4*>(a.out + e) = make_float4(0.f, 0.f, 0.f, 0.f);
            #pragma unroll
            for (int q = 0; q < 2; ++q)
                if (s.sp[q].hi)
                    for (int64_t e = e0 + 8 * t; e < e1; e += 8 * 256) {
                        *reinterpret_cast<uint4*>(s.sp[q].hi + e) = make_uint4(0, 0, 0, 0);
                        *reinterpret_cast<uint4*>(s.sp[q].lo + e) = make_uint4(0, 0, 0, 0);
                    }
            continue;
        }
    }
    unsigned mask = 0;
    int cnt = 0;
    const int32_t* clist = nullptr;                      // CSR: the row's present neighbours, read straight from global memory
    if (row < a.n_out) {
        if (CSR) {
            mask = __ldg(s.csr_mask + row);
            cnt = __popc(mask);
            clist = s.csr_list + __ldg(s.csr_off + row);
        } else if (UP) {
            const int p = __ldg(s.up_parent + row);
            if (p >= 0) { list[t] = p; mask = 1u << __ldg(s.up_koff + row); cnt = 1; }
        } else {
            #pragma unroll
            for (int k = 0; k < KV; ++k) {
                const int n = __ldg(a.nbr + (size_t)k * a.n_out + row);
                if (n >= 0) { list[cnt * 256 + t] = n; mask |= 1u << k; ++cnt; }
            }
        }
    }
    // (the list column of thread t is written and read by thread t only: no barrier)

    float2 acc[CO / 2];
    #pragma unroll
    for (int c = 0; c < CO / 2; ++c) acc[c] = make_float2(0.0f, 0.0f);
    auto src_ptr = [&](int n, int c0) -> const float4* {
        if (CAT) return reinterpret_cast<const float4*>((c0 < CA ? a.inA + (size_t)n * CA + c0 : a.inB + (size_t)n * CA + (c0 - CA)));
        return reinterpret_cast<const float4*>(a.inA + (size_t)n * CI + c0);
    };
    auto nb_at = [&](int i) -> int { return CSR ? __ldg(clist + i) : list[i * 256 + t]; };
    float4 xn[PREF ? CI / 4 : 1];
    int n_ahead = 0;                                     // CSR: index of neighbour i + 2, in flight behind the features of i + 1
    if (PREF && cnt > 0) {
        const int n = nb_at(0);
        #pragma unroll
        for (int c = 0; c < CI / 4; ++c) xn[c] = __ldg(src_ptr(n, 4 * c));
        if (CSR && cnt > 1) n_ahead = nb_at(1);
    }
    for (int i = 0; i < cnt; ++i) {
        float4 x[PREF ? CI / 4 : 1];
        int n_cur = 0;
        if (PREF) {
            #pragma unroll
            for (int c = 0; c < CI / 4; ++c) x[c] = xn[c];
            if (i + 1 < cnt) {                           // next neighbour's features in flight while this one is multiplied
                const int n = CSR ? n_ahead : nb_at(i + 1);
                #pragma unroll
                for (int c = 0; c < CI / 4; ++c) xn[c] = __ldg(src_ptr(n, 4 * c));
                if (CSR && i + 2 < cnt) n_ahead = nb_at(i + 2);
            }
        } else {
            n_cur = nb_at(i);
        }
        const int k = __ffs(mask) - 1;
        mask &= mask - 1;
        const float* wk = sw + k * WS;
        #pragma unroll
        for (int c0 = 0; c0 < CI; c0 += 4) {
            float4 v = PREF ? x[PREF ? c0 / 4 : 0] : __ldg(src_ptr(n_cur, c0));
            if (BN) {
                v.x = fmaxf(fmaf(v.x, sbn[c0 + 0], sbn[CI + c0 + 0]), 0.0f);
                v.y = fmaxf(fmaf(v.y, sbn[c0 + 1], sbn[CI + c0 + 1]), 0.0f);
                v.z = fmaxf(fmaf(v.z, sbn[c0 + 2], sbn[CI + c0 + 2]), 0.0f);
                v.w = fmaxf(fmaf(v.w, sbn[c0 + 3], sbn[CI + c0 + 3]), 0.0f);
            }
            #pragma unroll
            for (int cc = 0; cc < 4; ++cc) {
                const float sv = comp4(v, cc);
                const float2 ss = make_float2(sv, sv);
                const float* wrow = wk + (c0 + cc) * CO;
                #pragma unroll
                for (int q = 0; q < CO / 4; ++q) {
                    const float4 wv = *reinterpret_cast<const float4*>(wrow + 4 * q);
                    acc[2 * q + 0] = __ffma2_rn(ss, make_float2(wv.x, wv.y), acc[2 * q + 0]);
                    acc[2 * q + 1] = __ffma2_rn(ss, make_float2(wv.z, wv.w), acc[2 * q + 1]);
                }
            }
        }
    }
    if (row >= a.n_out) continue;
    #pragma unroll
    for (int q = 0; q < CO / 8; ++q) {
        float r[8] = {acc[4 * q].x, acc[4 * q].y, acc[4 * q + 1].x, acc[4 * q + 1].y,
                      acc[4 * q + 2].x, acc[4 * q + 2].y, acc[4 * q + 3].x, acc[4 * q + 3].y};
        sp_epilogue8(s, row, CO, 8 * q, r);
    }
  }
}

template <int CI, int CO, int KV, bool BN, bool CAT, bool UP, bool CSR = false>
static void launch_conv_sp(Ctx& cx, const SpArgs& s) {
    if (s.c.n_out <= 0) return;
    const size_t smem = (size_t)(128 + KV * (CI * CO + 4)) * 4 + (size_t)(CSR ? 0 : (UP ? 1 : KV)) * 256 * 4;
    auto kern = k_conv_sp<CI, CO, KV, BN, CAT, UP, CSR>;
    static int per_sm = 0;
    if (!per_sm) {
        SQSM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        SQSM_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 256, smem));
        if (per_sm < 1) per_sm = 1;
    }
    const int grid = (int)std::min<int64_t>(div_up(s.c.n_out, 256), (int64_t)kNumSMs * per_sm);
    kern<<<grid, 256, smem, cx.st>>>(s);
    cx.launches++;
}

// shapes served by the sparse-loop kernel
static bool conv_sparse_supported(const ConvW& w, bool cat, bool up) {
    const bool bn = w.bn != nullptr;
    // inverse convolutions: measured per C3 step 7.2 vs 24.2 ms (16 -> 8) and 18.7 vs 26.2 ms (32 -> 16) against the tensor-core
    // gather conv, but 36.0 vs 23.0 ms at 64 -> 32 (512 per-lane float4 weight reads per row): that one stays on the tensor cores
    if (up) return bn && w.kv == 8 && ((w.ci == 16 && w.co == 8) || (w.ci == 32 && w.co == 16) ||
                                       (w.ci == 64 && w.co == 32 && getenv("SQSM_UP_SPARSE_ALL") != nullptr));
    if (w.kv == 27 && w.co == 8) return (w.ci == 4 && !bn && !cat) || (w.ci == 8 && bn && !cat) || (w.ci == 16 && bn && cat);
    return w.kv == 8 && w.ci == 8 && w.co == 16 && bn && !cat;
}

static bool sparse_loop_on();
bool net_level0_uses_csr(const NetWeights& nw) {
    if (!sparse_loop_on() || getenv("SQSM_NO_CSR") != nullptr) return false;
    static const char* names[5] = {"input", "L0.blk.a", "L0.blk.b", "L0.tail.a", "L0.tail.b"};
    for (int i = 0; i < 5; ++i) {
        auto it = nw.conv.find(names[i]);
        if (it == nw.conv.end()) continue;
        if (it->second.kv != 27 || !conv_sparse_supported(it->second, i == 3, false)) return false;
    }
    return true;
}

static bool conv_sparse(Ctx& cx, const ConvW& w, bool cat, bool up, const SpArgs& s) {
    if (!conv_sparse_supported(w, cat, up)) return false;
    if (up) {
        if (w.ci == 16) launch_conv_sp<16, 8, 8, true, false, true>(cx, s);
        else if (w.ci == 32) launch_conv_sp<32, 16, 8, true, false, true>(cx, s);
        else launch_conv_sp<64, 32, 8, true, false, true>(cx, s);
    } else if (w.kv == 27 && s.csr_mask) {
        if (w.ci == 4) launch_conv_sp<4, 8, 27, false, false, false, true>(cx, s);
        else if (w.ci == 8) launch_conv_sp<8, 8, 27, true, false, false, true>(cx, s);
        else launch_conv_sp<16, 8, 27, true, true, false, true>(cx, s);
    } else if (w.kv == 27) {
        if (w.ci == 4) launch_conv_sp<4, 8, 27, false, false, false>(cx, s);
        else if (w.ci == 8) launch_conv_sp<8, 8, 27, true, false, false>(cx, s);
        else launch_conv_sp<16, 8, 27, true, true, false>(cx, s);
    } else {
        launch_conv_sp<8, 16, 8, true, false, false>(cx, s);
    }
    return true;
}

// ---------------------------------------------------------------------------------- heads + contraction

struct HeadParams { HeadW w; };

__device__ __forceinline__ float softplus_t20(float x) {       // F.softplus(beta=1, threshold=20)
    return x > 20.0f ? x : log1pf(expf(x));
}

// output_layer BN+ReLU, both MLP heads, normalize, softplus, pred = dir*dist,
// p' = p + pred, d' = d + pred for interior rows (spconv_based_contraction.py:103-108).
__global__ void __launch_bounds__(256) k_heads(const float* __restrict__ feat, const uint8_t* __restrict__ sel_flag,
                                               const float* __restrict__ sel_src, const float4* __restrict__ xyz,
                                               const float4* __restrict__ disp, const int32_t* __restrict__ out_pos, int n,
                                               const __grid_constant__ HeadParams hp, float* __restrict__ newP,
                                               float* __restrict__ newD, float4* __restrict__ pred_out,
                                               float4* __restrict__ raw_out) {
    int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    if (!pred_out && !raw_out && out_pos[r] < 0) return;       // halo row: its prediction is discarded (:105-108)
    const HeadW& w = hp.w;
    const float* f = (sel_flag && sel_flag[r] == 0) ? sel_src + (size_t)r * 8 : feat + (size_t)r * 8;
    float4 f0 = *reinterpret_cast<const float4*>(f), f1 = *reinterpret_cast<const float4*>(f + 4);
    float x[8] = {f0.x, f0.y, f0.z, f0.w, f1.x, f1.y, f1.z, f1.w};
    #pragma unroll
    for (int c = 0; c < 8; ++c) x[c] = fmaxf(fmaf(x[c], w.bn0[0][c], w.bn0[1][c]), 0.0f);
    float res[2][3];
    #pragma unroll
    for (int h = 0; h < 2; ++h) {
        float y[8];
        #pragma unroll
        for (int o = 0; o < 8; ++o) {
            float s = 0.0f;
            #pragma unroll
            for (int c = 0; c < 8; ++c) s = fmaf(x[c], w.w1[h][o][c], s);
            s += w.b1[h][o];
            y[o] = fmaxf(fmaf(s, w.a1[h][o], w.c1[h][o]), 0.0f);
        }
        float z[4];
        #pragma unroll
        for (int o = 0; o < 4; ++o) {
            float s = 0.0f;
            #pragma unroll
            for (int c = 0; c < 8; ++c) s = fmaf(y[c], w.w2[h][o][c], s);
            s += w.b2[h][o];
            z[o] = fmaxf(fmaf(s, w.a2[h][o], w.c2[h][o]), 0.0f);
        }
        #pragma unroll
        for (int o = 0; o < 3; ++o) {
            float s = 0.0f;
            #pragma unroll
            for (int c = 0; c < 4; ++c) s = fmaf(z[c], w.w3[h][o][c], s);
            res[h][o] = s + w.b3[h][o];
        }
    }
    float ox = res[0][0], oy = res[0][1], oz = res[0][2], reg = res[1][0];
    float nrm = sqrtf(ox * ox + oy * oy + oz * oz);
    float inv = 1.0f / fmaxf(nrm, 1e-12f);                     // F.normalize eps
    float dist = softplus_t20(reg);
    float px = ox * inv * dist, py = oy * inv * dist, pz = oz * inv * dist;
    if (pred_out) pred_out[r] = make_float4(px, py, pz, dist);
    if (raw_out) raw_out[r] = make_float4(ox, oy, oz, reg);
    int o = out_pos[r];
    if (o >= 0) {
        float4 p = xyz[r], d = disp[r];
        newP[3 * (size_t)o + 0] = p.x + px; newP[3 * (size_t)o + 1] = p.y + py; newP[3 * (size_t)o + 2] = p.z + pz;
        newD[3 * (size_t)o + 0] = px + d.x; newD[3 * (size_t)o + 1] = py + d.y; newD[3 * (size_t)o + 2] = pz + d.z;
    }
}

// ---------------------------------------------------------------------------------- weights

std::vector<float> make_tc_weight_image(const float* w_kcico, int kv, int ci, int co);   // conv_tc.cu
std::vector<float> make_tc_weight_image_f16(const float* w_kcico, int kv, int ci, int co);
std::vector<float> make_tma_weight_image(const float* w_kcico, int kv, int ci, int co, int cs, int g);   // conv_tma.cu
bool conv_tma_supported(int ci, int co, int kv, bool cat);
int tma_group(int cs, int kv);
void act_split_rows(Ctx& cx, const float* in, const float* bn, int c, int bn_off, int bn_stride, int64_t n, void* out, int* range_flag);
bool conv_tma(Ctx& cx, int ci, int co, int kv, bool cat, const void* splitA, const void* splitB, const int32_t* nbr, const float* wimg,
              const float* res, const float* sel_src, const uint8_t* sel_flag, float* out, int64_t n_out, int64_t n_in);

static const std::vector<float>& need(const std::map<std::string, std::vector<float>>& st, const std::string& k,
                                      size_t n) {
    auto it = st.find(k);
    if (it == st.end()) throw CudaError("missing weight tensor: " + k);
    if (it->second.size() != n) throw CudaError("wrong element count for weight tensor: " + k);
    return it->second;
}

static void fold_bn(const std::map<std::string, std::vector<float>>& st, const std::string& p, int c, float* alpha,
                    float* beta) {
    const auto& w = need(st, p + ".weight", c);
    const auto& b = need(st, p + ".bias", c);
    const auto& m = need(st, p + ".running_mean", c);
    const auto& v = need(st, p + ".running_var", c);
    for (int i = 0; i < c; ++i) {                               // BatchNorm1d(eps=1e-4), eval (App. A)
        volatile float invstd = 1.0f / std::sqrt(v[i] + 1e-4f);
        volatile float al = invstd * w[i];
        volatile float mb = m[i] * al;
        alpha[i] = al;
        beta[i] = b[i] - mb;
    }
}

void net_upload(Ctx& cx, NetWeights& nw, const std::map<std::string, std::vector<float>>& st, bool down_flip) {
    std::vector<float> blob;
    struct Pending { std::string name; size_t w_off, bn_off, img_off, img16_off; bool has_bn, has_img; int ci, co, kv; size_t tma_off = 0; bool has_tma = false; };
    std::vector<Pending> pend;
    auto add_conv = [&](const std::string& name, const std::string& wkey, const std::string& bnkey, int ci, int co, int kv,
                        int ci_pad, bool flip, bool cat = false) {
        const auto& W = need(st, wkey, (size_t)co * kv * ci);   // [co][k][ci]
        Pending p{name, blob.size(), 0, 0, 0, !bnkey.empty(), false, ci_pad, co, kv};
        blob.resize(blob.size() + (size_t)kv * ci_pad * co, 0.0f);
        for (int k = 0; k < kv; ++k)
            for (int c = 0; c < ci; ++c)
                for (int o = 0; o < co; ++o) {
                    int ks = flip ? (kv - 1 - k) : k;
                    blob[p.w_off + ((size_t)k * ci_pad + c) * co + o] = W[((size_t)o * kv + ks) * ci + c];
                }
        if (p.has_bn) {
            p.bn_off = blob.size();
            blob.resize(blob.size() + 2 * (size_t)ci_pad, 0.0f);
            fold_bn(st, bnkey, ci, &blob[p.bn_off], &blob[p.bn_off + ci_pad]);
        }
        while (blob.size() % 4) blob.push_back(0.0f);
        if (ci_pad % 8 == 0 && co <= 64) {                      // tensor-core weight image (conv_tc.cu)
            std::vector<float> img = make_tc_weight_image(&blob[p.w_off], kv, ci_pad, co);
            while (blob.size() % 64) blob.push_back(0.0f);      // 256-byte aligned
            p.img_off = blob.size();
            p.has_img = true;
            blob.insert(blob.end(), img.begin(), img.end());
            std::vector<float> img16 = make_tc_weight_image_f16(&blob[p.w_off], kv, ci_pad, co);
            while (blob.size() % 64) blob.push_back(0.0f);
            p.img16_off = blob.size();
            blob.insert(blob.end(), img16.begin(), img16.end());
        }
        if (conv_tma_supported(ci_pad, co, kv, cat)) {           // TMA-gather weight image (conv_tma.cu)
            const int cs = cat ? ci_pad / 2 : ci_pad;
            std::vector<float> img = make_tma_weight_image(&blob[p.w_off], kv, ci_pad, co, cs, tma_group(cs, kv));
            while (blob.size() % 64) blob.push_back(0.0f);
            p.tma_off = blob.size();
            p.has_tma = true;
            blob.insert(blob.end(), img.begin(), img.end());
        }
        pend.push_back(p);
    };
    add_conv("input", "input_conv.0.weight", "", 3, kPlanes[0], 27, 4, false);
    std::string prefix = "unet";
    for (int l = 0; l < kLevels; ++l) {
        int c = kPlanes[l];
        std::string L = "L" + std::to_string(l);
        add_conv(L + ".blk.a", prefix + ".blocks.block0.conv_branch.2.weight", prefix + ".blocks.block0.conv_branch.0", c, c, 27, c, false);
        add_conv(L + ".blk.b", prefix + ".blocks.block0.conv_branch.5.weight", prefix + ".blocks.block0.conv_branch.3", c, c, 27, c, false);
        if (l + 1 < kLevels) {
            int cn = kPlanes[l + 1];
            add_conv(L + ".down", prefix + ".conv.2.weight", prefix + ".conv.0", c, cn, 8, c, down_flip);
            add_conv(L + ".up", prefix + ".deconv.2.weight", prefix + ".deconv.0", cn, c, 8, cn, down_flip);
            add_conv(L + ".tail.i", prefix + ".blocks_tail.block0.i_branch.0.weight", "", 2 * c, c, 1, 2 * c, false, true);
            add_conv(L + ".tail.a", prefix + ".blocks_tail.block0.conv_branch.2.weight", prefix + ".blocks_tail.block0.conv_branch.0", 2 * c, c, 27, 2 * c, false, true);
            add_conv(L + ".tail.b", prefix + ".blocks_tail.block0.conv_branch.5.weight", prefix + ".blocks_tail.block0.conv_branch.3", c, c, 27, c, false);
        }
        prefix += ".u";
    }
    nw.blob.alloc((int64_t)blob.size() + 4, cx.st);
    nw.range_flag = reinterpret_cast<int*>(nw.blob.p + blob.size());     // FP16 range guard (last word of the blob)
    SQSM_CUDA(cudaMemsetAsync(nw.range_flag, 0, sizeof(int), cx.st));
    SQSM_CUDA(cudaMemcpyAsync(nw.blob.p, blob.data(), blob.size() * sizeof(float), cudaMemcpyHostToDevice, cx.st));
    SQSM_CUDA(cudaStreamSynchronize(cx.st));
    nw.conv.clear();
    for (const auto& p : pend) {
        ConvW c;
        c.ci = p.ci; c.co = p.co; c.kv = p.kv;
        c.w = nw.blob.p + p.w_off;
        c.bn = p.has_bn ? nw.blob.p + p.bn_off : nullptr;
        c.wimg = p.has_img ? nw.blob.p + p.img_off : nullptr;
        c.wimg16 = p.has_img ? nw.blob.p + p.img16_off : nullptr;
        c.wimg_tma = p.has_tma ? nw.blob.p + p.tma_off : nullptr;
        nw.conv[p.name] = c;
    }
    // heads (App. A #48-51)
    HeadW& h = nw.head;
    std::memset(&h, 0, sizeof(h));
    fold_bn(st, "output_layer.0", 8, h.bn0[0], h.bn0[1]);
    const char* heads[2] = {"offset_linear", "regression_linear"};
    for (int i = 0; i < 2; ++i) {
        std::string p = std::string(heads[i]) + ".sequence.";
        const auto& w1 = need(st, p + "0.weight", 64); const auto& b1 = need(st, p + "0.bias", 8);
        const auto& w2 = need(st, p + "3.weight", 32); const auto& b2 = need(st, p + "3.bias", 4);
        int last = (i == 0) ? 3 : 1;
        const auto& w3 = need(st, p + "6.weight", (size_t)last * 4); const auto& b3 = need(st, p + "6.bias", last);
        std::memcpy(h.w1[i], w1.data(), 64 * 4); std::memcpy(h.b1[i], b1.data(), 8 * 4);
        std::memcpy(h.w2[i], w2.data(), 32 * 4); std::memcpy(h.b2[i], b2.data(), 4 * 4);
        std::memcpy(h.w3[i], w3.data(), (size_t)last * 16); std::memcpy(h.b3[i], b3.data(), (size_t)last * 4);
        fold_bn(st, p + "1", 8, h.a1[i], h.c1[i]);
        fold_bn(st, p + "4", 4, h.a2[i], h.c2[i]);
    }
    nw.ready = true;
}

// ---------------------------------------------------------------------------------- forward

template <int CI, int CO, int KV, bool BN, bool CAT>
static void conv_dispatch(Ctx& cx, const ConvArgs& a) {
    if constexpr (CO == 8)       launch_conv<CI, CO, KV, 2, 8, BN, CAT>(cx, a);
    else if constexpr (CO == 16) launch_conv<CI, CO, KV, 2, 16, BN, CAT>(cx, a);
    else                         launch_conv<CI, CO, KV, 4, 16, BN, CAT>(cx, a);
}

// conv_tc.cu
bool conv_tc(Ctx& cx, int ci, int co, int kv, bool cat, const float* inA, const float* inB, const float* inA_lo,
             const float* inB_lo, const int32_t* nbr, const float* wimg, const float* res, const float* sel_src,
             const uint8_t* sel_flag, float* out, int64_t n_out, bool precise, bool f16, const TcFused* fused = nullptr,
             int* range_flag = nullptr, const int32_t* up_parent = nullptr, const int32_t* up_koff = nullptr,
             const int32_t* tile_list = nullptr, const int32_t* tile_count = nullptr);
void act_split(Ctx& cx, const float* in, const float* bn, int c, int bn_off, int bn_stride, int64_t n, float* hi, float* lo);
void act_split_f16(Ctx& cx, const float* in, const float* bn, int c, int bn_off, int bn_stride, int64_t n, void* hi, void* lo,
                   int* range_flag);
std::vector<float> make_tc_weight_image_f16(const float* w_kcico, int kv, int ci, int co);

static bool conv_ffma(Ctx& cx, const ConvW& w, bool cat, const ConvArgs& a) {
    const bool bn = w.bn != nullptr;
#define SQSM_FF_CASE(CI_, CO_, KV_, BN_, CAT_)                                                    \
    if (w.ci == CI_ && w.co == CO_ && w.kv == KV_ && bn == BN_ && cat == CAT_) {                  \
        conv_dispatch<CI_, CO_, KV_, BN_, CAT_>(cx, a);                                           \
        return true;                                                                              \
    }
    SQSM_FF_CASE(4, 8, 27, false, false)
    SQSM_FF_CASE(8, 8, 27, true, false) SQSM_FF_CASE(16, 16, 27, true, false) SQSM_FF_CASE(32, 32, 27, true, false)
    SQSM_FF_CASE(64, 64, 27, true, false)
    SQSM_FF_CASE(16, 8, 27, true, true) SQSM_FF_CASE(32, 16, 27, true, true) SQSM_FF_CASE(64, 32, 27, true, true)
    SQSM_FF_CASE(16, 8, 1, false, true) SQSM_FF_CASE(32, 16, 1, false, true) SQSM_FF_CASE(64, 32, 1, false, true)
    SQSM_FF_CASE(8, 16, 8, true, false) SQSM_FF_CASE(16, 32, 8, true, false) SQSM_FF_CASE(32, 64, 8, true, false)
    SQSM_FF_CASE(16, 8, 8, true, false) SQSM_FF_CASE(32, 16, 8, true, false) SQSM_FF_CASE(64, 32, 8, true, false)
#undef SQSM_FF_CASE
    return false;
}

// One convolution of the network: tensor-core path for C_out >= 16 (inputs pre-activated once), FFMA otherwise.
// Activation split handed from the convolution that produces a tensor to its consumers: the producer's epilogue applies
// the consumer's BatchNorm + ReLU and writes the FP16 hi / lo operands itself (conv_tc.cu; bulk stores of whole tiles), so
// the separate k_act_split_f16 pass and - when nothing else reads the tensor - its float32 round trip disappear.
struct FusedSplit {
    DBuf<float> hi, lo;
    bool ready = false;
};

static bool tc_f16_layer(const NetWeights& nw, const ConvW& w) { return nw.mode == 3 && w.wimg16 && w.co >= nw.tc_min_co; }

// tensor cores from tc_min_co output channels up, plus the level-0 decoder conv on the concatenated 16 channels
// (16 -> 8, 27 offsets: 51 vs 66 ms per C3 step on FFMA; the other 8-channel layers are faster on FFMA)
// SQSM_SPARSE_LOOP=0 turns the sparse-loop FFMA kernel (k_conv_sp) off: level 0 and the inverse convolutions then run as in
// round 1 / early round 2 (dense 27-offset FFMA kernel, tensor-core 16 -> 8 concat conv and inverse convolutions)
static bool sparse_loop_on() {
    const char* e = getenv("SQSM_SPARSE_LOOP");
    return !(e && e[0] == '0');
}

static bool layer_uses_tc(const NetWeights& nw, const ConvW& w, bool cat) {
    if (nw.mode == 0 || !w.wimg) return false;
    if (sparse_loop_on() && conv_sparse_supported(w, cat, false)) return false;     // level 0 and its down convolution
    return w.co >= nw.tc_min_co || (nw.mode >= 3 && cat && w.ci == 16 && w.co == 8 && w.kv == 27);
}

struct ConvOpts {
    const float* res = nullptr;          // residual added to the output
    const float* sel_src = nullptr;      // rows with sel_flag == 0 take this tensor instead ("batch did not descend")
    const uint8_t* sel_flag = nullptr;
    const FusedSplit* consume = nullptr;     // operands of the first (or only) source, already written by its producer
    const FusedSplit* consumeB = nullptr;    // ... of the second source of a concat
    FusedSplit* produce[2] = {nullptr, nullptr};     // operands to write for up to two consumers of the output
    const float* p_alpha[2] = {nullptr, nullptr};    // their BatchNorm scale / shift (nullptr: raw values)
    const float* p_beta[2] = {nullptr, nullptr};
    bool need_fp32 = true;               // false: the float32 output has no reader (only with a fused split)
    const int32_t* up_parent = nullptr;  // inverse convolution (kv = 8): rulebook from the down-sample tables
    const int32_t* up_koff = nullptr;
    const struct TileNeed* need = nullptr;   // decoder layers: only these tiles of 128 output rows are computed
    bool up_on_tc = false;               // inverse convolution on the tensor-core kernel (SQSM_UPCONV_TC=1) instead of k_conv_sp
    const LevelTables* csr = nullptr;    // level whose compact rulebook (nbr_mask / nbr_off / nbr_list) replaces `nbr`
};

// Which 128-row tiles of a level can influence an interior output (see need_tiles() below).
struct TileNeed {
    DBuf<uint8_t> flag;                  // [n_tiles] 1 = compute
    DBuf<int32_t> list;                  // needed tile numbers, ascending
    DBuf<int32_t> count;                 // [2]: needed tiles, skipped tiles
    int n_tiles = 0;
    bool valid = false;
};

// skipped tiles get defined contents: zeros (fp32 output and operand arrays alike)
__global__ void __launch_bounds__(128) k_zero_tiles(const uint8_t* __restrict__ flag, int64_t n_rows, int co, float* __restrict__ out,
                                                    uint2* __restrict__ h0, uint2* __restrict__ l0, uint2* __restrict__ h1,
                                                    uint2* __restrict__ l1) {
    if (flag[blockIdx.x]) return;
    const int64_t row = (int64_t)blockIdx.x * 128 + threadIdx.x;
    if (row >= n_rows) return;
    const int c4 = co / 4;
    for (int q = 0; q < c4; ++q) {
        const size_t e = (size_t)row * c4 + q;
        if (out) reinterpret_cast<float4*>(out)[e] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (h0) { h0[e] = make_uint2(0, 0); l0[e] = make_uint2(0, 0); }
        if (h1) { h1[e] = make_uint2(0, 0); l1[e] = make_uint2(0, 0); }
    }
}

static void conv_any(Ctx& cx, const NetWeights& nw, const std::string& name, bool cat, const float* inA, const float* inB,
                     const int32_t* nbr, float* out, int64_t n_out, int64_t n_in, const ConvOpts& op = ConvOpts()) {
    const ConvW& w = nw.conv.at(name);
    for (int q = 0; q < 2; ++q) if (op.produce[q]) op.produce[q]->ready = false;
    if (n_out <= 0) return;
    const bool up = op.up_parent != nullptr;
    const bool sparse = sparse_loop_on() && conv_sparse_supported(w, cat, up) && !(up && op.up_on_tc);
    const bool use_tc = !sparse && (up ? (nw.mode == 3 && w.wimg16) : layer_uses_tc(nw, w, cat));
    // per-kernel scope: algorithmic bytes of SURVEY 8(d): features in + out, rulebook entries, residual re-read
    char fam[64];
    snprintf(fam, sizeof(fam), "conv_%s%s_ci%d_co%d_k%d", use_tc ? "tc" : "ffma", up ? "_up" : "", w.ci, w.co, w.kv);
    const double alg_bytes = 4.0 * ((double)n_in * w.ci + (double)n_out * w.co) +
                             4.0 * (up ? 2.0 * n_out : ((nbr || op.csr) ? (double)w.kv * n_out : 0.0)) + (op.res ? 4.0 * (double)n_out * w.co : 0.0);
    static const char* tma_only = getenv("SQSM_TMA_ONLY");        // debug: restrict the TMA kernel to layers whose name contains this
    if (use_tc && !up && nw.mode == 4 && w.wimg_tma && conv_tma_supported(w.ci, w.co, w.kv, cat) &&
        (!tma_only || name.find(tma_only) != std::string::npos)) {
        // activation + FP16 hi/lo split into "split rows", then the TMA-gather tensor-core kernel
        const int ca = cat ? w.ci / 2 : w.ci;
        const int64_t halves = (n_in + kTmZeroRows) * 2 * ca;
        DBuf<float> sa((halves + 1) / 2, cx.st), sb;
        {
            ProfScope pa(cx, "act_split", (double)n_in * w.ci * 8.0, 0.0);
            act_split_rows(cx, inA, w.bn, ca, 0, w.ci, n_in, sa.p, nw.range_flag);
            if (cat) {
                sb.alloc((halves + 1) / 2, cx.st);
                act_split_rows(cx, inB, w.bn, ca, ca, w.ci, n_in, sb.p, nw.range_flag);
            }
        }
        snprintf(fam, sizeof(fam), "conv_tma_ci%d_co%d_k%d", w.ci, w.co, w.kv);
        ProfScope ps(cx, fam, alg_bytes, 0.0);
        if (conv_tma(cx, w.ci, w.co, w.kv, cat, sa.p, sb.p, nbr, w.wimg_tma, op.res, op.sel_src, op.sel_flag, out, n_out, n_in)) return;
        throw CudaError("conv_tma: unsupported shape for layer " + name);
    }
    if (use_tc) {
        // activation + hi/lo split once per element, then the tensor-core implicit GEMM
        const bool f16 = nw.mode == 3 || nw.mode == 4;
        const bool precise = nw.mode == 1 || f16;
        const int ca = cat ? w.ci / 2 : w.ci;
        // FP16 operands take half the space of TF32 ones: allocate in float units either way
        const int64_t elems = f16 ? (n_in * ca + 1) / 2 : n_in * ca;
        DBuf<float> a_hi, a_lo, b_hi, b_lo;
        const float *pa_hi = nullptr, *pa_lo = nullptr, *pb_hi = nullptr, *pb_lo = nullptr;
        const bool preA = op.consume && op.consume->ready && f16;
        const bool preB = cat && op.consumeB && op.consumeB->ready && f16;
        if (preA) { pa_hi = op.consume->hi.p; pa_lo = op.consume->lo.p; }
        if (preB) { pb_hi = op.consumeB->hi.p; pb_lo = op.consumeB->lo.p; }
        if (!preA || (cat && !preB)) {
            ProfScope pa(cx, "act_split", (double)n_in * ca * ((preA ? 0 : 1) + (cat && !preB ? 1 : 0)) * (f16 ? 8.0 : 4.0 * (precise ? 3 : 2)), 0.0);
            if (!preA) {
                a_hi.alloc(elems, cx.st);
                if (precise) a_lo.alloc(elems, cx.st);
                if (f16) act_split_f16(cx, inA, w.bn, ca, 0, w.ci, n_in, a_hi.p, a_lo.p, nw.range_flag);
                else act_split(cx, inA, w.bn, ca, 0, w.ci, n_in, a_hi.p, a_lo.p);
                pa_hi = a_hi.p;
                pa_lo = a_lo.p;
            }
            if (cat && !preB) {
                b_hi.alloc(elems, cx.st);
                if (precise) b_lo.alloc(elems, cx.st);
                if (f16) act_split_f16(cx, inB, w.bn, ca, ca, w.ci, n_in, b_hi.p, b_lo.p, nw.range_flag);
                else act_split(cx, inB, w.bn, ca, ca, w.ci, n_in, b_hi.p, b_lo.p);
                pb_hi = b_hi.p;
                pb_lo = b_lo.p;
            }
        }
        // fused splits for the consumers (FP16 mode only)
        TcFused fused[2];
        bool any_fused = false;
        if (f16 && nw.mode == 3) {
            for (int q = 0; q < 2; ++q) {
                if (!op.produce[q]) continue;
                const int64_t oe = (n_out * w.co + 1) / 2;
                op.produce[q]->hi.alloc(oe, cx.st);
                op.produce[q]->lo.alloc(oe, cx.st);
                op.produce[q]->ready = true;
                fused[q].hi = op.produce[q]->hi.p; fused[q].lo = op.produce[q]->lo.p;
                fused[q].alpha = op.p_alpha[q]; fused[q].beta = op.p_beta[q];
                any_fused = true;
            }
        }
        float* o_f32 = (any_fused && !op.need_fp32) ? nullptr : out;
        ProfScope ps(cx, fam, alg_bytes, 0.0);
        const bool skip = op.need && op.need->valid && f16 && w.co <= 32;       // bulk-epilogue kernels only
        if (skip) {
            k_zero_tiles<<<op.need->n_tiles, 128, 0, cx.st>>>(op.need->flag.p, n_out, w.co, o_f32, (uint2*)fused[0].hi, (uint2*)fused[0].lo,
                                                              (uint2*)fused[1].hi, (uint2*)fused[1].lo);
            cx.launches++;
        }
        if (conv_tc(cx, w.ci, w.co, w.kv, cat, pa_hi, pb_hi, pa_lo, pb_lo, nbr, f16 ? w.wimg16 : w.wimg, op.res, op.sel_src,
                    op.sel_flag, o_f32, n_out, precise, f16, any_fused ? fused : nullptr, nw.range_flag, op.up_parent, op.up_koff,
                    skip ? op.need->list.p : nullptr, skip ? op.need->count.p : nullptr))
            return;
        for (int q = 0; q < 2; ++q) if (op.produce[q]) op.produce[q]->ready = false;
        if (up) throw CudaError("no tensor-core kernel for inverse convolution " + name);
    }
    ConvArgs ar;
    ar.inA = inA; ar.inB = inB; ar.nbr = nbr; ar.w = w.w; ar.bn = w.bn; ar.res = op.res; ar.sel_src = op.sel_src;
    ar.sel_flag = op.sel_flag; ar.out = out; ar.n_out = (int)n_out; ar.kc = 0;
    ar.tile_need = (op.need && op.need->valid) ? op.need->flag.p : nullptr;
    if (sparse) {
        SpArgs sa;
        std::memset(&sa, 0, sizeof(sa));
        sa.up_parent = op.up_parent; sa.up_koff = op.up_koff; sa.range_flag = nw.range_flag;
        if (op.csr && op.csr->nbr_mask.p && w.kv == 27) {
            sa.csr_mask = op.csr->nbr_mask.p; sa.csr_off = op.csr->nbr_off.p; sa.csr_list = op.csr->nbr_list.p;
        }
        if (w.kv == 27 && !sa.csr_mask && !nbr) throw CudaError("layer " + name + ": no rulebook (neither dense nor compact)");
        bool any_fused = false;
        if (nw.mode == 3) {
            for (int q = 0; q < 2; ++q) {
                if (!op.produce[q]) continue;
                const int64_t oe = (n_out * w.co + 1) / 2;
                op.produce[q]->hi.alloc(oe, cx.st);
                op.produce[q]->lo.alloc(oe, cx.st);
                op.produce[q]->ready = true;
                sa.sp[q].hi = reinterpret_cast<__half*>(op.produce[q]->hi.p);
                sa.sp[q].lo = reinterpret_cast<__half*>(op.produce[q]->lo.p);
                sa.sp[q].alpha = op.p_alpha[q]; sa.sp[q].beta = op.p_beta[q];
                any_fused = true;
            }
        }
        if (any_fused && !op.need_fp32) ar.out = nullptr;
        sa.c = ar;
        ProfScope ps(cx, fam, alg_bytes, 0.0);
        if (conv_sparse(cx, w, cat, up, sa)) return;
        throw CudaError("no sparse-loop kernel for layer " + name);
    }
    if (up) throw CudaError("inverse convolution " + name + ": neither the sparse-loop nor the tensor-core kernel applies");
    ProfScope ps(cx, fam, alg_bytes, 0.0);
    if (!conv_ffma(cx, w, cat, ar)) throw CudaError("no convolution kernel for layer " + name);
}

static const char* lvl_family(int level) {
    return level == 0 ? "subm_l0" : level == 1 ? "subm_l1" : level == 2 ? "subm_l2" : "subm_l3";
}

static void keep_copy(Ctx& cx, ForwardBuffers& fb, const std::string& name, const float* src, int64_t n, int ch,
                      int level) {
    DBuf<float> b(n * ch, cx.st);
    SQSM_CUDA(cudaMemcpyAsync(b.p, src, sizeof(float) * n * ch, cudaMemcpyDeviceToDevice, cx.st));
    fb.keep[name] = std::move(b);
    fb.keep_meta[name] = {level, ch};
}

// ---------------------------------------------------------------------------------- decoder tile skipping
//
// A cube sample carries a 0.4 m halo (40 voxels) whose rows only exist to give the interior rows their receptive field
// (synthetic_tree.py:262-272); their own predictions are thrown away (spconv_based_contraction.py:104-108).  On the way DOWN
// the halo is needed almost entirely (the receptive field of the U-Net is ~45 level-0 voxels), but on the way UP a row matters
// only if an interior row can still see it: a 3^3 convolution at level l reaches one level-l voxel, an inverse convolution
// reads the parent only.  With I = the per-axis interval of voxel coordinates that interior rows occupy (measured on the
// rows of the pass, identical frame for every cube sample), the outputs of the decoder layers are needed inside
//     level 0: I +- 2          level 1: (level 0 >> 1) +- 2          level 2: (level 1 >> 1) +- 2
// (tail.b needs I, tail.a one voxel more, the inverse convolution and the skip operands one more; parents of a set [a, b]
// are [a >> 1, b >> 1]).  Tiles of 128 rows (brick order: spatially compact) without any row inside the interval are not
// computed - their outputs are zero-filled, so every tensor stays fully defined - which removes 20-40 % of the decoder
// work.  The result is bit-identical: a computed row sees exactly the inputs it saw before (tests/test_gpu_parity.py).

__global__ void __launch_bounds__(256) k_interior_bounds(const Brick* __restrict__ bricks, const int32_t* __restrict__ row_brick,
                                                         const uint16_t* __restrict__ row_pos, const uint8_t* __restrict__ interior,
                                                         int64_t n, int32_t* __restrict__ bounds) {
    // grid-stride: a few thousand atomics in total (one set per block), not one set per warp on the same six addresses
    __shared__ int s_lo[8][3], s_hi[8][3];
    int lo[3] = {1 << 30, 1 << 30, 1 << 30}, hi[3] = {-1, -1, -1};
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n; r += (int64_t)gridDim.x * blockDim.x) {
        if (!interior[r]) continue;
        const uint64_t key = bricks[row_brick[r]].key;
        const uint32_t pos = row_pos[r];
        const int c[3] = {(int)key_z(key) * 8 + (int)(pos >> 6), (int)key_y(key) * 8 + (int)((pos >> 3) & 7), (int)key_x(key) * 8 + (int)(pos & 7)};
        #pragma unroll
        for (int a = 0; a < 3; ++a) { lo[a] = min(lo[a], c[a]); hi[a] = max(hi[a], c[a]); }
    }
    const int w = threadIdx.x >> 5;
    #pragma unroll
    for (int a = 0; a < 3; ++a) {
        lo[a] = __reduce_min_sync(0xFFFFFFFFu, lo[a]);
        hi[a] = __reduce_max_sync(0xFFFFFFFFu, hi[a]);
        if ((threadIdx.x & 31) == 0) { s_lo[w][a] = lo[a]; s_hi[w][a] = hi[a]; }
    }
    __syncthreads();
    if (threadIdx.x < 3) {
        const int a = threadIdx.x;
        int l = s_lo[0][a], h = s_hi[0][a];
        for (int i = 1; i < 8; ++i) { l = min(l, s_lo[i][a]); h = max(h, s_hi[i][a]); }
        if (h >= 0) { atomicMin(&bounds[a], l); atomicMax(&bounds[3 + a], h); }
    }
}

// one block = one tile of 128 rows: flag = any row with all coordinates inside the level's interval
__global__ void __launch_bounds__(128) k_tile_need(const Brick* __restrict__ bricks, const int32_t* __restrict__ row_brick,
                                                   const uint16_t* __restrict__ row_pos, int64_t n, const int32_t* __restrict__ bounds,
                                                   int level, uint8_t* __restrict__ flag, int32_t* __restrict__ flag32) {
    const int64_t r = (int64_t)blockIdx.x * 128 + threadIdx.x;
    int lo[3], hi[3];
    #pragma unroll
    for (int a = 0; a < 3; ++a) {
        lo[a] = bounds[a] - 2;
        hi[a] = bounds[3 + a] + 2;
        for (int l = 0; l < level; ++l) { lo[a] = (lo[a] >> 1) - 2; hi[a] = (hi[a] >> 1) + 2; }
    }
    bool in = false;
    if (r < n) {
        const uint64_t key = bricks[row_brick[r]].key;
        const uint32_t pos = row_pos[r];
        const int c[3] = {(int)key_z(key) * 8 + (int)(pos >> 6), (int)key_y(key) * 8 + (int)((pos >> 3) & 7), (int)key_x(key) * 8 + (int)(pos & 7)};
        in = c[0] >= lo[0] && c[0] <= hi[0] && c[1] >= lo[1] && c[1] <= hi[1] && c[2] >= lo[2] && c[2] <= hi[2];
    }
    const int any = __syncthreads_or(in ? 1 : 0);
    if (threadIdx.x == 0) { flag[blockIdx.x] = any ? 1 : 0; flag32[blockIdx.x] = any ? 1 : 0; }
}

__global__ void __launch_bounds__(256) k_tile_list(const int32_t* __restrict__ flag32, const int32_t* __restrict__ scan, int n_tiles,
                                                   int32_t* __restrict__ list, int32_t* __restrict__ count) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i > n_tiles) return;
    if (i == n_tiles) { count[0] = scan[n_tiles]; count[1] = n_tiles - scan[n_tiles]; return; }
    if (flag32[i]) list[scan[i]] = i;
}

static void need_tiles(Ctx& cx, const LevelTables& L, int level, const int32_t* d_bounds, TileNeed& tn) {
    cudaStream_t st = cx.st;
    tn.n_tiles = div_up(L.n, 128);
    tn.flag.alloc(tn.n_tiles, st);
    tn.list.alloc(tn.n_tiles, st);
    tn.count.alloc(2, st);
    DBuf<int32_t> f32((int64_t)tn.n_tiles + 1, st), scan((int64_t)tn.n_tiles + 1, st);
    SQSM_CUDA(cudaMemsetAsync(f32.p + tn.n_tiles, 0, sizeof(int32_t), st));
    k_tile_need<<<tn.n_tiles, 128, 0, st>>>(L.map.bricks.p, L.row_brick.p, L.row_pos.p, L.n, d_bounds, level, tn.flag.p, f32.p);
    exclusive_scan_i32(cx, f32.p, scan.p, (int64_t)tn.n_tiles + 1);
    k_tile_list<<<div_up(tn.n_tiles + 1, 256), 256, 0, st>>>(f32.p, scan.p, tn.n_tiles, tn.list.p, tn.count.p);
    cx.launches += 2;
    tn.valid = true;
}

void net_forward(Ctx& cx, const NetWeights& nw, LevelTables* lv, const Level0Rows& rows, int n_levels,
                 ForwardBuffers& fb, bool keep, float* d_newP, float* d_newD) {
    cudaStream_t st = cx.st;
    SQSM_CHECK(nw.ready, "weights not finalised");
    DBuf<float> enc[kLevels], tmp[kLevels];
    // FP16 split mode: producers write their consumers' operands (FusedSplit); inverse convolutions on the tensor cores
    const bool fuse = nw.mode == 3 && getenv("SQSM_NO_FUSED_SPLIT") == nullptr;
    // inverse convolutions: sparse-loop FFMA kernel (default), tensor-core gather conv (SQSM_UPCONV_TC=1, or when the sparse
    // loop is off) or the offset-bucketed FFMA kernel of round 1 (SQSM_UPCONV_FFMA=1)
    const bool up_bucketed = getenv("SQSM_UPCONV_FFMA") != nullptr;
    const bool up_all_tc = nw.mode == 3 && !up_bucketed && (getenv("SQSM_UPCONV_TC") != nullptr || !sparse_loop_on());
    auto up_sparse_at = [&](int lvl) {
        return !up_bucketed && !up_all_tc && sparse_loop_on() && conv_sparse_supported(nw.conv.at("L" + std::to_string(lvl) + ".up"), false, true);
    };
    auto up_tc_at = [&](int lvl) {
        return nw.mode == 3 && !up_bucketed && nw.conv.at("L" + std::to_string(lvl) + ".up").wimg16 && !up_sparse_at(lvl);
    };
    FusedSplit sp_in;                        // operands of the level's first convolution, written by the down convolution
    FusedSplit sp_down[kLevels];             // enc[l] as operands of the down convolution of level l
    FusedSplit sp_skip[kLevels];             // enc[l] as operands (first half of the concat) of tail.a of level l
    FusedSplit sp_cur;                       // operands of the next inverse convolution, written by the producer of `cur`
    // decoder tile skipping (not when activations are captured for the parity tests: those compare every row)
    TileNeed need[kLevels];
    const bool no_skip = getenv("SQSM_NO_TILE_SKIP") != nullptr;
    if (!keep && !no_skip && n_levels > 1) {
        ProfScope pn(cx, "need_tiles", (double)lv[0].n * 7.0, 0.0);
        DBuf<int32_t> bounds(6, st);
        static const int32_t init[6] = {1 << 30, 1 << 30, 1 << 30, -1, -1, -1};      // static: outlives the asynchronous copy
        SQSM_CUDA(cudaMemcpyAsync(bounds.p, init, sizeof(init), cudaMemcpyHostToDevice, st));
        const int grid = (int)std::min<int64_t>(div_up(lv[0].n, 256), kNumSMs * 8);
        k_interior_bounds<<<grid, 256, 0, st>>>(lv[0].map.bricks.p, lv[0].row_brick.p, lv[0].row_pos.p, rows.interior.p, lv[0].n,
                                                bounds.p);
        cx.launches++;
        // levels 0 and 1 only: at level 2 the needed interval covers all but the outermost voxels (measured: no tile skipped)
        for (int l = 0; l < n_levels - 1 && l < 2; ++l) need_tiles(cx, lv[l], l, bounds.p, need[l]);
    }
    auto bn_a = [](const ConvW& w, int off) { return w.bn ? w.bn + off : nullptr; };
    auto bn_b = [](const ConvW& w, int off) { return w.bn ? w.bn + w.ci + off : nullptr; };
    // ---- encoder (smart_tree_xx.py:65-66, sparse_module.py:101-112)
    const int64_t n0 = lv[0].n;
    DBuf<float> cur_in(n0 * 8, st);
    {
        ProfScope pt(cx, "subm_l0", 4.0 * n0 * (3 + 8) + 4.0 * 27 * n0, 2.0 * (double)lv[0].pairs * 3 * 8);
        ConvOpts oi;
        oi.csr = &lv[0];
        conv_any(cx, nw, "input", false, (const float*)rows.xyz.p, nullptr, lv[0].nbr.p, cur_in.p, n0, n0, oi);
    }
    if (keep) keep_copy(cx, fb, "input_conv", cur_in.p, n0, 8, 0);
    for (int l = 0; l < n_levels; ++l) {
        const int64_t n = lv[l].n;
        const int c = kPlanes[l];
        const std::string L = "L" + std::to_string(l);
        const bool deepest = l == n_levels - 1;
        enc[l].alloc(n * c, st);
        tmp[l].alloc(n * c, st);
        {   // ResidualBlock with identity i_branch (sparse_module.py:32-39): out = convB(bnrelu(convA(bnrelu(x)))) + x
            ProfScope pt(cx, lvl_family(l), 2.0 * (4.0 * n * c * 2 + 4.0 * 27 * n) + 4.0 * n * c,
                         2.0 * 2.0 * (double)lv[l].pairs * c * c);
            FusedSplit sp_mid;                                  // bnrelu(convA(.)) as convB's operands, written by convA
            const ConvW& wb = nw.conv.at(L + ".blk.b");
            ConvOpts oa;
            oa.consume = &sp_in;
            if (fuse && tc_f16_layer(nw, wb)) { oa.produce[0] = &sp_mid; oa.p_alpha[0] = bn_a(wb, 0); oa.p_beta[0] = bn_b(wb, 0); oa.need_fp32 = false; }
            oa.csr = &lv[l];
            conv_any(cx, nw, L + ".blk.a", false, cur_in.p, nullptr, lv[l].nbr.p, tmp[l].p, n, n, oa);
            ConvOpts ob;
            ob.res = cur_in.p;
            ob.consume = &sp_mid;
            if (fuse && deepest && l > 0 && up_tc_at(l - 1)) {
                // the deepest level feeds the inverse convolution only: its epilogue writes that layer's operands
                const ConvW& wu = nw.conv.at("L" + std::to_string(l - 1) + ".up");
                ob.produce[0] = &sp_cur; ob.p_alpha[0] = bn_a(wu, 0); ob.p_beta[0] = bn_b(wu, 0);
                ob.need_fp32 = keep;
            } else if (fuse && !deepest) {
                // the skip tensor: operands of the down convolution and of the tail block's first convolution
                const ConvW& wd = nw.conv.at(L + ".down");
                const ConvW& wa = nw.conv.at(L + ".tail.a");
                if (layer_uses_tc(nw, wd, false)) { ob.produce[0] = &sp_down[l]; ob.p_alpha[0] = bn_a(wd, 0); ob.p_beta[0] = bn_b(wd, 0); }
                if (layer_uses_tc(nw, wa, true)) { ob.produce[1] = &sp_skip[l]; ob.p_alpha[1] = bn_a(wa, 0); ob.p_beta[1] = bn_b(wa, 0); }
            }
            ob.csr = &lv[l];
            conv_any(cx, nw, L + ".blk.b", false, tmp[l].p, nullptr, lv[l].nbr.p, enc[l].p, n, n, ob);
        }
        sp_in = FusedSplit();
        if (keep) keep_copy(cx, fb, "enc" + std::to_string(l), enc[l].p, n, c, l);
        if (!deepest) {
            const int64_t nn = lv[l + 1].n;
            DBuf<float> nxt(nn * kPlanes[l + 1], st);
            ProfScope pt(cx, "down_up", 4.0 * (n * c + nn * kPlanes[l + 1]) + 4.0 * n, 2.0 * (double)n * c * kPlanes[l + 1]);
            const ConvW& wn = nw.conv.at("L" + std::to_string(l + 1) + ".blk.a");
            ConvOpts od;
            od.consume = &sp_down[l];
            // (k = 8 convolutions are bound by their epilogue: writing the next layer's operands there costs more than the
            // separate split pass saves - measured 23.6 -> 37.1 ms per C3 step against 7.5 ms of act_split)
            static const bool fuse_down = getenv("SQSM_FUSE_DOWN") != nullptr;
            // the sparse-loop kernel (level 0 -> 1) holds a whole output row per thread: there the operand stores are cheap
            const bool down_sparse = sparse_loop_on() && conv_sparse_supported(nw.conv.at(L + ".down"), false, false);
            if (fuse && (fuse_down || down_sparse) && tc_f16_layer(nw, wn)) { od.produce[0] = &sp_in; od.p_alpha[0] = bn_a(wn, 0); od.p_beta[0] = bn_b(wn, 0); }
            conv_any(cx, nw, L + ".down", false, enc[l].p, nullptr, lv[l].child.p, nxt.p, nn, n, od);
            sp_down[l] = FusedSplit();
            cur_in = std::move(nxt);
        }
    }
    // ---- decoder (sparse_module.py:112-116)
    DBuf<float> cur = std::move(enc[n_levels - 1]);
    for (int l = n_levels - 2; l >= 0; --l) {
        const int64_t n = lv[l].n, nc = lv[l + 1].n;
        const int c = kPlanes[l], cn = kPlanes[l + 1];
        const std::string L = "L" + std::to_string(l);
        const ConvW& wu = nw.conv.at(L + ".up");
        const ConvW& wi = nw.conv.at(L + ".tail.i");
        const ConvW& wa = nw.conv.at(L + ".tail.a");
        const ConvW& wb = nw.conv.at(L + ".tail.b");
        DBuf<float> up, res(n * c, st), out(n * c, st);
        FusedSplit up_raw, up_act;           // the decoder half of the concat as operands of tail.i (raw) and tail.a (BN + ReLU)
        {
            ProfScope pt(cx, "down_up", 4.0 * (nc * cn + n * c) + 4.0 * n, 2.0 * (double)n * c * cn);
            if (up_tc_at(l) || up_sparse_at(l)) {
                // SparseInverseConv3d as a gather convolution with 8 offsets: the rulebook entry (k, fine row) exists for
                // k == koff[row] only (App. A); rows without a parent gather nothing and come out as 0 / relu(beta)
                const bool i_tc = layer_uses_tc(nw, wi, true), a_tc = layer_uses_tc(nw, wa, true);
                ConvOpts ou;
                ou.consume = &sp_cur;
                ou.up_parent = lv[l].parent.p;
                ou.up_koff = lv[l].koff.p;
                if (fuse && i_tc) { ou.produce[0] = &up_raw; }
                if (fuse && a_tc) { ou.produce[1] = &up_act; ou.p_alpha[1] = bn_a(wa, c); ou.p_beta[1] = bn_b(wa, c); }
                ou.need_fp32 = !(fuse && i_tc && a_tc);
                ou.need = &need[l];
                ou.up_on_tc = up_tc_at(l);
                if (ou.need_fp32) up.alloc(n * c, st);
                conv_any(cx, nw, L + ".up", false, cur.p, nullptr, nullptr, up.p, n, nc, ou);
            } else {
                up.alloc(n * c, st);
                if (l == 0) launch_upconv<16, 8>(cx, wu, cur.p, lv[l].parent.p, lv[l].koff.p, up.p, n);
                else if (l == 1) launch_upconv<32, 16>(cx, wu, cur.p, lv[l].parent.p, lv[l].koff.p, up.p, n);
                else launch_upconv<64, 32>(cx, wu, cur.p, lv[l].parent.p, lv[l].koff.p, up.p, n);
            }
            sp_cur = FusedSplit();
        }
        {   // blocks_tail on cat(enc, up); rows of batches that did not descend keep `enc`
            ProfScope pt(cx, lvl_family(l), 4.0 * n * (2 * c + c) * 2 + 4.0 * n * (c + c) + 2.0 * 4.0 * 27 * n + 4.0 * n * c,
                         2.0 * (double)lv[l].pairs * (2 * c * c + c * c) + 2.0 * (double)n * 2 * c * c);
            ConvOpts oi;
            oi.consumeB = &up_raw;
            oi.need = &need[l];
            conv_any(cx, nw, L + ".tail.i", true, enc[l].p, up.p, nullptr, res.p, n, n, oi);
            up_raw = FusedSplit();
            FusedSplit sp_mid;
            ConvOpts oa;
            oa.consume = &sp_skip[l];
            oa.consumeB = &up_act;
            oa.need = &need[l];
            if (fuse && tc_f16_layer(nw, wb)) { oa.produce[0] = &sp_mid; oa.p_alpha[0] = bn_a(wb, 0); oa.p_beta[0] = bn_b(wb, 0); oa.need_fp32 = false; }
            oa.csr = &lv[l];
            conv_any(cx, nw, L + ".tail.a", true, enc[l].p, up.p, lv[l].nbr.p, tmp[l].p, n, n, oa);
            sp_skip[l] = FusedSplit();
            up_act = FusedSplit();
            // the level's output feeds the next inverse convolution only (levels >= 1): fused operands, no float32 copy
            ConvOpts ob;
            ob.res = res.p; ob.sel_src = enc[l].p; ob.sel_flag = lv[l].row_desc.p;
            ob.consume = &sp_mid;
            ob.need = &need[l];
            if (fuse && l > 0 && up_tc_at(l - 1)) {
                const ConvW& wn = nw.conv.at("L" + std::to_string(l - 1) + ".up");
                ob.produce[0] = &sp_cur; ob.p_alpha[0] = bn_a(wn, 0); ob.p_beta[0] = bn_b(wn, 0);
                ob.need_fp32 = keep;
            }
            ob.csr = &lv[l];
            conv_any(cx, nw, L + ".tail.b", false, tmp[l].p, nullptr, lv[l].nbr.p, out.p, n, n, ob);
        }
        if (keep) keep_copy(cx, fb, "dec" + std::to_string(l), out.p, n, c, l);
        cur = std::move(out);
    }
    // ---- heads + contraction update
    if (keep) {
        fb.pred.alloc(n0 * 4, st);
        fb.offset.alloc(n0 * 4, st);
    }
    {
        ProfScope pt(cx, "heads", (double)n0 * (32 + 32 + 4 + 24), 2.0 * (double)n0 * 2 * (64 + 32 + 12));
        HeadParams hp;
        hp.w = nw.head;
        k_heads<<<div_up(n0, 256), 256, 0, st>>>(cur.p, nullptr, nullptr, rows.xyz.p, rows.disp.p, rows.out.p, (int)n0, hp,
                                                 d_newP, d_newD, keep ? (float4*)fb.pred.p : nullptr,
                                                 keep ? (float4*)fb.offset.p : nullptr);
        cx.launches++;
    }
}

}  // namespace sqsm
