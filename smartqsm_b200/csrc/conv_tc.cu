// Tensor-core sparse convolution: implicit GEMM on tcgen05 (5th-gen tensor cores, TMEM accumulators).
// sm_100a only.
//
// Every sparse convolution of the U-Net is a GEMM  D[rows, C_out] = A[rows, KV*C_in] . W[KV*C_in, C_out]
// whose A operand is gathered through the output-stationary rulebook (nbr[k][row] = input row or -1,
// zero when absent).  The 16/32/64-channel layers (and the 16 -> 8 decoder conv) run here:
//
//   * a persistent CTA per SM walks tiles of 128 output rows;
//   * the activated input is split once per element into "hi" and "lo" operand arrays - FP16 (default: 2-term split,
//     hi = fp16(x), lo = fp16(x - hi)) or TF32 - by k_act_split*, or directly by the epilogue of the convolution that
//     produced it when this layer is its only consumer;
//   * GR producer groups of 128 threads (GR = 2 or 3) gather the K dimension in chunks of 128 bytes per row with
//     cp.async (LDGSTS, 16 B, zero-fill for absent neighbours) straight into two SWIZZLE_128B K-major shared-memory
//     tiles (hi, lo); the matching weight chunk (pre-split, pre-swizzled on the host, W_hi rows then W_lo rows) is
//     copied next to them; completion is tracked by cp.async.mbarrier.arrive, so a ring of up to 6 chunks is in flight
//     without holding registers or scoreboard slots (a register-staged gather stalls in order on the dependent
//     rulebook -> row loads; measured in profiles/); rulebook entries are prefetched 7 chunks ahead;
//   * one thread issues tcgen05.mma (M = 128, 32 bytes of K per instruction) TWICE per K step:
//     A_hi . [W_hi | W_lo] with N = 2*C_out into two halves of the TMEM accumulator, then A_lo . W_hi with N = C_out
//     into the first half (hi*hi + lo*hi + hi*lo overall: error ~2^-22 per product, fp32-class accuracy; the reduced
//     single-pass TF32 mode of App. D.4 is a template switch), and releases the stage with tcgen05.commit;
//   * four epilogue warps tcgen05.ld the accumulator (two stages, so the next tile's MMAs overlap), sum the halves, add
//     the residual, apply the "batch did not descend" select, store - and optionally write the consumer's operands.
//
// Measured and dropped in round 2 (profiles/r02_conv_knobs.md): interleaved [hi | lo] operand rows (fewer neighbouring rows per
// 128-byte line: 127 -> 146 ms per C3 step for 16 -> 16), a "lean" gather loop in which every lane loads its own rulebook
// entries instead of receiving them by shuffle (1.7 k instead of 8.8 k SASS instructions, but 8 more LSU requests per chunk
// and warp: 127 -> 142 ms - the kernel is bound by the LSU / L2 request rate, not by issue slots), contiguous tile ranges per
// CTA, cp.async.ca, shallower rings.
//
// The gather uses LDGSTS rather than TMA: conv_tma.cu is the same kernel on cp.async.bulk.tensor tile::gather4 and is
// 10-25 % slower in the pipeline (profiles/r01_gather_microbench_notes.md); the weight chunk is small and L2 resident.
#include "engine.h"
#include "tc_ptx.cuh"
#include <cuda_fp16.h>
#include <cstring>

namespace sqsm {

// ------------------------------------------------------------------------------------------ kernel

struct TcSplitOut {            // fused activation split for one consumer of this output: hi/lo = split(act(out))
    __half* hi;                // [n_out][CO] or nullptr
    __half* lo;
    const float* alpha;        // [CO] scale / shift of the consumer's BatchNorm (+ ReLU); nullptr = raw values
    const float* beta;
};

struct TcArgs {
    const float* inA;         // [n_in][CI] TF32 hi part of the activated input (or first CI/2 channels when CAT)
    const float* inB;         // hi part of the second CI/2 channels when CAT
    const float* inA_lo;      // matching lo parts (PRECISE)
    const float* inB_lo;
    const int32_t* nbr;       // [KV][n_out] or nullptr (identity, KV == 1)
    const int32_t* up_parent; // UP (inverse convolution): coarse row of every output row or -1, and its kernel offset;
    const int32_t* up_koff;   //   the rulebook entry (k, row) is parent[row] when koff[row] == k, absent otherwise
    const float* wimg;        // per chunk: hi tile [NP][32] then lo tile [NP][32], SWIZZLE_128B images
    const float* res;         // [n_out][CO] or nullptr
    const float* sel_src;     // [n_out][CO] or nullptr
    const uint8_t* sel_flag;
    float* out;               // [n_out][CO], or nullptr when only the fused splits below are wanted
    TcSplitOut sp[2];         // F16 kernels: operands of up to two consumers, written by the epilogue
    int* range_flag;
    const int32_t* tile_list; // decoder tile skipping: the tiles to compute (ascending) and their number, or nullptr = all
    const int32_t* tile_count;
    int n_out;
    int n_tiles;
};

constexpr int kTcMaxStages = 6;       // shared-memory ring (A hi/lo + B hi/lo per stage), as many as fit 227 KB
// producer groups (template parameter GR): group g gathers chunks g, g+GR, ... (thread-level parallelism: one warp per SM
// sub-partition cannot hide its own issue latencies).  Measured per C3 step: 3 groups beat 2 for the 16-channel sources
// (16->16: 137 -> 124 ms, 32->16: 88 -> 81 ms) and lose for the wider ones (64->32: 71 -> 78 ms).
constexpr int kTcProducers = 128;     // threads per producer group; thread t owns tile row t
constexpr int kTcEpilogue = 128;      // next 4 warps: TMEM -> registers -> global; thread owns one tile row
__host__ __device__ constexpr int tc_threads(int gr) { return gr * kTcProducers + kTcEpilogue + 32; }   // last warp: MMA issuer
constexpr uint32_t kTmemCols = 256;   // two accumulator stages of up to 128 columns ([hi*hi + lo*hi | hi*lo] when PRECISE)

// Bulk epilogue (FP16 kernels with C_out <= 32): the 128 output rows of a tile are contiguous in global memory, so the
// epilogue stages the float32 tile and the fused operand tiles in shared memory and one thread stores each with a single
// cp.async.bulk (TMA); the residual tile arrives the same way.  A thread-per-row epilogue costs one LSU sector operation
// per 16 bytes (32 rows x 16 B per instruction) and competes with the cp.async gather for the same L1TEX path.
__host__ __device__ constexpr bool tc_bulk(int co, bool f16, bool precise) { return f16 && precise && co <= 32; }
__host__ __device__ constexpr int tc_staging_bytes(int co, bool bulk) { return bulk ? 128 * co * 16 : 0; }   // out, res, 2 x (hi, lo)
__host__ __device__ constexpr int tc_stages(int np, int staging = 0) {
    int fit = (227 * 1024 - 2048 - staging) / (2 * 128 * 128 + 2 * np * 128);
    return fit > kTcMaxStages ? kTcMaxStages : fit;
}

// F16 = false: TF32 operands (4 B), 32 K elements per 128-byte chunk row, tcgen05 kind::tf32
// F16 = true : FP16 operands (2 B), 64 K elements per chunk row, kind::f16 - half the bytes to gather, half the
//              LDGSTS instructions and half the shared-memory operand traffic per K element; the 2-term split
//              hi = fp16(x), lo = fp16(x - hi) carries the same 22 mantissa bits as the TF32 split
template <int CI, int CO, int KV, bool CAT, bool PRECISE, bool F16, int GR, bool UP>
__global__ void __launch_bounds__(tc_threads(GR), 1) k_conv_tc(const __grid_constant__ TcArgs a) {
    constexpr int kTcGroups = GR, kTcProdWarps = GR * 4;
    constexpr int KT = KV * CI;                       // GEMM K
    constexpr int ES = F16 ? 2 : 4;                   // operand element size
    constexpr int CE = 128 / ES;                      // K elements per chunk row (128 B)
    constexpr int PE = 16 / ES;                       // K elements per 16-byte piece
    constexpr int NCH = (KT + CE - 1) / CE;           // chunks
    constexpr int NP = CO < 16 ? 16 : CO;             // UMMA N (M=128 needs N % 16 == 0)
    constexpr int A_BYTES = 128 * 128;
    constexpr int B_BYTES = NP * 128;
    constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;
    constexpr bool BULK = tc_bulk(CO, F16, PRECISE);
    constexpr int kTcStages = tc_stages(NP, tc_staging_bytes(CO, BULK));
    constexpr uint32_t FMT = F16 ? 0u : 2u;           // a/b format: F16 = 0, TF32 = 2 (UMMA::F16F32Format)
    constexpr uint32_t IDESC = (1u << 4) | (FMT << 7) | (FMT << 10) | ((uint32_t)(NP >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    // PRECISE: the hi and lo weight tiles are adjacent in shared memory ([NP rows][128 B] each), so ONE instruction with
    // N = 2*NP computes A_hi.[W_hi | W_lo] into two accumulator halves and a second one adds A_lo.W_hi to the first half:
    // two MMAs and two reads of the A tile per K step instead of three (the tiny-N MMAs are bound by the A-tile fetch)
    constexpr uint32_t IDESC2 = (1u << 4) | (FMT << 7) | (FMT << 10) | ((uint32_t)((2 * NP) >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    constexpr uint32_t ACC = PRECISE ? 2 * NP : NP;   // TMEM columns of one accumulator stage
    static_assert(CI % 8 == 0 && NP <= 64 && B_BYTES % 1024 == 0, "tile shape");

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);     // SWIZZLE_128B needs 1024 B alignment
    __shared__ __align__(8) uint64_t bars[2 * kTcMaxStages + 5];
    __shared__ uint32_t tmem_slot;

    const int t = threadIdx.x;
    const int warp = t >> 5;
    const uint32_t full0 = smem_u32(&bars[0]), empty0 = smem_u32(&bars[kTcMaxStages]);
    const uint32_t tfull0 = smem_u32(&bars[2 * kTcMaxStages]), tempty0 = smem_u32(&bars[2 * kTcMaxStages + 2]);

    if (t == 0) {
        for (int s = 0; s < kTcStages; ++s) { mbar_init(full0 + 8 * s, kTcProducers); mbar_init(empty0 + 8 * s, 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(tfull0 + 8 * s, 1); mbar_init(tempty0 + 8 * s, kTcEpilogue); }
        mbar_init(smem_u32(&bars[2 * kTcMaxStages + 4]), 1);      // residual tile landed (bulk epilogue)
        fence_barrier_init();
    }
    if (warp == kTcProdWarps + 4) tmem_alloc(smem_u32(&tmem_slot), kTmemCols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_slot;
    const int n_tiles = a.tile_count ? __ldg(a.tile_count) : a.n_tiles;       // tiles this launch computes

    if (warp < kTcProdWarps) {
        // ===================== producers: rulebook rows -> cp.async gather into the swizzled smem tiles =====================
        // Nothing but rulebook entries passes through registers: the pre-split operands go global -> shared by LDGSTS and
        // the stage's mbarrier counts their arrival.
        const int n_my_tiles = (n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
        const int total = n_my_tiles * NCH;               // chunk sequence of this CTA
        const int grp = warp >> 2;                        // producer group
        const int tp = t & (kTcProducers - 1);            // thread inside the group = tile row
        const int my_total = (total - grp + kTcGroups - 1) / kTcGroups;   // chunks grp, grp+G, ... of that sequence
        // Lane l of a warp copies piece (l & 7) of tile rows (l >> 3) + 4m, m = 0..7: the eight lanes of a row
        // move one contiguous 128 B (or 2 x 64 B) span.  Rulebook entries are loaded by the lane that owns the
        // row, RI-1 chunks ahead (the only register loads left), and travel by shuffle.
        constexpr int KPC = (CI >= CE) ? 1 : CE / CI;    // kernel offsets covered by one chunk
        constexpr int RI = 8;
        const int lane = tp & 31;
        const int piece = lane & 7;
        const uint32_t smem_base = smem_u32(smem);
        int ibuf[RI][KPC];
        auto issue_idx = [&](int i, int* nb) {            // i-th chunk of this group
            const int c = grp + i * kTcGroups;
            const int tseq = (int)blockIdx.x + (c / NCH) * (int)gridDim.x;
            const int tile = a.tile_list ? __ldg(a.tile_list + tseq) : tseq;
            const int j = c % NCH;
            const int row = tile * 128 + tp;
            #pragma unroll
            for (int q = 0; q < KPC; ++q) {
                const int k = (j * CE) / CI + q;
                int n = -1;
                if (UP) {
                    if (row < a.n_out && k < KV) {
                        const int pr = __ldg(a.up_parent + row);
                        if (pr >= 0 && __ldg(a.up_koff + row) == k) n = pr;
                    }
                } else if (row < a.n_out && k < KV) n = a.nbr ? __ldg(a.nbr + (size_t)k * a.n_out + row) : row;
                nb[q] = n;
            }
        };
        #pragma unroll
        for (int d = 0; d < RI - 1; ++d)
            if (d < my_total) issue_idx(d, ibuf[d]);
        for (int i0 = 0; i0 < my_total; i0 += RI) {
            #pragma unroll
            for (int u = 0; u < RI; ++u) {
                const int i = i0 + u;
                if (i < my_total) {
                    const int c = grp + i * kTcGroups;
                    const uint32_t stage = (uint32_t)c % kTcStages, ph = ((uint32_t)c / kTcStages) & 1u;
                    const int j = c % NCH;
                    const uint32_t sA_hi = smem_base + stage * STAGE_BYTES;
                    const uint32_t sA_lo = sA_hi + A_BYTES;
                    const uint32_t sB = sA_lo + A_BYTES;
                    const int idx = j * CE + piece * PE;
                    const int ci = idx % CI;
                    mbar_wait(empty0 + 8 * stage, ph ^ 1u);
                    #pragma unroll
                    for (int m = 0; m < 8; ++m) {
                        const int rl = (lane >> 3) + 4 * m;              // row inside this warp's 32 rows
                        int n = -1;
                        #pragma unroll
                        for (int q = 0; q < KPC; ++q) {
                            const int nq = __shfl_sync(0xFFFFFFFFu, ibuf[u][q], rl);
                            if (KPC == 1 || (piece * PE) / CI == q) n = nq;
                        }
                        const bool on = (n >= 0) && (idx < KT);
                        const size_t nn = on ? (size_t)n : 0;
                        size_t eoff;
                        bool second = false;
                        if (CAT) { second = ci >= CI / 2; eoff = nn * (CI / 2) + (second ? ci - CI / 2 : ci); }
                        else eoff = nn * CI + ci;
                        const uint32_t rt = (uint32_t)(tp & ~31) + (uint32_t)rl;
                        const uint32_t off = rt * 128u + (uint32_t)((piece ^ (int)(rt & 7u)) << 4);
                        const uint32_t sz = on ? 16u : 0u;
                        // element offsets are in operand elements (float or __half)
                        cp_async16(sA_hi + off, (const char*)(second ? a.inB : a.inA) + eoff * ES, sz);
                        if (PRECISE) cp_async16(sA_lo + off, (const char*)(second ? a.inB_lo : a.inA_lo) + eoff * ES, sz);
                    }
                    {   // weight chunk: hi tile then lo tile, already swizzled images (L2 resident)
                        const float* src = a.wimg + (size_t)j * (2 * B_BYTES / 4);
                        constexpr int NPIECE = (PRECISE ? 2 : 1) * B_BYTES / 16;
                        #pragma unroll
                        for (int w = 0; w < (NPIECE + kTcProducers - 1) / kTcProducers; ++w) {
                            int q = tp + w * kTcProducers;
                            if (q < NPIECE) cp_async16(sB + 16u * q, src + 4 * q, 16u);
                        }
                    }
                    cp_async_arrive(full0 + 8 * stage);   // one of the 128 arrivals, fires when this thread's copies landed
                    if (i + RI - 1 < my_total) issue_idx(i + RI - 1, ibuf[(u + RI - 1) % RI]);
                }
            }
        }
    } else if (warp < kTcProdWarps + 4) {
        // ===================== epilogue: TMEM -> registers -> residual / select -> global =====================
        const int te = t - kTcGroups * kTcProducers;      // tile row, also the TMEM lane
        // staging tiles of the bulk epilogue, behind the stage ring (row-major like global memory: one bulk copy per tile)
        float* const out_s = reinterpret_cast<float*>(smem + (size_t)kTcStages * STAGE_BYTES);
        float* const res_s = out_s + 128 * CO;
        __half* const sp_s = reinterpret_cast<__half*>(res_s + 128 * CO);      // [2 splits][hi, lo][128][CO]
        const uint32_t res_bar = smem_u32(&bars[2 * kTcMaxStages + 4]);
        uint32_t tile_it = 0, res_it = 0;
        for (int tseq = blockIdx.x; tseq < n_tiles; tseq += gridDim.x, ++tile_it) {
            const int tile = a.tile_list ? __ldg(a.tile_list + tseq) : tseq;
            const uint32_t as = tile_it & 1u, aph = (tile_it >> 1) & 1u;
            const int row = tile * 128 + te;
            const bool live = row < a.n_out;
            const uint32_t rows_valid = (uint32_t)min(128, a.n_out - tile * 128);
            if (BULK && a.res && te == 0) {
                // every epilogue thread passed the barrier in front of the previous tile's stores, i.e. finished reading
                // res_s: the residual tile of this tile can land while its MMAs are still running
                mbar_arrive_expect_tx(res_bar, rows_valid * CO * 4u);
                bulk_g2s(smem_u32(res_s), a.res + (size_t)tile * 128 * CO, rows_valid * CO * 4u, res_bar);
            }
            mbar_wait(tfull0 + 8 * as, aph);
            tc_fence_after();
            if (BULK) {
                if (te == 0) bulk_wait_read();              // the previous tile's bulk stores have read the staging tiles
                named_barrier(1, kTcEpilogue);
                if (a.res) { mbar_wait(res_bar, res_it & 1u); ++res_it; }
            }
            const uint32_t taddr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + as * ACC;
            const size_t o = (size_t)row * CO;
            const bool take_src = live && a.sel_flag && a.sel_flag[row] == 0;
            #pragma unroll
            for (int c0 = 0; c0 < NP; c0 += 16) {           // 16 output channels at a time
                float v[16];
                tmem_ld16(taddr + c0, v);
                if (PRECISE) {
                    float w[16];
                    tmem_ld16(taddr + NP + c0, w);
                    #pragma unroll
                    for (int c = 0; c < 16; ++c) v[c] += w[c];
                }
                if (c0 + 16 >= NP) {
                    tc_fence_before();
                    mbar_arrive(tempty0 + 8 * as);          // all columns read: this accumulator stage may be overwritten
                }
                if ((BULK || live) && c0 < CO) {
                    #pragma unroll
                    for (int c = 0; c < 16 && c0 + c < CO; c += 8) {      // CO may be 8 (NP padded to 16)
                        const int ch = c0 + c;
                        float r[8];
                        #pragma unroll
                        for (int i = 0; i < 8; ++i) r[i] = v[c + i];
                        if (a.res) {
                            const float* rp = BULK ? res_s + te * CO + ch : a.res + o + ch;
                            const float4 e0 = *reinterpret_cast<const float4*>(rp);
                            const float4 e1 = *reinterpret_cast<const float4*>(rp + 4);
                            r[0] += e0.x; r[1] += e0.y; r[2] += e0.z; r[3] += e0.w;
                            r[4] += e1.x; r[5] += e1.y; r[6] += e1.z; r[7] += e1.w;
                        }
                        if (take_src) {
                            const float4 e0 = *reinterpret_cast<const float4*>(a.sel_src + o + ch);
                            const float4 e1 = *reinterpret_cast<const float4*>(a.sel_src + o + ch + 4);
                            r[0] = e0.x; r[1] = e0.y; r[2] = e0.z; r[3] = e0.w;
                            r[4] = e1.x; r[5] = e1.y; r[6] = e1.z; r[7] = e1.w;
                        }
                        if (a.out) {
                            float* op = BULK ? out_s + te * CO + ch : a.out + o + ch;
                            *reinterpret_cast<float4*>(op) = make_float4(r[0], r[1], r[2], r[3]);
                            *reinterpret_cast<float4*>(op + 4) = make_float4(r[4], r[5], r[6], r[7]);
                        }
                        if (F16) {
                            #pragma unroll
                            for (int q = 0; q < 2; ++q) {    // the consumers' BN + ReLU + FP16 hi/lo split, fused: 16-byte stores
                                if (!a.sp[q].hi) continue;
                                float x[8];
                                float mx = 0.f;
                                #pragma unroll
                                for (int i = 0; i < 8; ++i) {
                                    x[i] = a.sp[q].alpha ? fmaxf(fmaf(r[i], a.sp[q].alpha[ch + i], a.sp[q].beta[ch + i]), 0.f) : r[i];
                                    mx = fmaxf(mx, fabsf(x[i]));
                                }
                                if (live && !(mx < 3.0e4f)) *a.range_flag = 1;
                                uint32_t hw[4], lw[4];
                                #pragma unroll
                                for (int i = 0; i < 4; ++i) {
                                    const __half2 h = __floats2half2_rn(x[2 * i], x[2 * i + 1]);
                                    const float2 f = __half22float2(h);
                                    const __half2 l = __floats2half2_rn(x[2 * i] - f.x, x[2 * i + 1] - f.y);
                                    hw[i] = *reinterpret_cast<const uint32_t*>(&h);
                                    lw[i] = *reinterpret_cast<const uint32_t*>(&l);
                                }
                                __half* hp = BULK ? sp_s + (2 * q) * 128 * CO + te * CO + ch : a.sp[q].hi + o + ch;
                                __half* lp = BULK ? sp_s + (2 * q + 1) * 128 * CO + te * CO + ch : a.sp[q].lo + o + ch;
                                *reinterpret_cast<uint4*>(hp) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
                                *reinterpret_cast<uint4*>(lp) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
                            }
                        }
                    }
                }
            }
            if (BULK) {
                fence_proxy_async();                        // staging tiles were written through the generic proxy
                named_barrier(2, kTcEpilogue);
                if (te == 0) {
                    const size_t t0 = (size_t)tile * 128 * CO;
                    if (a.out) bulk_s2g(a.out + t0, smem_u32(out_s), rows_valid * CO * 4u);
                    #pragma unroll
                    for (int q = 0; q < 2; ++q)
                        if (a.sp[q].hi) {
                            bulk_s2g(a.sp[q].hi + t0, smem_u32(sp_s + (2 * q) * 128 * CO), rows_valid * CO * 2u);
                            bulk_s2g(a.sp[q].lo + t0, smem_u32(sp_s + (2 * q + 1) * 128 * CO), rows_valid * CO * 2u);
                        }
                    bulk_commit();
                }
            }
        }
        if (BULK && te == 0) bulk_wait_all();               // stores complete before the shared memory goes away
    } else if ((t & 31) == 0) {
        // ===================== MMA issuer: one thread =====================
        uint32_t it = 0, tile_it = 0;
        for (int tseq = blockIdx.x; tseq < n_tiles; tseq += gridDim.x, ++tile_it) {
            const uint32_t as = tile_it & 1u, aph = (tile_it >> 1) & 1u;
            mbar_wait(tempty0 + 8 * as, aph ^ 1u);      // epilogue has drained this accumulator stage
            tc_fence_after();
            const uint32_t tmem_d = tmem_base + as * ACC;
            #pragma unroll 1
            for (int j = 0; j < NCH; ++j, ++it) {
                const uint32_t stage = it % kTcStages, ph = (it / kTcStages) & 1u;
                mbar_wait(full0 + 8 * stage, ph);
                fence_proxy_async();                    // cp.async wrote through the generic proxy; the MMA reads through the async proxy
                tc_fence_after();
                const uint32_t sA_hi = smem_u32(smem + (size_t)stage * STAGE_BYTES);
                const uint32_t sA_lo = sA_hi + A_BYTES, sB_hi = sA_lo + A_BYTES;        // W_lo rows follow W_hi at sB_hi + B_BYTES
                #pragma unroll
                for (int ks = 0; ks < 4; ++ks) {        // 4 K steps of 8 tf32 (32 B) inside the 128 B swizzle span
                    const uint64_t ah = make_desc_sw128(sA_hi + ks * 32), bh = make_desc_sw128(sB_hi + ks * 32);
                    if (PRECISE) {
                        const uint64_t al = make_desc_sw128(sA_lo + ks * 32);
                        if (F16) { umma_f16(tmem_d, ah, bh, IDESC2, (j | ks) ? 1u : 0u); umma_f16(tmem_d, al, bh, IDESC, 1u); }
                        else { umma_tf32(tmem_d, ah, bh, IDESC2, (j | ks) ? 1u : 0u); umma_tf32(tmem_d, al, bh, IDESC, 1u); }
                    } else {
                        if (F16) umma_f16(tmem_d, ah, bh, IDESC, (j | ks) ? 1u : 0u);
                        else umma_tf32(tmem_d, ah, bh, IDESC, (j | ks) ? 1u : 0u);
                    }
                }
                umma_commit(empty0 + 8 * stage);       // stage reusable once these MMAs have read it
            }
            umma_commit(tfull0 + 8 * as);               // accumulator complete
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == kTcProdWarps + 4) {
        __syncwarp();
        tmem_dealloc(tmem_base, kTmemCols);
    }
}

// Optional BatchNorm(eval) + ReLU, then TF32 hi / lo split, once per element (the SIMT kernel re-applies the
// activation on every gather instead).
__global__ void __launch_bounds__(256) k_act_split(const float4* __restrict__ in, const float* __restrict__ bn, int c, int bn_off,
                                                   int bn_stride, int64_t n4, float4* __restrict__ hi, float4* __restrict__ lo) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n4) return;
    float4 x = in[i];
    if (bn) {
        const int ch = (int)((i * 4) % c) + bn_off;
        x.x = fmaxf(fmaf(x.x, bn[ch + 0], bn[bn_stride + ch + 0]), 0.f);
        x.y = fmaxf(fmaf(x.y, bn[ch + 1], bn[bn_stride + ch + 1]), 0.f);
        x.z = fmaxf(fmaf(x.z, bn[ch + 2], bn[bn_stride + ch + 2]), 0.f);
        x.w = fmaxf(fmaf(x.w, bn[ch + 3], bn[bn_stride + ch + 3]), 0.f);
    }
    float4 h = make_float4(tf32_rn(x.x), tf32_rn(x.y), tf32_rn(x.z), tf32_rn(x.w));
    hi[i] = h;
    if (lo) lo[i] = make_float4(tf32_rn(x.x - h.x), tf32_rn(x.y - h.y), tf32_rn(x.z - h.z), tf32_rn(x.w - h.w));
}

// FP16 variant of the split: hi = fp16(x), lo = fp16(x - hi).  |x| must stay inside the FP16 range; a value above
// 3e4 raises the flag that the engine turns into an error (measured activations of the checkpoints: <= ~1e3).
__global__ void __launch_bounds__(256) k_act_split_f16(const float4* __restrict__ in, const float* __restrict__ bn, int c, int bn_off,
                                                       int bn_stride, int64_t n4, uint2* __restrict__ hi, uint2* __restrict__ lo,
                                                       int* __restrict__ range_flag) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n4) return;
    float4 x = in[i];
    if (bn) {
        const int ch = (int)((i * 4) % c) + bn_off;
        x.x = fmaxf(fmaf(x.x, bn[ch + 0], bn[bn_stride + ch + 0]), 0.f);
        x.y = fmaxf(fmaf(x.y, bn[ch + 1], bn[bn_stride + ch + 1]), 0.f);
        x.z = fmaxf(fmaf(x.z, bn[ch + 2], bn[bn_stride + ch + 2]), 0.f);
        x.w = fmaxf(fmaf(x.w, bn[ch + 3], bn[bn_stride + ch + 3]), 0.f);
    }
    if (!(fmaxf(fmaxf(fabsf(x.x), fabsf(x.y)), fmaxf(fabsf(x.z), fabsf(x.w))) < 3.0e4f)) *range_flag = 1;
    const __half2 h01 = __floats2half2_rn(x.x, x.y), h23 = __floats2half2_rn(x.z, x.w);
    const float2 f01 = __half22float2(h01), f23 = __half22float2(h23);
    hi[i] = make_uint2(*reinterpret_cast<const uint32_t*>(&h01), *reinterpret_cast<const uint32_t*>(&h23));
    if (lo) {
        const __half2 l01 = __floats2half2_rn(x.x - f01.x, x.y - f01.y), l23 = __floats2half2_rn(x.z - f23.x, x.w - f23.y);
        lo[i] = make_uint2(*reinterpret_cast<const uint32_t*>(&l01), *reinterpret_cast<const uint32_t*>(&l23));
    }
}

void act_split_f16(Ctx& cx, const float* in, const float* bn, int c, int bn_off, int bn_stride, int64_t n, void* hi, void* lo,
                   int* range_flag) {
    int64_t n4 = n * c / 4;
    if (n4 <= 0) return;
    k_act_split_f16<<<div_up(n4, 256), 256, 0, cx.st>>>((const float4*)in, bn, c, bn_off, bn_stride, n4, (uint2*)hi, (uint2*)lo,
                                                        range_flag);
    cx.launches++;
}

void act_split(Ctx& cx, const float* in, const float* bn, int c, int bn_off, int bn_stride, int64_t n, float* hi, float* lo) {
    int64_t n4 = n * c / 4;
    if (n4 <= 0) return;
    k_act_split<<<div_up(n4, 256), 256, 0, cx.st>>>((const float4*)in, bn, c, bn_off, bn_stride, n4, (float4*)hi, (float4*)lo);
    cx.launches++;
}

// ------------------------------------------------------------------------------------------ host side

// Weight image for the tensor-core kernel: per chunk of 32 K indices a [NP][32] K-major tile, SWIZZLE_128B,
// TF32 "hi" (low 13 mantissa bits cleared) followed by the "lo" remainder.
static float host_tf32_rn(float x) {                  // cvt.rna.tf32.f32 on the host (round to nearest, ties away)
    uint32_t b;
    std::memcpy(&b, &x, 4);
    b = (b + 0x1000u) & 0xFFFFE000u;
    float r;
    std::memcpy(&r, &b, 4);
    return r;
}

std::vector<float> make_tc_weight_image(const float* w_kcico, int kv, int ci, int co) {
    const int KT = kv * ci, NCH = (KT + 31) / 32, NP = co < 16 ? 16 : co;
    std::vector<float> img((size_t)NCH * 2 * NP * 32, 0.0f);
    for (int j = 0; j < NCH; ++j)
        for (int n = 0; n < NP; ++n)
            for (int kk = 0; kk < 32; ++kk) {
                int idx = j * 32 + kk;
                float val = 0.0f;
                if (idx < KT && n < co) val = w_kcico[(size_t)idx * co + n];        // [k][ci][co] flattened: idx = k*ci + c
                float hi = host_tf32_rn(val);
                float lo = host_tf32_rn(val - hi);
                int p = kk / 4, e = kk % 4;
                size_t off = (size_t)n * 32 + (size_t)((p ^ (n & 7)) * 4) + e;
                img[(size_t)j * 2 * NP * 32 + off] = hi;
                img[(size_t)j * 2 * NP * 32 + (size_t)NP * 32 + off] = lo;
            }
    return img;
}

// FP16 weight image: per chunk of 64 K indices a [NP][64] K-major tile of __half, SWIZZLE_128B, hi then lo.
// Returned as floats (two halves per float) so that it can live in the same weight blob.
std::vector<float> make_tc_weight_image_f16(const float* w_kcico, int kv, int ci, int co) {
    const int KT = kv * ci, NCH = (KT + 63) / 64, NP = co < 16 ? 16 : co;
    std::vector<__half> img((size_t)NCH * 2 * NP * 64, __float2half(0.0f));
    for (int j = 0; j < NCH; ++j)
        for (int n = 0; n < NP; ++n)
            for (int kk = 0; kk < 64; ++kk) {
                int idx = j * 64 + kk;
                float val = 0.0f;
                if (idx < KT && n < co) val = w_kcico[(size_t)idx * co + n];
                __half hi = __float2half_rn(val);
                __half lo = __float2half_rn(val - __half2float(hi));
                int p = kk / 8, e = kk % 8;
                size_t off = (size_t)n * 64 + (size_t)((p ^ (n & 7)) * 8) + e;
                img[(size_t)j * 2 * NP * 64 + off] = hi;
                img[(size_t)j * 2 * NP * 64 + (size_t)NP * 64 + off] = lo;
            }
    std::vector<float> out((img.size() + 1) / 2, 0.0f);
    std::memcpy(out.data(), img.data(), img.size() * sizeof(__half));
    return out;
}

template <int CI, int CO, int KV, bool CAT, bool UP>
static void launch_tc(Ctx& cx, const TcArgs& a0, bool precise, bool f16) {
#ifndef SQSM_GR_NARROW
#define SQSM_GR_NARROW 3
#endif
#ifndef SQSM_GR_WIDE
#define SQSM_GR_WIDE 3
#endif
    constexpr int GR = ((CAT ? CI / 2 : CI) == 16 && KV == 27) ? SQSM_GR_NARROW : SQSM_GR_WIDE;      // FP16 kernels only (the default path)
    TcArgs a = a0;
    if (a.n_out <= 0) return;
    a.n_tiles = div_up(a.n_out, 128);
    constexpr int NP = CO < 16 ? 16 : CO;
    constexpr int STAGE = 2 * 128 * 128 + 2 * NP * 128;
    int grid = std::min(a.n_tiles, kNumSMs);
    if (f16) {
        constexpr int staging = tc_staging_bytes(CO, tc_bulk(CO, true, true));
        constexpr size_t smem = (size_t)tc_stages(NP, staging) * STAGE + staging + 1024;
        // the ring must hold at least as many stages as there are producer groups: with fewer, a group gets two ring cycles
        // ahead of a stage's "empty" barrier and its parity wait is satisfied by the stale phase (a hang)
        static_assert(tc_stages(NP, staging) >= GR, "ring shallower than the number of producer groups");
        auto kern = k_conv_tc<CI, CO, KV, CAT, true, true, GR, UP>;
        SQSM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<grid, tc_threads(GR), smem, cx.st>>>(a);
    } else if constexpr (!UP) {
        constexpr size_t smem = (size_t)tc_stages(NP) * STAGE + 1024;
        if (precise) {
            auto kern = k_conv_tc<CI, CO, KV, CAT, true, false, 2, false>;
            SQSM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            kern<<<grid, tc_threads(2), smem, cx.st>>>(a);
        } else {
            auto kern = k_conv_tc<CI, CO, KV, CAT, false, false, 2, false>;
            SQSM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            kern<<<grid, tc_threads(2), smem, cx.st>>>(a);
        }
    } else {
        throw CudaError("conv_tc: the inverse convolution runs in the FP16 split mode only");
    }
    cx.launches++;
}

// Dispatch on the shapes that occur in SmartTreeXX. Returns false for shapes it does not cover.
// `fused` (nullptr or two entries): operands the epilogue writes for the consumers of this output (FP16 modes).
// `up_parent` / `up_koff` non-null: SparseInverseConv3d(k = 2) - kv must be 8, the rulebook entry (k, row) is
// parent[row] when koff[row] == k (every fine row has exactly one pair), nbr is ignored.
bool conv_tc(Ctx& cx, int ci, int co, int kv, bool cat, const float* inA, const float* inB, const float* inA_lo,
             const float* inB_lo, const int32_t* nbr, const float* wimg, const float* res, const float* sel_src,
             const uint8_t* sel_flag, float* out, int64_t n_out, bool precise, bool f16, const TcFused* fused,
             int* range_flag, const int32_t* up_parent, const int32_t* up_koff, const int32_t* tile_list,
             const int32_t* tile_count) {
    TcArgs a;
    for (int q = 0; q < 2; ++q) {
        a.sp[q].hi = fused ? (__half*)fused[q].hi : nullptr;
        a.sp[q].lo = fused ? (__half*)fused[q].lo : nullptr;
        a.sp[q].alpha = fused ? fused[q].alpha : nullptr;
        a.sp[q].beta = fused ? fused[q].beta : nullptr;
    }
    a.range_flag = range_flag;
    a.inA = inA; a.inB = inB; a.inA_lo = inA_lo; a.inB_lo = inB_lo; a.nbr = nbr; a.wimg = wimg; a.res = res; a.sel_src = sel_src; a.sel_flag = sel_flag;
    a.up_parent = up_parent; a.up_koff = up_koff;
    a.tile_list = tile_list; a.tile_count = tile_count;
    a.out = out; a.n_out = (int)n_out; a.n_tiles = 0;
    const bool up = up_parent != nullptr;
#define SQSM_TC_CASE(CI_, CO_, KV_, CAT_, UP_)                                    \
    if (ci == CI_ && co == CO_ && kv == KV_ && cat == CAT_ && up == UP_) {        \
        launch_tc<CI_, CO_, KV_, CAT_, UP_>(cx, a, precise, f16);                 \
        return true;                                                              \
    }
    SQSM_TC_CASE(8, 8, 27, false, false) SQSM_TC_CASE(16, 8, 27, true, false) SQSM_TC_CASE(16, 8, 1, true, false)
    SQSM_TC_CASE(16, 16, 27, false, false) SQSM_TC_CASE(32, 32, 27, false, false) SQSM_TC_CASE(64, 64, 27, false, false)
    SQSM_TC_CASE(32, 16, 27, true, false) SQSM_TC_CASE(64, 32, 27, true, false)
    SQSM_TC_CASE(32, 16, 1, true, false) SQSM_TC_CASE(64, 32, 1, true, false)
    SQSM_TC_CASE(8, 16, 8, false, false) SQSM_TC_CASE(16, 32, 8, false, false) SQSM_TC_CASE(32, 64, 8, false, false)
    SQSM_TC_CASE(16, 8, 8, false, true) SQSM_TC_CASE(32, 16, 8, false, true) SQSM_TC_CASE(64, 32, 8, false, true)
#undef SQSM_TC_CASE
    return false;
}

}  // namespace sqsm
