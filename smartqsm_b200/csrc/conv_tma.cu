// Tensor-core sparse convolution with a TMA row gather (tcgen05 + cp.async.bulk.tensor tile::gather4).  sm_100a only.
//
// Same implicit GEMM as conv_tc.cu - D[rows, C_out] = A[rows, KV*C_in] . W with A gathered through the
// output-stationary rulebook, 2-term FP16 split (hi*hi + lo*hi + hi*lo, fp32-class accuracy), accumulators in TMEM -
// but the gather no longer goes through the LSU.  Measured (profiles/r01_gather_microbench_notes.md): 16-byte LDGSTS
// copies top out at ~4.5 TB/s of L2 -> SM traffic (1.2 ns x SM per 64-byte row slot in the production kernel), while
// UTMALDG.2D.GATHER4 issued from >= 16 warps moves a row slot in 0.7-0.9 ns x SM (64 B) / 1.1-1.7 ns x SM (128 B)
// provided absent neighbours are NOT encoded as out-of-bounds rows (4.5 ns) but point at in-bounds rows of zeros.
//
// Data layout.  The activated input of a convolution is written once per element as "split rows":
//   S[row][0..C) = fp16(x)  ("hi"),  S[row][C..2C) = fp16(x - hi)  ("lo"),  row pitch 4*C bytes,
// followed by kZeroRows rows of zeros; rulebook entry -1 (and rows past the end of the last tile) are redirected to one
// of those, chosen by a hash so that the reads spread over the L2 slices.  One gather4 moves four rows of <= 128 bytes
// into a K-major shared-memory tile with the swizzle of the tensor map (SWIZZLE_64B for C = 16, SWIZZLE_128B above),
// which is exactly the canonical layout of the tcgen05 shared-memory descriptors, so "hi" and "lo" operands are just
// two start addresses inside the same tile.  A concatenated input (decoder) is two tensor maps.
//
// Roles per CTA (persistent, one per SM, tiles of 128 output rows): up to 24 producer warps - every lane issues one gather4
// (4 rows) per unit, units are dealt round-robin to the warps; the weights of a stage arrive by one cp.async.bulk;
// 1 MMA thread; 4 epilogue warps (TMEM -> registers -> residual / select -> global), two accumulator stages.
#include "engine.h"
#include "tc_ptx.cuh"
#include <cuda_fp16.h>
#include <cstring>

namespace sqsm {

constexpr int kTmProdWarps = 24;
constexpr int kTmThreads = kTmProdWarps * 32 + 128 + 32;
constexpr int kTmMaxStages = 8;

struct alignas(64) TmArgs {
    CUtensorMap map[2];       // split rows of source 0 / 1
    const int32_t* nbr;       // [KV][n_out] or nullptr (identity)
    const uint8_t* wimg;      // per stage: B tiles in the order (g, source, part), each [NP][UB] swizzled
    const float* res;
    const float* sel_src;
    const uint8_t* sel_flag;
    float* out;
    int n_out, n_tiles, n_in;
};

template <int CS, int NS, int CO, int KV, int G>
struct TmShape {
    static constexpr int UB = CS == 16 ? 64 : 128;             // bytes per gathered box row
    static constexpr int PARTS = CS == 64 ? 2 : 1;             // boxes per split row
    static constexpr int NP = CO < 16 ? 16 : CO;
    static constexpr int UNIT = 128 * UB;                      // bytes of one gathered unit (128 rows)
    static constexpr int BT = 2 * NP * 2 * CS;                 // weight tile of one (offset, source): [W_hi rows ; W_lo rows][CS halves]
    static constexpr int UPS = G * NS * PARTS;                 // gather units per stage
    static constexpr int STAGE_A = UPS * UNIT, STAGE_B = G * NS * BT, STAGE = STAGE_A + STAGE_B;
    static constexpr int ACC = 2 * NP;                         // TMEM columns of one accumulator stage: [hi*hi + lo*hi | hi*lo]
    static constexpr uint32_t TMEM_COLS = 2 * ACC < 32 ? 32 : 2 * ACC;
    static constexpr int NST = KV / G;                         // stages per tile
    static constexpr int S = ((227 * 1024 - 2048) / STAGE) > kTmMaxStages ? kTmMaxStages : ((227 * 1024 - 2048) / STAGE);
    // issuing warps: consecutive items of a warp must be at most S stages apart, otherwise its parity wait on the "empty"
    // barrier could be satisfied by a phase two ring cycles old
    // ... and coprime to the items per stage, so that the (cheap) weight-copy item rotates over the warps instead of
    // parking a fixed subset of them
    static constexpr int gcd_(int x, int y) { return y == 0 ? x : gcd_(y, x % y); }
    static constexpr int pick_(int w) { return gcd_(w, UPS + 1) == 1 ? w : pick_(w - 1); }
    static constexpr int PW = pick_((S * (UPS + 1)) < kTmProdWarps ? (S * (UPS + 1)) : kTmProdWarps);
    static_assert(KV % G == 0 && S >= 2 && BT % 1024 == 0 && (CS == 16 || CS == 32 || CS == 64), "shape");
};

template <int CS, int NS, int CO, int KV, int G>
__global__ void __launch_bounds__(kTmThreads, 1) k_conv_tma(const __grid_constant__ TmArgs a) {
    using T = TmShape<CS, NS, CO, KV, G>;
    constexpr int UB = T::UB, PARTS = T::PARTS, NP = T::NP, UPS = T::UPS, NST = T::NST, S = T::S;
    // F16 x F16 -> F32, M 128; N = 2*NP for A_hi . [W_hi | W_lo], N = NP for A_lo . W_hi
    constexpr uint32_t IDESC_2N = (1u << 4) | ((uint32_t)((2 * NP) >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    constexpr uint32_t IDESC_N = (1u << 4) | ((uint32_t)(NP >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    __shared__ __align__(8) uint64_t bars[2 * kTmMaxStages + 4];
    __shared__ uint32_t tmem_slot;

    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
    const uint32_t full0 = smem_u32(&bars[0]), empty0 = smem_u32(&bars[kTmMaxStages]);
    const uint32_t tfull0 = smem_u32(&bars[2 * kTmMaxStages]), tempty0 = smem_u32(&bars[2 * kTmMaxStages + 2]);
    const uint32_t smem_base = smem_u32(smem);

    if (t == 0) {
        for (int s = 0; s < S; ++s) { mbar_init(full0 + 8 * s, UPS + 1); mbar_init(empty0 + 8 * s, 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(tfull0 + 8 * s, 1); mbar_init(tempty0 + 8 * s, 128); }
        fence_barrier_init();
    }
    if (warp == kTmProdWarps + 4) tmem_alloc(smem_u32(&tmem_slot), T::TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_slot;
    const int n_my_tiles = (a.n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;

    if (warp < T::PW) {
        // ===================== producers: rulebook rows -> TMA gather4 into the swizzled operand tiles =====================
        const int total = n_my_tiles * NST * (UPS + 1);            // unit sequence of this CTA (+1 per stage: the weights)
        // raw rulebook entries of an item (-1 = absent / past the end); the loads stay in flight until the item is issued
        auto load_rows = [&](int q, int* r) {
            const int sidx = q / (UPS + 1), u = q % (UPS + 1);
            r[0] = r[1] = r[2] = r[3] = -1;
            if (u == UPS) return;
            const int tile = (int)blockIdx.x + (sidx / NST) * (int)gridDim.x;
            const int k = (sidx % NST) * G + u / (NS * PARTS);
            const int row0 = tile * 128 + lane * 4;
            if (a.nbr) {
                const int32_t* src = a.nbr + (size_t)k * a.n_out + row0;
                #pragma unroll
                for (int i = 0; i < 4; ++i)
                    if (row0 + i < a.n_out) r[i] = __ldg(src + i);
            } else {
                #pragma unroll
                for (int i = 0; i < 4; ++i)
                    if (row0 + i < a.n_out) r[i] = row0 + i;
            }
        };
        constexpr int PF = 2;                                   // items of rulebook rows in flight per warp
        int rows[PF + 1][4];
        #pragma unroll
        for (int d = 0; d < PF; ++d)
            if (warp + d * T::PW < total) load_rows(warp + d * T::PW, rows[d]);
        for (int q0 = warp; q0 < total; q0 += (PF + 1) * T::PW) {
            #pragma unroll
            for (int d = 0; d <= PF; ++d) {                     // static register rotation
                const int q = q0 + d * T::PW;
                if (q >= total) break;
                if (q + PF * T::PW < total) load_rows(q + PF * T::PW, rows[(d + PF) % (PF + 1)]);
                const int sidx = q / (UPS + 1), u = q % (UPS + 1);
                const uint32_t stage = (uint32_t)sidx % S, ph = ((uint32_t)sidx / S) & 1u;
                const uint32_t sA = smem_base + stage * T::STAGE, sB = sA + T::STAGE_A;
                mbar_wait(empty0 + 8 * stage, ph ^ 1u);
                if (u == UPS) {
                    if (lane == 0) {
                        mbar_arrive_expect_tx(full0 + 8 * stage, T::STAGE_B);
                        bulk_g2s(sB, a.wimg + (size_t)(sidx % NST) * T::STAGE_B, T::STAGE_B, full0 + 8 * stage);
                    }
                } else {
                    const int s = (u / PARTS) % NS, p = u % PARTS;
                    const int k = (sidx % NST) * G + u / (NS * PARTS);
                    const int zbase = a.n_in + ((lane * 29 + k * 131 + sidx * 7) & (kTmZeroRows - 4));
                    int r[4];
                    #pragma unroll
                    for (int i = 0; i < 4; ++i) r[i] = rows[d][i] >= 0 ? rows[d][i] : zbase + i;      // absent -> a row of zeros
                    if (lane == 0) mbar_arrive_expect_tx(full0 + 8 * stage, T::UNIT);
                    __syncwarp();
                    tma_gather4(&a.map[s], full0 + 8 * stage, sA + (uint32_t)u * T::UNIT + (uint32_t)lane * 4u * UB, p * 64, r[0], r[1],
                                r[2], r[3]);
                }
            }
        }
    } else if (warp >= kTmProdWarps && warp < kTmProdWarps + 4) {
        // ===================== epilogue: TMEM -> registers -> residual / select -> global =====================
        const int te = t - kTmProdWarps * 32;              // tile row, also the TMEM lane
        uint32_t tile_it = 0;
        for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x, ++tile_it) {
            const uint32_t as = tile_it & 1u, aph = (tile_it >> 1) & 1u;
            const int row = tile * 128 + te;
            const bool live = row < a.n_out;
            mbar_wait(tfull0 + 8 * as, aph);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + as * (uint32_t)T::ACC;
            const size_t o = (size_t)row * CO;
            const bool take_src = live && a.sel_flag && a.sel_flag[row] == 0;
            #pragma unroll
            for (int c0 = 0; c0 < NP; c0 += 16) {               // 16 output channels at a time: two TMEM halves, summed
                float v[16], w[16];
                tmem_ld16(taddr + c0, v);
                tmem_ld16(taddr + NP + c0, w);
                if (c0 + 16 >= NP) {
                    tc_fence_before();
                    mbar_arrive(tempty0 + 8 * as);              // all columns read: the accumulator stage may be overwritten
                }
                if (live && c0 < CO) {
                    #pragma unroll
                    for (int c = 0; c < 16 && c0 + c < CO; c += 4) {      // CO may be 8 (NP padded to 16)
                        float4 r = make_float4(v[c] + w[c], v[c + 1] + w[c + 1], v[c + 2] + w[c + 2], v[c + 3] + w[c + 3]);
                        if (a.res) {
                            float4 e = *reinterpret_cast<const float4*>(a.res + o + c0 + c);
                            r.x += e.x; r.y += e.y; r.z += e.z; r.w += e.w;
                        }
                        if (take_src) r = *reinterpret_cast<const float4*>(a.sel_src + o + c0 + c);
                        *reinterpret_cast<float4*>(a.out + o + c0 + c) = r;
                    }
                }
            }
        }
    } else if (warp == kTmProdWarps + 4 && lane == 0) {
        // ===================== MMA issuer: one thread =====================
        uint32_t it = 0, tile_it = 0;
        for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x, ++tile_it) {
            const uint32_t as = tile_it & 1u, aph = (tile_it >> 1) & 1u;
            mbar_wait(tempty0 + 8 * as, aph ^ 1u);
            tc_fence_after();
            const uint32_t tmem_d = tmem_base + as * (uint32_t)T::ACC;
            #pragma unroll 1
            for (int j = 0; j < NST; ++j, ++it) {
                const uint32_t stage = it % S, ph = (it / S) & 1u;
                mbar_wait(full0 + 8 * stage, ph);               // TMA wrote through the async proxy: no proxy fence needed
                tc_fence_after();
                const uint32_t sA = smem_base + stage * T::STAGE, sB = sA + T::STAGE_A;
                #pragma unroll
                for (int gs = 0; gs < G * NS; ++gs) {
                    const uint32_t ua = sA + (uint32_t)(gs * PARTS) * T::UNIT, ub = sB + (uint32_t)gs * T::BT;
                    #pragma unroll
                    for (int ks = 0; ks < CS / 16; ++ks) {
                        uint64_t ah, al, b;
                        if (CS == 16) {
                            ah = make_desc_sw64(ua); al = make_desc_sw64(ua + 32);
                            b = make_desc_sw32(ub);
                        } else if (CS == 32) {
                            ah = make_desc_sw128(ua + ks * 32); al = make_desc_sw128(ua + 64 + ks * 32);
                            b = make_desc_sw64(ub + ks * 32);
                        } else {
                            ah = make_desc_sw128(ua + ks * 32); al = make_desc_sw128(ua + T::UNIT + ks * 32);
                            b = make_desc_sw128(ub + ks * 32);
                        }
                        umma_f16(tmem_d, ah, b, IDESC_2N, (j | gs | ks) ? 1u : 0u);     // [hi*hi | hi*lo]
                        umma_f16(tmem_d, al, b, IDESC_N, 1u);                            // lo*hi into the first half
                    }
                }
                umma_commit(empty0 + 8 * stage);
            }
            umma_commit(tfull0 + 8 * as);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == kTmProdWarps + 4) {
        __syncwarp();
        tmem_dealloc(tmem_base, T::TMEM_COLS);
    }
}

// BatchNorm(eval) + ReLU, FP16 hi / lo split, written as split rows [hi C | lo C]
__global__ void __launch_bounds__(256) k_act_split_rows(const float4* __restrict__ in, const float* __restrict__ bn, int c, int bn_off,
                                                        int bn_stride, int64_t n4, uint2* __restrict__ out, int* __restrict__ range_flag) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n4) return;
    float4 x = in[i];
    const int c4 = c >> 2;
    const int64_t row = i / c4;
    const int q = (int)(i - row * c4);                      // group of 4 channels inside the row
    if (bn) {
        const int ch = q * 4 + bn_off;
        x.x = fmaxf(fmaf(x.x, bn[ch + 0], bn[bn_stride + ch + 0]), 0.f);
        x.y = fmaxf(fmaf(x.y, bn[ch + 1], bn[bn_stride + ch + 1]), 0.f);
        x.z = fmaxf(fmaf(x.z, bn[ch + 2], bn[bn_stride + ch + 2]), 0.f);
        x.w = fmaxf(fmaf(x.w, bn[ch + 3], bn[bn_stride + ch + 3]), 0.f);
    }
    if (!(fmaxf(fmaxf(fabsf(x.x), fabsf(x.y)), fmaxf(fabsf(x.z), fabsf(x.w))) < 3.0e4f)) *range_flag = 1;
    const __half2 h01 = __floats2half2_rn(x.x, x.y), h23 = __floats2half2_rn(x.z, x.w);
    const float2 f01 = __half22float2(h01), f23 = __half22float2(h23);
    const __half2 l01 = __floats2half2_rn(x.x - f01.x, x.y - f01.y), l23 = __floats2half2_rn(x.z - f23.x, x.w - f23.y);
    uint2* o = out + row * (2 * c4) + q;                    // row pitch = 2C halves = 2*c4 uint2
    o[0] = make_uint2(*reinterpret_cast<const uint32_t*>(&h01), *reinterpret_cast<const uint32_t*>(&h23));
    o[c4] = make_uint2(*reinterpret_cast<const uint32_t*>(&l01), *reinterpret_cast<const uint32_t*>(&l23));
}

// split rows of `in` ([n][c] float32) into `out` ([n + kTmZeroRows][2c] half); the tail rows are zeros
void act_split_rows(Ctx& cx, const float* in, const float* bn, int c, int bn_off, int bn_stride, int64_t n, void* out, int* range_flag) {
    const int64_t n4 = n * c / 4;
    SQSM_CUDA(cudaMemsetAsync((char*)out + (size_t)n * c * 4, 0, (size_t)kTmZeroRows * c * 4, cx.st));
    if (n4 <= 0) return;
    k_act_split_rows<<<div_up(n4, 256), 256, 0, cx.st>>>((const float4*)in, bn, c, bn_off, bn_stride, n4, (uint2*)out, range_flag);
    cx.launches++;
}

// ------------------------------------------------------------------------------------------ host side

// Weight image: per stage j (G kernel offsets), per (g, source s) one K-major tile [2*NP rows][CS halves]: rows 0..NP-1
// hold W_hi[n][.], rows NP..2NP-1 hold W_lo[n][.]; row pitch 2*CS bytes, swizzled for that pitch (SWIZZLE_32B / 64B /
// 128B).  w_kcico is [k][ci][co] with ci = s * CS + channel.
std::vector<float> make_tma_weight_image(const float* w_kcico, int kv, int ci, int co, int cs, int g) {
    const int ns = ci / cs, np = co < 16 ? 16 : co, pitch = 2 * cs, bt = 2 * np * pitch, nst = kv / g;
    std::vector<uint8_t> img((size_t)nst * g * ns * bt, 0);
    for (int k = 0; k < kv; ++k)
        for (int s = 0; s < ns; ++s)
            for (int n = 0; n < co; ++n)
                for (int ch = 0; ch < cs; ++ch) {
                    const float val = w_kcico[((size_t)k * ci + (size_t)s * cs + ch) * co + n];
                    const __half hi = __float2half_rn(val);
                    const __half lo = __float2half_rn(val - __half2float(hi));
                    const int j = k / g, gg = k % g;
                    for (int which = 0; which < 2; ++which) {                // 0 = hi rows, 1 = lo rows
                        const int row = which * np + n, chunk = ch / 8, e = ch % 8;
                        const int swz = pitch == 128 ? (row & 7) : pitch == 64 ? ((row >> 1) & 3) : ((row >> 2) & 1);
                        const size_t off = ((size_t)j * g * ns + (size_t)(gg * ns + s)) * bt + (size_t)row * pitch +
                                           (size_t)((chunk ^ swz) * 16) + (size_t)e * 2;
                        const __half v = which ? lo : hi;
                        std::memcpy(&img[off], &v, 2);
                    }
                }
    std::vector<float> out((img.size() + 3) / 4, 0.0f);
    std::memcpy(out.data(), img.data(), img.size());
    return out;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_tiled() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        cudaDriverEntryPointQueryResult q;
        void* p = nullptr;
        SQSM_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
        SQSM_CHECK(p != nullptr, "cuTensorMapEncodeTiled is not available in this driver");
        fn = (EncodeTiledFn)p;
    }
    return fn;
}

// tensor map over split rows: [rows][2*cs] half, box = one row of min(2*cs, 64) halves
static void make_rows_map(CUtensorMap* tm, const void* base, int64_t rows, int cs) {
    const cuuint64_t dims[2] = {(cuuint64_t)(2 * cs), (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)(4 * cs)};
    const cuuint32_t box[2] = {(cuuint32_t)(cs == 16 ? 32 : 64), 1};
    const cuuint32_t es[2] = {1, 1};
    CUresult r = encode_tiled()(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), dims, strides, box, es,
                                CU_TENSOR_MAP_INTERLEAVE_NONE, cs == 16 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B,
                                CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    SQSM_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed");
}

template <int CS, int NS, int CO, int KV, int G>
static void launch_tma(Ctx& cx, TmArgs& a) {
    using T = TmShape<CS, NS, CO, KV, G>;
    constexpr size_t smem = (size_t)T::S * T::STAGE + 1024;
    auto kern = k_conv_tma<CS, NS, CO, KV, G>;
    SQSM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<std::min(a.n_tiles, kNumSMs), kTmThreads, smem, cx.st>>>(a);
    cx.launches++;
}

int tma_group(int cs, int kv) { return cs == 16 ? (kv == 27 ? 3 : kv == 8 ? 2 : 1) : 1; }

bool conv_tma_supported(int ci, int co, int kv, bool cat) {
    const int cs = cat ? ci / 2 : ci;
    if (!(cs == 16 || cs == 32 || cs == 64) || co < 16) return false;
    return kv == 27 || kv == 8 || kv == 1;
}

// splitA / splitB: split rows ([n_in + kTmZeroRows][2*cs] half) of the one or two sources
bool conv_tma(Ctx& cx, int ci, int co, int kv, bool cat, const void* splitA, const void* splitB, const int32_t* nbr, const float* wimg,
              const float* res, const float* sel_src, const uint8_t* sel_flag, float* out, int64_t n_out, int64_t n_in) {
    if (n_out <= 0) return true;
    const int cs = cat ? ci / 2 : ci;
    TmArgs a;
    memset(&a, 0, sizeof(a));
    make_rows_map(&a.map[0], splitA, n_in + kTmZeroRows, cs);
    if (cat) make_rows_map(&a.map[1], splitB, n_in + kTmZeroRows, cs);
    a.nbr = nbr; a.wimg = (const uint8_t*)wimg; a.res = res; a.sel_src = sel_src; a.sel_flag = sel_flag; a.out = out;
    a.n_out = (int)n_out; a.n_tiles = div_up(n_out, 128); a.n_in = (int)n_in;
#define SQSM_TM_CASE(CI_, CO_, KV_, CAT_, G_)                                                    \
    if (ci == CI_ && co == CO_ && kv == KV_ && cat == CAT_) {                                   \
        launch_tma<(CAT_ ? CI_ / 2 : CI_), (CAT_ ? 2 : 1), CO_, KV_, G_>(cx, a);                 \
        return true;                                                                            \
    }
    SQSM_TM_CASE(16, 16, 27, false, 3) SQSM_TM_CASE(32, 32, 27, false, 1) SQSM_TM_CASE(64, 64, 27, false, 1)
    SQSM_TM_CASE(32, 16, 27, true, 3) SQSM_TM_CASE(64, 32, 27, true, 1)
    SQSM_TM_CASE(32, 16, 1, true, 1) SQSM_TM_CASE(64, 32, 1, true, 1)
    SQSM_TM_CASE(16, 32, 8, false, 2) SQSM_TM_CASE(32, 64, 8, false, 1)
#undef SQSM_TM_CASE
    return false;
}

}  // namespace sqsm
