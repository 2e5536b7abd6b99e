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
    if (BN) {
        for (int i = t; i < 2 * CI; i += 256) sbn[i] = a.bn[i];
    }
    float acc[RPT][CPT];
    #pragma unroll
    for (int j = 0; j < RPT; ++j)
        #pragma unroll
        for (int c = 0; c < CPT; ++c) acc[j][c] = 0.0f;

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
                        #pragma unroll
                        for (int j = 0; j < RPT; ++j) {
                            const float s = comp4(v[j], cc);
                            acc[j][4 * q + 0] = fmaf(s, wv.x, acc[j][4 * q + 0]);
                            acc[j][4 * q + 1] = fmaf(s, wv.y, acc[j][4 * q + 1]);
                            acc[j][4 * q + 2] = fmaf(s, wv.z, acc[j][4 * q + 2]);
                            acc[j][4 * q + 3] = fmaf(s, wv.w, acc[j][4 * q + 3]);
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
            float4 r = make_float4(acc[j][4 * q], acc[j][4 * q + 1], acc[j][4 * q + 2], acc[j][4 * q + 3]);
            if (a.res) {
                float4 e = *reinterpret_cast<const float4*>(a.res + o + 4 * q);
                r.x += e.x; r.y += e.y; r.z += e.z; r.w += e.w;
            }
            if (take_src) r = *reinterpret_cast<const float4*>(a.sel_src + o + 4 * q);
            *reinterpret_cast<float4*>(a.out + o + 4 * q) = r;
        }
    }
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
    struct Pending { std::string name; size_t w_off, bn_off; bool has_bn; int ci, co, kv; };
    std::vector<Pending> pend;
    auto add_conv = [&](const std::string& name, const std::string& wkey, const std::string& bnkey, int ci, int co, int kv,
                        int ci_pad, bool flip) {
        const auto& W = need(st, wkey, (size_t)co * kv * ci);   // [co][k][ci]
        Pending p{name, blob.size(), 0, !bnkey.empty(), ci_pad, co, kv};
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
            add_conv(L + ".tail.i", prefix + ".blocks_tail.block0.i_branch.0.weight", "", 2 * c, c, 1, 2 * c, false);
            add_conv(L + ".tail.a", prefix + ".blocks_tail.block0.conv_branch.2.weight", prefix + ".blocks_tail.block0.conv_branch.0", 2 * c, c, 27, 2 * c, false);
            add_conv(L + ".tail.b", prefix + ".blocks_tail.block0.conv_branch.5.weight", prefix + ".blocks_tail.block0.conv_branch.3", c, c, 27, c, false);
        }
        prefix += ".u";
    }
    nw.blob.alloc((int64_t)blob.size(), cx.st);
    SQSM_CUDA(cudaMemcpyAsync(nw.blob.p, blob.data(), blob.size() * sizeof(float), cudaMemcpyHostToDevice, cx.st));
    SQSM_CUDA(cudaStreamSynchronize(cx.st));
    nw.conv.clear();
    for (const auto& p : pend) {
        ConvW c;
        c.ci = p.ci; c.co = p.co; c.kv = p.kv;
        c.w = nw.blob.p + p.w_off;
        c.bn = p.has_bn ? nw.blob.p + p.bn_off : nullptr;
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

struct ProfTimer {
    cudaEvent_t a = nullptr, b = nullptr;
    std::map<std::string, ConvProfile>* prof;
    std::string fam;
    cudaStream_t st;
    double bytes, flops;
    int64_t l0;
    Ctx& cx;
    ProfTimer(Ctx& cx_, std::map<std::string, ConvProfile>* p, const std::string& f, double by, double fl)
        : prof(p), fam(f), st(cx_.st), bytes(by), flops(fl), l0(cx_.launches), cx(cx_) {
        if (prof) { cudaEventCreate(&a); cudaEventCreate(&b); cudaEventRecord(a, st); }
    }
    ~ProfTimer() {
        if (!prof) return;
        cudaEventRecord(b, st);
        cudaEventSynchronize(b);
        float ms = 0;
        cudaEventElapsedTime(&ms, a, b);
        ConvProfile& c = (*prof)[fam];
        c.ms += ms; c.launches += cx.launches - l0; c.bytes += bytes; c.flops += flops;
        cudaEventDestroy(a); cudaEventDestroy(b);
    }
};

template <int CI, int CO, int KV, bool BN, bool CAT>
static void conv_dispatch(Ctx& cx, const ConvArgs& a) {
    if constexpr (CO == 8)       launch_conv<CI, CO, KV, 2, 8, BN, CAT>(cx, a);
    else if constexpr (CO == 16) launch_conv<CI, CO, KV, 2, 16, BN, CAT>(cx, a);
    else                         launch_conv<CI, CO, KV, 4, 16, BN, CAT>(cx, a);
}

static ConvArgs mk(const ConvW& w, const float* inA, const float* inB, const int32_t* nbr, float* out, int64_t n_out,
                   const float* res = nullptr, const float* sel_src = nullptr, const uint8_t* sel_flag = nullptr) {
    ConvArgs a;
    a.inA = inA; a.inB = inB; a.nbr = nbr; a.w = w.w; a.bn = w.bn; a.res = res; a.sel_src = sel_src;
    a.sel_flag = sel_flag; a.out = out; a.n_out = (int)n_out; a.kc = 0;
    return a;
}

template <int C>
static void run_block(Ctx& cx, const NetWeights& nw, const std::string& L, const LevelTables& lv, const float* x,
                      float* tmp, float* out, std::map<std::string, ConvProfile>* prof, int level) {
    // ResidualBlock with identity i_branch (sparse_module.py:32-39): out = convB(bnrelu(convA(bnrelu(x)))) + x
    double by = 2.0 * (4.0 * lv.n * C * 2 + 4.0 * 27 * lv.n) + 4.0 * lv.n * C;
    ProfTimer pt(cx, prof, "subm_l" + std::to_string(level), by, 2.0 * 2.0 * (double)lv.pairs * C * C);
    conv_dispatch<C, C, 27, true, false>(cx, mk(nw.conv.at(L + ".blk.a"), x, nullptr, lv.nbr.p, tmp, lv.n));
    conv_dispatch<C, C, 27, true, false>(cx, mk(nw.conv.at(L + ".blk.b"), tmp, nullptr, lv.nbr.p, out, lv.n, x));
}

template <int C>
static void run_tail(Ctx& cx, const NetWeights& nw, const std::string& L, const LevelTables& lv, const float* enc,
                     const float* up, float* res, float* tmp, float* out, std::map<std::string, ConvProfile>* prof,
                     int level) {
    // blocks_tail on cat(enc, up) (sparse_module.py:114-116); rows of batches that did not descend keep `enc`
    double by = 4.0 * lv.n * (2 * C + C) * 2 + 4.0 * lv.n * (C + C) + 2.0 * 4.0 * 27 * lv.n + 4.0 * lv.n * C;
    ProfTimer pt(cx, prof, "subm_l" + std::to_string(level), by,
                 2.0 * (double)lv.pairs * (2 * C * C + C * C) + 2.0 * (double)lv.n * 2 * C * C);
    conv_dispatch<2 * C, C, 1, false, true>(cx, mk(nw.conv.at(L + ".tail.i"), enc, up, nullptr, res, lv.n));
    conv_dispatch<2 * C, C, 27, true, true>(cx, mk(nw.conv.at(L + ".tail.a"), enc, up, lv.nbr.p, tmp, lv.n));
    conv_dispatch<C, C, 27, true, false>(cx, mk(nw.conv.at(L + ".tail.b"), tmp, nullptr, lv.nbr.p, out, lv.n, res, enc,
                                                lv.row_desc.p));
}

static void keep_copy(Ctx& cx, ForwardBuffers& fb, const std::string& name, const float* src, int64_t n, int ch,
                      int level) {
    DBuf<float> b(n * ch, cx.st);
    SQSM_CUDA(cudaMemcpyAsync(b.p, src, sizeof(float) * n * ch, cudaMemcpyDeviceToDevice, cx.st));
    fb.keep[name] = std::move(b);
    fb.keep_meta[name] = {level, ch};
}

void net_forward(Ctx& cx, const NetWeights& nw, LevelTables* lv, const Level0Rows& rows, int n_levels,
                 ForwardBuffers& fb, bool keep, float* d_newP, float* d_newD,
                 std::map<std::string, ConvProfile>* prof) {
    cudaStream_t st = cx.st;
    SQSM_CHECK(nw.ready, "weights not finalised");
    DBuf<float> enc[kLevels], tmp[kLevels];
    // ---- encoder
    const int64_t n0 = lv[0].n;
    DBuf<float> x0(n0 * 8, st);
    {
        ProfTimer pt(cx, prof, "subm_l0", 4.0 * n0 * (3 + 8) + 4.0 * 27 * n0, 2.0 * (double)lv[0].pairs * 3 * 8);
        conv_dispatch<4, 8, 27, false, false>(cx, mk(nw.conv.at("input"), (const float*)rows.xyz.p, nullptr, lv[0].nbr.p,
                                                     x0.p, n0));
    }
    if (keep) keep_copy(cx, fb, "input_conv", x0.p, n0, 8, 0);
    DBuf<float> cur_in = std::move(x0);
    for (int l = 0; l < n_levels; ++l) {
        const int64_t n = lv[l].n;
        const int c = kPlanes[l];
        std::string L = "L" + std::to_string(l);
        enc[l].alloc(n * c, st);
        tmp[l].alloc(n * c, st);
        switch (l) {
            case 0: run_block<8>(cx, nw, L, lv[l], cur_in.p, tmp[l].p, enc[l].p, prof, l); break;
            case 1: run_block<16>(cx, nw, L, lv[l], cur_in.p, tmp[l].p, enc[l].p, prof, l); break;
            case 2: run_block<32>(cx, nw, L, lv[l], cur_in.p, tmp[l].p, enc[l].p, prof, l); break;
            default: run_block<64>(cx, nw, L, lv[l], cur_in.p, tmp[l].p, enc[l].p, prof, l); break;
        }
        if (keep) keep_copy(cx, fb, "enc" + std::to_string(l), enc[l].p, n, c, l);
        if (l + 1 < n_levels) {
            const int64_t nn = lv[l + 1].n;
            DBuf<float> nxt(nn * kPlanes[l + 1], st);
            ProfTimer pt(cx, prof, "down_up", 4.0 * (n * c + nn * kPlanes[l + 1]) + 4.0 * n,
                         2.0 * (double)n * c * kPlanes[l + 1]);
            ConvArgs a = mk(nw.conv.at(L + ".down"), enc[l].p, nullptr, lv[l].child.p, nxt.p, nn);
            switch (l) {
                case 0: conv_dispatch<8, 16, 8, true, false>(cx, a); break;
                case 1: conv_dispatch<16, 32, 8, true, false>(cx, a); break;
                default: conv_dispatch<32, 64, 8, true, false>(cx, a); break;
            }
            cur_in = std::move(nxt);
        }
    }
    // ---- decoder
    DBuf<float> cur = std::move(enc[n_levels - 1]);
    for (int l = n_levels - 2; l >= 0; --l) {
        const int64_t n = lv[l].n;
        const int c = kPlanes[l];
        std::string L = "L" + std::to_string(l);
        DBuf<float> up(n * c, st), res(n * c, st), out(n * c, st);
        {
            ProfTimer pt(cx, prof, "down_up", 4.0 * (lv[l + 1].n * kPlanes[l + 1] + n * c) + 4.0 * n,
                         2.0 * (double)n * c * kPlanes[l + 1]);
            ConvArgs a = mk(nw.conv.at(L + ".up"), cur.p, nullptr, lv[l].inv.p, up.p, n);
            switch (l) {
                case 0: conv_dispatch<16, 8, 8, true, false>(cx, a); break;
                case 1: conv_dispatch<32, 16, 8, true, false>(cx, a); break;
                default: conv_dispatch<64, 32, 8, true, false>(cx, a); break;
            }
        }
        switch (l) {
            case 0: run_tail<8>(cx, nw, L, lv[l], enc[l].p, up.p, res.p, tmp[l].p, out.p, prof, l); break;
            case 1: run_tail<16>(cx, nw, L, lv[l], enc[l].p, up.p, res.p, tmp[l].p, out.p, prof, l); break;
            default: run_tail<32>(cx, nw, L, lv[l], enc[l].p, up.p, res.p, tmp[l].p, out.p, prof, l); break;
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
        ProfTimer pt(cx, prof, "heads", (double)n0 * (32 + 32 + 4 + 24), 2.0 * (double)n0 * 2 * (64 + 32 + 12));
        HeadParams hp;
        hp.w = nw.head;
        k_heads<<<div_up(n0, 256), 256, 0, st>>>(cur.p, nullptr, nullptr, rows.xyz.p, rows.disp.p, rows.out.p, (int)n0, hp,
                                                 d_newP, d_newD, keep ? (float4*)fb.pred.p : nullptr,
                                                 keep ? (float4*)fb.offset.p : nullptr);
        cx.launches++;
    }
}

}  // namespace sqsm
