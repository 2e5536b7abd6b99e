// Internal structures shared by the translation units of libsqsm_b200.
#pragma once
#include "brickmap.h"
#include "../../include/sqsm_b200.h"
#include <map>
#include <string>
#include <vector>

namespace sqsm {

constexpr int kLevels = 4;                      // u_channels = [8, 16, 32, 64]
constexpr int kPlanes[kLevels] = {8, 16, 32, 64};
constexpr int kMaxIdxF32 = 1 << 24;             // F7: sample-local index carried as float32

// One cube sample (synthetic_tree.py:252-281), float32 exactly as the reference computes it.
struct SampleInfo {
    float slo[3];   // halo box min (x, y, z)
    float shi[3];   // halo box max
    float lo[3];    // cube min
    float hi[3];    // cube max
    int32_t grid[3];// voxels per axis, round((shi-slo)/vsize)
    int32_t cube[3];// cube index (x, y, z)
    int32_t pad;
};

// Per U-Net level integer tables of one super-batch (internal row order).
struct LevelTables {
    BrickMap map;
    int64_t n = 0;                      // rows
    DBuf<int32_t>  row_brick;           // [n]
    DBuf<uint16_t> row_pos;             // [n]
    DBuf<int32_t>  nbr;                 // [27][n]  SubM rulebook, output-stationary: input row or -1 (absent when only
                                        //          the compact form below is wanted)
    // compact (CSR) form of the same rulebook for the sparse-loop kernel (level 0: 5 of 27 neighbours exist):
    DBuf<uint32_t> nbr_mask;            // [n] bit k = kernel offset k present
    DBuf<uint32_t> nbr_off;             // [n] first entry of the row in nbr_list
    DBuf<int32_t>  nbr_list;            // present input rows, ascending kernel offset inside a row
    DBuf<int32_t>  parent;              // [n]      coarse row or -1              (levels 0..2)
    DBuf<int32_t>  koff;                // [n]      kz*4+ky*2+kx                  (levels 0..2)
    DBuf<int32_t>  child;               // [8][n_next] fine row or -1 (down conv, output-stationary)
    DBuf<uint8_t>  row_desc;            // [n] 1 if the row's batch descended below this level
    DBuf<int32_t>  shape;               // [n_batches][3] spatial shape (z,y,x) of this level per batch
    DBuf<int32_t>  desc;                // [n_batches] batch descended from this level
    int64_t pairs = 0;                  // active SubM pairs (incl. centre)
};

// Level-0 row payload.
struct Level0Rows {
    DBuf<float4>  xyz;                  // [n0] features (x, y, z, 0)
    DBuf<float4>  disp;                 // [n0] accumulated displacement of the winning point
    DBuf<int32_t> pt;                   // [n0] winning input row
    DBuf<int32_t> out;                  // [n0] output position (interior rows) or -1
    DBuf<int32_t> canon;                // [n0] canonical row number within the super-batch
    DBuf<uint8_t> interior;             // [n0]
};

struct ConvW {                          // one convolution, device resident
    int ci = 0, co = 0, kv = 0;
    float* w = nullptr;                 // [kv][ci][co]
    float* bn = nullptr;                // [2][ci] alpha, beta of the BatchNorm+ReLU in FRONT of the conv (or null)
    float* wimg = nullptr;              // tensor-core weight image, TF32 hi/lo (conv_tc.cu) or null
    float* wimg16 = nullptr;            // tensor-core weight image, FP16 hi/lo
    float* wimg_tma = nullptr;          // weight image of the TMA-gather kernel (conv_tma.cu) or null
};

struct HeadW {                          // output_layer BN + both MLP heads; index 0 = offset, 1 = regression
    float bn0[2][8];                    // alpha, beta of output_layer.0
    float w1[2][8][8], b1[2][8], a1[2][8], c1[2][8];   // Linear(8,8) then BN(8) as alpha/beta
    float w2[2][4][8], b2[2][4], a2[2][4], c2[2][4];   // Linear(8,4) then BN(4)
    float w3[2][3][4], b3[2][3];                       // Linear(4,3) / Linear(4,1) (row 0)
};

struct NetWeights {
    std::map<std::string, ConvW> conv;
    DBuf<float> blob;
    HeadW head;
    bool ready = false;
    int mode = 3;                       // 0 = fp32 FFMA everywhere, 1 = tcgen05 3xTF32 (fp32-class), 2 = tcgen05 single-pass TF32,
                                        // 3 = tcgen05 3xFP16 (2-term FP16 split, fp32-class, half the operand bytes),
                                        // 4 = 3xFP16 with the TMA gather4 kernel where it covers the shape (else as 3)
    int* range_flag = nullptr;          // device flag: an activation left the FP16 range in mode 3
    int tc_min_co = 16;                 // layers with at least this many output channels use the tensor-core kernel
                                        // (measured: with FP16 operands the tcgen05 kernel wins from C=16 up; with TF32
                                        // operands only from C=32)
};

constexpr int kTmZeroRows = 1024;        // rows of zeros behind every split-row array (absent neighbours, conv_tma.cu)

struct Net;                             // conv.cu

// Operands a tensor-core epilogue writes for ONE consumer of its output: FP16 hi / lo split of relu(out * alpha + beta)
// (alpha == nullptr: of the raw output).  conv_tc.cu.
struct TcFused {
    void* hi = nullptr;
    void* lo = nullptr;
    const float* alpha = nullptr;
    const float* beta = nullptr;
};

// conv.cu
void net_upload(Ctx& cx, NetWeights& nw, const std::map<std::string, std::vector<float>>& state, bool down_flip);
struct ForwardBuffers {
    DBuf<float> pred;                   // [n0][4] direction*distance (w = distance)
    DBuf<float> offset;                 // debug [n0][4] (x,y,z offset, w = regression)
    std::map<std::string, DBuf<float>> keep;   // debug: named activations
    std::map<std::string, std::pair<int, int>> keep_meta;   // name -> (level, channels)
};
struct ConvProfile { double ms = 0; int64_t launches = 0; double bytes = 0; double flops = 0; };
void net_forward(Ctx& cx, const NetWeights& nw, LevelTables* lv, const Level0Rows& rows, int n_levels,
                 ForwardBuffers& fb, bool keep, float* d_newP, float* d_newD);

// structure.cu
struct TilingResult {
    int n_samples = 0, n_batches = 0;
    std::vector<SampleInfo> samples;            // host copy, cube order (lexicographic x,y,z)
    DBuf<SampleInfo> d_samples;
    DBuf<uint32_t> ent_sample;                  // [E] sorted by sample (stable: ascending point index inside)
    DBuf<uint32_t> ent_val;                     // [E] point index | interior << 31
    std::vector<int64_t> start;                 // [n_samples+1] entry range of every sample
    int64_t n_entries = 0;
};
void tile_points(Ctx& cx, const float* d_P, int64_t n, float cube_size, float buffer_size, float voxel_size,
                 int batch_size, TilingResult& out);
struct VoxelStats { int64_t far_face = 0; };
void build_level0(Ctx& cx, const float* d_P, const float* d_D, const TilingResult& tl, int s0, int s1,
                  int batch_size, float voxel_size, LevelTables& L0, Level0Rows& rows, int64_t out_base,
                  int64_t* n_out, DBuf<int32_t>& ent_row_out);
void build_subm(Ctx& cx, LevelTables& L, int batch0, int n_batches, int batch_size, int64_t* far_face,
                bool want_dense = true, bool want_csr = false);
bool net_level0_uses_csr(const struct NetWeights& nw);     // conv.cu: the level-0 layers run on the sparse-loop kernel
bool build_down(Ctx& cx, LevelTables& L, LevelTables& Lnext, int batch0, int n_batches, int batch_size);

// spatial.cu
double chamfer_distance(Ctx& cx, const float* d_P, int64_t np, const float* d_Q, int64_t nq, double voxel);
void chamfer_partial(Ctx& cx, const float* d_P, int64_t np, const float* d_Q, int64_t nq, double voxel, int shard,
                     int n_shards, double sums[2], int64_t counts[2]);
struct SkeletalResult {
    DBuf<int64_t> ids;                  // [m] ascending original index
    DBuf<float> pts;                    // [m][3]
    DBuf<float> radii;                  // [m]
    int64_t m = 0;
};
void skeletal_sample(Ctx& cx, const float* d_P, const float* d_D, int64_t n, double voxel, SkeletalResult& out);
// row f3: the ball of every node with its own radius (cKDTree.query_ball_point semantics), CSR, ascending indices
void ball_lists(Ctx& cx, const float* d_pts, int64_t m, const double* d_radii, double finest_cell, double largest_radius,
                DBuf<int64_t>& offsets, DBuf<uint32_t>& indices, int64_t* n_total);
void radius_graph(Ctx& cx, const float* d_pts, int64_t m, double r, DBuf<int64_t>& edges, int64_t* n_edges,
                  int64_t node_lo = 0, int64_t node_hi = -1);

// graph.cu (row a16 / f1: core/skeletonization_pipeline_base.py:75-238)
struct GraphState {
    DBuf<float> pts;                    // [m][3] skeletal points
    std::vector<float> h_pts;
    int64_t m = 0;
    DBuf<uint64_t> merged;              // merged edge keys (u << 32 | v, u < v), lexicographic (:158)
    int64_t n_edges = 0;
    int64_t delaunay_root = -1;
    int64_t n_spt = 0, n_mst = 0, n_delaunay = 0;
    bool valid = false;
};
void graph_construct(Ctx& cx, GraphState& gs, const float* d_pts, int64_t m, const int32_t* h_simplices, int64_t T,
                     const int64_t* d_radius, int64_t n_radius, const int64_t* h_preset, int64_t n_preset);
void graph_edges(Ctx& cx, const GraphState& gs, int64_t* h_out);
void graph_extract_skeleton(Ctx& cx, const GraphState& gs, int weight_fn, int64_t* root_out, int64_t* n_reached, int64_t* h_pred,
                            int64_t* h_order, float* h_weight);

// preprocess.cu (row f2: entrypoints/smartqsm.py:882-913)
struct PrepState {
    DBuf<double> pts;                   // [n_kept][3] shifted, down-sampled, filtered (not scaled)
    DBuf<double> vox, avg;              // keep_debug: voxel means [n_voxel][3], avg k-NN distance [n_voxel]
    DBuf<uint8_t> keep;                 // keep_debug: filter decision per voxel mean
    int64_t n_kept = 0, n_voxel = 0;
    double scale = 1.0;
    bool valid = false;
};
void prepare_cloud(Ctx& cx, const double* h_xyz, int64_t n, const SqsmPrepConfig& cfg, PrepState& ps, SqsmPrepStats& out);
void prepared_to_f32(Ctx& cx, const PrepState& ps, float* d_out);

}  // namespace sqsm
