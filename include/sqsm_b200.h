/*
 * sqsm_b200.h — C ABI of libsqsm_b200.so, the B200-native (sm_100a) replacement for the
 * `spconv-contraction` skeletonisation stage of project-lightlin/SmartQSM v2.1.3.
 *
 * The reference reaches this path through a Python plug-in class and un-vendored wheels
 * (spconv, Open3D, SciPy); it has no FFI of its own.  The entry points below are what a
 * ctypes binding of that plug-in binds; each one names the reference
 * interface it replaces (paths relative to the reference checkout).
 *
 * Conventions: plain C, no torch types; `h_` pointers are HOST memory, `d_` pointers are
 * DEVICE memory of the handle's GPU; sizes are element counts; every function returns 0 on
 * success and a non-zero status otherwise, with a message available from
 * sqsm_last_error().  A handle is not thread-safe and owns one CUDA stream; several handles
 * (one per process / GPU) can coexist.  There is no CPU fallback: sqsm_create() fails when
 * no sm_100 device is present.
 */
#ifndef SQSM_B200_H
#define SQSM_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct sqsm_engine* sqsm_handle;

/* core_algorithm_kwargs of configs/spconv-contraction-*.yaml:10-18 and the constructor
 * SpconvBasedContraction.__init__ (core/core_algorithms/spconv_based_contraction.py:51-73). */
typedef struct SqsmConfig {
    int32_t device;                     /* CUDA device ordinal ("device: cuda")              */
    int32_t batch_size;                 /* cubes per network batch (16)                       */
    int32_t max_iteration;              /* informational; the caller drives the iterations    */
    float   voxel_size;                 /* 0.01                                               */
    float   cube_size;                  /* 4.0                                                */
    float   buffer_size;                /* 0.4                                                */
    int32_t monitored_by_chamfer_distance; /* 0/1                                             */
    int32_t latch_early_stop;           /* 1: remember a Chamfer rejection (fixes F10; results
                                           identical); 0: recompute like the reference does  */
    int32_t down_kernel_flip;           /* 0: k index = (z&1,y&1,x&1) (SURVEY App. A default) */
    int32_t conv_mode;                  /* arithmetic of the 16/32/64-channel convolutions:
                                           0 = library default (4), 1 = fp32 FFMA everywhere,
                                           2 = tcgen05 3xTF32 split (fp32-class accuracy),
                                           3 = tcgen05 single-pass TF32 (reduced precision, App. D.4),
                                           4 = tcgen05 3xFP16 split (fp32-class; activations must stay < 3e4),
                                           5 = as 4 with the TMA tile::gather4 kernel where it covers the shape */
    int32_t reserved[6];
} SqsmConfig;

/* What one _contract() call did (spconv_based_contraction.py:75-140). */
typedef struct SqsmIterStats {
    int64_t n_in;                       /* points entering the iteration                      */
    int64_t n_entries;                  /* sum over cube samples of points in the halo box    */
    int64_t n_vox;                      /* level-0 voxels (network rows), all batches         */
    int64_t n_rows_l1, n_rows_l2, n_rows_l3;
    int64_t n_out;                      /* interior voxels = points leaving the iteration     */
    int64_t n_cubes;
    int64_t n_batches;
    int64_t far_face_rows;              /* rows on the far faces of a batch box (App. D.3)    */
    int64_t pairs_l0, pairs_l1, pairs_l2, pairs_l3; /* active SubM pairs incl. centre         */
    double  chamfer;                    /* NaN when the monitor is off                         */
    int32_t accepted;                   /* 1 = state replaced, 0 = rejected by the monitor    */
    int32_t skipped;                    /* 1 = latched early stop, nothing computed            */
    int32_t epoch;
    int32_t gpu_launches;               /* kernels launched by this call                       */
    float   ms_tiling, ms_voxel, ms_rulebook, ms_network, ms_chamfer; /* CUDA-event times       */
} SqsmIterStats;

/* ---- life cycle ------------------------------------------------------------------------ */
int  sqsm_create(const SqsmConfig* cfg, sqsm_handle* out);
void sqsm_destroy(sqsm_handle h);
const char* sqsm_last_error(sqsm_handle h);      /* h may be NULL: error of a failed create  */
int  sqsm_abi_version(void);

/* Weights: the reference loads `ckpt["model_state_dict"]` into SmartTreeXX
 * (spconv_based_contraction.py:62-69).  Tensors are passed one by one by their state-dict
 * name, float32, C-contiguous, original layout ([C_out,kz,ky,kx,C_in] for
 * convolutions, [out,in] for nn.Linear).  sqsm_finalize_weights() folds BatchNorm(eps=1e-4)
 * running statistics and uploads everything; it fails if a tensor is missing. */
int  sqsm_set_weight(sqsm_handle h, const char* name, const float* h_data, int64_t n_elem);
int  sqsm_finalize_weights(sqsm_handle h);

/* ---- the plug-in path ------------------------------------------------------------------ */
/* CoreAlgorithmBase.input(points=...) (core/core_algorithms/core_algorithm_base.py:31-34):
 * N x 3 float64, host (copied, cast to float32 like spconv_based_contraction.py:82) or
 * already on the device as float32 (device-resident variant used for kernel-only timing).
 * Resets epoch, displacements and the Chamfer state. */
int  sqsm_set_points_f64(sqsm_handle h, const double* h_xyz, int64_t n);
int  sqsm_set_points_dev_f32(sqsm_handle h, const float* d_xyz, int64_t n);
/* Resume from a saved state: float32 contracted points AND accumulated displacements (host),
 * i.e. the two arrays of SpconvBasedContraction.output() (:147-151).  Epoch restarts at 0. */
int  sqsm_set_state_f32(sqsm_handle h, const float* h_xyz, const float* h_disp, int64_t n);

/* One thunk of get_pipeline() = one SpconvBasedContraction._contract()
 * (spconv_based_contraction.py:75-140).  `stats` may be NULL. */
int  sqsm_contract_step(sqsm_handle h, SqsmIterStats* stats);

/* The same thunk in phases, for ONE tree sharded over several ranks.  Cube batches are
 * independent inside an iteration (synthetic_tree.py:252-281), so rank `shard` of `n_shards` runs voxelisation,
 * rulebooks and the U-Net for a contiguous range of batches only and yields a SEGMENT of the new cloud; the
 * concatenation of the segments in rank order is the new cloud in canonical order (identical to the
 * single-rank result).  The caller exchanges the segments (NCCL), installs the assembled cloud on every rank,
 * all-reduces the partial Chamfer sums and finishes with the same decision everywhere.
 *   sqsm_step_begin(h, r, R, stats)            -> stats.n_out = rows of this rank's segment
 *   sqsm_step_segment(h, &n, d_xyz, d_disp)    -> copy the segment to DEVICE buffers (NULL: size only)
 *   sqsm_step_set_candidate(h, d_xyz, d_disp, n)   assembled new cloud, DEVICE pointers
 *   sqsm_step_chamfer(h, r, R, sums[2], counts[2]) partial sums over rank r's share of the query rows
 *   sqsm_step_finish(h, monitored, chamfer, &accepted)   accept / reject like :123-131 */
int  sqsm_step_begin(sqsm_handle h, int32_t shard, int32_t n_shards, SqsmIterStats* stats);
int  sqsm_step_segment(sqsm_handle h, int64_t* n, float* d_dst_xyz, float* d_dst_disp);
int  sqsm_step_set_candidate(sqsm_handle h, const float* d_xyz, const float* d_disp, int64_t n);
int  sqsm_step_chamfer(sqsm_handle h, int32_t shard, int32_t n_shards, double* sums2, int64_t* counts2);
int  sqsm_step_finish(sqsm_handle h, int32_t monitored, double chamfer, int32_t* accepted);

/* SpconvBasedContraction.output() (spconv_based_contraction.py:147-151): sizes, then copy. */
int  sqsm_result_size(sqsm_handle h, int64_t* n);
int  sqsm_get_result(sqsm_handle h, float* h_contracted_xyz, float* h_displacements_xyz);

/* ThinningBasedSkeletonizationPipeline._produce_skeletal_points_and_radii
 * (core/thinning_based_skeletonization_pipeline.py:68-80) on the current state: one point
 * per `voxel` cell (lowest original index), radii = ||displacement||.  Node numbering =
 * ascending sample id.  Call with NULL outputs to obtain *m only. */
int  sqsm_skeletal_sample(sqsm_handle h, double voxel, int64_t* m,
                          int64_t* h_sample_ids, float* h_points_xyz, float* h_radii);

/* Radius-neighbour edge set of SkeletonizationPipelineBase._construct_graph
 * (core/skeletonization_pipeline_base.py:132-144, merged and sorted at :158) over the
 * skeletal points of the last sqsm_skeletal_sample(): all i<j with ||p_i-p_j|| <= r
 * (float64 test), lexicographic.  Call with h_edges == NULL to obtain *n_edges only. */
int  sqsm_radius_graph(sqsm_handle h, double r, int64_t* n_edges, int64_t* h_edges_ij);

/* The part of that edge list whose first node lies in [node_lo, node_hi): the list of the whole sample is the
 * concatenation over consecutive node ranges, so the ranks of a sharded tree (configuration C5) each produce and copy
 * out one slice.  `capacity` = pairs that fit h_edges_ij (< 0: unchecked).  Not cached. */
int  sqsm_radius_graph_range(sqsm_handle h, double r, int64_t node_lo, int64_t node_hi, int64_t* n_edges,
                             int64_t* h_edges_ij, int64_t capacity);

/* Row f3 - the ball occupancy queries of core/refinement.py:448-456
 * (kdtree.query_ball_point(points[path], r = reference_radii[path] * factor) inside the loop over the paths): the ball of
 * EVERY node with its own radius in one call, as a CSR (h_offsets: M + 1 entries; h_indices: ascending node indices, the
 * node itself included; closed ball, float64 d2 <= r * r like cKDTree; a negative radius acts like its magnitude, NaN gives an empty list).
 * The sequential loop of the reference then only unites precomputed lists.  Call with h_indices == NULL to obtain the
 * offsets and *n_total, then with a buffer of `capacity` >= *n_total entries. */
int  sqsm_ball_lists(sqsm_handle h, const float* h_points_xyz, int64_t m, const double* h_radii, int64_t* h_offsets,
                     int64_t* n_total, int32_t* h_indices, int64_t capacity);

/* Same operator on caller-provided points (host float32 M x 3), independent of the state. */
int  sqsm_radius_graph_points(sqsm_handle h, const float* h_points_xyz, int64_t m, double r,
                              int64_t* n_edges, int64_t* h_edges_ij, int64_t capacity);

/* ---- connectivity graph and skeleton extraction behind the path (SURVEY.md section 8 row a16 = f1) -----------------
 * SkeletonizationPipelineBase._construct_graph (core/skeletonization_pipeline_base.py:75-174): edges of the Delaunay
 * tetrahedra (`h_simplices` = scipy.spatial.Delaunay(points).simplices, int32 [t][4]: Qhull stays on the host, :77), its
 * shortest-path tree from the lowest node of the largest component (networkx dijkstra_predecessor_and_distance,
 * predecessors[0], :104-120) and its minimum spanning edges (Kruskal in networkx's edge order, :122-130), merged with the
 * radius-neighbour edges (r <= 0: none) and optional preset edges into the sorted unique list of :158.
 * `h_points_xyz` == NULL uses the skeletal points (and the cached radius graph) of the last sqsm_skeletal_sample(). */
int  sqsm_graph_construct(sqsm_handle h, const float* h_points_xyz, int64_t m, const int32_t* h_simplices,
                          int64_t n_simplices, double r, const int64_t* h_preset_edges_ij, int64_t n_preset,
                          int64_t* n_edges);
int  sqsm_graph_edges(sqsm_handle h, int64_t* h_edges_ij);              /* int64 [n_edges][2], i < j, lexicographic */
/* {root of the Delaunay graph, unique Delaunay edges, SPT edges, MST edges, merged edges} */
int  sqsm_graph_info(sqsm_handle h, int64_t* info5);
/* Weighted graph (:160-174; weight_fn 0 = "l1", 1 = "l2") and _extract_skeleton with topology_extractor == "spt"
 * (:190-215): root = lowest node of the largest component; h_pred[m] = predecessors[0] of every reached node (-1 for the
 * root and for unreached nodes); h_order[n_reached - 1] = nodes in the order networkx inserted them into the predecessor
 * dict (the reference adds the skeleton's edges in this order, :205-213); h_weight[m] = float32 length of the edge
 * (pred, node).  Any output pointer may be NULL. */
int  sqsm_graph_extract_skeleton(sqsm_handle h, int32_t weight_fn, int64_t* root, int64_t* n_reached, int64_t* h_pred,
                                 int64_t* h_order, float* h_weight);

/* ---- cloud preparation in front of the path ---------------------
 * Replaces entrypoints/smartqsm.py:882-913: bounding box -> global shift -> translate (:882-888),
 * Open3D voxel_down_sample, float64 voxel means (:889), Open3D remove_statistical_outlier
 * (:895-898), density-adaptation scale = voxel_down_size / mean nearest-neighbour distance
 * (SciPy KDTree, :908-913).  Output order: 8^3-voxel bricks by (z, y, x) brick index, then z, y, x
 * inside the brick (Open3D's unordered_map order is not reproducible).  Both filters of the entry point are provided:
 * statistical_outlier_removal (:895-898) and radius_outlier_removal (:899-902, Open3D: more than nb_points points,
 * itself included, with squared distance < radius^2). */
typedef struct SqsmPrepConfig {
    double  voxel_down_size;      /* configs/<name>.yaml: voxel_down_size (0.01) */
    double  std_ratio;            /* kwargs_for_filter.std_ratio (4.0) */
    int32_t filter;               /* 0 = none, 1 = statistical_outlier_removal, 2 = radius_outlier_removal */
    int32_t nb_neighbors;         /* kwargs_for_filter.nb_neighbors (10); 1..32 */
    int32_t density_adaptation;   /* using_density_adaptation_in_skeletonization */
    int32_t keep_debug;           /* keep the intermediate arrays for sqsm_debug_prepared */
    double  radius;               /* filter 2: kwargs_for_filter.radius */
    int32_t nb_points;            /* filter 2: kwargs_for_filter.nb_points */
    int32_t reserved[1];
} SqsmPrepConfig;

typedef struct SqsmPrepStats {
    int64_t n_in, n_voxel, n_kept, gpu_launches;
    int64_t n_far_filter, n_far_scale;   /* queries that needed the far (ring) search in each k-NN pass */
    double  global_shift[3];      /* smartqsm.py:883-887 */
    double  scale;                /* 1.0 without density adaptation */
    double  mean_nn;              /* mean distance to the nearest other point of the kept cloud */
    double  cloud_mean, std_dev, threshold;   /* statistics of remove_statistical_outlier */
    float   ms_total, ms_downsample, ms_filter, ms_scale;
} SqsmPrepStats;

/* Runs the whole preparation on host float64 N x 3 points; the prepared cloud stays on the device. */
int  sqsm_prepare_cloud_f64(sqsm_handle h, const double* h_points_xyz, int64_t n,
                            const SqsmPrepConfig* cfg, SqsmPrepStats* stats);
int  sqsm_prepared_size(sqsm_handle h, int64_t* n_kept, int64_t* n_voxel);
/* The cloud the reference holds in `points` after :904 (shifted, down-sampled, filtered; NOT scaled). */
int  sqsm_get_prepared(sqsm_handle h, double* h_points_xyz);
/* CoreAlgorithmBase.input(points = prepared * scale) without the host round trip (:922 + epoch-1 cast). */
int  sqsm_set_points_prepared(sqsm_handle h);
/* keep_debug only: voxel means float64[n_voxel,3], avg k-NN distance float64[n_voxel] (filter 1; with filter 2 the
 * neighbour count as a double), keep uint8[n_voxel] (the last two only with a filter); any pointer may be NULL. */
int  sqsm_debug_prepared(sqsm_handle h, double* h_voxel_means, double* h_avg_distance, uint8_t* h_keep);

/* ---- parity / inspection hooks (integer outputs of the LAST sqsm_contract_step) --------
 * Rows are reported in CANONICAL order: by cube sample, then ascending
 * winning point index.  Enable capture before the step; it keeps the per-level tables. */
int  sqsm_debug_capture(sqsm_handle h, int32_t enable);
int  sqsm_debug_level_size(sqsm_handle h, int32_t level, int64_t* n_rows);
/* level 0: indices = (sample, z, y, x) int32[n,4]; winner = input row int64[n]; mask uint8[n];
 * batch spatial_shape (max index z,y,x) int32[n_batches,3] via sqsm_debug_batch_shapes. */
int  sqsm_debug_voxels(sqsm_handle h, int32_t* h_indices, int64_t* h_winner, uint8_t* h_mask);
int  sqsm_debug_batch_shapes(sqsm_handle h, int32_t* h_shapes_zyx, int32_t* h_descend_flags);
/* per level: coordinates (sample, z, y, x) of every row in internal order, SubM table
 * nbr[27][n] (internal row numbers, -1 = none) and, for levels < 3, parent[n] / koff[n]. */
int  sqsm_debug_level_tables(sqsm_handle h, int32_t level, int32_t* h_coords, int32_t* h_nbr,
                             int32_t* h_parent, int32_t* h_koff);
/* level-0 permutation: canonical row -> internal row. */
int  sqsm_debug_canon_to_internal(sqsm_handle h, int32_t* h_perm);
/* network outputs per level-0 row, internal order: direction*distance float32[n,3], plus the
 * raw heads offset[n,3], regression[n] */
int  sqsm_debug_predictions(sqsm_handle h, float* h_pred, float* h_offset, float* h_regression);
/* activations after a named stage, internal row order of `level`; float32[n, channels] */
int  sqsm_debug_features(sqsm_handle h, const char* stage, int32_t* level, int32_t* channels,
                         float* h_data, int64_t capacity);

/* ---- measurement hooks ------------------------------------------------------------------ */
/* Kernel-family timing: CUDA events recorded on the handle's stream around every launch group of
 * a family ("tiling", "voxelise", "rulebook", "subm_l0".."subm_l3", "down_up", "heads", "chamfer")
 * while enabled; resolved lazily, so enabling it adds no synchronisation to a step.  Returns the
 * accumulated device time, launch count, algorithmic bytes and FLOPs since the last reset. */
int  sqsm_profile_enable(sqsm_handle h, int32_t enable);
int  sqsm_profile_reset(sqsm_handle h);
int  sqsm_profile_get(sqsm_handle h, const char* family, double* total_ms, int64_t* launches,
                      double* algorithmic_bytes, double* flops);
int  sqsm_profile_families(sqsm_handle h, char* buf, int64_t capacity);   /* comma separated */
/* Whole-step device timer: events on the handle's stream (torch's events only see torch's stream). */
int  sqsm_timer_start(sqsm_handle h);
int  sqsm_timer_stop(sqsm_handle h, double* elapsed_ms);
int  sqsm_launch_count(sqsm_handle h, int64_t* kernels_launched_since_create);
/* The handle's CUDA stream (cudaStream_t): a caller that exchanges segments between ranks (sqsm_step_*) enqueues its
 * collectives relative to it (e.g. torch.cuda.ExternalStream) instead of synchronising the device. */
int  sqsm_get_stream(sqsm_handle h, void** cuda_stream);
/* sizeof() of the public structs as the library was compiled: 0 SqsmConfig, 1 SqsmIterStats, 2 SqsmPrepConfig,
 * 3 SqsmPrepStats (-1 otherwise) - lets a binding verify its own declarations without a GPU. */
int  sqsm_sizeof(int32_t which);

#ifdef __cplusplus
}
#endif
#endif /* SQSM_B200_H */
