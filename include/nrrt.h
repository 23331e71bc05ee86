/*
 * nrrt.h — C ABI of libnrrt.so: the B200 (sm_100a) implementation of the neural-guided RRT* hot path of
 * cezarywawrzyniak/mgr_sampling_cnn.  This is the drop-in boundary: the reference has no FFI of its
 * own — its boundary is the Python class surface — so every entry point cites the reference lines it replaces.
 *
 * Conventions
 *   - Every pointer is a DEVICE pointer owned by the caller unless the name ends in `_host`.
 *   - Every call enqueues on `stream` (a cudaStream_t passed as void*) and returns; nothing synchronises.
 *   - Return value: 0 on success, a negative NrrtError otherwise; nrrt_last_error() gives the message (thread-local).
 *   - The library allocates nothing persistent and keeps no state besides that message.
 *   - Positions are (y,x) in 2D and (x,y,z) = array axes (0,1,2) in 3D, exactly as in the reference.
 */
#ifndef NRRT_H_
#define NRRT_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NRRT_VERSION 100

typedef enum NrrtError {
    NRRT_OK = 0,
    NRRT_ERR_BAD_ARG = -1,      /* bad shape / null pointer / unsupported size */
    NRRT_ERR_WORKSPACE = -2,    /* workspace too small */
    NRRT_ERR_CUDA = -3,         /* a CUDA runtime call failed (message has the CUDA error string) */
    NRRT_ERR_UNSUPPORTED = -4   /* device is not sm_100 or shared memory budget exceeded */
} NrrtError;

/* Per-task status written to NrrtPlanOut.status. */
typedef enum NrrtTaskStatus {
    NRRT_SOLVED = 0,               /* goal_reached() returned True            (rrt_star.py:203-208) */
    NRRT_MAX_ITER_BEST_EFFORT = 1, /* loop ran out; path ends at best_node    (rrt_star.py:210-211) */
    NRRT_NO_NODE_CONNECTED = 2,    /* best_node is None -> reference returns [] (find_path(None)) */
    NRRT_SAMPLER_STUCK = 3         /* bounded rejection hit max_reject; the reference would spin forever
                                      (rrt_star.py:64-68 / ThreeD_neural_rrt.py:81-108) */
} NrrtTaskStatus;

/* Planner configuration = the constructor arguments of the four reference RRTStar classes
 * (rrt_star.py:49-61, neural_rrt.py:50-64, ThreeD_rrt_star.py:50-62, ThreeD_neural_rrt.py:54-68). */
typedef struct NrrtPlanParams {
    int32_t dim;            /* 2 or 3 */
    int32_t shape[3];       /* occ_map.shape; 2D: (map_height, map_width, 1) */
    int32_t max_iterations; /* MAX_ITERATIONS (5000 in the reference) */
    int32_t node_stride;    /* rows per task in pos/cost/parent outputs; must be >= max_iterations + 2 */
    int32_t max_reject;     /* bound on consecutive rejected draws inside one sample */
    int32_t path_cap;       /* points per task in path_idx (0 = no path extraction) */
    double goal_threshold;  /* GOAL_THRESHOLD: 5.0 (2D) / 3.0 (3D) */
    double neural_bias;     /* neural_bias (0.75 in the reference); ignored when cdf == NULL */
    uint64_t seed;          /* Philox key of the sample stream */
    uint64_t draw_start;    /* index of the first draw to consume from every task's stream (0 = a fresh stream); with
                               n_draws this lets a caller continue one stream across calls, the way the reference's methods
                               all consume the one module-level `random` */
    int32_t sampler_mode;   /* nrrt_draw_samples: 0 = the main loop's `random.random() < neural_bias` switch per sample
                               (neural_rrt.py:225-228); 1 = generate_neural_sample() only, 2 = generate_random_sample() only
                               (the methods called directly: no bias draw is consumed) */
} NrrtPlanParams;

/* Per-task outputs; all arrays caller-allocated, fixed stride, any pointer marked optional may be NULL. */
typedef struct NrrtPlanOut {
    double* pos;        /* [B][node_stride][dim]  Node.position, insertion order (node index).  Required: it is also the
                           kernel's store for nodes with non-integer coordinates (integer nodes are int16 words in
                           shared memory), filled completely at the end of every task */
    double* cost;       /* required [B][node_stride]       Node.cost (also the kernel's working storage) */
    int32_t* parent;    /* required [B][node_stride]       index of Node.parent, -1 for the start node */
    int32_t* n_nodes;   /* [B] len(self.nodes) */
    int32_t* iterations;/* [B] iteration_no returned by rrt_star() (1-based) */
    int32_t* status;    /* [B] NrrtTaskStatus */
    int32_t* goal_node; /* [B] index of the goal-reaching node (== n_nodes, stored but not counted) or -1 */
    int32_t* best_node; /* [B] self.best_node index or -1 */
    double* best_dist;  /* [B] self.best_distance */
    int32_t* path_n;    /* optional [B] number of points of find_path() (rrt_star.py:163-172) */
    int32_t* path_idx;  /* optional [B][path_cap] node indices start -> end */
    double* path_len;   /* optional [B] calculate_length(path) (test_network_and_algorithms.py:52-56) */
    int32_t* trace;     /* optional [B][trace_cap][4] (iteration, kind 0=connect 1=loop A 2=loop B, node j, verdict):
                           the collision tests the reference would execute, in its order */
    int32_t* trace_n;   /* optional [B] number of trace records produced (may exceed trace_cap) */
    int32_t trace_cap;
} NrrtPlanOut;

int nrrt_version(void);
const char* nrrt_last_error(void);

/* Number of uint32 words of one bit-packed occupancy map (bit c of word c>>5 = cell c row-major, 1 = FREE). */
size_t nrrt_words_per_map(int64_t cells);

/* Bit-pack occupancy maps.  Replaces the per-cell tests `occ_map[y, x] != 0` (2D free, rrt_star.py:67,120) and
 * `occ_map[x, y, z] == 0` (3D free, ThreeD_rrt_star.py:69,128).  dtype: 0 = uint8, 1 = float32, 2 = float64.
 * occ: [n_maps][cells];  bits: [n_maps][nrrt_words_per_map(cells)]. */
int nrrt_pack_occupancy(const void* occ, int dtype, int free_is_nonzero, int n_maps, int64_t cells, uint32_t* bits,
                        void* stream);

/* CNN-output post-processing + CDF of the neural sampler, hoisted out of the per-draw path:
 *   2D (three_d = 0): h = clamp(logits,-3,1) + 10                      neural_rrt.py:351 + :64
 *   3D (three_d = 1): v=(l-min)/(max-min); v[v<=0.96]=0; h = v - 250   ThreeD_neural_rrt.py:339-346 + :68
 *   mode 2        : h = logits as given (already the planner-side heat map)
 *   then w = h - min(h); p = w / np.sum(w); cdf = np.cumsum(p)         neural_rrt.py:74-86 (fp32, NumPy rounding)
 * logits, cdf: [B][cells] float32.  workspace: nrrt_cdf_workspace_bytes(B, cells) bytes. */
size_t nrrt_cdf_workspace_bytes(int B, int64_t cells);
int nrrt_cdf_build(const float* logits, int mode, int B, int64_t cells, float* cdf, void* workspace, size_t ws_bytes,
                   void* stream);

/* Draw the sample point of every iteration of every task:
 * generate_random_sample (rrt_star.py:63-68, ThreeD_rrt_star.py:64-70), generate_neural_sample
 * (neural_rrt.py:73-99, ThreeD_neural_rrt.py:80-110) and the `random.random() < neural_bias` switch (neural_rrt.py:225).
 * cdf == NULL selects the uniform-only modules.  samples: [B][max_iterations][dim] int16.
 * task_ids[b] is the Philox stream id of task b; draw_status[b] = 0 or NRRT_SAMPLER_STUCK; n_draws optional. */
int nrrt_draw_samples(const NrrtPlanParams* p, const uint32_t* occ_bits, const int32_t* task_to_map, const float* cdf,
                      const int32_t* task_ids, int B, int16_t* samples, int8_t* neural_flag, int64_t* n_draws,
                      int32_t* draw_status, void* stream);

/* RRT* main loop for B independent tasks (rrt_star.py:184-215 and twins): nearest neighbour, steer, collision test,
 * near set, choose-parent, rewire, goal test; one task per CTA.  radius: [max_iterations] fp64 table of
 * compute_search_radius (rrt_star.py:174-182), computed by the host with Python `math`.
 * samples as produced by nrrt_draw_samples (or pre-drawn by the caller: parity level A).
 * draw_status may be NULL; tasks with draw_status != 0 are skipped and report that status.
 * queue_counter: one int32 of device scratch (zeroed by the call). */
int nrrt_plan_batch(const NrrtPlanParams* p, const uint32_t* occ_bits, const int32_t* task_to_map, const double* radius,
                    const int16_t* samples, const int32_t* starts, const int32_t* goals, const int32_t* draw_status,
                    const NrrtPlanOut* out, int B, int32_t* queue_counter, void* stream);

/* Batched stand-alone primitives behind the reference's individually callable methods.
 * nrrt_collision_free: is_collision_free(p1, p2) (rrt_star.py:108-123 / ThreeD_rrt_star.py:115-132) for S segments;
 *   seg_map[s] selects the occupancy map; p1, p2: [S][dim] fp64; verdict[s] = 1 (free) / 0.
 * nrrt_tree_search: for Q query points over one tree of n nodes ([n][dim] fp64, insertion order):
 *   nearest[q] = find_nearest_neighbor (rrt_star.py:70-80); near_mask[q][(n+31)/32] (optional) = bit j set iff
 *   euclid(query, node_j) <= radii[q] (find_nearby_nodes, rrt_star.py:149-154). */
int nrrt_collision_free(int dim, const int32_t* shape_host, const uint32_t* occ_bits, const int32_t* seg_map, const double* p1,
                        const double* p2, int S, int32_t* verdict, void* stream);
int nrrt_tree_search(int dim, const double* nodes, int n, const double* queries, const double* radii, int Q,
                     int32_t* nearest, uint32_t* near_mask, void* stream);

/* Shared memory one planner CTA needs for these params (bytes), and CTAs per SM the launch will get (info only). */
int nrrt_plan_smem_bytes(const NrrtPlanParams* p, size_t* bytes, int* ctas_per_sm);

/* ------------------------------------------------------------------------------------------------------------------
 * CNN forward (UNet_cooler.forward train.py:171-203, ThreeD_UNet_cooler.forward ThreeD_train.py:169-204).
 * Activations are fp16 channels-last (NHWC / NDHWC); accumulation is fp32 in tensor memory.
 * ------------------------------------------------------------------------------------------------------------------ */
typedef enum NrrtConvKind {
    NRRT_CONV_K3S1 = 0,  /* nn.Conv2d/3d(k=3, s=1, p=1): encoder stem, ResNet blocks, decoder convs */
    NRRT_CONV_K3S2 = 1,  /* nn.Conv2d/3d(k=3, s=2, p=1): first conv of layer2/3/4 */
    NRRT_CONV_K1S2 = 2,  /* nn.Conv2d/3d(k=1, s=2):      ResNet downsample branch */
    NRRT_CONV_K1S1 = 3,  /* nn.Conv2d/3d(k=1) */
    NRRT_CONVT_K2S2 = 4, /* nn.ConvTranspose2d/3d(k=2, s=2): upconv6/7/8 */
    NRRT_CONV_UPFOLD = 5 /* (2D) nn.ConvTranspose2d(k=2, s=2) composed with the 3x3 conv that consumes it (train.py:196-201:
                            upconv7 -> conv7[0], upconv8 -> conv8[0]; nothing non-linear in between).  in = the ConvTranspose's
                            input on its (in_h, in_w) grid, cin = its input channels; out / res / poly live on the
                            (2 in_h, 2 in_w) grid; weight = fp16 [cout][16 * cin + 128]: [4 output parities (row, col)][2x2 taps][cin]
                            composite weights, then 128 identity columns (row r has its 1 in column r % 128) through which the
                            tensor core accumulates the residual; poly = nrrt_poly_fold output [n][4][9][4][cout]; res is
                            required. */
} NrrtConvKind;

/* One fused layer: out = [relu2(post_scale * ] relu( scale * conv(in) + shift + residual ) [ + post_shift)].
 * weight: fp16 [rows][taps][cin], rows = cout (conv) or groups*cout (ConvTranspose: group g = (a*2+b) in 2D,
 *         ((a*2+b)*2+c) in 3D for output offset (a,b[,c])), taps in (kd, kh, kw) order; scale/shift: fp32 [rows]. */
typedef struct NrrtConvLayer {
    int32_t dims;            /* 2 or 3 */
    int32_t kind;            /* NrrtConvKind */
    int32_t n, in_d, in_h, in_w;   /* input grid (in_d ignored in 2D) */
    int32_t cin;             /* input channels read (multiple of 64) */
    int32_t cout;            /* output channels (multiple of 32; rows must tile by 64) */
    const void* in;          /* fp16, first channel of the slice to read */
    int64_t in_cstride;      /* channels per pixel of the input buffer */
    const void* weight;
    const float* scale;      /* folded eval-mode BatchNorm + bias: out = acc * scale[c] + shift[c].  Both NULL = identity */
    const float* shift;      /* (layers whose bias the caller has put into the constant term of `poly`; stored outputs only) */
    void* out;               /* fp16 output buffer base */
    int64_t out_cstride;     /* channels per pixel of the output buffer (concat target) */
    int64_t out_coffset;     /* first output channel inside that buffer */
    const void* res;         /* optional fp16 residual, same pixel grid as out */
    int64_t res_cstride, res_coffset;
    int32_t relu;
    const float* post_scale; /* optional second affine + ReLU (3D stem: conv, ReLU, BatchNorm3d, ReLU) */
    const float* post_shift;
    const float* poly;       /* optional (2D) fp32 [n][poly_cases][4][cout] from nrrt_coords_poly: the contribution of the
                                interpolated coordinate channels, added as E0 + ty*E1 + tx*E2 + ty*tx*E3 before the ReLU */
    int32_t poly_cases;      /* 9 for a 3x3 conv (border cases), 4 for a 2D ConvTranspose (one per output offset) */
    const float* head_weight;/* optional fused final_conv (1x1, cout -> 1): when `logits` is set the layer's activation is not
                                stored (out may be NULL); logits[pixel] = head_bias + sum_c act[c] * head_weight[c] */
    float head_bias;
    float* logits;           /* fp32 [n][pixels] */
    int32_t force_per_tap;   /* testing: use the generic per-tap kernel even where the slab kernel applies */
    /* Optional second input (K split): the layer's input channels are [in: cin channels | in2: cin2 channels], the
     * torch.cat of the reference (train.py:194,197,200) without a concatenated buffer.  in_index / in2_index (int32 [n],
     * optional) redirect sample b of a source to another row of its buffer: decoder layers read the encoder output of
     * map index[b], which exists once per map, not once per task.  in_n / in2_n = samples in those buffers (0 = n). */
    const void* in2;
    int64_t in2_cstride;
    int32_t cin2;
    int32_t in_n, in2_n;
    const int32_t* in_index;
    const int32_t* in2_index;
    /* Optional sample indirection of the residual (int32 [n]): sample b adds row res_index[b] of `res`.  The decoder's
     * first conv of every stage is linear in torch.cat((up, enc)) (train.py:194-201), so the encoder-side partial sum is
     * computed once per map and enters the per-task layer here, before the ReLU. */
    const int32_t* res_index;
    int32_t res_n;           /* samples in the residual buffer (0 = n) */
    int32_t no_poly_mma;     /* testing: evaluate the whole polynomial term in the epilogue (default: the interior polynomial
                                of residual-carrying 3x3 layers is one extra tensor-core MMA per tile) */
    int32_t no_tma_store;    /* testing: direct global stores from the epilogue threads instead of staged TMA stores */
    const float* poly3;      /* optional (3D, 128 output channels, residual, H and W multiples of 16): the coordinate field
                                E [n][54][4][cout] of nrrt_coords_poly_gemm (cases = 54) added before the ReLU in the epilogue,
                                instead of the nrrt_poly_relu_3d pass over the stored activation */
    int32_t res_mma;         /* NRRT_CONV_UPFOLD only: accumulate the residual tile with the tensor core (identity columns of the
                                weight) instead of staging it for the epilogue; same results */
    int32_t res_group;       /* scheduling hint (pair slab kernel): consecutive samples per group that share residual rows
                                (res_index: the tasks of one map); tiles run tile-major inside a group so that the shared
                                residual is read from DRAM once per group.  0 / 1 = sample-major.  Results do not depend on it. */
    int32_t tag;             /* caller's label of the layer (NrrtLayerTag for the model graphs); only reported back by the
                                launch profiler below */
} NrrtConvLayer;

int nrrt_conv_layer(const NrrtConvLayer* layer, void* stream);

/* Launch profiler of nrrt_conv_layer (measurement only; off by default; per host thread).  While it is on, EVERY launch is
 * counted with its algorithmic FLOPs (2 x MACs over the true input channels) and bytes (every operand and the output
 * once), and every `sample_every`-th launch is bracketed by CUDA events on its stream.  nrrt_conv_profile_read (after
 * the stream has been synchronised) returns the totals and, per timed launch, tag / milliseconds / FLOPs / bytes. */
int nrrt_conv_profile_start(int sample_every);   /* clears earlier records and turns counting on */
int nrrt_conv_profile_stop(void);                /* pause (records are kept) */
int nrrt_conv_profile_resume(void);              /* continue counting into the same records */
int nrrt_conv_profile_read(int64_t* launches_seen, double* flops_seen, double* bytes_seen, int max_records, int32_t* tags,
                           float* ms, double* flops, double* bytes, int32_t* n_records);

/* ------------------------------------------------------------------------------------------------------------------
 * The whole forward in ONE call: UNet_cooler.forward (train.py:171-203) / ThreeD_UNet_cooler.forward
 * (ThreeD_train.py:169-204) for a batch of tasks fed from resident occupancy maps, with the exact rewrites of DESIGN.md
 * 4.0 (encoder once per distinct map, first decoder convs split into per-map and per-task parts, coordinate channels as
 * epilogue polynomials, ConvTranspose folded into the conv behind it).  The layer graph lives in csrc/forward.cu.
 * ------------------------------------------------------------------------------------------------------------------ */
typedef struct NrrtConvW {   /* one packed layer (device pointers) */
    const void* weight;      /* fp16, layout as NrrtConvLayer.weight for `kind`; NULL = layer absent */
    const float* scale;      /* fp32 [rows] or NULL (identity, see NrrtConvLayer) */
    const float* shift;
    const float* post_scale; /* optional second affine (3D stem) */
    const float* post_shift;
    int32_t kind;            /* NrrtConvKind */
    int32_t cin, cin2, cout; /* channels read from the first / second source; output channels (per group for ConvTranspose) */
    int32_t relu;
    int32_t tag;             /* NrrtLayerTag */
} NrrtConvW;

typedef enum NrrtLayerTag {  /* labels of the model's tensor-core layers (profiler output, per-layer tables) */
    NRRT_TAG_STEM2 = 1,
    NRRT_TAG_ENC = 10,       /* + (stage * 2 + block) * 3 + {0: conv1, 1: conv2, 2: downsample}, stage 0..3 = conv2..conv5 */
    NRRT_TAG_UP6 = 40, NRRT_TAG_UP7, NRRT_TAG_UP8,
    NRRT_TAG_C6FULL = 45, NRRT_TAG_C6E, NRRT_TAG_C7E, NRRT_TAG_C8E,
    NRRT_TAG_C60 = 50, NRRT_TAG_C61, NRRT_TAG_C70, NRRT_TAG_C71, NRRT_TAG_C80, NRRT_TAG_C81,
    NRRT_TAG_C7F = 60, NRRT_TAG_C8F
} NrrtLayerTag;

/* Packed parameters of one model on one device.  Produced by the host packer (mgr_sampling_cnn_b200/cnn.py: UNetPlan,
 * which folds eval-mode BatchNorm, splits / composes the decoder weights and computes the tap moments in host
 * arithmetic) -- layouts are documented at the consuming entry points above.  Pointers not used by `dims` are NULL. */
typedef struct NrrtCnnWeights {
    int32_t dims;                 /* 2: UNet_cooler, 3: ThreeD_UNet_cooler */
    int32_t stem_cin, stem_cout;  /* first conv: 3 -> 32 (2D) / 1 -> 32 (3D) */
    const float* stem_w;          /* fp32 [stem_cout][stem_cin * taps] */
    const float* stem_b;
    NrrtConvW stem2;              /* conv1[2] (+ the 3D stem's BatchNorm + ReLU as post affine) */
    NrrtConvW enc[4][2][3];       /* conv2..conv5 = ResNet layer1..4: [stage][block][conv1, conv2, downsample] */
    const float* coords_w[4];     /* conv_coords_64/128/256/512: fp32 [tap][cin][cout] */
    const float* coords_b[4];
    NrrtConvW up6, up7, up8;      /* ConvTranspose k2 s2 (2D: up6 over x5 only, its coordinate rows are moments) */
    NrrtConvW c6full;             /* 2D: conv6[0] over (upconv6(x5) | x4), per map */
    NrrtConvW c6e, c7e, c8e;      /* per-map parts of conv6[0] (3D) / conv7[0] / conv8[0] over the encoder channels */
    NrrtConvW c60, c61, c70, c71, c80, c81;  /* per-task layers: conv6[0] (3D), conv6[2], conv7[0], conv7[2], conv8[0], conv8[2] */
    NrrtConvW c7f, c8f;           /* 2D: upconv7 -> conv7[0] and upconv8 -> conv8[0] folded (NRRT_CONV_UPFOLD) */
    const float* bc7;             /* 2D: fp32 [9][256] / [9][128] constants of the folded layers' border cases (biases) */
    const float* bc8;
    const float* mt_up6;          /* 2D: tap moments as nrrt_coords_poly_gemm operands: [512][4 * 1 * 512], */
    const float* mt_c6;           /*     [256][9 * 4 * 512], */
    const float* mt_c7;           /*     [128][9 * 4 * 256], */
    const float* mt_c8;           /*     [64][9 * 4 * 128] */
    const float* w6taps;          /* 2D: conv6[0] over the ConvTranspose channels, fp32 [512][9 * 512] (nrrt_poly16) */
    const float* b6;              /* 2D: conv6[0] bias */
    const float* mt3_c6;          /* 3D: [2 * 256][54 * 4 * 512], [2 * 128][54 * 4 * 256], [2 * 64][54 * 4 * 128] */
    const float* mt3_c7;
    const float* mt3_c8;
    const float* head_w;          /* final_conv: fp32 [64] */
    float head_bias;
} NrrtCnnWeights;

/* Workspace (device bytes) of nrrt_conv_forward_{2d,3d} for maps of (d, h, w) cells (d = 1 in 2D) and up to n_tasks tasks
 * per call.  It does not grow with the number of maps: the encoder runs 8 maps at a time. */
size_t nrrt_conv_forward_workspace_bytes(const NrrtCnnWeights* w, int d, int h, int w_, int n_tasks);

/* logits[t] = model(map task_to_row[t], coords[t]) for t < n_tasks (fp32 [n_tasks][d * h * w], raw network output).
 * occ_maps: device uint8 [*][cells], the task files' arrays (2D: (h, w); 3D: [x][y][z] with x = h, y = w, z = d sizes,
 *   presented to the network last axis first like the reference's dataset does, see nrrt_stem_conv_maps);
 * map_ids: HOST int32 [n_maps], rows of occ_maps of the distinct maps of this call;
 * task_to_row: HOST int32 [n_tasks], index into map_ids of every task, NON-DECREASING (tasks grouped by map, as the
 *   reference's task files are: generate_tasks.py:18-24);  both host arrays are copied on `stream` before use;
 * coords: device fp32 [n_tasks][2][2] (2D: [[sx, sy], [fx, fy]], train.py:57-60) / [n_tasks][2][3];
 * tasks_per_map: scheduling hint, tasks per map when it is the same for every map (10 in the reference), else 0.
 * Spatial sizes must be multiples of 8.  Everything is enqueued on `stream`; nothing is allocated. */
int nrrt_conv_forward_2d(const NrrtCnnWeights* w, const uint8_t* occ_maps, const int32_t* map_ids, int n_maps,
                         const int32_t* task_to_row, const float* coords, int n_tasks, int h, int w_, int tasks_per_map,
                         float* logits, void* workspace, size_t ws_bytes, void* stream);
int nrrt_conv_forward_3d(const NrrtCnnWeights* w, const uint8_t* occ_maps, const int32_t* map_ids, int n_maps,
                         const int32_t* task_to_row, const float* coords, int n_tasks, int d, int h, int w_, int tasks_per_map,
                         float* logits, void* workspace, size_t ws_bytes, void* stream);

/* First layer (Cin = 3 in 2D, 1 in 3D; fp32 NCHW / NCDHW input as the reference feeds it) -> fp16 channels-last with
 * `cout_pad` channels (cout real, the rest zero), + bias, ReLU.  weight fp32 [cout][cin][taps], bias fp32 [cout]. */
int nrrt_stem_conv(int dims, const float* x, int n, int cin, int d, int h, int w, const float* weight, const float* bias,
                   int cout, int cout_pad, void* out, void* stream);

/* Same layer fed straight from the resident uint8 occupancy maps: sample b reads map task_to_map[b] and every input
 * channel is the dataset's {0,1} normalisation of it, x = (occ != 0) (train.py:41-44, ThreeD_train.py:46-50), so the
 * (B,3,H,W) fp32 tensor the reference materialises per task never exists.  3D: a map is the task file's array
 * [x][y][z] (sizes h, w, d) and the network sees it as the dataset's ToTensorV2 presents it, last axis first:
 * voxel (d, h, w) of the network input = map[h][w][d] (ThreeD_train.py:59-66; the output is consumed un-transposed,
 * ThreeD_test_network_and_algorithms.py:98-105). */
int nrrt_stem_conv_maps(int dims, const uint8_t* occ_maps, const int32_t* task_to_map, int n, int cin, int d, int h, int w,
                        const float* weight, const float* bias, int cout, int cout_pad, void* out, void* stream);

/* Coordinate branch (train.py:179-190): the four conv_coords_* layers (3x3 Conv2d + ReLU) on the (B,1,2,2) / (B,1,2,3)
 * coords tensor in fp32.  weights: 4 pointers, fp32 [tap = ky*3+kx][cin][cout] (transposed from PyTorch's [cout][cin][3][3]); biases: 4 pointers.  feats: 4 output pointers,
 * fp32 [B][gh*gw][C] with C = 64, 128, 256, 512. */
int nrrt_coords_branch(const float* coords, int B, int gh, int gw, const float* const* weights, const float* const* biases,
                       float* const* feats, void* stream);

/* F.interpolate(feat, size, mode='bilinear'/'trilinear', align_corners=True) of a (gh x gw [x 1]) coordinate feature,
 * written as fp16 into channels [out_coffset, out_coffset + C) of a channels-last buffer (the torch.cat target). */
int nrrt_coords_fill(const float* feat, int B, int gh, int gw, int C, int dims, int od, int oh, int ow, void* out,
                     int64_t out_cstride, int64_t out_coffset, void* stream);

/* Coordinate channels folded into an epilogue polynomial (2D).  F.interpolate(..., align_corners=True) of the 2x2
 * coordinate feature (train.py:180,183,186,189) is bilinear in the pixel position, so a convolution over those channels
 * is a bilinear function of the output position with per-task coefficients E (derivation in csrc/conv_aux.cu).
 * feat: fp32 [B][4][Cc] (nrrt_coords_branch output); moments: fp32 [cases][n_uv][Cc][Cout] tap moments of the
 * coordinate slice of the consumer's weight (n_uv = 4 for 3x3 convs with cases = 9 border cases; n_uv = 1 for
 * ConvTranspose with cases = 4 output offsets); a = 1/(H-1), b = 1/(W-1) of the consumer's input grid;
 * E: fp32 [B][cases][4][Cout]. */
int nrrt_coords_poly(const float* feat, int B, int Cc, const float* moments, int cases, int n_uv, int Cout, float a, float b,
                     float* E, void* stream);

/* nrrt_coords_poly as one fp32 GEMM (the production path for large batches): moments_t = the same moments laid out
 * [Cc][cases][n_uv][Cout]; work: fp32 scratch [B*4*Cc + B*4*cases*n_uv*Cout]. */
int nrrt_coords_poly_gemm(const float* feat, int B, int Cc, const float* moments_t, int cases, int n_uv, int Cout, float a,
                          float b, float* work, float* E, void* stream);

/* 3D coordinate channels as a post-pass polynomial (derivation in csrc/conv_aux.cu; ThreeD_train.py:180-191).
 * nrrt_coords_pieces_3d: feat fp32 [B][6][Cc] (nrrt_coords_branch output on the (2,3) grid) -> A fp32 [B][4][2*Cc], the
 *   bilinear coefficients of the two h-pieces stacked along channels; nrrt_coords_poly_gemm(A, B, 2*Cc, moments_t, cases = 54,
 *   n_uv = 4, ...) then yields E fp32 [B][54][4][Cout] (case = (d-border * 6 + h-class) * 3 + w-border).
 * nrrt_poly_relu_3d: x[b] = fp16(relu(x[b] + E[b](d, h, w))) in place on a channels-last (D, H, W, C) fp16 buffer. */
int nrrt_coords_pieces_3d(const float* feat, int B, int Cc, float* A, void* stream);
int nrrt_poly_relu_3d(void* x, const float* E, int B, int D, int H, int W, int C, void* stream);

/* Epilogue polynomial of a 3x3 conv (nrrt_coords_poly output E [B][9][4][Cout] on the (2H, 2W) grid) re-expressed per
 * output parity in the tile space of the NRRT_CONV_UPFOLD layer, the (H, W) grid: Ef [B][4][9][4][Cout].  cases_const
 * (optional, fp32 [9][Cout]) is added to the constant term of border case k: the ConvTranspose bias (train.py:148,154)
 * seen through the conv taps that fall inside the image. */
int nrrt_poly_fold(const float* E, const float* cases_const, int B, int Cout, int H, int W, float* Ef, void* stream);

/* conv6.0 hoisted out of the per-task path (2D; train.py:192-195).  The layer is linear in its input and only the
 * coordinate channels depend on the task, so its pre-activation = a per-map tensor (nrrt_conv_layer once per map) + a
 * per-task piecewise-bilinear field with 16 cases (4 row classes x 4 column classes: first / interior even / interior
 * odd / last) -- derivation in csrc/conv_aux.cu.
 * nrrt_poly16: e_up fp32 [B][4][4][Cm] = nrrt_coords_poly output of the ConvTranspose (upconv6) whose result feeds the
 *   conv; w_taps fp32 [Cm][9][Cout] = the conv's weight over the ConvTranspose channels (tap = ky*3+kx); e_conv fp32
 *   [B][9][4][Cout] = nrrt_coords_poly output for the conv's own coordinate channels; bias fp32 [Cout];
 *   (up_h, up_w) = ConvTranspose input grid; work: fp32 scratch [B*16*9*Cout]; P16: fp32 [B][16][4][Cout].
 * nrrt_poly_relu: out[b] = fp16(relu(pre[pre_index[b]] + P16[b](pixel))) on an (H, W) grid, channels-last, C channels. */
int nrrt_poly16(const float* e_up, const float* w_taps, const float* e_conv, const float* bias, int B, int Cm, int Cout,
                int up_h, int up_w, float* work, float* P16, void* stream);
int nrrt_poly_relu(const void* pre, const int32_t* pre_index, const float* P16, void* out, int B, int H, int W, int C,
                   void* stream);

/* Encoder reuse: the encoder (conv1 .. conv5) sees only the occupancy map, and the benchmark has 10 tasks per map, so it
 * runs once per distinct map; this copies channels [src_coffset, +C) of sample index[b] of a channels-last fp16 buffer
 * into channels [dst_coffset, +C) of sample b of another (the task's torch.cat buffer, train.py:181,184,187,190). */
int nrrt_gather_channels(const void* src, int64_t src_cstride, int64_t src_coffset, void* dst, int64_t dst_cstride,
                         int64_t dst_coffset, int C, int64_t pixels, const int32_t* index, int B, void* stream);

/* final_conv (1x1, C -> 1): fp16 channels-last in, fp32 logits out [B][pixels]. */
int nrrt_head_1x1(const void* in, int64_t pixels, int C, int64_t in_cstride, const float* weight, float bias, float* logits,
                  void* stream);

/* Batched exact shortest grid paths: generate_paths.astar (generate_paths.py:11-74) and ThreeD_generate_paths.astar
 * (ThreeD_generate_paths.py:14-97) for B tasks at once -- the "A*" column of the reference's results table
 * (test_network_and_algorithms.py:58-70,148-183) and its task filter (tasks without a path are deleted,
 * generate_paths.py:111-113).  2D: 8-connected moves, cost 1 / sqrt(2), a diagonal step only when both orthogonal neighbours
 * are free.  3D: 26-connected moves, cost 1 / sqrt(2) / sqrt(3); every move needs its three axis-projected cells free (the
 * reference's `dx != 255` tests are always true, ThreeD_generate_paths.py:55-72).  A* with the consistent Euclidean
 * heuristic returns a shortest path and every shortest path has the same number of steps of each kind, so the kernel
 * returns those integers per task, n_steps [B][3] = steps of cost 1, sqrt(2), sqrt(3) (-1 when there is no path, length
 * NaN; -2 and length = +inf when the search reached the cost limit of the label format -- 8191 in 2D, 1771 in 3D --
 * before settling the goal), and length = n1 + n2 sqrt(2) + n3 sqrt(3): the reference's path length up to the order of
 * its additions.
 * occ_bits as nrrt_pack_occupancy writes them (1 = free; 3D maps hold {0 = free, 255}); starts / goals [B][dims] in array-axis
 * order ((y, x) in 2D, (x, y, z) in 3D, as the planners take them); (s0, s1, s2) = array shape (s2 ignored in 2D);
 * axis_only = 1 restricts the moves to the axis steps: reachability by 4- / 6-connected free cells (the task filter);
 * work: [n_slots][cells] uint64 scratch (one slot per resident CTA, e.g. the SM count); queue: one zeroed int32. */
int nrrt_grid_paths(int dims, const uint32_t* occ_bits, int words_per_map, const int32_t* task_to_map, const int32_t* starts,
                    const int32_t* goals, int B, int s0, int s1, int s2, int axis_only, void* work, int n_slots, int32_t* queue,
                    int32_t* n_steps, double* length, void* stream);

/* Same search, also returning ONE shortest path per task (what astar() returns to generate_paths(), the input of the mask
 * makers): paths int32 [B][path_cap][dims] from start to goal in array-axis order, path_n [B] = points on it (0 = no path,
 * -(points) = longer than path_cap).  The path is read off the final labels (predecessor = lowest legal neighbour whose step
 * counts are one move smaller); among equal-cost paths the reference's A* returns whichever its heap order meets first, so
 * length and step kinds agree with it, the individual cells may not. */
int nrrt_grid_paths_ex(int dims, const uint32_t* occ_bits, int words_per_map, const int32_t* task_to_map, const int32_t* starts,
                       const int32_t* goals, int B, int s0, int s1, int s2, int axis_only, void* work, int n_slots, int32_t* queue,
                       int32_t* n_steps, double* length, int32_t* paths, int path_cap, int32_t* path_n, void* stream);

/* ------------------------------------------------------------------------------------------------------------------
 * On-device workload generation (csrc/generate.cu): the reference's generators with CPython's `random` (MT19937,
 * random.seed(int), randint) reproduced bit for bit -- seed s here gives the map / the tasks `random.seed(s)` gives there.
 * ------------------------------------------------------------------------------------------------------------------ */
/* generate_maps.create_map (generate_maps.py:7-44; 2D: (s0, s1) = (height, width), 255 = free, 0 = obstacle / frame) and
 * ThreeD_generate_maps.create_map (ThreeD_generate_maps.py:5-41; (s0, s1, s2) = array axes x, y, z; 255 = obstacle).
 * seeds: device int64 [n_maps], one random.seed() per map;  occ_maps: device uint8 [n_maps][cells]. */
int nrrt_generate_maps(int dims, const int64_t* seeds, int n_maps, int s0, int s1, int s2, uint8_t* occ_maps, void* stream);

/* draw_start_and_finish (generate_tasks.py:27-46, ThreeD_generate_tasks.py:25-48) called n_cand times in a row per map
 * after random.seed(seeds[m]): candidate (start, goal) pairs in the planners' axis order ((y, x) / (x, y, z)),
 * [n_maps][n_cand][dims].  n_done[m] = n_cand, or -1 if the map has no admissible cells (the reference would spin). */
int nrrt_generate_task_candidates(int dims, const int64_t* seeds, int n_maps, int s0, int s1, int s2, const uint8_t* occ_maps,
                                  int n_cand, int32_t* starts, int32_t* goals, int32_t* n_done, void* stream);

/* The reference pipeline deletes tasks its A* cannot solve (generate_paths.py:111-113), which does not touch the random
 * stream: keep, per map and in candidate order, the first `per_map` candidates whose reach_steps[q][0] >= 0
 * (nrrt_grid_paths output for the candidates; NULL = keep all).  Writes the task arrays of the first n_tasks tasks
 * (the last map may be truncated): task_to_map, starts, goals [n_tasks][dims], coords fp32 [n_tasks][2][dims] (the CNN
 * input, (x, y[, z]) order: train.py:57-60), *redrawn = candidates skipped, *short_maps = maps that ran out of
 * candidates (call again with a larger n_cand). */
int nrrt_select_tasks(int dims, int n_maps, int n_cand, int per_map, int n_tasks, const int32_t* cand_starts,
                      const int32_t* cand_goals, const int32_t* reach_steps, int32_t* task_to_map, int32_t* starts,
                      int32_t* goals, float* coords, int32_t* redrawn, int32_t* short_maps, void* stream);

/* ------------------------------------------------------------------------------------------------------------------
 * Training-mask makers (csrc/masks.cu): the label images the reference's data pipeline writes next to the
 * task files and MapsDataset reads back as `mask` (train.py:47-52, ThreeD_train.py:52-57).
 * ------------------------------------------------------------------------------------------------------------------ */
/* generate_paths.visualize_path (generate_paths.py:118-169), the saved `path_brightness` image of B tasks:
 * filled circles of `radius` (10) at every path point (cv2.circle, :129-131), cv2.GaussianBlur with a (ksize, ksize) =
 * (33, 33) window and sigmaX = sigma = 3 (:134; 8-bit fixed-point arithmetic, BORDER_REFLECT_101), path pixels set to 255
 * (:136-137), everything zeroed where the map is not above free_above = 240 (:148,155).  Bit-identical to OpenCV 4.x.
 * occ_maps: uint8 [*][h][w] (255 = free); paths: int32 [B][path_cap][2] (y, x) as astar() returns them
 * (generate_paths.py:11-74); path_n: int32 [B]; masks: uint8 [B][h][w]. */
int nrrt_path_masks_2d(const uint8_t* occ_maps, const int32_t* task_to_map, const int32_t* paths, const int32_t* path_n,
                       int path_cap, int B, int h, int w, int radius, int ksize, double sigma, int free_above, uint8_t* masks,
                       void* stream);
/* Host-only: the two tables the kernel above is built on -- half_width[0..radius] of cv2.circle's filled disk per |dy| and
 * the ksize fixed-point (8.8) Gaussian taps.  No device work. */
int nrrt_mask_tables_host(int radius, int ksize, double sigma, int32_t* half_width, int32_t* taps);

/* ThreeD_generate_paths.save_and_visualize_path:142-150: vols[b] = zeros with 255 on the voxels of path b.
 * paths: int32 [B][path_cap][3] (x, y, z) = array axes; vols: uint8 [B][s0][s1][s2]. */
int nrrt_paths_to_volume(const int32_t* paths, const int32_t* path_n, int path_cap, int B, int s0, int s1, int s2, uint8_t* vols,
                         void* stream);
/* ThreeD_enlarge_path.enlarge_the_path (ThreeD_enlarge_path.py:25-65) for B path volumes: every voxel that is free
 * (occ != 255) and at Chebyshev distance d < layers from a voxel with path == 255 becomes max(path, values[d]);
 * values_host: HOST uint8 [layers], the reference's `path_value` per shell as its uint8 array stores it (206, 157, 108, 59,
 * 10 for LAYERS = 5).  out must not alias path_vols.  layers <= 6. */
int nrrt_enlarge_paths_3d(const uint8_t* occ_maps, const int32_t* task_to_map, const uint8_t* path_vols, int B, int s0, int s1,
                          int s2, int layers, const uint8_t* values_host, uint8_t* out, void* stream);

/* ------------------------------------------------------------------------------------------------------------------
 * Training step: UNet_cooler.training_step (train.py:205-210) = forward in train mode, nn.BCEWithLogitsLoss
 * (:113), loss.backward(), torch.optim.Adam(lr=1e-3) (:228-229).  The forward convolutions and the data gradients are
 * nrrt_conv_layer calls (a data gradient is the same convolution with the weights of nrrt_pack_dgrad_weight); the weight
 * gradients are nrrt_conv_wgrad_nhwc (csrc/wgrad_tc.cu, tcgen05); the rest are one-pass kernels of csrc/train_ops.cu.
 * Activations and activation gradients are fp16 channels-last; parameters, their gradients and the Adam state are fp32.
 * Gradients are computed for loss * grad_scale_inverse (the caller picks the scale: nrrt_bce_with_logits' grad_scale) and
 * unscaled by nrrt_adam_step.  The host graph is mgr_sampling_cnn_b200/trainer.py.
 * ------------------------------------------------------------------------------------------------------------------ */
/* Channels-last slice [n][h][w][in_cstride] (channels [0, c) at `in`) -> channel-major fp16, optionally masked by the ReLU
 * output it belongs to (value kept where relu_mask > 0; NULL = no mask):
 *   mode 0: out [copies][c][n][h][w]; copies = 1, or 3 for the X operand of a 3x3 weight gradient: copy k holds the image
 *   shifted by k - 1 pixels along w, out[k][..][x] = in[..][x + k - 1] (zero outside), because a TMA tile row cannot start
 *   at an odd pixel;  mode 1: out [c][n][2h][2w], value at the even positions, zeros elsewhere (the gradient of a
 *   stride-2 convolution seen on its input grid);  mode 2: out [4][c][n][h/2][w/2], group (y & 1) * 2 + (x & 1) (the
 *   gradient of a ConvTranspose k2 s2 seen on its input grid). */
int nrrt_to_channel_major(const void* in, int64_t in_cstride, const void* relu_mask, int64_t mask_cstride, int n, int h, int w, int c,
                          int mode, int copies, void* out, void* stream);
/* dw[co][tap][ci] += sum over pixels dy[co][pixel] * x[ci][pixel + tap] for a ksize x ksize (1 or 3), stride 1, pad ksize / 2
 * convolution; dy_cm [cout][n][h][w] and x_cm [ksize][cin][n][h][w] (the shifted copies) channel-major fp16
 * (nrrt_to_channel_major); dw fp32
 * [cout][ksize * ksize][ci_stride] (taps in (ky, kx) order; the caller zeroes it).  w must be a multiple of 64. */
int nrrt_conv_wgrad(const void* dy_cm, int cout, const void* x_cm, int cin, int n, int h, int w, int ksize, float* dw, int ci_stride,
                    void* stream);
/* The same weight gradient read straight from the channels-last buffers (no nrrt_to_channel_major): dy [n][h][w][dy_cstride]
 * (channels [0, cout) at `dy`, already masked by the layer's ReLU), x [n][h][w][x_cstride] (channels [0, cin) at `x`).  The
 * tiles are MN-major tensor-core operands, so both tap shifts are TMA coordinates.  Any w (rows are walked in 64-pixel blocks; a
 * partial block is zero-filled by TMA). */
int nrrt_conv_wgrad_nhwc(const void* dy, int64_t dy_cstride, int cout, const void* x, int64_t x_cstride, int cin, int n, int h, int w,
                         int ksize, float* dw, int ci_stride, void* stream);
/* g = dy where y > 0 else 0 (y = the stored output of a conv + bias + ReLU layer; NULL = no mask), dbias[c] += sum g
 * (dbias may be NULL; g may be NULL or alias dy). */
int nrrt_relu_bwd_sum(const void* dy, int64_t dy_cs, const void* y, int64_t y_cs, void* g, int64_t g_cs, float* dbias, int64_t pixels,
                      int c, void* stream);
/* nn.BatchNorm2d in train mode over a raw conv output z: batch statistics (biased variance), running statistics updated
 * with `momentum` (unbiased variance), y = [relu](gamma * (z - mean) * rstd + beta [+ res]).  save_mean / save_rstd: fp32 [c]
 * for the backward; scale_shift: fp32 [2c] scratch; work: fp64 [2c] scratch. */
int nrrt_bn_train_forward(const void* z, int64_t z_cs, int64_t pixels, int c, const float* gamma, const float* beta, float eps,
                          float momentum, float* running_mean, float* running_var, const void* res, int64_t res_cs, int relu, void* y,
                          int64_t y_cs, float* save_mean, float* save_rstd, float* scale_shift, double* work, void* stream);
/* The same layer in eval mode (validation_step under Lightning's model.eval(), train.py:212-217): running statistics, nothing
 * updated.  scale_shift: fp32 [2c] scratch. */
int nrrt_bn_eval_forward(const void* z, int64_t z_cs, int64_t pixels, int c, const float* gamma, const float* beta, float eps,
                         const float* running_mean, const float* running_var, const void* res, int64_t res_cs, int relu, void* y,
                         int64_t y_cs, float* scale_shift, void* stream);
/* Its backward: g = dy masked by y_mask > 0 (NULL = none); dbeta += sum g; dgamma += sum g * xhat;
 * dz = gamma * rstd * (g - mean(g) - xhat * mean(g * xhat)); g_out (optional) = g, the gradient of the residual branch. */
int nrrt_bn_train_backward(const void* dy, int64_t dy_cs, const void* y_mask, int64_t y_cs, const void* z, int64_t z_cs, int64_t pixels,
                           int c, const float* gamma, const float* save_mean, const float* save_rstd, void* dz, int64_t dz_cs,
                           void* g_out, int64_t g_cs, float* dgamma, float* dbeta, double* work, void* stream);
/* Channels-last pixel remaps: mode 0 = space to depth, in [n][2h][2w][c] -> out [n][h][w][4c] (channel (a * 2 + b) * c + k =
 * pixel (2y + a, 2x + b)); mode 1 = zero-stuffing, in [n][h][w][c] -> out [n][2h][2w][c] dense. */
int nrrt_remap_channels_last(const void* in, int64_t in_cs, void* out, int64_t out_cs, int n, int h, int w, int c, int mode, void* stream);
/* nn.BCEWithLogitsLoss: *loss_sum = sum of the per-element losses (divide by count for the mean, train.py:113,208);
 * dlogits (optional) = (sigmoid(z) - t) * grad_scale (or * scale_state[0] when scale_state, the device-resident dynamic loss
 * scale below, is given).  logit_bias (optional, one device float: final_conv's bias, which
 * lives in the device-resident parameter buffer) is added to logits in place first. */
int nrrt_bce_with_logits(float* logits, const float* logit_bias, const float* target, int64_t count, float grad_scale,
                         const float* scale_state, float* dlogits, double* loss_sum, void* stream);
/* MapsDataset's mask normalisation (train.py:47-50): target = (mask - min) / (max - min) per sample, float64 then float32. */
int nrrt_mask_targets(const uint8_t* masks, int B, int64_t cells, float* target, void* stream);
/* final_conv (1x1, c -> 1) backward: g = x > 0 ? dlogits * weight[k] : 0 (x = the ReLU output feeding the head);
 * dweight[k] += sum dlogits * x; *dbias += sum dlogits. */
int nrrt_head_backward(const float* dlogits, const void* x, int64_t x_cs, const float* weight, int64_t pixels, int c, void* g, int64_t g_cs,
                       float* dweight, float* dbias, void* stream);
/* torch.optim.Adam (no weight decay / amsgrad) on a flat fp32 buffer; the gradient is multiplied by grad_scale first.
 * Dynamic loss scaling without host synchronisation (activation gradients are fp16): scale_state = 4 device floats
 * [scale carried by the gradients, good steps in a row, overflow flag, optimizer steps taken].  nrrt_grad_check raises the
 * flag when a gradient is not finite; nrrt_adam_step with scale_state skips the step when the flag is up, divides the
 * gradients by scale_state[0] (on top of grad_scale) and takes its bias corrections from scale_state[3] + 1 (`step` is
 * then ignored); nrrt_loss_scale_update halves the scale after an overflow (not below min_scale), else counts the step and
 * doubles the scale after growth_interval good steps (not above max_scale), and clears the flag. */
int nrrt_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t count, int step, float lr, float beta1,
                   float beta2, float eps, float grad_scale, const float* scale_state, void* stream);
int nrrt_grad_check(const float* grad, int64_t count, float* scale_state, void* stream);
int nrrt_loss_scale_update(float* scale_state, int growth_interval, float max_scale, float min_scale, void* stream);
int nrrt_to_half(const float* src, void* dst, int64_t count, void* stream);
/* Data-gradient weights: dst fp16 [ci_rows][taps][co_cols], dst[ci][taps - 1 - t][co] = w[co][t][ci] (w fp32
 * [cout][taps][ci_stride]); rows / columns beyond the source are zero. */
int nrrt_pack_dgrad_weight(const float* w, int cout, int taps, int ci_stride, int ci_rows, int co_cols, void* dst, void* stream);
/* nrrt_pack_dgrad_weight for all layers of a model in one launch: descs_dev = device int64 [n_layers][8] =
 * {source offset into params (floats), destination offset into dst_base (halfs), cout, taps, ci_stride, ci_rows, co_cols, 0}. */
int nrrt_pack_dgrad_weights(const float* params, void* dst_base, const int64_t* descs_dev, int n_layers, void* stream);
/* First conv (cin -> cout from the fp32 NCHW input, train.py:118): dweight[co][ci][tap] += sum g[pixel][co] x[ci][pixel + tap];
 * dbias[co] += sum g.  g: fp16 channels-last, already masked by the ReLU. */
int nrrt_stem_wgrad(const float* x, const void* g, int64_t g_cs, int n, int cin, int h, int w, int cout, float* dweight, float* dbias,
                    void* stream);
/* Coordinate branch backward (train.py:179-190).  nrrt_coords_fill_backward: dfeat[b][q][k] += sum over pixels of the
 * bilinear weight of corner q times dcat[b][pixel][k] (the transpose of nrrt_coords_fill on a 2x2 source).
 * nrrt_coords_layer_backward: one conv_coords_* layer (3x3 + ReLU on the 2x2 grid): g = dfeat_out masked by f_out > 0;
 * dweight [tap][cin][cout] += ..., dbias += ..., dfeat_in (optional) += W^T g. */
int nrrt_coords_fill_backward(const void* dcat, int64_t cs, int B, int h, int w, int c, float* dfeat, void* stream);
int nrrt_coords_layer_backward(const float* dfeat_out, const float* f_out, const float* f_in, const float* weight, int B, int cin, int cout,
                               float* dweight, float* dbias, float* dfeat_in, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* NRRT_H_ */
