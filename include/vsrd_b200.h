/* vsrd_b200 — C ABI of the B200-native VSRD silhouette-renderer hot path.
 *
 * The reference (skmhrk1209/VSRD) is pure Python/PyTorch and has no FFI layer of its own
 * (SURVEY.md §8b).  Each entry point below replaces the named reference function; the Python
 * host layer (vsrd_b200/, vsrd/) binds them with ctypes and re-exposes the reference's own
 * call signatures on top.  INTEGRATION.md shows the binding.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer to contiguous fp32 (or int64 where stated) unless the
 *     parameter name starts with `host_`;
 *   - all calls are asynchronous on `stream` (a cudaStream_t passed as void*), perform no host
 *     synchronisation and are CUDA-graph capturable;
 *   - buffers are borrowed for the duration of the call; outputs are caller-allocated;
 *   - return value: 0 on success, non-zero on error; vsrd_last_error() returns a thread-local,
 *     human-readable message for the last failure on the calling thread;
 *   - layouts are RAY-MAJOR: distances[R][M+1], per-sample outputs [R][M][...];
 *     the reference's sample-major tensors ([M, R, ...]) are permuted views of these.
 *
 * Symbols: R = rays, S = num_samples, M = intervals per ray (S-1 in pass 1, 2S-1 in pass 2),
 *          N = instances, NW = 1617 residual-MLP weights per instance (48-16-16-16-16-1).
 */
#ifndef VSRD_B200_H_
#define VSRD_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VSRD_MLP_WEIGHTS 1617
#define VSRD_GRAD_STRIDE 1632   /* 1617 weights + 15 pose gradients (t, half extents, R), padded */
#define VSRD_MAX_INSTANCES 32
#define VSRD_MAX_INTERVALS 512

/* Device-resident per-step scalars.  scripts/main.py:420-431 re-derives the union temperature, the SDF
 * standard deviation and the cosine ratio on the host every step; keeping them (and the sampler seed)
 * in device memory lets one captured CUDA graph be replayed for every optimisation step.  Wherever a
 * call accepts a `step_state` DEVICE pointer, a non-NULL value overrides the by-value scalars. */
typedef struct VsrdStepState {
    float temperature;            /* sdf_union_temperature                                    */
    float std_deviation;          /* sdf_std_deviation                                        */
    float cosine_ratio;           /* step / num_steps                                         */
    float eikonal_weight;         /* loss weight, 0 while step < warmup_steps (main.py:677)   */
    uint64_t seed;                /* counter-based generator key of this step                 */
    int64_t step;
} VsrdStepState;

/* Host-side description of the schedule (config.json:226-238, optimization.num_steps/warmup_steps). */
typedef struct VsrdSchedule {
    int64_t num_steps;
    int64_t warmup_steps;
    float max_temperature, min_temperature;
    float max_std_deviation, min_std_deviation;
    float eikonal_weight;
    float _pad;
    uint64_t seed;
} VsrdSchedule;

/* Decoded per-frame parameters: what scripts/main.py:530-618 closes over when it composes
 * soft_union(translation(rotation(instance_field(box [+ residual])))). */
typedef struct VsrdScene {
    int32_t num_instances;        /* N <= VSRD_MAX_INSTANCES */
    int32_t _pad;
    const float* locations;       /* [N,3]   world_outputs.locations                         */
    const float* rotations;       /* [N,3,3] world_outputs.orientations, p = (x - t) @ R      */
    const float* half_extents;    /* [N,3]   world_outputs.dimensions                         */
    const float* mlp_weights;     /* [N,NW]  hyper_distance_field(embeddings) or NULL (warm-up)*/
    float temperature;            /* sdf_union_temperature (main.py:422-426)                  */
    float scale;                  /* max(distance_range) (main.py:441)                        */
    const VsrdStepState* step_state; /* DEVICE pointer or NULL: overrides temperature, std_deviation,
                                        cosine_ratio and the eikonal weight of the calls below   */
} VsrdScene;

typedef struct VsrdRays {
    int32_t num_rays;             /* R */
    int32_t num_intervals;        /* M <= VSRD_MAX_INTERVALS; distances hold M+1 entries per ray */
    const float* origins;         /* [R,3]   ray_positions                                    */
    const float* directions;      /* [R,3]   ray_directions                                   */
    const float* distances;       /* [R,M+1] sampled_distances, ascending                     */
    /* Optional instance culling.  The residual lies in (0,1), so an instance whose BOX SDF exceeds the lowest
     * box SDF + 1 by more than VSRD_CULL_LOG_EPS * temperature has a soft-min weight w below exp(-VSRD_CULL_LOG_EPS) = 2e-9
     * of the dominant one, and its union-gradient coefficient |w (1 - (d_i - d) / T)| stays below 20 exp(-20) = 4e-8 < 2^-24:
     * neither term can change the fp32 sums of the union it is added to.  vsrd_cull_samples() tests
     * that per (instance, sample), writes the box value and gradient of the pairs that pass into `field`, and lists the
     * others in forward_samples: a header of VSRD_CULL_HEADER_INTS int32 (the live-sample count of instance i at
     * [i * VSRD_CULL_COUNT_STRIDE], one cache line apart), then [N][R*M] flat sample indices (in no particular order;
     * every sample's result is independent of its neighbours in the list).  vsrd_field_forward then
     * evaluates the residual MLP on the listed samples only, an equal share of them per thread block.  NULL disables it.
     * cull_stats (DEVICE uint64[4], may be NULL) accumulates over all field launches {backward tiles skipped, backward
     * tiles visited (vsrd_backward_tile_rows() samples each), forward (sample, instance) pairs skipped, visited}. */
    int32_t* forward_samples;
    unsigned long long* cull_stats;
    /* Backward side of the culling: live_tiles, vsrd_live_tiles_bytes() bytes ([N][ceil(R*M / vsrd_backward_tile_rows())] marks,
     * then int32 counts of live samples per block of VSRD_CENSUS_BLOCK_TILES tiles), ZEROED by the caller
     * before vsrd_composite_backward, which (a) writes exact zeros for the adjoints of instances whose soft-min weight
     * is below exp(-VSRD_CULL_LOG_EPS) and (b) marks every warp tile of vsrd_field_backward that received a non-zero
     * adjoint.  vsrd_field_backward then visits only the marked tiles (compacted in order per thread block, so the
     * work is balanced and the accumulation order stays deterministic).  NULL disables it. */
    uint8_t* live_tiles;
} VsrdRays;
/* Bytes of VsrdRays::live_tiles for (N, R, M): the marks plus the per-block census behind them. */
size_t vsrd_live_tiles_bytes(int num_instances, int num_rays, int num_intervals);
#define VSRD_CULL_COUNT_STRIDE 32
#define VSRD_CENSUS_BLOCK_TILES 128
#define VSRD_CULL_HEADER_INTS (VSRD_MAX_INSTANCES * VSRD_CULL_COUNT_STRIDE)
#define VSRD_CULL_LOG_EPS 20.0f      /* x exp(-x) < 2^-24 = 5.96e-8 for x >= 20: value AND gradient terms of the soft-min */
/* Samples per warp tile of the backward field kernel (16 or 32), for sizing VsrdRays::live_tiles. */
int vsrd_backward_tile_rows(void);

/* Arguments of hierarchical_volumetric_rendering (vsrd/rendering/renderers.py:177-188). */
typedef struct VsrdRenderParams {
    float std_deviation;          /* sdf_std_deviation */
    float cosine_ratio;
    float epsilon;                /* 1e-6 */
    float _pad;
} VsrdRenderParams;

/* Optional in-kernel loss (scripts/main.py:653-687, 855): loss = sil_w * mean BCE(clamp(labels), targets)
 * + eik_w * mean((|grad| - 1)^2).  targets == NULL disables it. */
typedef struct VsrdLoss {
    const float* targets;         /* [R,N] soft-mask targets in label order, or NULL            */
    float silhouette_weight;
    float eikonal_weight;         /* 0 during warm-up                                           */
} VsrdLoss;

/* The V views of a target frame (target + sources, main.py:339-367). */
typedef struct VsrdViews {
    int32_t num_views;            /* V */
    int32_t target_view;          /* index of relative frame 0 (matching uses this view only)   */
    int32_t height, width;        /* image size for clip_boxes_to_image                         */
    const float* extrinsics;      /* [V,4,4] world -> camera                                    */
    const float* intrinsics;      /* [V,3,3]                                                    */
} VsrdViews;

int vsrd_version(void);
const char* vsrd_last_error(void);

/* Number of thread blocks per instance the backward field kernel uses for (N,R,M); the caller
 * provides `partials` with N * that * VSRD_GRAD_STRIDE floats. */
int vsrd_backward_blocks_per_instance(int num_instances, int num_rays, int num_intervals);

/* ---- a1: vsrd.rendering.ray_casting (vsrd/rendering/utils.py:5-18) ----------------------
 * inv_projection[V][9] = inv(E)[:3,:3] @ inv(K) (row-major); out directions[V][H][W][3]. */
int vsrd_ray_directions(const float* inv_projection, int num_views, int height, int width,
                        float* directions, void* stream);

/* Rays for `num_rays` flat pixel indices into [V,H,W] (scripts/main.py:632-633 gathers them from
 * the [V,H,W,3] tensors; here they are generated on the fly).  camera_positions[V][3]. */
int vsrd_gather_rays(const float* inv_projection, const float* camera_positions, const int64_t* pixel_indices,
                     int num_rays, int num_views, int height, int width,
                     float* origins, float* directions, void* stream);

/* ---- a9: quadrature_sampler (vsrd/rendering/samplers.py:5-8) ----------------------------
 * bins[S+1]; jitter[R][S] in [0,1) or NULL to draw from the counter-based generator (seed).
 * `step_state` (DEVICE pointer or NULL) supplies the seed instead of the by-value one.
 * out distances[R][S]. */
int vsrd_place_coarse(const float* bins, const float* jitter, uint64_t seed, const VsrdStepState* step_state,
                      int num_rays, int num_samples, float* distances, void* stream);

/* ---- a10: inverse_transform_sampler + concat + sort (samplers.py:11-36, renderers.py:196-210)
 * coarse_distances[R][S], coarse_weights[R][S-1]; sorted_uniforms[R][S] or NULL (generator);
 * out distances[R][2S], ascending. */
int vsrd_place_fine(const float* coarse_distances, const float* coarse_weights, const float* sorted_uniforms,
                    uint64_t seed, const VsrdStepState* step_state, int num_rays, int num_samples,
                    float* distances, void* stream);

/* The culling pre-pass described at VsrdRays::forward_samples, to be enqueued before vsrd_field_forward on the same
 * `field` buffer (reads origins / directions / distances and the box parameters of `scene`; rays->forward_samples itself
 * is ignored, the list goes to `forward_samples`: VSRD_CULL_HEADER_INTS + N * R * M int32). */
int vsrd_cull_samples(const VsrdScene* scene, const VsrdRays* rays, float* field, int32_t* forward_samples, void* stream);

/* ---- a5-a8: per-(sample, instance) field: box SDF + residual MLP, value and spatial gradient.
 * out field[N][R*M] as float4 (d_i, dd_i/dx, dd_i/dy, dd_i/dz). */
int vsrd_field_forward(const VsrdScene* scene, const VsrdRays* rays, float* field, void* stream);

/* ---- a8 (soft union) + a11 (renderers.py:218-263): union, SDF->opacity, front-to-back compositing.
 * out labels[R][N], gradients[R][M][3] (un-normalised union gradient), weights[R][M].
 * If loss->targets != NULL also accumulates loss_out[0] += BCE sum / (R*N) * w, loss_out[1] += eikonal
 * mean * w (loss_out must be zeroed by the caller); loss/loss_out may be NULL. */
int vsrd_composite_forward(const VsrdScene* scene, const VsrdRays* rays, const VsrdRenderParams* params,
                           const float* field, float* labels, float* gradients, float* weights,
                           const VsrdLoss* loss, float* loss_out, void* stream);

/* ---- backward of the above (replaces the autograd double-backward replay, a16).
 * Upstream gradients grad_labels[R][N], grad_gradients[R][M][3], grad_weights[R][M]; each may be
 * NULL.  If loss->targets != NULL the loss gradients are generated in-kernel from `labels`
 * (the forward output) instead and added to the explicit ones.
 * out adjoint[N][R*M] as float4: adjoints of (d_i, grad d_i). */
int vsrd_composite_backward(const VsrdScene* scene, const VsrdRays* rays, const VsrdRenderParams* params,
                            const float* field, const float* grad_labels, const float* grad_gradients,
                            const float* grad_weights, const VsrdLoss* loss, const float* labels,
                            float* adjoint, void* stream);

/* Reduces adjoint[N][R*M] into parameter gradients (recomputing the per-sample forward):
 * grad_locations[N][3], grad_rotations[N][9], grad_half_extents[N][3], grad_mlp_weights[N][NW]
 * (NULL when scene->mlp_weights is NULL).  `partials` is scratch, see
 * vsrd_backward_blocks_per_instance(). Outputs are overwritten (not accumulated). */
int vsrd_field_backward(const VsrdScene* scene, const VsrdRays* rays, const float* adjoint, float* partials,
                        float* grad_locations, float* grad_rotations, float* grad_half_extents,
                        float* grad_mlp_weights, void* stream);

/* EXPERIMENT, not called by anything the package ships: vsrd_field_backward's contract served by a
 * tcgen05 / TMEM kernel (residual instances only, R*M > 0).  Measured slower than the shipped kernel, and its MLP weight
 * gradients carry the rounding of bf16 operand staging (2e-3 relative); kept so that the measurement can be repeated
 * (tools/compare_backward.py). */
int vsrd_experimental_field_backward_tcgen05(const VsrdScene* scene, const VsrdRays* rays, const float* adjoint,
                                             float* partials, float* grad_locations, float* grad_rotations,
                                             float* grad_half_extents, float* grad_mlp_weights, void* stream);

/* ---- a14 + a15: multi-view box projection, matching and projection losses, forward and adjoint in
 * ONE launch (scripts/main.py:339-415; operations/geometric_operations.py:343-389 project_box_3d and
 * clip_lines_to_front; torchvision clip_boxes_to_image / distance_box_iou / distance_box_iou_loss;
 * scipy linear_sum_assignment; nn.functional.smooth_l1_loss).
 *   world_boxes [N,8,3]        detector corners (box_parameters.py:73-91)
 *   gt_boxes_2d [V,N,4]        x1 y1 x2 y2 per view in TARGET instance order, or NULL: projection only
 *   visible     [V,N] uint8    source-view visibility of each target instance (main.py:248-251), NULL = all
 *   fixed_gt_indices [N]       NULL: solve the assignment on -DIoU of the target view; else use these
 * out boxes_2d [V,N,4] (clipped to the image); gt_indices [N] int64 (ground-truth index matched to
 * prediction k; pd_indices are 0..N-1 as scipy returns them for a square cost); losses[2] =
 * (iou_projection_loss, l1_projection_loss) as means over the visible matched pairs;
 * grad_world_boxes [2,N,8,3] = d losses[k] / d world_boxes (NULL to skip the adjoint).
 * scratch: vsrd_projection_scratch_floats(V, N) floats. */
size_t vsrd_projection_scratch_floats(int num_views, int num_instances);
int vsrd_projection_step(const VsrdViews* views, int num_instances, const float* world_boxes,
                         const float* gt_boxes_2d, const uint8_t* visible, const int64_t* fixed_gt_indices,
                         float* boxes_2d, int64_t* gt_indices, float* losses, float* grad_world_boxes,
                         float* scratch, void* stream);

/* vsrd.operations.project_box_3d (vsrd/operations/geometric_operations.py:343-389; scripts/main.py:346-355 calls it once
 * per view and instance, 136 times per step) for `num_boxes` CAMERA-frame boxes [B,8,3] sharing one intrinsic matrix
 * [3,3]: the 12 edges of main.py's LINE_INDICES are clipped to z > 0 and projected; boxes_2d [B,4] = (min u, min v,
 * max u, max v), zeros when the box is entirely behind the camera.  No host synchronisation (the reference's
 * `torch.any(masks)` forces one per call).  The backward maps grad_boxes_2d [B,4] to grad_boxes_3d [B,8,3]. */
int vsrd_project_box_3d(const float* boxes_3d, int num_boxes, const float* intrinsic_matrix, float epsilon,
                        float* boxes_2d, void* stream);
int vsrd_project_box_3d_backward(const float* boxes_3d, int num_boxes, const float* intrinsic_matrix, float epsilon,
                                 const float* grad_boxes_2d, float* grad_boxes_3d, void* stream);

/* ---- a2: ray selection (scripts/main.py:620-627: torch.multinomial(max_n soft_masks, num_rays, no
 * replacement) over all V*H*W pixels, every step).  The weights do not change within a frame, so the
 * max over instances and its inclusive CDF (double) are built once per frame:
 *   soft_masks [P,N] (P = V*H*W, instance-minor as main.py:300-315 stacks them) -> cdf [P] double;
 *   scratch: vsrd_ray_cdf_scratch_doubles(P) doubles. */
size_t vsrd_ray_cdf_scratch_doubles(int64_t num_pixels);
int vsrd_ray_cdf_build(const float* soft_masks, int64_t num_pixels, int num_instances, double* cdf, double* scratch,
                       void* stream);
/* Per step: draws pixels i.i.d. from the CDF and rejects repeats in draw order, which IS sequential
 * sampling without replacement.  uniforms [max_draws] double in [0,1) or NULL (counter-based generator
 * keyed by `seed`, or by step_state->seed when step_state is a non-NULL DEVICE pointer).
 * out pixel_indices [R] int64 in acceptance order; status[0] (device int32, may be NULL) = number of
 * rays that could NOT be drawn (0 on success; >0 when the uniforms ran out or fewer than R pixels have
 * non-zero weight, where torch.multinomial raises). */
int vsrd_select_rays(const double* cdf, int64_t num_pixels, const double* uniforms, int max_draws, uint64_t seed,
                     const VsrdStepState* step_state, int num_rays, int64_t* pixel_indices, int32_t* status, void* stream);
/* targets[r][k] = soft_masks[pixel_indices[r]][gt_indices[k]] (main.py:656; gt_indices NULL = identity). */
int vsrd_gather_targets(const float* soft_masks, const int64_t* pixel_indices, const int64_t* gt_indices,
                        int num_rays, int num_instances, float* targets, void* stream);

/* ---- soft masks of the synthetic frames (transforms/geometric_transforms.py:267-309 SoftRasterizer):
 * sigmoid(signed pixel distance to the instance polygon / temperature); polygons [V,N,max_vertices,2]
 * (x = column, y = row), polygon_sizes [V,N] int32 (< 3: instance absent, mask 0);
 * out soft_masks [V,H,W,N]. */
int vsrd_soft_masks(const float* polygons, const int32_t* polygon_sizes, int num_views, int num_instances,
                    int max_vertices, int height, int width, float temperature, float* soft_masks, void* stream);

/* ---- inference / logging renderers (scripts/main.py:1011-1041): the per-instance field at ARBITRARY points,
 * the soft union there, and one iteration of vsrd.rendering.sphere_tracing (rendering/renderers.py:21-76).
 * surface_normal (renderers.py:79-113) is the normalised union gradient these return.
 *   points [P,3] -> field [N][P] float4 (d_i, grad d_i): the same kernels as vsrd_field_forward. */
int vsrd_field_points(const VsrdScene* scene, const float* points, int num_points, float* field, void* stream);
/* soft union (scripts/main.py:477-492) of field [N][P]: union_out [P] float4 (d, grad d) and/or
 * weights [P,N] (softmin weights = soft instance labels); either output may be NULL. */
int vsrd_union_points(const VsrdScene* scene, const float* field, int num_points, float* union_out, float* weights,
                      void* stream);
/* One sphere-tracing iteration (renderers.py:45-55) given union_out [P] float4 evaluated at `positions`:
 *   positions += directions * d where foreground & ~converged;  foreground &= |positions| < bounding_radius
 *   (bounding_radius <= 0: no bound);  converged = |d| < convergence_criteria.
 * directions [P,3] (directions_per_ray != 0) or one shared [3]; foreground / converged [P] uint8 in/out.
 * active [>= iteration + 1] int32, zero-initialised: active[iteration] receives the number of rays still
 * foreground and not converged; an iteration whose predecessor counted 0 is a no-op, which reproduces the
 * reference's global `break` (renderers.py:55) without a host round trip per iteration. */
int vsrd_sphere_trace_step(const float* union_out, const float* directions, int directions_per_ray, int num_rays,
                           float convergence_criteria, float bounding_radius, float* positions, uint8_t* foreground,
                           uint8_t* converged, int32_t* active, int iteration, void* stream);

/* ---- a3 / a4 / a16: the per-frame models and their optimiser (scripts/main.py:174-199, 330-332, 527, 859-865).
 * In the reference these are nn.Modules stepped by autograd + torch.optim.Adam: ~150 launches of 2-5 us per
 * optimisation step on [N,.] tensors.  The entry points below run them as a dozen latency-bound launches. */
#define VSRD_HYPER_WIDTH 256        /* hyper_in_channels == hyper_out_channels_list[i] (config.json:142-162) */
#define VSRD_HYPER_MAX_LAYERS 5
#define VSRD_MAX_PARAM_GROUPS 8

/* One weight-normed Linear of HyperDistanceField.hypernetwork (models/fields/hyper_distance_field.py:30-55)
 * and the LayerNorm that follows it (NULL for the output layer).  State-dict names in parentheses. */
typedef struct VsrdHyperLayer {
    const float* weight_v;        /* [out,in]  (hypernetwork.L.0.weight_v)                       */
    const float* weight_g;        /* [out]     (hypernetwork.L.0.weight_g, stored [out,1])       */
    const float* bias;            /* [out]     (hypernetwork.L.0.bias)                           */
    const float* ln_weight;       /* [out]     (hypernetwork.L.1.weight) or NULL                 */
    const float* ln_bias;         /* [out]     (hypernetwork.L.1.bias) or NULL                   */
    int32_t in_features, out_features;
} VsrdHyperLayer;
typedef struct VsrdHyperNet {
    int32_t num_layers;           /* 5 in configs/kitti_360: 4 hidden + the 1617-wide output layer */
    int32_t _pad;
    VsrdHyperLayer layers[VSRD_HYPER_MAX_LAYERS];
} VsrdHyperNet;
typedef struct VsrdHyperLayerGrads {
    float* weight_v; float* weight_g; float* bias; float* ln_weight; float* ln_bias;
} VsrdHyperLayerGrads;
typedef struct VsrdHyperNetGrads {
    int32_t num_layers;
    int32_t _pad;
    VsrdHyperLayerGrads layers[VSRD_HYPER_MAX_LAYERS];
} VsrdHyperNetGrads;

/* HyperDistanceField.forward (hyper_distance_field.py:75-77): embeddings [N,256] -> mlp_weights [N,out_last].
 * activations [num_layers-1][N][256] receives the pre-LayerNorm outputs of the hidden Linears (kept for the
 * backward). */
int vsrd_hyper_forward(const VsrdHyperNet* net, const float* embeddings, int num_instances,
                       float* activations, float* mlp_weights, void* stream);
/* Its backward (what autograd replays for main.py:859): grad_mlp_weights [N,out_last] -> every gradient in `grads`
 * and grad_embeddings [N,256] (all overwritten).  scratch: vsrd_hyper_scratch_floats(N) floats. */
size_t vsrd_hyper_scratch_floats(int num_instances);
int vsrd_hyper_backward(const VsrdHyperNet* net, const VsrdHyperNetGrads* grads, const float* embeddings,
                        int num_instances, const float* activations, const float* grad_mlp_weights,
                        float* grad_embeddings, float* scratch, void* stream);

/* BoxParameters3D buffers `location_range` / `dimension_range` (box_parameters.py:19-30). */
typedef struct VsrdBoxRanges {
    float location_min[3], location_max[3];
    float dimension_min[3], dimension_max[3];
} VsrdBoxRanges;
/* BoxParameters3D.forward (box_parameters.py:60-91, 124-146): raw locations [N,3], dimensions [N,3],
 * orientations [N,2] -> locations [N,3] = lerp(range, sigmoid(raw)), half_extents [N,3] likewise,
 * rotations [N,3,3] = rotation_matrix_y(normalize(raw)), boxes_3d [N,8,3] corners. */
int vsrd_decode_boxes(const VsrdBoxRanges* ranges, const float* raw_locations, const float* raw_dimensions,
                      const float* raw_orientations, int num_instances, float* locations, float* half_extents,
                      float* rotations, float* boxes_3d, void* stream);
/* Its backward.  Inputs: gradients w.r.t. the decoded locations / half extents / rotations (from
 * vsrd_field_backward) and, optionally, grad_boxes_3d [2,N,8,3] from vsrd_projection_step weighted by
 * (iou_weight, l1_weight) (config.json:120-127).  If `losses` is non-NULL it also records
 * losses[5] = (total, silhouette, eikonal, iou, l1) from render_loss_parts[2] (weighted, as
 * vsrd_composite_forward accumulates them) and projection_losses[2] (unweighted; NULL = none). */
int vsrd_decode_boxes_backward(const VsrdBoxRanges* ranges, const float* raw_locations, const float* raw_dimensions,
                               const float* raw_orientations, int num_instances, const float* half_extents,
                               const float* rotations, const float* grad_locations, const float* grad_half_extents,
                               const float* grad_rotations, const float* grad_boxes_3d, float iou_weight, float l1_weight,
                               float* grad_raw_locations, float* grad_raw_dimensions, float* grad_raw_orientations,
                               const float* render_loss_parts, const float* projection_losses, float* losses, void* stream);

/* torch.optim.Adam (amsgrad off, no weight decay) + ExponentialLR (config.json:177-215; main.py:863-865) over
 * ONE flat arena of `numel` floats holding every parameter, group after group.  Group k covers
 * [group_end[k-1], group_end[k]); its learning rate at optimisation step s is base_lr[k] * exp(log_gamma * s);
 * its own Adam step count is s - first_step[k] + 1 (the reference's Adam skips parameters without a gradient,
 * so the hypernetwork groups start counting after the warm-up); groups with s < first_step[k] are untouched.
 * s = step_state->step when step_state is a non-NULL DEVICE pointer, else `step`. */
typedef struct VsrdAdamGroups {
    int32_t num_groups;
    int32_t _pad;
    int64_t group_end[VSRD_MAX_PARAM_GROUPS];
    int64_t first_step[VSRD_MAX_PARAM_GROUPS];
    float base_lr[VSRD_MAX_PARAM_GROUPS];
    float beta1, beta2, eps, _pad2;
    double log_gamma;
} VsrdAdamGroups;
int vsrd_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t numel,
                   const VsrdAdamGroups* groups, const VsrdStepState* step_state, int64_t step, void* stream);

/* ---- schedule (scripts/main.py:420-431, 677): set_step >= 0 jumps to that step, < 0 advances by one;
 * recomputes temperature / std_deviation (cosine annealing), cosine_ratio, the eikonal switch and the
 * per-step seed in DEVICE memory. */
int vsrd_step_state_update(VsrdStepState* step_state, const VsrdSchedule* schedule, int64_t set_step, void* stream);

#ifdef __cplusplus
}
#endif
#endif  /* VSRD_B200_H_ */
