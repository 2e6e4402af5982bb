/*
 * ts2d.h -- C ABI of the B200-native 2D triangle-splatting rasterizer (libts2d.so).
 *
 * This is the drop-in boundary for the hot path BASELINE.json:north_star names.  It replaces the
 * reference's host orchestration + kernels behind its pybind module
 *     R2D = submodules/diff-triangle-rasterization-2D
 *     R2D/ext.cpp:4-9                        rasterize_triangles / rasterize_triangles_backward
 *     R2D/src/extension_interface.cu:19-260  (tensor checks, output allocation)
 *     R2D/src/rasterizer.cu:101-358          Rasterizer::forward / Rasterizer::backward
 *     R2D/src/param_struct.h:127-194         CameraInfo / GeometryInfo / ForwardOutput / BackwardInput / LossInput / BackwardOutput
 * with plain pointers and sizes: no torch (or any C++) types cross this boundary.  Every pointer
 * is a DEVICE pointer unless the name ends in _host.  All arrays are fp32 / int32 / uint8, dense,
 * in the layouts of the reference's tensors (extension_interface.cu:99-128, 229-233).
 *
 * Memory is caller-owned.  The three opaque "state" blobs mirror the reference's geometryBuffer /
 * binningBuffer / imageBuffer tensors (param_struct.h:44-123): the caller asks for their sizes,
 * allocates them (the Python host side uses torch uint8 tensors so that autograd keeps them alive
 * exactly like the reference, __init__.py:100), and passes them to forward and again to backward.
 * Their internal layout is private to this library (ts2d_export_* decode them for parity tests).
 *
 * The size of the binning state depends on the number of (triangle, tile) instances R (`num_rendered`), which is only known
 * after the per-triangle preprocess -- the reference blocks on a device->host copy of R at that point and resizes its
 * binningBuffer tensor (rasterizer.cu:189-195).  Two ways to drive the forward pass:
 *     ts2d_forward()           -> ONE enqueue, no host synchronisation: the caller sizes the binning state for a capacity it
 *                                 expects to suffice (e.g. R of the previous frame plus a margin); every kernel takes R from device
 *                                 memory.  R comes back asynchronously in ts2d_frame_counters; if it exceeds the capacity the frame is
 *                                 invalid (nothing is written out of bounds) and the caller repeats ts2d_forward_render() with a
 *                                 larger state -- the geometry state of the first attempt stays valid.
 *     ts2d_forward_geometry()  -> K1 preprocess + SH colour, depth sort, scan; returns R on the host (one synchronisation)
 *     ts2d_forward_render()    -> key emission, tile binning, tile ranges, front-to-back composite
 * Backward is one call:
 *     ts2d_backward()          -> reverse-walk composite gradients + preprocess/SH backward
 * The capacity of a binning state is a function of its size in bytes (ts2d_binning_capacity()); forward and backward both
 * derive the array layout from that, so the instance count itself never has to cross the ABI.
 *
 * Error convention: every entry point returns 0 on success; a negative TS2D_E_* code for argument
 * errors (the cases where the reference raises via AT_ERROR, extension_interface.cu:53-81); or a
 * positive cudaError_t.  ts2d_error_string() explains either.  When `debug` is non-zero every
 * launch is followed by a stream synchronisation + error check (reference: CHECK_CUDA,
 * auxiliary.h:358-367).  All work is enqueued on `stream` (a cudaStream_t passed as void*); the
 * only host synchronisation is the one inside ts2d_forward_geometry() that returns R (ts2d_forward() has none).
 *
 * Gradients are bit-reproducible: the fast composite backward writes no atomics (every (sub-tile, list entry) pair owns a 64 B
 * row of the scratch, rows are summed per triangle in a fixed order), unlike the reference's fp32 atomicAdd accumulation
 * (backward.cu:412,482-490).
 *
 * Multi-GPU (image-space tile sharding, no counterpart in the reference): `shard_rank`/`shard_world`
 * restrict key emission and compositing to the tiles with tile_id % shard_world == shard_rank.
 * shard_world == 1 is the single-GPU path.  Collectives live above this ABI (torch.distributed/NCCL).
 */
#ifndef TS2D_H_
#define TS2D_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TS2D_ABI_VERSION 7
#define TS2D_TILE 16          /* R2D/src/config.h:4-5  BLOCK_X = BLOCK_Y = 16 */
#define TS2D_MAX_CHANNELS 3   /* R2D/src/config.h:3 */

enum {
    TS2D_OK = 0,
    TS2D_E_BAD_VERTEX = -1,    /* extension_interface.cu:53-56 */
    TS2D_E_BAD_FEATURE = -2,   /* :57-60 */
    TS2D_E_BAD_SHS = -3,       /* :61-64 */
    TS2D_E_CHANNELS = -4,      /* :65-68  C > MAX_CHANNELS */
    TS2D_E_BACKGROUND = -5,    /* :69-72 */
    TS2D_E_GAMMA = -6,         /* :73-76 */
    TS2D_E_NULL = -7,          /* required pointer missing */
    TS2D_E_STATE_SIZE = -8,    /* state blob smaller than ts2d_*_state_bytes() */
    TS2D_E_SH_DEGREE = -9,     /* (sh_degree+1)^2 > M or sh_degree > 3 */
    TS2D_E_SHARD = -10,        /* shard_rank/shard_world invalid */
    TS2D_E_SIZE = -11,         /* image or primitive count out of range */
    TS2D_E_PRIMITIVE = -12,    /* flags.primitive is neither TS2D_PRIMITIVE_2D nor TS2D_PRIMITIVE_3D */
    TS2D_E_MODEL = -13,        /* geometry.model / backward_out.model inconsistent (missing pointer, use_shs == 0, M == 0) */
    TS2D_E_FABRIC = -14        /* ts2d_exchange_*: bad rank / world / operation, or a misaligned base, first or count */
};

/* R2D/src/param_struct.h:127-137 (CameraInfo).  Matrices are the 16 floats of the (contiguous)
 * torch tensors the reference receives: element [i] = row i/4, col i%4 of world_view_transform /
 * full_proj_transform, i.e. column-major w.r.t. the mathematical W2C matrix (auxiliary.h:40-58). */
typedef struct ts2d_camera {
    int32_t width, height;
    float tan_fovx, tan_fovy;
    const float *viewmatrix;   /* [16] */
    const float *projmatrix;   /* [16] */
    const float *campos;       /* [3]  */
} ts2d_camera;

/* R2D/src/param_struct.h:139-155 (GeometryInfo) + the three flags of Rasterizer::forward. */
typedef struct ts2d_geometry {
    int32_t P;                 /* triangles */
    int32_t sh_degree;         /* D: active SH degree 0..3 */
    int32_t M;                 /* stored SH coefficients per triangle (shs is [P][M][3]) */
    int32_t C;                 /* channels: 3 in SH mode, feature.size(1) otherwise (<= 3) */
    int32_t use_shs;           /* 1: colour from SH (shs), 0: `feature` [P][C] given */
    float gamma;               /* compactness exponent: power = -0.5 * ecc^(2*gamma) */
    float scale_modifier;      /* carried for API parity; unused by the reference kernels too */
    float background_depth;
    const float *background;   /* [C] */
    const float *vertex;       /* [P][3][3] */
    const float *shs;          /* [P][M][3] or NULL */
    const float *feature;      /* [P][C]   or NULL */
    const float *opacity;      /* [P] (post-activation) */
    const struct ts2d_model_inputs *model;  /* NULL: the reference-shaped call above.  Non-NULL: parameter-space inputs, see below */
} ts2d_geometry;

/* Parameter-space inputs.  diff_recon's model runs a Python preamble in front of every
 * rasterizer call (src/diff_recon/models/VanillaTS_model.py:608-647): opacity = sigmoid(_opacity) (:84,:612), optional
 * straight-through binarisation (:620-621), shs = cat(_f_dc, _f_rest) (:80,:611), optional rescale of every triangle about
 * its centre (gamma_rescale, :615-618 + _rescale_triangles :431-447) and background_depth = max ||campos - vertex|| (:623,
 * followed by an implicit device->host read when the 0-dim tensor is converted to the `float` the pybind call takes).
 * With `model` set the per-triangle kernels read the raw parameters directly and do that arithmetic in registers -- same
 * operation order as the torch kernels the preamble launches -- so nothing is materialised and nothing is read back:
 *   geometry.vertex  = _vertex (raw), geometry.shs / geometry.opacity are ignored (use_shs must be 1, C == 3),
 *   backward writes dL/d_vertex, dL/d_f_dc, dL/d_f_rest, dL/d_opacity (the LOGIT) through ts2d_model_grads. */
typedef struct ts2d_model_inputs {
    const float *f_dc;           /* [P][1][3]    SH coefficient 0 */
    const float *f_rest;         /* [P][M-1][3]  SH coefficients 1..M-1; NULL iff M == 1 */
    const float *opacity_logit;  /* [P]          pre-activation */
    float ste_threshold;         /* >= 0: forward uses ((sigmoid > thr) - sigmoid) + sigmoid, gradient passes straight through; < 0: off */
    float rescale_ratio;         /* v' = (v - mean(v)) * ratio + mean(v); 1 = off */
    int32_t bg_depth_from_vertices; /* 1: background_depth = max over all vertices of ||campos - v|| (un-rescaled v), computed on
                                       the device and consumed by the composite kernels from device memory; 0: geometry.background_depth */
} ts2d_model_inputs;

/* Gradients w.r.t. the raw parameters (autograd through the preamble: CatBackward = split, SigmoidBackward, the linear
 * rescale map) + the optional training statistics of VanillaTS_model.py:347-363, updated in place
 * for the triangles with radii > 0 (`visible_mask`, :674).  Any of the six statistics pointers may be NULL. */
typedef struct ts2d_model_grads {
    float *dL_df_dc;           /* [P][1][3] */
    float *dL_df_rest;         /* [P][M-1][3] */
    float *gradient_accum;     /* [P]  += ||dL_dcenter2D||   (:358) */
    float *gradient_denom;     /* [P]  += 1                  (:359) */
    float *contrib_sum;        /* [P]  = max(., forward contrib_sum)  (:360) */
    float *contrib_max;        /* [P]  = max(., forward contrib_max)  (:361) */
    float *contrib_denom;      /* [P]  += 1                  (:362) */
    float *max_radii2D;        /* [P]  = max(., radii / radii_div)    (:363; radii // render_up_scale, :651) */
    const float *fwd_contrib_sum;  /* [P] forward outputs feeding :360-361 (required when contrib_sum / contrib_max are set) */
    const float *fwd_contrib_max;
    int32_t radii_div;         /* render_up_scale (>= 1) */
} ts2d_model_grads;

typedef struct ts2d_flags {
    int32_t back_culling;
    int32_t rich_info;
    int32_t debug;
    int32_t shard_rank;        /* tile ownership for multi-GPU; 0 */
    int32_t shard_world;       /* 1 = all tiles */
    int32_t exact;             /* 1: per-pair arithmetic mirrors the reference op-for-op (IEEE div, powf,
                                  expf); 0: fast path with exact re-evaluation inside the decision bands */
    int32_t primitive;         /* TS2D_PRIMITIVE_2D: screen-space triangles (R2D, the north-star path);
                                  TS2D_PRIMITIVE_3D: ray / triangle-plane intersection in view space, the
                                  reference's second rasterizer package R3D = submodules/diff-triangle-rasterization-3D
                                  (same pybind API: R3D/ext.cpp:4-9, R3D/src/extension_interface.h:7-62; kernels
                                  R3D/src/forward.cu:61-306, R3D/src/backward.cu:144-454).  Same entry points, same
                                  state blobs, same outputs; dL_dcenter2D is then the view-space xy of the summed
                                  vertex gradients (R3D/src/backward.cu:211-213). */
} ts2d_flags;

/* Multi-GPU exchange over NVLink peer memory (no counterpart in the reference: it has no distributed code).  The composite kernels
 * know nothing about other GPUs: a rank renders the tiles it owns (shard_rank / shard_world above) into ITS replica of a SYMMETRIC
 * buffer (same layout on every rank, mapped by every peer and behind one NVSwitch multicast alias; cuMem + cuMulticast, e.g. torch
 * symmetric memory), and two exchange kernels, launched behind them on the same stream, complete the frame and the sums everywhere:
 *   ts2d_exchange_tiles      the 64-byte rows of the owned tiles of `n_planes` planar [H][W] images: replica -> every replica
 *                            (multimem.st.v4; disjoint tiles, so all ranks end up with the same frame, bit for bit);
 *   ts2d_exchange_allreduce  this rank's slice [first, first + count) of an array: combined over all replicas inside the switch
 *                            (multimem.ld_reduce) and written back to all of them (multimem.st).  One rank computes each element and
 *                            the switch combines in a fixed order: same bits on every rank and in every frame.
 * The caller owns the protocol: rendezvous of the ranks after the local results are complete and before anybody publishes (which
 * also says that nobody still reads the previous frame), exchange kernels, rendezvous.  Two rendezvous per pass. */
#define TS2D_MAX_RANKS 8
#define TS2D_EXCHANGE_ADD_F32 0   /* fp32 sum; first, count multiples of 4, 16-byte aligned base */
#define TS2D_EXCHANGE_MAX_U32 1   /* u32 maximum (contrib_max: non-negative floats, bit order == value order) */

#define TS2D_PRIMITIVE_2D 0
#define TS2D_PRIMITIVE_3D 1

/* R2D/src/param_struct.h:157-169 (ForwardOutput), minus the three state tensors. */
typedef struct ts2d_forward_out {
    float *out_feature;        /* [C][H][W] planar */
    int32_t *radii;            /* [P] */
    float *depth;              /* [H][W]      (rich_info) */
    float *normal;             /* [3][H][W]   (rich_info) */
    float *contrib_sum;        /* [P]         (rich_info) */
    float *contrib_max;        /* [P]         (rich_info) */
} ts2d_forward_out;

/* R2D/src/param_struct.h:180-185 (LossInput). dL_dout_depth / dL_dout_normal may be NULL when
 * rich_info == 0 (the reference's Python backward would raise there, __init__.py:114-117,141-142). */
typedef struct ts2d_loss_in {
    const float *dL_dout_feature;  /* [C][H][W] */
    const float *dL_dout_depth;    /* [H][W] */
    const float *dL_dout_normal;   /* [3][H][W] */
} ts2d_loss_in;

/* R2D/src/param_struct.h:187-194 (BackwardOutput).  Every element is written (no pre-zeroing
 * needed; the reference relies on torch::zeros, extension_interface.cu:229-233). */
typedef struct ts2d_backward_out {
    float *dL_dvertex;     /* [P][3][3] */
    float *dL_dcenter2D;   /* [P][2] */
    float *dL_dshs;        /* [P][M][3] (M may be 0) */
    float *dL_dfeature;    /* [P][C]   (SH mode: dL/d rgb, like the reference's scratch use) */
    float *dL_dopacity;    /* [P]  (model inputs: gradient w.r.t. the logit) */
    const ts2d_model_grads *model;  /* required iff geometry.model is set; dL_dshs is then ignored */
} ts2d_backward_out;

/* Device-side counters of a frame.  A ts2d_counters object owns pinned host memory for them plus the events that say when the
 * asynchronous copies have landed; pass one to ts2d_forward() / ts2d_forward_render() and ask it afterwards:
 *   ts2d_counters_num_rendered()   waits only for the scan (the rest of the frame is still queued on the GPU: no bubble),
 *   ts2d_counters_backward_rows()  waits for the end of that forward pass.
 * One object per forward pass in flight; not thread-safe.  ts2d_read_counters() is the blocking alternative (reads the geometry
 * state, synchronises the stream). */
typedef struct ts2d_counters ts2d_counters;
typedef struct ts2d_frame_counters {
    int64_t num_rendered;    /* R: (triangle, tile) instances of the frame (the reference's return value, rasterizer.cu:190-193).
                                R > ts2d_binning_capacity(binning_state_bytes): the frame is invalid, render again with a larger state */
    int64_t backward_rows;   /* (sub-tile, list entry) pairs the composite backward will write a 64 B row for: sizes its scratch
                                (ts2d_backward_scratch_bytes).  0 with the mirror kernels (flags.exact), which use the reference's atomics */
} ts2d_frame_counters;

int ts2d_counters_create(ts2d_counters **out);
void ts2d_counters_destroy(ts2d_counters *c);
int ts2d_counters_num_rendered(ts2d_counters *c, int64_t *num_rendered);
int ts2d_counters_backward_rows(ts2d_counters *c, int64_t *backward_rows);
int ts2d_read_counters(const void *geometry_state, int32_t P, ts2d_frame_counters *out_host, void *stream);

int ts2d_abi_version(void);
const char *ts2d_error_string(int code);

/* Sizes of the opaque state blobs (bytes).  cf. BaseDataBuffer::requiredSize, param_struct.h:36-40. */
size_t ts2d_geometry_state_bytes(int32_t P);
size_t ts2d_binning_state_bytes(int64_t capacity /* instances */, int32_t width, int32_t height);
int64_t ts2d_binning_capacity(size_t binning_state_bytes);  /* instances a binning state of that size holds (>= the capacity it was sized for) */
size_t ts2d_image_state_bytes(int32_t width, int32_t height);
/* Scratch for backward: per-triangle screen-space gradient accumulators (the reference's dL_dv*_2D, dL_dnormal_view, dL_dv_depth
 * temporaries, rasterizer.cu:289-300) at its START (16 floats per triangle), followed by the row storage of the atomics-free
 * write-back: `backward_rows` from ts2d_frame_counters (the mirror kernels need no rows: pass 0). */
size_t ts2d_backward_scratch_bytes(int32_t P, size_t binning_state_bytes, int64_t backward_rows);

/* Replaces rasterizer.cu:116-267 in ONE enqueue without host synchronisation (see the header comment).  Writes radii[P] and
 * every array of `out`. */
int ts2d_forward(const ts2d_camera *cam, const ts2d_geometry *geom, const ts2d_flags *flags, int32_t *radii, void *geometry_state,
                 size_t geometry_state_bytes, void *binning_state, size_t binning_state_bytes, void *image_state, size_t image_state_bytes,
                 const ts2d_forward_out *out, ts2d_counters *counters /* may be NULL */, void *stream);

/* Replaces rasterizer.cu:116-193: preprocess (forward.cu:61-193), then ordering + scan.
 * Writes radii[P]; returns num_rendered through *num_rendered_host (one stream synchronisation). */
int ts2d_forward_geometry(const ts2d_camera *cam, const ts2d_geometry *geom, const ts2d_flags *flags,
                          int32_t *radii, void *geometry_state, size_t geometry_state_bytes,
                          int64_t *num_rendered_host, void *stream);

/* Replaces rasterizer.cu:195-266: duplicateWithKeys, SortPairs, identifyTileRanges, FORWARD::renderCUDA.
 * num_rendered: what ts2d_forward_geometry() returned, or -1 when only the device knows it (then the grids are sized for the
 * capacity of the binning state).  `counters` may be NULL. */
int ts2d_forward_render(const ts2d_camera *cam, const ts2d_geometry *geom, const ts2d_flags *flags,
                        int64_t num_rendered, const void *geometry_state, void *binning_state, size_t binning_state_bytes,
                        void *image_state, size_t image_state_bytes, const ts2d_forward_out *out, ts2d_counters *counters,
                        void *stream);

/* Replaces rasterizer.cu:269-358: BACKWARD::renderCUDA + BACKWARD::preprocessCUDA.  The binning state is not const: the backward
 * pass narrows the sub-tile coverage bits of the instance keys to the pairs it visits (idempotent). */
int ts2d_backward(const ts2d_camera *cam, const ts2d_geometry *geom, const ts2d_flags *flags,
                  const int32_t *radii, const void *geometry_state, void *binning_state, size_t binning_state_bytes, const void *image_state,
                  const ts2d_loss_in *loss, const ts2d_backward_out *out, void *scratch, size_t scratch_bytes, void *stream);

/* ts2d_backward split in two for tile-sharded multi-GPU (no counterpart in the reference): each rank runs the composite
 * backward over its own tiles; the first 16 * P floats of `scratch` then hold its per-triangle partial sums (the reference's
 * dL_dv*_2D / dL_dnormal_view / dL_dv_depth / dL_drgb / dL_dopacity temporaries, rasterizer.cu:289-300, in this library's layout),
 * the caller sum-reduces those across ranks, then every rank runs the per-triangle stage on the sums.  `accumulators` != NULL puts
 * the 16 * P floats there instead (e.g. into a symmetric buffer for ts2d_exchange_allreduce); ts2d_backward_geometry reads them
 * from wherever its `scratch` argument points. */
int ts2d_backward_composite(const ts2d_camera *cam, const ts2d_geometry *geom, const ts2d_flags *flags,
                            const void *geometry_state, void *binning_state, size_t binning_state_bytes, const void *image_state,
                            const ts2d_loss_in *loss, void *scratch, size_t scratch_bytes, float *accumulators, void *stream);
int ts2d_backward_geometry(const ts2d_camera *cam, const ts2d_geometry *geom, const ts2d_flags *flags, const int32_t *radii,
                           const void *geometry_state, const ts2d_backward_out *out, const void *scratch, size_t scratch_bytes,
                           void *stream);

/* render_up_scale epilogue (VanillaTS_model.py:647-655): F.interpolate(x, size=(H/s, W/s), mode="bilinear") of `planes`
 * planar images rendered at s times the target resolution, align_corners = False, no antialiasing -- i.e. each output
 * pixel samples the source at s * (d + 0.5) - 0.5: the mean of the two middle source rows/columns for even s, the
 * middle one for odd s (same lerp order as torch's upsample_bilinear2d kernel).  `in` is [planes][H*s][W*s], `out` is
 * [planes][H][W].  ts2d_downsample_bwd is its adjoint (writes every element of dL_din). */
int ts2d_downsample(const float *in, float *out, int32_t planes, int32_t out_width, int32_t out_height, int32_t s, void *stream);
int ts2d_downsample_bwd(const float *dL_dout, float *dL_din, int32_t planes, int32_t out_width, int32_t out_height, int32_t s, void *stream);

/* Fused depth-normal consistency loss, the geometry term of the trainer (DepthNormalLoss, src/diff_recon/trainers/trainer_utils.py:203-257,
 * used at VanillaTS_trainer.py:84,111): mean((1 - <normalize(normal), normal_from_depth(depth)>) * mask), mask = pixels whose depth-gradient
 * norm (Scharr, :159-185) lies below its `quantile` over the frame (torch.quantile semantics, linear interpolation; 0.9 in the reference).
 * depth [H][W], normal [3][H][W].  half_resolution = 1 evaluates the depth surface on the bilinearly halved depth map and up-samples the
 * result (the reference's scale_factor = 0.5, what its shipped config uses); 0 = scale_factor None / 1.  forward() writes the loss to the
 * device scalar *loss and leaves what backward() needs in `scratch`; backward() writes dL/ddepth and / or dL/dnormal (either may be NULL:
 * depth_grad / normal_grad = False), scaled by the device scalar *grad_loss (NULL = 1).  No host synchronisation, no float atomics in the
 * gradients (bit-reproducible). */
size_t ts2d_depth_normal_loss_scratch_bytes(int32_t width, int32_t height, int32_t half_resolution);
int ts2d_depth_normal_loss_forward(const float *depth, const float *normal, int32_t width, int32_t height, float tan_fovx, float tan_fovy,
                                   int32_t half_resolution, float quantile, float *loss, void *scratch, size_t scratch_bytes, void *stream);
int ts2d_depth_normal_loss_backward(int32_t width, int32_t height, float tan_fovx, float tan_fovy, int32_t half_resolution, const float *grad_loss,
                                    void *scratch, size_t scratch_bytes, float *dL_ddepth, float *dL_dnormal, void *stream);

/* Multi-GPU exchange, see above.  `local` / `multicast`: this rank's replica and the multicast alias of the same symmetric buffer
 * ([n_planes][height][width] floats for the tiles; element index `first` of the 4-byte array for the all-reduce). */
int ts2d_exchange_tiles(const float *local, float *multicast, int32_t n_planes, int32_t width, int32_t height, int32_t rank, int32_t world,
                        void *stream);
int ts2d_exchange_allreduce(void *multicast, int64_t first, int64_t count, int32_t op, void *stream);

/* Fused image loss on the rendered frame (first step):
 *     loss = w_l1 * mean|image - gt| + w_ssim * (1 - mean SSIM(image, gt))
 * i.e. w_L1 * L1(image, gt) + w_ssim * SSIMLoss()(image, gt) of the trainer (src/diff_recon/trainers/VanillaTS_trainer.py:74-75,108;
 * trainer_utils.py:323-324 and :9-103: 11x11 Gaussian window, sigma 1.5, zero padding, C1 = 0.01^2, C2 = 0.03^2) in one forward and one
 * backward kernel instead of ~30 torch kernels and ten depth-wise convolutions.  `image`, `gt`: planar [channels][height][width].
 * forward writes loss[0] (and terms = {L1, 1 - SSIM} when non-NULL) and keeps three derivative maps in `scratch`
 * (ts2d_image_loss_scratch_bytes()); backward turns them into dL/d image, scaled by the device scalar *grad_loss (NULL = 1). */
size_t ts2d_image_loss_scratch_bytes(int32_t channels, int32_t width, int32_t height);
int ts2d_image_loss_forward(const float *image, const float *gt, int32_t channels, int32_t width, int32_t height, float w_l1, float w_ssim,
                            float *loss, float *terms, void *scratch, size_t scratch_bytes, void *stream);
int ts2d_image_loss_backward(const float *image, const float *gt, int32_t channels, int32_t width, int32_t height, float w_l1, float w_ssim,
                             const float *grad_loss, const void *scratch, size_t scratch_bytes, float *dL_dimage, void *stream);

/* ---- state decoding, for parity tests against the reference's buffers ----
 * Each writes arrays in the reference's own element types/order.  Any output pointer may be NULL. */
int ts2d_export_geometry(const void *geometry_state, int32_t P,
                         float *v2d /*[P][3][2]*/, float *area2 /*[P]*/, float *normal_view /*[P][3]*/, float *v_depth /*[P][3]*/,
                         float *depth /*[P]*/, float *rgb /*[P][3]*/, uint8_t *clamped /*[P][3]*/, uint32_t *tiles_touched /*[P]*/,
                         uint32_t *rect_min /*[P][2]*/, uint32_t *rect_max /*[P][2]*/, void *stream);
/* 3D primitive (R3D/src/param_struct.h:44-75): view-space vertices, UN-normalised plane normal. */
int ts2d_export_geometry3d(const void *geometry_state, int32_t P,
                           float *v_view /*[P][3][3]*/, float *normal_view /*[P][3]*/, float *depth /*[P]*/, float *rgb /*[P][3]*/,
                           uint8_t *clamped /*[P][3]*/, uint32_t *tiles_touched /*[P]*/, uint32_t *rect_min /*[P][2]*/,
                           uint32_t *rect_max /*[P][2]*/, void *stream);
/* model inputs: the activated (sigmoid / STE) opacity the kernels used, and the device-computed background depth */
int ts2d_export_model(const void *geometry_state, int32_t P, int32_t primitive, float *opacity /*[P]*/, float *background_depth /*[1]*/, void *stream);
int ts2d_export_binning(const void *geometry_state, const void *binning_state, size_t binning_state_bytes, const void *image_state, int32_t P,
                        int64_t num_rendered, int32_t width, int32_t height, uint64_t *keys_sorted /*[R]*/, uint32_t *point_list /*[R]*/,
                        uint32_t *ranges /*[tiles][2]*/, void *stream);
int ts2d_export_image(const void *image_state, int32_t width, int32_t height, uint32_t *n_contrib /*[H][W]*/, float *final_T /*[H][W]*/,
                      void *stream);

/* ---- per-stage device timing (measurement only; used by bench.py for the roofline object) ----
 * When enabled, every stage launched by the three entry points above is bracketed by CUDA events on
 * the caller's stream.  ts2d_profile_read() synchronises, adds up the elapsed time per stage over
 * all calls since the last read (ms) and the number of launches, then resets. */
enum {
    TS2D_STAGE_PREPROCESS = 0,     /* K1  preprocess + SH colour */
    TS2D_STAGE_ORDER_SCAN = 1,     /* K2/K3 depth sort of P keys + scan */
    TS2D_STAGE_BINNING = 2,        /* K4-K6 emit, tile radix sort, ranges */
    TS2D_STAGE_RENDER_FWD = 3,     /* K7 */
    TS2D_STAGE_RENDER_BWD = 4,     /* K8: composite backward */
    TS2D_STAGE_PREPROCESS_BWD = 5, /* K9 */
    TS2D_STAGE_BWD_PREPARE = 6,    /* row marking + scan in front of K8 (atomics-free write-back) */
    TS2D_STAGE_BWD_REDUCE = 7,     /* per-triangle row reduction behind K8 */
    TS2D_NUM_STAGES = 8
};
int ts2d_profile_enable(int enable);
int ts2d_profile_read(float *ms_out /*[TS2D_NUM_STAGES]*/, int32_t *launches_out /*[TS2D_NUM_STAGES]*/);

#ifdef __cplusplus
}
#endif
#endif /* TS2D_H_ */
