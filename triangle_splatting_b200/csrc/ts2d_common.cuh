// ts2d_common.cuh -- shared device-side definitions for libts2d (sm_100a only).
//
// State layout in HBM (all private to this library, decoded only by ts2d_export_*):
//
//   geometry state (per triangle, P entries; written by k_preprocess)
//     rec0   float4[3P]   48 B "raster record": {v1.x v1.y v2.x v2.y} {v3.x v3.y 1/area2 opacity} {r g b area2}
//     rec1   float4[2P]   32 B rich record    : {n.x n.y n.z vd1} {vd2 vd3 0 0}        (rich_info only)
//     dkey   u32[P]       fp32 bit pattern of view depth, 0xFFFFFFFF for culled triangles
//     dkey2  u32[P]       sorted depth keys;  ids2 u32[P]  triangle ids in depth-rank order  (tmpk / ids: ping-pong buffers of the sort)
//     tiles  u32[P]       number of (owned) tiles in the triangle's rect
//     rect   ushort4[P]   {min.x, min.y, max.x, max.y} in tiles
//     offs   u32[P]       inclusive scan of tiles[] in depth-rank order
//     binrec uint4[P]     {rect.x | rect.y << 16, rect.z | rect.w << 16, first emission index (offs[rank] - tiles), tiles}, by triangle
//                         id: what the backward's row marking needs about the triangle of an instance, in ONE 16-byte gather
//     clamp  u8[P]        SH clamp mask (bit c set <=> channel c clamped at 0)
//     hdr    GeomHeader   R (num_rendered) and the other device-side counters of a frame
//     status / sstatus    look-back words of the depth sort's passes / of the scan (ts2d_sort.cuh)
//   binning state (per instance; sized for a CAPACITY >= R chosen by the caller: R itself may only be known on the device)
//     tkey[0]/tval[0] u32[cap]  (tile id << 8 | sub-tile mask) / triangle id of each instance in emission (depth-rank) order
//     tkey[1]/tval[1] u32[cap]  the same after the stable tile sort; tval[1] is the per-tile list
//     status                    look-back words of the tile sort's passes
//   image state
//     ranges  uint2[tiles]  [start,end) of each tile in the sorted list
//     n_contrib u32[H*W], final_T f32[H*W]
//     lastw   u32[tiles*8]  per 8x4 sub-tile: max n_contrib of its pixels (where the backward starts; written by the fast K7)
//
// The raster record replaces the reference's nine SoA arrays (R2D/src/param_struct.h:44-57) so the
// composite kernels stage one contiguous 48 B (+32 B) record per list entry instead of 9 indirect loads.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include "../../include/ts2d.h"

#define TS2D_BLOCK 256
#define TS2D_DEFAULT_CTA_WARPS 1
#define TS2D_EPS 1e-8f  // R2D/src/auxiliary.h:8
#define TS2D_MASK_BITS 8  // low bits of an instance key: coverage of the tile's eight 8x4 sub-tiles (ts2d_fast.cuh)

struct GeomHeader {
    int64_t num_rendered;  // R: written by the scan (ts2d_sort.cuh: k_scan_gather); may exceed the binning capacity (then the frame is invalid)
    uint32_t bg_bits;  // model inputs: fp32 bit pattern of max ||campos - v|| (norms are >= 0, so unsigned order == float order)
    uint32_t pad;
    uint32_t tickets[8];      // block tickets of the depth sort's passes and of the scan
    uint32_t hist[4][256];    // digit histograms of the depth sort's passes
    struct Render {           // what a render (binning + composite) on this geometry state uses: cleared at its start, so it can be repeated
        int64_t bwd_rows;     // number of (sub-tile, list entry) rows the fast composite backward will produce (counted by the fast K7)
        uint32_t tickets[8];  // tile sort passes 0..2, backward row scan
        uint32_t hist[3][256];  // digit histograms of the tile sort's passes
    } render;
};
enum { TS2D_TICKET_DEPTH0 = 0, TS2D_TICKET_SCAN = 4, TS2D_TICKET_TILE0 = 0, TS2D_TICKET_BWD_SCAN = 4 };  // depth/scan index GeomHeader::tickets, tile/backward index Render::tickets

// Device-side view of ts2d_model_inputs (include/ts2d.h); `on == 0` is the reference-shaped call.
struct ModelIn {
    const float *f_dc, *f_rest, *logit;
    float ste_thr, ratio;
    uint32_t *bg_bits;  // non-NULL: K1 maxes ||campos - v|| into it
    int on;
};
// Device-side view of ts2d_model_grads.
struct ModelOut {
    float *g_dc, *g_rest;
    float *grad_accum, *grad_denom, *csum, *cmax, *cdenom, *max_radii;
    const float *fwd_csum, *fwd_cmax;
    int radii_div;
};

struct GeomState {
    float4 *rec0;
    float4 *rec1;
    uint32_t *dkey, *dkey2, *ids, *ids2, *tmpk;
    uint32_t *tiles;
    ushort4 *rect;
    uint32_t *offs;
    uint4 *binrec;
    unsigned long long *csum64;  // contrib_sum in 2^-32 fixed point (fast forward kernels)
    uint8_t *clamp;
    GeomHeader *hdr;
    unsigned long long *status;   // [rs_tiles(P)][256]
    unsigned long long *sstatus;  // [sc_tiles(P)]
    size_t status_bytes, sstatus_bytes;
};

struct BinState {
    int64_t cap;  // instances the arrays hold
    uint32_t *tkey[2];
    uint32_t *tval[2];
    unsigned long long *status;  // [rs_tiles(cap)][256]
    size_t status_bytes;
};

struct ImageState {
    uint2 *ranges;
    uint32_t *n_contrib;
    float *final_T;
    uint32_t *lastw;
};

// Scratch of the backward pass (ts2d_backward_scratch_bytes): the per-triangle accumulators K9 reads, and -- fast kernels only --
// the atomics-free write-back of the composite backward: every (sub-tile, list entry) pair the forward blended owns one 64 B
// row; rows are laid out in EMISSION order (all rows of a triangle are consecutive), so the per-triangle sums are plain
// sequential reductions in a fixed order (ts2d_bwd_reduce.cu).
struct BwdScratch {
    int64_t cap;          // instance capacity (== BinState.cap of the frame)
    float *gacc;          // [16 P]
    uint32_t *ei;         // [cap]   emission index of sorted list position
    uint8_t *cnt;         // [cap]   rows of emission index e (popcount of its live sub-tile bits)
    uint32_t *sbase;      // [cap + 1] exclusive scan of cnt
    unsigned long long *sstatus;  // scan look-back words
    size_t sstatus_bytes;
    float4 *rows;         // [rows_cap][4]
    int64_t rows_cap;
};

// ---- small vector helpers (own naming; semantics are plain component-wise fp32) ----
struct f2 { float x, y; };
struct f3 { float x, y, z; };
__device__ __forceinline__ f2 mk2(float x, float y) { f2 r; r.x = x; r.y = y; return r; }
__device__ __forceinline__ f3 mk3(float x, float y, float z) { f3 r; r.x = x; r.y = y; r.z = z; return r; }
__device__ __forceinline__ f2 operator+(f2 a, f2 b) { return mk2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ f2 operator-(f2 a, f2 b) { return mk2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ f2 operator*(f2 a, f2 b) { return mk2(a.x * b.x, a.y * b.y); }
__device__ __forceinline__ f2 operator*(f2 a, float s) { return mk2(a.x * s, a.y * s); }
__device__ __forceinline__ f2 operator*(float s, f2 a) { return mk2(s * a.x, s * a.y); }
__device__ __forceinline__ f2 operator+(f2 a, float s) { return mk2(a.x + s, a.y + s); }
__device__ __forceinline__ f3 operator+(f3 a, f3 b) { return mk3(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ f3 operator-(f3 a, f3 b) { return mk3(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ f3 operator-(f3 a) { return mk3(-a.x, -a.y, -a.z); }
__device__ __forceinline__ f3 operator*(f3 a, float s) { return mk3(a.x * s, a.y * s, a.z * s); }
__device__ __forceinline__ f3 operator*(float s, f3 a) { return mk3(s * a.x, s * a.y, s * a.z); }
__device__ __forceinline__ f3 operator/(f3 a, float s) { return mk3(a.x / s, a.y / s, a.z / s); }
__device__ __forceinline__ f3 operator+(f3 a, float s) { return mk3(a.x + s, a.y + s, a.z + s); }
__device__ __forceinline__ float dot3(f3 a, f3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ float len3(f3 a) { return sqrtf(dot3(a, a)); }
__device__ __forceinline__ float len2(f2 a) { return sqrtf(a.x * a.x + a.y * a.y); }
__device__ __forceinline__ float cross2(f2 a, f2 b) { return a.x * b.y - a.y * b.x; }
__device__ __forceinline__ f3 cross3(f3 a, f3 b) { return mk3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
__device__ __forceinline__ f2 perp2(f2 a) { return mk2(a.y, -a.x); }

// Affine maps with the 16-float matrices of ts2d_camera (element [4*c + r]); the summation order
// (x-term + y-term + z-term + translation) is the one the reference's parity depends on
// (R2D/src/auxiliary.h:40-95).
__device__ __forceinline__ f3 xf_point(const float *m, f3 p)
{
    return mk3(m[0] * p.x + m[4] * p.y + m[8] * p.z + m[12], m[1] * p.x + m[5] * p.y + m[9] * p.z + m[13],
               m[2] * p.x + m[6] * p.y + m[10] * p.z + m[14]);
}
__device__ __forceinline__ f3 xf_vec(const float *m, f3 p)
{
    return mk3(m[0] * p.x + m[4] * p.y + m[8] * p.z, m[1] * p.x + m[5] * p.y + m[9] * p.z, m[2] * p.x + m[6] * p.y + m[10] * p.z);
}
__device__ __forceinline__ f3 xf_vec_T(const float *m, f3 p)
{
    return mk3(m[0] * p.x + m[1] * p.y + m[2] * p.z, m[4] * p.x + m[5] * p.y + m[6] * p.z, m[8] * p.x + m[9] * p.y + m[10] * p.z);
}
__device__ __forceinline__ float4 xf_hom(const float *m, f3 p)
{
    return make_float4(m[0] * p.x + m[4] * p.y + m[8] * p.z + m[12], m[1] * p.x + m[5] * p.y + m[9] * p.z + m[13],
                       m[2] * p.x + m[6] * p.y + m[10] * p.z + m[14], m[3] * p.x + m[7] * p.y + m[11] * p.z + m[15]);
}
__device__ __forceinline__ f3 project_center(const float *m, f3 p)
{
    float4 h = xf_hom(m, p);
    float winv = 1.0f / (fabsf(h.w) + TS2D_EPS);
    return mk3(h.x * winv, h.y * winv, h.z * winv);
}
// First-order projection of a view-space offset d at view-space point p (R2D/src/auxiliary.h:97-118).
__device__ __forceinline__ f2 project_offset(f3 p, f3 d, float tfx, float tfy)
{
    return mk2((d.x - d.z * p.x / p.z) / (p.z * tfx), (d.y - d.z * p.y / p.z) / (p.z * tfy));
}
// The reference's one fp64 hop (R2D/src/auxiliary.h:35-38); kept so centre pixels round identically.
__device__ __forceinline__ float ndc_to_pix(float v, int S) { return ((v + 1.0) * S - 1.0) * 0.5; }

// ------------------------------------------------------------------------------------------------
// Per-(pixel, triangle) evaluation.
//
// eval_exact() is the op-for-op mirror of the reference's per-pair arithmetic
// (R2D/src/forward.cu:299-314 == backward.cu:382-401) with the FMA contraction the reference's
// sm_100 build has (checked against its SASS: t = pv2.y*pv3.x; fma(pv2.x, pv3.y, -t); IEEE divide;
// ecc = fma(-3, min3, 1); libdevice powf / expf).  Written with explicit round-to-nearest
// intrinsics so that no surrounding code can change the contraction.  All skip decisions
// (ecc range, alpha < 1/255) and the value of alpha are therefore bit-identical to the reference.
// ------------------------------------------------------------------------------------------------
struct PairEval {
    float a1, a2, a3, ecc, power, G, alpha;
    float pv1x, pv1y, pv2x, pv2y, pv3x, pv3y;
};

__device__ __forceinline__ bool eval_exact(float v1x, float v1y, float v2x, float v2y, float v3x, float v3y, float area2, float op,
                                           float two_gamma, float px, float py, PairEval &e)
{
    e.pv1x = __fsub_rn(v1x, px); e.pv1y = __fsub_rn(v1y, py);
    e.pv2x = __fsub_rn(v2x, px); e.pv2y = __fsub_rn(v2y, py);
    e.pv3x = __fsub_rn(v3x, px); e.pv3y = __fsub_rn(v3y, py);
    const float c1 = __fmaf_rn(e.pv2x, e.pv3y, -__fmul_rn(e.pv2y, e.pv3x));
    const float c2 = __fmaf_rn(e.pv3x, e.pv1y, -__fmul_rn(e.pv3y, e.pv1x));
    e.a1 = __fdiv_rn(c1, area2);
    e.a2 = __fdiv_rn(c2, area2);
    e.a3 = __fsub_rn(__fsub_rn(1.0f, e.a1), e.a2);
    e.ecc = __fmaf_rn(fminf(fminf(e.a1, e.a2), e.a3), -3.0f, 1.0f);
    if (e.ecc < 0.0f || e.ecc > 10.0f) return false;
    e.power = __fmul_rn(-0.5f, powf(e.ecc, two_gamma));
    e.G = expf(e.power);
    e.alpha = fminf(0.99f, __fmul_rn(op, e.G));
    return !(e.alpha < 1.0f / 255.0f);
}

// Per-triangle gradient accumulator written by the composite backward and consumed by K9:
// 16 floats = one 64 B line per triangle.
//   [0..5] dL/d(v1.xy, v2.xy, v3.xy)   [6] dL/d opacity   [7] dL/d n.x
//   [8..10] dL/d rgb                   [11] dL/d n.y      [12] dL/d n.z   [13..15] dL/d v_depth
#define GACC_STRIDE 16

// Warp helpers
__device__ __forceinline__ float warp_sum(float v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

#define TS2D_CUDA_TRY(expr)                    \
    do {                                       \
        cudaError_t _e = (expr);               \
        if (_e != cudaSuccess) return (int)_e; \
    } while (0)

// ---- programmatic dependent launch inside the short-kernel chains of a frame -------------------------------------------------
// A frame is a chain of ~22 kernels, most of them short (radix passes, scans, tables): at the small configurations the drain /
// launch / ramp-up gap between two of them is a visible share of the frame.  Every kernel of the frame starts with
// ts2d_grid_chain(): it lets the NEXT kernel of the stream be scheduled as soon as all of this kernel's blocks have started
// (griddepcontrol.launch_dependents) and then waits until the PREVIOUS kernel has completed and its writes are visible
// (griddepcontrol.wait) -- before touching global memory, so the data dependences (and the anti-dependences: buffers are
// reused along the chain) are exactly those of plain stream order; only block scheduling and the kernel prologue overlap the
// predecessor's tail.  Completion is transitive (a kernel cannot complete before its own wait returned), so kernel N + 2 sees
// everything kernel N wrote.  Both instructions are no-ops for a kernel launched without the attribute.
// WHICH launches carry the attribute was measured (B200; profiles/r02/README.md): inside the sort / scan / table chains it takes
// 3.5 % off the C4 frame (0.599 -> 0.578 ms) and 14 us off the C3 frame; on the edges into and out of the long kernels (K1, K7, K8,
// row reduction, K9) it gave nothing at C4 and cost 45 us at C3 -- so only the short kernels are launched with it
// (ts2d_launch_chained), the long ones and whatever follows them plainly (ts2d_launch).  TS2D_PDL=0 turns it off everywhere, TS2D_PDL=2
// puts it on every launch of the frame.
__device__ __forceinline__ void ts2d_grid_chain()
{
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
}
static inline int ts2d_pdl()  // 0 = off, 1 = the short-kernel chains (default), 2 = every kernel of the frame (the measured-and-rejected variant)
{
    static int v = -1;
    if (v < 0) {
        const char *e = getenv("TS2D_PDL");
        v = e ? atoi(e) : 1;
        if (v < 0 || v > 2) v = 1;
    }
    return v;
}
template <typename... KArgs, typename... Args>
static inline cudaError_t ts2d_launch_impl(bool chained, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args &&...args)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = ((chained && ts2d_pdl() >= 1) || ts2d_pdl() >= 2) ? 1u : 0u;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
// plain launch: <<<grid, block, smem, s>>>
template <typename... KArgs, typename... Args>
static inline cudaError_t ts2d_launch(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args &&...args)
{
    return ts2d_launch_impl(false, kernel, grid, block, smem, s, static_cast<Args &&>(args)...);
}
// launch of a short kernel behind another kernel of the frame: may be scheduled while its predecessor drains
template <typename... KArgs, typename... Args>
static inline cudaError_t ts2d_launch_chained(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args &&...args)
{
    return ts2d_launch_impl(true, kernel, grid, block, smem, s, static_cast<Args &&>(args)...);
}

// Host-side helpers shared by the .cu translation units
static inline size_t ts2d_align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }
static inline ModelIn ts2d_model_in(const ts2d_geometry *g, const GeomHeader *hdr)
{
    ModelIn m = {nullptr, nullptr, nullptr, -1.0f, 1.0f, nullptr, 0};
    if (g->model) {
        m.f_dc = g->model->f_dc;
        m.f_rest = g->model->f_rest;
        m.logit = g->model->opacity_logit;
        m.ste_thr = g->model->ste_threshold;
        m.ratio = g->model->rescale_ratio;
        m.bg_bits = g->model->bg_depth_from_vertices ? const_cast<uint32_t *>(&hdr->bg_bits) : nullptr;
        m.on = 1;
    }
    return m;
}
static inline ModelOut ts2d_model_out(const ts2d_backward_out *o)
{
    ModelOut m = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, 1};
    if (o->model) {
        const ts2d_model_grads *q = o->model;
        m.g_dc = q->dL_df_dc;
        m.g_rest = q->dL_df_rest;
        m.grad_accum = q->gradient_accum;
        m.grad_denom = q->gradient_denom;
        m.csum = q->contrib_sum;
        m.cmax = q->contrib_max;
        m.cdenom = q->contrib_denom;
        m.max_radii = q->max_radii2D;
        m.fwd_csum = q->fwd_contrib_sum;
        m.fwd_cmax = q->fwd_contrib_max;
        m.radii_div = q->radii_div > 0 ? q->radii_div : 1;
    }
    return m;
}
// tile sort: the tile id sits above the 8 mask bits of an instance key and is sorted in 8-bit digits
static inline int ts2d_tile_bits(int n_tiles)
{
    int b = 1;
    while (b < 24 && (1 << b) < n_tiles) b++;
    return b;
}
static inline int ts2d_tile_sort_passes(int n_tiles) { return (ts2d_tile_bits(n_tiles) + 7) / 8; }
// which of the two (tkey, tval) buffers of the binning state holds the sorted list (the passes ping-pong, starting from buffer 0)
static inline int ts2d_sorted_buf(int n_tiles) { return ts2d_tile_sort_passes(n_tiles) & 1; }
// warps per CTA of the fast composite kernels (tuning knob for experiments: TS2D_CTA_WARPS = 1 | 2 | 4 | 8)
static inline int ts2d_cta_warps()
{
    static int v = -1;
    if (v < 0) {
        const char *e = getenv("TS2D_CTA_WARPS");
        v = e ? atoi(e) : TS2D_DEFAULT_CTA_WARPS;
        if (v != 1 && v != 2 && v != 4 && v != 8) v = TS2D_DEFAULT_CTA_WARPS;
    }
    return v;
}
// background depth the composite kernels use: the scalar of the reference-shaped call, or the device-computed maximum
static inline const float *ts2d_bg_ptr(const ts2d_geometry *g, const GeomState &gs)
{
    return (g->model && g->model->bg_depth_from_vertices) ? reinterpret_cast<const float *>(&gs.hdr->bg_bits) : nullptr;
}


// stage launchers (each returns 0 / cudaError_t)
int ts2d_launch_preprocess(const ts2d_camera *cam, const ts2d_geometry *g, const ts2d_flags *f, int32_t *radii, GeomState gs, cudaStream_t s);
// K2/K3: depth order + scan.  R_host != NULL: R is copied to the host and the stream is synchronised (two-call forward);
// NULL: nothing leaves the device (one-enqueue forward).
int ts2d_launch_order_and_scan(int32_t P, GeomState gs, int64_t *R_host, cudaStream_t s);
// K4-K6.  R_host >= 0: the instance count is known on the host (grids are sized for it); < 0: it only exists in gs.hdr.
// pre_cleared: the caller zeroed the tile ranges and the look-back words of the tile sort at the start of the frame
// (ts2d_forward_clear), so that no memset node sits between the kernels of the frame (ts2d_grid_chain)
int ts2d_launch_binning(const ts2d_camera *cam, const ts2d_geometry *g, const ts2d_flags *f, int64_t R_host, GeomState gs, BinState bs, ImageState is,
                        bool pre_cleared, cudaStream_t s);
int ts2d_launch_contrib_finish(int32_t P, const unsigned long long *csum64, float *contrib_sum, cudaStream_t s);
size_t ts2d_sort_status_bytes(int64_t n_cap);
size_t ts2d_scan_status_bytes(int64_t n_cap);
// atomics-free gradient write-back (ts2d_bwd_reduce.cu)
int ts2d_launch_bwd_rows_prepare(const ts2d_camera *cam, const ts2d_flags *f, GeomState gs, BinState bs, ImageState is, BwdScratch sc, cudaStream_t s);
int ts2d_launch_bwd_rows_reduce(int32_t P, GeomState gs, BwdScratch sc, cudaStream_t s);
int ts2d_launch_render_fwd(const ts2d_camera *cam, const ts2d_geometry *g, const ts2d_flags *f, GeomState gs, const uint32_t *list,
                           ImageState is, const ts2d_forward_out *out, cudaStream_t s);
// fast kernels: `keys` = sorted instance keys (tile << 8 | sub-tile mask), `list` = triangle ids, both in tile-list order
int ts2d_launch_render_fwd_fast(const ts2d_camera *cam, const ts2d_geometry *g, const ts2d_flags *f, GeomState gs, const uint32_t *keys,
                                const uint32_t *list, ImageState is, const ts2d_forward_out *out, bool pre_cleared, cudaStream_t s);
// (the fast backward kernels write per-(sub-tile, entry) rows into `sc`; ts2d_launch_bwd_rows_reduce() turns them into sc.gacc)
int ts2d_launch_render_bwd_fast(const ts2d_camera *cam, const ts2d_geometry *g, const ts2d_flags *f, GeomState gs, const uint32_t *keys,
                                const uint32_t *list, ImageState is, const ts2d_loss_in *loss, BwdScratch sc, cudaStream_t s);
// The fast kernels cover the gamma range the trainer schedules (1..50, VanillaTS_model.py:549-554) with margin;
// outside it (for gamma < 0.6 the ecc <= 10 cut starts to matter; gamma -> 0 makes ecc^(2 gamma) degenerate) the exact mirror kernels are used.
static inline bool ts2d_use_fast(const ts2d_geometry *g, const ts2d_flags *f) { return !f->exact && g->gamma >= 0.6f && g->gamma <= 64.0f; }
// 3D primitive (ts2d_prim3d.cu)
int ts2d_launch_preprocess3d(const ts2d_camera *cam, const ts2d_geometry *g, const ts2d_flags *f, int32_t *radii, GeomState gs, cudaStream_t s);
int ts2d_launch_render3d_fwd(const ts2d_camera *cam, const ts2d_geometry *g, const ts2d_flags *f, GeomState gs, const uint32_t *list,
                             ImageState is, const ts2d_forward_out *out, cudaStream_t s);
int ts2d_launch_render3d_bwd(const ts2d_camera *cam, const ts2d_geometry *g, const ts2d_flags *f, GeomState gs, const uint32_t *list,
                             ImageState is, const ts2d_loss_in *loss, float *gacc, cudaStream_t s);
int ts2d_launch_preprocess3d_bwd(const ts2d_camera *cam, const ts2d_geometry *g, const ts2d_flags *f, const int32_t *radii, GeomState gs,
                                 const float *gacc, const ts2d_backward_out *out, cudaStream_t s);
// fast kernels of the 3D primitive (ts2d_prim3d_fast.cu); same roles as ts2d_launch_render_{fwd,bwd}_fast
int ts2d_launch_render3d_fwd_fast(const ts2d_camera *cam, const ts2d_geometry *g, const ts2d_flags *f, GeomState gs, const uint32_t *keys,
                                  const uint32_t *list, ImageState is, const ts2d_forward_out *out, bool pre_cleared, cudaStream_t s);
int ts2d_launch_render3d_bwd_fast(const ts2d_camera *cam, const ts2d_geometry *g, const ts2d_flags *f, GeomState gs, const uint32_t *keys,
                                  const uint32_t *list, ImageState is, const ts2d_loss_in *loss, BwdScratch sc, cudaStream_t s);
int ts2d_launch_export_geometry3d(int P, GeomState gs, float *v_view, float *normal_view, float *depth, float *rgb, uint8_t *clamped,
                                  uint32_t *tiles_touched, uint32_t *rect_min, uint32_t *rect_max, cudaStream_t s);
int ts2d_launch_render_bwd(const ts2d_camera *cam, const ts2d_geometry *g, const ts2d_flags *f, GeomState gs, const uint32_t *list,
                           ImageState is, const ts2d_loss_in *loss, float *gacc, cudaStream_t s);
int ts2d_launch_preprocess_bwd(const ts2d_camera *cam, const ts2d_geometry *g, const ts2d_flags *f, const int32_t *radii, GeomState gs,
                               const float *gacc, const ts2d_backward_out *out, cudaStream_t s);
