// ts2d_api.cu -- extern "C" entry points of libts2d.so (declared in include/ts2d.h), argument
// validation (R2D/src/extension_interface.cu:53-81) and carving of the opaque state blobs
// (replaces Params::*State::fromChunk, R2D/src/param_struct.h:11-125).
#include <stdio.h>

#include <mutex>
#include <vector>

#include "ts2d_common.cuh"

namespace {

// ---- optional per-stage timing (ts2d_profile_*) ----
struct StageEvent {
    int stage;
    cudaEvent_t a, b;
};
bool g_profile = false;
std::vector<StageEvent> g_events;
std::mutex g_profile_mu;

struct StageTimer {
    int stage;
    cudaStream_t s;
    cudaEvent_t a = nullptr, b = nullptr;
    StageTimer(int stage_, cudaStream_t s_) : stage(stage_), s(s_)
    {
        if (!g_profile) return;
        cudaEventCreate(&a);
        cudaEventCreate(&b);
        cudaEventRecord(a, s);
    }
    ~StageTimer()
    {
        if (!a) return;
        cudaEventRecord(b, s);
        std::lock_guard<std::mutex> lk(g_profile_mu);
        g_events.push_back({stage, a, b});
    }
};

}  // namespace

// Host side of a frame's device counters: pinned memory the asynchronous copies land in, and the two events that say when.
struct ts2d_counters {
    ts2d_frame_counters *host = nullptr;  // cudaHostAlloc
    cudaEvent_t ev_r = nullptr;           // num_rendered has landed (recorded right behind the scan: the rest of the frame is still queued)
    cudaEvent_t ev_all = nullptr;         // backward_rows has landed (end of the forward pass)
};

namespace {

struct Carver {
    char *p;
    size_t used;
    explicit Carver(void *base) : p((char *)base), used(0) {}
    template <typename T>
    T *take(size_t count)
    {
        used = ts2d_align_up(used, 256);
        T *r = p ? (T *)(p + used) : (T *)nullptr;
        used += sizeof(T) * count;
        return r;
    }
};

size_t carve_geometry(void *blob, int32_t P, GeomState *gs)
{
    const size_t n = P > 0 ? (size_t)P : 1;
    Carver c(blob);
    GeomState g;
    // header + look-back words first: one memset at the start of a frame clears them all (zero_bytes below)
    g.hdr = c.take<GeomHeader>(1);
    g.sstatus_bytes = ts2d_scan_status_bytes((int64_t)n);
    g.sstatus = (unsigned long long *)c.take<char>(g.sstatus_bytes);
    g.status_bytes = ts2d_sort_status_bytes((int64_t)n);
    g.status = (unsigned long long *)c.take<char>(g.status_bytes);
    g.rec0 = c.take<float4>(3 * n);
    g.rec1 = c.take<float4>(2 * n);
    g.dkey = c.take<uint32_t>(n);
    g.dkey2 = c.take<uint32_t>(n);
    g.tmpk = c.take<uint32_t>(n);
    g.ids = c.take<uint32_t>(n);
    g.ids2 = c.take<uint32_t>(n);
    g.tiles = c.take<uint32_t>(n);
    g.rect = c.take<ushort4>(n);
    g.offs = c.take<uint32_t>(n);
    g.binrec = c.take<uint4>(n);
    g.csum64 = c.take<unsigned long long>(n);
    g.clamp = c.take<uint8_t>(n);
    if (gs) *gs = g;
    return ts2d_align_up(c.used, 256);
}
// bytes from the start of the geometry state that must be zero when a frame starts: header (counters, tickets, histograms)
// and the look-back words of the depth sort and of the scan
size_t geometry_zero_bytes(const GeomState &g) { return (size_t)((char *)g.status - (char *)g.hdr) + g.status_bytes; }

size_t carve_binning(void *blob, int64_t cap, BinState *bs)
{
    const size_t n = cap > 0 ? (size_t)cap : 1;
    Carver c(blob);
    BinState b;
    b.cap = (int64_t)n;
    b.status_bytes = ts2d_sort_status_bytes((int64_t)n);
    b.status = (unsigned long long *)c.take<char>(b.status_bytes);
    b.tkey[0] = c.take<uint32_t>(n);
    b.tkey[1] = c.take<uint32_t>(n);
    b.tval[0] = c.take<uint32_t>(n);
    b.tval[1] = c.take<uint32_t>(n);
    if (bs) *bs = b;
    return ts2d_align_up(c.used, 256);
}
// largest capacity whose binning state fits in `bytes` (the inverse of ts2d_binning_state_bytes: forward and backward both derive
// the array layout from the size of the caller's blob, so the instance count itself never has to be known on the host)
int64_t binning_capacity(size_t bytes)
{
    int64_t lo = 0, hi = (int64_t)(bytes / 16) + 1;
    if (carve_binning(nullptr, 1, nullptr) > bytes) return 0;
    lo = 1;
    while (lo < hi) {
        const int64_t mid = lo + (hi - lo + 1) / 2;
        if (carve_binning(nullptr, mid, nullptr) <= bytes) lo = mid;
        else hi = mid - 1;
    }
    return lo;
}

size_t carve_image(void *blob, int32_t W, int32_t H, ImageState *is)
{
    const size_t gx = (W + TS2D_TILE - 1) / TS2D_TILE, gy = (H + TS2D_TILE - 1) / TS2D_TILE;
    const size_t N = (size_t)W * H;
    Carver c(blob);
    ImageState i;
    i.ranges = c.take<uint2>(gx * gy);
    i.n_contrib = c.take<uint32_t>(N);
    i.final_T = c.take<float>(N);
    i.lastw = c.take<uint32_t>(8 * gx * gy);
    if (is) *is = i;
    return ts2d_align_up(c.used, 256);
}

// backward scratch: accumulators, and (fast kernels) the row machinery of ts2d_bwd_reduce.cu
size_t carve_scratch(void *blob, int32_t P, int64_t cap, int64_t rows, BwdScratch *out)
{
    const size_t n = P > 0 ? (size_t)P : 1;
    Carver c(blob);
    BwdScratch b = {};
    b.gacc = c.take<float>(GACC_STRIDE * n);
    b.cap = cap;
    b.rows_cap = rows;
    if (cap > 0) {
        b.ei = c.take<uint32_t>((size_t)cap);
        b.cnt = c.take<uint8_t>((size_t)cap);
        b.sbase = c.take<uint32_t>((size_t)cap + 1);
        b.sstatus_bytes = ts2d_scan_status_bytes(cap);
        b.sstatus = (unsigned long long *)c.take<char>(b.sstatus_bytes);
        b.rows = c.take<float4>(4 * (size_t)(rows > 0 ? rows : 1));
    }
    if (out) *out = b;
    return ts2d_align_up(c.used, 256);
}

int validate(const ts2d_camera *cam, const ts2d_geometry *g, const ts2d_flags *f)
{
    if (!cam || !g || !f) return TS2D_E_NULL;
    if (cam->width <= 0 || cam->height <= 0 || g->P < 0) return TS2D_E_SIZE;
    if ((int64_t)cam->width * cam->height > (int64_t)1 << 30) return TS2D_E_SIZE;
    if ((cam->width + TS2D_TILE - 1) / TS2D_TILE > 65535 || (cam->height + TS2D_TILE - 1) / TS2D_TILE > 65535) return TS2D_E_SIZE;
    if ((int64_t)((cam->width + TS2D_TILE - 1) / TS2D_TILE) * ((cam->height + TS2D_TILE - 1) / TS2D_TILE) > ((int64_t)1 << 24)) return TS2D_E_SIZE;  // 24 tile bits in an instance key
    if (g->C < 1 || g->C > TS2D_MAX_CHANNELS) return TS2D_E_CHANNELS;
    if (g->gamma < 0.0f) return TS2D_E_GAMMA;
    if (f->shard_world < 1 || f->shard_rank < 0 || f->shard_rank >= f->shard_world) return TS2D_E_SHARD;
    if (f->primitive != TS2D_PRIMITIVE_2D && f->primitive != TS2D_PRIMITIVE_3D) return TS2D_E_PRIMITIVE;
    if (g->P == 0) return 0;
    if (!cam->viewmatrix || !cam->projmatrix || !cam->campos || !g->background || !g->vertex) return TS2D_E_NULL;
    if (g->model) {  // parameter-space inputs replace shs / opacity
        const ts2d_model_inputs *m = g->model;
        if (!g->use_shs || g->M < 1 || !m->f_dc || !m->opacity_logit || (g->M > 1 && !m->f_rest)) return TS2D_E_MODEL;
        if (!(m->rescale_ratio > 0.0f)) return TS2D_E_MODEL;
    } else if (!g->opacity) {
        return TS2D_E_NULL;
    }
    if (g->use_shs) {
        if (!g->model && !g->shs) return TS2D_E_BAD_SHS;
        if (g->C != 3) return TS2D_E_BACKGROUND;
        if (g->sh_degree < 0 || g->sh_degree > 3 || (g->sh_degree + 1) * (g->sh_degree + 1) > g->M) return TS2D_E_SH_DEGREE;
    } else {
        if (!g->feature) return TS2D_E_BAD_FEATURE;
    }
    return 0;
}


int validate_backward_out(const ts2d_geometry *g, const ts2d_backward_out *out)
{
    if (!out->dL_dvertex || !out->dL_dcenter2D || !out->dL_dfeature || !out->dL_dopacity) return TS2D_E_NULL;
    if ((g->model != nullptr) != (out->model != nullptr)) return TS2D_E_MODEL;
    if (g->model) {
        const ts2d_model_grads *q = out->model;
        if (!q->dL_df_dc || (g->M > 1 && !q->dL_df_rest)) return TS2D_E_MODEL;
        if ((q->contrib_sum && !q->fwd_contrib_sum) || (q->contrib_max && !q->fwd_contrib_max)) return TS2D_E_MODEL;
    } else if (g->M > 0 && !out->dL_dshs) {
        return TS2D_E_NULL;
    }
    return 0;
}

inline int dbg_sync(const ts2d_flags *f, cudaStream_t s)
{
    if (!f->debug) return 0;
    cudaError_t e = cudaStreamSynchronize(s);
    if (e != cudaSuccess) return (int)e;
    return (int)cudaGetLastError();
}

#define TS2D_STAGE(stage_id, call)        \
    do {                                 \
        int _r;                          \
        {                                \
            StageTimer _t(stage_id, s);  \
            _r = (call);                 \
        }                                \
        if (_r) return _r;               \
        _r = dbg_sync(flags, s);         \
        if (_r) return _r;               \
    } while (0)

// ---- export kernels ----
__global__ void k_export_geometry(int P, const float4 *rec0, const float4 *rec1, const uint32_t *dkey, const uint32_t *tiles,
                                  const ushort4 *rect, const uint8_t *clamp, float *v2d, float *area2, float *normal_view, float *v_depth,
                                  float *depth, float *rgb, uint8_t *clamped, uint32_t *tiles_touched, uint32_t *rect_min, uint32_t *rect_max)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P) return;
    const bool vis = dkey[i] != 0xFFFFFFFFu;
    float4 a = make_float4(0, 0, 0, 0), b = a, c = a, q0 = a, q1 = a;
    if (vis) {
        a = rec0[3 * (size_t)i];
        b = rec0[3 * (size_t)i + 1];
        c = rec0[3 * (size_t)i + 2];
        if (rec1) {
            q0 = rec1[2 * (size_t)i];
            q1 = rec1[2 * (size_t)i + 1];
        }
    }
    if (v2d) { float *o = v2d + 6 * (size_t)i; o[0] = a.x; o[1] = a.y; o[2] = a.z; o[3] = a.w; o[4] = b.x; o[5] = b.y; }
    if (area2) area2[i] = c.w;
    if (normal_view) { normal_view[3 * i] = q0.x; normal_view[3 * i + 1] = q0.y; normal_view[3 * i + 2] = q0.z; }
    if (v_depth) { v_depth[3 * i] = q0.w; v_depth[3 * i + 1] = q1.x; v_depth[3 * i + 2] = q1.y; }
    if (depth) depth[i] = vis ? __uint_as_float(dkey[i]) : 0.0f;  // the depth sort key IS the fp32 bit pattern of the view depth
    if (rgb) { rgb[3 * i] = c.x; rgb[3 * i + 1] = c.y; rgb[3 * i + 2] = c.z; }
    const uint8_t m = vis ? clamp[i] : 0;
    if (clamped) { clamped[3 * i] = m & 1; clamped[3 * i + 1] = (m >> 1) & 1; clamped[3 * i + 2] = (m >> 2) & 1; }
    if (tiles_touched) tiles_touched[i] = tiles[i];
    const ushort4 r = rect[i];
    if (rect_min) { rect_min[2 * i] = r.x; rect_min[2 * i + 1] = r.y; }
    if (rect_max) { rect_max[2 * i] = r.z; rect_max[2 * i + 1] = r.w; }
}

__global__ void k_export_keys(int64_t R, const uint32_t *tkey, const uint32_t *tval, const uint32_t *dkey, uint64_t *keys, uint32_t *list)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= R) return;
    const uint32_t id = tval[i];
    if (keys) keys[i] = ((uint64_t)(tkey[i] >> TS2D_MASK_BITS) << 32) | dkey[id];  // rasterizer.cu:67-69
    if (list) list[i] = id;
}

__global__ void k_export_opacity(int P, const float4 *rec, int stride, int slot, const uint32_t *dkey, float *opacity)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P) return;
    opacity[i] = dkey[i] != 0xFFFFFFFFu ? rec[(size_t)stride * i + slot].w : 0.0f;
}

// ---- render_up_scale epilogue: bilinear resize by an integer factor, align_corners = False (torch upsample_bilinear2d) ----
// Source index of output pixel d: sc * (d + 0.5) - 0.5 = sc*d + (sc-1)/2: integral for odd sc (one tap), half-way between the two
// middle source pixels for even sc (two taps, weights 0.5 / 0.5).  torch evaluates w_y0 * (w_x0 * a + w_x1 * b) + w_y1 * (w_x0 * c + w_x1 * d).
__global__ void k_downsample(const float *__restrict__ in, float *__restrict__ out, int W, int H, int sc)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= W || y >= H) return;
    const int Wi = W * sc, Hi = H * sc;
    const float *src = in + (size_t)blockIdx.z * Wi * Hi;
    const int x0 = sc * x + (sc - 1) / 2, y0 = sc * y + (sc - 1) / 2;
    const float l1 = (sc & 1) ? 0.0f : 0.5f, l0 = 1.0f - l1;
    const int x1 = min(x0 + 1, Wi - 1), y1 = min(y0 + 1, Hi - 1);
    const float a = src[(size_t)y0 * Wi + x0], b = src[(size_t)y0 * Wi + x1], c = src[(size_t)y1 * Wi + x0], d = src[(size_t)y1 * Wi + x1];
    const float top = __fadd_rn(__fmul_rn(l0, a), __fmul_rn(l1, b)), bot = __fadd_rn(__fmul_rn(l0, c), __fmul_rn(l1, d));
    out[((size_t)blockIdx.z * H + y) * W + x] = __fadd_rn(__fmul_rn(l0, top), __fmul_rn(l1, bot));
}
// adjoint: source pixel (xi, yi) receives w_y * w_x * dL_dout[yi / sc][xi / sc] when it is one of that output pixel's taps, else 0
__global__ void k_downsample_bwd(const float *__restrict__ g_out, float *__restrict__ g_in, int W, int H, int sc)
{
    const int xi = blockIdx.x * blockDim.x + threadIdx.x, yi = blockIdx.y * blockDim.y + threadIdx.y;
    const int Wi = W * sc, Hi = H * sc;
    if (xi >= Wi || yi >= Hi) return;
    const int x = xi / sc, y = yi / sc, rx = xi - x * sc, ry = yi - y * sc, t0 = (sc - 1) / 2;
    const bool odd = sc & 1;
    const float wx = odd ? (rx == t0 ? 1.0f : 0.0f) : ((rx == t0 || rx == t0 + 1) ? 0.5f : 0.0f);
    const float wy = odd ? (ry == t0 ? 1.0f : 0.0f) : ((ry == t0 || ry == t0 + 1) ? 0.5f : 0.0f);
    const float w = wx * wy;
    g_in[((size_t)blockIdx.z * Hi + yi) * Wi + xi] = w != 0.0f ? w * g_out[((size_t)blockIdx.z * H + y) * W + x] : 0.0f;
}

}  // namespace

extern "C" {

int ts2d_abi_version(void) { return TS2D_ABI_VERSION; }

const char *ts2d_error_string(int code)
{
    switch (code) {
    case TS2D_OK: return "ok";
    case TS2D_E_BAD_VERTEX: return "vertex must have dimensions (num_points, 3, 3)";
    case TS2D_E_BAD_FEATURE: return "feature must have dimensions (num_points, num_channels)";
    case TS2D_E_BAD_SHS: return "shs must have dimensions (num_points, (1 + sh_degree) ** 2, 3)";
    case TS2D_E_CHANNELS: return "feature's num_channels can't be larger than MAX_CHANNELS";
    case TS2D_E_BACKGROUND: return "background must have the same number of channels as feature";
    case TS2D_E_GAMMA: return "gamma must be larger than 0";
    case TS2D_E_NULL: return "required pointer is NULL";
    case TS2D_E_STATE_SIZE: return "state buffer smaller than ts2d_*_state_bytes()";
    case TS2D_E_SH_DEGREE: return "sh_degree must be in 0..3 and (sh_degree + 1) ** 2 <= shs.size(1)";
    case TS2D_E_SHARD: return "invalid shard_rank / shard_world";
    case TS2D_E_SIZE: return "image size or primitive count out of range";
    case TS2D_E_PRIMITIVE: return "flags.primitive must be TS2D_PRIMITIVE_2D or TS2D_PRIMITIVE_3D";
    case TS2D_E_FABRIC: return "ts2d_exchange_*: 2 <= world <= 8, 0 <= rank < world, a known operation, 16-byte aligned base, first and count multiples of 4 for fp32 sums";
    case TS2D_E_MODEL: return "model inputs / model gradients inconsistent (need use_shs, f_dc, f_rest for M > 1, opacity_logit, ratio > 0)";
    default: break;
    }
    if (code > 0) return cudaGetErrorString((cudaError_t)code);
    return "unknown ts2d error";
}

size_t ts2d_geometry_state_bytes(int32_t P) { return carve_geometry(nullptr, P, nullptr); }
size_t ts2d_binning_state_bytes(int64_t capacity, int32_t, int32_t) { return carve_binning(nullptr, capacity, nullptr); }
int64_t ts2d_binning_capacity(size_t binning_state_bytes) { return binning_capacity(binning_state_bytes); }
size_t ts2d_image_state_bytes(int32_t W, int32_t H) { return carve_image(nullptr, W, H, nullptr); }
size_t ts2d_backward_scratch_bytes(int32_t P, size_t binning_state_bytes, int64_t backward_rows)
{
    return carve_scratch(nullptr, P, binning_capacity(binning_state_bytes), backward_rows > 0 ? backward_rows : 1, nullptr);
}

}  // extern "C"

namespace {

// rows the scratch has room for: everything behind the start of the row array
BwdScratch scratch_view(void *scratch, size_t scratch_bytes, int32_t P, int64_t cap)
{
    BwdScratch sc;
    carve_scratch(scratch, P, cap, 1, &sc);
    const size_t rows_off = (size_t)((char *)sc.rows - (char *)scratch);
    sc.rows_cap = (cap > 0 && scratch_bytes > rows_off) ? (int64_t)((scratch_bytes - rows_off) / 64) : 0;
    return sc;
}

int forward_geometry_impl(const ts2d_camera *cam, const ts2d_geometry *geom, const ts2d_flags *flags, int32_t *radii, GeomState gs,
                          int64_t *num_rendered_host, cudaStream_t s)
{
    // one memset per frame: counters, tickets, digit histograms, look-back words of the depth sort and of the scan
    TS2D_CUDA_TRY(cudaMemsetAsync(gs.hdr, 0, geometry_zero_bytes(gs), s));
    if (flags->primitive == TS2D_PRIMITIVE_3D)
        TS2D_STAGE(TS2D_STAGE_PREPROCESS, ts2d_launch_preprocess3d(cam, geom, flags, radii, gs, s));
    else
        TS2D_STAGE(TS2D_STAGE_PREPROCESS, ts2d_launch_preprocess(cam, geom, flags, radii, gs, s));
    TS2D_STAGE(TS2D_STAGE_ORDER_SCAN, ts2d_launch_order_and_scan(geom->P, gs, num_rendered_host, s));
    return 0;
}

// One-enqueue forward: everything the render half of the frame needs zeroed (tile ranges, look-back words of the tile sort, the
// fixed-point contrib_sum and contrib_max of the fast kernels) is cleared HERE, in front of K1, so that the kernels of the frame
// follow each other without a memset node in between (ts2d_grid_chain: each kernel's launch overlaps its predecessor's tail); the one
// copy node left is R going to the host right behind the scan (the host layer must not wait for the end of the pass to see it).
// The render half of the geometry header is part of the frame's first memset (forward_geometry_impl).
int forward_clear_impl(const ts2d_camera *cam, const ts2d_geometry *geom, const ts2d_flags *flags, GeomState gs, BinState bs, ImageState is,
                       const ts2d_forward_out *out, cudaStream_t s)
{
    const int n_tiles = ((cam->width + TS2D_TILE - 1) / TS2D_TILE) * ((cam->height + TS2D_TILE - 1) / TS2D_TILE);
    TS2D_CUDA_TRY(cudaMemsetAsync(is.ranges, 0, sizeof(uint2) * (size_t)n_tiles, s));
    TS2D_CUDA_TRY(cudaMemsetAsync(bs.status, 0, ts2d_sort_status_bytes(bs.cap), s));
    if (flags->rich_info && ts2d_use_fast(geom, flags)) {
        TS2D_CUDA_TRY(cudaMemsetAsync(gs.csum64, 0, sizeof(unsigned long long) * (size_t)geom->P, s));
        TS2D_CUDA_TRY(cudaMemsetAsync(out->contrib_max, 0, sizeof(float) * (size_t)geom->P, s));
    }
    return 0;
}

// pre_cleared: called behind forward_clear_impl + forward_geometry_impl in the same enqueue (ts2d_forward)
int forward_render_impl(const ts2d_camera *cam, const ts2d_geometry *geom, const ts2d_flags *flags, int64_t num_rendered, GeomState gs, BinState bs,
                        ImageState is, const ts2d_forward_out *out, ts2d_counters *ctr, bool pre_cleared, cudaStream_t s)
{
    const int n_tiles = ((cam->width + TS2D_TILE - 1) / TS2D_TILE) * ((cam->height + TS2D_TILE - 1) / TS2D_TILE);
    const int sb = ts2d_sorted_buf(n_tiles);
    if (ctr) {  // R first: the host layer waits for it (to return it and to detect an overflow) while the rest of the frame is still queued
        TS2D_CUDA_TRY(cudaMemcpyAsync(&ctr->host->num_rendered, &gs.hdr->num_rendered, sizeof(int64_t), cudaMemcpyDeviceToHost, s));
        TS2D_CUDA_TRY(cudaEventRecord(ctr->ev_r, s));
    }
    // the render half of the header (row counter, tickets and histograms of the tile sort): cleared here so that a render can be
    // repeated on the same geometry state (binning state too small the first time)
    if (!pre_cleared) TS2D_CUDA_TRY(cudaMemsetAsync(&gs.hdr->render, 0, sizeof(gs.hdr->render), s));
    TS2D_STAGE(TS2D_STAGE_BINNING, ts2d_launch_binning(cam, geom, flags, num_rendered, gs, bs, is, pre_cleared, s));
    if (flags->primitive == TS2D_PRIMITIVE_3D && ts2d_use_fast(geom, flags))
        TS2D_STAGE(TS2D_STAGE_RENDER_FWD, ts2d_launch_render3d_fwd_fast(cam, geom, flags, gs, bs.tkey[sb], bs.tval[sb], is, out, pre_cleared, s));
    else if (flags->primitive == TS2D_PRIMITIVE_3D)
        TS2D_STAGE(TS2D_STAGE_RENDER_FWD, ts2d_launch_render3d_fwd(cam, geom, flags, gs, bs.tval[sb], is, out, s));
    else if (ts2d_use_fast(geom, flags))
        TS2D_STAGE(TS2D_STAGE_RENDER_FWD, ts2d_launch_render_fwd_fast(cam, geom, flags, gs, bs.tkey[sb], bs.tval[sb], is, out, pre_cleared, s));
    else
        TS2D_STAGE(TS2D_STAGE_RENDER_FWD, ts2d_launch_render_fwd(cam, geom, flags, gs, bs.tval[sb], is, out, s));
    if (ctr) {
        TS2D_CUDA_TRY(cudaMemcpyAsync(&ctr->host->backward_rows, &gs.hdr->render.bwd_rows, sizeof(int64_t), cudaMemcpyDeviceToHost, s));
        TS2D_CUDA_TRY(cudaEventRecord(ctr->ev_all, s));
    }
    return 0;
}

// P == 0: no work is enqueued; the counters read zero
int counters_clear(ts2d_counters *c, cudaStream_t s)
{
    if (!c) return 0;
    c->host->num_rendered = c->host->backward_rows = 0;
    TS2D_CUDA_TRY(cudaEventRecord(c->ev_r, s));
    TS2D_CUDA_TRY(cudaEventRecord(c->ev_all, s));
    return 0;
}

int validate_render(const ts2d_geometry *geom, const ts2d_flags *flags, const ts2d_forward_out *out)
{
    if (!out) return TS2D_E_NULL;
    if (!out->out_feature) return TS2D_E_NULL;
    if (flags->rich_info && (!out->depth || !out->normal || !out->contrib_sum || !out->contrib_max)) return TS2D_E_NULL;
    return 0;
}

// composite backward into sc.gacc: fast kernels -> rows -> fixed-order reduction; mirror kernels -> the reference's atomics
int backward_composite_impl(const ts2d_camera *cam, const ts2d_geometry *geom, const ts2d_flags *flags, GeomState gs, BinState bs, ImageState is,
                            const ts2d_loss_in *loss, BwdScratch sc, cudaStream_t s)
{
    const int n_tiles = ((cam->width + TS2D_TILE - 1) / TS2D_TILE) * ((cam->height + TS2D_TILE - 1) / TS2D_TILE);
    const int sb = ts2d_sorted_buf(n_tiles);
    if (ts2d_use_fast(geom, flags)) {
        if (sc.rows_cap <= 0) return TS2D_E_STATE_SIZE;
        TS2D_STAGE(TS2D_STAGE_BWD_PREPARE, ts2d_launch_bwd_rows_prepare(cam, flags, gs, bs, is, sc, s));
        if (flags->primitive == TS2D_PRIMITIVE_3D)
            TS2D_STAGE(TS2D_STAGE_RENDER_BWD, ts2d_launch_render3d_bwd_fast(cam, geom, flags, gs, bs.tkey[sb], bs.tval[sb], is, loss, sc, s));
        else
            TS2D_STAGE(TS2D_STAGE_RENDER_BWD, ts2d_launch_render_bwd_fast(cam, geom, flags, gs, bs.tkey[sb], bs.tval[sb], is, loss, sc, s));
        TS2D_STAGE(TS2D_STAGE_BWD_REDUCE, ts2d_launch_bwd_rows_reduce(geom->P, gs, sc, s));
        return dbg_sync(flags, s);
    }
    TS2D_CUDA_TRY(cudaMemsetAsync(sc.gacc, 0, sizeof(float) * GACC_STRIDE * (size_t)geom->P, s));
    if (flags->primitive == TS2D_PRIMITIVE_3D)
        TS2D_STAGE(TS2D_STAGE_RENDER_BWD, ts2d_launch_render3d_bwd(cam, geom, flags, gs, bs.tval[sb], is, loss, sc.gacc, s));
    else
        TS2D_STAGE(TS2D_STAGE_RENDER_BWD, ts2d_launch_render_bwd(cam, geom, flags, gs, bs.tval[sb], is, loss, sc.gacc, s));
    return 0;
}

int backward_geometry_impl(const ts2d_camera *cam, const ts2d_geometry *geom, const ts2d_flags *flags, const int32_t *radii, GeomState gs,
                           const float *gacc, const ts2d_backward_out *out, cudaStream_t s)
{
    if (flags->primitive == TS2D_PRIMITIVE_3D)
        TS2D_STAGE(TS2D_STAGE_PREPROCESS_BWD, ts2d_launch_preprocess3d_bwd(cam, geom, flags, radii, gs, gacc, out, s));
    else
        TS2D_STAGE(TS2D_STAGE_PREPROCESS_BWD, ts2d_launch_preprocess_bwd(cam, geom, flags, radii, gs, gacc, out, s));
    return 0;
}

}  // namespace

extern "C" {

int ts2d_forward_geometry(const ts2d_camera *cam, const ts2d_geometry *geom, const ts2d_flags *flags, int32_t *radii, void *geometry_state,
                          size_t geometry_state_bytes, int64_t *num_rendered_host, void *stream)
{
    int rc = validate(cam, geom, flags);
    if (rc) return rc;
    if (!num_rendered_host) return TS2D_E_NULL;
    *num_rendered_host = 0;
    if (geom->P == 0) return 0;
    if (!radii || !geometry_state) return TS2D_E_NULL;
    GeomState gs;
    if (carve_geometry(geometry_state, geom->P, &gs) > geometry_state_bytes) return TS2D_E_STATE_SIZE;
    return forward_geometry_impl(cam, geom, flags, radii, gs, num_rendered_host, (cudaStream_t)stream);
}

int ts2d_forward_render(const ts2d_camera *cam, const ts2d_geometry *geom, const ts2d_flags *flags, int64_t num_rendered,
                        const void *geometry_state, void *binning_state, size_t binning_state_bytes, void *image_state,
                        size_t image_state_bytes, const ts2d_forward_out *out, ts2d_counters *counters, void *stream)
{
    int rc = validate(cam, geom, flags);
    if (rc) return rc;
    if (geom->P == 0) return counters_clear(counters, (cudaStream_t)stream);
    if (!geometry_state || !binning_state || !image_state) return TS2D_E_NULL;
    if ((rc = validate_render(geom, flags, out)) != 0) return rc;
    if (num_rendered >= ((int64_t)1 << 31)) return TS2D_E_SIZE;
    GeomState gs;
    BinState bs;
    ImageState is;
    carve_geometry(const_cast<void *>(geometry_state), geom->P, &gs);
    const int64_t cap = binning_capacity(binning_state_bytes);
    if (cap < 1 || (num_rendered > 0 && cap < num_rendered)) return TS2D_E_STATE_SIZE;
    carve_binning(binning_state, cap, &bs);
    if (carve_image(image_state, cam->width, cam->height, &is) > image_state_bytes) return TS2D_E_STATE_SIZE;
    return forward_render_impl(cam, geom, flags, num_rendered, gs, bs, is, out, counters, false, (cudaStream_t)stream);
}

int ts2d_forward(const ts2d_camera *cam, const ts2d_geometry *geom, const ts2d_flags *flags, int32_t *radii, void *geometry_state,
                 size_t geometry_state_bytes, void *binning_state, size_t binning_state_bytes, void *image_state, size_t image_state_bytes,
                 const ts2d_forward_out *out, ts2d_counters *counters, void *stream)
{
    int rc = validate(cam, geom, flags);
    if (rc) return rc;
    if (geom->P == 0) return counters_clear(counters, (cudaStream_t)stream);
    if (!radii || !geometry_state || !binning_state || !image_state) return TS2D_E_NULL;
    if ((rc = validate_render(geom, flags, out)) != 0) return rc;
    GeomState gs;
    BinState bs;
    ImageState is;
    if (carve_geometry(geometry_state, geom->P, &gs) > geometry_state_bytes) return TS2D_E_STATE_SIZE;
    const int64_t cap = binning_capacity(binning_state_bytes);
    if (cap < 1) return TS2D_E_STATE_SIZE;
    carve_binning(binning_state, cap, &bs);
    if (carve_image(image_state, cam->width, cam->height, &is) > image_state_bytes) return TS2D_E_STATE_SIZE;
    cudaStream_t s = (cudaStream_t)stream;
    if ((rc = forward_clear_impl(cam, geom, flags, gs, bs, is, out, s)) != 0) return rc;
    if ((rc = forward_geometry_impl(cam, geom, flags, radii, gs, nullptr, s)) != 0) return rc;
    return forward_render_impl(cam, geom, flags, -1, gs, bs, is, out, counters, true, s);
}

int ts2d_counters_create(ts2d_counters **out)
{
    if (!out) return TS2D_E_NULL;
    ts2d_counters *c = new ts2d_counters();
    cudaError_t e = cudaHostAlloc((void **)&c->host, sizeof(ts2d_frame_counters), cudaHostAllocDefault);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->ev_r, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->ev_all, cudaEventDisableTiming);
    if (e != cudaSuccess) {
        ts2d_counters_destroy(c);
        return (int)e;
    }
    c->host->num_rendered = c->host->backward_rows = 0;
    *out = c;
    return 0;
}

void ts2d_counters_destroy(ts2d_counters *c)
{
    if (!c) return;
    if (c->ev_r) cudaEventDestroy(c->ev_r);
    if (c->ev_all) cudaEventDestroy(c->ev_all);
    if (c->host) cudaFreeHost(c->host);
    delete c;
}

int ts2d_counters_num_rendered(ts2d_counters *c, int64_t *num_rendered)
{
    if (!c || !num_rendered) return TS2D_E_NULL;
    TS2D_CUDA_TRY(cudaEventSynchronize(c->ev_r));
    *num_rendered = c->host->num_rendered;
    return 0;
}

int ts2d_counters_backward_rows(ts2d_counters *c, int64_t *backward_rows)
{
    if (!c || !backward_rows) return TS2D_E_NULL;
    TS2D_CUDA_TRY(cudaEventSynchronize(c->ev_all));
    *backward_rows = c->host->backward_rows;
    return 0;
}

int ts2d_read_counters(const void *geometry_state, int32_t P, ts2d_frame_counters *out_host, void *stream)
{
    if (!out_host) return TS2D_E_NULL;
    out_host->num_rendered = out_host->backward_rows = 0;
    if (P <= 0) return 0;
    if (!geometry_state) return TS2D_E_NULL;
    GeomState gs;
    carve_geometry(const_cast<void *>(geometry_state), P, &gs);
    cudaStream_t s = (cudaStream_t)stream;
    TS2D_CUDA_TRY(cudaMemcpyAsync(&out_host->num_rendered, &gs.hdr->num_rendered, sizeof(int64_t), cudaMemcpyDeviceToHost, s));
    TS2D_CUDA_TRY(cudaMemcpyAsync(&out_host->backward_rows, &gs.hdr->render.bwd_rows, sizeof(int64_t), cudaMemcpyDeviceToHost, s));
    TS2D_CUDA_TRY(cudaStreamSynchronize(s));
    return 0;
}

int ts2d_backward(const ts2d_camera *cam, const ts2d_geometry *geom, const ts2d_flags *flags, const int32_t *radii,
                  const void *geometry_state, void *binning_state, size_t binning_state_bytes, const void *image_state, const ts2d_loss_in *loss,
                  const ts2d_backward_out *out, void *scratch, size_t scratch_bytes, void *stream)
{
    int rc = validate(cam, geom, flags);
    if (rc) return rc;
    if (geom->P == 0) return 0;
    if (!radii || !geometry_state || !binning_state || !image_state || !loss || !out || !scratch) return TS2D_E_NULL;
    if (!loss->dL_dout_feature) return TS2D_E_NULL;
    if ((rc = validate_backward_out(geom, out)) != 0) return rc;
    if (flags->rich_info && (!loss->dL_dout_depth || !loss->dL_dout_normal)) return TS2D_E_NULL;
    GeomState gs;
    BinState bs;
    ImageState is;
    carve_geometry(const_cast<void *>(geometry_state), geom->P, &gs);
    const int64_t cap = binning_capacity(binning_state_bytes);
    if (cap < 1) return TS2D_E_STATE_SIZE;
    carve_binning(binning_state, cap, &bs);
    carve_image(const_cast<void *>(image_state), cam->width, cam->height, &is);
    if (scratch_bytes < carve_scratch(nullptr, geom->P, cap, 1, nullptr)) return TS2D_E_STATE_SIZE;
    const BwdScratch sc = scratch_view(scratch, scratch_bytes, geom->P, cap);
    cudaStream_t s = (cudaStream_t)stream;
    if ((rc = backward_composite_impl(cam, geom, flags, gs, bs, is, loss, sc, s)) != 0) return rc;
    return backward_geometry_impl(cam, geom, flags, radii, gs, sc.gacc, out, s);
}

// The two halves of ts2d_backward, for tile-sharded multi-GPU: the per-triangle accumulators at the START of `scratch` (16 floats
// per triangle, P triangles) are what the ranks sum between the two calls -- 64 B/triangle instead of the 240+ B/triangle of the
// final gradients, and K9 then runs replicated on identical data.
int ts2d_backward_composite(const ts2d_camera *cam, const ts2d_geometry *geom, const ts2d_flags *flags, const void *geometry_state,
                            void *binning_state, size_t binning_state_bytes, const void *image_state, const ts2d_loss_in *loss, void *scratch,
                            size_t scratch_bytes, float *accumulators, void *stream)
{
    int rc = validate(cam, geom, flags);
    if (rc) return rc;
    if (geom->P == 0) return 0;
    if (!geometry_state || !binning_state || !image_state || !loss || !scratch || !loss->dL_dout_feature) return TS2D_E_NULL;
    if (flags->rich_info && (!loss->dL_dout_depth || !loss->dL_dout_normal)) return TS2D_E_NULL;
    GeomState gs;
    BinState bs;
    ImageState is;
    carve_geometry(const_cast<void *>(geometry_state), geom->P, &gs);
    const int64_t cap = binning_capacity(binning_state_bytes);
    if (cap < 1) return TS2D_E_STATE_SIZE;
    carve_binning(binning_state, cap, &bs);
    carve_image(const_cast<void *>(image_state), cam->width, cam->height, &is);
    if (scratch_bytes < carve_scratch(nullptr, geom->P, cap, 1, nullptr)) return TS2D_E_STATE_SIZE;
    BwdScratch sc = scratch_view(scratch, scratch_bytes, geom->P, cap);
    if (accumulators) sc.gacc = accumulators;
    return backward_composite_impl(cam, geom, flags, gs, bs, is, loss, sc, (cudaStream_t)stream);
}

int ts2d_backward_geometry(const ts2d_camera *cam, const ts2d_geometry *geom, const ts2d_flags *flags, const int32_t *radii,
                           const void *geometry_state, const ts2d_backward_out *out, const void *scratch, size_t scratch_bytes, void *stream)
{
    int rc = validate(cam, geom, flags);
    if (rc) return rc;
    if (geom->P == 0) return 0;
    if (!radii || !geometry_state || !out || !scratch) return TS2D_E_NULL;
    if ((rc = validate_backward_out(geom, out)) != 0) return rc;
    if (scratch_bytes < sizeof(float) * GACC_STRIDE * (size_t)geom->P) return TS2D_E_STATE_SIZE;
    GeomState gs;
    carve_geometry(const_cast<void *>(geometry_state), geom->P, &gs);
    return backward_geometry_impl(cam, geom, flags, radii, gs, (const float *)scratch, out, (cudaStream_t)stream);
}

int ts2d_export_geometry(const void *geometry_state, int32_t P, float *v2d, float *area2, float *normal_view, float *v_depth, float *depth,
                         float *rgb, uint8_t *clamped, uint32_t *tiles_touched, uint32_t *rect_min, uint32_t *rect_max, void *stream)
{
    if (P <= 0) return 0;
    if (!geometry_state) return TS2D_E_NULL;
    GeomState gs;
    carve_geometry(const_cast<void *>(geometry_state), P, &gs);
    k_export_geometry<<<(P + 255) / 256, 256, 0, (cudaStream_t)stream>>>(P, gs.rec0, (normal_view || v_depth) ? gs.rec1 : nullptr, gs.dkey,
                                                                        gs.tiles, gs.rect, gs.clamp, v2d, area2, normal_view, v_depth, depth,
                                                                        rgb, clamped, tiles_touched, rect_min, rect_max);
    return (int)cudaGetLastError();
}

int ts2d_export_geometry3d(const void *geometry_state, int32_t P, float *v_view, float *normal_view, float *depth, float *rgb, uint8_t *clamped,
                           uint32_t *tiles_touched, uint32_t *rect_min, uint32_t *rect_max, void *stream)
{
    if (P <= 0) return 0;
    if (!geometry_state) return TS2D_E_NULL;
    GeomState gs;
    carve_geometry(const_cast<void *>(geometry_state), P, &gs);
    return ts2d_launch_export_geometry3d(P, gs, v_view, normal_view, depth, rgb, clamped, tiles_touched, rect_min, rect_max, (cudaStream_t)stream);
}

int ts2d_export_model(const void *geometry_state, int32_t P, int32_t primitive, float *opacity, float *background_depth, void *stream)
{
    if (P <= 0) return 0;
    if (!geometry_state) return TS2D_E_NULL;
    GeomState gs;
    carve_geometry(const_cast<void *>(geometry_state), P, &gs);
    cudaStream_t s = (cudaStream_t)stream;
    if (opacity) {
        // the activated opacity sits in the raster record: rec0[3i+1].w (2D) / rec1[2i].w (3D); only visible triangles have one
        k_export_opacity<<<(P + 255) / 256, 256, 0, s>>>(P, primitive == TS2D_PRIMITIVE_3D ? gs.rec1 : gs.rec0, primitive == TS2D_PRIMITIVE_3D ? 2 : 3,
                                                         primitive == TS2D_PRIMITIVE_3D ? 0 : 1, gs.dkey, opacity);
        TS2D_CUDA_TRY(cudaGetLastError());
    }
    if (background_depth) TS2D_CUDA_TRY(cudaMemcpyAsync(background_depth, &gs.hdr->bg_bits, sizeof(float), cudaMemcpyDeviceToDevice, s));
    return 0;
}

int ts2d_downsample(const float *in, float *out, int32_t planes, int32_t out_width, int32_t out_height, int32_t sc, void *stream)
{
    if (!in || !out) return TS2D_E_NULL;
    if (planes < 1 || out_width < 1 || out_height < 1 || sc < 1 || (int64_t)out_width * sc * out_height * sc > ((int64_t)1 << 30)) return TS2D_E_SIZE;
    const dim3 grid((out_width + 31) / 32, (out_height + 7) / 8, planes), block(32, 8);
    k_downsample<<<grid, block, 0, (cudaStream_t)stream>>>(in, out, out_width, out_height, sc);
    return (int)cudaGetLastError();
}

int ts2d_downsample_bwd(const float *dL_dout, float *dL_din, int32_t planes, int32_t out_width, int32_t out_height, int32_t sc, void *stream)
{
    if (!dL_dout || !dL_din) return TS2D_E_NULL;
    if (planes < 1 || out_width < 1 || out_height < 1 || sc < 1 || (int64_t)out_width * sc * out_height * sc > ((int64_t)1 << 30)) return TS2D_E_SIZE;
    const dim3 grid((out_width * sc + 31) / 32, (out_height * sc + 7) / 8, planes), block(32, 8);
    k_downsample_bwd<<<grid, block, 0, (cudaStream_t)stream>>>(dL_dout, dL_din, out_width, out_height, sc);
    return (int)cudaGetLastError();
}

int ts2d_export_binning(const void *geometry_state, const void *binning_state, size_t binning_state_bytes, const void *image_state, int32_t P,
                        int64_t R, int32_t W, int32_t H, uint64_t *keys_sorted, uint32_t *point_list, uint32_t *ranges, void *stream)
{
    if (!geometry_state || !binning_state || !image_state) return TS2D_E_NULL;
    GeomState gs;
    BinState bs;
    ImageState is;
    carve_geometry(const_cast<void *>(geometry_state), P, &gs);
    const int64_t cap = binning_capacity(binning_state_bytes);
    if (cap < R) return TS2D_E_STATE_SIZE;
    carve_binning(const_cast<void *>(binning_state), cap, &bs);
    carve_image(const_cast<void *>(image_state), W, H, &is);
    cudaStream_t s = (cudaStream_t)stream;
    const int sb = ts2d_sorted_buf((int)(((W + TS2D_TILE - 1) / TS2D_TILE) * ((H + TS2D_TILE - 1) / TS2D_TILE)));
    if (R > 0 && (keys_sorted || point_list)) {
        k_export_keys<<<(unsigned)((R + 255) / 256), 256, 0, s>>>(R, bs.tkey[sb], bs.tval[sb], gs.dkey, keys_sorted, point_list);
        TS2D_CUDA_TRY(cudaGetLastError());
    }
    if (ranges) {
        const size_t gx = (W + TS2D_TILE - 1) / TS2D_TILE, gy = (H + TS2D_TILE - 1) / TS2D_TILE;
        TS2D_CUDA_TRY(cudaMemcpyAsync(ranges, is.ranges, sizeof(uint2) * gx * gy, cudaMemcpyDeviceToDevice, s));
    }
    return 0;
}

int ts2d_export_image(const void *image_state, int32_t W, int32_t H, uint32_t *n_contrib, float *final_T, void *stream)
{
    if (!image_state) return TS2D_E_NULL;
    ImageState is;
    carve_image(const_cast<void *>(image_state), W, H, &is);
    cudaStream_t s = (cudaStream_t)stream;
    const size_t N = (size_t)W * H;
    if (n_contrib) TS2D_CUDA_TRY(cudaMemcpyAsync(n_contrib, is.n_contrib, sizeof(uint32_t) * N, cudaMemcpyDeviceToDevice, s));
    if (final_T) TS2D_CUDA_TRY(cudaMemcpyAsync(final_T, is.final_T, sizeof(float) * N, cudaMemcpyDeviceToDevice, s));
    return 0;
}

int ts2d_profile_enable(int enable)
{
    std::lock_guard<std::mutex> lk(g_profile_mu);
    g_profile = enable != 0;
    return 0;
}

int ts2d_profile_read(float *ms_out, int32_t *launches_out)
{
    std::lock_guard<std::mutex> lk(g_profile_mu);
    for (int i = 0; i < TS2D_NUM_STAGES; i++) {
        if (ms_out) ms_out[i] = 0.0f;
        if (launches_out) launches_out[i] = 0;
    }
    int rc = 0;
    for (auto &e : g_events) {
        float ms = 0.0f;
        cudaError_t err = cudaEventSynchronize(e.b);
        if (err == cudaSuccess) err = cudaEventElapsedTime(&ms, e.a, e.b);
        if (err != cudaSuccess) rc = (int)err;
        if (ms_out) ms_out[e.stage] += ms;
        if (launches_out) launches_out[e.stage] += 1;
        cudaEventDestroy(e.a);
        cudaEventDestroy(e.b);
    }
    g_events.clear();
    return rc;
}

}  // extern "C"
