// ts2d_geometry_loss.cu -- fused depth-normal consistency loss (the geometry term of the trainer's loss).
//
//   loss = mean( (1 - <normalize(normal), normal_from_depth(depth)>) * mask )
//
// Replaces DepthNormalLoss.forward + autograd (src/diff_recon/trainers/trainer_utils.py:203-257, called at
// VanillaTS_trainer.py:84 with the rendered depth and normal maps): optional bilinear down-sampling of the depth by 2 (:217; the
// shipped MatrixCity config uses scale_factor 0.5), Scharr gradients of the depth (ScharrFilter, :159-185, zero padding), the
// normal of the depth surface in camera space (:220-229), bilinear up-sampling back to the frame (:232-233), normalisation, a
// mask of the pixels whose depth-gradient norm lies below its 0.9 quantile over the frame (:236-243, torch.quantile = a sort of
// two million values), the normalised rendered normal (:254) and the masked mean (:255) -- about 40 torch kernels forward and
// backward, one of them a full sort.  Here:
//   forward   k_dn_down, k_dn_geom (half-resolution depth, surface normal, gradient norm), k_dn_up_gn, a 4-pass radix SELECT of the
//             two order statistics torch.quantile interpolates between (no sort, no host round trip), k_dn_loss (per pixel: mask,
//             both unit vectors, the loss term; fp64-atomic sum; and the two per-pixel gradient maps the backward needs);
//   backward  k_dn_bwd_half (adjoint of the up-sampling as a gather, chain rule to the Scharr outputs), k_dn_bwd_dh (adjoint of the
//             Scharr correlation), k_dn_bwd_full (adjoint of the down-sampling; scaling by the upstream gradient on the device).
// Every adjoint is a gather: no atomics on floats except the one fp64 loss sum, so the gradients are bit-reproducible.
// The mask is piecewise constant (no gradient), exactly as in the reference (a comparison followed by .float()).
#include "ts2d_common.cuh"

namespace {

struct DnDims {
    int W0, H0;    // frame
    int W, H;      // resolution the depth surface is evaluated at (W0 / 2, H0 / 2 or the frame itself)
    float sx, sy;  // W / W0, H / H0 as torch computes them for interpolate(size=...) (float division)
    int half;
};

// torch upsample_bilinear2d, align_corners = False: source index max(scale (dst + 0.5) - 0.5, 0), neighbour clamped at the border
__device__ __forceinline__ void up_taps(int dst, int n_in, float scale, int &i0, int &i1, float &w1)
{
    const float src = fmaxf(scale * ((float)dst + 0.5f) - 0.5f, 0.0f);
    i0 = min((int)src, n_in - 1);
    i1 = min(i0 + 1, n_in - 1);
    w1 = src - (float)i0;
}

__device__ __forceinline__ float at0(const float *__restrict__ p, int W, int H, int x, int y)  // zero padding
{
    return (x >= 0 && y >= 0 && x < W && y < H) ? p[(size_t)y * W + x] : 0.0f;
}

// Scharr cross-correlation (trainer_utils.py:162-163, /32), zero padding
__device__ __forceinline__ void scharr(const float *__restrict__ d, int W, int H, int x, int y, float &gx, float &gy)
{
    const float a = at0(d, W, H, x - 1, y - 1), b = at0(d, W, H, x, y - 1), c = at0(d, W, H, x + 1, y - 1);
    const float e = at0(d, W, H, x - 1, y), f = at0(d, W, H, x + 1, y);
    const float g = at0(d, W, H, x - 1, y + 1), h = at0(d, W, H, x, y + 1), i = at0(d, W, H, x + 1, y + 1);
    gx = (3.0f * (c - a) + 10.0f * (f - e) + 3.0f * (i - g)) * (1.0f / 32.0f);
    gy = (3.0f * (g - a) + 10.0f * (h - b) + 3.0f * (i - c)) * (1.0f / 32.0f);
}

__global__ void __launch_bounds__(256) k_dn_down(const float *__restrict__ depth, DnDims m, float *__restrict__ d_h)
{
    const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (x >= m.W || y >= m.H) return;
    if (!m.half) {
        d_h[(size_t)y * m.W + x] = depth[(size_t)y * m.W0 + x];
        return;
    }
    const float *r0 = depth + (size_t)(2 * y) * m.W0 + 2 * x, *r1 = r0 + m.W0;  // F.interpolate(scale_factor = 0.5): the 2x2 mean
    d_h[(size_t)y * m.W + x] = 0.5f * (0.5f * r0[0] + 0.5f * r0[1]) + 0.5f * (0.5f * r1[0] + 0.5f * r1[1]);
}

// surface normal of the depth map (un-normalised, :226-229) and the norm of its Scharr gradient, at the working resolution
__global__ void __launch_bounds__(256) k_dn_geom(const float *__restrict__ d_h, DnDims m, float kx, float ky, float *__restrict__ nh, float *__restrict__ gnh)
{
    const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (x >= m.W || y >= m.H) return;
    float gx, gy;
    scharr(d_h, m.W, m.H, x, y, gx, gy);
    const size_t o = (size_t)y * m.W + x, n = (size_t)m.W * m.H;
    const float d = d_h[o];
    const float Dx = gx / d, Dy = gy / d;
    const float cx = (float)x - 0.5f * (float)m.W + 0.5f, cy = (float)y - 0.5f * (float)m.H + 0.5f;
    nh[o] = kx * Dx;            // W Dx / (2 tan_fovx)
    nh[n + o] = ky * Dy;        // H Dy / (2 tan_fovy)
    nh[2 * n + o] = -(1.0f + cx * Dx + cy * Dy);
    gnh[o] = sqrtf(gx * gx + gy * gy);
}

__device__ __forceinline__ float up_sample(const float *__restrict__ p, const DnDims &m, int x0, int x1, float wx, int y0, int y1, float wy)
{
    const float *r0 = p + (size_t)y0 * m.W, *r1 = p + (size_t)y1 * m.W;
    return (1.0f - wy) * ((1.0f - wx) * r0[x0] + wx * r0[x1]) + wy * ((1.0f - wx) * r1[x0] + wx * r1[x1]);
}

__global__ void __launch_bounds__(256) k_dn_up_gn(const float *__restrict__ gnh, DnDims m, float *__restrict__ gn_f)
{
    const int X = blockIdx.x * 32 + (threadIdx.x & 31), Y = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (X >= m.W0 || Y >= m.H0) return;
    float v;
    if (m.half) {
        int x0, x1, y0, y1;
        float wx, wy;
        up_taps(X, m.W, m.sx, x0, x1, wx);
        up_taps(Y, m.H, m.sy, y0, y1, wy);
        v = up_sample(gnh, m, x0, x1, wx, y0, y1, wy);
    } else {
        v = gnh[(size_t)Y * m.W + X];
    }
    gn_f[(size_t)Y * m.W0 + X] = v;
}

// ---- radix select of two order statistics of n non-negative floats (bit order == value order) -----------------------------------
struct SelState {
    uint32_t prefix[2];      // high bits found so far
    uint32_t rank[2];        // rank still to be resolved inside the prefix bucket
    uint32_t hist[2][256];
    float threshold;
};

__global__ void k_sel_init(SelState *st, uint32_t lo, uint32_t hi)
{
    const int t = threadIdx.x;
    st->hist[0][t] = st->hist[1][t] = 0;
    if (t == 0) {
        st->prefix[0] = st->prefix[1] = 0;
        st->rank[0] = lo;
        st->rank[1] = hi;
    }
}

__global__ void __launch_bounds__(256) k_sel_hist(const float *__restrict__ v, int64_t n, int pass, SelState *st)
{
    __shared__ uint32_t s_h[2][256];
    s_h[0][threadIdx.x] = s_h[1][threadIdx.x] = 0;
    __syncthreads();
    const int shift = 24 - 8 * pass;
    const uint32_t p0 = st->prefix[0], p1 = st->prefix[1];
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const uint32_t b = __float_as_uint(v[i]);
        const uint32_t hi = pass ? (b >> (shift + 8)) : 0u, dg = (b >> shift) & 0xFFu;
        if (hi == p0) atomicAdd(&s_h[0][dg], 1u);
        if (hi == p1) atomicAdd(&s_h[1][dg], 1u);
    }
    __syncthreads();
    if (s_h[0][threadIdx.x]) atomicAdd(&st->hist[0][threadIdx.x], s_h[0][threadIdx.x]);
    if (s_h[1][threadIdx.x]) atomicAdd(&st->hist[1][threadIdx.x], s_h[1][threadIdx.x]);
}

// one block: pick the digit of both order statistics, clear the histograms; after the last pass interpolate like torch.quantile
__global__ void k_sel_pick(SelState *st, int pass, float weight)
{
    __shared__ uint32_t s_c[2][256];
    const int t = threadIdx.x;
    s_c[0][t] = st->hist[0][t];
    s_c[1][t] = st->hist[1][t];
    __syncthreads();
    if (t < 2) {
        uint32_t r = st->rank[t], cum = 0;
        int d = 0;
        for (; d < 255; d++) {
            if (r < cum + s_c[t][d]) break;
            cum += s_c[t][d];
        }
        st->prefix[t] = (st->prefix[t] << 8) | (uint32_t)d;
        st->rank[t] = r - cum;
    }
    st->hist[0][t] = st->hist[1][t] = 0;
    __syncthreads();
    if (pass == 3 && t == 0) {
        const float a = __uint_as_float(st->prefix[0]), b = __uint_as_float(st->prefix[1]);
        st->threshold = weight < 0.5f ? a + weight * (b - a) : b - (b - a) * (1.0f - weight);  // at::lerp
    }
}

// ---- per-pixel loss term and the two gradient maps ------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_dn_loss(const float *__restrict__ nh, const float *__restrict__ gn_f, const SelState *__restrict__ st,
                                                 const float *__restrict__ normal, DnDims m, float inv_n, float *__restrict__ g_normal,
                                                 float *__restrict__ g_nf, double *loss_sum)
{
    __shared__ float red[8];
    const int X = blockIdx.x * 32 + (threadIdx.x & 31), Y = blockIdx.y * 8 + (threadIdx.x >> 5);
    float term = 0.0f;
    if (X < m.W0 && Y < m.H0) {
        const size_t N0 = (size_t)m.W0 * m.H0, o = (size_t)Y * m.W0 + X, nhp = (size_t)m.W * m.H;
        float n0, n1, n2;
        if (m.half) {
            int x0, x1, y0, y1;
            float wx, wy;
            up_taps(X, m.W, m.sx, x0, x1, wx);
            up_taps(Y, m.H, m.sy, y0, y1, wy);
            n0 = up_sample(nh, m, x0, x1, wx, y0, y1, wy);
            n1 = up_sample(nh + nhp, m, x0, x1, wx, y0, y1, wy);
            n2 = up_sample(nh + 2 * nhp, m, x0, x1, wx, y0, y1, wy);
        } else {
            n0 = nh[o];
            n1 = nh[nhp + o];
            n2 = nh[2 * nhp + o];
        }
        const float inv_lf = 1.0f / sqrtf(n0 * n0 + n1 * n1 + n2 * n2);
        const float d0 = n0 * inv_lf, d1 = n1 * inv_lf, d2 = n2 * inv_lf;                       // normal of the depth surface
        const float r0 = normal[o], r1 = normal[N0 + o], r2 = normal[2 * N0 + o];
        const float inv_ln = 1.0f / fmaxf(sqrtf(r0 * r0 + r1 * r1 + r2 * r2), 1.0e-8f);         // F.normalize, eps = 1e-8
        const float u0 = r0 * inv_ln, u1 = r1 * inv_ln, u2 = r2 * inv_ln;
        const float mask = gn_f[o] < st->threshold ? 1.0f : 0.0f;
        const float dot = u0 * d0 + u1 * d1 + u2 * d2;
        term = (1.0f - dot) * mask;
        // d loss / d (rendered normal): g = -mask d / N through u = r / |r|;  d loss / d (surface normal before normalisation): same with u, d swapped
        const float c = -mask * inv_n;
        g_normal[o] = c * (d0 - u0 * dot) * inv_ln;
        g_normal[N0 + o] = c * (d1 - u1 * dot) * inv_ln;
        g_normal[2 * N0 + o] = c * (d2 - u2 * dot) * inv_ln;
        g_nf[o] = c * (u0 - d0 * dot) * inv_lf;
        g_nf[N0 + o] = c * (u1 - d1 * dot) * inv_lf;
        g_nf[2 * N0 + o] = c * (u2 - d2 * dot) * inv_lf;
    }
    term = warp_sum(term);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = term;
    __syncthreads();
    if (threadIdx.x < 32) {
        float t = threadIdx.x < 8 ? red[threadIdx.x] : 0.0f;
        t = warp_sum(t);
        if (threadIdx.x == 0 && t != 0.0f) atomicAdd(loss_sum, (double)t);
    }
}

__global__ void k_dn_finalize(const double *loss_sum, double n, float *loss) { *loss = (float)(*loss_sum / n); }

// adjoint of the up-sampling (gather: the few frame pixels whose taps touch this working-resolution pixel), then the chain rule
// through n = (kx Dx, ky Dy, -(1 + cx Dx + cy Dy)), Dx = gx / d, Dy = gy / d  ->  gradients w.r.t. gx, gy and the direct one w.r.t. d
__global__ void __launch_bounds__(256) k_dn_bwd_half(const float *__restrict__ g_nf, const float *__restrict__ d_h, DnDims m, float kx, float ky,
                                                     float *__restrict__ g3)
{
    const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (x >= m.W || y >= m.H) return;
    const size_t N0 = (size_t)m.W0 * m.H0, nhp = (size_t)m.W * m.H, o = (size_t)y * m.W + x;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f;
    if (m.half) {
        // frame pixels Y with a tap on row y satisfy |scale (Y + 0.5) - 0.5 - y| < 1 (plus the clamped borders): a window of
        // 2 / scale + 2 candidates, each tested with the forward's own tap function
        const int Ylo = max(0, (int)floorf(((float)y - 1.0f + 0.5f) / m.sy - 0.5f) - 1), Yhi = min(m.H0 - 1, (int)ceilf(((float)y + 1.0f + 0.5f) / m.sy - 0.5f) + 1);
        const int Xlo = max(0, (int)floorf(((float)x - 1.0f + 0.5f) / m.sx - 0.5f) - 1), Xhi = min(m.W0 - 1, (int)ceilf(((float)x + 1.0f + 0.5f) / m.sx - 0.5f) + 1);
        for (int Y = Ylo; Y <= Yhi; Y++) {
            int y0, y1;
            float wy;
            up_taps(Y, m.H, m.sy, y0, y1, wy);
            const float fy = (y0 == y ? 1.0f - wy : 0.0f) + (y1 == y ? wy : 0.0f);
            if (fy == 0.0f) continue;
            for (int X = Xlo; X <= Xhi; X++) {
                int x0, x1;
                float wx;
                up_taps(X, m.W, m.sx, x0, x1, wx);
                const float f = fy * ((x0 == x ? 1.0f - wx : 0.0f) + (x1 == x ? wx : 0.0f));
                if (f == 0.0f) continue;
                const size_t q = (size_t)Y * m.W0 + X;
                a0 = fmaf(f, g_nf[q], a0);
                a1 = fmaf(f, g_nf[N0 + q], a1);
                a2 = fmaf(f, g_nf[2 * N0 + q], a2);
            }
        }
    } else {
        a0 = g_nf[o];
        a1 = g_nf[N0 + o];
        a2 = g_nf[2 * N0 + o];
    }
    float gx, gy;
    scharr(d_h, m.W, m.H, x, y, gx, gy);
    const float d = d_h[o], inv_d = 1.0f / d;
    const float cx = (float)x - 0.5f * (float)m.W + 0.5f, cy = (float)y - 0.5f * (float)m.H + 0.5f;
    const float gDx = a0 * kx - a2 * cx, gDy = a1 * ky - a2 * cy;
    g3[o] = gDx * inv_d;                                          // d / d gx
    g3[nhp + o] = gDy * inv_d;                                    // d / d gy
    g3[2 * nhp + o] = -(gDx * gx + gDy * gy) * inv_d * inv_d;     // direct d / d depth of Dx = gx / d, Dy = gy / d
}

// adjoint of the Scharr correlation: sum_k K[k] g(p - k)
__global__ void __launch_bounds__(256) k_dn_bwd_dh(const float *__restrict__ g3, DnDims m, float *__restrict__ g_dh)
{
    const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (x >= m.W || y >= m.H) return;
    const size_t nhp = (size_t)m.W * m.H;
    const float *px = g3, *py = g3 + nhp;
    // out(p) = sum over taps (i, j) of K[i][j] g(p - (i - 1, j - 1)); K_x[i][j] = cx[j] * wv[i], K_y[i][j] = wv[j] * cy[i], cx = cy = (-1, 0, 1), wv = (3, 10, 3) / 32
    float acc = 0.0f;
#pragma unroll
    for (int i = 0; i < 3; i++) {
#pragma unroll
        for (int j = 0; j < 3; j++) {
            const float wv_i = (i == 1) ? 10.0f : 3.0f, wv_j = (j == 1) ? 10.0f : 3.0f;
            const float kx = (float)(j - 1) * wv_i, ky = (float)(i - 1) * wv_j;
            const int sx = x - (j - 1), sy = y - (i - 1);
            if (kx != 0.0f) acc = fmaf(kx, at0(px, m.W, m.H, sx, sy), acc);
            if (ky != 0.0f) acc = fmaf(ky, at0(py, m.W, m.H, sx, sy), acc);
        }
    }
    g_dh[(size_t)y * m.W + x] = acc * (1.0f / 32.0f) + g3[2 * nhp + (size_t)y * m.W + x];
}

__global__ void __launch_bounds__(256) k_dn_bwd_full(const float *__restrict__ g_dh, const float *__restrict__ g_normal, DnDims m,
                                                     const float *__restrict__ grad_loss, float *__restrict__ dL_ddepth, float *__restrict__ dL_dnormal)
{
    const int X = blockIdx.x * 32 + (threadIdx.x & 31), Y = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (X >= m.W0 || Y >= m.H0) return;
    const float up = grad_loss ? __ldg(grad_loss) : 1.0f;
    const size_t N0 = (size_t)m.W0 * m.H0, o = (size_t)Y * m.W0 + X;
    if (dL_ddepth) {
        float g;
        if (m.half) {  // adjoint of the 2x2 mean; a last odd row / column is not seen by the loss
            const int x = X >> 1, y = Y >> 1;
            g = (x < m.W && y < m.H) ? 0.25f * g_dh[(size_t)y * m.W + x] : 0.0f;
        } else {
            g = g_dh[o];
        }
        dL_ddepth[o] = up * g;
    }
    if (dL_dnormal) {
        dL_dnormal[o] = up * g_normal[o];
        dL_dnormal[N0 + o] = up * g_normal[N0 + o];
        dL_dnormal[2 * N0 + o] = up * g_normal[2 * N0 + o];
    }
}

struct DnScratch {
    float *d_h, *nh, *gnh, *g3, *g_dh;   // working resolution: 1 + 3 + 1 + 3 + 1 planes
    float *gn_f, *g_normal, *g_nf;       // frame: 1 + 3 + 3 planes
    double *loss_sum;
    SelState *sel;
};

DnDims dn_dims(int W0, int H0, int half)
{
    DnDims m;
    m.W0 = W0;
    m.H0 = H0;
    m.half = half;
    m.W = half ? W0 / 2 : W0;
    m.H = half ? H0 / 2 : H0;
    m.sx = (float)m.W / (float)W0;
    m.sy = (float)m.H / (float)H0;
    return m;
}

size_t dn_carve(void *base, const DnDims &m, DnScratch *sc)
{
    const size_t nh = (size_t)m.W * m.H, n0 = (size_t)m.W0 * m.H0;
    size_t used = 0;
    auto take = [&](size_t bytes) {
        used = ts2d_align_up(used, 256);
        char *p = base ? (char *)base + used : nullptr;
        used += bytes;
        return p;
    };
    DnScratch s;
    s.d_h = (float *)take(4 * nh);
    s.nh = (float *)take(4 * 3 * nh);
    s.gnh = (float *)take(4 * nh);
    s.g3 = (float *)take(4 * 3 * nh);
    s.g_dh = (float *)take(4 * nh);
    s.gn_f = (float *)take(4 * n0);
    s.g_normal = (float *)take(4 * 3 * n0);
    s.g_nf = (float *)take(4 * 3 * n0);
    s.loss_sum = (double *)take(sizeof(double));
    s.sel = (SelState *)take(sizeof(SelState));
    if (sc) *sc = s;
    return ts2d_align_up(used, 256);
}

bool dn_size_ok(int W0, int H0, int half) { return W0 >= 1 && H0 >= 1 && (int64_t)W0 * H0 <= ((int64_t)1 << 30) && (!half || (W0 >= 2 && H0 >= 2)); }

}  // namespace

extern "C" {

size_t ts2d_depth_normal_loss_scratch_bytes(int32_t width, int32_t height, int32_t half_resolution)
{
    if (!dn_size_ok(width, height, half_resolution)) return 0;
    return dn_carve(nullptr, dn_dims(width, height, half_resolution), nullptr);
}

int ts2d_depth_normal_loss_forward(const float *depth, const float *normal, int32_t width, int32_t height, float tan_fovx, float tan_fovy,
                                   int32_t half_resolution, float quantile, float *loss, void *scratch, size_t scratch_bytes, void *stream)
{
    if (!dn_size_ok(width, height, half_resolution) || !(quantile >= 0.0f && quantile <= 1.0f) || !(tan_fovx > 0.0f) || !(tan_fovy > 0.0f)) return TS2D_E_SIZE;
    if (!depth || !normal || !loss || !scratch) return TS2D_E_NULL;
    const DnDims m = dn_dims(width, height, half_resolution);
    DnScratch sc;
    if (scratch_bytes < dn_carve(scratch, m, &sc)) return TS2D_E_STATE_SIZE;
    cudaStream_t s = (cudaStream_t)stream;
    const dim3 gh((m.W + 31) / 32, (m.H + 7) / 8), gf((m.W0 + 31) / 32, (m.H0 + 7) / 8);
    const float kx = (float)m.W / (2.0f * tan_fovx), ky = (float)m.H / (2.0f * tan_fovy);
    const int64_t n0 = (int64_t)m.W0 * m.H0;
    TS2D_CUDA_TRY(cudaMemsetAsync(sc.loss_sum, 0, sizeof(double), s));
    k_dn_down<<<gh, 256, 0, s>>>(depth, m, sc.d_h);
    k_dn_geom<<<gh, 256, 0, s>>>(sc.d_h, m, kx, ky, sc.nh, sc.gnh);
    k_dn_up_gn<<<gf, 256, 0, s>>>(sc.gnh, m, sc.gn_f);
    // torch.quantile: ranks = q * (n - 1) in the input's dtype (fp32), the result interpolates between floor and ceil of it
    const float rank = quantile * (float)(n0 - 1);
    const float lo = floorf(rank), hi = ceilf(rank);
    k_sel_init<<<1, 256, 0, s>>>(sc.sel, (uint32_t)lo, (uint32_t)(hi < (float)(n0 - 1) ? hi : (float)(n0 - 1)));
    const int hb = (int)((n0 + 4095) / 4096 < 592 ? (n0 + 4095) / 4096 : 592);
    for (int pass = 0; pass < 4; pass++) {
        k_sel_hist<<<hb, 256, 0, s>>>(sc.gn_f, n0, pass, sc.sel);
        k_sel_pick<<<1, 256, 0, s>>>(sc.sel, pass, rank - lo);
    }
    k_dn_loss<<<gf, 256, 0, s>>>(sc.nh, sc.gn_f, sc.sel, normal, m, 1.0f / (float)n0, sc.g_normal, sc.g_nf, sc.loss_sum);
    k_dn_finalize<<<1, 1, 0, s>>>(sc.loss_sum, (double)n0, loss);
    return (int)cudaGetLastError();
}

int ts2d_depth_normal_loss_backward(int32_t width, int32_t height, float tan_fovx, float tan_fovy, int32_t half_resolution, const float *grad_loss,
                                    void *scratch, size_t scratch_bytes, float *dL_ddepth, float *dL_dnormal, void *stream)
{
    if (!dn_size_ok(width, height, half_resolution) || !(tan_fovx > 0.0f) || !(tan_fovy > 0.0f)) return TS2D_E_SIZE;
    if (!scratch) return TS2D_E_NULL;
    const DnDims m = dn_dims(width, height, half_resolution);
    DnScratch sc;
    if (scratch_bytes < dn_carve(scratch, m, &sc)) return TS2D_E_STATE_SIZE;
    cudaStream_t s = (cudaStream_t)stream;
    const dim3 gh((m.W + 31) / 32, (m.H + 7) / 8), gf((m.W0 + 31) / 32, (m.H0 + 7) / 8);
    const float kx = (float)m.W / (2.0f * tan_fovx), ky = (float)m.H / (2.0f * tan_fovy);
    if (dL_ddepth) {
        k_dn_bwd_half<<<gh, 256, 0, s>>>(sc.g_nf, sc.d_h, m, kx, ky, sc.g3);
        k_dn_bwd_dh<<<gh, 256, 0, s>>>(sc.g3, m, sc.g_dh);
    }
    if (dL_ddepth || dL_dnormal) k_dn_bwd_full<<<gf, 256, 0, s>>>(sc.g_dh, sc.g_normal, m, grad_loss, dL_ddepth, dL_dnormal);
    return (int)cudaGetLastError();
}

}  // extern "C"
