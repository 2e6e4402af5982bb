// ts2d_loss.cu -- fused image loss on the rendered frame (first step: single GPU, whole frame).
//
//   loss = w_l1 * mean|image - gt| + w_ssim * (1 - mean SSIM(image, gt))
//
// Replaces, for the pixel-wise part of the trainer's loss (src/diff_recon/trainers/VanillaTS_trainer.py:74-75,108), the torch
// composition  L1(image, gt)  (trainer_utils.py:323-324)  and  SSIMLoss()(image, gt)  (trainer_utils.py:45-103): five depth-wise
// 11x11 Gaussian convolutions (window sigma 1.5, zero padding, kernel normalised by its sum, :17-41) of image, gt, image^2, gt^2,
// image*gt, the SSIM map with C1 = 0.01^2, C2 = 0.03^2 (:72-80), its mean -- about 15 kernels forward and as many backward -- by
// one forward and one backward kernel.  The window is separable (exp(-(dx^2 + dy^2) / 2 sigma^2) = g(dx) g(dy)), so each kernel
// runs a horizontal and a vertical 11-tap pass over a shared-memory tile with a 5-pixel halo.
//
//   forward   per pixel: mu1, mu2, raw second moments -> ssim; block-reduced sums of ssim and |x - y| (fp64 atomics: the loss
//             value does not depend on the block order beyond 1e-16), and the three maps the backward needs:
//               d1 = d ssim / d mu1 (raw moments held), d2 = d ssim / d E[x^2], d3 = d ssim / d E[xy];
//   backward  dL/dx = -w_ssim / N * (W * d1 + 2 x (W * d2) + y (W * d3)) + w_l1 / N * sign(x - y)   (W * = the same window).
//
// fp32 throughout, IEEE divide in the SSIM ratio.  Note: the reference's convolutions go through cuDNN with TF32 allowed by default,
// so its own values carry ~1e-3 relative error; tests compare against torch with TF32 off.
#include "ts2d_common.cuh"

namespace {

constexpr int LT = 16;            // tile edge (pixels)
constexpr int LR = 5;             // window radius
constexpr int LH = LT + 2 * LR;   // tile + halo = 26
constexpr float SSIM_C1 = 0.01f * 0.01f, SSIM_C2 = 0.03f * 0.03f;

struct Window {
    float g[2 * LR + 1];  // 1-D weights, normalised so that the 2-D window sums to one
};

__device__ __forceinline__ float block_sum(float v, float *red)
{
    v = warp_sum(v);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) red[warp] = v;
    __syncthreads();
    float t = (threadIdx.x < (LT * LT) / 32) ? red[threadIdx.x] : 0.0f;
    if (warp == 0) t = warp_sum(t);
    __syncthreads();
    return t;  // valid in thread 0
}

__global__ void __launch_bounds__(LT *LT)
k_loss_fwd(const float *__restrict__ img, const float *__restrict__ gt, int W, int H, Window win, float *__restrict__ d1, float *__restrict__ d2,
           float *__restrict__ d3, double *__restrict__ sums)
{
    __shared__ float sx[LH][LH + 1], sy[LH][LH + 1];
    __shared__ float hx[LH][LT], hy[LH][LT], hxx[LH][LT], hyy[LH][LT], hxy[LH][LT];
    __shared__ float red[(LT * LT) / 32];
    const size_t plane = (size_t)blockIdx.z * W * H;
    const int x0 = blockIdx.x * LT - LR, y0 = blockIdx.y * LT - LR;
    for (int i = threadIdx.x; i < LH * LH; i += LT * LT) {
        const int ty = i / LH, tx = i - ty * LH, gx = x0 + tx, gy = y0 + ty;
        const bool in = gx >= 0 && gx < W && gy >= 0 && gy < H;
        sx[ty][tx] = in ? img[plane + (size_t)gy * W + gx] : 0.0f;  // zero padding (F.conv2d(padding=5))
        sy[ty][tx] = in ? gt[plane + (size_t)gy * W + gx] : 0.0f;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < LH * LT; i += LT * LT) {  // horizontal pass
        const int ty = i / LT, tx = i - ty * LT;
        float ax = 0.f, ay = 0.f, axx = 0.f, ayy = 0.f, axy = 0.f;
#pragma unroll
        for (int k = 0; k <= 2 * LR; k++) {
            const float w = win.g[k], a = sx[ty][tx + k], b = sy[ty][tx + k];
            ax = fmaf(w, a, ax);
            ay = fmaf(w, b, ay);
            axx = fmaf(w, a * a, axx);
            ayy = fmaf(w, b * b, ayy);
            axy = fmaf(w, a * b, axy);
        }
        hx[ty][tx] = ax; hy[ty][tx] = ay; hxx[ty][tx] = axx; hyy[ty][tx] = ayy; hxy[ty][tx] = axy;
    }
    __syncthreads();
    const int tx = threadIdx.x % LT, ty = threadIdx.x / LT, px = blockIdx.x * LT + tx, py = blockIdx.y * LT + ty;
    float mu1 = 0.f, mu2 = 0.f, e11 = 0.f, e22 = 0.f, e12 = 0.f;
#pragma unroll
    for (int k = 0; k <= 2 * LR; k++) {  // vertical pass
        const float w = win.g[k];
        mu1 = fmaf(w, hx[ty + k][tx], mu1);
        mu2 = fmaf(w, hy[ty + k][tx], mu2);
        e11 = fmaf(w, hxx[ty + k][tx], e11);
        e22 = fmaf(w, hyy[ty + k][tx], e22);
        e12 = fmaf(w, hxy[ty + k][tx], e12);
    }
    float ssim = 0.0f, l1 = 0.0f;
    if (px < W && py < H) {
        const float s11 = e11 - mu1 * mu1, s22 = e22 - mu2 * mu2, s12 = e12 - mu1 * mu2;
        const float A = 2.0f * mu1 * mu2 + SSIM_C1, B = 2.0f * s12 + SSIM_C2;
        const float D = mu1 * mu1 + mu2 * mu2 + SSIM_C1, E = s11 + s22 + SSIM_C2;
        const float inv_DE = 1.0f / (D * E);
        ssim = A * B * inv_DE;
        // partials of ssim w.r.t. (mu1, s11, s12) at fixed central moments, then to raw moments:
        const float g_s11 = -ssim / E, g_s12 = 2.0f * A * inv_DE;
        const float g_mu1 = 2.0f * mu2 * B * inv_DE - 2.0f * mu1 * ssim / D;
        const size_t o = plane + (size_t)py * W + px;
        d1[o] = g_mu1 - 2.0f * mu1 * g_s11 - mu2 * g_s12;
        d2[o] = g_s11;
        d3[o] = g_s12;
        l1 = fabsf(sx[ty + LR][tx + LR] - sy[ty + LR][tx + LR]);
    }
    const float bs = block_sum(ssim, red), bl = block_sum(l1, red);
    if (threadIdx.x == 0) {
        atomicAdd(sums + 0, (double)bl);
        atomicAdd(sums + 1, (double)bs);
    }
}

__global__ void k_loss_finalize(const double *sums, double n, float w_l1, float w_ssim, float *loss, float *terms)
{
    const double l1 = sums[0] / n, ssim_loss = 1.0 - sums[1] / n;
    loss[0] = (float)((double)w_l1 * l1 + (double)w_ssim * ssim_loss);
    if (terms) {
        terms[0] = (float)l1;
        terms[1] = (float)ssim_loss;
    }
}

__global__ void __launch_bounds__(LT *LT)
k_loss_bwd(const float *__restrict__ img, const float *__restrict__ gt, int W, int H, Window win, const float *__restrict__ d1,
           const float *__restrict__ d2, const float *__restrict__ d3, float c_ssim, float c_l1, const float *__restrict__ grad_loss,
           float *__restrict__ out)
{
    __shared__ float s1[LH][LH + 1], s2[LH][LH + 1], s3[LH][LH + 1];
    __shared__ float h1[LH][LT], h2[LH][LT], h3[LH][LT];
    const size_t plane = (size_t)blockIdx.z * W * H;
    const int x0 = blockIdx.x * LT - LR, y0 = blockIdx.y * LT - LR;
    for (int i = threadIdx.x; i < LH * LH; i += LT * LT) {
        const int ty = i / LH, tx = i - ty * LH, gx = x0 + tx, gy = y0 + ty;
        const bool in = gx >= 0 && gx < W && gy >= 0 && gy < H;
        const size_t o = plane + (size_t)gy * W + gx;
        s1[ty][tx] = in ? d1[o] : 0.0f;  // pixels outside the frame have no ssim term
        s2[ty][tx] = in ? d2[o] : 0.0f;
        s3[ty][tx] = in ? d3[o] : 0.0f;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < LH * LT; i += LT * LT) {
        const int ty = i / LT, tx = i - ty * LT;
        float a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
        for (int k = 0; k <= 2 * LR; k++) {
            const float w = win.g[k];
            a1 = fmaf(w, s1[ty][tx + k], a1);
            a2 = fmaf(w, s2[ty][tx + k], a2);
            a3 = fmaf(w, s3[ty][tx + k], a3);
        }
        h1[ty][tx] = a1; h2[ty][tx] = a2; h3[ty][tx] = a3;
    }
    __syncthreads();
    const int tx = threadIdx.x % LT, ty = threadIdx.x / LT, px = blockIdx.x * LT + tx, py = blockIdx.y * LT + ty;
    if (px >= W || py >= H) return;
    float c1 = 0.f, c2 = 0.f, c3 = 0.f;
#pragma unroll
    for (int k = 0; k <= 2 * LR; k++) {
        const float w = win.g[k];
        c1 = fmaf(w, h1[ty + k][tx], c1);
        c2 = fmaf(w, h2[ty + k][tx], c2);
        c3 = fmaf(w, h3[ty + k][tx], c3);
    }
    const size_t o = plane + (size_t)py * W + px;
    const float x = img[o], y = gt[o], d = x - y;
    const float up = grad_loss ? __ldg(grad_loss) : 1.0f;
    const float sgn = d > 0.0f ? 1.0f : (d < 0.0f ? -1.0f : 0.0f);  // torch.abs backward: sign(0) = 0
    out[o] = up * (c_ssim * (c1 + 2.0f * x * c2 + y * c3) + c_l1 * sgn);
}

Window make_window()
{
    // trainer_utils.py:17-30 with kernel_size 11, sigma 1.5: exp(-((i-5)^2 + (j-5)^2) / (2 sigma^2)) / sum = g(i) g(j) / (sum g)^2
    Window w;
    double g[2 * LR + 1], s = 0.0;
    for (int i = 0; i <= 2 * LR; i++) {
        g[i] = exp(-(double)((i - LR) * (i - LR)) / (2.0 * 1.5 * 1.5));
        s += g[i];
    }
    for (int i = 0; i <= 2 * LR; i++) w.g[i] = (float)(g[i] / s);
    return w;
}

}  // namespace

extern "C" {

size_t ts2d_image_loss_scratch_bytes(int32_t channels, int32_t width, int32_t height)
{
    if (channels < 1 || width < 1 || height < 1) return 0;
    return ts2d_align_up(sizeof(float) * 3 * (size_t)channels * width * height, 256) + 256;  // three maps + the two fp64 sums
}

int ts2d_image_loss_forward(const float *image, const float *gt, int32_t channels, int32_t width, int32_t height, float w_l1, float w_ssim,
                            float *loss, float *terms, void *scratch, size_t scratch_bytes, void *stream)
{
    if (channels < 1 || width < 1 || height < 1 || (int64_t)width * height > ((int64_t)1 << 30) || channels > 65535) return TS2D_E_SIZE;
    if (!image || !gt || !loss || !scratch) return TS2D_E_NULL;
    if (scratch_bytes < ts2d_image_loss_scratch_bytes(channels, width, height)) return TS2D_E_STATE_SIZE;
    cudaStream_t s = (cudaStream_t)stream;
    const size_t n = (size_t)channels * width * height;
    float *d1 = (float *)scratch, *d2 = d1 + n, *d3 = d2 + n;
    double *sums = (double *)((char *)scratch + ts2d_align_up(sizeof(float) * 3 * n, 256));
    TS2D_CUDA_TRY(cudaMemsetAsync(sums, 0, 2 * sizeof(double), s));
    const dim3 grid((width + LT - 1) / LT, (height + LT - 1) / LT, channels);
    k_loss_fwd<<<grid, LT * LT, 0, s>>>(image, gt, width, height, make_window(), d1, d2, d3, sums);
    k_loss_finalize<<<1, 1, 0, s>>>(sums, (double)n, w_l1, w_ssim, loss, terms);
    return (int)cudaGetLastError();
}

int ts2d_image_loss_backward(const float *image, const float *gt, int32_t channels, int32_t width, int32_t height, float w_l1, float w_ssim,
                             const float *grad_loss, const void *scratch, size_t scratch_bytes, float *dL_dimage, void *stream)
{
    if (channels < 1 || width < 1 || height < 1 || (int64_t)width * height > ((int64_t)1 << 30) || channels > 65535) return TS2D_E_SIZE;
    if (!image || !gt || !scratch || !dL_dimage) return TS2D_E_NULL;
    if (scratch_bytes < ts2d_image_loss_scratch_bytes(channels, width, height)) return TS2D_E_STATE_SIZE;
    const size_t n = (size_t)channels * width * height;
    const float *d1 = (const float *)scratch, *d2 = d1 + n, *d3 = d2 + n;
    const dim3 grid((width + LT - 1) / LT, (height + LT - 1) / LT, channels);
    k_loss_bwd<<<grid, LT * LT, 0, (cudaStream_t)stream>>>(image, gt, width, height, make_window(), d1, d2, d3, -w_ssim / (float)n, w_l1 / (float)n,
                                                         grad_loss, dL_dimage);
    return (int)cudaGetLastError();
}

}  // extern "C"
