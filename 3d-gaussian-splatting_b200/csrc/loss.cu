// Training loss on the device: L1 + SSIM of the rendered image against the ground
// truth, forward AND backward, producing the image gradient directly in the layout the blend
// backward consumes ([H, W, 3], gs_render_backward_final's `grad_final`).
//
// Replaces reference train.py:99-107:
//     l1_loss   = (rendered_img - ground_truth).abs().mean()
//     ssim_loss = 1 - StructuralSimilarityIndexMeasure(data_range=1.0)(img NCHW, gt NCHW)     (torchmetrics)
//     loss      = (1 - w) * l1_loss + w * ssim_loss
// and the autograd graph behind it (~25 full-image torch kernels: permutes, reflect pads, a grouped
// conv2d over 5 stacked maps, its backward, abs backward ...).
// SSIM definition (torchmetrics functional/image/ssim.py `_ssim_update`, gaussian_kernel=True,
// kernel_size=11, sigma=1.5, k1=0.01, k2=0.03): per channel, 11x11 Gaussian window (outer product of the
// normalised 1-D window exp(-(d/1.5)^2/2), d=-5..5); mu, E[x^2], E[y^2], E[xy] by that window;
//     ssim = (2 mu_x mu_y + c1)(2 s_xy + c2) / ((mu_x^2 + mu_y^2 + c1)(s_xx + s_yy + c2)),  c1 = 1e-4, c2 = 9e-4
// averaged over the pixels whose window lies inside the image (torchmetrics reflect-pads by 5 and then
// crops the same 5-pixel border again, so the padding never contributes) and over the 3 channels.
//
// Two passes over 16x16 pixel tiles with a 5-pixel halo staged in shared memory (separable window):
//   1. window statistics -> ssim value + its partial derivatives wrt (mu_x, E[x^2], E[xy]) per pixel/channel;
//      per-block partial sums of |x-y| and ssim (summed later in a fixed order: deterministic loss value);
//   2. the same window applied to the three derivative maps gives d(sum ssim)/dx; combined with the L1
//      sign term into grad_image.
#include <cuda_fp16.h>

#include "internal.h"

namespace {

constexpr int LT = 16;             // tile edge
constexpr int LH = 5;              // window half width
constexpr int LW = LT + 2 * LH;    // tile + halo
constexpr int LK = 2 * LH + 1;

struct GsWin {
  float w[LK];
};

__device__ __forceinline__ float load_target(const float* t, size_t i) { return t[i]; }
__device__ __forceinline__ float load_target(const __half* t, size_t i) { return __half2float(t[i]); }

template <typename GT>
__global__ void __launch_bounds__(256) ssim_stats_kernel(const float* __restrict__ img, const GT* __restrict__ gt, int H,
                                                          int W, GsWin win, float c1, float c2,
                                                          float* __restrict__ abc, float* __restrict__ partial) {
  __shared__ float xs[LW][LW * 3], ys[LW][LW * 3];
  __shared__ float hs[5][LW][LT * 3];
  __shared__ float red[2][8];
  const int bx = blockIdx.x * LT, by = blockIdx.y * LT, tid = threadIdx.x;
  for (int i = tid; i < LW * LW * 3; i += 256) {
    const int r = i / (LW * 3), cc = i % (LW * 3), col = cc / 3, ch = cc % 3;
    const int gy = min(max(by + r - LH, 0), H - 1), gx = min(max(bx + col - LH, 0), W - 1);   // halo outside the image:
    const size_t idx = ((size_t)gy * W + gx) * 3 + ch;                                         // any finite value (unused)
    xs[r][cc] = img[idx];
    ys[r][cc] = load_target(gt, idx);
  }
  __syncthreads();
  for (int i = tid; i < LW * LT * 3; i += 256) {          // horizontal pass
    const int r = i / (LT * 3), cc = i % (LT * 3), col = cc / 3, ch = cc % 3;
    float sx = 0.f, sy = 0.f, sxx = 0.f, syy = 0.f, sxy = 0.f;
#pragma unroll
    for (int k = 0; k < LK; ++k) {
      const float x = xs[r][(col + k) * 3 + ch], y = ys[r][(col + k) * 3 + ch], w = win.w[k];
      sx = fmaf(w, x, sx);
      sy = fmaf(w, y, sy);
      sxx = fmaf(w * x, x, sxx);
      syy = fmaf(w * y, y, syy);
      sxy = fmaf(w * x, y, sxy);
    }
    hs[0][r][cc] = sx;
    hs[1][r][cc] = sy;
    hs[2][r][cc] = sxx;
    hs[3][r][cc] = syy;
    hs[4][r][cc] = sxy;
  }
  __syncthreads();
  float l1 = 0.f, ss = 0.f;
  for (int i = tid; i < LT * LT * 3; i += 256) {          // vertical pass + ssim and its derivatives
    const int r = i / (LT * 3), cc = i % (LT * 3), col = cc / 3, ch = cc % 3;
    const int gy = by + r, gx = bx + col;
    if (gy >= H || gx >= W) continue;
    const float x = xs[r + LH][(col + LH) * 3 + ch], y = ys[r + LH][(col + LH) * 3 + ch];
    l1 += fabsf(x - y);
    float A = 0.f, B = 0.f, C = 0.f;
    if (gy >= LH && gy < H - LH && gx >= LH && gx < W - LH) {
      float mx = 0.f, my = 0.f, exx = 0.f, eyy = 0.f, exy = 0.f;
#pragma unroll
      for (int k = 0; k < LK; ++k) {
        const float w = win.w[k];
        mx = fmaf(w, hs[0][r + k][cc], mx);
        my = fmaf(w, hs[1][r + k][cc], my);
        exx = fmaf(w, hs[2][r + k][cc], exx);
        eyy = fmaf(w, hs[3][r + k][cc], eyy);
        exy = fmaf(w, hs[4][r + k][cc], exy);
      }
      const float sxx = exx - mx * mx, syy = eyy - my * my, sxy = exy - mx * my;
      const float n1 = 2.f * mx * my + c1, n2 = 2.f * sxy + c2;
      const float d1 = mx * mx + my * my + c1, d2 = sxx + syy + c2;
      const float inv = 1.f / (d1 * d2);
      const float s = n1 * n2 * inv;
      ss += s;
      // s as a function of (mu_x, E[x^2], E[xy]) with s_xx = E[x^2] - mu_x^2, s_xy = E[xy] - mu_x mu_y
      A = 2.f * my * (n2 - n1) * inv - 2.f * mx * s * (d2 - d1) * inv;
      B = -s / d2;
      C = 2.f * n1 * inv;
    }
    float* o = abc + (((size_t)gy * W + gx) * 3 + ch) * 3;
    o[0] = A;
    o[1] = B;
    o[2] = C;
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) {
    l1 += __shfl_xor_sync(0xffffffffu, l1, o);
    ss += __shfl_xor_sync(0xffffffffu, ss, o);
  }
  if ((tid & 31) == 0) {
    red[0][tid >> 5] = l1;
    red[1][tid >> 5] = ss;
  }
  __syncthreads();
  if (tid == 0) {
    float a = 0.f, b = 0.f;
    for (int w = 0; w < 8; ++w) {
      a += red[0][w];
      b += red[1][w];
    }
    const int blk = blockIdx.y * gridDim.x + blockIdx.x;
    partial[2 * blk] = a;
    partial[2 * blk + 1] = b;
  }
}

template <typename GT>
__global__ void __launch_bounds__(256) ssim_grad_kernel(const float* __restrict__ img, const GT* __restrict__ gt, int H,
                                                         int W, GsWin win, const float* __restrict__ abc, float k_l1,
                                                         float k_ssim, float* __restrict__ grad) {
  __shared__ float ms[LW][LW * 9];
  __shared__ float hs[LW][LT * 9];
  const int bx = blockIdx.x * LT, by = blockIdx.y * LT, tid = threadIdx.x;
  for (int i = tid; i < LW * LW * 9; i += 256) {
    const int r = i / (LW * 9), cc = i % (LW * 9);
    const int gy = by + r - LH, gx = bx + cc / 9 - LH;
    ms[r][cc] = (gy >= 0 && gy < H && gx >= 0 && gx < W) ? abc[((size_t)gy * W + gx) * 9 + cc % 9] : 0.f;
  }
  __syncthreads();
  for (int i = tid; i < LW * LT * 9; i += 256) {
    const int r = i / (LT * 9), cc = i % (LT * 9), col = cc / 9, m = cc % 9;
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < LK; ++k) s = fmaf(win.w[k], ms[r][(col + k) * 9 + m], s);
    hs[r][cc] = s;
  }
  __syncthreads();
  for (int i = tid; i < LT * LT * 3; i += 256) {
    const int r = i / (LT * 3), cc = i % (LT * 3), col = cc / 3, ch = cc % 3;
    const int gy = by + r, gx = bx + col;
    if (gy >= H || gx >= W) continue;
    float bA = 0.f, bB = 0.f, bC = 0.f;
#pragma unroll
    for (int k = 0; k < LK; ++k) {
      const float w = win.w[k];
      const float* h = &hs[r + k][col * 9 + ch * 3];
      bA = fmaf(w, h[0], bA);
      bB = fmaf(w, h[1], bB);
      bC = fmaf(w, h[2], bC);
    }
    const size_t idx = ((size_t)gy * W + gx) * 3 + ch;
    const float x = img[idx], y = load_target(gt, idx);
    const float sgn = x > y ? 1.f : (x < y ? -1.f : 0.f);                 // torch: d|z|/dz = sign(z), 0 at 0
    grad[idx] = k_ssim * (bA + 2.f * x * bB + y * bC) + k_l1 * sgn;
  }
}

__global__ void __launch_bounds__(256) loss_finalize_kernel(const float* __restrict__ partial, int nblk, double inv_all,
                                                             double inv_inner, float w_l1, float w_ssim, float bias,
                                                             float* __restrict__ out3) {
  __shared__ double red[2][256];
  double a = 0.0, b = 0.0;
  for (int i = threadIdx.x; i < nblk; i += 256) {       // fixed assignment + fixed tree: deterministic
    a += (double)partial[2 * i];
    b += (double)partial[2 * i + 1];
  }
  red[0][threadIdx.x] = a;
  red[1][threadIdx.x] = b;
  __syncthreads();
  for (int s = 128; s; s >>= 1) {
    if (threadIdx.x < s) {
      red[0][threadIdx.x] += red[0][threadIdx.x + s];
      red[1][threadIdx.x] += red[1][threadIdx.x + s];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const float l1 = (float)(red[0][0] * inv_all), ssim = (float)(red[1][0] * inv_inner);
    out3[0] = w_l1 * l1 + w_ssim * ssim + bias;
    out3[1] = l1;
    out3[2] = ssim;
  }
}

GsWin make_window() {
  GsWin g;
  double s = 0.0, v[LK];
  for (int k = 0; k < LK; ++k) {
    const double d = (k - LH) / 1.5;
    v[k] = exp(-0.5 * d * d);
    s += v[k];
  }
  for (int k = 0; k < LK; ++k) g.w[k] = (float)(v[k] / s);
  return g;
}

inline size_t up256(size_t x) { return (x + 255) / 256 * 256; }

}  // namespace

extern "C" size_t gs_loss_workspace_bytes(int height, int width) {
  if (height <= 0 || width <= 0) return 0;
  const size_t nblk = (size_t)((height + LT - 1) / LT) * ((width + LT - 1) / LT);
  return up256((size_t)height * width * 9 * sizeof(float)) + up256(nblk * 2 * sizeof(float));
}

extern "C" int gs_loss_l1_ssim(const float* image, const void* target, int target_is_half, int height, int width,
                               float w_l1, float w_ssim, float bias, float* grad_image, float* out3, void* workspace,
                               size_t workspace_bytes, gs_stream_t stream) {
  if (!image || !target || !out3 || !workspace)
    return gs_set_error_msg(GS_ERR_INVALID_ARG, "gs_loss_l1_ssim: null argument");
  if (height <= 2 * LH || width <= 2 * LH)
    return gs_set_error_msg(GS_ERR_INVALID_ARG, "gs_loss_l1_ssim: the image must be larger than the 11x11 SSIM window");
  if (workspace_bytes < gs_loss_workspace_bytes(height, width))
    return gs_set_error_msg(GS_ERR_INVALID_ARG, "gs_loss_l1_ssim: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  static const GsWin win = make_window();
  float* abc = static_cast<float*>(workspace);
  float* partial = reinterpret_cast<float*>(static_cast<char*>(workspace) + up256((size_t)height * width * 9 * sizeof(float)));
  const dim3 grid((width + LT - 1) / LT, (height + LT - 1) / LT);
  const int nblk = (int)(grid.x * grid.y);
  const float c1 = 0.01f * 0.01f, c2 = 0.03f * 0.03f;                    // (k * data_range)^2, data_range = 1
  const double n_all = (double)height * width * 3, n_inner = (double)(height - 2 * LH) * (width - 2 * LH) * 3;
  if (target_is_half)
    ssim_stats_kernel<__half><<<grid, 256, 0, st>>>(image, static_cast<const __half*>(target), height, width, win, c1, c2,
                                                    abc, partial);
  else
    ssim_stats_kernel<float><<<grid, 256, 0, st>>>(image, static_cast<const float*>(target), height, width, win, c1, c2,
                                                   abc, partial);
  GS_CUDA_TRY(cudaGetLastError());
  gs_count_launch();
  loss_finalize_kernel<<<1, 256, 0, st>>>(partial, nblk, 1.0 / n_all, 1.0 / n_inner, w_l1, w_ssim, bias, out3);
  GS_CUDA_TRY(cudaGetLastError());
  gs_count_launch();
  if (grad_image) {
    const float k_l1 = (float)((double)w_l1 / n_all), k_ssim = (float)((double)w_ssim / n_inner);
    if (target_is_half)
      ssim_grad_kernel<__half><<<grid, 256, 0, st>>>(image, static_cast<const __half*>(target), height, width, win, abc,
                                                     k_l1, k_ssim, grad_image);
    else
      ssim_grad_kernel<float><<<grid, 256, 0, st>>>(image, static_cast<const float*>(target), height, width, win, abc,
                                                    k_l1, k_ssim, grad_image);
    GS_CUDA_TRY(cudaGetLastError());
    gs_count_launch();
  }
  return 0;
}
