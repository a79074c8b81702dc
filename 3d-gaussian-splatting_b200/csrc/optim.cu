// Fused Adam over the flat parameter / gradient buckets.
// The reference steps torch.optim.Adam with 5 parameter groups (train.py:56-64: opa, rgb, pos,
// scale, quat; betas (0.9, 0.99), eps 1e-8, no weight decay) right after the rasterizer backward.
// Here the five gradients already live in ONE flat buffer written by fused_project_bwd_kernel
// (and all-reduced in place), so the optimizer is a single HBM-bound pass: 4 streams read
// (p, g, m, v), 3 written.  Same update as torch's `_single_tensor_adam`:
//   m += (1-b1)(g-m);  v = b2 v + (1-b2) g^2;  p -= (lr / (1-b1^t)) * m / (sqrt(v)/sqrt(1-b2^t) + eps)
#include "internal.h"

#include <cmath>

namespace {

constexpr int kMaxSeg = 8;
struct AdamSegs {
  long long end[kMaxSeg];    // exclusive end (in floats) of each segment of the flat buffer
  float step_size[kMaxSeg];  // lr / bias_correction1 per segment
  int n;
};

__global__ void __launch_bounds__(256) adam_kernel(float4* __restrict__ p, const float4* __restrict__ g,
                                                    float4* __restrict__ m, float4* __restrict__ v, long long n4,
                                                    AdamSegs segs, float beta1, float beta2, float inv_bc2_sqrt,
                                                    float eps) {
  const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
  if (i >= n4) return;
  // segments are 16-byte aligned (renderer._flat_grads), so a float4 never straddles two of them
  const long long e0 = i * 4;
  float step = segs.step_size[0];
#pragma unroll
  for (int s = 1; s < kMaxSeg; ++s)
    if (s < segs.n && e0 >= segs.end[s - 1]) step = segs.step_size[s];
  const float4 gg = g[i];
  float4 mm = m[i], vv = v[i], pp = p[i];
  const float om1 = 1.f - beta1, om2 = 1.f - beta2;
#define GS_ADAM(C)                                            \
  mm.C = fmaf(om1, gg.C - mm.C, mm.C);                        \
  vv.C = fmaf(om2 * gg.C, gg.C, beta2 * vv.C);                \
  pp.C -= step * (mm.C / (sqrtf(vv.C) * inv_bc2_sqrt + eps));
  GS_ADAM(x) GS_ADAM(y) GS_ADAM(z) GS_ADAM(w)
#undef GS_ADAM
  m[i] = mm;
  v[i] = vv;
  p[i] = pp;
}

}  // namespace

extern "C" int gs_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, long long n,
                            const long long* seg_end_host, const float* lr_host, int n_seg, float beta1, float beta2,
                            float eps, int step, gs_stream_t stream) {
  if (n < 0 || n_seg < 1 || n_seg > kMaxSeg || step < 1 || !seg_end_host || !lr_host)
    return gs_set_error_msg(GS_ERR_INVALID_ARG, "gs_adam_step: bad arguments");
  if (n % 4) return gs_set_error_msg(GS_ERR_INVALID_ARG, "gs_adam_step: flat length must be a multiple of 4 floats");
  if (n == 0) return 0;
  AdamSegs segs{};
  segs.n = n_seg;
  const double bc1 = 1.0 - std::pow((double)beta1, (double)step);
  const double bc2 = 1.0 - std::pow((double)beta2, (double)step);
  for (int s = 0; s < n_seg; ++s) {
    if (seg_end_host[s] % 4 || (s && seg_end_host[s] < seg_end_host[s - 1]))
      return gs_set_error_msg(GS_ERR_INVALID_ARG, "gs_adam_step: segment ends must be ascending multiples of 4");
    segs.end[s] = seg_end_host[s];
    segs.step_size[s] = (float)((double)lr_host[s] / bc1);
  }
  const long long n4 = n / 4;
  adam_kernel<<<(unsigned)((n4 + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<float4*>(param), reinterpret_cast<const float4*>(grad), reinterpret_cast<float4*>(exp_avg),
      reinterpret_cast<float4*>(exp_avg_sq), n4, segs, beta1, beta2, (float)(1.0 / std::sqrt(bc2)), eps);
  GS_CUDA_TRY(cudaGetLastError());
  gs_count_launch();
  return 0;
}
