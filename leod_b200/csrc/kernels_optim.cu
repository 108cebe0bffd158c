// Fused AdamW + gradient clip-by-value + teacher EMA over flat fp32 buffers (one launch per step
// instead of ~4 kernels x 376 tensors).  torch.optim.AdamW arithmetic (modules/detection.py:485-518,
// train.py:236-237) and modules/utils/ssod.py:429-438.
#include <math.h>

#include "common.cuh"

namespace {

__global__ void adamw_ema_kernel(float *__restrict__ p, const float *__restrict__ g, float *__restrict__ m, float *__restrict__ v,
                                 float *__restrict__ ema, int64_t n, float lr, float beta1, float beta2, float eps, float wd,
                                 float clip, float step_size, float bc2_sqrt, float ema_alpha) {
  pdl_prologue();
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float gi = g[i];
    if (clip > 0.f) gi = fminf(fmaxf(gi, -clip), clip);
    float pi = p[i] * (1.f - lr * wd);
    float mi = m[i];
    mi = mi + (1.f - beta1) * (gi - mi);            // exp_avg.lerp_(grad, 1 - beta1)
    const float vi = v[i] * beta2 + (1.f - beta2) * gi * gi;
    const float denom = sqrtf(vi) / bc2_sqrt + eps;
    pi = pi - step_size * (mi / denom);
    p[i] = pi; m[i] = mi; v[i] = vi;
    if (ema) ema[i] = ema[i] * ema_alpha + pi * (1.f - ema_alpha);
  }
}

}  // namespace

extern "C" int leod_adamw_ema(float *p, const float *g, float *m, float *v, float *ema, int64_t n, int step, float lr, float beta1,
                              float beta2, float eps, float weight_decay, float clip_value, float ema_alpha, void *stream) {
  LEOD_REQUIRE(p && g && m && v && n >= 0 && step >= 1, "leod_adamw_ema: bad argument");
  if (n == 0) return 0;
  const double bc1 = 1.0 - pow((double)beta1, step), bc2 = 1.0 - pow((double)beta2, step);
  int blocks = (int)((n + 255) / 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  LEOD_LAUNCH((adamw_ema_kernel), blocks, 256, 0, (cudaStream_t)stream, p, g, m, v, ema, n, lr, beta1, beta2, eps, weight_decay, clip_value,
                                                             (float)(lr / bc1), (float)sqrt(bc2), ema_alpha);
  LEOD_LAUNCH_CHECK();
  return 0;
}
