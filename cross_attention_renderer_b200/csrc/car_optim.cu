// Fused Adam over ONE flat fp32 parameter buffer (training step of reference training.py:124-136:
// average_gradients -> clip_grad_norm_(1.0) -> torch.optim.Adam.step, Adam(lr, betas=(0.99, 0.999)),
// train_realestate10k.py:86).  The host keeps every parameter and every gradient as a view into a flat buffer
// (cross_attention_renderer_b200/optim.py), so the gradient exchange is one all-reduce of that buffer, the clip
// scale one device scalar, and the update this one kernel - instead of ~60 per-tensor launches of each.
//
// Arithmetic = torch.optim.Adam (amsgrad off, maximize off), single-tensor path:
//   g' = g * grad_scale (+ weight_decay * p)
//   m = b1 m + (1 - b1) g';  v = b2 v + (1 - b2) g'^2
//   p -= (lr / (1 - b1^t)) * m / (sqrt(v) / sqrt(1 - b2^t) + eps)
#include <math.h>

#include "car_common.cuh"

namespace car {
namespace {

__global__ void __launch_bounds__(256)
k_adam(float *__restrict__ p, const float *__restrict__ g, float *__restrict__ m, float *__restrict__ v, size_t n,
       float lr, float b1, float b2, float eps, float wd, float bc1, float bc2_sqrt, const float *__restrict__ gscale) {
  const float gs = gscale ? *gscale : 1.0f;
  const float step_size = lr / bc1;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    float gi = g[i] * gs;
    const float pi = p[i];
    if (wd != 0.f) gi = fmaf(wd, pi, gi);
    // torch: exp_avg.lerp_(grad, 1 - beta1)  ==  m + (g - m) * (1 - b1)
    const float mi = m[i] + (gi - m[i]) * (1.0f - b1);
    const float vi = v[i] * b2 + (1.0f - b2) * gi * gi;      // mul_(b2).addcmul_(g, g, value = 1 - b2)
    m[i] = mi;
    v[i] = vi;
    const float denom = sqrtf(vi) / bc2_sqrt + eps;
    p[i] = pi - step_size * (mi / denom);                  // addcdiv_(m, denom, value = -step_size)
  }
}

}  // namespace
}  // namespace car

extern "C" int car_adam_step(float *param, const float *grad, float *exp_avg, float *exp_avg_sq, size_t n, float lr,
                             float beta1, float beta2, float eps, float weight_decay, int step, const float *grad_scale,
                             void *stream) {
  using namespace car;
  if (!param || !grad || !exp_avg || !exp_avg_sq || step < 1) { set_error("car_adam_step: bad argument"); return -1; }
  if (n == 0) return 0;
  const float bc1 = 1.0f - powf(beta1, (float)step);
  const float bc2_sqrt = sqrtf(1.0f - powf(beta2, (float)step));
  size_t blocks = (n + 255) / 256;
  const size_t cap = (size_t)sm_count() * 16;
  if (blocks > cap) blocks = cap;
  k_adam<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(param, grad, exp_avg, exp_avg_sq, n, lr, beta1, beta2, eps,
                                                              weight_decay, bc1, bc2_sqrt, grad_scale);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { set_error("car_adam_step: %s", cudaGetErrorString(e)); return (int)e; }
  return 0;
}
