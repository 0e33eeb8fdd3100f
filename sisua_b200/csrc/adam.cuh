// Fused Keras/TF-formulation Adam with clipnorm over the flat parameter buffer
// (configs/base.yaml:46-50; formulas SURVEY.md Appendix A; Q5 per-variable vs global clip).
#pragma once
#include "device_math.cuh"

namespace sisua {

constexpr int kMaxSegments = 48;
struct SegTable {
  int n;
  long long off[kMaxSegments];
  long long size[kMaxSegments];   // floats incl. row padding (padding carries zero gradients)
};

// per-variable sum of squared gradients (double); block (0,0) also advances the device step counter
__global__ void __launch_bounds__(256) grad_sqnorm_kernel(const float* __restrict__ g, SegTable st,
                                                          double* __restrict__ sq, long long* step,
                                                          long long step_override, float lr, float b1, float b2,
                                                          float* __restrict__ lr_t_out) {
  __shared__ double scratch[33];
  int s = blockIdx.y;
  const float* p = g + st.off[s];
  long long n = st.size[s];
  double acc = 0.0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    double v = (double)p[i];
    acc += v * v;
  }
  acc = block_sum(acc, scratch);
  if (threadIdx.x == 0) {
    atomicAdd(&sq[s], acc);
    if (blockIdx.x == 0 && blockIdx.y == 0) {
      long long t = step_override > 0 ? step_override : (*step + 1);
      *step = t;
      // Keras / TF formulation: lr_t = lr * sqrt(1 - beta2^t) / (1 - beta1^t)   (one thread; read by adam_kernel)
      *lr_t_out = (float)((double)lr * sqrt(1.0 - pow((double)b2, (double)t)) / (1.0 - pow((double)b1, (double)t)));
    }
  }
}

__global__ void __launch_bounds__(256) adam_kernel(float* __restrict__ p, const float* __restrict__ g,
                                                   float* __restrict__ m, float* __restrict__ v, SegTable st,
                                                   const double* __restrict__ sq, const float* __restrict__ lr_t_in,
                                                   float b1, float b2, float eps_hat, float clipnorm,
                                                   int clip_mode, float grad_scale) {
  int s = blockIdx.y;
  float scale = grad_scale;
  if (clipnorm > 0.f) {
    double n2 = 0.0;
    if (clip_mode == 0) n2 = sq[s]; else for (int i = 0; i < st.n; ++i) n2 += sq[i];
    double nrm = sqrt(n2) * (double)fabsf(grad_scale);
    if (nrm > (double)clipnorm) scale *= (float)((double)clipnorm / nrm);
  }
  const float lr_t = *lr_t_in;
  long long off = st.off[s], n = st.size[s];
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    long long j = off + i;
    float gv = g[j] * scale;
    float mv = b1 * m[j] + (1.f - b1) * gv;
    float vv = b2 * v[j] + (1.f - b2) * gv * gv;
    m[j] = mv; v[j] = vv;
    p[j] -= lr_t * mv / (sqrtf(vv) + eps_hat);
  }
}

}  // namespace sisua
