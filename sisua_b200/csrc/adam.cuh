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

// segment (variable) that owns flat offset `o`; tensors start on 64-float boundaries so a float4 never straddles two
__device__ __forceinline__ int find_segment(const SegTable& st, long long o) {
  int lo = 0, hi = st.n - 1;
  while (lo < hi) {
    int mid = (lo + hi + 1) >> 1;
    if (st.off[mid] <= o) lo = mid; else hi = mid - 1;
  }
  return lo;
}

constexpr int kAdamChunk = 2048;   // floats per block

// per-variable sum of squared gradients (double); block 0 also advances the device step counter and publishes lr_t
__global__ void __launch_bounds__(256) grad_sqnorm_kernel(const float* __restrict__ g, SegTable st, long long total,
                                                          double* __restrict__ sq, long long* step,
                                                          long long step_override, float lr, float b1, float b2,
                                                          float* __restrict__ lr_t_out) {
  __shared__ double part[kMaxSegments];
  for (int i = threadIdx.x; i < st.n; i += blockDim.x) part[i] = 0.0;
  __syncthreads();
  const long long base = (long long)blockIdx.x * kAdamChunk;
#pragma unroll
  for (int j = 0; j < kAdamChunk / (256 * 4); ++j) {
    long long o = base + (long long)(j * 256 + threadIdx.x) * 4;
    double a = 0.0;
    int seg = -1;
    if (o < total) {
      float4 v = *reinterpret_cast<const float4*>(g + o);
      a = (double)v.x * v.x + (double)v.y * v.y + (double)v.z * v.z + (double)v.w * v.w;
      seg = find_segment(st, o);
    }
    // a warp covers 128 consecutive floats: almost always one variable -> one shared atomic per warp
    const int seg0 = __shfl_sync(0xffffffffu, seg, 0);
    if (__all_sync(0xffffffffu, seg == seg0)) {
      a = warp_sum(a);
      if ((threadIdx.x & 31) == 0 && seg0 >= 0 && a != 0.0) atomicAdd(&part[seg0], a);
    } else if (seg >= 0 && a != 0.0) {
      atomicAdd(&part[seg], a);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < st.n; i += blockDim.x)
    if (part[i] != 0.0) atomicAdd(&sq[i], part[i]);
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    long long t = step_override > 0 ? step_override : (*step + 1);
    *step = t;
    // Keras / TF formulation: lr_t = lr * sqrt(1 - beta2^t) / (1 - beta1^t)   (one thread; read by adam_kernel)
    *lr_t_out = (float)((double)lr * sqrt(1.0 - pow((double)b2, (double)t)) / (1.0 - pow((double)b1, (double)t)));
  }
}

__global__ void __launch_bounds__(256) adam_kernel(float* __restrict__ p, const float* __restrict__ g,
                                                   float* __restrict__ m, float* __restrict__ v, SegTable st, long long total,
                                                   const double* __restrict__ sq, const float* __restrict__ lr_t_in,
                                                   float b1, float b2, float eps_hat, float clipnorm,
                                                   int clip_mode, float grad_scale) {
  __shared__ float seg_scale[kMaxSegments];
  if (threadIdx.x < st.n) {
    float scale = grad_scale;
    if (clipnorm > 0.f) {
      double n2 = 0.0;
      if (clip_mode == 0) n2 = sq[threadIdx.x]; else for (int i = 0; i < st.n; ++i) n2 += sq[i];
      double nrm = sqrt(n2) * (double)fabsf(grad_scale);
      if (nrm > (double)clipnorm) scale *= (float)((double)clipnorm / nrm);
    }
    seg_scale[threadIdx.x] = scale;
  }
  __syncthreads();
  const float lr_t = *lr_t_in;
  const long long base = (long long)blockIdx.x * kAdamChunk;
#pragma unroll
  for (int j = 0; j < kAdamChunk / (256 * 4); ++j) {
    long long o = base + (long long)(j * 256 + threadIdx.x) * 4;
    if (o < total) {
      const float scale = seg_scale[find_segment(st, o)];
      float4 gv = *reinterpret_cast<const float4*>(g + o);
      float4 mv = *reinterpret_cast<const float4*>(m + o);
      float4 vv = *reinterpret_cast<const float4*>(v + o);
      float4 pv = *reinterpret_cast<const float4*>(p + o);
      float ga[4] = {gv.x * scale, gv.y * scale, gv.z * scale, gv.w * scale};
      float ma[4] = {mv.x, mv.y, mv.z, mv.w}, va[4] = {vv.x, vv.y, vv.z, vv.w}, pa[4] = {pv.x, pv.y, pv.z, pv.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        ma[k] = b1 * ma[k] + (1.f - b1) * ga[k];
        va[k] = b2 * va[k] + (1.f - b2) * ga[k] * ga[k];
        pa[k] -= lr_t * ma[k] / (sqrtf(va[k]) + eps_hat);
      }
      *reinterpret_cast<float4*>(m + o) = make_float4(ma[0], ma[1], ma[2], ma[3]);
      *reinterpret_cast<float4*>(v + o) = make_float4(va[0], va[1], va[2], va[3]);
      *reinterpret_cast<float4*>(p + o) = make_float4(pa[0], pa[1], pa[2], pa[3]);
    }
  }
}

}  // namespace sisua
