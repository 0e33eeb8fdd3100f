// Pipe-rate micro-benchmarks for the epilogue design (B200, sm_100a): how many warp-instructions per clock per SM the
// FMA / packed-FMA (FFMA2) / ALU / MUFU pipes sustain, alone and mixed.  Build + run (GPU box):
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/ubench tools/ubench.cu && gpurun_out/ubench
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdio.h>

constexpr int kIters = 2048;
constexpr int kChains = 8;

__device__ __forceinline__ float mufu_ex2(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float mufu_lg2(float x) { float y; asm volatile("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float mufu_rcp(float x) { float y; asm volatile("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

enum Op { FFMA = 0, FFMA2, FADD2, FMUL2, EX2, LG2, RCP, FMNMX, FSEL, IMAD, MIX_EX2_FFMA2, MIX_EX2_4FFMA2, MIX_EX2_8FFMA, MIX_FMNMX_FFMA, F2FP, NUM_OPS };
const char* kNames[NUM_OPS] = {"FFMA", "FFMA2", "FADD2", "FMUL2", "MUFU.EX2", "MUFU.LG2", "MUFU.RCP", "FMNMX", "FSEL", "IMAD",
                               "1 EX2 + 1 FFMA2", "1 EX2 + 4 FFMA2", "1 EX2 + 8 FFMA", "1 FMNMX + 1 FFMA", "F2FP.PACK"};
const int kInstrPerIter[NUM_OPS] = {1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 2, 5, 9, 2, 1};

template <int OP>
__global__ void __launch_bounds__(1024) bench(float* out, float seed) {
  float a[kChains];
  float2 p[kChains];
  int q[kChains];
#pragma unroll
  for (int i = 0; i < kChains; ++i) { a[i] = seed + threadIdx.x * 1e-3f + i; p[i] = make_float2(a[i], a[i] + 0.5f); q[i] = threadIdx.x + i; }
  const float c0 = seed * 0.999f, c1 = seed * 1e-3f;
  const float2 c2 = make_float2(c0, c0), c3 = make_float2(c1, c1);
#pragma unroll 1
  for (int it = 0; it < kIters; ++it) {
#pragma unroll
    for (int i = 0; i < kChains; ++i) {
      if (OP == FFMA) a[i] = fmaf(a[i], c0, c1);
      if (OP == FFMA2) p[i] = __ffma2_rn(p[i], c2, c3);
      if (OP == FADD2) p[i] = __fadd2_rn(p[i], c3);
      if (OP == FMUL2) p[i] = __fmul2_rn(p[i], c2);
      if (OP == EX2) a[i] = mufu_ex2(a[i]);
      if (OP == LG2) a[i] = mufu_lg2(a[i]);
      if (OP == RCP) a[i] = mufu_rcp(a[i]);
      if (OP == FMNMX) a[i] = fminf(a[i], c0 + (float)it);
      if (OP == FSEL) a[i] = (it & (1 << (i & 3))) ? a[i] : c0;
      if (OP == IMAD) q[i] = q[i] * 0x800000 + it;
      if (OP == MIX_EX2_FFMA2) { a[i] = mufu_ex2(a[i]); p[i] = __ffma2_rn(p[i], c2, c3); }
      if (OP == MIX_EX2_4FFMA2) {
        a[i] = mufu_ex2(a[i]);
        p[i] = __ffma2_rn(p[i], c2, c3); p[i] = __ffma2_rn(p[i], c3, c2); p[i] = __ffma2_rn(p[i], c2, c3); p[i] = __ffma2_rn(p[i], c3, c2);
      }
      if (OP == MIX_EX2_8FFMA) {
        a[i] = mufu_ex2(a[i]);
        float t = p[i].x, u = p[i].y;
        t = fmaf(t, c0, c1); u = fmaf(u, c0, c1); t = fmaf(t, c1, c0); u = fmaf(u, c1, c0);
        t = fmaf(t, c0, c1); u = fmaf(u, c0, c1); t = fmaf(t, c1, c0); u = fmaf(u, c1, c0);
        p[i] = make_float2(t, u);
      }
      if (OP == MIX_FMNMX_FFMA) { a[i] = fminf(a[i], c0 + (float)it); p[i].x = fmaf(p[i].x, c0, c1); }
      if (OP == F2FP) {
        __half2 h = __floats2half2_rn(a[i], p[i].x);
        a[i] += __half2float(__low2half(h)) * 0.f + (float)(*reinterpret_cast<unsigned*>(&h) & 1u);
      }
    }
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < kChains; ++i) s += a[i] + p[i].x + p[i].y + (float)q[i];
  if (s == 123.456f) out[0] = s;
}

template <int OP>
void run(int sms, float clock_ghz, float* d_out, int warps_per_sm) {
  const int threads = 256, blocks_per_sm = warps_per_sm * 32 / threads;
  dim3 grid(sms * blocks_per_sm), block(threads);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  bench<OP><<<grid, block>>>(d_out, 1.0001f);
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  for (int r = 0; r < 5; ++r) bench<OP><<<grid, block>>>(d_out, 1.0001f);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  ms /= 5;
  const double warp_instr = (double)sms * warps_per_sm * kIters * kChains * kInstrPerIter[OP];
  const double cycles = ms * 1e-3 * clock_ghz * 1e9;
  printf("%-18s warps/SM %2d : %7.3f ms  %6.3f warp-instr/clk/SM (assuming %.3f GHz)\n", kNames[OP], warps_per_sm, ms,
         warp_instr / cycles / sms, clock_ghz);
}

int main() {
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, 0);
  const int sms = prop.multiProcessorCount;
  const float ghz = prop.clockRate * 1e-6f;
  printf("%s, %d SMs, clockRate %.3f GHz\n", prop.name, sms, ghz);
  float* d_out;
  cudaMalloc(&d_out, 4);
  for (int w : {16, 32, 64}) {
    run<FFMA>(sms, ghz, d_out, w); run<FFMA2>(sms, ghz, d_out, w); run<FADD2>(sms, ghz, d_out, w); run<FMUL2>(sms, ghz, d_out, w);
    run<EX2>(sms, ghz, d_out, w); run<LG2>(sms, ghz, d_out, w); run<RCP>(sms, ghz, d_out, w);
    run<FMNMX>(sms, ghz, d_out, w); run<FSEL>(sms, ghz, d_out, w); run<IMAD>(sms, ghz, d_out, w); run<F2FP>(sms, ghz, d_out, w);
    run<MIX_EX2_FFMA2>(sms, ghz, d_out, w); run<MIX_EX2_4FFMA2>(sms, ghz, d_out, w); run<MIX_EX2_8FFMA>(sms, ghz, d_out, w);
    run<MIX_FMNMX_FFMA>(sms, ghz, d_out, w);
  }
  return 0;
}
