// Fused tcgen05 kernel for the decoder output heads:
//   out = d . W_out^T + b   (tensor cores, accumulators in TMEM, never written to HBM)
//   -> mean / dispersion / dropout-logit activations -> ZINB / NB log-likelihood per cell (epilogue)
//   -> [train] d llk / d out as an fp16 operand tile in shared memory -> two more tcgen05 GEMMs:
//        dD    += G . W_out          (gradient wrt the decoder activations, TMEM-resident across gene tiles)
//        dW_out = G^T . [d | 1]      (weight and bias gradient of the gene tile, flushed with vector reds)
// Forward products are error-compensated 3xFP16 (d = d1 + d2, w = w1 + w2; d1w1 + d1w2 + d2w1 accumulated in
// fp32), i.e. fp32-grade logits; the two gradient GEMMs use single fp16 operands (2^-11 relative).
// One CTA owns a tile of 128 cells and walks a chunk of 32-gene tiles; roles: 16 epilogue warps (TMEM lane
// quarter x 8-gene slice), 1 MMA-issuing thread, 1 bulk-copy (TMA) thread streaming pre-packed weight tiles.
#pragma once
#include "device_math.cuh"
#include "kernels_mid.cuh"
#include "pair_math.cuh"
#include "tc_ptx.cuh"

namespace sisua {
namespace tc {

constexpr int kCellTile = 128;
constexpr int kGeneTile = 32;
constexpr int kK = 64;                 // hidden width = contraction length
constexpr int kEpiWarps = 16;
constexpr int kEpiThreads = kEpiWarps * 32;
constexpr int kOutThreads = kEpiThreads + 64;   // 16 epilogue warps + MMA warp + loader warp
constexpr int kMmaWarp = kEpiWarps, kLoadWarp = kEpiWarps + 1;
constexpr int kTmemCols = 512;
constexpr int kTmemDD = 192, kTmemDWO = 256, kDwoCols = 80;

__host__ __device__ constexpr int w_tile_bytes(int nh) { return nh * 32 * kK * 2; }                 // one fp16 copy
__host__ __device__ constexpr int packed_tile_bytes(int nh) { return 2 * w_tile_bytes(nh) + nh * 32 * 4; }   // w1 | w2 | bias
__host__ __device__ constexpr int packed_tile_stride(int nh) { return (packed_tile_bytes(nh) + 127) / 128 * 128; }

// ---- weight pre-pack: W_out[nh*G, 64] fp32 -> per gene tile (w1 | w2 | bias) in the canonical UMMA layout ----
__device__ __forceinline__ void pack_wout_tile(const float* __restrict__ W, const float* __restrict__ bias,
                                               uint8_t* __restrict__ packed, int G, int nh, int tile) {
  const int rows = nh * 32;
  const int RS = 128, CS = rows / 8 * 128;
  uint8_t* base = packed + (size_t)tile * packed_tile_stride(nh);
  for (int i = threadIdx.x; i < rows * 8; i += blockDim.x) {     // (row, 8-column group)
    int n = i % rows, cg = i / rows;
    int h = n / 32, g = tile * kGeneTile + (n % 32);
    uint32_t hi[4], lo[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float v0 = 0.f, v1 = 0.f;
      if (g < G) {
        const float* src = W + ((size_t)h * G + g) * kK + cg * 8 + 2 * j;
        v0 = src[0]; v1 = src[1];
      }
      split_f16x2(v0, v1, hi[j], lo[j]);
    }
    uint32_t off = (n >> 3) * RS + cg * CS + (n & 7) * 16;
    *reinterpret_cast<uint4*>(base + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    *reinterpret_cast<uint4*>(base + w_tile_bytes(nh) + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
  }
  float* bdst = reinterpret_cast<float*>(base + 2 * w_tile_bytes(nh));
  for (int n = threadIdx.x; n < rows; n += blockDim.x) {
    int h = n / 32, g = tile * kGeneTile + (n % 32);
    bdst[n] = g < G ? bias[(size_t)h * G + g] : 0.f;
  }
}
__global__ void __launch_bounds__(256) pack_wout_kernel(const float* __restrict__ W, const float* __restrict__ bias,
                                                        uint8_t* __restrict__ packed, int G, int nh, int n_tiles) {
  pack_wout_tile(W, bias, packed, G, nh, blockIdx.x);
}

struct OutHeadsArgs {
  const float* D;          // [R, ldD] decoder output: activated (fuse_norm = 0) or the last unit's pre-activation
  int ldD, fuse_norm;      //          whose norm / bias + ReLU + dropout is then applied on load (ns)
  NormSpec ns;
  const float* x;          // [B, G] counts
  const uint8_t* packed;   // pre-packed weight tiles
  float* llk_x;            // [R], zeroed by the caller (gene chunks add atomically)
  // inference outputs (nullable)
  float* out_mean; float* out_disp; float* out_pi;
  // training outputs
  float* dD;               // [R, 64] += d loss / d D
  float* dW;               // [nh*G, 64] += d loss / d W_out
  float* db;               // [nh*G]     += d loss / d b_out
  int R, B, G, n_tiles, tiles_per_chunk;
  int mean_act, disp_act;
  float upstream;          // d loss / d llk_x = -1 / R
  float gscale, inv_gscale;   // power of two <= 1 applied to d llk / d out before the fp16 operand tile (|g| <= max count, and
                           // fp16 ends at 65504): 1 unless the caller declared larger counts (sisua_set_count_bound)
  // scVI (gene softmax over head 0, library-scaled mean, exp dispersion): see the MODE table below
  float2* lse_part;        // [R, n_lse_parts] running (max, sum exp) of the head-0 logits per gene chunk and slice
  int n_lse_parts;
  const float* lib;        // [R] sampled log library size
  float clip_library;
  float* Trow;             // [R] sum_g s_raw_g * d llk / d s_raw_g   (zeroed by the caller)
  float* dlibsum;          // [R] sum_g (d llk / d mean_g) * mean_g     (zeroed by the caller)
  float* dLib;             // [R] d loss / d library (written by the training pass)
};

// MODE: 0 = VAE / DCA / SISUA heads (elementwise links);  scVI passes over the same weight tiles:
//   1 = logsumexp of the head-0 logits (per-row partials), 2 = evaluation (llk, parameters),
//   3 = llk + the two row sums the softmax / library gradients need, 4 = training pass (G tiles + gradient GEMMs)
enum OutMode { MODE_PLAIN = 0, MODE_SCVI_LSE = 1, MODE_SCVI_EVAL = 2, MODE_SCVI_SUMS = 3, MODE_SCVI_TRAIN = 4 };

struct OutSmem {     // offsets into dynamic shared memory (bytes)
  static constexpr int dA1 = 0;                          // [128][80] fp16 (cols 64.. = ones / zero pad, train)
  static constexpr int dA2 = dA1 + 10 * 2048;            // [128][64] fp16
  static constexpr int W0 = dA2 + 8 * 2048;              // 2 stages of packed tiles
  __host__ __device__ static constexpr int Wstage(int nh) { return packed_tile_stride(nh); }
  __host__ __device__ static constexpr int G0(int nh) { return W0 + 2 * Wstage(nh); }       // [128][128] fp16 x 2 stages (train)
  static constexpr int Gstage = 16 * 2048;                                                 // two G stages when training
  __host__ __device__ static constexpr int XS(int nh, bool train) { return G0(nh) + (train ? 2 * Gstage : 0); }   // [8][512] count stash
  __host__ __device__ static constexpr int LLK(int nh, bool train) { return XS(nh, train) + 8 * kEpiThreads * 4; }
  __host__ __device__ static constexpr int BAR(int nh, bool train) { return LLK(nh, train) + (3 * kCellTile + 2 * kK) * 4; }   // llk | T | dlib | norm scale, shift
  __host__ __device__ static constexpr int total(int nh, bool train) { return BAR(nh, train) + 32 * 8 + 16; }
};

enum OutBar { W_FULL = 0, W_FREE = 2, ACC_FULL = 4, ACC_FREE = 6, G_FULL = 8, G_FREE = 10, DWO_FULL = 12, DWO_FREE = 14, DD_FULL = 16, NUM_BARS = 17 };

template <int NH, bool TRAIN, bool VEC, bool FAST, int MODE = MODE_PLAIN>
__global__ void __launch_bounds__(kOutThreads, 1) out_heads_kernel(OutHeadsArgs a) {
  static_assert(TRAIN == (MODE == MODE_SCVI_TRAIN) || MODE == MODE_PLAIN, "only the plain and scVI-train modes run the gradient GEMMs");
  extern __shared__ __align__(128) uint8_t smem[];
  constexpr int N = NH * 32;
  constexpr int NF = MODE == MODE_SCVI_LSE ? 32 : N;     // the logsumexp pass only needs head 0 (rows 0..31 of a tile)
  constexpr bool SCVI = MODE >= MODE_SCVI_EVAL;
  constexpr bool ZI = NH == 3;
  constexpr int W_CS = NH * 512;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OutSmem::BAR(NH, TRAIN));
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + OutSmem::BAR(NH, TRAIN) + 32 * 8);
  float* llk_s = reinterpret_cast<float*>(smem + OutSmem::LLK(NH, TRAIN));
  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  const int row0 = blockIdx.x * kCellTile;
  const int tile_begin = blockIdx.y * a.tiles_per_chunk;
  const int tile_end = min(a.n_tiles, tile_begin + a.tiles_per_chunk);
  const int nt = tile_end - tile_begin;
  if (nt <= 0) return;

  // ---------------- prologue: barriers, TMEM, decoder-activation tile (hi/lo fp16), pads ----------------
  if (t == 0) {
    mbar_init(&bars[W_FULL], 1); mbar_init(&bars[W_FULL + 1], 1);
    // a weight stage is released by the MMA commit and, when there is no backward (whose commit already follows the
    // epilogue), by every epilogue warp once it has read the stage's bias values
    mbar_init(&bars[W_FREE], TRAIN ? 1 : 1 + kEpiWarps); mbar_init(&bars[W_FREE + 1], TRAIN ? 1 : 1 + kEpiWarps);
    mbar_init(&bars[ACC_FULL], 1); mbar_init(&bars[ACC_FULL + 1], 1);
    mbar_init(&bars[ACC_FREE], kEpiWarps); mbar_init(&bars[ACC_FREE + 1], kEpiWarps);
    mbar_init(&bars[G_FULL], kEpiWarps); mbar_init(&bars[G_FULL + 1], kEpiWarps);
    mbar_init(&bars[G_FREE], 1); mbar_init(&bars[G_FREE + 1], 1);
    mbar_init(&bars[DWO_FULL], 1); mbar_init(&bars[DWO_FULL + 1], 1);
    mbar_init(&bars[DWO_FREE], kEpiWarps); mbar_init(&bars[DWO_FREE + 1], kEpiWarps);
    mbar_init(&bars[DD_FULL], 1);
    fence_barrier_init();
  }
  if (warp == kMmaWarp) tmem_alloc(tmem_slot, kTmemCols);
  float* nsc = llk_s + 3 * kCellTile;      // [64] scale, [64] shift of the fused norm
  if (a.fuse_norm) {
    if (t < kK) { float4 q = norm_coeffs4(a.ns, t); nsc[t] = q.x; nsc[kK + t] = q.y; }
    __syncthreads();
  }
  for (int item = t; item < kCellTile * 8; item += kOutThreads) {
    int r = item % kCellTile, cg = item / kCellTile;
    float v[8];
    if (row0 + r < a.R) {
      const float4* src = reinterpret_cast<const float4*>(a.D + (size_t)(row0 + r) * a.ldD + cg * 8);
      float4 p = src[0], q = src[1];
      v[0] = p.x; v[1] = p.y; v[2] = p.z; v[3] = p.w; v[4] = q.x; v[5] = q.y; v[6] = q.z; v[7] = q.w;
      if (a.fuse_norm) {            // h = dropout(relu(a * sc + sh)): what norm_relu_kernel would have written to HBM
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          v[j] = fmaf(v[j], nsc[cg * 8 + j], nsc[kK + cg * 8 + j]);
          if (a.ns.mode != NORM_RAW) v[j] = fmaxf(v[j], 0.f);
        }
        if (a.ns.mode != NORM_RAW && a.ns.drop.rate > 0.f) {
          DropMult8 m = dropout_mult8(a.ns.drop, (uint32_t)(row0 + r), (uint32_t)cg);
#pragma unroll
          for (int j = 0; j < 8; ++j) v[j] *= m.m[j];
        }
      }
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = 0.f;
    }
    uint32_t hi[4], lo[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) split_f16x2(v[2 * j], v[2 * j + 1], hi[j], lo[j]);
    uint32_t off = cg * 2048 + (r >> 3) * 128 + (r & 7) * 16;
    *reinterpret_cast<uint4*>(smem + OutSmem::dA1 + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    *reinterpret_cast<uint4*>(smem + OutSmem::dA2 + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
  }
  if (TRAIN) {
    for (int item = t; item < kCellTile * 2; item += kOutThreads) {      // column groups 8 (ones | 0..) and 9 (zeros)
      int r = item % kCellTile, cg = 8 + item / kCellTile;
      uint32_t off = cg * 2048 + (r >> 3) * 128 + (r & 7) * 16;
      uint32_t first = (cg == 8 && row0 + r < a.R) ? 0x00003C00u : 0u;  // fp16 1.0 in column 64
      *reinterpret_cast<uint4*>(smem + OutSmem::dA1 + off) = make_uint4(first, 0u, 0u, 0u);
    }
    for (int item = t; item < 2 * OutSmem::Gstage / 16; item += kOutThreads)      // zero both G stages once (pad columns stay 0)
      *reinterpret_cast<uint4*>(smem + OutSmem::G0(NH) + item * 16) = make_uint4(0u, 0u, 0u, 0u);
  }
  if (t < 3 * kCellTile) llk_s[t] = 0.f;
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == kLoadWarp) {
    // =========================== loader: pre-packed weight tiles via 1-D TMA bulk copies ===========================
    if (lane == 0) {
      for (int i = 0; i < nt; ++i) {
        const int s = i & 1;
        if (i >= 2) mbar_wait_backoff(&bars[W_FREE + s], ((i >> 1) - 1) & 1);
        mbar_arrive_expect_tx(&bars[W_FULL + s], packed_tile_bytes(NH));
        bulk_copy_g2s(smem + OutSmem::W0 + s * OutSmem::Wstage(NH),
                      a.packed + (size_t)(tile_begin + i) * packed_tile_stride(NH), packed_tile_bytes(NH), &bars[W_FULL + s]);
      }
    }
  } else if (warp == kMmaWarp) {
    // =========================== MMA issuer (one thread) ===========================
    if (lane == 0) {
      const uint32_t idesc_fwd = make_idesc_f16(kCellTile, NF, 0, 0);
      const uint32_t idesc_dd = make_idesc_f16(kCellTile, kK, 0, 1);
      const uint32_t idesc_dwo = make_idesc_f16(kCellTile, kDwoCols, 1, 1);
      const uint32_t idesc_dwo_lo = make_idesc_f16(kCellTile, kK, 1, 1);
      const uint32_t sA1 = smem_u32(smem + OutSmem::dA1), sA2 = smem_u32(smem + OutSmem::dA2);
      const uint32_t sG0 = smem_u32(smem + OutSmem::G0(NH));
      auto fwd = [&](int i) {
        const int s = i & 1;
        const uint32_t sW1 = smem_u32(smem + OutSmem::W0 + s * OutSmem::Wstage(NH)), sW2 = sW1 + w_tile_bytes(NH);
        mbar_wait_backoff(&bars[W_FULL + s], (i >> 1) & 1);
        if (i >= 2) mbar_wait_backoff(&bars[ACC_FREE + s], ((i >> 1) - 1) & 1);
        tc_fence_after();
        const uint32_t d_t = tmem + (uint32_t)(s * N);
        uint32_t acc = 0;
#pragma unroll
        for (int p = 0; p < 3; ++p) {
          const uint32_t sa = (p == 2) ? sA2 : sA1, sb = (p == 1) ? sW2 : sW1;
#pragma unroll
          for (int ks = 0; ks < kK / 16; ++ks) {
            umma_f16(d_t, make_smem_desc(sa + ks * 4096, 2048, 128), make_smem_desc(sb + ks * 2 * W_CS, W_CS, 128), idesc_fwd, acc);
            acc = 1;
          }
        }
        umma_commit(&bars[ACC_FULL + s]);
        if (!TRAIN) umma_commit(&bars[W_FREE + s]);
      };
      fwd(0);
      for (int i = 0; i < nt; ++i) {
        if (i + 1 < nt) fwd(i + 1);
        if (TRAIN) {
          const int s = i & 1;
          const uint32_t sW1 = smem_u32(smem + OutSmem::W0 + s * OutSmem::Wstage(NH));
          const uint32_t sG = sG0 + s * OutSmem::Gstage;
          mbar_wait_backoff(&bars[G_FULL + s], (i >> 1) & 1);
          if (i >= 2) mbar_wait_backoff(&bars[DWO_FREE + s], ((i >> 1) - 1) & 1);
          tc_fence_after();
          // dD[cells, k] += G[cells, n] . W[n, k]   (A: G K-major, B: w1 then w2, MN-major).  Both halves of the weight
          // split are used: the fp16 rounding of a weight is the same for every cell, so with w1 alone the error of dD
          // is correlated across the batch and survives the sums of the backward pass below (measured: 6e-3 of the
          // largest dec.0.W gradient at 18 944 cells); the rounding of G is independent per entry and averages out.
#pragma unroll
          for (int ks = 0; ks < N / 16; ++ks)
            umma_f16(tmem + kTmemDD, make_smem_desc(sG + ks * 4096, 2048, 128), make_smem_desc(sW1 + ks * 256, 128, W_CS),
                     idesc_dd, (i > 0 || ks > 0) ? 1u : 0u);
#pragma unroll
          for (int ks = 0; ks < N / 16; ++ks)
            umma_f16(tmem + kTmemDD, make_smem_desc(sG + ks * 4096, 2048, 128),
                     make_smem_desc(sW1 + w_tile_bytes(NH) + ks * 256, 128, W_CS), idesc_dd, 1u);
          // dW[n, k | 1] = G^T[n, cells] . [d | 1][cells, k]   (A: G MN-major, B: d1 (+ ones column) then d2, MN-major)
#pragma unroll
          for (int ks = 0; ks < kCellTile / 16; ++ks)
            umma_f16(tmem + kTmemDWO + s * kDwoCols, make_smem_desc(sG + ks * 256, 128, 2048),
                     make_smem_desc(sA1 + ks * 256, 128, 2048), idesc_dwo, ks > 0 ? 1u : 0u);
#pragma unroll
          for (int ks = 0; ks < kCellTile / 16; ++ks)
            umma_f16(tmem + kTmemDWO + s * kDwoCols, make_smem_desc(sG + ks * 256, 128, 2048),
                     make_smem_desc(sA2 + ks * 256, 128, 2048), idesc_dwo_lo, 1u);
          umma_commit(&bars[G_FREE + s]);
          umma_commit(&bars[DWO_FULL + s]);
          umma_commit(&bars[W_FREE + s]);
          if (i == nt - 1) umma_commit(&bars[DD_FULL]);
        }
      }
    }
  } else {
    // =========================== epilogue warps ===========================
    const int q = warp & 3, sub = warp >> 2;      // TMEM lane quarter, 8-gene slice of the 32-gene tile
    const int cell = q * 32 + lane;
    const int row = row0 + cell;
    const bool row_ok = row < a.R;
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    // scVI row state: logsumexp of the gene logits (merged from the MODE 1 partials), exp(clipped library)
    float lse = 0.f, eL = 1.f, t_row = 0.f;
    float m_run = -1e30f, s_run = 0.f;
    bool lib_open = false;
    if (SCVI && row_ok) {
      const float2* pp = a.lse_part + (size_t)row * a.n_lse_parts;
      float m = -1e30f;
      for (int i = 0; i < a.n_lse_parts; ++i) m = fmaxf(m, pp[i].x);
      float ssum = 0.f;
      for (int i = 0; i < a.n_lse_parts; ++i) { float2 v = pp[i]; ssum += v.y * __expf(v.x - m); }
      lse = m + logf(ssum);
      const float lr = a.lib[row];
      lib_open = lr >= 0.f && lr <= a.clip_library;
      eL = expf(fminf(fmaxf(lr, 0.f), a.clip_library));
      if (MODE == MODE_SCVI_TRAIN) t_row = a.Trow[row];
    }

    // this thread's 8 counts of tile i straight from global memory (one 32-byte sector per lane; prefetched a
    // tile ahead into registers, so no shared-memory staging and no CTA-wide barriers in the tile loop)
    const float* xrow = a.x + (size_t)((row_ok ? row : 0) % a.B) * a.G;
    auto load_x = [&](int i, float* xv) {
      const int g = (tile_begin + i) * kGeneTile + sub * 8;
      if (VEC) {
        float4 p0 = make_float4(0.f, 0.f, 0.f, 0.f), p1 = p0;
        if (row_ok && g < a.G) p0 = __ldg(reinterpret_cast<const float4*>(xrow + g));
        if (row_ok && g + 4 < a.G) p1 = __ldg(reinterpret_cast<const float4*>(xrow + g + 4));
        xv[0] = p0.x; xv[1] = p0.y; xv[2] = p0.z; xv[3] = p0.w; xv[4] = p1.x; xv[5] = p1.y; xv[6] = p1.z; xv[7] = p1.w;
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) xv[j] = (row_ok && g + j < a.G) ? __ldg(xrow + g + j) : 0.f;
      }
    };

    auto flush_dwo = [&](int i) {            // tile i's weight / bias gradient: TMEM -> vector reds
      const int s = i & 1;
      mbar_wait(&bars[DWO_FULL + s], (i >> 1) & 1);
      tc_fence_after();
      const int n = cell;                    // TMEM lane = output-unit row of the tile
      const int h = n >> 5, g = (tile_begin + i) * kGeneTile + (n & 31);
      const bool ok = n < N && g < a.G;
      float v[16], vb[16];
      tmem_ld16(tmem + lane_addr + kTmemDWO + s * kDwoCols + sub * 16, v);
      if (sub == 0) tmem_ld16(tmem + lane_addr + kTmemDWO + s * kDwoCols + 64, vb);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars[DWO_FREE + s]);
      if (ok) {
        float* dst = a.dW + ((size_t)h * a.G + g) * kK + sub * 16;
        const float sc = a.upstream * a.inv_gscale;
#pragma unroll
        for (int j = 0; j < 4; ++j) red_add_v4(dst + 4 * j, sc * v[4 * j], sc * v[4 * j + 1], sc * v[4 * j + 2], sc * v[4 * j + 3]);
        if (sub == 0) atomicAdd(a.db + (size_t)h * a.G + g, sc * vb[0]);
      }
    };

    float xnext[8];
    if (MODE != MODE_SCVI_LSE) load_x(0, xnext);
    // this thread's column of the [4 gene pairs][512 threads] count stash
    float2* xs = reinterpret_cast<float2*>(smem + OutSmem::XS(NH, TRAIN)) + t;
    const bool rows_full = row0 + kCellTile <= a.R;
    pm::F2 llk2 = pm::bc(0.f), t2 = pm::bc(0.f), dl2 = pm::bc(0.f);
    for (int i = 0; i < nt; ++i) {
      const int s = i & 1;
      const int g0 = (tile_begin + i) * kGeneTile + sub * 8;
      // no masking inside tiles that lie fully inside the matrix (all but the last cell tile / gene tile)
      const bool full = rows_full && (tile_begin + i + 1) * kGeneTile <= a.G;
      if (MODE != MODE_SCVI_LSE) {
#pragma unroll
        for (int j = 0; j < 4; ++j) xs[j * kEpiThreads] = make_float2(xnext[2 * j], xnext[2 * j + 1]);
        if (i + 1 < nt) load_x(i + 1, xnext);
      }
      mbar_wait(&bars[W_FULL + s], (i >> 1) & 1);   // bias values of this stage (bulk copy) visible to this thread
      mbar_wait(&bars[ACC_FULL + s], (i >> 1) & 1);
      if (TRAIN && i >= 2) mbar_wait(&bars[G_FREE + s], ((i >> 1) - 1) & 1);   // gradient GEMMs of tile i-2 consumed this G stage
      tc_fence_after();
      const uint32_t tb = tmem + lane_addr + (uint32_t)(s * N + sub * 8);
      const float* bias_s = reinterpret_cast<const float*>(smem + OutSmem::W0 + s * OutSmem::Wstage(NH) + 2 * w_tile_bytes(NH)) + sub * 8;
      // this thread's 16-byte slot in column group `sub` of each head of the G stage
      uint8_t* gt = smem + OutSmem::G0(NH) + s * OutSmem::Gstage + sub * 2048 + (cell >> 3) * 128 + (cell & 7) * 16;
      const size_t o = (size_t)row * a.G + g0;
      if (MODE == MODE_SCVI_LSE) {
        // running (max, sum exp) of this thread's slice of the head-0 logits; finite sentinels keep it NaN-free
        float u[8];
        tmem_ld8(tb, u);
        tmem_ld_wait();
        float tm = -3e38f;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          u[j] = (g0 + j) < a.G ? u[j] + bias_s[j] : -3e38f;
          tm = fmaxf(tm, u[j]);
        }
        if (tm > m_run) { s_run *= mufu_ex2((m_run - tm) * kLog2e); m_run = tm; }
#pragma unroll
        for (int j = 0; j < 8; ++j) s_run += mufu_ex2((u[j] - m_run) * kLog2e);
      }
      // Two genes per trip, evaluated in lock-step with packed fp32x2 arithmetic (pair_math.cuh); rolled on purpose: one
      // trip's code stays inside the instruction cache.  The TMEM reads of trip j+1 are in flight while trip j is evaluated.
      float pa[2], pb[2], pl[2];
      if (MODE != MODE_SCVI_LSE) {
        tmem_ld2(tb, pa);
        tmem_ld2(tb + 32, pb);
        if (ZI) tmem_ld2(tb + 64, pl);
      }
#pragma unroll 1
      for (int j = 0; j < (MODE == MODE_SCVI_LSE ? 0 : 8); j += 2) {
        tmem_ld_wait_tie<2>(pa); tmem_ld_tie<2>(pb);
        if (ZI) tmem_ld_tie<2>(pl);
        const float2 ba = *reinterpret_cast<const float2*>(bias_s + j), bb = *reinterpret_cast<const float2*>(bias_s + 32 + j);
        pm::F2 ra = pm::add(pm::mk(pa[0], pa[1]), pm::mk(ba.x, ba.y));
        pm::F2 rb = pm::add(pm::mk(pb[0], pb[1]), pm::mk(bb.x, bb.y));
        pm::F2 pi = pm::bc(0.f);
        if (ZI) {
          const float2 bl = *reinterpret_cast<const float2*>(bias_s + 64 + j);
          pi = pm::add(pm::mk(pl[0], pl[1]), pm::mk(bl.x, bl.y));
        }
        if (j + 2 < 8) {
          tmem_ld2(tb + j + 2, pa);
          tmem_ld2(tb + 32 + j + 2, pb);
          if (ZI) tmem_ld2(tb + 64 + j + 2, pl);
        }
        const float2 xv = xs[(j >> 1) * kEpiThreads];
        const pm::F2 x2 = pm::mk(xv.x, xv.y);
        pm::F2 llk, ga, gb, gl, mu, th;
        if (SCVI) {
          const pm::Scvi2 e = pm::elem_pair_scvi<ZI, (TRAIN || MODE == MODE_SCVI_SUMS)>(pm::sub(ra, pm::bc(lse)), rb, pi, x2, eL);
          llk = e.llk; mu = e.mu; th = e.th; gb = e.gb; gl = e.gl;
          ga = pm::mul(e.s_raw, pm::sub(e.t, pm::bc(t_row)));          // softmax Jacobian (row sum from the MODE 3 pass)
          if (MODE == MODE_SCVI_SUMS) {
            pm::F2 st = pm::mul(e.s_raw, e.t), dl = e.gmu_mu;
            if (!full) {
              const bool ok0 = row_ok && (g0 + j) < a.G, ok1 = row_ok && (g0 + j + 1) < a.G;
              st = pm::mk(ok0 ? st.x : 0.f, ok1 ? st.y : 0.f); dl = pm::mk(ok0 ? dl.x : 0.f, ok1 ? dl.y : 0.f);
            }
            t2 = pm::add(t2, st); dl2 = pm::add(dl2, dl);
          }
        } else if (FAST) {
          const pm::Elem2 e = pm::elem_pair_softplus<ZI, TRAIN>(ra, rb, pi, x2);
          llk = e.llk; ga = e.ga; gb = e.gb; gl = e.gl; mu = e.mu; th = e.th;
        } else {
          // other link functions (SURVEY.md section 8a Q1 alternatives): generic scalar evaluation
          float l_[2], ga_[2], gb_[2], gl_[2], mu_[2], th_[2];
          const float ra_[2] = {ra.x, ra.y}, rb_[2] = {rb.x, rb.y}, pi_[2] = {pi.x, pi.y}, x_[2] = {x2.x, x2.y};
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            float dmu, dth;
            activation(a.mean_act, ra_[u], mu_[u], dmu);
            activation(a.disp_act, rb_[u], th_[u], dth);
            CountGrad cg;
            cg.dmu = cg.dth = cg.dpi = 0.f;
            l_[u] = count_llk<ZI, TRAIN>(x_[u], mu_[u], th_[u], pi_[u], cg);
            ga_[u] = cg.dmu * dmu; gb_[u] = cg.dth * dth; gl_[u] = cg.dpi;
          }
          llk = pm::mk(l_[0], l_[1]); ga = pm::mk(ga_[0], ga_[1]); gb = pm::mk(gb_[0], gb_[1]); gl = pm::mk(gl_[0], gl_[1]);
          mu = pm::mk(mu_[0], mu_[1]); th = pm::mk(th_[0], th_[1]);
        }
        if (!full) {
          const bool ok0 = row_ok && (g0 + j) < a.G, ok1 = row_ok && (g0 + j + 1) < a.G;
          llk = pm::mk(ok0 ? llk.x : 0.f, ok1 ? llk.y : 0.f);
          if (TRAIN) {
            ga = pm::mk(ok0 ? ga.x : 0.f, ok1 ? ga.y : 0.f); gb = pm::mk(ok0 ? gb.x : 0.f, ok1 ? gb.y : 0.f);
            gl = pm::mk(ok0 ? gl.x : 0.f, ok1 ? gl.y : 0.f);
          } else {
            if (ok0) {
              if (a.out_mean) a.out_mean[o + j] = mu.x;
              if (a.out_disp) a.out_disp[o + j] = th.x;
              if (ZI && a.out_pi) a.out_pi[o + j] = pi.x;
            }
            if (ok1) {
              if (a.out_mean) a.out_mean[o + j + 1] = mu.y;
              if (a.out_disp) a.out_disp[o + j + 1] = th.y;
              if (ZI && a.out_pi) a.out_pi[o + j + 1] = pi.y;
            }
          }
        } else if (!TRAIN) {
          if (VEC) {       // even gene count and 16-byte aligned rows: (row * G + g0 + j) is even
            if (a.out_mean) *reinterpret_cast<float2*>(a.out_mean + o + j) = make_float2(mu.x, mu.y);
            if (a.out_disp) *reinterpret_cast<float2*>(a.out_disp + o + j) = make_float2(th.x, th.y);
            if (ZI && a.out_pi) *reinterpret_cast<float2*>(a.out_pi + o + j) = make_float2(pi.x, pi.y);
          } else {
            if (a.out_mean) { a.out_mean[o + j] = mu.x; a.out_mean[o + j + 1] = mu.y; }
            if (a.out_disp) { a.out_disp[o + j] = th.x; a.out_disp[o + j + 1] = th.y; }
            if (ZI && a.out_pi) { a.out_pi[o + j] = pi.x; a.out_pi[o + j + 1] = pi.y; }
          }
        }
        llk2 = pm::add(llk2, llk);
        if (TRAIN) {   // two genes -> one packed fp16x2 word per head
          if (a.gscale != 1.f) { ga = pm::mul(ga, pm::bc(a.gscale)); gb = pm::mul(gb, pm::bc(a.gscale)); gl = pm::mul(gl, pm::bc(a.gscale)); }
          *reinterpret_cast<__half2*>(gt + 0 * 4 * 2048 + j * 2) = __floats2half2_rn(ga.x, ga.y);
          *reinterpret_cast<__half2*>(gt + 1 * 4 * 2048 + j * 2) = __floats2half2_rn(gb.x, gb.y);
          if (ZI) *reinterpret_cast<__half2*>(gt + 2 * 4 * 2048 + j * 2) = __floats2half2_rn(gl.x, gl.y);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(&bars[ACC_FREE + s]);
        if (!TRAIN) mbar_arrive(&bars[W_FREE + s]);   // bias of this stage consumed
      }
      if (TRAIN) {
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars[G_FULL + s]);
        if (i >= 1) flush_dwo(i - 1);
      }
    }
    // per-cell log-likelihood: four gene slices per cell -> shared -> one atomic per cell and chunk
    if (MODE == MODE_SCVI_LSE) {
      if (row_ok) a.lse_part[(size_t)row * a.n_lse_parts + blockIdx.y * 4 + sub] = make_float2(m_run, s_run);
    } else if (MODE != MODE_SCVI_TRAIN) {      // the scVI training pass re-walks the tiles: llk came from MODE 3
      atomicAdd(&llk_s[cell], llk2.x + llk2.y);
      if (MODE == MODE_SCVI_SUMS) {
        atomicAdd(&llk_s[kCellTile + cell], t2.x + t2.y);
        atomicAdd(&llk_s[2 * kCellTile + cell], dl2.x + dl2.y);
      }
      named_bar_sync(1, kEpiThreads);
      if (sub == 0 && row_ok) {
        atomicAdd(a.llk_x + row, llk_s[cell]);
        if (MODE == MODE_SCVI_SUMS) {
          atomicAdd(a.Trow + row, llk_s[kCellTile + cell]);
          atomicAdd(a.dlibsum + row, llk_s[2 * kCellTile + cell]);
        }
      }
    } else if (sub == 0 && row_ok && blockIdx.y == 0) {
      a.dLib[row] = lib_open ? a.upstream * a.dlibsum[row] : 0.f;
    }
    if (TRAIN) {
      flush_dwo(nt - 1);
      mbar_wait(&bars[DD_FULL], 0);
      tc_fence_after();
      float v[16];
      tmem_ld16(tmem + lane_addr + kTmemDD + sub * 16, v);
      tmem_ld_wait();
      if (row_ok) {
        float* dst = a.dD + (size_t)row * kK + sub * 16;
        const float sc = a.upstream * a.inv_gscale;
#pragma unroll
        for (int j = 0; j < 4; ++j) red_add_v4(dst + 4 * j, sc * v[4 * j], sc * v[4 * j + 1], sc * v[4 * j + 2], sc * v[4 * j + 3]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) tmem_dealloc(tmem, kTmemCols);
}

}  // namespace tc
}  // namespace sisua
