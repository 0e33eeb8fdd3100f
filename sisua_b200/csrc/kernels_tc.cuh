// Fused tcgen05 kernel for the decoder output heads:
//   out = d . W_out^T + b   (tensor cores, accumulators in TMEM, never written to HBM)
//   -> mean / dispersion / dropout-logit activations -> ZINB / NB log-likelihood per cell (epilogue)
//   -> [train] d llk / d out as an fp16 operand tile in shared memory -> two more tcgen05 GEMMs:
//        dD    += G . W_out          (gradient wrt the decoder activations, TMEM-resident across gene tiles)
//        dW_out = G^T . [d | 1]      (weight and bias gradient of the gene tile, flushed with vector reds)
// Forward products are error-compensated 3xFP16 (d = d1 + d2, w = w1 + w2; d1w1 + d1w2 + d2w1 accumulated in
// fp32), i.e. fp32-grade logits; the two gradient GEMMs take G as one fp16 operand and both halves of W / d.
// One CTA owns a tile of 128 cells and walks a chunk of 32-gene tiles; roles: 16 epilogue warps (TMEM lane
// quarter x 8-gene slice) and a service warpgroup: forward-MMA issuer, bulk-copy (TMA) thread streaming pre-packed weight
// tiles, gradient-MMA issuer.  The counts arrive as fp32 or (XU16) uint16 and are widened in the epilogue.
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
// 16 epilogue warps + one service warpgroup (MMA warp, loader warp, two idle warps).  The service warps sit in a warpgroup of
// their own so that `setmaxnreg` can move registers from them to the epilogue warpgroups: the block is launched with 96
// registers per thread (the most 20 warps can get), the service warpgroup drops to 32 and the four epilogue warpgroups
// grow to 112 -- enough to keep two gene pairs in flight per thread.
constexpr int kOutThreads = kEpiThreads + 128;
constexpr int kOutRegsService = 32, kOutRegsEpilogue = 112;
constexpr int kMmaWarp = kEpiWarps, kLoadWarp = kEpiWarps + 1, kGradWarp = kEpiWarps + 2;

// Development aid (-DSISUA_OUT_TRACE): CTA 0 stamps clock64() at the hand-over points of every role for each gene tile;
// tools/trace_out_heads.py prints the timeline.  Compiled out of the product.
#ifdef SISUA_OUT_TRACE
__device__ long long g_out_trace[4][64][8];
#define OUT_TR(role, i, k) do { if (blockIdx.x == 0 && (i) < 64) g_out_trace[role][i][k] = clock64(); } while (0)
#else
#define OUT_TR(role, i, k) do { } while (0)
#endif
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
  const float* x;          // [B, G] counts, or the resident [N, G] matrix when ridx is given
  const int* ridx;         // [B] rows of x that make up this minibatch (nullable: rows 0 .. B-1)
  const uint8_t* packed;   // pre-packed weight tiles
  float* llk_x;            // [R], zeroed by the caller (gene chunks add atomically)
  // inference outputs (nullable)
  float* out_mean; float* out_disp; float* out_pi;
  float* out_mean_avg;     // [B, G] += mean / S  (mean over the Monte-Carlo samples of the NB mean; zeroed by the caller)
  float inv_S;
  int nozi;                // inference: likelihood of the count distribution WITHOUT zero inflation ("imputed", posterior.py:210-220)
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

// Weight tiles have THREE stages: a stage is released by the gradient GEMMs of its tile, i.e. only after the slowest
// epilogue warp has finished that tile; with two stages the next-but-one tile's forward product (and with it every
// faster warp) waited for that, which made each tile a barrier across the 16 warps.
constexpr int kWStages = 3;
struct OutSmem {     // offsets into dynamic shared memory (bytes)
  static constexpr int dA1 = 0;                          // [128][80] fp16 (cols 64.. = ones / zero pad, train)
  static constexpr int dA2 = dA1 + 10 * 2048;            // [128][64] fp16
  static constexpr int W0 = dA2 + 8 * 2048;              // 2 stages of packed tiles
  __host__ __device__ static constexpr int Wstage(int nh) { return packed_tile_stride(nh); }
  __host__ __device__ static constexpr int G0(int nh) { return W0 + kWStages * Wstage(nh); }       // [128][128] fp16 x 2 stages (train)
  static constexpr int Gstage = 16 * 2048;                                                 // two G stages when training
  __host__ __device__ static constexpr int XS(int nh, bool train) { return G0(nh) + (train ? 2 * Gstage : 0); }   // [8][512] count stash
  __host__ __device__ static constexpr int LLK(int nh, bool train) { return XS(nh, train) + 2 * 8 * kEpiThreads * 4; }   // two stash stages
  __host__ __device__ static constexpr int BAR(int nh, bool train) { return LLK(nh, train) + (3 * kCellTile + 2 * kK) * 4; }   // llk | T | dlib | norm scale, shift
  __host__ __device__ static constexpr int total(int nh, bool train) { return BAR(nh, train) + 32 * 8 + 16; }
};

enum OutBar { ACC_FULL = 4, ACC_FREE = 6, G_FULL = 8, G_FREE = 10, DWO_FULL = 12, DWO_FREE = 14, DD_FULL = 16, W_FULL = 17, W_FREE = 20, NUM_BARS = 23 };

// LINK: how the raw head outputs become (mean, inverse dispersion): the default softplus pair (packed fast path), TFP's
// total_count / logits form of the 'zinb' / 'nb' enums (packed), or any other activation pair (scalar evaluation)
enum OutLink { LINK_GENERIC = 0, LINK_SOFTPLUS = 1, LINK_TFP = 2 };
// XU16 (VEC only): a.x points at uint16 counts (exact for count data; half the bytes of the likelihood's second read of x)
template <int NH, bool TRAIN, bool VEC, int LINK, int MODE = MODE_PLAIN, bool XU16 = false>
__global__ void __launch_bounds__(kOutThreads, 1) out_heads_kernel(OutHeadsArgs a) {
  static_assert(TRAIN == (MODE == MODE_SCVI_TRAIN) || MODE == MODE_PLAIN, "only the plain and scVI-train modes run the gradient GEMMs");
  static_assert(VEC || !XU16, "uint16 counts are implemented for the 16-byte aligned (VEC) geometry");
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
    // a weight stage is released by the forward issuer's commit plus, when training, the gradient issuer's commit; without
    // a backward pass, by every epilogue warp once it has read the stage's bias values
    for (int ws = 0; ws < kWStages; ++ws) { mbar_init(&bars[W_FULL + ws], 1); mbar_init(&bars[W_FREE + ws], TRAIN ? 2 : 1 + kEpiWarps); }
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

  // (each role branch starts with its own setmaxnreg and the branches only meet again at the end of the kernel: that is
  // what lets ptxas give the epilogue code the larger budget)
  if (warp > kGradWarp) {
    setmaxnreg_dec<kOutRegsService>();      // idle warp of the service warpgroup
  } else if (warp == kLoadWarp) {
    setmaxnreg_dec<kOutRegsService>();
    // =========================== loader: pre-packed weight tiles via 1-D TMA bulk copies ===========================
    if (lane == 0) {
      for (int i = 0; i < nt; ++i) {
        const int ws = i % kWStages;
        if (i >= kWStages) mbar_wait_backoff(&bars[W_FREE + ws], ((i / kWStages) - 1) & 1);
        mbar_arrive_expect_tx(&bars[W_FULL + ws], packed_tile_bytes(NH));
        bulk_copy_g2s(smem + OutSmem::W0 + ws * OutSmem::Wstage(NH),
                      a.packed + (size_t)(tile_begin + i) * packed_tile_stride(NH), packed_tile_bytes(NH), &bars[W_FULL + ws]);
      }
    }
  } else if (warp == kMmaWarp) {
    setmaxnreg_dec<kOutRegsService>();
    // =========================== forward MMA issuer (one thread) ===========================
    // Its own thread, so that the product of tile i+2 is issued as soon as ITS inputs are there (weight tile landed,
    // accumulator stage handed back in the middle of tile i) and never queues behind the gradient GEMMs of tile i, which
    // wait for the slowest epilogue warp.
    if (lane == 0) {
      const uint32_t idesc_fwd = make_idesc_f16(kCellTile, NF, 0, 0);
      const uint32_t sA1 = smem_u32(smem + OutSmem::dA1), sA2 = smem_u32(smem + OutSmem::dA2);
      for (int i = 0; i < nt; ++i) {
        const int s = i & 1, ws = i % kWStages;
        const uint32_t sW1 = smem_u32(smem + OutSmem::W0 + ws * OutSmem::Wstage(NH)), sW2 = sW1 + w_tile_bytes(NH);
        OUT_TR(1, i, 0);
        mbar_wait_backoff(&bars[W_FULL + ws], (i / kWStages) & 1);
        OUT_TR(1, i, 1);
        if (i >= 2) mbar_wait_backoff(&bars[ACC_FREE + s], ((i >> 1) - 1) & 1);
        OUT_TR(1, i, 2);
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
        umma_commit(&bars[W_FREE + ws]);      // (training: the gradient issuer's commit is the stage's second arrival)
        OUT_TR(1, i, 3);
      }
    }
  } else if (warp == kGradWarp) {
    setmaxnreg_dec<kOutRegsService>();
    // =========================== gradient MMA issuer (one thread) ===========================
    if (TRAIN && lane == 0) {
      const uint32_t idesc_dd = make_idesc_f16(kCellTile, kK, 0, 1);
      const uint32_t idesc_dwo = make_idesc_f16(kCellTile, kDwoCols, 1, 1);
      const uint32_t idesc_dwo_lo = make_idesc_f16(kCellTile, kK, 1, 1);
      const uint32_t sA1 = smem_u32(smem + OutSmem::dA1), sA2 = smem_u32(smem + OutSmem::dA2);
      const uint32_t sG0 = smem_u32(smem + OutSmem::G0(NH));
      for (int i = 0; i < nt; ++i) {
        const int s = i & 1, ws = i % kWStages;
        const uint32_t sW1 = smem_u32(smem + OutSmem::W0 + ws * OutSmem::Wstage(NH));
        const uint32_t sG = sG0 + s * OutSmem::Gstage;
        OUT_TR(2, i, 0);
        mbar_wait_backoff(&bars[G_FULL + s], (i >> 1) & 1);
        OUT_TR(2, i, 1);
        if (i >= 2) mbar_wait_backoff(&bars[DWO_FREE + s], ((i >> 1) - 1) & 1);
        OUT_TR(2, i, 2);
        tc_fence_after();
        // dD[cells, k] += G[cells, n] . W[n, k]   (A: G K-major, B: w1 then w2, MN-major).  Both halves of the weight
        // split are used: the fp16 rounding of a weight is the same for every cell, so with w1 alone the error of dD
        // is correlated across the batch and survives the sums of the backward pass below; the rounding of G is
        // independent per entry and averages out.
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
        umma_commit(&bars[W_FREE + ws]);
        if (i == nt - 1) umma_commit(&bars[DD_FULL]);
        OUT_TR(2, i, 3);
      }
    }
  } else {
    // =========================== epilogue warps ===========================
    setmaxnreg_inc<kOutRegsEpilogue>();
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
    const int xb = (row_ok ? row : 0) % a.B;
    const float* xrow = a.x + (size_t)(a.ridx ? a.ridx[xb] : xb) * a.G;
    auto load_x = [&](int i, float* xv) {
      const int g = (tile_begin + i) * kGeneTile + sub * 8;
      if (VEC) {
        float4 p0 = make_float4(0.f, 0.f, 0.f, 0.f), p1 = p0;
        if (row_ok && g < a.G && a.x) p0 = __ldg(reinterpret_cast<const float4*>(xrow + g));
        if (row_ok && g + 4 < a.G && a.x) p1 = __ldg(reinterpret_cast<const float4*>(xrow + g + 4));
        xv[0] = p0.x; xv[1] = p0.y; xv[2] = p0.z; xv[3] = p0.w; xv[4] = p1.x; xv[5] = p1.y; xv[6] = p1.z; xv[7] = p1.w;
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) xv[j] = (row_ok && g + j < a.G && a.x) ? __ldg(xrow + g + j) : 0.f;
      }
    };

    auto flush_dwo = [&](int i) {            // tile i's weight / bias gradient: TMEM -> vector reds
      const int s = i & 1;
      mbar_wait_backoff(&bars[DWO_FULL + s], (i >> 1) & 1);
      if (t == 0) OUT_TR(0, i + 2, 7);       // (stamped in the row of the tile whose trip does this flush)
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

    // Count stash: this thread's 8 counts of a tile as two 16-byte slots ([2 stages][2 halves][512 threads] x 16 B).  With
    // 16-byte aligned rows (VEC) the slots of tile i+1 are filled by cp.async while tile i is evaluated: no registers
    // hold prefetched counts and nobody but the owning thread touches a slot, so no barrier is needed; otherwise the
    // counts travel through registers.
    uint8_t* xs_base = smem + OutSmem::XS(NH, TRAIN) + t * 16;
    constexpr int kXsHalf = kEpiThreads * 16, kXsStage = 2 * kXsHalf;
    auto prefetch_x = [&](int i) {          // VEC only
      const int g = (tile_begin + i) * kGeneTile + sub * 8;
      uint8_t* dst = xs_base + (i & 1) * kXsStage;
      const bool ok0 = row_ok && g < a.G && a.x, ok1 = row_ok && g + 4 < a.G && a.x;      // (x == NULL: decode only, zeros)
      if (XU16) {          // the 8 counts of this thread are ONE 16-byte copy (G is a multiple of 8: whole or nothing)
        const uint16_t* xrow16 = reinterpret_cast<const uint16_t*>(a.x) + (size_t)(a.ridx ? a.ridx[xb] : xb) * a.G;
        cp_async_16_zfill(dst, ok0 ? (const void*)(xrow16 + g) : (const void*)a.x, ok0 ? 16u : 0u);
      } else {
        cp_async_16_zfill(dst, ok0 ? (const void*)(xrow + g) : (const void*)a.x, ok0 ? 16u : 0u);
        cp_async_16_zfill(dst + kXsHalf, ok1 ? (const void*)(xrow + g + 4) : (const void*)a.x, ok1 ? 16u : 0u);
      }
      cp_async_commit();
    };
    float xnext[VEC ? 1 : 8];
    if (MODE != MODE_SCVI_LSE) {
      if (VEC) prefetch_x(0); else load_x(0, xnext);
    }
    const bool rows_full = row0 + kCellTile <= a.R;
    pm::F2 llk2 = pm::bc(0.f), t2 = pm::bc(0.f), dl2 = pm::bc(0.f);
    for (int i = 0; i < nt; ++i) {
      const int s = i & 1;
      const int g0 = (tile_begin + i) * kGeneTile + sub * 8;
      // no masking inside tiles that lie fully inside the matrix (all but the last cell tile / gene tile)
      const bool full = rows_full && (tile_begin + i + 1) * kGeneTile <= a.G;
      const uint8_t* xs = xs_base + (i & 1) * kXsStage;
      if (MODE != MODE_SCVI_LSE) {
        if (VEC) {
          cp_async_wait<0>();               // this tile's counts have landed (issued one tile ago)
          if (i + 1 < nt) prefetch_x(i + 1);
        } else {
          *reinterpret_cast<float4*>(xs_base + (i & 1) * kXsStage) = make_float4(xnext[0], xnext[1], xnext[2], xnext[3]);
          *reinterpret_cast<float4*>(xs_base + (i & 1) * kXsStage + kXsHalf) = make_float4(xnext[VEC ? 0 : 4], xnext[VEC ? 0 : 5], xnext[VEC ? 0 : 6], xnext[VEC ? 0 : 7]);
          if (i + 1 < nt) load_x(i + 1, xnext);
        }
      }
      const int ws = i % kWStages;
      const int tr_role = (t == 0) ? 0 : 3;
      const bool tr_on = (t == 0) || (t == kEpiThreads - 32);
      if (tr_on) OUT_TR(tr_role, i, 0);
      mbar_wait_backoff(&bars[W_FULL + ws], (i / kWStages) & 1);   // bias values of this stage (bulk copy) visible to this thread
      if (tr_on) OUT_TR(tr_role, i, 1);
      mbar_wait_backoff(&bars[ACC_FULL + s], (i >> 1) & 1);
      if (tr_on) OUT_TR(tr_role, i, 2);
      if (TRAIN && i >= 2) mbar_wait_backoff(&bars[G_FREE + s], ((i >> 1) - 1) & 1);   // gradient GEMMs of tile i-2 consumed this G stage
      if (tr_on) OUT_TR(tr_role, i, 3);
      tc_fence_after();
      const uint32_t tb = tmem + lane_addr + (uint32_t)(s * N + sub * 8);
      const float* bias_s = reinterpret_cast<const float*>(smem + OutSmem::W0 + ws * OutSmem::Wstage(NH) + 2 * w_tile_bytes(NH)) + sub * 8;
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
      // kNP gene pairs per trip: each pair advances in lock-step through packed fp32x2 arithmetic (pair_math.cuh) and the
      // pairs are independent dependency chains the compiler interleaves; the trip loop is rolled on purpose (its code
      // stays inside the instruction cache).  The TMEM reads of trip j+1 are in flight while trip j is evaluated.
#ifndef SISUA_OUT_NP
#define SISUA_OUT_NP 2       // measured on the B200 (18 944 x 2 000, ZINB train): 1 pair per trip 0.303 ms, 2 pairs 0.290 ms
#endif
      constexpr int kNP = SISUA_OUT_NP, kGT = 2 * kNP;     // genes per trip
      float pa[kGT], pb[kGT], pl[kGT];
      if (MODE != MODE_SCVI_LSE) {
        tmem_ldn<kGT>(tb, pa);
        tmem_ldn<kGT>(tb + 32, pb);
        if (ZI) tmem_ldn<kGT>(tb + 64, pl);
      }
#pragma unroll 1
      for (int j = 0; j < (MODE == MODE_SCVI_LSE ? 0 : 8); j += kGT) {
        tmem_ld_wait_tie<kGT>(pa); tmem_ld_tie<kGT>(pb);
        if (ZI) tmem_ld_tie<kGT>(pl);
        if (j + kGT >= 8) {       // the last accumulator columns of this tile are in registers: hand the TMEM stage back now, so
          tc_fence_before();      // the forward product of tile i+2 can start while this trip is still being evaluated
          __syncwarp();
          if (lane == 0) mbar_arrive(&bars[ACC_FREE + s]);
        }
        pm::F2 ra[kNP], rb[kNP], pi[kNP], x2[kNP];
#pragma unroll
        for (int p = 0; p < kNP; ++p) {
          const float2 ba = *reinterpret_cast<const float2*>(bias_s + j + 2 * p), bb = *reinterpret_cast<const float2*>(bias_s + 32 + j + 2 * p);
          ra[p] = pm::add(pm::mk(pa[2 * p], pa[2 * p + 1]), pm::mk(ba.x, ba.y));
          rb[p] = pm::add(pm::mk(pb[2 * p], pb[2 * p + 1]), pm::mk(bb.x, bb.y));
          pi[p] = pm::bc(0.f);
          if (ZI) {
            const float2 bl = *reinterpret_cast<const float2*>(bias_s + 64 + j + 2 * p);
            pi[p] = pm::add(pm::mk(pl[2 * p], pl[2 * p + 1]), pm::mk(bl.x, bl.y));
          }
          const int pj = (j >> 1) + p;      // gene pair 0..3 of this thread's slice
          if (XU16) {
            const uint32_t w = *reinterpret_cast<const uint32_t*>(xs + pj * 4);
            x2[p] = pm::mk(u16_to_float(w & 0xffffu), u16_to_float(w >> 16));
          } else {
            const float2 xv = *reinterpret_cast<const float2*>(xs + (pj >> 1) * kXsHalf + (pj & 1) * 8);
            x2[p] = pm::mk(xv.x, xv.y);
          }
        }
        if (j + kGT < 8) {
          tmem_ldn<kGT>(tb + j + kGT, pa);
          tmem_ldn<kGT>(tb + 32 + j + kGT, pb);
          if (ZI) tmem_ldn<kGT>(tb + 64 + j + kGT, pl);
        }
        pm::F2 llk[kNP], ga[kNP], gb[kNP], gl[kNP], mu[kNP], th[kNP];
        if (SCVI) {
          pm::Scvi2 e[kNP];
          pm::F2 ul[kNP];
#pragma unroll
          for (int p = 0; p < kNP; ++p) ul[p] = pm::sub(ra[p], pm::bc(lse));
          pm::elem_multi_scvi<ZI, (TRAIN || MODE == MODE_SCVI_SUMS), kNP>(ul, rb, pi, x2, eL, e, (TRAIN || MODE == MODE_SCVI_SUMS) ? false : a.nozi != 0);
#pragma unroll
          for (int p = 0; p < kNP; ++p) {
            llk[p] = e[p].llk; mu[p] = e[p].mu; th[p] = e[p].th; gb[p] = e[p].gb; gl[p] = e[p].gl;
            ga[p] = pm::mul(e[p].s_raw, pm::sub(e[p].t, pm::bc(t_row)));          // softmax Jacobian (row sum from the MODE 3 pass)
            if (MODE == MODE_SCVI_SUMS) {
              pm::F2 st = pm::mul(e[p].s_raw, e[p].t), dl = e[p].gmu_mu;
              if (!full) {
                const bool ok0 = row_ok && (g0 + j + 2 * p) < a.G, ok1 = row_ok && (g0 + j + 2 * p + 1) < a.G;
                st = pm::mk(ok0 ? st.x : 0.f, ok1 ? st.y : 0.f); dl = pm::mk(ok0 ? dl.x : 0.f, ok1 ? dl.y : 0.f);
              }
              t2 = pm::add(t2, st); dl2 = pm::add(dl2, dl);
            }
          }
        } else if (LINK != LINK_GENERIC) {
          pm::Elem2 e[kNP];
          if (LINK == LINK_TFP) pm::elem_multi_tfp<ZI, TRAIN, kNP>(ra, rb, pi, x2, e, TRAIN ? false : a.nozi != 0);
          else pm::elem_multi_softplus<ZI, TRAIN, kNP>(ra, rb, pi, x2, e, TRAIN ? false : a.nozi != 0);
#pragma unroll
          for (int p = 0; p < kNP; ++p) { llk[p] = e[p].llk; ga[p] = e[p].ga; gb[p] = e[p].gb; gl[p] = e[p].gl; mu[p] = e[p].mu; th[p] = e[p].th; }
        } else {
          // other link functions (SURVEY.md section 8a Q1 alternatives): generic scalar evaluation
#pragma unroll
          for (int p = 0; p < kNP; ++p) {
            float l_[2], ga_[2], gb_[2], gl_[2], mu_[2], th_[2];
            const float ra_[2] = {ra[p].x, ra[p].y}, rb_[2] = {rb[p].x, rb[p].y}, pi_[2] = {pi[p].x, pi[p].y}, x_[2] = {x2[p].x, x2[p].y};
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
            llk[p] = pm::mk(l_[0], l_[1]); ga[p] = pm::mk(ga_[0], ga_[1]); gb[p] = pm::mk(gb_[0], gb_[1]); gl[p] = pm::mk(gl_[0], gl_[1]);
            mu[p] = pm::mk(mu_[0], mu_[1]); th[p] = pm::mk(th_[0], th_[1]);
          }
        }
#pragma unroll
        for (int p = 0; p < kNP; ++p) {
          const int jj = j + 2 * p;
          if (!full) {
            const bool ok0 = row_ok && (g0 + jj) < a.G, ok1 = row_ok && (g0 + jj + 1) < a.G;
            llk[p] = pm::mk(ok0 ? llk[p].x : 0.f, ok1 ? llk[p].y : 0.f);
            if (TRAIN) {
              ga[p] = pm::mk(ok0 ? ga[p].x : 0.f, ok1 ? ga[p].y : 0.f); gb[p] = pm::mk(ok0 ? gb[p].x : 0.f, ok1 ? gb[p].y : 0.f);
              gl[p] = pm::mk(ok0 ? gl[p].x : 0.f, ok1 ? gl[p].y : 0.f);
            } else {
              const size_t oa = (size_t)(row % a.B) * a.G + g0;
              if (ok0 && a.out_mean_avg) atomicAdd(a.out_mean_avg + oa + jj, mu[p].x * a.inv_S);
              if (ok1 && a.out_mean_avg) atomicAdd(a.out_mean_avg + oa + jj + 1, mu[p].y * a.inv_S);
              if (ok0) {
                if (a.out_mean) a.out_mean[o + jj] = mu[p].x;
                if (a.out_disp) a.out_disp[o + jj] = th[p].x;
                if (ZI && a.out_pi) a.out_pi[o + jj] = pi[p].x;
              }
              if (ok1) {
                if (a.out_mean) a.out_mean[o + jj + 1] = mu[p].y;
                if (a.out_disp) a.out_disp[o + jj + 1] = th[p].y;
                if (ZI && a.out_pi) a.out_pi[o + jj + 1] = pi[p].y;
              }
            }
          } else if (!TRAIN) {
            if (a.out_mean_avg) {
              const size_t oa = (size_t)(row % a.B) * a.G + g0;
              atomicAdd(a.out_mean_avg + oa + jj, mu[p].x * a.inv_S);
              atomicAdd(a.out_mean_avg + oa + jj + 1, mu[p].y * a.inv_S);
            }
            if (VEC) {       // even gene count and 16-byte aligned rows: (row * G + g0 + jj) is even
              if (a.out_mean) *reinterpret_cast<float2*>(a.out_mean + o + jj) = make_float2(mu[p].x, mu[p].y);
              if (a.out_disp) *reinterpret_cast<float2*>(a.out_disp + o + jj) = make_float2(th[p].x, th[p].y);
              if (ZI && a.out_pi) *reinterpret_cast<float2*>(a.out_pi + o + jj) = make_float2(pi[p].x, pi[p].y);
            } else {
              if (a.out_mean) { a.out_mean[o + jj] = mu[p].x; a.out_mean[o + jj + 1] = mu[p].y; }
              if (a.out_disp) { a.out_disp[o + jj] = th[p].x; a.out_disp[o + jj + 1] = th[p].y; }
              if (ZI && a.out_pi) { a.out_pi[o + jj] = pi[p].x; a.out_pi[o + jj + 1] = pi[p].y; }
            }
          }
          llk2 = pm::add(llk2, llk[p]);
          if (TRAIN) {   // two genes -> one packed fp16x2 word per head
            if (a.gscale != 1.f) { ga[p] = pm::mul(ga[p], pm::bc(a.gscale)); gb[p] = pm::mul(gb[p], pm::bc(a.gscale)); gl[p] = pm::mul(gl[p], pm::bc(a.gscale)); }
            *reinterpret_cast<__half2*>(gt + 0 * 4 * 2048 + jj * 2) = __floats2half2_rn(ga[p].x, ga[p].y);
            *reinterpret_cast<__half2*>(gt + 1 * 4 * 2048 + jj * 2) = __floats2half2_rn(gb[p].x, gb[p].y);
            if (ZI) *reinterpret_cast<__half2*>(gt + 2 * 4 * 2048 + jj * 2) = __floats2half2_rn(gl[p].x, gl[p].y);
          }
        }
      }
      if (tr_on) OUT_TR(tr_role, i, 4);      // element math of this tile done
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (MODE == MODE_SCVI_LSE) mbar_arrive(&bars[ACC_FREE + s]);     // (the other modes released it inside the trip loop)
        if (!TRAIN) mbar_arrive(&bars[W_FREE + ws]);   // bias of this stage consumed
      }
      if (TRAIN) {
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars[G_FULL + s]);
        if (tr_on) OUT_TR(tr_role, i, 5);
        // the weight-gradient tile of tile i-2 (not i-1): its MMAs were committed together with the release of the G stage
        // this warp waited for at the top of tile i, so the flush never waits for the slowest warp of the CTA (flushing
        // tile i-1 here made every tile a barrier across the 16 epilogue warps: 9 % of the kernel's samples were polls)
        if (i >= 2) flush_dwo(i - 2);
        if (tr_on) OUT_TR(tr_role, i, 6);
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
      if (nt >= 2) flush_dwo(nt - 2);
      flush_dwo(nt - 1);
      mbar_wait_backoff(&bars[DD_FULL], 0);
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
