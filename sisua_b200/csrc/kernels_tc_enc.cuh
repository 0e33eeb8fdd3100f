// tcgen05 kernels for the genes x hidden contraction of the first encoder layer:
//   forward   A0[cells, N0] = dropout(log1p(x))[cells, G] . W1[N0, G]^T        (N0 = 64, or 128 for scVI's two encoders)
//   backward  dW1[N0, G]   = delta1[cells, N0]^T . dropout(log1p(x))[cells, G]
// The forward streams the count matrix once from HBM (4 G bytes per cell against 2*G*N0 flops): converter warps load
// coalesced fp32 rows, apply log1p (+ Philox input dropout), split to fp16 (hi, lo) and write canonical no-swizzle UMMA
// tiles; one thread issues the MMAs, accumulators stay in TMEM for the whole K range; the hi tiles are also written
// back to HBM with TMA bulk stores, and the backward re-loads them with TMA (read MN-major) instead of converting the
// counts a second time.  Forward products are 3xFP16 compensated (fp32-grade pre-activations);
// the weight gradient uses single fp16 operands with delta1 pre-scaled by the batch size to stay in range.
#pragma once
#include "device_math.cuh"
#include "tc_ptx.cuh"

namespace sisua {
namespace tc {

constexpr int kEncFwdConvWarps = 16;    // forward: converter / epilogue warps (the conversion, not HBM, is what limits it)
constexpr int kEncFwdThreads = (kEncFwdConvWarps + 3) * 32;   // + MMA warp + weight-loader warp + tile-store warp
constexpr int kEncThreads = 320;        // backward: 8 delta-converter / epilogue warps + MMA warp + TMA warp
constexpr int kEncStages = 3;
constexpr int kPadCS = 2064;            // column-group stride of thread-written tiles: 128 rows * 16 B + 16 B (bank spread)
constexpr int kXtTile = 8 * kPadCS;     // bytes of one normalised fp16 count tile [128 cells][64 genes] as kept for the backward

__host__ __device__ constexpr int w1_tile_bytes(int n0) { return n0 * 64 * 2; }          // one fp16 copy of a 64-gene k-block
__host__ __device__ constexpr int w1_block_bytes(int n0) { return 2 * w1_tile_bytes(n0); }   // hi | lo

// W1[N0, Gp] fp32 -> per 64-gene k-block (hi | lo) fp16 tiles T[n][k], RS = 128, CS = N0/8*128
__device__ __forceinline__ void pack_w1_block(const float* __restrict__ W, int ldw, uint8_t* __restrict__ packed, int G, int n0,
                                              int kb) {
  const int CS = n0 / 8 * 128;
  uint8_t* base = packed + (size_t)kb * w1_block_bytes(n0);
  for (int i = threadIdx.x; i < n0 * 8; i += blockDim.x) {
    int n = i % n0, cg = i / n0;
    uint32_t hi[4], lo[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int g = kb * 64 + cg * 8 + 2 * j;
      float v0 = g < G ? W[(size_t)n * ldw + g] : 0.f;
      float v1 = g + 1 < G ? W[(size_t)n * ldw + g + 1] : 0.f;
      split_f16x2(v0, v1, hi[j], lo[j]);
    }
    uint32_t off = (n >> 3) * 128 + cg * CS + (n & 7) * 16;
    *reinterpret_cast<uint4*>(base + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    *reinterpret_cast<uint4*>(base + w1_tile_bytes(n0) + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
  }
}
// one launch per step re-packs both tensor-core weight operands after the optimiser moved them: blocks [0, n_kblocks)
// the first-layer k-blocks, the rest the output-head gene tiles (W_out == nullptr: first layer only)
__global__ void __launch_bounds__(256) pack_weights_kernel(const float* __restrict__ W1, int ldw, uint8_t* __restrict__ packed_w1,
                                                           int G, int n0, int n_kblocks, const float* __restrict__ W_out,
                                                           const float* __restrict__ b_out, uint8_t* __restrict__ packed_wout,
                                                           int nh) {
  if ((int)blockIdx.x < n_kblocks) pack_w1_block(W1, ldw, packed_w1, G, n0, blockIdx.x);
  else pack_wout_tile(W_out, b_out, packed_wout, G, nh, blockIdx.x - n_kblocks);
}

// log(1 + x): |error| <= 1 ulp of (1 + x) in absolute terms, far below the 2^-22 of the hi/lo operand split
__device__ __forceinline__ float log1p_count(float x) { return kLn2 * mufu_lg2(1.f + x); }

// 8 consecutive columns c0..c0+7 of row r of X[rows, ld]; zero outside [0,rows) x [0,cols)
template <bool VEC>
__device__ __forceinline__ void load8(const float* __restrict__ X, int ld, int rows, int cols, int r, int c0, float* v) {
  if (VEC) {
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
    if (r < rows) {
      // volatile asm keeps the loads where they are written: the software prefetch one block ahead must not be
      // sunk next to the first use by the compiler
      const float* p = X + (size_t)r * ld + c0;
      if (c0 < cols) asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w) : "l"(p));
      if (c0 + 4 < cols) asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w) : "l"(p + 4));
    }
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  } else {
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = (r < rows && c0 + j < cols) ? __ldg(X + (size_t)r * ld + c0 + j) : 0.f;
  }
}

// log1p + dropout on 8 count values of (row, c0..c0+7), in place (`step`: dropout_step(drop), read once per thread)
__device__ __forceinline__ void normalise8(float* v, int log_norm, const DropSpec& drop, uint32_t step, int row, int c0) {
  if (log_norm) {
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = log1p_count(v[j]);
  }
  if (drop.rate > 0.f) {     // c0 is a multiple of 8: one Philox call masks the whole group
    float m[8];
    dropout_mult8_inline(drop, step, (uint32_t)row, (uint32_t)(c0 >> 3), m);
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] *= m[j];
  }
}

__device__ __forceinline__ void store8_hi_lo(uint8_t* hi_tile, uint8_t* lo_tile, uint32_t off, const float* v) {
  uint32_t hi[4], lo[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) split_f16x2(v[2 * j], v[2 * j + 1], hi[j], lo[j]);
  *reinterpret_cast<uint4*>(hi_tile + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
  if (lo_tile) *reinterpret_cast<uint4*>(lo_tile + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
}
__device__ __forceinline__ void store8_hi(uint8_t* tile, uint32_t off, const float* v, float scale) {
  uint32_t hi[4];
#pragma unroll
  for (int j = 0; j < 4; ++j)
  {
    const __half2 h = __floats2half2_rn(fminf(fmaxf(v[2 * j] * scale, -60000.f), 60000.f),
                                        fminf(fmaxf(v[2 * j + 1] * scale, -60000.f), 60000.f));
    hi[j] = *reinterpret_cast<const uint32_t*>(&h);
  }
  *reinterpret_cast<uint4*>(tile + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
}

struct EncFwdArgs {
  const float* x;           // [B, G], or the resident [N, G] matrix when ridx is given (uint16 entries in the XU16 kernels)
  const int* ridx;          // [B] rows of x that make up this minibatch (nullable: rows 0 .. B-1)
  const uint8_t* packed;    // packed W1 k-blocks
  float* A0;                // [B, ld0] pre-activations (zeroed by the caller when k_chunks > 1)
  int B, G, ld0, n_kblocks, kblocks_per_chunk, atomic_out, log_norm;
  DropSpec drop;
  uint8_t* xt;              // [cell_tiles][xt_kblocks] tiles of dropout(log1p(x)) in fp16, exactly as the MMA consumed them
  int xt_kblocks;           // (training) kept for the weight-gradient kernel so the counts are not converted twice
};

constexpr int kRawTile = 128 * 64 * 4;   // one raw fp32 count tile [128 cells][64 genes] as the bulk copies land it

// VEC (row pitch and base 16-byte aligned): the counts arrive by 16-byte async copies (cp.async) into raw fp32 stages,
// issued by the loader warp several k-blocks ahead, so no converter warp ever waits on a global load (one bulk copy per
// 256-byte row segment was tried first: the TMA unit's per-request cost made it 2x slower); otherwise the converter
// warps load through registers.
template <int N0, bool VEC>
struct EncFwdSmem {
  static constexpr int kStages = VEC ? 2 : 3;                // operand stages (hi | lo | weights)
  static constexpr int kRaw = VEC ? (N0 == 64 ? 3 : 2) : 0;  // raw stages
  static constexpr int A = 8 * kPadCS;                       // one fp16 tile [128][64]
  static constexpr int stage = 2 * A + w1_block_bytes(N0);
  static constexpr int raw = kStages * stage;
  static constexpr int bar = raw + kRaw * kRawTile;
  static constexpr int total = bar + 16 * 8 + 16;
};

enum EncBar { EB_A_FULL = 0, EB_W_FULL = 3, EB_STAGE_FREE = 6, EB_ACC_FULL = 9, EB_RAW_FULL = 10, EB_RAW_FREE = 13 };

// XU16 (VEC only): the count matrix is stored as uint16 (exact for count data, half the HBM bytes); a row segment of 64
// genes is 128 bytes = 8 async copies, and an item's 8 counts are ONE 16-byte shared-memory read.
template <int N0, bool VEC, bool XU16 = false>
__global__ void __launch_bounds__(kEncFwdThreads, 1) enc_first_fwd_kernel(EncFwdArgs a) {
  static_assert(VEC || !XU16, "uint16 counts are implemented for the 16-byte aligned (VEC) geometry");
  constexpr int CW = kEncFwdConvWarps, kMma = CW, kLoad = CW + 1, kStore = CW + 2;
  extern __shared__ __align__(128) uint8_t smem[];
  using S = EncFwdSmem<N0, VEC>;
  constexpr int NS = S::kStages, NR = S::kRaw > 0 ? S::kRaw : 1;
  constexpr int W_CS = N0 / 8 * 128;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S::bar);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + S::bar + 16 * 8);
  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  const int row0 = blockIdx.x * 128;
  const int kb_begin = blockIdx.y * a.kblocks_per_chunk;
  const int nkb = min(a.n_kblocks, kb_begin + a.kblocks_per_chunk) - kb_begin;
  if (nkb <= 0) return;
  if (t == 0) {
    for (int s = 0; s < 3; ++s) {
      mbar_init(&bars[EB_A_FULL + s], CW); mbar_init(&bars[EB_W_FULL + s], 1); mbar_init(&bars[EB_STAGE_FREE + s], a.xt ? 2 : 1);
      mbar_init(&bars[EB_RAW_FULL + s], CW * 32); mbar_init(&bars[EB_RAW_FREE + s], CW);
    }
    mbar_init(&bars[EB_ACC_FULL], 1);
    fence_barrier_init();
  }
  if (warp == kMma) tmem_alloc(tmem_slot, N0 < 32 ? 32 : N0);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == kLoad) {
    if (lane == 0) {
      for (int i = 0; i < nkb; ++i) {
        const int s = i % NS;
        if (i >= NS) mbar_wait(&bars[EB_STAGE_FREE + s], ((i / NS) - 1) & 1);
        mbar_arrive_expect_tx(&bars[EB_W_FULL + s], w1_block_bytes(N0));
        bulk_copy_g2s(smem + s * S::stage + 2 * S::A, a.packed + (size_t)(kb_begin + i) * w1_block_bytes(N0), w1_block_bytes(N0),
                      &bars[EB_W_FULL + s]);
      }
    }
  } else if (warp == kStore) {
    // ---- tile store: the hi tile of every stage goes to HBM once (TMA bulk store) for enc_first_bwd_kernel ----
    if (lane == 0 && a.xt) {
      for (int i = 0; i < nkb; ++i) {
        const int s = i % NS;
        mbar_wait(&bars[EB_A_FULL + s], (i / NS) & 1);
        bulk_store_s2g(a.xt + ((size_t)blockIdx.x * a.xt_kblocks + kb_begin + i) * kXtTile, smem + s * S::stage, kXtTile);
        bulk_store_wait_read();
        mbar_arrive(&bars[EB_STAGE_FREE + s]);
      }
      bulk_store_wait_all();
    }
  } else if (warp == kMma) {
    if (lane == 0) {
      const uint32_t idesc = make_idesc_f16(128, N0, 0, 0);
      for (int i = 0; i < nkb; ++i) {
        const int s = i % NS;
        const uint32_t ph = (i / NS) & 1;
        mbar_wait(&bars[EB_A_FULL + s], ph);
        mbar_wait(&bars[EB_W_FULL + s], ph);
        tc_fence_after();
        const uint32_t sA1 = smem_u32(smem + s * S::stage), sA2 = sA1 + S::A;
        const uint32_t sW1 = sA1 + 2 * S::A, sW2 = sW1 + w1_tile_bytes(N0);
#pragma unroll
        for (int p = 0; p < 3; ++p) {
          const uint32_t sa = (p == 2) ? sA2 : sA1, sb = (p == 1) ? sW2 : sW1;
#pragma unroll
          for (int ks = 0; ks < 4; ++ks)
            umma_f16(tmem, make_smem_desc(sa + ks * 2 * kPadCS, kPadCS, 128), make_smem_desc(sb + ks * 2 * W_CS, W_CS, 128), idesc,
                     (i > 0 || p > 0 || ks > 0) ? 1u : 0u);
        }
        umma_commit(&bars[EB_STAGE_FREE + s]);
      }
      umma_commit(&bars[EB_ACC_FULL]);
    }
  } else if (warp < CW) {
    // ---- converter warps: rows r = (t >> 3) + 64 j, column group cg = t & 7 ----
    constexpr int RPT = 128 * 8 / (CW * 32);   // (row, column-group) items per thread
    constexpr int RSTEP = CW * 32 / 8;
    const int cg = t & 7, rbase = t >> 3;
    const uint32_t step = a.drop.rate > 0.f ? dropout_step(a.drop) : 0u;
    float cur[RPT][8];
    auto fetch = [&](int i, float (*dst)[8]) {
      const int c0 = (kb_begin + i) * 64 + cg * 8;
#pragma unroll
      for (int j = 0; j < RPT; ++j) {
        const int r = row0 + rbase + RSTEP * j;
        const int src = (a.ridx && r < a.B) ? a.ridx[r] : r;
        // rows are addressed through their source index; the bound check stays on the minibatch row
        load8<false>(a.x + ((size_t)src - (size_t)r) * a.G, a.G, a.B, a.G, r, c0, dst[j]);
      }
    };
    // VEC: every converter thread issues four 16-byte async copies per k-block (rows (t >> 4) + 32 j, chunk t & 15 of the
    // 256-byte row segment), NR - 1 k-blocks ahead of the one it converts; cells / genes outside the matrix are zero-filled.
    // (A single loader warp issuing all 2048 copies of a tile was issue-bound: 1.5 TB/s.)
    // (XU16: two copies per thread -- rows (t >> 3) + 64 j, chunk t & 7 of the 128-byte row segment)
    constexpr int NCP = XU16 ? 2 : 4, CHUNKS = XU16 ? 8 : 16, RSEG = XU16 ? 128 : 256, EPC = XU16 ? 8 : 4;   // copies, chunks / row, bytes / row, entries / chunk
    constexpr int ESZ = XU16 ? 2 : 4;
    const int cp_r = t / CHUNKS, cp_c = t % CHUNKS;
    const uint8_t* cp_rows[NCP];          // this thread's source rows (gathered through ridx when given)
#pragma unroll
    for (int j = 0; j < NCP; ++j) {
      const int r = min(row0 + cp_r + (512 / CHUNKS) * j, a.B - 1);
      cp_rows[j] = reinterpret_cast<const uint8_t*>(a.x) + ((size_t)(a.ridx ? a.ridx[r] : r) * a.G + kb_begin * 64 + cp_c * EPC) * ESZ;
    }
    uint8_t* cp_dst = smem + S::raw + cp_r * RSEG + cp_c * 16;
    auto issue = [&](int i) {
      const int rs = i % NR;
      const bool c_ok = (kb_begin + i) * 64 + cp_c * EPC < a.G;
#pragma unroll
      for (int j = 0; j < NCP; ++j) {
        const bool ok = c_ok && row0 + cp_r + (512 / CHUNKS) * j < a.B;
        cp_async_16_zfill(cp_dst + rs * kRawTile + j * (512 / CHUNKS) * RSEG, ok ? (const void*)(cp_rows[j] + (size_t)i * 64 * ESZ) : (const void*)a.x,
                          ok ? 16u : 0u);
      }
      cp_async_mbar_arrive_noinc(&bars[EB_RAW_FULL + rs]);
    };
    if (VEC) {
      for (int i = 0; i < NR - 1 && i < nkb; ++i) issue(i);
    } else {
      fetch(0, cur);
    }
    for (int i = 0; i < nkb; ++i) {
      const int s = i % NS;
      const int c0 = (kb_begin + i) * 64 + cg * 8;
      float nxt[RPT][8];
      if (VEC) {
        const int rs = i % NR;
        if (i + NR - 1 < nkb) {      // refill the stage every warp finished reading one k-block ago
          if (i >= 1) mbar_wait(&bars[EB_RAW_FREE + (i - 1) % NR], ((i - 1) / NR) & 1);
          issue(i + NR - 1);
        }
        mbar_wait(&bars[EB_RAW_FULL + rs], (i / NR) & 1);
        // two 16-byte halves per item; column groups 4..7 read theirs in the opposite order, which spreads a quarter
        // warp over all 32 banks
        const bool swap = (cg >> 2) & 1;
#pragma unroll
        for (int j = 0; j < RPT; ++j) {
          const int r = rbase + RSTEP * j;
          if (XU16) {         // 8 counts = one 16-byte read; a quarter warp (8 column groups of a row) covers 128 contiguous bytes
            const uint4 w = *reinterpret_cast<const uint4*>(smem + S::raw + rs * kRawTile + r * 128 + cg * 16);
            cur[j][0] = u16_to_float(w.x & 0xffffu); cur[j][1] = u16_to_float(w.x >> 16);
            cur[j][2] = u16_to_float(w.y & 0xffffu); cur[j][3] = u16_to_float(w.y >> 16);
            cur[j][4] = u16_to_float(w.z & 0xffffu); cur[j][5] = u16_to_float(w.z >> 16);
            cur[j][6] = u16_to_float(w.w & 0xffffu); cur[j][7] = u16_to_float(w.w >> 16);
            continue;
          }
          const uint8_t* src = smem + S::raw + rs * kRawTile + r * 256 + cg * 32;
          const float4 p = *reinterpret_cast<const float4*>(src + (swap ? 16 : 0));
          const float4 q = *reinterpret_cast<const float4*>(src + (swap ? 0 : 16));
          const float4 lo = swap ? q : p, hi = swap ? p : q;       // cells / genes outside the matrix were zero-filled
          cur[j][0] = lo.x; cur[j][1] = lo.y; cur[j][2] = lo.z; cur[j][3] = lo.w;
          cur[j][4] = hi.x; cur[j][5] = hi.y; cur[j][6] = hi.z; cur[j][7] = hi.w;
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars[EB_RAW_FREE + rs]);     // the raw stage can be refilled while this tile is converted
      } else if (i + 1 < nkb) {
        fetch(i + 1, nxt);
      }
      if (i >= NS) mbar_wait(&bars[EB_STAGE_FREE + s], ((i / NS) - 1) & 1);
      uint8_t* A1 = smem + s * S::stage;
#pragma unroll
      for (int j = 0; j < RPT; ++j) {
        const int r = rbase + RSTEP * j;
        normalise8(cur[j], a.log_norm, a.drop, step, row0 + r, c0);
        store8_hi_lo(A1, A1 + S::A, cg * kPadCS + (r >> 3) * 128 + (r & 7) * 16, cur[j]);
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars[EB_A_FULL + s]);
      if (!VEC && i + 1 < nkb) {
#pragma unroll
        for (int j = 0; j < RPT; ++j)
#pragma unroll
          for (int k = 0; k < 8; ++k) cur[j][k] = nxt[j][k];
      }
    }
    // ---- epilogue: TMEM -> A0 ----
    mbar_wait(&bars[EB_ACC_FULL], 0);
    tc_fence_after();
    const int q = warp & 3, half = warp >> 2;       // TMEM lane quarter, column slice
    const int row = row0 + q * 32 + lane;
    constexpr int CPT = N0 / (CW / 4);     // columns per thread
#pragma unroll
    for (int c = 0; c < CPT; c += 16) {
      float v[16];
      tmem_ld16(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(half * CPT + c), v);
      tmem_ld_wait();
      if (row < a.B) {
        float* dst = a.A0 + (size_t)row * a.ld0 + half * CPT + c;
        if (a.atomic_out) {
#pragma unroll
          for (int j = 0; j < 4; ++j) red_add_v4(dst + 4 * j, v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
        } else {
#pragma unroll
          for (int j = 0; j < 4; ++j) *reinterpret_cast<float4*>(dst + 4 * j) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kMma) tmem_dealloc(tmem, N0 < 32 ? 32 : N0);
}

// ------------------------------------------------------------------------------------------------
struct EncBwdArgs {
  const uint8_t* xt;        // fp16 tiles written by enc_first_fwd_kernel: [cell_tiles][xt_kblocks][kXtTile]
  int xt_kblocks;           // even; gene tile g of this kernel = k-blocks 2g, 2g+1 (adjacent in memory)
  const float* delta;       // [B, ld0] d loss / d pre-activation of the first layer
  float* dW;                // [N0, Gp] += delta^T . x~
  int B, G, Gp, ld0, n_cell_tiles, tiles_per_chunk, log_norm;
  float in_scale, out_scale;   // delta is multiplied by in_scale before fp16, the result by out_scale = 1/in_scale
  DropSpec drop;
};

template <int N0>
struct EncBwdSmem {
  static constexpr int A = 16 * kPadCS;                      // x~ tile [128 cells][128 genes] fp16
  static constexpr int Bt = (N0 / 8) * kPadCS;               // delta tile [128 cells][N0] fp16
  static constexpr int stage = A + Bt;
  static constexpr int bar = kEncStages * stage;
  static constexpr int total = bar + 16 * 8 + 16;
};

template <int N0, bool VEC>
__global__ void __launch_bounds__(kEncThreads, 1) enc_first_bwd_kernel(EncBwdArgs a) {
  extern __shared__ __align__(128) uint8_t smem[];
  using S = EncBwdSmem<N0>;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S::bar);   // [0..2] full, [3..5] free, [6] acc_full
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + S::bar + 16 * 8);
  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  const int g0 = blockIdx.x * 128;
  const int ct_begin = blockIdx.y * a.tiles_per_chunk;
  const int nct = min(a.n_cell_tiles, ct_begin + a.tiles_per_chunk) - ct_begin;
  if (nct <= 0) return;
  if (t == 0) {
    for (int s = 0; s < kEncStages; ++s) { mbar_init(&bars[s], 8 + 1); mbar_init(&bars[3 + s], 1); }   // 8 converter warps + the TMA arrival
    mbar_init(&bars[6], 1);
    fence_barrier_init();
  }
  if (warp == 8) tmem_alloc(tmem_slot, N0 < 32 ? 32 : N0);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 8) {
    if (lane == 0) {
      const uint32_t idesc = make_idesc_f16(128, N0, 1, 1);
      for (int i = 0; i < nct; ++i) {
        const int s = i % kEncStages;
        mbar_wait(&bars[s], (i / kEncStages) & 1);
        tc_fence_after();
        const uint32_t sA = smem_u32(smem + s * S::stage), sB = sA + S::A;
#pragma unroll
        for (int ks = 0; ks < 8; ++ks)   // 16 cells per step: two 8-row groups
          umma_f16(tmem, make_smem_desc(sA + ks * 256, 128, kPadCS), make_smem_desc(sB + ks * 256, 128, kPadCS), idesc,
                   (i > 0 || ks > 0) ? 1u : 0u);
        umma_commit(&bars[3 + s]);
      }
      umma_commit(&bars[6]);
    }
  } else if (warp == 9) {
    // ---- TMA: two adjacent 64-gene tiles of normalised counts per 128-cell tile = one 33 KB bulk copy ----
    if (lane == 0) {
      for (int i = 0; i < nct; ++i) {
        const int s = i % kEncStages;
        if (i >= kEncStages) mbar_wait(&bars[3 + s], ((i / kEncStages) - 1) & 1);
        mbar_arrive_expect_tx(&bars[s], 2 * kXtTile);
        bulk_copy_g2s(smem + s * S::stage, a.xt + ((size_t)(ct_begin + i) * a.xt_kblocks + 2 * blockIdx.x) * kXtTile, 2 * kXtTile,
                      &bars[s]);
      }
    }
  } else if (warp < 8) {
    const int cgd = t & 7, rd = t >> 3;           // delta tile (per 64 columns): 8 column groups, rows rd + 32 j (j < 4)
    for (int i = 0; i < nct; ++i) {
      const int s = i % kEncStages;
      const int row0 = (ct_begin + i) * 128;
      float dv[N0 / 64][4][8];
#pragma unroll
      for (int blk = 0; blk < N0 / 64; ++blk)
#pragma unroll
        for (int j = 0; j < 4; ++j) load8<true>(a.delta, a.ld0, a.B, a.ld0, row0 + rd + 32 * j, blk * 64 + cgd * 8, dv[blk][j]);
      if (i >= kEncStages) mbar_wait(&bars[3 + s], ((i / kEncStages) - 1) & 1);
      uint8_t* Bt = smem + s * S::stage + S::A;
#pragma unroll
      for (int blk = 0; blk < N0 / 64; ++blk)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int r = rd + 32 * j;
          store8_hi(Bt, (blk * 8 + cgd) * kPadCS + (r >> 3) * 128 + (r & 7) * 16, dv[blk][j], a.in_scale);
        }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars[s]);
    }
    // ---- epilogue: TMEM lane = gene, column = output unit n ----
    mbar_wait(&bars[6], 0);
    tc_fence_after();
    const int q = warp & 3, half = warp >> 2;
    const int g = g0 + q * 32 + lane;
    constexpr int CPT = N0 / 2;
#pragma unroll
    for (int c = 0; c < CPT; c += 16) {
      float v[16];
      tmem_ld16(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(half * CPT + c), v);
      tmem_ld_wait();
      if (g < a.G) {
#pragma unroll
        for (int j = 0; j < 16; ++j) atomicAdd(a.dW + (size_t)(half * CPT + c + j) * a.Gp + g, v[j] * a.out_scale);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 8) tmem_dealloc(tmem, N0 < 32 ? 32 : N0);
}

}  // namespace tc
}  // namespace sisua
