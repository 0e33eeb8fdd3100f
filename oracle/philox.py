"""NumPy replica of the product's counter-based dropout masks.  TEST INFRASTRUCTURE ONLY.

The CUDA step draws its dropout masks from Philox4x32-10(key = seed; counter = (row, col // 8, step,
stream)), 16 bits per column (sisua_b200/csrc/device_math.cuh: philox4x32_10 / dropout_mult).  Philox is the
published counter-based generator of Salmon et al. (SC'11); this file restates it so the oracle can
be fed the very same masks.  Known-answer vectors from the Random123 distribution pin it
(tests/test_oracle_kat.py)."""
import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = 0x9E3779B9, 0xBB67AE85
MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
  c0, c1, c2, c3 = (np.asarray(c, dtype=np.uint64) & MASK for c in (c0, c1, c2, c3))
  k0, k1 = int(k0) & 0xFFFFFFFF, int(k1) & 0xFFFFFFFF
  for _ in range(10):
    p0 = M0 * c0
    p1 = M1 * c2
    hi0, lo0 = p0 >> np.uint64(32), p0 & MASK
    hi1, lo1 = p1 >> np.uint64(32), p1 & MASK
    c0, c1, c2, c3 = (hi1 ^ c1 ^ np.uint64(k0)) & MASK, lo1, (hi0 ^ c3 ^ np.uint64(k1)) & MASK, lo0
    k0 = (k0 + W0) & 0xFFFFFFFF
    k1 = (k1 + W1) & 0xFFFFFFFF
  return c0, c1, c2, c3


def dropout_mask(rows: int, cols: int, rate: float, seed: int, step: int, stream: int) -> np.ndarray:
  """[rows, cols] array of {0, 1}: 1 = kept.  One Philox call covers 8 consecutive columns: column j of the group
  takes the (low if j even else high) 16 bits of output word j // 2 and is kept when that is >= floor(rate * 65536)."""
  c8 = (cols + 7) // 8
  r = np.repeat(np.arange(rows, dtype=np.uint64)[:, None], c8, axis=1)
  c = np.repeat(np.arange(c8, dtype=np.uint64)[None, :], rows, axis=0)
  out = philox4x32_10(r, c, np.full_like(r, step & 0xFFFFFFFF), np.full_like(r, stream),
                      seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
  w = np.stack(out, axis=-1)                                    # [rows, c8, 4]
  lo = w & np.uint64(0xFFFF)
  hi = (w >> np.uint64(16)) & np.uint64(0xFFFF)
  u = np.stack([lo, hi], axis=-1).reshape(rows, c8 * 8)[:, :cols]   # word-major, low half first
  thr = np.uint64(int(np.float32(rate) * np.float32(65536.0)))
  return (u >= thr).astype(np.float64)


NOISE_STREAM_Z, NOISE_STREAM_L = 0x100, 0x101


def normal_noise(rows: int, cols: int, seed: int, step: int, stream: int) -> np.ndarray:
  """[rows, cols] standard normals as the CUDA step draws them when no eps is injected
  (sisua_b200/csrc/device_math.cuh: philox_normal4): one Philox4x32-10 call per (row, column group of 4),
  counter = (row, col // 4, step, stream); Box-Muller on (w0, w1) and (w2, w3) with
  u1 = ((w >> 8) + 0.5) 2^-24, u2 = (w >> 8) 2^-24: n = sqrt(-2 ln u1) (cos, sin)(2 pi u2)."""
  c4 = (cols + 3) // 4
  r = np.repeat(np.arange(rows, dtype=np.uint64)[:, None], c4, axis=1)
  c = np.repeat(np.arange(c4, dtype=np.uint64)[None, :], rows, axis=0)
  w = philox4x32_10(r, c, np.full_like(r, step & 0xFFFFFFFF), np.full_like(r, stream), seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
  k24 = 2.0 ** -24
  u = [(x >> np.uint64(8)).astype(np.float64) for x in w]
  ra = np.sqrt(-2.0 * np.log((u[0] + 0.5) * k24)); rb = np.sqrt(-2.0 * np.log((u[2] + 0.5) * k24))
  ta = 2.0 * np.pi * u[1] * k24; tb = 2.0 * np.pi * u[3] * k24
  out = np.stack([ra * np.cos(ta), ra * np.sin(ta), rb * np.cos(tb), rb * np.sin(tb)], axis=-1).reshape(rows, c4 * 4)
  return out[:, :cols]


CORRUPT_STREAM = 0x200


def corrupt_counts(x: np.ndarray, dropout: float, retain_rate: float = 0.2, distribution: str = "binomial", seed: int = 8) -> np.ndarray:
  """Restatement of sisua_corrupt_counts (sisua_b200/csrc/abi.cu: corrupt_kernel), integers only, so bit-exact: entry
  (r, c) > 0 is selected when word 0 of Philox(seed; r, c, 0, 0x200) < floor(float32(dropout) 2^32); a selected count n
  becomes the number of successes among n trials ('binomial') or n times one trial ('uniform'); trial t takes 16 bits
  of call (r, c, 1 + t // 8, 0x200) -- low half of word (t % 8) // 2 first -- and succeeds below
  floor(float32(retain_rate) 65536).  (The algorithm being imitated: sisua/data/utils.py:168-228.)"""
  x = np.asarray(x, dtype=np.float32)
  out = x.copy()
  k0, k1 = seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF
  sel_thr = min(4294967295, int(np.floor(float(np.float32(dropout)) * 4294967296.0)))
  keep_thr = int(np.floor(float(np.float32(retain_rate)) * 65536.0))
  r, c = np.nonzero(x > 0)
  if r.size == 0:
    return out
  r64, c64 = r.astype(np.uint64), c.astype(np.uint64)
  w = philox4x32_10(r64, c64, np.zeros_like(r64), np.full_like(r64, CORRUPT_STREAM), k0, k1)
  sel = w[0] < np.uint64(sel_thr)
  r, c, r64, c64 = r[sel], c[sel], r64[sel], c64[sel]
  n = x[r, c].astype(np.int64)
  trials = np.ones_like(n) if distribution == "uniform" else n
  k = np.zeros_like(n)
  for blk in range(int((trials.max() + 7) // 8) if trials.size else 0):
    live = trials > 8 * blk
    if not live.any():
      break
    w = philox4x32_10(r64[live], c64[live], np.full(int(live.sum()), 1 + blk, dtype=np.uint64), np.full(int(live.sum()), CORRUPT_STREAM, dtype=np.uint64), k0, k1)
    for j in range(8):
      word = w[j >> 1]
      u = (word >> np.uint64(16)) if (j & 1) else (word & np.uint64(0xFFFF))
      ok = (8 * blk + j < trials[live]) & (u < np.uint64(keep_thr))
      k[live] += ok.astype(np.int64)
  out[r, c] = np.where(k > 0, x[r, c], 0.0) if distribution == "uniform" else k.astype(np.float32)
  return out
