"""Host-side minibatch formats (sisua_b200/pipeline.py) — no GPU needed: what is shipped over PCIe must decode to
exactly the float32 matrix the reference's pipeline would have handed over."""
import numpy as np
import pytest
import torch

from sisua_b200 import synthetic as SY
from sisua_b200.pipeline import CsrBatch, quantize_counts


def _counts(rows=37, genes=211, seed=0):
  return SY.realistic_counts(rows, genes, 0, 6.42, 0.28, seed=SY.DATA_SEED + seed)["x"]


def test_csr_batch_decodes_to_the_dense_matrix():
  X = _counts()
  c = CsrBatch(X)
  assert c.rows == X.shape[0] and c.genes == X.shape[1]
  ip = c.indptr.numpy()
  assert ip[0] == 0 and ip[-1] == c.cols.numel() == c.vals.numel() == int((X != 0).sum())
  assert np.all(np.diff(ip) >= 0)
  dense = np.zeros_like(X)
  cols = c.cols.numpy().view(np.uint16).astype(np.int64)
  vals = c.vals.numpy().view(np.uint16).astype(np.float32)
  for r in range(c.rows):
    dense[r, cols[ip[r]:ip[r + 1]]] = vals[ip[r]:ip[r + 1]]
    assert np.all(np.diff(cols[ip[r]:ip[r + 1]]) > 0)          # column ids ascending within a row
  np.testing.assert_array_equal(dense, X)
  assert c.nbytes == ip.size * 4 + cols.size * 2 + vals.size * 2 < X.nbytes
  assert c.indptr.is_pinned() == c.cols.is_pinned() == c.vals.is_pinned()


def test_csr_batch_edge_cases():
  Z = np.zeros((5, 9), dtype=np.float32)
  c = CsrBatch(Z)                                   # all-zero batch: empty value arrays, flat row pointers
  assert c.cols.numel() == 0 and np.array_equal(c.indptr.numpy(), np.zeros(6, dtype=np.int32))
  big = np.zeros((2, 3), dtype=np.float32); big[1, 2] = 65535.0
  assert CsrBatch(big).vals.numpy().view(np.uint16)[0] == 65535
  for bad in (np.array([[0.5]], dtype=np.float32), np.array([[-1.0]], dtype=np.float32), np.array([[65536.0]], dtype=np.float32)):
    with pytest.raises(ValueError):
      CsrBatch(bad)


def test_quantize_counts_is_lossless_or_declines():
  X = _counts(seed=1)
  q = quantize_counts(X)
  assert q.dtype == torch.int16 and q.shape == X.shape
  np.testing.assert_array_equal(q.numpy().view(np.uint16).astype(np.float32), X)
  for Y in (X + 0.25, -X - 1.0, X + 70000.0):        # non-integer, negative, too large: stays float32, unchanged
    f = quantize_counts(Y.astype(np.float32))
    assert f.dtype == torch.float32
    np.testing.assert_array_equal(f.numpy(), Y.astype(np.float32))


def test_packed_csr_round_trip_and_size():
  """Packed CSR ("delta-8", 2 bytes per non-zero): gaps of 255 and more go through skip words, counts of 255 and more
  through the per-row escape list; decoding returns the dense matrix; half the bytes of the 4-byte CSR form."""
  from sisua_b200.pipeline import Csr8Batch, decode_csr8, encode_csr8
  X = _counts(rows=64, genes=2003)
  X[3, :] = 0; X[3, 2002] = 7                       # one entry behind a gap of 2002 (7 skip words)
  X[4, :] = 0                                       # empty row
  X[5, 0] = 255; X[5, 1] = 254; X[5, 300] = 65535   # escapes next to the largest direct value
  X[6, :] = 0; X[6, 255] = 1; X[6, 510] = 2         # gaps that are exact multiples of 255 (advance 0 after the skips)
  ip, ents, bp, big = encode_csr8(X)
  np.testing.assert_array_equal(decode_csr8(ip, ents, bp, big, X.shape[1]), X)
  assert big.tolist() == [255, 65535] or sorted(big.tolist())[-2:] == [255, 65535]
  assert ip[4] == ip[5] and ip[4] - ip[3] == 8       # empty row; 7 skips + the entry
  c8, c = Csr8Batch(X), CsrBatch(X)
  assert c8.nbytes < 0.56 * c.nbytes
  np.testing.assert_array_equal(decode_csr8(c8.indptr.numpy(), c8.ents.numpy().view(np.uint16), c8.big_ptr.numpy(),
                                            c8.big.numpy().view(np.uint16), X.shape[1]), X)
  with pytest.raises(ValueError):
    Csr8Batch(np.array([[0.5]], dtype=np.float32))


def test_host_dataset_packed_batches_are_the_permuted_rows():
  from sisua_b200.models import SingleCellData
  from sisua_b200.pipeline import Csr8Batch, HostDataset, decode_csr8
  X = _counts(rows=96, genes=300, seed=3)
  X[10, 17] = 900.0
  hds = HostDataset(SingleCellData(X, name="toy"), 32, shuffle=True, seed=5, packed=True)
  assert hds.packed and len(hds) == 3
  for s in range(3):
    b, extras = hds.batch(0, s)
    assert isinstance(b, Csr8Batch) and not extras and b.rows == 32
    got = decode_csr8(b.indptr.numpy(), b.ents.numpy().view(np.uint16), b.big_ptr.numpy(), b.big.numpy().view(np.uint16), 300)
    np.testing.assert_array_equal(got, X[hds.order[32 * s:32 * (s + 1)]])
  assert hds.h2d_bytes == sum((33 * 8 + 2 * (int(hds.indptr[32 * (s + 1)] - hds.indptr[32 * s]) + max(1, int(hds.big_ptr[32 * (s + 1)] - hds.big_ptr[32 * s])))) for s in range(3))
