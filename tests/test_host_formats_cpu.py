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
