"""Thin owner of one C-ABI handle plus the device buffers it is bound to.  torch is used only to
hold device memory and the CUDA stream; every computation goes through libsisua_b200.so."""
from __future__ import annotations

import ctypes
from typing import Dict, Optional

import numpy as np
import torch

from . import _lib
from .config import StepConfig, bn_layer_names, param_layout
from .params import init_bn_moving, init_flat_params


def _ptr(t: Optional[torch.Tensor]):
  return ctypes.c_void_p(0) if t is None else ctypes.c_void_p(t.data_ptr())


class Engine:
  """One model replica on one GPU."""

  def __init__(self, cfg: StepConfig, device: int = 0, flat_params: Optional[np.ndarray] = None,
               bn_moving: Optional[np.ndarray] = None, seed: int = 8, train: bool = True):
    if not torch.cuda.is_available():
      raise _lib.SisuaError("no CUDA device: the SISUA B200 step has no CPU path")
    self.lib = _lib.load()
    self.cfg = cfg
    self.device = torch.device("cuda", device)
    self.handle = ctypes.c_void_p()
    with torch.cuda.device(self.device):
      rc = self.lib.sisua_create(ctypes.byref(cfg), device, ctypes.byref(self.handle))
    if rc != 0:
      msg = self.lib.sisua_last_error(self.handle).decode() if self.handle else "create failed"
      if self.handle:
        self.lib.sisua_destroy(self.handle)
        self.handle = ctypes.c_void_p()
      raise _lib.SisuaError(f"sisua_create: {msg}")
    self.entries, self.total = param_layout(cfg)
    n = ctypes.c_int(0)
    tot = ctypes.c_int64(0)
    self._check(self.lib.sisua_param_layout(self.handle, None, ctypes.byref(n), ctypes.byref(tot)))
    if tot.value != self.total or n.value != len(self.entries):
      raise _lib.SisuaError("host / library parameter layouts disagree")
    flat = init_flat_params(cfg, seed) if flat_params is None else np.asarray(flat_params, dtype=np.float32)
    mov = init_bn_moving(cfg) if bn_moving is None else np.asarray(bn_moving, dtype=np.float32)
    self.params = torch.from_numpy(flat.copy()).to(self.device)
    self.bn_moving = torch.from_numpy(mov.copy()).to(self.device)
    if train:
      self.grads = torch.zeros_like(self.params)
      self.adam_m = torch.zeros_like(self.params)
      self.adam_v = torch.zeros_like(self.params)
    else:
      self.grads = self.adam_m = self.adam_v = None
    self._check(self.lib.sisua_bind_buffers(self.handle, _ptr(self.params), _ptr(self.grads), _ptr(self.adam_m),
                                            _ptr(self.adam_v), _ptr(self.bn_moving)))
    self.step_count = 0

  # -------------------------------------------------------------------------------------------
  def _check(self, rc: int):
    if rc != 0:
      raise _lib.SisuaError(self.lib.sisua_last_error(self.handle).decode())

  def close(self):
    if getattr(self, "handle", None):
      self.lib.sisua_destroy(self.handle)
      self.handle = ctypes.c_void_p()

  def __del__(self):
    try:
      self.close()
    except Exception:
      pass

  def layout(self):
    n = ctypes.c_int(64)
    arr = (_lib.ParamDesc * 64)()
    tot = ctypes.c_int64(0)
    self._check(self.lib.sisua_param_layout(self.handle, arr, ctypes.byref(n), ctypes.byref(tot)))
    return [(arr[i].name.decode(), arr[i].offset, arr[i].rows, arr[i].cols, arr[i].ld, arr[i].kind)
            for i in range(n.value)], tot.value

  def _dev(self, a, dtype=torch.float32):
    if a is None:
      return None
    if isinstance(a, torch.Tensor):
      t = a.to(device=self.device, dtype=dtype)
    else:
      t = torch.from_numpy(np.ascontiguousarray(a)).to(device=self.device, dtype=dtype)
    return t.contiguous()

  def _stream(self):
    return ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

  # -------------------------------------------------------------------------------------------
  def train_step(self, x, y=None, library=None, mask=None, eps_z=None, eps_l=None, terms=None, loss=None,
                 seed: int = 0, step: int = -1):
    """forward + backward on device tensors (or host arrays, copied). Returns (terms [5,B], loss [1]).  A 16-bit integer
    device matrix `x` is taken as uint16 counts (sisua_train_step_gather_u16 without a row index)."""
    if isinstance(x, torch.Tensor) and x.is_cuda and x.dtype in (torch.int16, torch.uint16):
      return self._train_step_u16(x, y, library, mask, eps_z, eps_l, terms, loss, seed, step)
    x = self._dev(x); y = self._dev(y); library = self._dev(library); eps_z = self._dev(eps_z); eps_l = self._dev(eps_l)
    mask = self._dev(mask, torch.uint8)
    B = x.shape[0]
    if terms is None:
      terms = torch.empty((5, B), dtype=torch.float32, device=self.device)
    if loss is None:
      loss = torch.empty((1,), dtype=torch.float32, device=self.device)
    with torch.cuda.device(self.device):
      self._check(self.lib.sisua_train_step(self.handle, _ptr(x), _ptr(y), _ptr(library), _ptr(mask), _ptr(eps_z),
                                            _ptr(eps_l), B, int(seed), int(step), _ptr(terms), _ptr(loss), self._stream()))
    return terms, loss

  def _train_step_u16(self, x, y, library, mask, eps_z, eps_l, terms, loss, seed, step):
    y = self._dev(y); library = self._dev(library); eps_z = self._dev(eps_z); eps_l = self._dev(eps_l); mask = self._dev(mask, torch.uint8)
    if not x.is_contiguous():
      raise ValueError("train_step: 16-bit counts must be contiguous")
    B = x.shape[0]
    terms = torch.empty((5, B), dtype=torch.float32, device=self.device) if terms is None else terms
    loss = torch.empty((1,), dtype=torch.float32, device=self.device) if loss is None else loss
    with torch.cuda.device(self.device):
      self._check(self.lib.sisua_train_step_gather_u16(self.handle, _ptr(x), _ptr(y), _ptr(library), _ptr(mask), None, _ptr(eps_z),
                                                       _ptr(eps_l), B, int(seed), int(step), _ptr(terms), _ptr(loss), self._stream()))
    return terms, loss

  def train_step_gather(self, x_all, rows, y_all=None, library_all=None, mask_all=None, eps_z=None, eps_l=None, terms=None,
                        loss=None, seed: int = 0, step: int = -1):
    """forward + backward on the minibatch `rows` (int32 device tensor [B]) of matrices resident in HBM
    (sisua_train_step_gather): no gathered copy of the count rows is made."""
    if rows.dtype != torch.int32 or not rows.is_cuda or not rows.is_contiguous():
      raise ValueError("train_step_gather: rows must be a contiguous int32 device tensor")
    for t in (x_all, y_all, library_all, mask_all, eps_z, eps_l):
      if t is not None and (not t.is_cuda or not t.is_contiguous()):
        raise ValueError("train_step_gather: inputs must be contiguous device tensors")
    # a 16-bit integer matrix is the compact (uint16) resident storage of the counts (sisua_train_step_gather_u16)
    if x_all.dtype in (torch.uint16, torch.int16):
      entry = self.lib.sisua_train_step_gather_u16
    elif x_all.dtype == torch.float32:
      entry = self.lib.sisua_train_step_gather
    else:
      raise ValueError(f"train_step_gather: counts must be float32 or 16-bit integers, got {x_all.dtype}")
    B = rows.numel()
    if terms is None:
      terms = torch.empty((5, B), dtype=torch.float32, device=self.device)
    if loss is None:
      loss = torch.empty((1,), dtype=torch.float32, device=self.device)
    with torch.cuda.device(self.device):
      self._check(entry(self.handle, _ptr(x_all), _ptr(y_all), _ptr(library_all), _ptr(mask_all), _ptr(rows), _ptr(eps_z), _ptr(eps_l), B,
                        int(seed), int(step), _ptr(terms), _ptr(loss), self._stream()))
    return terms, loss

  def widen_rows(self, x_u16: torch.Tensor, rows: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """fp32 copy of rows of a uint16 resident count matrix (sisua_widen_rows_u16); rows=None: the whole matrix."""
    if x_u16.dtype not in (torch.uint16, torch.int16) or not x_u16.is_cuda or not x_u16.is_contiguous():
      raise ValueError("widen_rows: x must be a contiguous 16-bit integer device matrix")
    n = x_u16.shape[0] if rows is None else rows.numel()
    if rows is not None and (rows.dtype != torch.int32 or not rows.is_cuda or not rows.is_contiguous()):
      raise ValueError("widen_rows: rows must be a contiguous int32 device tensor")
    out = torch.empty((n, x_u16.shape[1]), dtype=torch.float32, device=self.device) if out is None else out
    with torch.cuda.device(self.device):
      self._check(self.lib.sisua_widen_rows_u16(self.handle, _ptr(x_u16), _ptr(rows), _ptr(out), int(n), self._stream()))
    return out

  def adam_step(self, lr=1e-3, beta1=0.9, beta2=0.999, eps_hat=1e-7, clipnorm=100.0, grad_scale=1.0, t=0):
    with torch.cuda.device(self.device):
      self._check(self.lib.sisua_adam_step(self.handle, lr, beta1, beta2, eps_hat, clipnorm or 0.0, grad_scale,
                                           int(t), self._stream()))
    self.step_count = t if t > 0 else self.step_count + 1

  # ---- data-parallel optimiser step over peer memory ---------------------------------------------------------------
  def rebind(self, params: torch.Tensor, grads: torch.Tensor):
    """Move the flat parameter / gradient buffers into caller-provided (symmetric) device memory."""
    params.copy_(self.params)
    grads.zero_()
    self.params, self.grads = params, grads
    self._check(self.lib.sisua_bind_buffers(self.handle, _ptr(self.params), _ptr(self.grads), _ptr(self.adam_m), _ptr(self.adam_v),
                                            _ptr(self.bn_moving)))

  def dp_bind(self, rank: int, world: int, peer_grads, peer_params, peer_sq, peer_flags, grid: int = 0):
    """sisua_dp_bind: lists of `world` device pointers (ints) to every rank's symmetric buffers."""
    arr = lambda ps: (ctypes.c_void_p * world)(*[ctypes.c_void_p(int(p)) for p in ps])
    self._dp_keep = (arr(peer_grads), arr(peer_params), arr(peer_sq), arr(peer_flags))
    self._check(self.lib.sisua_dp_bind(self.handle, int(rank), int(world), *self._dp_keep, int(grid)))

  def adam_step_dp(self, lr=1e-3, beta1=0.9, beta2=0.999, eps_hat=1e-7, clipnorm=100.0, t=0):
    with torch.cuda.device(self.device):
      self._check(self.lib.sisua_adam_step_dp(self.handle, lr, beta1, beta2, eps_hat, clipnorm or 0.0, int(t), self._stream()))
    self.step_count = t if t > 0 else self.step_count + 1

  def dp_shard(self):
    b, e = ctypes.c_int64(0), ctypes.c_int64(0)
    self._check(self.lib.sisua_dp_shard(self.handle, ctypes.byref(b), ctypes.byref(e)))
    return int(b.value), int(e.value)

  def infer(self, x, y=None, library=None, mask=None, eps_z=None, eps_l=None, S=1, want_mean=True,
            want_disp=False, want_pi=False) -> Dict[str, torch.Tensor]:
    cfg = self.cfg
    x = self._dev(x); y = self._dev(y); library = self._dev(library); eps_z = self._dev(eps_z); eps_l = self._dev(eps_l)
    mask = self._dev(mask, torch.uint8)
    B, G, Z, P = x.shape[0], cfg.n_genes, cfg.n_latent, cfg.n_proteins
    R = S * B
    f = lambda *shape: torch.empty(shape, dtype=torch.float32, device=self.device)
    out = dict(terms=f(5, R), z_loc=f(B, Z), z_scale=f(B, Z))
    out["mean"] = f(R, G) if want_mean else None
    out["disp"] = f(R, G) if want_disp else None
    out["pi_logit"] = f(R, G) if (want_pi and cfg.n_out_heads == 3) else None
    out["lib_loc"] = f(B) if cfg.model_kind == 1 else None
    out["lib_scale"] = f(B) if cfg.model_kind == 1 else None
    out["y_mean"] = f(R, P) if P > 0 else None
    with torch.cuda.device(self.device):
      self._check(self.lib.sisua_infer(self.handle, _ptr(x), _ptr(y), _ptr(library), _ptr(mask), _ptr(eps_z),
                                       _ptr(eps_l), B, S, _ptr(out["terms"]), _ptr(out["z_loc"]), _ptr(out["z_scale"]),
                                       _ptr(out["lib_loc"]), _ptr(out["lib_scale"]), _ptr(out["mean"]),
                                       _ptr(out["disp"]), _ptr(out["pi_logit"]), _ptr(out["y_mean"]), self._stream()))
    return out

  def infer_ex(self, x, x_eval=None, y=None, library=None, mask=None, eps_z=None, eps_l=None, S=1, strip_zi=False, want_mean=False,
               want_disp=False, want_pi=False, want_mean_avg=False, want_logw=False, want_latent=True) -> Dict[str, torch.Tensor]:
    """sisua_infer_ex: inference step with the Posterior options -- likelihood evaluated on `x_eval` while the encoder
    reads `x`, zero inflation stripped, Monte-Carlo mean of the NB mean [B,G], importance weights [S*B]."""
    cfg = self.cfg
    x = self._dev(x); x_eval = self._dev(x_eval); y = self._dev(y); library = self._dev(library)
    eps_z = self._dev(eps_z); eps_l = self._dev(eps_l); mask = self._dev(mask, torch.uint8)
    B, G, Z, P = x.shape[0], cfg.n_genes, cfg.n_latent, cfg.n_proteins
    R = S * B
    f = lambda *shape: torch.empty(shape, dtype=torch.float32, device=self.device)
    out = dict(terms=f(5, R))
    out["z_loc"] = f(B, Z) if want_latent else None
    out["z_scale"] = f(B, Z) if want_latent else None
    out["mean"] = f(R, G) if want_mean else None
    out["disp"] = f(R, G) if want_disp else None
    out["pi_logit"] = f(R, G) if (want_pi and cfg.n_out_heads == 3) else None
    out["lib_loc"] = f(B) if (cfg.model_kind == 1 and want_latent) else None
    out["lib_scale"] = f(B) if (cfg.model_kind == 1 and want_latent) else None
    out["y_mean"] = f(R, P) if P > 0 else None
    out["mean_avg"] = f(B, G) if want_mean_avg else None
    out["logw"] = f(R) if want_logw else None
    with torch.cuda.device(self.device):
      self._check(self.lib.sisua_infer_ex(self.handle, _ptr(x), _ptr(x_eval), _ptr(y), _ptr(library), _ptr(mask), _ptr(eps_z),
                                          _ptr(eps_l), B, S, 1 if strip_zi else 0, _ptr(out["terms"]), _ptr(out["z_loc"]),
                                          _ptr(out["z_scale"]), _ptr(out["lib_loc"]), _ptr(out["lib_scale"]), _ptr(out["mean"]),
                                          _ptr(out["disp"]), _ptr(out["pi_logit"]), _ptr(out["y_mean"]), _ptr(out["mean_avg"]),
                                          _ptr(out["logw"]), self._stream()))
    return out

  def train_forward(self, x, y=None, library=None, mask=None, eps_z=None, eps_l=None, seed: int = 0, step: int = 0):
    """sisua_forward_train_mode: training-mode forward (batch statistics, dropout), no gradients; the outputs are kept
    for `last_forward_outputs`."""
    cfg = self.cfg
    x = self._dev(x); y = self._dev(y); library = self._dev(library); eps_z = self._dev(eps_z); eps_l = self._dev(eps_l)
    mask = self._dev(mask, torch.uint8)
    B, G, Z, P = x.shape[0], cfg.n_genes, cfg.n_latent, cfg.n_proteins
    f = lambda *shape: torch.empty(shape, dtype=torch.float32, device=self.device)
    out = dict(terms=f(5, B), z_loc=f(B, Z), z_scale=f(B, Z), mean=f(B, G), disp=f(B, G))
    out["pi_logit"] = f(B, G) if cfg.n_out_heads == 3 else None
    out["lib_loc"] = f(B) if cfg.model_kind == 1 else None
    out["lib_scale"] = f(B) if cfg.model_kind == 1 else None
    out["y_mean"] = f(B, P) if P > 0 else None
    with torch.cuda.device(self.device):
      self._check(self.lib.sisua_forward_train_mode(self.handle, _ptr(x), _ptr(y), _ptr(library), _ptr(mask), _ptr(eps_z),
                                                    _ptr(eps_l), B, int(seed), int(step), _ptr(out["terms"]), _ptr(out["z_loc"]),
                                                    _ptr(out["z_scale"]), _ptr(out["lib_loc"]), _ptr(out["lib_scale"]),
                                                    _ptr(out["mean"]), _ptr(out["disp"]), _ptr(out["pi_logit"]), _ptr(out["y_mean"]),
                                                    self._stream()))
    self._last_forward = out
    return out

  def last_forward_outputs(self, B: int):
    return self._last_forward

  def decode(self, z: torch.Tensor, lib: Optional[torch.Tensor] = None) -> Dict[str, torch.Tensor]:
    """sisua_decode: latent samples [R, z] (+ scVI log-library [R]) -> output parameters, chunked by max_batch."""
    cfg = self.cfg
    R, G, P = z.shape[0], cfg.n_genes, cfg.n_proteins
    f = lambda *shape: torch.empty(shape, dtype=torch.float32, device=self.device)
    out = dict(mean=f(R, G), disp=f(R, G), pi_logit=f(R, G) if cfg.n_out_heads == 3 else None, y_mean=f(R, P) if P > 0 else None)
    with torch.cuda.device(self.device):
      for s in range(0, R, cfg.max_batch):
        e = min(R, s + cfg.max_batch)
        sub = lambda t: None if t is None else t[s:e]
        self._check(self.lib.sisua_decode(self.handle, _ptr(z[s:e].contiguous()), _ptr(None if lib is None else lib[s:e].contiguous()),
                                          e - s, _ptr(sub(out["mean"])), _ptr(sub(out["disp"])), _ptr(sub(out["pi_logit"])),
                                          _ptr(sub(out["y_mean"])), self._stream()))
    return out

  def marginal_llk(self, x, y=None, library=None, mask=None, eps_z=None, eps_l=None, S=100):
    """sisua_marginal_llk: (mllk [B], llk_x [B], llk_y [B] or None) for one minibatch, S importance samples."""
    x = self._dev(x); y = self._dev(y); library = self._dev(library); eps_z = self._dev(eps_z); eps_l = self._dev(eps_l)
    mask = self._dev(mask, torch.uint8)
    B = x.shape[0]
    f = lambda: torch.empty((B,), dtype=torch.float32, device=self.device)
    mllk, llk_x = f(), f()
    llk_y = f() if self.cfg.n_proteins > 0 else None
    with torch.cuda.device(self.device):
      self._check(self.lib.sisua_marginal_llk(self.handle, _ptr(x), _ptr(y), _ptr(library), _ptr(mask), _ptr(eps_z), _ptr(eps_l),
                                              B, int(S), _ptr(mllk), _ptr(llk_x), _ptr(llk_y), self._stream()))
    return mllk, llk_x, llk_y

  # -------------------------------------------------------------------------------------------
  def unpack_counts_u16(self, src_u16: torch.Tensor, dst_f32: torch.Tensor):
    """device uint16 counts -> device fp32 counts (same shape), on the current stream."""
    if src_u16.dtype not in (torch.uint16, torch.int16) or dst_f32.dtype != torch.float32 or src_u16.numel() != dst_f32.numel():
      raise ValueError("unpack_counts_u16: need uint16 source and fp32 destination of equal size")
    with torch.cuda.device(self.device):
      self._check(self.lib.sisua_unpack_counts_u16(self.handle, _ptr(src_u16), _ptr(dst_f32), src_u16.numel(), self._stream()))

  def enable_grad_overlap(self):
    """Returns (event, (begin, end)): `event` fires inside train_step once grads[begin:end] (the output heads) are
    final; see distributed.OverlappedAllReduce."""
    ev = torch.cuda.Event()
    ev.record(torch.cuda.current_stream(self.device))      # forces creation of the underlying cudaEvent_t
    self._check(self.lib.sisua_set_grad_ready_event(self.handle, ctypes.c_void_p(ev.cuda_event)))
    self._grad_event = ev
    ents = {e.name: e for e in self.entries}
    begin = ents["out.W"].offset
    end = self.total if self.entries[-1].name == "out.b" else ents["out.b"].offset + ents["out.b"].size
    return ev, (begin, end)

  def unpack_counts_csr(self, indptr: torch.Tensor, cols: torch.Tensor, vals: torch.Tensor, dst: torch.Tensor):
    """device CSR minibatch (int32 indptr [rows+1], 16-bit cols / vals) -> dense [rows, G] on the current stream; `dst`
    float32, or 16-bit integers (the compact form `train_step` accepts directly)."""
    rows = indptr.numel() - 1
    if dst.shape != (rows, self.cfg.n_genes) or dst.dtype not in (torch.float32, torch.int16, torch.uint16) or indptr.dtype != torch.int32:
      raise ValueError("unpack_counts_csr: bad shapes / dtypes")
    fn = self.lib.sisua_unpack_counts_csr if dst.dtype == torch.float32 else self.lib.sisua_unpack_counts_csr_u16
    with torch.cuda.device(self.device):
      self._check(fn(self.handle, _ptr(indptr), _ptr(cols), _ptr(vals), _ptr(dst), rows, self._stream()))

  def unpack_counts_csr8(self, indptr: torch.Tensor, big_ptr: torch.Tensor, ents: torch.Tensor, big: torch.Tensor, dst_u16: torch.Tensor):
    """device packed-CSR minibatch (pipeline.Csr8Batch layout) -> dense 16-bit [rows, G] on the current stream."""
    rows = indptr.numel() - 1
    if (dst_u16.shape != (rows, self.cfg.n_genes) or dst_u16.dtype not in (torch.int16, torch.uint16) or indptr.dtype != torch.int32
        or big_ptr.dtype != torch.int32 or big_ptr.numel() != rows + 1):
      raise ValueError("unpack_counts_csr8: bad shapes / dtypes")
    with torch.cuda.device(self.device):
      self._check(self.lib.sisua_unpack_counts_csr8_u16(self.handle, _ptr(indptr), _ptr(big_ptr), _ptr(ents), _ptr(big), _ptr(dst_u16), rows,
                                                        self._stream()))

  def train_step_host(self, x, *, y=None, library=None, mask=None, eps_z=None, eps_l=None, host_loss=None, host_terms=None,
                      seed: int = 0, step: int = -1):
    """One train step from HOST tensors (sisua_train_step_host): `x` is a pinned float32 / 16-bit [B,G] tensor or a
    `pipeline.CsrBatch`; the other arguments are host tensors as in `train_step`.  The library stages them on its
    own copy stream (double-buffered), so consecutive calls overlap H2D with compute.  Asynchronous: `host_loss`
    (pinned, 1 float) / `host_terms` (pinned, [5,B]) are valid after the current stream has been synchronised, and
    the host tensors must stay alive until then."""
    hb = _lib.HostBatch()
    hp = lambda t: None if t is None else ctypes.c_void_p(t.data_ptr())
    if hasattr(x, "indptr"):
      hb.format, hb.B, hb.nnz = _lib.HOST_CSR, x.rows, x.cols.numel()
      hb.indptr, hb.cols, hb.vals = hp(x.indptr), hp(x.cols), hp(x.vals)
      if x.genes != self.cfg.n_genes:
        raise ValueError("train_step_host: CSR batch has a different gene count")
    else:
      if x.is_cuda or x.dim() != 2 or x.shape[1] != self.cfg.n_genes or not x.is_contiguous():
        raise ValueError("train_step_host: x must be a contiguous host [B, n_genes] tensor")
      if x.dtype == torch.float32:
        hb.format = _lib.HOST_F32
      elif x.dtype in (torch.int16, torch.uint16):
        hb.format = _lib.HOST_U16
      else:
        raise ValueError("train_step_host: x must be float32 or a 16-bit integer tensor")
      hb.B, hb.x = x.shape[0], hp(x)
    for name, t, dt in (("y", y, torch.float32), ("library", library, torch.float32), ("mask", mask, torch.uint8),
                        ("eps_z", eps_z, torch.float32), ("eps_l", eps_l, torch.float32)):
      if t is not None:
        if t.is_cuda or t.dtype != dt or not t.is_contiguous():
          raise ValueError(f"train_step_host: {name} must be a contiguous host tensor of dtype {dt}")
        setattr(hb, name, hp(t))
    with torch.cuda.device(self.device):
      self._check(self.lib.sisua_train_step_host(self.handle, ctypes.byref(hb), int(seed), int(step), hp(host_loss),
                                                 hp(host_terms), self._stream()))

  def corrupt_counts(self, x: torch.Tensor, dropout: float, retain_rate: float = 0.2, distribution: str = "binomial", seed: int = 8,
                     out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Corrupted copy of the device count matrix `x` [rows, genes] (sisua_corrupt_counts; sisua/data/utils.py:168-228)."""
    dist = {"binomial": 0, "uniform": 1}.get(str(distribution).lower())
    if dist is None:
      raise ValueError(f"Only support 2 corruption distribution: 'uniform' and 'binomial', but given: '{distribution}'")
    x = self._dev(x)
    if x.dim() != 2 or x.stride(1) != 1:
      raise ValueError("corrupt_counts: x must be a [rows, genes] matrix with contiguous rows")
    out = torch.empty_like(x, memory_format=torch.contiguous_format) if out is None else out
    with torch.cuda.device(self.device):
      self._check(self.lib.sisua_corrupt_counts(self.handle, x.data_ptr(), out.data_ptr(), x.shape[0], x.shape[1], x.stride(0), out.stride(0),
                                                float(dropout), float(retain_rate), dist, int(seed) & (2 ** 64 - 1), self._stream()))
    return out

  def nonfinite(self, reset: bool = False) -> bool:
    """True once a training step produced a non-finite loss (sisua_nonfinite_flag; sticky, read without a sync)."""
    v = self.lib.sisua_nonfinite_flag(self.handle, 1 if reset else 0)
    if v < 0:
      raise _lib.SisuaError("sisua_nonfinite_flag: bad handle")
    return bool(v)

  def set_count_bound(self, max_count: float):
    """Largest count the step will see (sisua_set_count_bound): keeps the fp16 gradient operand tiles in range."""
    self._check(self.lib.sisua_set_count_bound(self.handle, float(max_count)))

  def set_infer_seed(self, seed: int, call_index: int = 0):
    """Seed / call index of the in-kernel noise of `infer` calls without injected eps (sisua_set_infer_seed)."""
    self._check(self.lib.sisua_set_infer_seed(self.handle, int(seed), int(call_index)))

  def reset_step_counter(self, t: int):
    """Sets the device-side optimiser step counter (what `adam_step(t=0)` and `train_step(step=-1)` follow)."""
    with torch.cuda.device(self.device):
      self._check(self.lib.sisua_set_step(self.handle, int(t), self._stream()))
    self.step_count = int(t)

  SECTIONS = ("enc_first", "mid_fwd", "out_heads", "mid_bwd", "enc_first_bwd", "adam")

  def launch_count(self) -> int:
    return int(self.lib.sisua_launch_count(self.handle))

  def profile(self, on: bool):
    self._check(self.lib.sisua_profile_enable(self.handle, 1 if on else 0))

  def profile_read(self):
    ms = (ctypes.c_float * 8)()
    cnt = (ctypes.c_int * 8)()
    with torch.cuda.device(self.device):
      self._check(self.lib.sisua_profile_read(self.handle, ms, cnt))
    return {n: (float(ms[i]), int(cnt[i])) for i, n in enumerate(self.SECTIONS)}

  def geometry(self, B: int) -> Dict[str, int]:
    """Launch geometry of the tcgen05 kernels for a train step of B cells (sisua_debug_geometry; tests only)."""
    out = (ctypes.c_int32 * 9)()
    self._check(self.lib.sisua_debug_geometry(self.handle, int(B), out))
    keys = ("enc_cell_tiles", "enc_chunks", "enc_kblocks_per_chunk", "out_cell_tiles", "out_chunks", "out_tiles_per_chunk",
            "bwd_gene_tiles", "bwd_chunks", "bwd_cell_tiles_per_chunk")
    return {k: int(out[i]) for i, k in enumerate(keys)}

  def force_chunks(self, out_chunks: int = 0, enc_chunks: int = 0, bwd_chunks: int = 0):
    """Override the chunk heuristics of the tcgen05 kernels (tests only; 0 = automatic)."""
    self._check(self.lib.sisua_debug_force_chunks(self.handle, int(out_chunks), int(enc_chunks), int(bwd_chunks)))

  def debug_buffer(self, name: str, rows: int, cols: int) -> torch.Tensor:
    """Copy of a private workspace buffer (tests only)."""
    out = torch.empty((rows, cols), dtype=torch.float32, device=self.device)
    with torch.cuda.device(self.device):
      self._check(self.lib.sisua_debug_copy(self.handle, name.encode(), _ptr(out), rows * cols, self._stream()))
    return out

  def params_dict(self) -> Dict[str, np.ndarray]:
    from .params import flat_to_dict
    return flat_to_dict(self.cfg, self.params.detach().cpu().numpy())

  def grads_dict(self) -> Dict[str, np.ndarray]:
    from .params import flat_to_dict
    return flat_to_dict(self.cfg, self.grads.detach().cpu().numpy())
