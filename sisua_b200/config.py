"""Host-side configuration objects of the hot path.

* ``RVmeta`` / ``NetConf`` mirror the two odin-ai dataclasses the reference
  builds its models from (call sites: sisua/models/single_cell_model.py:74-86,
  sisua/models/scvi.py:33-48, sisua/train.py:75-89).
* ``StepConfig`` is the POD that crosses the C ABI (include/sisua_b200.h); its
  enums carry the open odin-ai questions Q1-Q6 of SURVEY.md section 8a so that a
  different answer is a flag flip.
* ``param_layout`` is the single definition of the flat fp32 parameter buffer;
  the C library computes the same table (``sisua_param_layout``) and a test
  checks both agree.
"""
from __future__ import annotations

import ctypes
import dataclasses
from typing import Any, Dict, List, Optional, Sequence, Tuple

# ----------------------------------------------------------------------------
# enums shared with include/sisua_b200.h
# ----------------------------------------------------------------------------
MODEL_VAE, MODEL_SCVI, MODEL_DCA, MODEL_SISUA = 0, 1, 2, 3
MODEL_NAMES = {MODEL_VAE: "vae", MODEL_SCVI: "scvi", MODEL_DCA: "dca", MODEL_SISUA: "sisua"}

XDIST_ZINBD, XDIST_NBD, XDIST_ZINB, XDIST_NB = 0, 1, 2, 3
YDIST_NB, YDIST_NBD = 0, 1

ACT_SOFTPLUS, ACT_SOFTPLUS1, ACT_SOFTPLUS_P1, ACT_EXP, ACT_IDENTITY = 0, 1, 2, 3, 4
ACT_NAMES = {"softplus": ACT_SOFTPLUS, "softplus1": ACT_SOFTPLUS1, "softplus+1": ACT_SOFTPLUS_P1,
             "exp": ACT_EXP, "identity": ACT_IDENTITY, "linear": ACT_IDENTITY}

MASKNORM_ALL, MASKNORM_LABELLED = 0, 1
CLIP_PER_VARIABLE, CLIP_GLOBAL = 0, 1

# arithmetic of the three big contractions (genes x hidden, hidden x genes)
GEMM_FP32_UNFUSED = 0   # CUDA-core FFMA, un-fused kernels (GPU-side cross-check path)
GEMM_TC_3XFP16 = 1      # fused tcgen05 kernels: error-compensated 3xFP16 forward GEMMs (fp32-grade), fp16 gradient GEMMs

SOFTPLUS1_SHIFT = 0.5413248546129181  # log(e - 1): softplus(x + shift) == 1 at x == 0

MAX_LAYERS = 4


# ----------------------------------------------------------------------------
# mirrors of the odin-ai dataclasses the reference passes around
# ----------------------------------------------------------------------------
@dataclasses.dataclass
class RVmeta:
  """Random-variable description. Reference usage:
  ``RVmeta(10, 'diag', True, 'Latents')`` single_cell_model.py:77,
  ``RVmeta(rna_dim, 'zinbd', projection=True, name='RNA')`` vae.py:30-31."""
  event_shape: Any = 10
  posterior: str = "diag"
  projection: bool = True
  name: str = "RVmeta"
  kwargs: Dict[str, Any] = dataclasses.field(default_factory=dict)

  def __post_init__(self):
    if isinstance(self.event_shape, (tuple, list)):
      if len(self.event_shape) != 1:
        raise ValueError(f"only 1-D event_shape is supported, given {self.event_shape}")
      self.event_shape = int(self.event_shape[0])
    self.event_shape = int(self.event_shape)
    self.posterior = str(self.posterior).lower()

  @property
  def dim(self) -> int:
    return int(self.event_shape)

  @property
  def is_deterministic(self) -> bool:
    return self.posterior in ("relu", "linear", "deterministic", "identity")

  @property
  def is_zero_inflated(self) -> bool:
    return self.posterior.startswith("zi")

  def copy(self, **kw) -> "RVmeta":
    return dataclasses.replace(self, **kw)


@dataclasses.dataclass
class NetConf:
  """Dense-network description. Reference usage:
  ``NetConf([64, 64], batchnorm=True, input_dropout=0.3)`` single_cell_model.py:78-81;
  YAML path ``units: [64, 64], batchnorm: True, dropout: 0.1`` configs/base.yaml:10-17."""
  units: Sequence[int] = (64, 64)
  activation: str = "relu"
  use_bias: bool = True
  batchnorm: bool = False
  input_dropout: float = 0.0
  dropout: float = 0.0
  name: Optional[str] = None
  kwargs: Dict[str, Any] = dataclasses.field(default_factory=dict)

  def __post_init__(self):
    if isinstance(self.units, int):
      self.units = [self.units]
    self.units = [int(u) for u in self.units]
    if self.activation != "relu":
      raise ValueError("the B200 hot path implements the reference's 'relu' networks only")

  def copy(self, **kw) -> "NetConf":
    return dataclasses.replace(self, **kw)


# ----------------------------------------------------------------------------
# the POD that crosses the C ABI
# ----------------------------------------------------------------------------
class StepConfig(ctypes.Structure):
  """Field-for-field image of ``sisua_step_config`` in include/sisua_b200.h."""
  _fields_ = [
      ("model_kind", ctypes.c_int32),
      ("n_genes", ctypes.c_int32),
      ("n_proteins", ctypes.c_int32),
      ("n_latent", ctypes.c_int32),
      ("n_hidden", ctypes.c_int32),
      ("n_enc_layers", ctypes.c_int32),
      ("n_dec_layers", ctypes.c_int32),
      ("n_encl_layers", ctypes.c_int32),
      ("batchnorm", ctypes.c_int32),
      ("log_norm", ctypes.c_int32),
      ("x_dist", ctypes.c_int32),
      ("y_dist", ctypes.c_int32),
      ("mean_act", ctypes.c_int32),
      ("disp_act", ctypes.c_int32),
      ("scale_act", ctypes.c_int32),
      ("scvi_reapply_act", ctypes.c_int32),
      ("mask_norm", ctypes.c_int32),
      ("clip_mode", ctypes.c_int32),
      ("gemm_mode", ctypes.c_int32),
      ("max_batch", ctypes.c_int32),
      ("latent_linear", ctypes.c_int32),
      ("bn_eps", ctypes.c_float),
      ("bn_momentum", ctypes.c_float),
      ("input_dropout", ctypes.c_float),
      ("enc_dropout", ctypes.c_float),
      ("dec_dropout", ctypes.c_float),
      ("encl_dropout", ctypes.c_float),
      ("beta", ctypes.c_float),
      ("alpha", ctypes.c_float),
      ("clip_library", ctypes.c_float),
  ]

  def as_dict(self) -> Dict[str, Any]:
    return {k: getattr(self, k) for k, _ in self._fields_}

  def clone(self, **kw) -> "StepConfig":
    c = StepConfig(**self.as_dict())
    for k, v in kw.items():
      if not hasattr(c, k):
        raise AttributeError(k)
      setattr(c, k, v)
    return c

  @property
  def n_out_heads(self) -> int:
    return 3 if self.x_dist in (XDIST_ZINBD, XDIST_ZINB) else 2

  @property
  def genes_padded(self) -> int:
    return (self.n_genes + 3) // 4 * 4

  @property
  def latent_params(self) -> int:
    return self.n_latent if self.model_kind == MODEL_DCA else 2 * self.n_latent


def make_step_config(model: str = "vae", n_genes: int = 2000, n_proteins: int = 0, n_latent: int = 10,
                     n_hidden: int = 64, n_enc_layers: int = 2, n_dec_layers: int = 2,
                     n_encl_layers: int = 1, batchnorm: bool = True, log_norm: bool = True,
                     x_dist: str = "zinbd", y_dist: str = "nb", mean_act: str = "softplus",
                     disp_act: str = "softplus1", scale_act: str = "softplus1",
                     scvi_reapply_act: bool = False, mask_norm: int = MASKNORM_ALL,
                     clip_mode: int = CLIP_PER_VARIABLE, gemm_mode: int = GEMM_TC_3XFP16,
                     max_batch: int = 8192, latent_linear: bool = False, bn_eps: float = 1e-3, bn_momentum: float = 0.99,
                     input_dropout: float = 0.0, enc_dropout: float = 0.0, dec_dropout: float = 0.0,
                     encl_dropout: float = 0.0, beta: float = 1.0, alpha: float = 10.0,
                     clip_library: float = 1e3) -> StepConfig:
  kinds = {"vae": MODEL_VAE, "scvi": MODEL_SCVI, "dca": MODEL_DCA, "sisua": MODEL_SISUA}
  if model not in kinds:
    raise ValueError(f"unknown model kind '{model}'")
  xd = {"zinbd": XDIST_ZINBD, "nbd": XDIST_NBD, "zinb": XDIST_ZINB, "nb": XDIST_NB}
  yd = {"nb": YDIST_NB, "nbd": YDIST_NBD}
  if x_dist not in xd:
    raise ValueError(f"gene-count distribution '{x_dist}' is not on the B200 hot path (zinbd, nbd, zinb, nb)")
  if y_dist not in yd:
    raise ValueError(f"protein distribution '{y_dist}' is not on the B200 hot path (nb, nbd)")
  kind = kinds[model]
  if kind == MODEL_SCVI and x_dist not in ("zinbd", "nbd"):
    raise ValueError("scVI only supports 'zinbd' / 'nbd' (sisua/models/scvi.py:50-52)")
  if kind != MODEL_SISUA:
    n_proteins = 0
  if kind == MODEL_SISUA and n_proteins <= 0:
    raise ValueError("SISUA needs n_proteins > 0")
  cfg = StepConfig(
      model_kind=kind, n_genes=int(n_genes), n_proteins=int(n_proteins), n_latent=int(n_latent),
      n_hidden=int(n_hidden), n_enc_layers=int(n_enc_layers), n_dec_layers=int(n_dec_layers),
      n_encl_layers=int(n_encl_layers) if kind == MODEL_SCVI else 0,
      batchnorm=int(bool(batchnorm)), log_norm=int(bool(log_norm)),
      x_dist=xd[x_dist], y_dist=yd[y_dist], mean_act=ACT_NAMES[mean_act], disp_act=ACT_NAMES[disp_act],
      scale_act=ACT_NAMES[scale_act], scvi_reapply_act=int(bool(scvi_reapply_act)),
      mask_norm=int(mask_norm), clip_mode=int(clip_mode), gemm_mode=int(gemm_mode),
      max_batch=int(max_batch), latent_linear=int(bool(latent_linear) and kind == MODEL_DCA), bn_eps=float(bn_eps), bn_momentum=float(bn_momentum),
      input_dropout=float(input_dropout), enc_dropout=float(enc_dropout), dec_dropout=float(dec_dropout),
      encl_dropout=float(encl_dropout), beta=float(beta), alpha=float(alpha),
      clip_library=float(clip_library))
  validate(cfg)
  return cfg


def validate(cfg: StepConfig) -> None:
  if cfg.n_hidden != 64:
    raise ValueError("the sm_100a kernels are built for 64 hidden units (reference default "
                     "single_cell_model.py:78-81, configs/base.yaml:11,15)")
  if not (1 <= cfg.n_latent <= 32):
    raise ValueError("n_latent must be in [1, 32]")
  if not (0 <= cfg.n_proteins <= 32):
    raise ValueError("n_proteins must be in [0, 32]")
  for n in (cfg.n_enc_layers, cfg.n_dec_layers):
    if not (1 <= n <= MAX_LAYERS):
      raise ValueError(f"hidden layer count must be in [1, {MAX_LAYERS}]")
  if cfg.model_kind == MODEL_SCVI and not (1 <= cfg.n_encl_layers <= MAX_LAYERS):
    raise ValueError("scVI needs 1..4 library-encoder layers")
  if cfg.n_genes < 1:
    raise ValueError("n_genes must be positive")


# ----------------------------------------------------------------------------
# flat parameter buffer
# ----------------------------------------------------------------------------
@dataclasses.dataclass
class ParamEntry:
  name: str
  offset: int          # in floats, from the start of the flat buffer
  shape: Tuple[int, ...]   # logical shape (rows, cols) or (n,)
  ld: int              # leading dimension in floats (== cols unless padded)
  kind: str            # 'weight' | 'bias' | 'gamma' | 'beta'
  fan_in: int = 0
  fan_out: int = 0

  @property
  def size(self) -> int:   # floats occupied (incl. padding)
    return self.shape[0] * self.ld if len(self.shape) == 2 else self.shape[0]


ALIGN = 64  # floats (256 B): every tensor starts on a TMA/vector friendly boundary


def _align(n: int) -> int:
  return (n + ALIGN - 1) // ALIGN * ALIGN


def param_layout(cfg: StepConfig) -> Tuple[List[ParamEntry], int]:
  """All weights use the [out_features, in_features] convention (rows are output
  units, the contraction index is contiguous = "K-major", which is what the
  tcgen05 B operand wants).  Order: first-layer encoder weights (z encoder then
  library encoder, adjacent so scVI streams the counts once through one
  [2H, G] operand), remaining encoder layers, latent heads, decoder, output
  protein head, and last the output heads (mean | dispersion | dropout-logit blocks of G rows each)."""
  H, G, Z, P = cfg.n_hidden, cfg.n_genes, cfg.n_latent, cfg.n_proteins
  Gp = cfg.genes_padded
  bn = bool(cfg.batchnorm)
  entries: List[ParamEntry] = []
  off = 0

  def add(name, shape, ld, kind, fan_in=0, fan_out=0):
    nonlocal off
    e = ParamEntry(name, off, tuple(shape), ld, kind, fan_in, fan_out)
    entries.append(e)
    off = _align(off + e.size)

  def add_norm_or_bias(prefix):
    if bn:
      add(prefix + ".gamma", (H,), H, "gamma")
      add(prefix + ".beta", (H,), H, "beta")
    else:
      add(prefix + ".b", (H,), H, "bias")

  add("enc.0.W", (H, G), Gp, "weight", G, H)
  if cfg.model_kind == MODEL_SCVI:
    add("encl.0.W", (H, G), Gp, "weight", G, H)
  add_norm_or_bias("enc.0")
  for i in range(1, cfg.n_enc_layers):
    add(f"enc.{i}.W", (H, H), H, "weight", H, H)
    add_norm_or_bias(f"enc.{i}")
  if cfg.model_kind == MODEL_SCVI:
    add_norm_or_bias("encl.0")
    for i in range(1, cfg.n_encl_layers):
      add(f"encl.{i}.W", (H, H), H, "weight", H, H)
      add_norm_or_bias(f"encl.{i}")
  ZP = cfg.latent_params
  add("lat.W", (ZP, H), H, "weight", H, ZP)
  add("lat.b", (ZP,), ZP, "bias")
  if cfg.model_kind == MODEL_SCVI:
    add("lib.W", (2, H), H, "weight", H, 2)
    add("lib.b", (2,), 2, "bias")
  add("dec.0.W", (H, Z), Z, "weight", Z, H)
  add_norm_or_bias("dec.0")
  for i in range(1, cfg.n_dec_layers):
    add(f"dec.{i}.W", (H, H), H, "weight", H, H)
    add_norm_or_bias(f"dec.{i}")
  if P > 0:
    add("y.W", (2 * P, H), H, "weight", H, 2 * P)
    add("y.b", (2 * P,), 2 * P, "bias")
  # the output heads come last: their gradients (about 3/4 of all bytes) are final early in the backward pass
  # and are all-reduced as one contiguous bucket while the rest of the backward runs
  NO = cfg.n_out_heads * G
  # scVI builds three separate Dense(64 -> G) layers (scvi.py:67-83): glorot fan_out = G
  fo = G if cfg.model_kind == MODEL_SCVI else NO
  add("out.W", (NO, H), H, "weight", H, fo)
  add("out.b", (NO,), NO, "bias")
  return entries, off


def bn_layer_names(cfg: StepConfig) -> List[str]:
  """Order of the (mean[H], var[H]) pairs inside the BN moving-statistics buffer."""
  if not cfg.batchnorm:
    return []
  names = [f"enc.{i}" for i in range(cfg.n_enc_layers)]
  if cfg.model_kind == MODEL_SCVI:
    names += [f"encl.{i}" for i in range(cfg.n_encl_layers)]
  names += [f"dec.{i}" for i in range(cfg.n_dec_layers)]
  return names
