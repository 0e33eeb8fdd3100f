"""ctypes binding of include/sisua_b200.h.  There is no CPU fallback: if the shared library is
missing or a call fails the error is raised."""
from __future__ import annotations

import ctypes
import os

from .config import StepConfig

_LIB = None

EXPORTS = ["sisua_create", "sisua_destroy", "sisua_param_layout", "sisua_bind_buffers", "sisua_train_step", "sisua_train_step_gather",
           "sisua_train_step_gather_u16", "sisua_widen_rows_u16", "sisua_unpack_counts_csr_u16", "sisua_unpack_counts_csr8_u16", "sisua_infer", "sisua_infer_ex", "sisua_forward_train_mode", "sisua_decode", "sisua_marginal_llk", "sisua_adam_step", "sisua_dp_bind", "sisua_adam_step_dp", "sisua_dp_shard", "sisua_debug_buffer", "sisua_debug_copy", "sisua_debug_geometry", "sisua_debug_force_chunks", "sisua_launch_count", "sisua_set_step", "sisua_set_infer_seed", "sisua_set_count_bound", "sisua_nonfinite_flag", "sisua_corrupt_counts", "sisua_set_grad_ready_event", "sisua_unpack_counts_u16", "sisua_unpack_counts_csr", "sisua_train_step_host", "sisua_tc_selftest", "sisua_profile_enable", "sisua_profile_read", "sisua_last_error", "sisua_version"]


class ParamDesc(ctypes.Structure):
  _fields_ = [("name", ctypes.c_char * 32), ("offset", ctypes.c_int64), ("rows", ctypes.c_int32),
              ("cols", ctypes.c_int32), ("ld", ctypes.c_int32), ("kind", ctypes.c_int32)]


class HostBatch(ctypes.Structure):
  """sisua_host_batch (include/sisua_b200.h): one minibatch in host memory."""
  _fields_ = [("format", ctypes.c_int32), ("B", ctypes.c_int32), ("x", ctypes.c_void_p), ("indptr", ctypes.c_void_p),
              ("cols", ctypes.c_void_p), ("vals", ctypes.c_void_p), ("nnz", ctypes.c_int64), ("y", ctypes.c_void_p),
              ("library", ctypes.c_void_p), ("mask", ctypes.c_void_p), ("eps_z", ctypes.c_void_p), ("eps_l", ctypes.c_void_p)]


HOST_F32, HOST_U16, HOST_CSR = 0, 1, 2


class SisuaError(RuntimeError):
  pass


def lib_path() -> str:
  return os.path.join(os.path.dirname(os.path.abspath(__file__)), "libsisua_b200.so")


def load():
  global _LIB
  if _LIB is not None:
    return _LIB
  path = lib_path()
  if not os.path.exists(path):
    raise SisuaError(f"{path} is missing: run `python -m sisua_b200.build` (there is no CPU fallback)")
  L = ctypes.CDLL(path)
  vp, ci, cf = ctypes.c_void_p, ctypes.c_int, ctypes.c_float
  L.sisua_create.argtypes = [ctypes.POINTER(StepConfig), ci, ctypes.POINTER(vp)]
  L.sisua_create.restype = ci
  L.sisua_destroy.argtypes = [vp]
  L.sisua_destroy.restype = ci
  L.sisua_param_layout.argtypes = [vp, ctypes.POINTER(ParamDesc), ctypes.POINTER(ci), ctypes.POINTER(ctypes.c_int64)]
  L.sisua_param_layout.restype = ci
  L.sisua_bind_buffers.argtypes = [vp, vp, vp, vp, vp, vp]
  L.sisua_bind_buffers.restype = ci
  L.sisua_train_step.argtypes = [vp, vp, vp, vp, vp, vp, vp, ci, ctypes.c_uint64, ctypes.c_int64, vp, vp, vp]
  L.sisua_train_step.restype = ci
  L.sisua_train_step_gather.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp, ci, ctypes.c_uint64, ctypes.c_int64, vp, vp, vp]
  L.sisua_train_step_gather.restype = ci
  L.sisua_train_step_gather_u16.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp, ci, ctypes.c_uint64, ctypes.c_int64, vp, vp, vp]
  L.sisua_train_step_gather_u16.restype = ci
  L.sisua_widen_rows_u16.argtypes = [vp, vp, vp, vp, ci, vp]
  L.sisua_widen_rows_u16.restype = ci
  L.sisua_unpack_counts_csr_u16.argtypes = [vp, vp, vp, vp, vp, ci, vp]
  L.sisua_unpack_counts_csr_u16.restype = ci
  L.sisua_unpack_counts_csr8_u16.argtypes = [vp, vp, vp, vp, vp, vp, ci, vp]
  L.sisua_unpack_counts_csr8_u16.restype = ci
  L.sisua_infer.argtypes = [vp, vp, vp, vp, vp, vp, vp, ci, ci] + [vp] * 10
  L.sisua_infer.restype = ci
  L.sisua_infer_ex.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp, ci, ci, ci] + [vp] * 12
  L.sisua_infer_ex.restype = ci
  L.sisua_marginal_llk.argtypes = [vp, vp, vp, vp, vp, vp, vp, ci, ci, vp, vp, vp, vp]
  L.sisua_marginal_llk.restype = ci
  L.sisua_forward_train_mode.argtypes = [vp, vp, vp, vp, vp, vp, vp, ci, ctypes.c_uint64, ctypes.c_int64] + [vp] * 10
  L.sisua_forward_train_mode.restype = ci
  L.sisua_decode.argtypes = [vp, vp, vp, ci, vp, vp, vp, vp, vp]
  L.sisua_decode.restype = ci
  L.sisua_adam_step.argtypes = [vp, cf, cf, cf, cf, cf, cf, ctypes.c_int64, vp]
  L.sisua_adam_step.restype = ci
  L.sisua_dp_bind.argtypes = [vp, ci, ci, ctypes.POINTER(vp), ctypes.POINTER(vp), ctypes.POINTER(vp), ctypes.POINTER(vp), ci]
  L.sisua_dp_bind.restype = ci
  L.sisua_adam_step_dp.argtypes = [vp, cf, cf, cf, cf, cf, ctypes.c_int64, vp]
  L.sisua_adam_step_dp.restype = ci
  L.sisua_dp_shard.argtypes = [vp, ctypes.POINTER(ctypes.c_int64), ctypes.POINTER(ctypes.c_int64)]
  L.sisua_dp_shard.restype = ci
  L.sisua_debug_buffer.argtypes = [vp, ctypes.c_char_p]
  L.sisua_debug_buffer.restype = vp
  L.sisua_debug_copy.argtypes = [vp, ctypes.c_char_p, vp, ctypes.c_int64, vp]
  L.sisua_debug_copy.restype = ci
  L.sisua_debug_geometry.argtypes = [vp, ci, ctypes.POINTER(ctypes.c_int32)]
  L.sisua_debug_geometry.restype = ci
  L.sisua_debug_force_chunks.argtypes = [vp, ci, ci, ci]
  L.sisua_debug_force_chunks.restype = ci
  L.sisua_launch_count.argtypes = [vp]
  L.sisua_launch_count.restype = ctypes.c_int64
  L.sisua_profile_enable.argtypes = [vp, ci]
  L.sisua_profile_enable.restype = ci
  L.sisua_profile_read.argtypes = [vp, ctypes.POINTER(cf), ctypes.POINTER(ci)]
  L.sisua_profile_read.restype = ci
  L.sisua_tc_selftest.argtypes = [vp, vp, vp, ci, ci, ci, ci, vp]
  L.sisua_tc_selftest.restype = ci
  L.sisua_unpack_counts_u16.argtypes = [vp, vp, vp, ctypes.c_int64, vp]
  L.sisua_unpack_counts_u16.restype = ci
  L.sisua_set_grad_ready_event.argtypes = [vp, vp]
  L.sisua_set_grad_ready_event.restype = ci
  L.sisua_unpack_counts_csr.argtypes = [vp, vp, vp, vp, vp, ci, vp]
  L.sisua_unpack_counts_csr.restype = ci
  L.sisua_train_step_host.argtypes = [vp, ctypes.POINTER(HostBatch), ctypes.c_uint64, ctypes.c_int64, vp, vp, vp]
  L.sisua_train_step_host.restype = ci
  L.sisua_corrupt_counts.argtypes = [vp, vp, vp, ctypes.c_int64, ci, ctypes.c_int64, ctypes.c_int64, cf, cf, ci, ctypes.c_uint64, vp]
  L.sisua_corrupt_counts.restype = ci
  L.sisua_nonfinite_flag.argtypes = [vp, ci]
  L.sisua_nonfinite_flag.restype = ci
  L.sisua_set_count_bound.argtypes = [vp, cf]
  L.sisua_set_count_bound.restype = ci
  L.sisua_set_infer_seed.argtypes = [vp, ctypes.c_uint64, ctypes.c_int64]
  L.sisua_set_infer_seed.restype = ci
  L.sisua_set_step.argtypes = [vp, ctypes.c_int64, vp]
  L.sisua_set_step.restype = ci
  L.sisua_last_error.argtypes = [vp]
  L.sisua_last_error.restype = ctypes.c_char_p
  L.sisua_version.argtypes = []
  L.sisua_version.restype = ctypes.c_char_p
  _LIB = L
  return L
