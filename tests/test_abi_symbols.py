"""CPU-side checks of the drop-in boundary: the shared library loads and exports every function
include/sisua_b200.h declares; the ctypes StepConfig mirrors the C struct."""
import ctypes
import os
import re

from sisua_b200 import _lib, build, config as C

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header():
  return open(os.path.join(ROOT, "include", "sisua_b200.h")).read()


def test_library_exports_every_declared_symbol():
  build.build()
  L = _lib.load()
  names = set(re.findall(r"\b(sisua_[a-z0-9_]+)\s*\(", _header()))
  assert {"sisua_create", "sisua_train_step", "sisua_infer", "sisua_adam_step"} <= names
  for n in names:
    assert hasattr(L, n), n
  assert set(_lib.EXPORTS) <= names
  assert b"sm_100a" in L.sisua_version()


def test_step_config_matches_header_field_order():
  body = re.search(r"typedef struct sisua_step_config \{(.*?)\} sisua_step_config;", _header(), re.S).group(1)
  body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
  fields = []
  for decl in body.split(";"):
    decl = decl.strip()
    if not decl:
      continue
    typ, rest = decl.split(None, 1)
    for name in rest.split(","):
      fields.append((name.strip(), typ))
  py = [(n, "int32_t" if t is ctypes.c_int32 else "float") for n, t in C.StepConfig._fields_]
  assert fields == py
  assert ctypes.sizeof(C.StepConfig) == 4 * len(py)


def test_create_without_gpu_or_bad_config_fails_loudly():
  import pytest
  import torch
  from sisua_b200.engine import Engine
  cfg = C.make_step_config("vae", n_genes=32)
  if not torch.cuda.is_available():
    with pytest.raises(_lib.SisuaError):
      Engine(cfg)
  with pytest.raises(ValueError):
    C.make_step_config("vae", n_genes=32, n_hidden=48)
  with pytest.raises(ValueError):
    C.make_step_config("vae", n_genes=32, x_dist="poisson")


def test_param_layout_properties():
  for model, kw in [("vae", {}), ("scvi", {}), ("dca", {}), ("sisua", dict(n_proteins=10))]:
    cfg = C.make_step_config(model, n_genes=558, **kw)
    entries, total = C.param_layout(cfg)
    offs = [e.offset for e in entries]
    assert offs == sorted(offs) and all(o % 64 == 0 for o in offs)
    for a, b in zip(entries[:-1], entries[1:]):
      assert a.offset + a.size <= b.offset
    assert entries[-1].offset + entries[-1].size <= total
    names = [e.name for e in entries]
    assert len(set(names)) == len(names)
    if model == "scvi":   # the two first-layer operands must be adjacent ([2H, G] streamed once)
      e0, e1 = entries[0], entries[1]
      assert (e0.name, e1.name) == ("enc.0.W", "encl.0.W") and e1.offset == e0.offset + e0.size


def test_ctypes_structs_match_the_c_header_layout(tmp_path):
  """sizeof / offsetof of every struct that crosses the C ABI, as gcc sees include/sisua_b200.h, equal the ctypes
  mirrors (sisua_step_config, sisua_param_desc, sisua_host_batch)."""
  import shutil
  import subprocess
  import pytest
  gcc = shutil.which("gcc")
  if gcc is None:
    pytest.skip("gcc not available")
  pairs = [("sisua_step_config", C.StepConfig), ("sisua_param_desc", _lib.ParamDesc), ("sisua_host_batch", _lib.HostBatch)]
  lines = ['#include <stdio.h>', '#include <stddef.h>', f'#include "{os.path.join(ROOT, "include", "sisua_b200.h")}"',
           'int main(void) {']
  for cname, st in pairs:
    lines.append(f'  printf("{cname} %zu\\n", sizeof({cname}));')
    for fname, _ in st._fields_:
      lines.append(f'  printf("{cname}.{fname} %zu\\n", offsetof({cname}, {fname}));')
  lines += ['  return 0;', '}']
  src = tmp_path / "layout.c"
  src.write_text("\n".join(lines))
  exe = tmp_path / "layout"
  subprocess.run([gcc, "-std=c99", "-o", str(exe), str(src)], check=True)
  got = dict(l.split() for l in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.splitlines())
  for cname, st in pairs:
    assert int(got[cname]) == ctypes.sizeof(st), cname
    for fname, _ in st._fields_:
      assert int(got[f"{cname}.{fname}"]) == getattr(st, fname).offset, f"{cname}.{fname}"


def test_product_never_touches_the_oracle():
  """The oracle is test infrastructure: no module of the package imports it (comments may cite it), and importing the
  package does not pull it in."""
  import ast
  import pathlib
  import subprocess
  import sys
  root = pathlib.Path(__file__).resolve().parents[1]
  for path in (root / "sisua_b200").glob("*.py"):
    tree = ast.parse(path.read_text())
    for node in ast.walk(tree):
      names = []
      if isinstance(node, ast.Import):
        names = [a.name for a in node.names]
      elif isinstance(node, ast.ImportFrom):
        names = [node.module or ""]
      assert not any(n == "oracle" or n.startswith("oracle.") for n in names), f"{path.name} imports the oracle"
  code = "import sys; import sisua_b200.models, sisua_b200.posterior, sisua_b200.pipeline, sisua_b200.streamed; " \
         "print(any(m == 'oracle' or m.startswith('oracle.') for m in sys.modules))"
  out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd=str(root))
  assert out.returncode == 0 and out.stdout.strip() == "False", out.stdout + out.stderr
