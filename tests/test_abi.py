"""CPU checks of the drop-in boundary: the C-ABI library builds/loads without a GPU, exports
every symbol include/nvfi_b200.h declares, and the ctypes struct mirrors have the C layout.
No compute call is made here (there is no GPU in the build container)."""
import ctypes as C
import os
import re
import subprocess
import tempfile

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "nvfi_b200.h")


def _declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(nvfi_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from nvfi_b200 import _lib, build
    build.build()
    lib = C.CDLL(_lib.LIB_PATH)
    names = _declared_symbols()
    assert len(names) >= 20
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing
    # the binding covers the whole header
    unbound = [n for n in names if n not in _lib.SIGNATURES]
    assert not unbound, unbound
    assert lib.nvfi_abi_version() == _lib.ABI_VERSION


def test_abi_version_matches_header():
    from nvfi_b200 import _lib
    m = re.search(r"#define\s+NVFI_ABI_VERSION\s+(\d+)", open(HEADER).read())
    assert int(m.group(1)) == _lib.ABI_VERSION


def test_struct_layouts_match_c():
    """sizeof/offsetof of the ctypes mirrors against a C program compiled from the header."""
    from nvfi_b200 import _lib
    if not any(os.access(os.path.join(p, "gcc"), os.X_OK) for p in os.environ.get("PATH", "").split(":")):
        pytest.skip("gcc not available")
    probes = [("NvfiLinear", _lib.NvfiLinear, ["umma", "ummaT", "in_dim", "ummaT_rows"]),
              ("NvfiField", _lib.NvfiField, ["grid", "dplane_space", "basis_mat", "vel_net", "gate_lo",
                                             "alpha_volume", "mask_net", "mlp_mode"]),
              ("NvfiRenderArgs", _lib.NvfiRenderArgs, ["jitter", "chunk_bg", "t", "advect"]),
              ("NvfiRenderBuffers", _lib.NvfiRenderBuffers, ["x_adv", "sigma", "stats"]),
              ("NvfiRenderGrads", _lib.NvfiRenderGrads, ["g_basis_mat", "g_vel_b", "workspace_bytes"]),
              ("NvfiParamGrads", _lib.NvfiParamGrads, ["aplane_time", "basis_mat", "render_b", "vel_b"]),
              ("NvfiPdeGrads", _lib.NvfiPdeGrads, ["g_acc_w", "g_acc_pts", "workspace_bytes"]),
              ("NvfiProfileEntry", _lib.NvfiProfileEntry, ["ms", "launches"])]
    lines = ['#include <stdio.h>', '#include <stddef.h>', f'#include "{HEADER}"', "int main(void){"]
    for name, _, fields in probes:
        lines.append(f'printf("%zu\\n", sizeof({name}));')
        for f in fields:
            lines.append(f'printf("%zu\\n", offsetof({name}, {f}));')
    lines.append("return 0;}")
    with tempfile.TemporaryDirectory() as d:
        src, exe = os.path.join(d, "p.c"), os.path.join(d, "p")
        open(src, "w").write("\n".join(lines))
        subprocess.run(["gcc", "-std=c99", src, "-o", exe], check=True)
        got = [int(x) for x in subprocess.run([exe], capture_output=True, text=True, check=True).stdout.split()]
    want = []
    for _, cls, fields in probes:
        want.append(C.sizeof(cls))
        want += [getattr(cls, f).offset for f in fields]
    assert got == want


def test_product_path_fails_loudly_without_library(tmp_path):
    from nvfi_b200 import _lib
    with pytest.raises(RuntimeError, match="no CPU fallback|not found"):
        _lib.load(str(tmp_path / "missing.so"))


def test_product_path_never_imports_oracle():
    """oracle/ is test infrastructure: nothing under nvfi_b200/ may import it."""
    pkg = os.path.join(ROOT, "nvfi_b200")
    bad = []
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                txt = open(os.path.join(dp, f)).read()
                if re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M):
                    bad.append(os.path.join(dp, f))
    assert not bad, bad
