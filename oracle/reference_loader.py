"""Locate and import the UNMODIFIED reference (vLAR-group/NVFi) for the CPU arm of bench.py and for
tests that compare against it.  Test / measurement infrastructure only: nothing under nvfi_b200/ may
import this module.

The reference is pure Python without a setup.py, so "installing" it is a copy of its packages:
``python tools/install_reference.py`` copies models/, utils/ and config/ from /root/reference into
baseline/_ref/ (git-ignored, but shipped to the GPU box by gpurun).  On the GPU box /root/reference
does not exist; baseline/_ref is the only copy there.  When neither is present the callers fall back
to the oracle port (oracle/nvfi_oracle.py), which tests/golden/* pin to the reference.
"""
from __future__ import annotations

import json
import os
import sys
import types
from typing import Optional

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
INSTALLED = os.path.join(ROOT, "baseline", "_ref")


def find_reference(allow_checkout: bool = True) -> Optional[str]:
    """Directory holding the reference's ``models`` package, or None."""
    cands = [os.environ.get("NVFI_REFERENCE"), INSTALLED]
    if allow_checkout:
        cands.append("/root/reference")
    for c in cands:
        if c and os.path.isfile(os.path.join(c, "models", "nvfi.py")):
            return c
    return None


def import_reference(path: str):
    """Returns the reference's ``models`` package and ``CfgNode`` (the metric /
    image libraries the hot path never touches are stubbed)."""
    for m in ("lpips", "imageio", "matplotlib", "matplotlib.pyplot"):
        sys.modules.setdefault(m, types.ModuleType(m))
    if path not in sys.path:
        sys.path.insert(0, path)
    import models as ref_models          # noqa: E402
    from utils import CfgNode            # noqa: E402
    if not os.path.abspath(ref_models.__file__).startswith(os.path.abspath(path)):
        raise RuntimeError(f"a different 'models' package is already imported: {ref_models.__file__}")
    return ref_models, CfgNode


def build_reference(path: str, cfg, grid, K: int, sd, device: str = "cpu"):
    """The reference's NVFi carrying the synthetic parameters `sd` (nvfi_b200.synth.synth_state)."""
    from nvfi_b200 import synth
    ref_models, CfgNode = import_reference(path)
    c = CfgNode(json.loads(json.dumps(cfg)))
    c.nvfi.num_keyframes = K
    aabb = synth.aabb_from_cfg(cfg)
    nv = ref_models.NVFi(c, device, aabb, list(grid), [cfg.dataset.near, cfg.dataset.far])
    missing, unexpected = nv.load_state_dict({"nvfi." + k: v for k, v in sd.items()}, strict=False)
    assert not unexpected, unexpected
    bad = [m for m in missing if "frequency_bands" not in m and ".vel.vel_net." not in m and m != "nvfi.aabb"]
    assert not bad, bad
    return ref_models, c, nv
