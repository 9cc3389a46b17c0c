"""Import the UNMODIFIED reference modules from /root/reference (TEST INFRASTRUCTURE).

Only usable where the reference is mounted (the authoring container); the GPU
box has no /root/reference, so nothing under ``-m gpu``, ``smoke()`` or
``bench.py`` may call this.  Two shims are needed:
  * vit.py:3 imports ``torchsummary`` (unused, not installed);
  * layers.py:12 -> nnmf/optimizer.py:8 imports the private
    ``torch.optim.optimizer._dispatch_sqrt`` removed in torch 2.11.
"""
from __future__ import annotations

import math
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("VITB_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "vit.py"))


def import_reference():
    """Returns (vit, layers, criterions) reference modules."""
    if not reference_available():
        raise FileNotFoundError(f"reference not mounted at {REFERENCE_ROOT}")
    import torch
    import torch.optim.optimizer as opt_mod

    sys.modules.setdefault("torchsummary", types.ModuleType("torchsummary"))
    if not hasattr(opt_mod, "_dispatch_sqrt"):
        opt_mod._dispatch_sqrt = lambda x: x.sqrt() if torch.is_tensor(x) else math.sqrt(x)
    # The product package also ships modules called `vit`, `layers`, `criterions`
    # inside its own namespace; the reference's are top-level, so there is no clash.
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import vit as ref_vit  # noqa: E402
    import layers as ref_layers  # noqa: E402
    import criterions as ref_criterions  # noqa: E402

    assert os.path.dirname(ref_vit.__file__) == REFERENCE_ROOT, ref_vit.__file__
    return ref_vit, ref_layers, ref_criterions
