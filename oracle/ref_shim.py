"""Import the UNMODIFIED reference modules (TEST INFRASTRUCTURE).

From /root/reference where it is mounted (the authoring container), else from
``oracle/_ref`` — the verbatim copy ``oracle/make_ref.py`` makes there, which is
git-ignored but travels to the GPU box.  Nothing on the GPU box reads
/root/reference.  Two shims are needed:
  * vit.py:3 imports ``torchsummary`` (unused, not installed);
  * layers.py:12 -> nnmf/optimizer.py:8 imports the private
    ``torch.optim.optimizer._dispatch_sqrt`` removed in torch 2.11.
"""
from __future__ import annotations

import math
import os
import sys
import types

_COPY = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")
REFERENCE_ROOT = os.environ.get("VITB_REFERENCE_ROOT", "/root/reference")
if not os.path.isfile(os.path.join(REFERENCE_ROOT, "vit.py")) and os.path.isfile(os.path.join(_COPY, "vit.py")):
    REFERENCE_ROOT = _COPY


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
