"""Flat parameter storage.

All parameters of a model (or of a stand-alone block) live in ONE contiguous fp32 buffer; every
``nn.Parameter`` is a view into it.  That gives: a fused QKV weight ([Wq;Wk;Wv] adjacent -> one
(3H,H) GEMM operand), a single-launch Adam over the whole model, one bf16
shadow buffer for the tensor-core GEMMs, and contiguous per-layer gradient buckets for the all-reduce.
The ``state_dict`` names/shapes stay those of the reference (vit.py:44-63, layers.py:26-39, 81-85).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn as nn

ALIGN = 64  # elements: 256 B for fp32, 128 B for the bf16 shadow (TMA needs 16 B)

LAYER_FIELDS = ("ln1_w", "ln1_b", "wqkv", "bqkv", "wo", "bo", "ln2_w", "ln2_b", "w1", "b1", "w2", "b2")


@dataclass
class Slot:
    off: int
    numel: int
    shape: Tuple[int, ...]


def _round_up(x: int, a: int = ALIGN) -> int:
    return (x + a - 1) // a * a


class FlatLayout:
    """Ordered (name -> Slot) map with aligned offsets; `active_end` = end of the optimised region."""

    def __init__(self, entries: Sequence[Tuple[str, Tuple[int, ...], bool]]):
        self.slots: Dict[str, Slot] = {}
        off = 0
        inactive = []
        for name, shape, active in entries:
            if not active:
                inactive.append((name, shape))
                continue
            n = int(torch.Size(shape).numel())
            self.slots[name] = Slot(off, n, tuple(shape))
            off = _round_up(off + n)
        self.active_end = off
        for name, shape in inactive:  # parameters the forward never touches (Adam skips them, like torch)
            n = int(torch.Size(shape).numel())
            self.slots[name] = Slot(off, n, tuple(shape))
            off = _round_up(off + n)
        self.total = off

    def view(self, buf: torch.Tensor, name: str) -> torch.Tensor:
        s = self.slots[name]
        return buf[s.off:s.off + s.numel].view(s.shape)

    def span(self, buf: torch.Tensor, first: str, last: str, shape: Tuple[int, ...]) -> torch.Tensor:
        """One view over several adjacent slots (e.g. Wq.weight..Wv.weight -> (3H,H)); checks adjacency."""
        a, b = self.slots[first], self.slots[last]
        n = int(torch.Size(shape).numel())
        if b.off + b.numel - a.off != n:
            raise AssertionError(f"slots {first}..{last} are not densely adjacent ({b.off + b.numel - a.off} != {n})")
        return buf[a.off:a.off + n].view(shape)


def layer_entries(prefix: str, H: int, M: int, use_mlp: bool, with_ln: bool = True,
                  attn_prefix: Optional[str] = None) -> List[Tuple[str, Tuple[int, ...], bool]]:
    """Storage order of one TransformerEncoder's tensors (not the registration order).
    with_ln=False + attn_prefix="" describes a stand-alone MultiHeadSelfAttention."""
    a = prefix + "attention." if attn_prefix is None else attn_prefix
    e: List[Tuple[str, Tuple[int, ...], bool]] = []
    if with_ln:
        e += [(prefix + "la1.weight", (H,), True), (prefix + "la1.bias", (H,), True)]
    e += [(a + "Wq.weight", (H, H), True), (a + "Wk.weight", (H, H), True), (a + "Wv.weight", (H, H), True),
          (a + "Wq.bias", (H,), True), (a + "Wk.bias", (H,), True), (a + "Wv.bias", (H,), True),
          (a + "out_project.weight", (H, H), True), (a + "out_project.bias", (H,), True)]
    if with_ln:
        # la2 exists even without an MLP (layers.py:30) but is then never used (layers.py:46)
        e += [(prefix + "la2.weight", (H,), use_mlp), (prefix + "la2.bias", (H,), use_mlp)]
        if use_mlp:
            e += [(prefix + "mlp.0.weight", (M, H), True), (prefix + "mlp.0.bias", (M,), True),
                  (prefix + "mlp.3.weight", (H, M), True), (prefix + "mlp.3.bias", (H,), True)]
    return e


class LayerViews:
    """Views of one encoder layer inside a flat buffer (fp32 master, compute-dtype copy, or grads)."""

    __slots__ = LAYER_FIELDS

    def __init__(self, layout: FlatLayout, buf: torch.Tensor, prefix: str, H: int, M: int, use_mlp: bool, with_ln: bool = True,
                 attn_prefix: Optional[str] = None):
        a = prefix + "attention." if attn_prefix is None else attn_prefix
        v = layout.view
        self.ln1_w = v(buf, prefix + "la1.weight") if with_ln else None
        self.ln1_b = v(buf, prefix + "la1.bias") if with_ln else None
        self.wqkv = layout.span(buf, a + "Wq.weight", a + "Wv.weight", (3 * H, H))
        self.bqkv = layout.span(buf, a + "Wq.bias", a + "Wv.bias", (3 * H,))
        self.wo = v(buf, a + "out_project.weight")
        self.bo = v(buf, a + "out_project.bias")
        self.ln2_w = v(buf, prefix + "la2.weight") if with_ln else None
        self.ln2_b = v(buf, prefix + "la2.bias") if with_ln else None
        if use_mlp and with_ln:
            self.w1 = v(buf, prefix + "mlp.0.weight")
            self.b1 = v(buf, prefix + "mlp.0.bias")
            self.w2 = v(buf, prefix + "mlp.3.weight")
            self.b2 = v(buf, prefix + "mlp.3.bias")
        else:
            self.w1 = self.b1 = self.w2 = self.b2 = None


class FlatStore:
    """Owns the flat fp32 parameter buffer of a module tree and (lazily) its bf16 shadow."""

    def __init__(self, module: nn.Module, layout: FlatLayout):
        self.layout = layout
        named = dict(module.named_parameters())
        missing = [k for k in layout.slots if k not in named]
        extra = [k for k in named if k not in layout.slots]
        if missing or extra:
            raise AssertionError(f"layout/module mismatch: missing={missing} extra={extra}")
        dev = next(iter(named.values())).device
        self.device = dev
        self.flat = torch.zeros(layout.total, dtype=torch.float32, device=dev)
        with torch.no_grad():
            for k, s in layout.slots.items():
                p = named[k]
                dst = layout.view(self.flat, k)
                dst.copy_(p.detach().to(torch.float32))
                p.data = dst  # the nn.Parameter is now a view of the flat buffer
        self._shadow: Optional[torch.Tensor] = None
        self._first = named[next(iter(layout.slots))]
        self._last = named[list(layout.slots)[-1]]
        self._first_off = layout.slots[next(iter(layout.slots))].off
        self._last_off = layout.slots[list(layout.slots)[-1]].off

    def consistent(self) -> bool:
        """Are the module's parameters still views of this buffer (False after .to()/.cuda()/re-assignment)?"""
        base = self.flat.data_ptr()
        return (self._first.data_ptr() == base + 4 * self._first_off and self._last.data_ptr() == base + 4 * self._last_off
                and self._first.dtype == torch.float32)

    def shadow(self) -> torch.Tensor:
        if self._shadow is None:
            self._shadow = torch.zeros(self.layout.total, dtype=torch.bfloat16, device=self.device)
        return self._shadow
