"""Epoch-level pieces around the hot step: the reference's learning-rate schedule, its
validation step and its checkpoint format.  Host code only.

Learning rate — network.py:113-122: `CosineAnnealingLR(T_max=max_epochs, eta_min=min_lr)` wrapped in
`warmup_scheduler.GradualWarmupScheduler(multiplier=1.0, total_epoch=warmup_epoch)`, stepped once per epoch by Lightning.
The wrapper is the un-vendored dependency ildoonet/pytorch-gradual-warmup-lr (git HEAD, setup.sh:5); its published
algorithm with multiplier == 1.0 is restated here in closed form (tests/test_host_logic.py replays the library's class
against torch's real CosineAnnealingLR and compares):

    e = number of scheduler steps taken so far = index of the epoch about to be trained
    e <= W            lr = base * e / W              (epoch 0 trains with lr = 0, as upstream does)
    e == W + 1        lr = base                      (the cosine schedule's own epoch 0)
    e >  W + 1        lr = min_lr + (base - min_lr) * (1 + cos(pi * (e - W - 1) / T_max)) / 2
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import torch


def warmup_cosine_lr(epoch: int, base_lr: float, min_lr: float, max_epochs: int, warmup_epoch: int) -> float:
    """Learning rate in force while epoch `epoch` (0-based) is trained."""
    if epoch < 0:
        raise ValueError("epoch must be >= 0")
    W = int(warmup_epoch)
    if W > 0 and epoch <= W:
        return base_lr * (float(epoch) / W)
    # upstream divides by zero for warmup_epoch == 0; here that simply means "no warm-up": the cosine schedule from epoch 0
    t = epoch - (W + 1) if W > 0 else epoch
    return min_lr + (base_lr - min_lr) * (1.0 + math.cos(math.pi * t / max_epochs)) / 2.0


class WarmupCosine:
    """`engine.set_lr(sched(epoch))` once per epoch reproduces network.py:113-122 for a TrainEngine / FusedAdam."""

    def __init__(self, base_lr: float = 1e-3, min_lr: float = 1e-5, max_epochs: int = 200, warmup_epoch: int = 5):
        self.base_lr, self.min_lr, self.max_epochs, self.warmup_epoch = float(base_lr), float(min_lr), int(max_epochs), int(warmup_epoch)

    def __call__(self, epoch: int) -> float:
        return warmup_cosine_lr(epoch, self.base_lr, self.min_lr, self.max_epochs, self.warmup_epoch)


@torch.no_grad()
def evaluate(model, criterion, batches) -> Dict[str, float]:
    """network.py:388-395 over an iterable of (img, label): mean loss and accuracy (each batch weighted by its size)."""
    was_training = model.training
    model.eval()
    dev = next(model.parameters()).device
    n, loss_sum, correct = 0, 0.0, 0
    for img, label in batches:
        img, label = img.to(dev, non_blocking=True), label.to(dev, non_blocking=True)
        out = model(img)
        loss_sum += float(criterion(out, label)) * img.shape[0]
        correct += int(torch.eq(out.argmax(-1), label).sum())
        n += img.shape[0]
    model.train(was_training)
    return {"val_loss": loss_sum / max(n, 1), "val_acc": correct / max(n, 1), "n": n}


# ---------------------------------------------------------------------------------------------
# checkpoints: what `trainer.save_checkpoint` writes (main.py:234-237) and run_model.py:12-37 reads — a dict with
# "state_dict" (keys prefixed by the LightningModule attribute, "model.") and "hyper_parameters"
# ---------------------------------------------------------------------------------------------
def to_lightning_checkpoint(model, hyper_parameters: Optional[dict] = None, prefix: str = "model.") -> dict:
    sd = {prefix + k: v.detach().to("cpu", copy=True) for k, v in model.state_dict().items()}
    return {"state_dict": sd, "hyper_parameters": dict(hyper_parameters or {})}


def save_checkpoint(model, path: str, hyper_parameters: Optional[dict] = None) -> None:
    torch.save(to_lightning_checkpoint(model, hyper_parameters), path)


def load_checkpoint(model, ckpt, strict: bool = False, prefix: str = "model."):
    """`ckpt`: a path or an already loaded dict in the reference's format (or a bare state_dict).  Keys are matched after
    stripping `prefix`; with strict=False (run_model.py:37) unknown / missing keys are reported, not fatal."""
    if isinstance(ckpt, (str, bytes)) or hasattr(ckpt, "__fspath__"):
        ckpt = torch.load(ckpt, map_location="cpu")
    sd = ckpt.get("state_dict", ckpt)
    sd = {(k[len(prefix):] if k.startswith(prefix) else k): v for k, v in sd.items()}
    return model.load_state_dict(sd, strict=strict)
