"""Generate golden fixtures from the UNMODIFIED reference (run in the authoring container only).

    python tests/golden/make_golden.py

Imports ``vit.ViT`` / ``criterions.LabelSmoothingCrossEntropyLoss`` from
/root/reference (via oracle/ref_shim.py), loads hash-derived weights and inputs
(oracle.hash_init_ / oracle.hash_inputs: integer arithmetic, reproducible bit for
bit anywhere), runs forward + LS-CE + backward + torch.optim.Adam (the optimiser
network.py:71-77 configures) in fp32 on CPU and stores the results.

"full" fixtures store every tensor; "summary" fixtures (the 6.3 M-parameter
model) store logits, loss and per-tensor norms + leading elements to stay small.
"""
from __future__ import annotations

import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import ViTConfig, hash_init_, hash_inputs  # noqa: E402
from oracle.ref_shim import import_reference  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))

ADAM = dict(lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=5e-5)  # main.py:48-56
SMOOTHING = 0.1  # main.py:62

CASES = {
    # name: (cfg kwargs, batch, mode)
    "tiny65": (dict(num_classes=10, img_size=32, patch=8, num_layers=2, hidden=128, mlp_hidden=128, head=4), 4, "full"),
    "tiny17c100": (dict(num_classes=100, img_size=32, patch=4, num_layers=1, hidden=128, mlp_hidden=256, head=2), 3, "full"),
    "nocls_nomlp": (dict(num_classes=10, img_size=32, patch=4, num_layers=1, hidden=128, mlp_hidden=128, head=4,
                         is_cls_token=False, encoder_mlp=False), 2, "full"),
    "full65": (dict(num_classes=10, img_size=32, patch=8, num_layers=7, hidden=384, mlp_hidden=384, head=12), 4, "summary"),
    "full17c100": (dict(num_classes=100, img_size=32, patch=4, num_layers=7, hidden=384, mlp_hidden=384, head=12), 4, "summary"),
}


def summarize(d):
    return {k: dict(norm=v.double().norm().item(), sum=v.double().sum().item(), head=v.flatten()[:16].clone())
            for k, v in d.items()}


def main():
    ref_vit, ref_layers, ref_crit = import_reference()
    torch.set_num_threads(1)  # fixed reduction order
    for name, (kw, batch, mode) in CASES.items():
        cfg = ViTConfig(**kw)
        model = ref_vit.ViT(
            3, cfg.num_classes, img_size=cfg.img_size, patch=cfg.patch, dropout=0.0,
            num_layers=cfg.num_layers, hidden=cfg.hidden, encoder_mlp=cfg.encoder_mlp,
            mlp_hidden=cfg.mlp_hidden, head=cfg.head, is_cls_token=cfg.is_cls_token)
        sd = model.state_dict()
        assert list(sd.keys()) == list(cfg.param_shapes().keys()), (list(sd.keys()), list(cfg.param_shapes().keys()))
        hash_init_(sd, seed=0)  # in place -> model weights
        x, y = hash_inputs(cfg, batch, seed=1)
        crit = ref_crit.LabelSmoothingCrossEntropyLoss(cfg.num_classes, smoothing=SMOOTHING)
        opt = torch.optim.Adam(model.parameters(), **ADAM)

        # attention maps via the reference's own save_attn_map protocol (layers.py:50-65)
        for m in model.modules():
            if hasattr(m, "save_attn_map"):
                m.save_attn_map = True
        logits = model(x)
        attn = torch.stack([blk.get_attention_map().detach() for blk in model.enc])
        loss = crit(logits, y)
        opt.zero_grad()
        loss.backward()
        # parameters the forward never touches (la2 when encoder_mlp=False, layers.py:30,46) keep grad None
        # and torch's Adam skips them; stored as zeros.
        grads = {k: (p.grad.detach().clone() if p.grad is not None else torch.zeros_like(p))
                 for k, p in model.named_parameters()}
        losses = [loss.item()]
        opt.step()
        for _ in range(2):  # two more steps on the same batch -> params after 3 Adam steps
            opt.zero_grad()
            l2 = crit(model(x), y)
            l2.backward()
            losses.append(l2.item())
            opt.step()
        params3 = {k: v.detach().clone() for k, v in model.state_dict().items()}

        out = dict(cfg=kw, batch=batch, mode=mode, smoothing=SMOOTHING, adam=ADAM,
                   logits=logits.detach().clone(), loss=losses[0], losses=losses,
                   torch_version=torch.__version__)
        if mode == "full":
            out.update(grads=grads, params3=params3, attn=attn)
        else:
            out.update(grads=summarize(grads), params3=summarize(params3),
                       attn=dict(norm=attn.double().norm().item(), head=attn[0, 0, 0, :2].clone(),
                                 shape=tuple(attn.shape)))
        path = os.path.join(HERE, f"{name}.pt")
        torch.save(out, path)
        print(f"{name}: loss={losses} logits[0,:3]={logits[0, :3].tolist()} -> {path} "
              f"({os.path.getsize(path) / 1e6:.2f} MB)")


def dropout_block():
    """One reference TransformerEncoder in TRAINING mode with dropout 0.2 (layers.py:35, 38, 102): the masks torch drew at the three
    nn.Dropout sites (captured by forward hooks), the output, and every gradient for a fixed cotangent."""
    _, ref_layers, _ = import_reference()
    torch.set_num_threads(1)
    p_drop, F_, M, head, B, T = 0.2, 128, 256, 4, 2, 17
    torch.manual_seed(5)
    blk = ref_layers.TransformerEncoder(F_, M, head=head, dropout=p_drop)
    sd = blk.state_dict()
    hash_init_(sd, seed=3)
    blk.train()
    x = torch.randn(B, T, F_, requires_grad=True)
    w = torch.randn(B, T, F_)
    masks = {}
    sites = {blk.attention.dropout: 0, blk.mlp[2]: 1, blk.mlp[5]: 2}
    hooks = [m.register_forward_hook(lambda mod, inp, out, s=s: masks.__setitem__(s, ((out != 0) | (inp[0] == 0)).clone()))
             for m, s in sites.items()]
    y = blk(x)
    for h in hooks:
        h.remove()
    (y * w).sum().backward()
    out = dict(p=p_drop, features=F_, mlp_hidden=M, head=head, state_dict={k: v.detach().clone() for k, v in blk.state_dict().items()},
               x=x.detach().clone(), w=w, masks=masks, y=y.detach().clone(), dx=x.grad.clone(),
               grads={k: v.grad.detach().clone() for k, v in blk.named_parameters()}, torch_version=torch.__version__)
    path = os.path.join(HERE, "dropout_block.pt")
    torch.save(out, path)
    print(f"dropout_block: keep rates {[round(m.float().mean().item(), 3) for m in masks.values()]} -> {path} ({os.path.getsize(path) / 1e6:.2f} MB)")


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "dropout":  # only the newer fixture; the others stay byte-identical
        dropout_block()
    else:
        main()
        dropout_block()
