"""GPU: every libvitb200 kernel, called through the C ABI (ctypes), against an fp32 CPU restatement.

Tolerances: fp32 check mode 1e-4 relative (north_star), bf16 mode 2e-2 relative on outputs computed from the
SAME bf16-rounded inputs (so only accumulation order / output rounding differ; most land near 4e-3).
"""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

DTYPES = [torch.float32, torch.bfloat16]


def tol(dtype):
    return 1e-4 if dtype == torch.float32 else 2e-2


def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def rnd(shape, dtype, seed, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    t = (torch.randn(shape, generator=g) * scale).to(dtype)
    return t  # CPU tensor already rounded to the storage dtype


@pytest.fixture(scope="module")
def ops():
    import vit_cifar_b200  # noqa: F401
    from vit_cifar_b200 import ops as o
    o.require_device()
    return o


def cu(t):
    return t.cuda().contiguous()


# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("rows,H", [(260, 128), (1040, 384), (77, 768)])
def test_layernorm_fwd_bwd(ops, dtype, rows, H):
    x = rnd((rows, H), dtype, 1, 2.0)
    gam = rnd((H,), torch.float32, 2) * 0.2 + 1
    bet = rnd((H,), torch.float32, 3) * 0.2
    dy = rnd((rows, H), dtype, 4)
    dres = rnd((rows, H), dtype, 5)
    xr = x.float().requires_grad_(True)
    gr, br = gam.clone().requires_grad_(True), bet.clone().requires_grad_(True)
    yr = F.layer_norm(xr, (H,), gr, br, 1e-5)
    yr.backward(dy.float())
    dx_ref = xr.grad + dres.float()

    y = torch.empty_like(cu(x)); mean = torch.empty(rows, device="cuda"); rstd = torch.empty(rows, device="cuda")
    ops.layernorm_fwd(cu(x), H, cu(gam), cu(bet), y, mean, rstd, rows, H)
    assert rel(y, yr.detach()) < tol(dtype)
    dx = torch.empty_like(y); dg = torch.empty(H, device="cuda"); db = torch.empty(H, device="cuda"); dc = torch.empty(H, device="cuda")
    ops.layernorm_bwd(cu(dy), cu(x), H, cu(gam), mean, rstd, cu(dres), dx, H, dg, db, dc, rows, H)
    assert rel(dx, dx_ref) < tol(dtype)
    assert rel(dg, gr.grad) < tol(dtype)
    assert rel(db, br.grad) < tol(dtype)
    assert rel(dc, dx_ref.sum(0)) < (1e-4 if dtype == torch.float32 else 5e-3)  # bias gradient of the producing Linear


@pytest.mark.parametrize("dtype", DTYPES)
def test_layernorm_strided_rows(ops, dtype):
    B, T, H = 6, 17, 128
    x = rnd((B, T, H), dtype, 11)
    gam = torch.ones(H); bet = torch.zeros(H)
    y = torch.empty((B, H), dtype=dtype, device="cuda"); mean = torch.empty(B, device="cuda"); rstd = torch.empty(B, device="cuda")
    ops.layernorm_fwd(cu(x), T * H, cu(gam), cu(bet), y, mean, rstd, B, H)
    assert rel(y, F.layer_norm(x[:, 0].float(), (H,))) < tol(dtype)


# ---------------------------------------------------------------------------------------------
GEMM_SHAPES = [(260, 128, 128), (260, 384, 128), (1040, 1152, 384), (520, 384, 384), (130, 256, 128), (128, 128, 64), (100, 128, 128),
               (33, 256, 64), (20000, 384, 384), (4, 10, 128)]


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("M,N,K", GEMM_SHAPES)
@pytest.mark.parametrize("variant", ["plain", "gelu_res_pre"])
def test_gemm_fwd(ops, dtype, M, N, K, variant):
    a = rnd((M, K), dtype, 1); w = rnd((N, K), dtype, 2, 1 / math.sqrt(K)); bias = rnd((N,), torch.float32, 3)
    res = rnd((M, N), dtype, 4)
    z_ref = a.float() @ w.float().t() + bias
    out = torch.empty((M, N), dtype=dtype, device="cuda")
    if variant == "plain":
        ops.gemm_fwd(cu(a), cu(w), cu(bias), None, out, None, M, N, K)
        assert rel(out, z_ref) < tol(dtype)
    else:
        pre = torch.empty_like(out)
        ops.gemm_fwd(cu(a), cu(w), cu(bias), cu(res), out, pre, M, N, K, gelu=True)
        assert rel(pre, z_ref) < tol(dtype)
        assert rel(out, F.gelu(z_ref) + res.float()) < tol(dtype)


@pytest.mark.parametrize("dtype", DTYPES)
def test_gemm_fwd_out_f32_head(ops, dtype):
    M, N, K = 7, 100, 384
    a = rnd((M, K), dtype, 1); w = rnd((N, K), dtype, 2, 0.05); bias = rnd((N,), torch.float32, 3)
    out = torch.empty((M, N), dtype=torch.float32, device="cuda")
    ops.gemm_fwd(cu(a), cu(w), cu(bias), None, out, None, M, N, K, out_f32=True)
    assert rel(out, a.float() @ w.float().t() + bias) < 1e-4  # fp32 accumulate and store in both modes


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("M,N,K", GEMM_SHAPES[:-1])
@pytest.mark.parametrize("with_z", [False, True])
def test_gemm_dgrad(ops, dtype, M, N, K, with_z):
    dy = rnd((M, N), dtype, 1); w = rnd((N, K), dtype, 2, 1 / math.sqrt(N)); z = rnd((M, K), dtype, 3)
    ref = dy.float() @ w.float()
    if with_z:
        zz = z.float().requires_grad_(True)
        F.gelu(zz).sum().backward()
        ref = ref * zz.grad
    dx = torch.empty((M, K), dtype=dtype, device="cuda")
    ops.gemm_dgrad(cu(dy), cu(w), cu(z) if with_z else None, dx, M, N, K)
    assert rel(dx, ref) < tol(dtype)


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("M,N,K", GEMM_SHAPES[:-1] + [(8320, 384, 384), (66560, 1152, 384)])
def test_gemm_wgrad_dbias(ops, dtype, M, N, K):
    dy = rnd((M, N), dtype, 1); x = rnd((M, K), dtype, 2)
    dw = torch.empty((N, K), dtype=torch.float32, device="cuda"); db = torch.empty((N,), dtype=torch.float32, device="cuda")
    ops.gemm_wgrad(cu(dy), cu(x), dw, db, M, N, K)
    assert rel(dw, dy.float().t() @ x.float()) < 1e-4   # fp32 accumulation of exact bf16 products
    assert rel(db, dy.float().sum(0)) < 1e-4
    dw2 = torch.empty_like(dw)
    ops.gemm_wgrad(cu(dy), cu(x), dw2, None, M, N, K)
    assert torch.equal(dw, dw2)  # deterministic split-K


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("B,C,H", [(1024, 10, 384), (130, 100, 384), (9, 128, 768), (1023, 100, 384), (17, 10, 128), (64, 7, 200)])
def test_classifier_head_kernels(ops, dtype, B, C, H):
    """fc[1] (vit.py:63,76) forward / dgrad / wgrad+dbias through the small-N kernels (head.cu), at CIFAR-10 and -100 widths."""
    hn = rnd((B, H), dtype, 1); w = rnd((C, H), dtype, 2, 0.05); bias = rnd((C,), torch.float32, 3)
    dl = rnd((B, C), torch.float32, 4, 1.0 / B)
    logits = torch.empty((B, C), device="cuda")
    ops.gemm_fwd(cu(hn), cu(w), cu(bias), None, logits, None, B, C, H, out_f32=True)
    assert rel(logits, hn.float() @ w.float().t() + bias) < 1e-4
    dw = torch.empty((C, H), device="cuda"); db = torch.empty((C,), device="cuda")
    ops.gemm_wgrad(cu(dl), cu(hn), dw, db, B, C, H, dy_f32=True)
    assert rel(dw, dl.t() @ hn.float()) < 1e-4 and rel(db, dl.sum(0)) < 1e-4
    dw2 = torch.empty_like(dw); db2 = torch.empty_like(db)
    ops.gemm_wgrad(cu(dl), cu(hn), dw2, db2, B, C, H, dy_f32=True)
    assert torch.equal(dw, dw2) and torch.equal(db, db2)  # fixed-order reductions
    dx = torch.empty((B, H), dtype=dtype, device="cuda")
    ops.gemm_dgrad(cu(dl), cu(w), None, dx, B, C, H, dy_f32=True)
    assert rel(dx, dl @ w.float()) < tol(dtype)


@pytest.mark.parametrize("dtype", DTYPES)
def test_gemm_head_backward_f32_dy(ops, dtype):
    B, C, H = 5, 10, 128
    dl = rnd((B, C), torch.float32, 1); hn = rnd((B, H), dtype, 2); w = rnd((C, H), dtype, 3)
    dw = torch.empty((C, H), device="cuda"); db = torch.empty((C,), device="cuda")
    ops.gemm_wgrad(cu(dl), cu(hn), dw, db, B, C, H, dy_f32=True)
    assert rel(dw, dl.t() @ hn.float()) < 1e-4 and rel(db, dl.sum(0)) < 1e-5
    dx = torch.empty((B, H), dtype=dtype, device="cuda")
    ops.gemm_dgrad(cu(dl), cu(w), None, dx, B, C, H, dy_f32=True)
    assert rel(dx, dl @ w.float()) < tol(dtype)


# ---------------------------------------------------------------------------------------------
def attn_ref(qkv, B, T, heads, d, scale, do=None):
    H = heads * d
    qkv = qkv.float().view(B, T, 3, heads, d).requires_grad_(do is not None)
    q, k, v = qkv[:, :, 0].transpose(1, 2), qkv[:, :, 1].transpose(1, 2), qkv[:, :, 2].transpose(1, 2)
    s = torch.einsum("bhif,bhjf->bhij", q, k) * scale
    p = s.softmax(-1)
    o = torch.einsum("bhij,bhjf->bihf", p, v).reshape(B, T, H)
    lse = torch.logsumexp(s, -1)
    if do is None:
        return o, lse, p
    o.backward(do.float())
    return o.detach(), lse.detach(), p.detach(), qkv.grad.reshape(B, T, 3 * H)


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("B,T,heads,d", [(3, 65, 12, 32), (2, 17, 12, 32), (2, 16, 4, 32), (2, 64, 2, 64), (1, 65, 12, 64), (1, 100, 2, 32),
                                         # many (image, head) items per persistent forward CTA: 1,560 / 4,800 items on <= 592 / 1,184 CTAs
                                         (130, 65, 12, 32), (400, 17, 12, 32)])
def test_attention_fwd_bwd(ops, dtype, B, T, heads, d):
    H = heads * d
    scale = 1.0 / math.sqrt(H)
    qkv = rnd((B, T, 3 * H), dtype, 1, 2.0)
    do = rnd((B, T, H), dtype, 2)
    o_ref, lse_ref, p_ref, dqkv_ref = attn_ref(qkv, B, T, heads, d, scale, do)
    o = torch.empty((B, T, H), dtype=dtype, device="cuda"); lse = torch.empty((B, heads, T), device="cuda")
    am = torch.empty((B, heads, T, T), device="cuda")
    ops.attn_fwd(cu(qkv), o, lse, am, B, T, heads, d, scale)
    assert rel(o, o_ref) < tol(dtype)
    assert rel(lse, lse_ref) < 1e-4
    assert rel(am, p_ref) < (1e-4 if dtype == torch.float32 else 5e-3)
    o2 = torch.empty_like(o); lse2 = torch.empty_like(lse)
    ops.attn_fwd(cu(qkv), o2, lse2, None, B, T, heads, d, scale)
    assert torch.equal(o, o2)
    dqkv = torch.empty((B, T, 3 * H), dtype=dtype, device="cuda")
    ops.attn_bwd(cu(qkv), o, cu(do), lse, dqkv, B, T, heads, d, scale)
    assert rel(dqkv, dqkv_ref) < tol(dtype)


# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("rows,cols", [(260, 128), (1040, 384), (999, 1152), (33, 3072)])
def test_gelu_bwd_and_colsum(ops, dtype, rows, cols):
    dy = rnd((rows, cols), dtype, 1); z = rnd((rows, cols), dtype, 2, 1.5)
    zz = z.float().requires_grad_(True)
    F.gelu(zz).backward(dy.float())
    dz = torch.empty((rows, cols), dtype=dtype, device="cuda"); cs = torch.empty(cols, device="cuda")
    ops.gelu_bwd_colsum(cu(dy), cu(z), dz, cs, rows, cols)
    assert rel(dz, zz.grad) < tol(dtype)
    assert rel(cs, zz.grad.sum(0)) < (1e-4 if dtype == torch.float32 else 5e-3)
    cs2 = torch.empty(cols, device="cuda")
    ops.colsum(cu(dy), cs2, rows, cols)
    assert rel(cs2, dy.float().sum(0)) < 1e-4


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("P,has_cls", [(8, True), (4, True), (4, False)])
@pytest.mark.parametrize("B", [5, 70])
def test_patch_embed_fwd_bwd(ops, dtype, P, has_cls, B):
    import oracle
    S, H = 32, 128
    cfg = oracle.ViTConfig(patch=P, hidden=H, is_cls_token=has_cls)
    K, T = cfg.patch_len, cfg.num_tokens
    img = rnd((B, 3, S, S), torch.float32, 1)
    w = (rnd((H, K), torch.float32, 2) / math.sqrt(K)).requires_grad_(True)
    b = rnd((H,), torch.float32, 3).requires_grad_(True)
    cls = rnd((1, 1, H), torch.float32, 4).requires_grad_(True)
    pos = rnd((1, T, H), torch.float32, 5).requires_grad_(True)
    out_ref = F.linear(oracle.to_words(img, cfg), w, b)
    if has_cls:
        out_ref = torch.cat([cls.repeat(B, 1, 1), out_ref], 1)
    out_ref = out_ref + pos
    dout = rnd((B, T, H), dtype, 6)
    out_ref.backward(dout.float())
    out = torch.empty((B * T, H), dtype=dtype, device="cuda")
    words = torch.empty((B * P * P, K), dtype=dtype, device="cuda") if dtype == torch.bfloat16 else None
    w_act = cu(w.detach().to(dtype)) if dtype == torch.bfloat16 else None
    ops.patch_embed_fwd(cu(img), cu(w.detach()), w_act, cu(b.detach()), cu(cls.detach().view(-1)) if has_cls else None,
                        cu(pos.detach().view(T, H)), out, words, P, has_cls)
    assert rel(out.view(B, T, H), out_ref.detach()) < (1e-5 if dtype == torch.float32 else 5e-3)
    if words is not None:
        assert rel(words.view(B, P * P, K), oracle.to_words(img, cfg)) < 3e-3
    dw = torch.empty((H, K), device="cuda"); db = torch.empty(H, device="cuda"); dpos = torch.empty((T, H), device="cuda")
    dcls = torch.empty(H, device="cuda") if has_cls else None
    ops.patch_embed_bwd(cu(img), words, cu(dout).view(B * T, H), dw, db, dcls, dpos, P, has_cls)
    tol_w = 1e-4 if dtype == torch.float32 else 5e-3  # bf16 path multiplies the bf16-rounded patch matrix
    assert rel(dw, w.grad) < tol_w and rel(db, b.grad) < 1e-4 and rel(dpos, pos.grad.view(T, H)) < 1e-4
    if has_cls:
        assert rel(dcls, cls.grad.view(-1)) < 1e-4


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("mode", [0, 1])
def test_pool(ops, dtype, mode):
    B, T, H = 4, 17, 128
    x = rnd((B, T, H), dtype, 1); dy = rnd((B, H), dtype, 2)
    y = torch.empty((B, H), dtype=dtype, device="cuda")
    ops.pool_fwd(cu(x), y, B, T, H, mode)
    ref = x.float()[:, 0] if mode == 0 else x.float().mean(1)
    assert rel(y, ref) < tol(dtype)
    dx = torch.full((B, T, H), 7.0, dtype=dtype, device="cuda")
    ops.pool_bwd(cu(dy), dx, B, T, H, mode)
    ref = torch.zeros(B, T, H)
    if mode == 0:
        ref[:, 0] = dy.float()
    else:
        ref[:] = dy.float()[:, None] / T
    assert rel(dx, ref) < tol(dtype)
    if mode == 0:  # mode 2: cls rows only, into a buffer the caller keeps zero elsewhere
        dx2 = torch.zeros((B, T, H), dtype=dtype, device="cuda")
        ops.pool_bwd(cu(dy), dx2, B, T, H, 2)
        assert torch.equal(dx2, dx)
        dx2[:, 1:] = 3.0
        ops.pool_bwd(cu(dy), dx2, B, T, H, 2)
        assert torch.equal(dx2[:, 0], dx[:, 0]) and bool((dx2[:, 1:] == 3.0).all())  # really writes nothing else


@pytest.mark.parametrize("B,C", [(4, 10), (128, 100), (1000, 10), (3, 1000), (1024, 100), (1023, 17), (5, 16), (33, 256), (2, 257)])
def test_ls_ce(ops, B, C):
    import oracle
    z = rnd((B, C), torch.float32, 1, 3.0); y = torch.randint(0, C, (B,), generator=torch.Generator().manual_seed(2))
    loss = torch.empty((), device="cuda"); dl = torch.empty((B, C), device="cuda")
    for fn in (ops.ls_ce, ops.ls_ce_single_block):  # many blocks + workspace, and the workspace-free single block
        loss.fill_(-1.0); dl.fill_(7.0)
        fn(cu(z), cu(y), loss, dl, 0.1, 1.0)
        assert abs(loss.item() - oracle.ls_ce_loss(z, y, C, 0.1).item()) < 1e-5 * max(1.0, abs(loss.item()))
        assert rel(dl, oracle.ls_ce_dlogits(z, y, C, 0.1)) < 1e-5
    # deterministic (block partials are added in block order) and reusable: the arrival counter is back at zero
    l2 = torch.empty((), device="cuda"); d2 = torch.empty_like(dl)
    for _ in range(3):
        ops.ls_ce(cu(z), cu(y), l2, d2, 0.1, 1.0)
        ops.ls_ce(cu(z), cu(y), loss, dl, 0.1, 1.0)
        assert l2.item() == loss.item() and torch.equal(d2, dl)


def test_adam_matches_oracle(ops):
    import oracle
    from vit_cifar_b200 import adam_hyper
    n = 10007
    p = rnd((n,), torch.float32, 1)
    ref = {"p": p.clone()}
    m = {"p": torch.zeros(n)}; v = {"p": torch.zeros(n)}
    pd, md, vd = cu(p), torch.zeros(n, device="cuda"), torch.zeros(n, device="cuda")
    sh = torch.zeros(n, dtype=torch.bfloat16, device="cuda")
    hd = torch.zeros(16, device="cuda")
    for t in range(1, 5):
        g = rnd((n,), torch.float32, 10 + t)
        oracle.adam_step(ref, {"p": g}, m, v, t)
        h = adam_hyper(t, 1e-3, 0.9, 0.999, 1e-8, 5e-5)
        if t % 2:
            ops.adam(pd, cu(g), md, vd, sh, hyper_host=h)
        else:  # device-resident hyper-parameters (the CUDA-graph path)
            hd.copy_(torch.tensor(h + [0.0] * (16 - len(h))))
            ops.adam(pd, cu(g), md, vd, sh, hyper_dev=hd)
    assert rel(pd, ref["p"]) < 1e-6
    assert rel(md, m["p"]) < 1e-6 and rel(vd, v["p"]) < 1e-6
    assert torch.equal(sh.cpu(), pd.cpu().to(torch.bfloat16))


def test_bad_arguments_raise(ops):
    from vit_cifar_b200 import VitbError
    x = torch.zeros((4, 100), device="cuda")
    with pytest.raises(VitbError):
        ops.layernorm_fwd(x, 100, x, x, x, x, x, 4, 100)  # H not a multiple of 128
    with pytest.raises(VitbError):
        ops.attn_fwd(x, x, x, None, 1, 200, 2, 32, 1.0)  # T > 128
    with pytest.raises(VitbError):
        ops.cast_f32_to_bf16(torch.zeros(4), torch.zeros(4, dtype=torch.bfloat16))  # CPU tensors: no fallback


@pytest.mark.parametrize("B,C,lam", [(64, 10, 0.3), (1000, 100, 0.85), (5, 10, 1.0), (7, 10, 0.0)])
def test_ls_ce_two_targets(ops, B, C, lam):
    """CutMix / MixUp objective (network.py:149-167) in the fused loss kernel vs the oracle and its autograd gradient;
    lam from the host argument and from device memory; lam = 1 reproduces the plain loss bit for bit."""
    import oracle
    g = torch.Generator().manual_seed(B + C)
    z = torch.randn(B, C, generator=g) * 3
    ya = torch.randint(0, C, (B,), generator=g); yb = torch.randint(0, C, (B,), generator=g)
    zr = z.clone().requires_grad_(True)
    ref = oracle.mixed_ls_ce_loss(zr, ya, yb, lam, C, 0.1)
    ref.backward()
    for use_dev in (False, True):
        loss = torch.zeros((), device="cuda"); dl = torch.empty((B, C), device="cuda")
        lam_dev = torch.tensor([lam], device="cuda") if use_dev else None
        ops.ls_ce(cu(z), cu(ya), loss, dl, 0.1, 1.0, labels_b=cu(yb), lam=(0.5 if use_dev else lam), lam_dev=lam_dev)
        assert abs(loss.item() - ref.item()) < 1e-5 * max(1.0, abs(ref.item()))
        assert rel(dl, zr.grad) < 1e-5
    if lam == 1.0:
        loss1 = torch.zeros((), device="cuda"); dl1 = torch.empty((B, C), device="cuda")
        ops.ls_ce(cu(z), cu(ya), loss1, dl1, 0.1, 1.0)
        assert loss1.item() == loss.item() and torch.equal(dl1, dl)


@pytest.mark.parametrize("B,S,pad", [(37, 32, 4), (4, 32, 0), (1024, 32, 4)])
def test_augment_crop_flip_normalize(ops, B, S, pad):
    """Device input pipeline (utils.py:337-355) vs the oracle for the same random draws; bit-exact up to one fp32 rounding."""
    import oracle
    g = torch.Generator().manual_seed(B)
    img = torch.randint(0, 256, (B, S, S, 3), generator=g, dtype=torch.uint8)
    dx = torch.randint(0, 2 * pad + 1, (B,), generator=g, dtype=torch.int32)
    dy = torch.randint(0, 2 * pad + 1, (B,), generator=g, dtype=torch.int32)
    fl = torch.randint(0, 2, (B,), generator=g, dtype=torch.uint8)
    mean, std = (0.5071, 0.4867, 0.4408), (0.2675, 0.2565, 0.2761)
    out = torch.empty((B, 3, S, S), device="cuda")
    ops.augment(img.cuda(), dx.cuda(), dy.cuda(), fl.cuda(), mean, std, out, pad)
    nb = min(B, 64)  # the oracle loops over images in Python
    ref = oracle.augment_crop_flip_normalize(img[:nb], dx[:nb], dy[:nb], fl[:nb], mean, std, pad)
    assert (out[:nb].cpu() - ref).abs().max().item() < 1e-6
    out2 = torch.empty_like(out)
    ops.augment(img.cuda(), None, None, None, mean, std, out2, pad)   # no draws = centre crop, no flip = plain ToTensor + Normalize
    plain = (img.permute(0, 3, 1, 2).float() / 255.0 - torch.tensor(mean).view(1, 3, 1, 1)) / torch.tensor(std).view(1, 3, 1, 1)
    assert (out2.cpu() - plain).abs().max().item() < 1e-6


def test_gpu_augment_class_statistics(ops):
    import vit_cifar_b200 as vb
    aug = vb.GpuAugment(size=32, padding=4, seed=1)
    img = torch.full((512, 32, 32, 3), 255, dtype=torch.uint8, device="cuda")
    out = aug(img)
    assert out.shape == (512, 3, 32, 32) and out.dtype == torch.float32
    hi = [(1.0 - m) / s for m, s in zip(vb.schedule.CIFAR10_MEAN, vb.schedule.CIFAR10_STD)]
    lo = [(0.0 - m) / s for m, s in zip(vb.schedule.CIFAR10_MEAN, vb.schedule.CIFAR10_STD)]
    # every pixel is either white or padding-black after normalisation, and on average (1 - 2/8)^2-ish of the crop is image
    for c in range(3):
        ch = out[:, c]
        assert torch.all(((ch - hi[c]).abs() < 1e-5) | ((ch - lo[c]).abs() < 1e-5))
    frac = ((out[:, 0] - hi[0]).abs() < 1e-5).float().mean().item()
    assert 0.85 < frac < 0.93   # E[(32 - |d|) / 32]^2 with d uniform in [-4, 4]: 0.879


# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dtype", DTYPES)
def test_dropout_is_the_documented_philox_stream(ops, dtype):
    """vitb_dropout: the keep mask equals the numpy restatement of Philox4x32-10 bit for bit; kept values are scaled by
    1/(1-p) (nn.Dropout, layers.py:35, 38, 102); residual add, in-place use, device-side step, backward = same call."""
    import oracle
    n, p, seed, site, step = 8 * 4099, 0.3, 0x1234567890ABCDEF, 2, 77
    x = rnd((n,), dtype, 1); x[x == 0] = 1.0
    res = rnd((n,), dtype, 2)
    keep = torch.from_numpy(oracle.dropout_keep_mask(n, p, seed, site, step))
    out = torch.empty(n, dtype=dtype, device="cuda")
    ops.dropout(cu(x), None, out, p, seed, site, step)
    assert torch.equal((out != 0).cpu(), keep)
    ref = x.float() * keep / (1 - p)
    assert rel(out, ref) < (1e-6 if dtype == torch.float32 else 4e-3)
    step_dev = torch.tensor([step], dtype=torch.int32, device="cuda")
    y = cu(x).clone()
    ops.dropout(y, cu(res), y, p, seed, site, 0, step_dev)  # in place, residual, step read on the device
    assert rel(y, ref + res.float()) < (1e-6 if dtype == torch.float32 else 4e-3)
    for other in [dict(site=1), dict(step=78), dict(seed=seed + 1)]:
        kw = dict(seed=seed, site=site, step=step); kw.update(other)
        o2 = torch.empty_like(out)
        ops.dropout(cu(x), None, o2, p, kw["seed"], kw["site"], kw["step"])
        assert torch.equal((o2 != 0).cpu(), torch.from_numpy(oracle.dropout_keep_mask(n, p, kw["seed"], kw["site"], kw["step"])))
        assert not torch.equal(o2 != 0, out != 0)
    ops.dropout(cu(x), None, out, 0.0, seed, site, step)
    assert torch.equal(out.cpu(), x)
    with pytest.raises(Exception):
        ops.dropout(cu(x), None, out, 1.0, seed, site, step)
    with pytest.raises(Exception):
        ops.dropout(cu(x)[:12], None, out[:12], p, seed, site, step)


def test_dropout_keep_rate_at_training_size(ops):
    """Activation-sized tensor (66 560 x 384 bf16): keep rate within 5 sigma of 1 - round(p * 65536) / 65536, mean preserved."""
    n, p = 66560 * 384, 0.1
    x = torch.ones(n, dtype=torch.bfloat16, device="cuda")
    out = torch.empty_like(x)
    ops.dropout(x, None, out, p, 42, 0, 1)
    kept = (out != 0).float().mean().item()
    q = 1.0 - round(p * 65536) / 65536
    assert abs(kept - q) < 5 * math.sqrt(q * (1 - q) / n)
    assert abs(out.float().mean().item() - q / (1 - p)) < 1e-2


def test_cutmix_mixup_on_device(ops):
    """vb.GpuCutMix / vb.GpuMixUp (da.py:51-93): with the generators seeded like the reference's they return the oracle's batches
    bit for bit (the paste is a copy; the blend uses torch's rounding: two products, one sum), the shuffled labels and lambda."""
    import numpy as np
    import oracle
    import vit_cifar_b200 as vb
    g = torch.Generator().manual_seed(3)
    img = torch.randn(64, 3, 32, 32, generator=g); label = torch.randint(0, 10, (64,), generator=g)
    for seed in range(5):
        np.random.seed(seed); torch.manual_seed(seed)
        out, la, lb, lam = vb.GpuCutMix(32, 1.0)((img.cuda(), label.cuda()))
        np.random.seed(seed); torch.manual_seed(seed)
        perm = torch.randperm(64)
        box, lam_ref = vb.cutmix_box(32, np.random.beta(1.0, 1.0), np.random.uniform(0, 32), np.random.uniform(0, 32))
        assert lam == lam_ref and torch.equal(lb.cpu(), label[perm]) and torch.equal(la.cpu(), label)
        assert torch.equal(out.cpu(), oracle.cutmix_apply(img, perm, box))
        np.random.seed(seed); torch.manual_seed(seed)
        mx, ya, yb, lam_m = vb.GpuMixUp(0.4)((img.cuda(), label.cuda()))
        np.random.seed(seed); torch.manual_seed(seed)
        lam2 = np.random.beta(0.4, 0.4); index = torch.randperm(64)
        assert lam_m == lam2 and torch.equal(yb.cpu(), label[index])
        assert torch.equal(mx.cpu(), oracle.mixup_apply(img, index, lam2))
    with pytest.raises(Exception):
        ops.batch_mix(img.cuda(), torch.arange(64, dtype=torch.int32, device="cuda"), img.cuda()[:, :, :, :30].contiguous(), 0)
