"""SURVEY §8f N1 — the training-step glue on the device against torch's own autograd / optimizer in fp32
(the reference's loop is plain torch: methods/semi_supervised_learning/textual_prompt.py:93-135).
Tolerances: loss and gradients are fp32 sums over ≤ 2048 rows in a different order than torch's — relative
L2 error ≤ 2e-5; the SGD update is element-wise fp32 — ≤ 1e-6 absolute; a whole fused step inherits the fp16
prompt-only backward's tolerance (relative 1.5e-2)."""
import importlib

import pytest
import torch

from oracle import clip_ref, synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def mods():
    return {n: importlib.import_module(f"menghini-neurips23-code_b200.{n}")
            for n in ("clip", "models", "engine", "training")}


@pytest.fixture(scope="module")
def model(mods):
    m, _ = mods["clip"].load("ViT-B/32", "cuda:0", state_dict=clip_ref.synth_state_dict(seed=1234))
    return m


def _torch_ce(imfn16, text, labels, coef, scale):
    text = text.clone().requires_grad_(True)
    tn = text / text.norm(dim=-1, keepdim=True)
    logits = scale * imfn16.float() @ tn.t()
    per = torch.nn.functional.cross_entropy(logits, labels.long(), reduction="none")
    loss = per.mean() if coef is None else (per * coef).sum()
    loss.backward()
    return loss.detach(), text.grad, logits.argmax(dim=1)


@pytest.mark.parametrize("B,C,weighted", [(16, 10, False), (1000, 45, False), (2048, 102, True), (7, 3, True)])
def test_ce_text_grad_matches_autograd(model, B, C, weighted):
    g = torch.Generator().manual_seed(B * 131 + C)
    f, _ = synth.pool(B, C, peaked=0.2)
    imfn16 = f.half().cuda()
    text = (torch.randn(C, 512, generator=g) * 0.7).cuda()
    labels = torch.randint(0, C, (B,), generator=g).cuda()
    coef = None
    if weighted:   # two groups with their own means and a balance factor (textual_fpl.py:123-165)
        grp = torch.rand(B, generator=g) < 0.3
        n1, n0 = max(int(grp.sum()), 1), max(int((~grp).sum()), 1)
        coef = torch.where(grp, torch.tensor(2.5 / n1), torch.tensor(1.0 / n0)).cuda()
    loss, dtext, pred = model.engine.ce_text_grad(imfn16, text, labels, coef, scale=100.0, want_pred=True)
    w_loss, w_grad, w_pred = _torch_ce(imfn16, text, labels, coef, 100.0)
    assert abs(loss.item() - w_loss.item()) <= 2e-5 * max(1.0, abs(w_loss.item()))
    rel = ((dtext - w_grad).norm() / w_grad.norm()).item()
    assert rel <= 2e-5, rel
    assert (pred.long() == w_pred).float().mean().item() >= 0.999   # fp32 near-ties may flip
    # deterministic: a second call gives the same bits
    loss2, dtext2, _ = model.engine.ce_text_grad(imfn16, text, labels, coef, scale=100.0)
    assert torch.equal(dtext, dtext2) and torch.equal(loss, loss2)


def test_sgd_step_matches_torch_optim(model):
    g = torch.Generator().manual_seed(3)
    p0 = torch.randn(16 * 512, generator=g).cuda()
    grads = [torch.randn(16 * 512, generator=g).cuda() for _ in range(4)]
    for mu, wd in ((0.9, 0.1), (0.0, 0.0), (0.9, 0.0)):
        ref = torch.nn.Parameter(p0.clone())
        opt = torch.optim.SGD([ref], lr=0.1, momentum=mu, weight_decay=wd)
        mine, buf = p0.clone(), torch.zeros_like(p0)
        for i, gr in enumerate(grads):
            ref.grad = gr.clone()
            opt.step()
            model.engine.sgd_step(mine, gr, buf, 0.1, mu, wd, first_step=i == 0)
        assert (mine - ref.data).abs().max().item() <= 1e-6


def test_fused_coop_step_tracks_the_autograd_loop(model, mods):
    """Three optimisation steps through CoOpStep vs the reference-shaped loop (TextPrefixModel → normalise →
    logits → CrossEntropyLoss → backward → torch.optim.SGD) on the same cached features."""
    classes = [" ".join(c.split("_")) for c in synth.class_names(6, seed=1)]
    cte = mods["models"].CustomTextEncoder(model, "cuda:0", torch.float16)
    with torch.no_grad():
        _, imfn16, _ = model.engine.vit_forward(synth.images(24, seed=5).cuda(), None, want_feat=False, want_featn=True)
    labels = (torch.arange(24) % 6).cuda()
    a = mods["models"].TextPrefixModel(synth.text_prefix(16).cuda(), cte, classes, device="cuda:0")
    b = mods["models"].TextPrefixModel(synth.text_prefix(16).cuda(), cte, classes, device="cuda:0")
    opt = torch.optim.SGD([a.prefix], lr=0.002, momentum=0.9, weight_decay=0.1)   # a stable trajectory
    fused = mods["training"].CoOpStep(b, lr=0.002, weight_decay=0.1, momentum=0.9, warmup_epochs=0, epochs=10)
    scale = model.logit_scale.exp().float()   # what the reference's loop multiplies by (:105)
    assert abs(scale.item() - model.engine.logit_scale_exp) <= 1e-4 * scale.item()
    for _ in range(3):
        tf = a(classes)
        tfn = tf / tf.norm(dim=-1, keepdim=True)
        loss_a = torch.nn.functional.cross_entropy(scale * imfn16.float() @ tfn.t(), labels)
        opt.zero_grad()
        loss_a.backward()
        opt.step()
        loss_b, _ = fused.step(imfn16, labels)
        print(f"step loss autograd {loss_a.item():.6f} fused {loss_b.item():.6f}")
        assert abs(loss_a.item() - loss_b.item()) <= 5e-3 * max(1.0, abs(loss_a.item()))
    # the two paths hand the fp16 prompt-only backward gradients that differ by ~1e-5; its rounding noise
    # (relative 7e-3 against fp32 autograd, see test_gpu_towers.py) is input dependent, so the updates agree
    # to that noise level, not to 1e-5
    rel = ((a.prefix.data - b.prefix.data).norm() / (a.prefix.data - synth.text_prefix(16).cuda()).norm()).item()
    assert rel <= 1.5e-2, rel


@pytest.mark.parametrize("B,C,weighted", [(16, 10, False), (983, 102, True), (5, 3, False)])
def test_ce_image_grad_matches_autograd(model, B, C, weighted):
    """gb_ce_image_grad: the same cosine-logit CE with BOTH sides un-normalised fp32 and gradients for both
    (visual_prompt.py:122-135, multimodal_prompt.py:103-121) against torch autograd in fp32."""
    g = torch.Generator().manual_seed(B * 17 + C)
    image = (torch.randn(B, 512, generator=g) * 3.0).cuda()
    text = (torch.randn(C, 512, generator=g) * 0.7).cuda()
    labels = torch.randint(0, C, (B,), generator=g).cuda()
    coef = None
    if weighted:
        grp = torch.rand(B, generator=g) < 0.3
        n1, n0 = max(int(grp.sum()), 1), max(int((~grp).sum()), 1)
        coef = torch.where(grp, torch.tensor(0.5 / n1), torch.tensor(1.0 / n0)).cuda()
    loss, dimage, dtext, pred = model.engine.ce_image_grad(image, text, labels, coef, scale=100.0, want_dtext=True,
                                                           want_pred=True)
    im, tx = image.clone().requires_grad_(True), text.clone().requires_grad_(True)
    logits = 100.0 * (im / im.norm(dim=-1, keepdim=True)) @ (tx / tx.norm(dim=-1, keepdim=True)).t()
    per = torch.nn.functional.cross_entropy(logits, labels.long(), reduction="none")
    w_loss = per.mean() if coef is None else (per * coef).sum()
    w_loss.backward()
    assert abs(loss.item() - w_loss.item()) <= 2e-5 * max(1.0, abs(w_loss.item()))
    assert ((dimage - im.grad).norm() / im.grad.norm()).item() <= 2e-5
    assert ((dtext - tx.grad).norm() / tx.grad.norm()).item() <= 2e-5
    assert (pred.long() == logits.argmax(dim=1)).float().mean().item() >= 0.999
    loss2, dimage2, none_t, _ = model.engine.ce_image_grad(image, text, labels, coef, scale=100.0)
    assert none_t is None and torch.equal(dimage, dimage2) and torch.equal(loss, loss2)


def test_fused_vpt_step_tracks_the_autograd_loop(model, mods):
    """VPTStep vs the reference-shaped loop (ImagePrefixModel → normalise → logits → CE → backward → SGD)."""
    C, B, P = 7, 12, 16
    g = torch.Generator().manual_seed(11)
    text = torch.randn(C, 512, generator=g).cuda()
    tn = text / text.norm(dim=-1, keepdim=True)
    img = synth.images(B, seed=6).cuda()
    labels = (torch.arange(B) % C).cuda()
    cie = mods["models"].CustomImageEncoder(model.visual)
    p0 = (768 ** -0.5) * torch.randn(P, 768, generator=g)
    a = mods["models"].ImagePrefixModel(p0.clone().cuda(), cie, device="cuda:0")
    b = mods["models"].ImagePrefixModel(p0.clone().cuda(), cie, device="cuda:0")
    opt = torch.optim.SGD([a.prefix], lr=0.01, momentum=0.9, weight_decay=0.1)
    fused = mods["training"].VPTStep(b, text, lr=0.01, weight_decay=0.1, momentum=0.9, warmup_epochs=0, epochs=10)
    scale = model.engine.logit_scale_exp
    for _ in range(3):
        vf = a(img)
        vfn = vf / vf.norm(dim=-1, keepdim=True)
        loss_a = torch.nn.functional.cross_entropy(scale * vfn @ tn.t(), labels)
        opt.zero_grad()
        loss_a.backward()
        opt.step()
        loss_b, _, _ = fused.step(img, labels)
        assert abs(loss_a.item() - loss_b.item()) <= 5e-3 * max(1.0, abs(loss_a.item()))
    rel = ((a.prefix.data - b.prefix.data).norm() / (a.prefix.data - p0.cuda()).norm()).item()
    assert rel <= 1.5e-2, rel


def test_fused_upt_step_tracks_the_autograd_loop(model, mods):
    """UPTStep vs the reference-shaped loop (UPTModel.forward → normalise both → logits → CE → backward → SGD over
    the coupling head and both prompt tensors, multimodal_prompt.py:103-127)."""
    classes = [" ".join(c.split("_")) for c in synth.class_names(5, seed=2)]
    B = 10
    img = synth.images(B, seed=7).cuda()
    labels = (torch.arange(B) % 5).cuda()
    cte = mods["models"].CustomTextEncoder(model, "cuda:0", torch.float32)
    cie = mods["models"].CustomImageEncoder(model.visual)

    def make():
        torch.manual_seed(4)
        g = torch.Generator().manual_seed(4)
        coop = (0.02 * torch.randn(1, 4, 512, generator=g)).cuda()
        vpt = ((768 ** -0.5) * torch.randn(1, 4, 768, generator=g)).cuda()
        return mods["models"].UPTModel(coop, vpt, None, cie, cte, classes, 128, device="cuda:0", dtype=torch.float32)

    a, b = make(), make()
    b.load_state_dict(a.state_dict())
    oa = torch.optim.SGD(a.parameters(), lr=0.01, momentum=0.9, weight_decay=0.1)
    ob = torch.optim.SGD(b.parameters(), lr=0.01, momentum=0.9, weight_decay=0.1)
    fused = mods["training"].UPTStep(b, ob)
    scale = model.engine.logit_scale_exp
    for _ in range(3):
        tf, vf = a(img, classes)
        tfn, vfn = tf / tf.norm(dim=-1, keepdim=True), vf / vf.norm(dim=-1, keepdim=True)
        loss_a = torch.nn.functional.cross_entropy(scale * vfn @ tfn.t(), labels)
        oa.zero_grad()
        loss_a.backward()
        oa.step()
        loss_b, _, _ = fused.step(img, labels)
        assert abs(loss_a.item() - loss_b.item()) <= 5e-3 * max(1.0, abs(loss_a.item()))
    for (n, pa), (_, pb) in zip(a.named_parameters(), b.named_parameters()):
        assert (pa.data - pb.data).abs().max().item() <= 2e-2 * max(1e-3, pa.data.abs().max().item()), n
