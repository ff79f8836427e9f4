"""Op-level parity on a real B200: every kernel against a plain torch fp32 reference of the same op
(floating-point kernels), through the C ABI."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx(pkg):
    return pkg.Context.get(0)


def _gemm_case(ctx, M, N, K, bias=False, resid=False, act=0, out_f32=False, inplace=False):
    g = torch.Generator(device="cuda").manual_seed(M * 7 + N * 3 + K)
    A = (torch.randn(M, K, device="cuda", generator=g) * 0.5).half()
    W = (torch.randn(N, K, device="cuda", generator=g) * K ** -0.5).half()
    b = torch.randn(N, device="cuda", generator=g) * 0.1 if bias else None
    r = torch.randn(M, N, device="cuda", generator=g).half() if resid else None
    want = A.float() @ W.float().t()
    if b is not None:
        want = want + b
    if act == 1:
        want = want * torch.sigmoid(1.702 * want)
    if r is not None:
        want = want + r.float()
    if inplace:
        out = r.clone()
        ctx.gemm(A, W, b, out, out=out, act=act)
    else:
        out = ctx.gemm(A, W, b, r, act=act, out_f32=out_f32)
    torch.cuda.synchronize()
    err = (out.float() - want).abs().max().item()
    scale = want.abs().max().item()
    tol = 1e-4 * scale + 1e-5 if out_f32 else 2e-3 * scale  # fp32 accumulate; fp16 store rounding
    assert err <= tol, (M, N, K, err, tol)


@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (50, 768, 768), (77, 512, 512), (800, 2304, 768),
                                   (3300, 3072, 768), (3300, 768, 3072), (1, 512, 768),
                                   (130, 128, 64), (7700, 1536, 512), (19999, 768, 3072)])
def test_gemm_shapes(ctx, M, N, K):
    _gemm_case(ctx, M, N, K, out_f32=True)
    _gemm_case(ctx, M, N, K)


def test_gemm_epilogues(ctx):
    _gemm_case(ctx, 1000, 768, 768, bias=True)
    _gemm_case(ctx, 1000, 3072, 768, bias=True, act=1)
    _gemm_case(ctx, 1000, 768, 3072, bias=True, resid=True)
    _gemm_case(ctx, 1000, 768, 3072, bias=True, resid=True, inplace=True)
    _gemm_case(ctx, 333, 512, 768, out_f32=True)


def test_gemm_epilogues_many_tiles_per_cta(ctx):
    """Every epilogue flavour at a size where a persistent CTA pair walks 5-19 tiles: exercises the look-ahead
    paths (next tile's column constants, residual half carried across tiles, incremental tile walk)."""
    _gemm_case(ctx, 30000, 768, 768)
    _gemm_case(ctx, 30000, 2304, 768, bias=True)
    _gemm_case(ctx, 30000, 3072, 768, bias=True, act=1)
    _gemm_case(ctx, 30000, 768, 768, bias=True, resid=True)
    _gemm_case(ctx, 30000, 768, 3072, bias=True, resid=True, inplace=True)
    _gemm_case(ctx, 29999, 768, 3072, resid=True)


def test_gemm_rows_do_not_depend_on_batch(ctx):
    """Bit-identical rows whatever M is (tile shape is a function of N only)."""
    g = torch.Generator(device="cuda").manual_seed(1)
    for N, K in ((3072, 768), (768, 3072), (2304, 768), (512, 768), (1536, 512)):
        A = (torch.randn(20000, K, device="cuda", generator=g) * 0.5).half()
        W = (torch.randn(N, K, device="cuda", generator=g) * K ** -0.5).half()
        big = ctx.gemm(A, W, out_f32=True)
        for m in (1, 50, 128, 1850):
            small = ctx.gemm(A[:m].contiguous(), W, out_f32=True)
            assert torch.equal(small, big[:m]), (N, K, m)


def test_gemm_rejects_bad_shapes(ctx, pkg):
    A = torch.zeros(16, 100, device="cuda", dtype=torch.float16)
    W = torch.zeros(128, 100, device="cuda", dtype=torch.float16)
    with pytest.raises(pkg.GripB200Error):
        ctx.gemm(A, W)


@pytest.mark.parametrize("D", [512, 768])
def test_layernorm(ctx, D):
    g = torch.Generator(device="cuda").manual_seed(D)
    x = (torch.randn(1000, D, device="cuda", generator=g) * 3 + 0.5).half()
    gamma = 1 + 0.1 * torch.randn(D, device="cuda", generator=g)
    beta = 0.1 * torch.randn(D, device="cuda", generator=g)
    want = torch.nn.functional.layer_norm(x.float(), (D,), gamma, beta, 1e-5)
    y32 = ctx.layernorm(x, gamma, beta, out_f32=True)
    assert (y32 - want).abs().max().item() < 1e-4
    y16 = ctx.layernorm(x, gamma, beta)
    assert (y16.float() - want).abs().max().item() < 5e-3
    # gathered rows (CLS / EOT gather)
    idx = torch.tensor([5, 999, 0, 17], device="cuda", dtype=torch.int32)
    yg = ctx.layernorm(x, gamma, beta, row_idx=idx, out_f32=True)
    assert (yg - want[idx.long()]).abs().max().item() < 1e-4
    ys = ctx.layernorm(x, gamma, beta, in_row_mul=50, rows=20, out_f32=True)
    assert (ys - want[::50]).abs().max().item() < 1e-4


def test_l2norm(ctx):
    x = torch.randn(777, 512, device="cuda") * 3
    y16, y32 = ctx.l2norm512(x, want16=True, want32=True)
    want = x / x.norm(dim=-1, keepdim=True)
    assert (y32 - want).abs().max().item() < 1e-6
    assert (y16.float() - want).abs().max().item() < 1e-3


def _attn_ref(qkv, B, L, D, causal):
    H = D // 64
    q, k, v = qkv.float().reshape(B, L, 3, H, 64).permute(2, 0, 3, 1, 4)
    s = q @ k.transpose(-1, -2) / 8.0
    if causal:
        s = s + torch.full((L, L), float("-inf"), device=qkv.device).triu_(1)
    p = s.softmax(-1)
    return (p @ v).permute(0, 2, 1, 3).reshape(B * L, D)


@pytest.mark.parametrize("B,L,D,causal", [(3, 50, 768, 0), (2, 54, 768, 0), (5, 66, 768, 0),
                                          (4, 77, 512, 1), (7, 24, 512, 1), (2, 9, 512, 1),
                                          (1, 16, 512, 0), (2, 96, 768, 0), (3, 33, 512, 1),
                                          (301, 50, 768, 0), (200, 66, 768, 0), (150, 77, 512, 1), (1, 128, 512, 1),
                                          (64, 64, 768, 1)])
def test_attention_forward(ctx, B, L, D, causal):
    g = torch.Generator(device="cuda").manual_seed(L)
    qkv = torch.randn(B * L, 3 * D, device="cuda", generator=g).half()
    out = ctx.attention_fwd(qkv, B, L, D, causal)
    want = _attn_ref(qkv, B, L, D, causal)
    err = (out.float() - want).abs().max().item()
    assert err < 4e-3, err  # fp16 probabilities and output


@pytest.mark.parametrize("B,L,D,causal", [(2, 50, 768, 0), (3, 66, 768, 0), (3, 77, 512, 1),
                                          (4, 24, 512, 1), (2, 9, 512, 1), (1, 96, 768, 0),
                                          (301, 50, 768, 0), (151, 66, 768, 0), (75, 77, 512, 1), (2, 96, 512, 1),
                                          (3, 80, 768, 0), (65, 64, 768, 1), (333, 11, 512, 1), (5, 33, 768, 0)])
def test_attention_backward(ctx, B, L, D, causal):
    g = torch.Generator(device="cuda").manual_seed(100 + L)
    qkv = torch.randn(B * L, 3 * D, device="cuda", generator=g).half()
    dout = torch.randn(B * L, D, device="cuda", generator=g).half()
    dqkv = ctx.attention_bwd(qkv, dout, B, L, D, causal)
    ref_in = qkv.float().requires_grad_(True)
    _attn_ref(ref_in, B, L, D, causal).backward(dout.float())
    err = (dqkv.float() - ref_in.grad).abs().max().item()
    scale = ref_in.grad.abs().max().item()
    assert err < 6e-3 * max(scale, 1.0), (err, scale)
