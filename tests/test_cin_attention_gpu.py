"""GPU parity of the CIN and field self-attention kernels against the golden fixtures produced by
the unmodified reference modules and against the numpy oracle (fp32: forward 1e-5, gradients 1e-4
per-tensor max-norm relative); reference tests/test_layers.py:143-210 facts."""

import numpy as np
import pytest
import torch

from deepfm_b200.layers.attention import MultiHeadSelfAttention
from deepfm_b200.layers.cin import CIN
from oracle import deepfm_oracle as O
from tests.helpers import assert_close_rel, load_golden, split_prefixed

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("tag", ["split", "nosplit", "one"])
def test_cin_golden(tag):
    g = load_golden(f"cin_{tag}.npz")
    sizes, split = [int(v) for v in g["sizes"]], bool(g["split"])
    cin = CIN(num_fields=4, embed_dim=6, layer_sizes=sizes, split_half=split)
    with torch.no_grad():
        for i, c in enumerate(cin.conv_layers):
            c.weight.copy_(torch.from_numpy(g[f"w{i}"]))
            c.bias.copy_(torch.from_numpy(g[f"b{i}"]))
    cin = cin.cuda()
    x = torch.from_numpy(g["x"]).cuda().requires_grad_(True)
    out = cin(x)
    assert out.shape == g["out"].shape
    assert_close_rel(out.detach().cpu(), g["out"], 1e-5, "cin out")
    out.backward(torch.from_numpy(g["g"]).cuda())
    assert_close_rel(x.grad.cpu(), g["gx"], 1e-4, "cin gx")
    for i, c in enumerate(cin.conv_layers):
        assert_close_rel(c.weight.grad.cpu(), g[f"gw{i}"], 1e-4, f"gw{i}")
        assert_close_rel(c.bias.grad.cpu(), g[f"gb{i}"], 1e-4, f"gb{i}")


def test_cin_reference_shape_facts():
    # tests/test_layers.py:143-168
    x = torch.randn(4, 5, 8, device="cuda")
    assert CIN(5, 8, [64, 64], split_half=False).cuda()(x).shape == (4, 128)
    c = CIN(5, 8, [64, 64], split_half=True).cuda()
    assert c.output_dim == 32 + 64 and c(x).shape == (4, c.output_dim)
    xg = x.clone().requires_grad_(True)
    c(xg).sum().backward()
    assert xg.grad is not None and all(p.grad is not None for p in c.parameters())
    assert CIN(5, 8).layer_sizes == [128, 128]


@pytest.mark.parametrize("B,F,D,sizes,split", [(300, 16, 16, [64], True), (130, 16, 16, [128, 128, 64], True),
                                                (65, 39, 64, [24, 20], True), (37, 7, 12, [9, 5, 3], False)])
def test_cin_vs_oracle_larger(B, F, D, sizes, split):
    # Inputs are re-drawn (fixed seeds) until no pre-activation sits within fp32 rounding of the ReLU kink:
    # there the fp32 kernel and the fp64 oracle may disagree on the mask, which is not a kernel error.
    for attempt in range(40):
        torch.manual_seed(1000 * attempt + B)
        rng = np.random.default_rng(1000 * attempt + B)
        cin = CIN(F, D, sizes, split)
        x = (rng.standard_normal((B, F, D)) * 0.5).astype(np.float32)
        W = [c.weight.detach().numpy()[:, :, 0].astype(np.float64) for c in cin.conv_layers]
        b = [c.bias.detach().numpy().astype(np.float64) for c in cin.conv_layers]
        if O.cin_relu_margin(x.astype(np.float64), W, b, split) > 2e-7:      # ~10x the fp32 rounding of a pre-activation
            break
    else:
        pytest.fail("no draw with a clear ReLU margin")
    cin = cin.cuda()
    cin.precision = "fp32"                      # this test pins the CUDA-core path ("auto" may pick tf32 at B*D >= 4096)
    xt = torch.from_numpy(x).cuda().requires_grad_(True)
    out = cin(xt)
    want = O.cin_forward(x.astype(np.float64), W, b, split)
    assert_close_rel(out.detach().cpu(), want, 2e-5, "cin out")
    g = rng.standard_normal(want.shape).astype(np.float32)
    out.backward(torch.from_numpy(g).cuda())
    gx, gW, gb = O.cin_backward(x.astype(np.float64), W, b, split, g.astype(np.float64))
    assert_close_rel(xt.grad.cpu(), gx, 1e-4, "gx")
    for i, c in enumerate(cin.conv_layers):
        assert_close_rel(c.weight.grad.cpu()[:, :, 0], gW[i], 1e-4, f"gW{i}")
        assert_close_rel(c.bias.grad.cpu(), gb[i], 1e-4, f"gb{i}")


@pytest.mark.parametrize("tag", ["res", "nores"])
def test_attention_golden(tag):
    g = load_golden(f"attention_{tag}.npz")
    params, ref = split_prefixed(g, "param/"), split_prefixed(g, "grad/")
    D = g["x"].shape[2]
    A = params["layers.0.W_q.weight"].shape[0]
    att = MultiHeadSelfAttention(embed_dim=D, num_heads=int(g["num_heads"]), attention_dim=A,
                                 num_layers=int(g["num_layers"]), use_residual=bool(g["use_residual"]))
    att.load_state_dict({k: torch.from_numpy(v) for k, v in params.items()})
    att = att.cuda()
    x = torch.from_numpy(g["x"]).cuda().requires_grad_(True)
    out = att(x)
    assert out.shape == x.shape                                         # tests/test_layers.py:175-179
    assert_close_rel(out.detach().cpu(), g["out"], 1e-5, "attention out")
    out.backward(torch.from_numpy(g["g"]).cuda())
    assert_close_rel(x.grad.cpu(), g["gx"], 1e-4, "gx")
    for k, p in att.named_parameters():
        # W_k.bias grad is analytically zero (softmax shift invariance): absolute tolerance only
        if k.endswith("W_k.bias"):
            assert np.abs(p.grad.cpu().numpy() - ref[k]).max() < 1e-5, k
        else:
            assert_close_rel(p.grad.cpu(), ref[k], 1e-4, k)


# (1000,16,16,64,4), (257,11,16,32,2), (70,16,16,128,8), (5,1,16,16,1): the warp-per-sample kernels (D = 16, head dim 16,
# F <= 16; F = 11 and F = 1 exercise the masked rows); the others: the general shared-memory kernels
@pytest.mark.parametrize("B,F,D,A,H,res", [(1000, 16, 16, 64, 4, True), (77, 39, 64, 64, 4, True), (33, 5, 6, 12, 3, False),
                                            (257, 11, 16, 32, 2, False), (70, 16, 16, 128, 8, True), (5, 1, 16, 16, 1, True)])
def test_attention_vs_oracle_larger(B, F, D, A, H, res):
    rng = np.random.default_rng(B + F)
    att = MultiHeadSelfAttention(D, H, A, 1, res).cuda()
    with torch.no_grad():
        for p in att.parameters():
            p.add_(0.1 * torch.randn_like(p))
    p64 = {k[len("layers.0."):]: v.detach().cpu().numpy().astype(np.float64) for k, v in att.state_dict().items()}
    x = rng.standard_normal((B, F, D)).astype(np.float32)
    xt = torch.from_numpy(x).cuda().requires_grad_(True)
    out = att(xt)
    want = O.attn_block_forward(x.astype(np.float64), p64, H, res)
    assert_close_rel(out.detach().cpu(), want, 2e-5, "out")
    g = rng.standard_normal(x.shape).astype(np.float32)
    out.backward(torch.from_numpy(g).cuda())
    gx, pg = O.attn_block_backward(x.astype(np.float64), p64, H, res, g.astype(np.float64))
    assert_close_rel(xt.grad.cpu(), gx, 1e-4, "gx")
    for k, v in pg.items():
        got = dict(att.named_parameters())[f"layers.0.{k}"].grad.cpu().numpy()
        if k == "W_k.bias":
            assert np.abs(got - v).max() < 1e-4 * max(1.0, np.abs(pg["W_q.bias"]).max())
        else:
            assert_close_rel(got, v, 1e-4, k)


def test_attention_reference_shape_facts():
    # tests/test_layers.py:181-210: 3 layers; no residual with 2 heads / attention_dim 16
    x = torch.randn(4, 5, 16, device="cuda")
    assert MultiHeadSelfAttention(16, num_heads=2, attention_dim=16, num_layers=3).cuda()(x).shape == (4, 5, 16)
    att = MultiHeadSelfAttention(16, num_heads=2, attention_dim=16, use_residual=False).cuda()
    assert att(x).shape == (4, 5, 16) and not hasattr(att.layers[0], "layer_norm")
    xg = x.clone().requires_grad_(True)
    att(xg).sum().backward()
    assert xg.grad is not None


# ----------------------------------------------------------------- tcgen05 (TF32) CIN forward
# TF32 inputs (10-bit mantissa), FP32 accumulation in TMEM: tolerance 2e-3 per-tensor max-norm relative
# (~1.3e-4 .. 1e-3 expected); the backward runs on the tensor cores too (pinned entry-wise further down).

@pytest.mark.parametrize("B,F,D,sizes,split", [(4, 16, 16, [64], True), (300, 16, 16, [64], True),
                                                (130, 16, 16, [128, 128, 64], True), (65, 39, 64, [24, 20], True),
                                                (37, 7, 12, [9, 5, 3], False), (513, 39, 64, [128, 128], True)])
def test_cin_tcgen05_tf32_vs_oracle(B, F, D, sizes, split):
    rng = np.random.default_rng(B + F)
    torch.manual_seed(B + F)                      # fixed weights: ReLU-boundary flips are input-dependent
    cin = CIN(F, D, sizes, split).cuda()
    cin.precision = "tf32"
    x = (rng.standard_normal((B, F, D)) * 0.5).astype(np.float32)
    W = [c.weight.detach().cpu().numpy()[:, :, 0].astype(np.float64) for c in cin.conv_layers]
    b = [c.bias.detach().cpu().numpy().astype(np.float64) for c in cin.conv_layers]
    xt = torch.from_numpy(x).cuda().requires_grad_(True)
    out = cin(xt)
    want = O.cin_forward(x.astype(np.float64), W, b, split)
    assert_close_rel(out.detach().cpu(), want, 2e-3, "cin tf32 out")
    g = rng.standard_normal(want.shape).astype(np.float32)
    out.backward(torch.from_numpy(g).cuda())
    # The backward runs in fp32 on the TF32 activations.  A pre-activation within ~1e-3 of zero can land on
    # the other side of the ReLU than in the fp64 oracle, which changes single gradient entries by O(1):
    # gradients are therefore compared in relative L2 norm (5e-2), not entry-wise; the tensor-core backward itself
    # is pinned at 3e-3 by the next test, which shares the ReLU masks.
    gx, gW, gb = O.cin_backward(x.astype(np.float64), W, b, split, g.astype(np.float64))
    rel_l2 = lambda a_, b_: float(np.linalg.norm(np.asarray(a_, np.float64) - b_) / (np.linalg.norm(b_) + 1e-12))
    assert rel_l2(xt.grad.cpu().numpy(), gx) < 5e-2
    for i, c in enumerate(cin.conv_layers):
        assert rel_l2(c.weight.grad.cpu().numpy()[:, :, 0], gW[i]) < 5e-2, i


@pytest.mark.parametrize("B,F,D,sizes", [(130, 16, 16, [128, 128, 64]), (70, 39, 64, [128, 128]), (33, 20, 8, [24, 16])])
def test_cin_tcgen05_backward_matches_fp32_backward_on_same_activations(B, F, D, sizes):
    """Isolates the tensor-core backward: ONE tf32 forward (so the ReLU masks are shared), then the same
    graph is back-propagated twice -- tcgen05 (TF32) and CUDA-core fp32.  Only TF32 rounding differs:
    per-tensor max-norm relative error <= 3e-3."""
    torch.manual_seed(B)
    cin = CIN(F, D, sizes, True).cuda()
    cin.precision = "tf32"
    x = (torch.randn(B, F, D, device="cuda") * 0.5).requires_grad_(True)
    out = cin(x)
    g = torch.randn_like(out)
    grads = {}
    for prec in ("tf32", "fp32"):
        cin.precision = prec                      # read at backward time by the autograd Function
        cin.zero_grad(set_to_none=True)
        x.grad = None
        out.backward(g, retain_graph=True)
        grads[prec] = (x.grad.clone(), [c.weight.grad.clone() for c in cin.conv_layers],
                       [c.bias.grad.clone() for c in cin.conv_layers])
    assert_close_rel(grads["tf32"][0].cpu(), grads["fp32"][0].cpu(), 3e-3, "gx")
    for i in range(len(sizes)):
        assert_close_rel(grads["tf32"][1][i].cpu(), grads["fp32"][1][i].cpu(), 3e-3, f"gW{i}")
        assert_close_rel(grads["tf32"][2][i].cpu(), grads["fp32"][2][i].cpu(), 3e-3, f"gb{i}")


@pytest.mark.parametrize("B,F,D,sizes", [(130, 16, 16, [128, 128, 64]), (70, 39, 64, [128, 128])])
def test_cin_tcgen05_backward_vs_oracle_on_the_kernels_own_relu_decisions(B, F, D, sizes):
    """the TF32 tensor-core backward against the fp64 ORACLE (not the repo's fp32 kernel), entry
    by entry.  The only legitimate difference between a TF32 forward and the oracle is which side of zero a
    pre-activation within rounding distance of it lands on, so the oracle is evaluated with the ReLU decisions the
    kernel took (its post-ReLU activations > 0); what is left is TF32 rounding: 3e-3 per-tensor max-norm relative."""
    rng = np.random.default_rng(B + F)
    torch.manual_seed(B + F)
    cin = CIN(F, D, sizes, True).cuda()
    cin.precision = "tf32"
    cin.keep_activations = True
    x = (rng.standard_normal((B, F, D)) * 0.5).astype(np.float32)
    W = [c.weight.detach().cpu().numpy()[:, :, 0].astype(np.float64) for c in cin.conv_layers]
    b = [c.bias.detach().cpu().numpy().astype(np.float64) for c in cin.conv_layers]
    xt = torch.from_numpy(x).cuda().requires_grad_(True)
    out = cin(xt)
    masks = [a.detach().cpu().numpy() > 0 for a in cin.last_activations]
    g = rng.standard_normal(tuple(out.shape)).astype(np.float32)
    out.backward(torch.from_numpy(g).cuda())
    want_out = O.cin_forward(x.astype(np.float64), W, b, True, masks=masks)
    assert_close_rel(out.detach().cpu(), want_out, 2e-3, "cin tf32 out (shared ReLU decisions)")
    gx, gW, gb = O.cin_backward(x.astype(np.float64), W, b, True, g.astype(np.float64), masks=masks)
    assert_close_rel(xt.grad.cpu(), gx, 3e-3, "gx")
    for i, c in enumerate(cin.conv_layers):
        assert_close_rel(c.weight.grad.cpu().numpy()[:, :, 0], gW[i], 3e-3, f"gW{i}")
        assert_close_rel(c.bias.grad.cpu(), gb[i], 3e-3, f"gb{i}")
