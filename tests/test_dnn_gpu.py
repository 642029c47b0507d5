"""DNN tower on the repo's own kernels against torch in fp64.

dfm_gemm3 (tcgen05, 3xTF32) in its three operand layouts: max-norm relative error <= 3e-6 vs the fp64 product (the
error class of an fp32 SIMT sgemm; a single TF32 pass would be ~1e-3).  The fused Linear -> BatchNorm1d -> act ->
Dropout block and the head Linear against the eager nn.Sequential of the same module evaluated in fp64: forward 1e-5,
gradients 1e-4 (reference: deepfm/models/layers/dnn.py:45-59)."""

import numpy as np
import pytest
import torch

from deepfm_b200 import _lib
from deepfm_b200.layers.dnn import DNN, _gemm3, linear_head
from tests.helpers import assert_close_rel

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("M,N,K", [(128, 64, 32), (300, 100, 36), (4096, 256, 2496), (1000, 2496, 256), (777, 64, 128),
                                   (65, 32, 64), (33, 8, 256)])
def test_gemm3_nt_nn_tn_vs_fp64(M, N, K):
    torch.manual_seed(M + N + K)
    x = torch.randn(M, K, device="cuda")
    w = torch.randn(N, K, device="cuda") * 0.1
    b = torch.randn(N, device="cuda")
    y = _gemm3(0, x, w, torch.empty(M, N, device="cuda"), b, M, N, K)
    ref = x.double() @ w.double().t() + b.double()
    assert_close_rel(y.cpu(), ref.cpu(), 3e-6, "NT (Linear forward)")
    dy = torch.randn(M, N, device="cuda")
    dx = _gemm3(1, dy, w, torch.empty(M, K, device="cuda"), None, M, K, N)
    assert_close_rel(dx.cpu(), (dy.double() @ w.double()).cpu(), 3e-6, "NN (grad input)")
    dw = _gemm3(2, dy, x, torch.empty(N, K, device="cuda"), None, N, K, M)
    assert_close_rel(dw.cpu(), (dy.double().t() @ x.double()).cpu(), 3e-6, "TN (grad weight)")
    dw2 = _gemm3(2, dy, x, torch.empty(N, K, device="cuda"), None, N, K, M)
    assert torch.equal(dw, dw2)                                   # split-K with a fixed-order reduce: deterministic


def test_gemm3_rejects_unaligned_extent():
    x = torch.randn(64, 30, device="cuda")
    w = torch.randn(16, 30, device="cuda")
    with pytest.raises(NotImplementedError):
        _gemm3(0, x, w, torch.empty(64, 16, device="cuda"), None, 64, 16, 30)


@pytest.mark.parametrize("single", [True, False])
@pytest.mark.parametrize("act,bn,p", [("relu", True, 0.0), ("gelu", True, 0.0), ("tanh", False, 0.0), ("leaky_relu", True, 0.0)])
def test_tower_matches_eager_fp64(act, bn, p, single, monkeypatch):
    # single = True: the whole tower as one autograd node (dfm_tower_fwd / dfm_tower_bwd); False: one node per block
    monkeypatch.setattr(DNN, "single_call", single)
    torch.manual_seed(3)
    B, width = 2048, 364
    tower = DNN(width, [256, 128, 64], activation=act, dropout=p, use_batch_norm=bn).cuda().train()
    head = torch.nn.Linear(64, 1).cuda()
    ref = DNN(width, [256, 128, 64], activation=act, dropout=p, use_batch_norm=bn).double().cuda().train()
    ref.load_state_dict({k: v.double() for k, v in tower.state_dict().items()})
    ref_head = torch.nn.Linear(64, 1).double().cuda()
    ref_head.load_state_dict({k: v.double() for k, v in head.state_dict().items()})
    x = torch.randn(B, width, device="cuda", requires_grad=True)
    xd = x.detach().double().requires_grad_(True)
    g = torch.randn(B, 1, device="cuda")
    out = linear_head(head, tower(x))
    out.backward(g)
    DNN.fused = False
    try:
        want = ref_head(ref(xd))
        want.backward(g.double())
    finally:
        DNN.fused = True
    assert_close_rel(out.detach().cpu(), want.detach().cpu(), 1e-5, "tower output")
    assert_close_rel(x.grad.cpu(), xd.grad.cpu(), 1e-4, "grad input")
    for (k, pm), (_, pr) in zip(list(tower.named_parameters()) + list(head.named_parameters()),
                                list(ref.named_parameters()) + list(ref_head.named_parameters())):
        # a Linear bias in front of a training-mode BatchNorm has an analytically zero gradient (the fp64 reference says
        # 1e-14): what is left is the fp32 rounding noise of a 2048-term sum of +-1-sized terms
        zero_bias = bn and k.endswith(".bias") and k.startswith("mlp.") and int(k.split(".")[1]) % 4 == 0
        assert_close_rel(pm.grad.cpu(), pr.grad.cpu(), 1e-4, k, floor=3e-5 if zero_bias else 1e-6)
    for (k, bm), (_, br) in zip(tower.named_buffers(), ref.named_buffers()):     # running statistics, batch counter
        assert_close_rel(bm.detach().double().cpu(), br.detach().cpu(), 1e-5, k)
    # eval mode: running statistics, no dropout
    tower.eval(); ref.eval()
    with torch.no_grad():
        e1 = tower(x.detach())
        DNN.fused = False
        try:
            e2 = ref(xd.detach())
        finally:
            DNN.fused = True
    assert_close_rel(e1.cpu(), e2.cpu(), 1e-5, "eval output")


@pytest.mark.parametrize("single", [True, False])
def test_tower_dropout_mask_is_regenerated_in_backward(single, monkeypatch):
    monkeypatch.setattr(DNN, "single_call", single)
    torch.manual_seed(5)
    B = 4096
    tower = DNN(64, [128], activation="relu", dropout=0.5, use_batch_norm=False).cuda().train()
    x = torch.randn(B, 64, device="cuda", requires_grad=True)
    out = tower(x)
    zero = (out == 0).float().mean().item()
    assert 0.65 < zero < 0.85, zero                              # relu zeros (~50 %) + dropped half of the rest
    pre = torch.relu(x.detach().double() @ tower.mlp[0].weight.detach().double().t() + tower.mlp[0].bias.detach().double())
    kept = out.detach() != 0
    assert_close_rel(out.detach()[kept].cpu(), (2.0 * pre[kept]).cpu(), 1e-5, "kept activations are scaled by 1/(1-p)")
    out.sum().backward()
    # the gradient flows exactly through the kept, positive units
    want = (kept.double() * 2.0) @ tower.mlp[0].weight.detach().double()
    assert_close_rel(x.grad.cpu(), want.cpu(), 1e-5, "dropout backward uses the same mask")
