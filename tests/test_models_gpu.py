"""Whole-model parity on the GPU against the golden fixtures of the unmodified reference:
logits, loss (BCE mean + L2) and every parameter gradient; reference tests/test_models.py facts."""

import numpy as np
import pytest
import torch

from deepfm_b200.models import create_model
from tests.golden import spec
from tests.helpers import assert_close_rel, load_golden, split_prefixed, to_dev

pytestmark = pytest.mark.gpu


def _model(name):
    g = load_golden(f"model_{name}.npz")
    model = create_model(name, spec.golden_schema(), spec.golden_config())
    model.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in split_prefixed(g, "param/").items()})
    return g, model.cuda()


@pytest.mark.parametrize("name", ["deepfm", "xdeepfm", "attention_deepfm"])
def test_model_logits_loss_and_grads_match_reference(name):
    g, model = _model(name)
    batch = to_dev(spec.golden_batch())
    labels = torch.from_numpy(spec.golden_labels()).cuda()
    model.train()
    logits = model(batch)
    assert logits.shape == (6, 1)
    assert_close_rel(logits.detach().cpu(), g["logits"], 1e-5, "logits", floor=1e-6)
    loss = torch.nn.BCEWithLogitsLoss()(logits.squeeze(1), labels) + model.get_l2_reg_loss()
    assert abs(loss.item() - float(g["loss"])) < 1e-5 * abs(float(g["loss"]))
    loss.backward()
    ref = split_prefixed(g, "grad/")
    for k, p in model.named_parameters():
        assert p.grad is not None, k
        # analytically-zero grads (W_k.bias; Linear bias feeding BatchNorm) are rounding noise
        assert_close_rel(p.grad.cpu(), ref[k], 1e-4, k, floor=1e-6)
    model.eval()
    with torch.no_grad():
        probs = model.predict(batch)
    assert probs.min() >= 0 and probs.max() <= 1                      # tests/test_models.py:36-41
    assert_close_rel(probs.cpu(), g["probs_eval"], 1e-5, "probs")
    assert model.get_l2_reg_loss().item() > 0                         # tests/test_models.py:43-46


def test_train_step_changes_weights_like_reference_trainer():
    """trainer.py:219-237 step body (BCE + L2, clip, Adam) runs unmodified on the drop-in model."""
    _, model = _model("deepfm")
    batch = to_dev(spec.golden_batch())
    labels = torch.from_numpy(spec.golden_labels()).cuda()
    opt = torch.optim.Adam(model.parameters(), lr=1e-3)
    before = {k: v.clone() for k, v in model.state_dict().items()}
    model.train()
    for _ in range(2):
        loss = torch.nn.BCEWithLogitsLoss()(model(batch).squeeze(1), labels) + model.get_l2_reg_loss()
        opt.zero_grad()
        loss.backward()
        torch.nn.utils.clip_grad_norm_(model.parameters(), 1.0)
        opt.step()
    changed = [k for k, v in model.state_dict().items() if v.dtype.is_floating_point and not torch.equal(v, before[k])]
    assert any(k.startswith("embedding.second_order") for k in changed) and any(k.startswith("dnn") for k in changed)


def test_model_parity_holds_under_cublas_fp32_emulation():
    """bench.py runs the (library) DNN GEMMs on cuBLAS 12.9's FP32 emulation (BF16x9 on the tensor cores, see
    deepfm_b200/fp32_emulation.py).  The swap has to happen before torch is imported, so the whole-model golden
    tests are re-run in a fresh interpreter with it enabled: same fixtures, same tolerances."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = ("import sys\n"
            "from deepfm_b200 import fp32_emulation as E\n"
            "if not E.enable():\n"
            "    print('EMULATION-UNAVAILABLE', E.status()); sys.exit(0)\n"
            "import pytest\n"
            "rc = pytest.main(['-q', '-x', 'tests/test_models_gpu.py', '-m', 'gpu', '-k', 'logits_loss_and_grads'])\n"
            "print('CUBLAS', E.cublas_version())\n"
            "sys.exit(int(rc))\n")
    r = subprocess.run([sys.executable, "-c", code], cwd=root, capture_output=True, text=True, timeout=900)
    if "EMULATION-UNAVAILABLE" in r.stdout:
        pytest.skip(r.stdout.strip())
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    assert "CUBLAS 12.9" in r.stdout or "CUBLAS 12.1" in r.stdout or "CUBLAS 13" in r.stdout, r.stdout[-500:]


def test_l2_value_is_cached_per_table_version_and_sees_every_update():
    """get_l2_reg_loss keeps sum ||W||^2 of the id tables in a device fp64 scalar keyed on (storage, version)
    (layers/l2.py TableNormCache): unchanged tables are not re-read, any write to a table is seen, value and
    gradients are exactly those of the exact reduction."""
    _, model = _model("deepfm")
    batch = to_dev(spec.golden_batch())
    labels = torch.from_numpy(spec.golden_labels()).cuda()
    model.train()
    vals, grads = [], []
    for step in range(3):
        model.zero_grad(set_to_none=True)
        torch.manual_seed(7)                                  # same dropout mask every step
        logits = model(batch)
        l2 = model.get_l2_reg_loss()
        (torch.nn.BCEWithLogitsLoss()(logits.squeeze(1), labels) + l2).backward()
        vals.append(l2.item())
        grads.append(model.embedding.second_order_embeddings["u"].weight.grad.clone())
    assert model.embedding._l2_cache.refreshes == 1           # one exact pass, two cache hits
    assert vals[0] == vals[1] == vals[2]
    assert torch.equal(grads[0], grads[1]) and torch.equal(grads[0], grads[2])
    with torch.no_grad():
        model.embedding.second_order_embeddings["u"].weight.mul_(2.0)
    model.zero_grad(set_to_none=True)
    model(batch)
    assert model.get_l2_reg_loss().item() > vals[0]
    assert model.embedding._l2_cache.refreshes == 2           # the in-place write bumped the version


def test_l2_fold_is_not_double_counted_on_a_second_backward():
    """back-propagating the same graph twice must give the same gradients twice."""
    _, model = _model("deepfm")
    batch = to_dev(spec.golden_batch())
    labels = torch.from_numpy(spec.golden_labels()).cuda()
    model.eval()                                              # deterministic tower
    loss = torch.nn.BCEWithLogitsLoss()(model(batch).squeeze(1), labels) + model.get_l2_reg_loss()
    loss.backward(retain_graph=True)
    first = {k: p.grad.clone() for k, p in model.named_parameters()}
    model.zero_grad(set_to_none=True)
    loss.backward()
    for k, p in model.named_parameters():
        assert torch.equal(p.grad, first[k]), k
