"""Generate tests/golden/*.npz from the UNMODIFIED reference (build container only).

    python tests/golden/make_golden.py

Imports ``/root/reference`` (read-only) with a stand-in for the one missing third-party
module (``dacite.from_dict`` is only called by ``load_config``, never here), runs the
reference's own modules on the seeded cases of ``spec.py`` and stores inputs, weights, outputs
and autograd gradients.  The committed fixtures are what the oracle and the CUDA path are
pinned to; ``/root/reference`` is not needed to run the tests.
"""

import os
import sys
import tempfile
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
shim = types.ModuleType("dacite")
shim.from_dict = lambda data_class, data: (_ for _ in ()).throw(NotImplementedError)
sys.modules["dacite"] = shim
sys.path.insert(0, "/root/reference")
sys.path.insert(0, ROOT)   # ahead of the reference so `tests` resolves to this repo

from deepfm.config import ExperimentConfig as RefConfig                       # noqa: E402
from deepfm.data.schema import DatasetSchema as RefSchema                     # noqa: E402
from deepfm.data.schema import FeatureType as RefType, FieldSchema as RefField  # noqa: E402
from deepfm.models import create_model                                         # noqa: E402
from deepfm.models.layers.attention import MultiHeadSelfAttention             # noqa: E402
from deepfm.models.layers.cin import CIN                                       # noqa: E402
from deepfm.models.layers.embedding import FeatureEmbedding                   # noqa: E402
from deepfm.models.layers.fm import FMInteraction                             # noqa: E402

from tests.golden import spec                                                  # noqa: E402


def npd(state):
    return {k: v.detach().cpu().numpy().copy() for k, v in state.items()}


def save(name, **arrays):
    path = os.path.join(HERE, name)
    np.savez_compressed(path, **arrays)
    print(f"wrote {path}: {len(arrays)} arrays, {os.path.getsize(path)} bytes")


def main():
    torch.manual_seed(1234)
    torch.set_num_threads(1)
    schema = spec.golden_schema(RefSchema, RefField, RefType)
    batch_np = spec.golden_batch()
    batch = {k: torch.from_numpy(v) for k, v in batch_np.items()}
    B = len(batch_np["u"])
    D = spec.FM_DIM
    F, T = schema.num_fields, schema.total_embedding_dim

    # ---- FeatureEmbedding: outputs and grads for random upstream gradients
    emb = FeatureEmbedding(schema, fm_embed_dim=D)
    with torch.no_grad():  # make row 0 non-zero in one table: forward must return it as stored
        emb.second_order_embeddings["u"].weight[0] = 0.25
        emb.first_order_embeddings["i"].weight[0] = -0.5
    fo, fe, fl = emb(batch)
    g_fo, g_fe, g_fl = torch.randn_like(fo), torch.randn_like(fe), torch.randn_like(fl)
    l2 = spec.L2_REG * sum(p.norm(2).pow(2) for p in emb.parameters())
    ((fo * g_fo).sum() + (fe * g_fe).sum() + (fl * g_fl).sum() + l2).backward()
    arrays = {f"param/{k}": v for k, v in npd(emb.state_dict()).items()}
    arrays.update({f"grad/{k}": p.grad.numpy() for k, p in emb.named_parameters()})
    arrays.update({f"batch/{k}": v for k, v in batch_np.items()})
    arrays.update(first_order=fo.detach().numpy(), field_embeddings=fe.detach().numpy(),
                  flat=fl.detach().numpy(), g_first=g_fo.numpy(), g_field=g_fe.numpy(),
                  g_flat=g_fl.numpy(), l2_loss=np.float64(l2.item()))
    save("embedding.npz", **arrays)

    # ---- FM
    x = torch.randn(5, 7, 6, requires_grad=True)
    out = FMInteraction()(x)
    g = torch.randn_like(out)
    out.backward(g)
    ex = torch.tensor([[[1., 2.], [3., 4.], [5., 6.]]])
    save("fm.npz", x=x.detach().numpy(), out=out.detach().numpy(), g=g.numpy(), gx=x.grad.numpy(),
         example_in=ex.numpy(), example_out=FMInteraction()(ex).numpy())

    # ---- CIN (split and no-split)
    for tag, sizes, split in (("split", [6, 5, 4], True), ("nosplit", [4, 3], False), ("one", [5], True)):
        cin = CIN(num_fields=4, embed_dim=6, layer_sizes=sizes, split_half=split)
        x = torch.randn(3, 4, 6, requires_grad=True)
        out = cin(x)
        g = torch.randn_like(out)
        out.backward(g)
        arrays = dict(x=x.detach().numpy(), out=out.detach().numpy(), g=g.numpy(), gx=x.grad.numpy(),
                      split=np.int64(split), sizes=np.array(sizes))
        for i, c in enumerate(cin.conv_layers):
            arrays[f"w{i}"] = c.weight.detach().numpy()
            arrays[f"b{i}"] = c.bias.detach().numpy()
            arrays[f"gw{i}"] = c.weight.grad.numpy()
            arrays[f"gb{i}"] = c.bias.grad.numpy()
        save(f"cin_{tag}.npz", **arrays)

    # ---- attention (2 layers residual; 1 layer no residual)
    for tag, kw in (("res", dict(embed_dim=8, num_heads=2, attention_dim=8, num_layers=2, use_residual=True)),
                    ("nores", dict(embed_dim=6, num_heads=3, attention_dim=12, num_layers=1, use_residual=False))):
        att = MultiHeadSelfAttention(**kw)
        with torch.no_grad():
            for p in att.parameters():          # default LN weight=1/bias=0 would hide bugs
                p.add_(0.1 * torch.randn_like(p))
        x = torch.randn(3, 5, kw["embed_dim"], requires_grad=True)
        out = att(x)
        g = torch.randn_like(out)
        out.backward(g)
        arrays = dict(x=x.detach().numpy(), out=out.detach().numpy(), g=g.numpy(), gx=x.grad.numpy(),
                      num_heads=np.int64(kw["num_heads"]), use_residual=np.int64(kw["use_residual"]),
                      num_layers=np.int64(kw["num_layers"]))
        arrays.update({f"param/{k}": v for k, v in npd(att.state_dict()).items()})
        arrays.update({f"grad/{k}": p.grad.numpy() for k, p in att.named_parameters()})
        save(f"attention_{tag}.npz", **arrays)

    # ---- whole models: logits, loss (BCE mean + L2) and every parameter gradient
    cfg = spec.golden_config(RefConfig)
    labels = torch.from_numpy(spec.golden_labels())
    for name in ("deepfm", "xdeepfm", "attention_deepfm"):
        torch.manual_seed(99)
        model = create_model(name, schema, cfg)
        model.train()
        before = npd(model.state_dict())      # BatchNorm running stats BEFORE the training forward
        logits = model(batch)
        loss = torch.nn.BCEWithLogitsLoss()(logits.squeeze(1), labels) + model.get_l2_reg_loss()
        loss.backward()
        arrays = {f"param/{k}": v for k, v in before.items()}
        arrays.update({f"grad/{k}": p.grad.numpy() for k, p in model.named_parameters()})
        arrays.update(logits=logits.detach().numpy(), loss=np.float64(loss.item()))
        model.eval()
        with torch.no_grad():
            arrays["probs_eval"] = model.predict(batch).numpy()
        save(f"model_{name}.npz", **arrays)


if __name__ == "__main__":
    main()
