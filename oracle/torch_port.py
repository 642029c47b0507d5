"""PyTorch-CPU port of the reference's train step (TEST / BASELINE INFRASTRUCTURE ONLY).

The reference is pure PyTorch and cannot travel to the GPU box (``/root/reference`` exists
only in the build container), so the CPU baseline timed by ``bench.py`` is this port: the
same ATen CPU ops, issued in the same order and granularity as the reference's modules
(one ``embedding``/``embedding_bag``/``linear`` per field, un-fused FM, materialised CIN
outer product + ``conv1d(k=1)``, matmul/softmax attention, Linear-BatchNorm-act-Dropout
tower, BCE-with-logits mean loss, the all-parameters L2 term, global-norm clip, dense Adam).
It is written functionally over a flat ``{state_dict key: tensor}`` dict instead of
``nn.Module`` classes.  ``tests/test_oracle.py`` checks it against the golden fixtures made
from the unmodified reference (``tests/golden/make_golden.py``).

Citations are to the reference tree: embedding.py:76-126, fm.py:18-23, cin.py:66-105,
attention.py:91-120, dnn.py:45-59, deepfm.py:30-42, xdeepfm.py:36-48,
attention_deepfm.py:48-66, base.py:78-83, trainer.py:212-240.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py`` (cpu_baseline / --impl reference)
may import this file.
"""

from __future__ import annotations

import math
from typing import Dict, List

import torch
import torch.nn.functional as F


def _kind(field) -> str:
    ft = field.feature_type
    return ft.value if hasattr(ft, "value") else str(ft)


# ------------------------------------------------------------------ parameter construction

def _xavier(shape, gen):
    fan_out, fan_in = shape[0], shape[1]
    bound = math.sqrt(6.0 / (fan_in + fan_out))
    return (torch.rand(shape, generator=gen) * 2 - 1) * bound


def init_embedding_params(schema, fm_embed_dim: int, gen: torch.Generator) -> Dict[str, torch.Tensor]:
    """Same parameter set / shapes / init family as FeatureEmbedding (embedding.py:20-74):
    xavier-uniform on rows 1.. of every table (row 0 stays zero), xavier + zero bias Linears."""
    p: Dict[str, torch.Tensor] = {}
    for name, f in schema.fields.items():
        k, d = _kind(f), f.embedding_dim
        if k == "dense":
            p[f"second_order_embeddings.{name}.weight"] = _xavier((d, 1), gen)
            p[f"second_order_embeddings.{name}.bias"] = torch.zeros(d)
            p[f"first_order_embeddings.{name}.weight"] = _xavier((1, 1), gen)
            p[f"first_order_embeddings.{name}.bias"] = torch.zeros(1)
        else:
            V = f.vocabulary_size
            w2 = torch.zeros(V, d)
            w2[1:] = _xavier((V - 1, d), gen)
            w1 = torch.zeros(V, 1)
            w1[1:] = _xavier((V - 1, 1), gen)
            p[f"second_order_embeddings.{name}.weight"] = w2
            p[f"first_order_embeddings.{name}.weight"] = w1
        if d != fm_embed_dim:
            p[f"projections.{name}.weight"] = _xavier((fm_embed_dim, d), gen)
    return p


def init_dnn_params(prefix: str, in_dim: int, hidden: List[int], use_bn: bool, gen) -> Dict[str, torch.Tensor]:
    """nn.Sequential(Linear, [BatchNorm1d], act, Dropout)* keys as dnn.py:45-55 lays them out."""
    p = {}
    idx = 0
    for h in hidden:
        bound = 1.0 / math.sqrt(in_dim)
        p[f"{prefix}.mlp.{idx}.weight"] = (torch.rand(h, in_dim, generator=gen) * 2 - 1) * bound
        p[f"{prefix}.mlp.{idx}.bias"] = (torch.rand(h, generator=gen) * 2 - 1) * bound
        idx += 1
        if use_bn:
            p[f"{prefix}.mlp.{idx}.weight"] = torch.ones(h)
            p[f"{prefix}.mlp.{idx}.bias"] = torch.zeros(h)
            p[f"{prefix}.mlp.{idx}.running_mean"] = torch.zeros(h)
            p[f"{prefix}.mlp.{idx}.running_var"] = torch.ones(h)
            idx += 1
        idx += 2  # activation, dropout
        in_dim = h
    return p


def init_linear(prefix: str, in_dim: int, out_dim: int, gen) -> Dict[str, torch.Tensor]:
    bound = 1.0 / math.sqrt(in_dim)
    return {f"{prefix}.weight": (torch.rand(out_dim, in_dim, generator=gen) * 2 - 1) * bound,
            f"{prefix}.bias": (torch.rand(out_dim, generator=gen) * 2 - 1) * bound}


# ------------------------------------------------------------------ forward pieces

def embedding_views(schema, p, batch, prefix: str = "embedding."):
    """FeatureEmbedding.forward (embedding.py:76-126): python loop over fields, 2-3 ATen calls each."""
    fo_parts, fe_parts, flat_parts = [], [], []
    for name, f in schema.fields.items():
        k = _kind(f)
        x = batch[name]
        w2 = p[f"{prefix}second_order_embeddings.{name}.weight"]
        w1 = p[f"{prefix}first_order_embeddings.{name}.weight"]
        if k == "dense":
            xin = x.unsqueeze(-1)
            raw = F.linear(xin, w2, p[f"{prefix}second_order_embeddings.{name}.bias"])
            fo = F.linear(xin, w1, p[f"{prefix}first_order_embeddings.{name}.bias"])
        elif k == "sequence":
            raw = F.embedding_bag(x, w2, mode=f.combiner, padding_idx=0)
            fo = F.embedding_bag(x, w1, mode=f.combiner, padding_idx=0)
        else:
            raw = F.embedding(x, w2, padding_idx=0)
            fo = F.embedding(x, w1, padding_idx=0)
        fo_parts.append(fo)
        flat_parts.append(raw)
        pk = f"{prefix}projections.{name}.weight"
        fe_parts.append(F.linear(raw, p[pk]) if pk in p else raw)
    first_order = torch.stack(fo_parts, dim=1).sum(dim=1)
    return first_order, torch.stack(fe_parts, dim=1), torch.cat(flat_parts, dim=-1)


def fm_interaction(e):
    """fm.py:18-23, un-fused (7 ATen ops)."""
    sq_of_sum = e.sum(dim=1).pow(2)
    sum_of_sq = e.pow(2).sum(dim=1)
    return 0.5 * (sq_of_sum - sum_of_sq).sum(dim=1, keepdim=True)


def cin_network(x0, weights, biases, direct_sizes, next_sizes, split_half):
    """cin.py:66-105: einsum outer product, reshape, conv1d(k=1), relu, split, sum-pool."""
    B, _, D = x0.shape
    hidden, outs = x0, []
    n = len(weights)
    for i in range(n):
        outer = torch.einsum("bhd,bfd->bhfd", hidden, x0).reshape(B, -1, D)
        comp = torch.relu(F.conv1d(outer, weights[i], biases[i]))
        if split_half and i < n - 1:
            direct, hidden = comp.split([direct_sizes[i], next_sizes[i]], dim=1)
        else:
            direct = hidden = comp
        outs.append(direct.sum(dim=2))
    return torch.cat(outs, dim=1)


def attention_block(x, p, prefix, num_heads, use_residual):
    """attention.py:91-120."""
    B, Fn, D = x.shape
    q = F.linear(x, p[f"{prefix}W_q.weight"], p[f"{prefix}W_q.bias"])
    k = F.linear(x, p[f"{prefix}W_k.weight"], p[f"{prefix}W_k.bias"])
    v = F.linear(x, p[f"{prefix}W_v.weight"], p[f"{prefix}W_v.bias"])
    A = q.shape[-1]
    hd = A // num_heads
    q = q.view(B, Fn, num_heads, hd).transpose(1, 2)
    k = k.view(B, Fn, num_heads, hd).transpose(1, 2)
    v = v.view(B, Fn, num_heads, hd).transpose(1, 2)
    w = torch.softmax(torch.matmul(q, k.transpose(-2, -1)) / math.sqrt(hd), dim=-1)
    o = torch.matmul(w, v).transpose(1, 2).contiguous().view(B, Fn, -1)
    out = F.linear(o, p[f"{prefix}W_out.weight"], p[f"{prefix}W_out.bias"])
    if use_residual:
        out = F.layer_norm(out + x, (D,), p[f"{prefix}layer_norm.weight"], p[f"{prefix}layer_norm.bias"])
    return out


_ACT = {"relu": torch.relu, "leaky_relu": F.leaky_relu, "gelu": F.gelu, "tanh": torch.tanh}


def dnn_tower(x, p, prefix, n_hidden, activation, dropout, use_bn, training):
    """dnn.py:45-59: Linear -> BatchNorm1d -> act -> Dropout, stacked."""
    idx = 0
    for _ in range(n_hidden):
        x = F.linear(x, p[f"{prefix}.mlp.{idx}.weight"], p[f"{prefix}.mlp.{idx}.bias"])
        idx += 1
        if use_bn:
            x = F.batch_norm(x, p[f"{prefix}.mlp.{idx}.running_mean"], p[f"{prefix}.mlp.{idx}.running_var"],
                             p[f"{prefix}.mlp.{idx}.weight"], p[f"{prefix}.mlp.{idx}.bias"],
                             training=training, momentum=0.1, eps=1e-5)
            idx += 1
        x = _ACT[activation](x)
        x = F.dropout(x, dropout, training=training)
        idx += 2
    return x


# ------------------------------------------------------------------ whole models

class PortedModel:
    """Functional stand-in for the reference's DeepFM / xDeepFM / AttentionDeepFM.

    ``params`` is keyed exactly like the reference model's ``state_dict`` so weights can be
    loaded from the golden fixtures (or from the CUDA drop-in) verbatim.
    """

    def __init__(self, name: str, schema, config, seed: int = 0, params=None):
        if name not in ("deepfm", "xdeepfm", "attention_deepfm"):
            raise ValueError(f"Unknown model: {name}")          # models/__init__.py:32-35
        self.name, self.schema, self.cfg = name, schema, config
        gen = torch.Generator().manual_seed(seed)
        D = config.feature.fm_embed_dim
        Fn, T = len(schema.fields), sum(f.embedding_dim for f in schema.fields.values())
        p = {f"embedding.{k}": v for k, v in init_embedding_params(schema, D, gen).items()}
        dnn_in = T
        if name == "xdeepfm":
            from .deepfm_oracle import cin_plan
            self.direct, self.next, ks = cin_plan(Fn, config.cin.layer_sizes, config.cin.split_half)
            for i, (L, K) in enumerate(zip(config.cin.layer_sizes, ks)):
                b = 1.0 / math.sqrt(K)
                p[f"cin.conv_layers.{i}.weight"] = (torch.rand(L, K, 1, generator=gen) * 2 - 1) * b
                p[f"cin.conv_layers.{i}.bias"] = (torch.rand(L, generator=gen) * 2 - 1) * b
            p.update(init_linear("cin_linear", sum(self.direct), 1, gen))
        if name == "attention_deepfm":
            a = config.attention
            if a.attention_dim % a.num_heads:
                raise ValueError("attention_dim must be divisible by num_heads")
            for li in range(a.num_layers):
                pre = f"attention.layers.{li}."
                for w in ("W_q", "W_k", "W_v"):
                    p.update(init_linear(pre + w, D, a.attention_dim, gen))
                p.update(init_linear(pre + "W_out", a.attention_dim, D, gen))
                if a.use_residual:
                    p[pre + "layer_norm.weight"] = torch.ones(D)
                    p[pre + "layer_norm.bias"] = torch.zeros(D)
            dnn_in = Fn * D + T
        p.update(init_dnn_params("dnn", dnn_in, config.dnn.hidden_units, config.dnn.use_batch_norm, gen))
        head = "dnn_linear" if name == "xdeepfm" else "output_linear"
        p.update(init_linear(head, config.dnn.hidden_units[-1], 1, gen))
        self.head = head
        if params is not None:
            missing = set(p) - set(params)
            extra = {k for k in set(params) - set(p) if "num_batches_tracked" not in k}
            if missing or extra:
                raise KeyError(f"state_dict mismatch: missing {sorted(missing)} extra {sorted(extra)}")
            p = {k: torch.as_tensor(params[k]).clone() for k in p}
        self.buffers = {k for k in p if "running_" in k}
        self.params = {k: (v if k in self.buffers else v.requires_grad_(True)) for k, v in p.items()}
        self.training = True

    def trainable(self):
        return [v for k, v in self.params.items() if k not in self.buffers]

    def forward(self, batch):
        p, c = self.params, self.cfg
        fo, fe, flat = embedding_views(self.schema, p, batch)
        d = c.dnn
        if self.name == "deepfm":
            tower = dnn_tower(flat, p, "dnn", len(d.hidden_units), d.activation, d.dropout, d.use_batch_norm, self.training)
            return fo + fm_interaction(fe) + F.linear(tower, p["output_linear.weight"], p["output_linear.bias"])
        if self.name == "xdeepfm":
            n = len(c.cin.layer_sizes)
            cin = cin_network(fe, [p[f"cin.conv_layers.{i}.weight"] for i in range(n)],
                              [p[f"cin.conv_layers.{i}.bias"] for i in range(n)],
                              self.direct, self.next, c.cin.split_half)
            tower = dnn_tower(flat, p, "dnn", len(d.hidden_units), d.activation, d.dropout, d.use_batch_norm, self.training)
            return (fo + F.linear(cin, p["cin_linear.weight"], p["cin_linear.bias"])
                    + F.linear(tower, p["dnn_linear.weight"], p["dnn_linear.bias"]))
        if self.name == "attention_deepfm":
            a = c.attention
            x = fe
            for li in range(a.num_layers):
                x = attention_block(x, p, f"attention.layers.{li}.", a.num_heads, a.use_residual)
            dnn_in = torch.cat([x.reshape(x.size(0), -1), flat], dim=1)
            tower = dnn_tower(dnn_in, p, "dnn", len(d.hidden_units), d.activation, d.dropout, d.use_batch_norm, self.training)
            return fo + fm_interaction(fe) + F.linear(tower, p["output_linear.weight"], p["output_linear.bias"])
        raise ValueError(f"Unknown model: {self.name}")

    def l2_reg_loss(self):
        """base.py:78-83: python loop of norm(2).pow(2) over every embedding parameter."""
        total = torch.tensor(0.0)
        for k, v in self.params.items():
            if k.startswith("embedding."):
                total = total + v.norm(2).pow(2)
        return self.cfg.feature.embedding_l2_reg * total

    def loss(self, batch, labels):
        """trainer.py:219-225."""
        logits = self.forward(batch).squeeze(1)
        loss = F.binary_cross_entropy_with_logits(logits, labels)
        if self.cfg.feature.embedding_l2_reg > 0:
            loss = loss + self.l2_reg_loss()
        return loss

    def train_step(self, batch, labels, optimizer):
        """trainer.py:219-237: fwd, loss, L2, zero_grad, backward, clip, step."""
        loss = self.loss(batch, labels)
        optimizer.zero_grad()
        loss.backward()
        clip = self.cfg.training.gradient_clip_norm
        if clip > 0:
            torch.nn.utils.clip_grad_norm_(self.trainable(), clip)
        optimizer.step()
        return loss
