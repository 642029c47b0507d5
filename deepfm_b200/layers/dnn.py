"""MLP tower (reference: deepfm/models/layers/dnn.py:8-59) on the repo's own kernels.

Same constructor, ``mlp`` ``nn.Sequential`` layout (hence ``state_dict`` keys ``mlp.<i>.weight`` ...), ``output_dim``
and ``ValueError``s as the reference.  The ``nn.Linear`` / ``nn.BatchNorm1d`` / activation / ``nn.Dropout`` children
are parameter containers; on CUDA tensors every ``Linear -> BatchNorm1d -> act -> Dropout`` block is ONE autograd node:

  forward   y = x W^T + b            dfm_gemm3 mode 0 (tcgen05, 3xTF32: fp32-accurate on the tensor cores)
            mean / rstd              dfm_bn_stats (training-mode batch statistics, running stats updated)
            a = dropout(act(BN(y)))  dfm_bn_act_fwd (one pass; mask = counter-based function of (seed, index))
  backward  dy, dgamma, dbeta, db    dfm_bn_act_bwd (z, xhat, mask recomputed from y: nothing else is saved)
            dW = dy^T x              dfm_gemm3 mode 2 (both operands MN-major straight out of memory, split-K,
                                     fixed-order reduction: deterministic)
            dx = dy W                dfm_gemm3 mode 1

``linear_head(module, x)`` runs the models' final ``nn.Linear(., 1)`` the same way (dfm_head_fwd / dfm_head_bwd).
CPU tensors (and ``DNN.fused = False``) take the plain ``nn.Sequential`` path -- the tower is outside the embedding
hot path, so unlike the embedding kernels it keeps an eager route (used by CPU-side bookkeeping tests).
"""

from __future__ import annotations

from typing import List, Optional

import torch
import torch.nn as nn

from .. import _lib

_ACTIVATIONS = {"relu": nn.ReLU, "leaky_relu": nn.LeakyReLU, "gelu": nn.GELU, "tanh": nn.Tanh}
_ACT_CODE = {"relu": 0, "leaky_relu": 1, "gelu": 2, "tanh": 3}
BN_NONE, BN_BATCH, BN_FIXED = 0, 1, 2


def _gemm3(mode: int, a: torch.Tensor, b: torch.Tensor, out: torch.Tensor, bias: Optional[torch.Tensor], M: int, N: int, K: int):
    lib = _lib.lib()
    ws_bytes = lib.dfm_gemm3_workspace_bytes(mode, M, N, K)
    ws = torch.empty((max(ws_bytes, 16),), device=a.device, dtype=torch.uint8)
    _lib.check(lib.dfm_gemm3(mode, a.data_ptr(), b.data_ptr(), out.data_ptr(), _lib.ptr(bias), M, N, K, ws.data_ptr(),
                             ws.numel(), _lib.stream_ptr()), "dfm_gemm3")
    return out


def gemm3_supported(M: int, N: int, K: int) -> bool:
    """All three products of a Linear(K -> N) on M rows need 16-byte-aligned contiguous extents."""
    return M > 0 and K % 4 == 0 and N % 4 == 0


class _LinearBnActFn(torch.autograd.Function):
    """One tower block: Linear (+ BatchNorm1d) + activation + dropout."""

    @staticmethod
    def forward(ctx, x, weight, bias, gamma, beta, cfg):
        # cfg: (bn_mode, act, drop_p, seed, eps, momentum, running_mean, running_var, fixed_mean, fixed_rstd)
        bn, act, p, seed, eps, momentum, run_mean, run_var, fmean, frstd = cfg
        lib = _lib.lib()
        x = x.contiguous()
        M, K = x.shape
        N = weight.shape[0]
        dev = x.device
        y = torch.empty((M, N), device=dev, dtype=torch.float32)
        _gemm3(0, x, weight, y, bias, M, N, K)
        mean = rstd = None
        ws = torch.empty((lib.dfm_tower_workspace_bytes(M, N),), device=dev, dtype=torch.uint8)
        if bn == BN_BATCH:
            mean = torch.empty((N,), device=dev, dtype=torch.float32)
            rstd = torch.empty((N,), device=dev, dtype=torch.float32)
            _lib.check(lib.dfm_bn_stats(y.data_ptr(), M, N, float(eps), mean.data_ptr(), rstd.data_ptr(), _lib.ptr(run_mean),
                                        _lib.ptr(run_var), float(momentum), ws.data_ptr(), ws.numel(), _lib.stream_ptr()),
                       "dfm_bn_stats")
        elif bn == BN_FIXED:
            mean, rstd = fmean, frstd
        out = torch.empty((M, N), device=dev, dtype=torch.float32)
        _lib.check(lib.dfm_bn_act_fwd(y.data_ptr(), M, N, bn, act, _lib.ptr(mean), _lib.ptr(rstd), _lib.ptr(gamma),
                                      _lib.ptr(beta), float(p), int(seed), out.data_ptr(), _lib.stream_ptr()), "dfm_bn_act_fwd")
        ctx.cfg = (bn, act, p, seed)
        ctx.has_bias = bias is not None
        ctx.save_for_backward(x, weight, y, mean, rstd, gamma, beta)
        return out

    @staticmethod
    def backward(ctx, g_out):
        x, weight, y, mean, rstd, gamma, beta = ctx.saved_tensors
        bn, act, p, seed = ctx.cfg
        lib = _lib.lib()
        g_out = g_out.contiguous()
        M, K = x.shape
        N = weight.shape[0]
        dev = x.device
        dy = torch.empty((M, N), device=dev, dtype=torch.float32)
        dgamma = torch.empty((N,), device=dev, dtype=torch.float32) if bn != BN_NONE else None
        dbeta = torch.empty((N,), device=dev, dtype=torch.float32) if bn != BN_NONE else None
        dbias = torch.empty((N,), device=dev, dtype=torch.float32) if ctx.has_bias else None
        ws = torch.empty((lib.dfm_tower_workspace_bytes(M, N),), device=dev, dtype=torch.uint8)
        _lib.check(lib.dfm_bn_act_bwd(g_out.data_ptr(), y.data_ptr(), M, N, bn, act, _lib.ptr(mean), _lib.ptr(rstd),
                                      _lib.ptr(gamma), _lib.ptr(beta), float(p), int(seed), dy.data_ptr(), _lib.ptr(dgamma),
                                      _lib.ptr(dbeta), _lib.ptr(dbias), ws.data_ptr(), ws.numel(), _lib.stream_ptr()),
                   "dfm_bn_act_bwd")
        dw = dx = None
        if ctx.needs_input_grad[1]:
            dw = torch.empty((N, K), device=dev, dtype=torch.float32)
            _gemm3(2, dy, x, dw, None, N, K, M)              # dW[n][k] = sum_m dy[m][n] x[m][k]
        if ctx.needs_input_grad[0]:
            dx = torch.empty((M, K), device=dev, dtype=torch.float32)
            _gemm3(1, dy, weight, dx, None, M, K, N)         # dx[m][k] = sum_n dy[m][n] W[n][k]
        return dx, dw, dbias, dgamma, dbeta, None


class _TowerFn(torch.autograd.Function):
    """The whole tower as ONE autograd node over two C-ABI calls (dfm_tower_fwd / dfm_tower_bwd sequence the same per-block
    kernels): the host side of a step -- three nodes, ~45 ctypes calls, ~50 allocations -- was what bounded multi-rank steps."""

    @staticmethod
    def forward(ctx, x, cfg, *params):
        # cfg: (dims, bn modes, act, p, seeds, eps list, momentum list, running tensors [2 per block or None], fixed pairs)
        dims, bn, act, p, seeds, eps, mom, running, fixed = cfg
        lib = _lib.lib()
        n = len(bn)
        x = x.contiguous()
        M = x.shape[0]
        dev = x.device
        c_dims = _lib.i64_array(dims)
        store = torch.empty((lib.dfm_tower_store_floats(n, c_dims, M, None),), device=dev, dtype=torch.float32)
        out = torch.empty((M, dims[-1]), device=dev, dtype=torch.float32)
        ws = torch.empty((lib.dfm_tower_seq_workspace_bytes(n, c_dims, M, 0),), device=dev, dtype=torch.uint8)
        run_or_fixed = [(fixed[i] if bn[i // 2] == BN_FIXED else running[i]) for i in range(2 * n)]
        c_params = _lib.ptr_array(list(params))
        c_bn = _lib.i32_array(bn)
        c_seeds = (_lib.C.c_uint64 * n)(*[int(v) for v in seeds])
        _lib.check(lib.dfm_tower_fwd(n, c_dims, M, x.data_ptr(), c_params, c_bn, _lib.ptr_array(run_or_fixed),
                                     (_lib.C.c_float * n)(*eps), (_lib.C.c_float * n)(*mom), act, float(p), c_seeds,
                                     store.data_ptr(), out.data_ptr(), ws.data_ptr(), ws.numel(), _lib.stream_ptr()), "dfm_tower_fwd")
        ctx.cfg = (dims, bn, act, p, seeds, fixed)
        ctx.params = params                      # inputs of the node: plain references
        ctx.save_for_backward(x, store)
        return out

    @staticmethod
    def backward(ctx, g_out):
        x, store = ctx.saved_tensors
        dims, bn, act, p, seeds, fixed = ctx.cfg
        params = ctx.params
        lib = _lib.lib()
        n = len(bn)
        M = x.shape[0]
        dev = x.device
        g_out = g_out.contiguous()
        c_dims = _lib.i64_array(dims)
        grads = [None if t is None else torch.empty_like(t) for t in params]
        dx = torch.empty((M, dims[0]), device=dev, dtype=torch.float32) if ctx.needs_input_grad[0] else None
        ws = torch.empty((lib.dfm_tower_seq_workspace_bytes(n, c_dims, M, 1),), device=dev, dtype=torch.uint8)
        _lib.check(lib.dfm_tower_bwd(n, c_dims, M, x.data_ptr(), _lib.ptr_array(list(params)), _lib.i32_array(bn),
                                     _lib.ptr_array(list(fixed)), act, float(p), (_lib.C.c_uint64 * n)(*[int(v) for v in seeds]),
                                     store.data_ptr(), g_out.data_ptr(), _lib.ptr(dx), _lib.ptr_array(grads), ws.data_ptr(),
                                     ws.numel(), _lib.stream_ptr()), "dfm_tower_bwd")
        return (dx, None, *grads)


class _HeadFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias):
        lib = _lib.lib()
        x = x.contiguous()
        M, Cn = x.shape
        out = torch.empty((M, 1), device=x.device, dtype=torch.float32)
        _lib.check(lib.dfm_head_fwd(x.data_ptr(), weight.data_ptr(), _lib.ptr(bias), M, Cn, out.data_ptr(), _lib.stream_ptr()),
                   "dfm_head_fwd")
        ctx.has_bias = bias is not None
        ctx.save_for_backward(x, weight)
        return out

    @staticmethod
    def backward(ctx, g):
        x, weight = ctx.saved_tensors
        lib = _lib.lib()
        g = g.contiguous()
        M, Cn = x.shape
        dev = x.device
        dx = torch.empty_like(x) if ctx.needs_input_grad[0] else None
        dw = torch.empty_like(weight)
        db = torch.empty((1,), device=dev, dtype=torch.float32) if ctx.has_bias else None
        ws = torch.empty((lib.dfm_tower_workspace_bytes(M, Cn),), device=dev, dtype=torch.uint8)
        _lib.check(lib.dfm_head_bwd(x.data_ptr(), weight.data_ptr(), g.data_ptr(), M, Cn, _lib.ptr(dx), dw.data_ptr(),
                                    _lib.ptr(db), ws.data_ptr(), ws.numel(), _lib.stream_ptr()), "dfm_head_bwd")
        return dx, dw, db


def linear_head(module: nn.Linear, x: torch.Tensor) -> torch.Tensor:
    """``module(x)`` for the models' final ``nn.Linear(width, 1)`` on the repo's kernels (CUDA, fp32, 2-D input)."""
    if (not x.is_cuda) or module.out_features != 1 or x.dim() != 2 or x.dtype != torch.float32 or x.shape[0] == 0 \
            or not DNN.fused:
        return module(x)
    return _HeadFn.apply(x, module.weight, module.bias)


class DNN(nn.Module):
    ACTIVATIONS = _ACTIVATIONS
    fused = True            # class-wide switch: False runs the plain nn.Sequential (library GEMMs) everywhere
    single_call = True      # the whole tower as one autograd node over dfm_tower_fwd / dfm_tower_bwd (False: one node per block)

    def __init__(self, input_dim: int, hidden_units: List[int], activation: str = "relu",
                 dropout: float = 0.1, use_batch_norm: bool = True) -> None:
        super().__init__()
        if not hidden_units:
            raise ValueError("hidden_units must be non-empty")
        act = _ACTIVATIONS.get(activation.lower())
        if act is None:
            raise ValueError(f"Unknown activation: {activation}. Choose from {list(_ACTIVATIONS)}")
        blocks: List[nn.Module] = []
        width = input_dim
        self._blocks = []       # (linear index, bn index or None, dropout index) inside self.mlp
        for units in hidden_units:
            li = len(blocks)
            blocks.append(nn.Linear(width, units))
            bi = None
            if use_batch_norm:
                bi = len(blocks)
                blocks.append(nn.BatchNorm1d(units))
            blocks += [act(), nn.Dropout(p=dropout)]
            self._blocks.append((li, bi, len(blocks) - 1))
            width = units
        self.mlp = nn.Sequential(*blocks)
        self.output_dim = hidden_units[-1]
        self._act_code = _ACT_CODE[activation.lower()]

    def _block_fused(self, x: torch.Tensor, li: int, bi: Optional[int], di: int) -> torch.Tensor:
        lin: nn.Linear = self.mlp[li]
        bn_mod: Optional[nn.BatchNorm1d] = self.mlp[bi] if bi is not None else None
        p = float(self.mlp[di].p) if self.training else 0.0
        seed = 0
        if p > 0.0:             # one 63-bit seed per block and step from torch's CPU generator (follows torch.manual_seed)
            seed = int(torch.randint(0, 2 ** 62, (1,)).item())
        bn, gamma, beta = BN_NONE, None, None
        eps, momentum, run_mean, run_var, fmean, frstd = 1e-5, 0.1, None, None, None, None
        if bn_mod is not None:
            gamma, beta, eps = bn_mod.weight, bn_mod.bias, bn_mod.eps
            use_batch = self.training or bn_mod.running_mean is None
            if use_batch:
                bn = BN_BATCH
                if self.training and bn_mod.track_running_stats and bn_mod.running_mean is not None:
                    bn_mod.num_batches_tracked.add_(1)
                    momentum = bn_mod.momentum if bn_mod.momentum is not None else 1.0 / float(bn_mod.num_batches_tracked.item())
                    run_mean, run_var = bn_mod.running_mean, bn_mod.running_var
            else:
                bn = BN_FIXED
                fmean = bn_mod.running_mean
                frstd = torch.rsqrt(bn_mod.running_var + eps)
        cfg = (bn, self._act_code, p, seed, eps, momentum, run_mean, run_var, fmean, frstd)
        return _LinearBnActFn.apply(x, lin.weight, lin.bias, gamma, beta, cfg)

    def _tower_single_call(self, x: torch.Tensor) -> Optional[torch.Tensor]:
        """Every block in one autograd node (``_TowerFn``); None when some block needs the per-block route."""
        dims, bn, eps, mom, running, fixed, params = [x.shape[1]], [], [], [], [], [], []
        p = None
        for li, bi, di in self._blocks:
            lin: nn.Linear = self.mlp[li]
            bn_mod = self.mlp[bi] if bi is not None else None
            if not gemm3_supported(x.shape[0], lin.out_features, lin.in_features):
                return None
            pp = float(self.mlp[di].p) if self.training else 0.0
            if p is not None and pp != p:
                return None                      # one dropout probability per tower (the reference's DNN has one)
            p = pp
            mode, e_, m_, rm, rv, fm, fr = BN_NONE, 1e-5, 0.1, None, None, None, None
            if bn_mod is not None:
                if not (bn_mod.affine and bn_mod.weight is not None):
                    return None
                e_ = bn_mod.eps
                if self.training or bn_mod.running_mean is None:
                    mode = BN_BATCH
                    if self.training and bn_mod.track_running_stats and bn_mod.running_mean is not None:
                        if bn_mod.momentum is None:
                            return None          # cumulative moving average needs the step count on the host
                        m_, rm, rv = bn_mod.momentum, bn_mod.running_mean, bn_mod.running_var
                else:
                    mode = BN_FIXED
                    fm, fr = bn_mod.running_mean, torch.rsqrt(bn_mod.running_var + e_)
            dims.append(lin.out_features)
            bn.append(mode); eps.append(float(e_)); mom.append(float(m_))
            running += [rm, rv]; fixed += [fm, fr]
            params += [lin.weight, lin.bias, bn_mod.weight if bn_mod is not None else None,
                       bn_mod.bias if bn_mod is not None else None]
        seeds = [0] * len(bn)
        if p and p > 0.0:        # one 63-bit seed per block and step from torch's CPU generator (follows torch.manual_seed)
            seeds = torch.randint(0, 2 ** 62, (len(bn),)).tolist()
        if self.training:
            nbt = [self.mlp[bi].num_batches_tracked for _, bi, _ in self._blocks
                   if bi is not None and self.mlp[bi].track_running_stats and self.mlp[bi].num_batches_tracked is not None]
            if nbt:
                torch._foreach_add_(nbt, 1)
        cfg = (dims, bn, self._act_code, p or 0.0, seeds, eps, mom, running, fixed)
        return _TowerFn.apply(x, cfg, *params)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if not (DNN.fused and x.is_cuda and x.dim() == 2 and x.dtype == torch.float32 and x.shape[0] > 0):
            return self.mlp(x)
        if DNN.single_call:
            out = self._tower_single_call(x)
            if out is not None:
                return out
        for li, bi, di in self._blocks:
            lin = self.mlp[li]
            bn_ok = bi is None or (self.mlp[bi].affine and self.mlp[bi].weight is not None)
            if gemm3_supported(x.shape[0], lin.out_features, lin.in_features) and bn_ok:
                x = self._block_fused(x, li, bi, di)
            else:               # a width the TMA path cannot address (not a multiple of 4): library route for this block
                for k in range(li, di + 1):
                    x = self.mlp[k](x)
        return x
