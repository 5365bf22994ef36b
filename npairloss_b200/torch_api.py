"""PyTorch autograd surface over the C ABI ("so the layer is usable outside Caffe").

    loss_fn = NPairLoss(margin_diff=-0.05, an_method=synth.HARD, ...)      # NPairLossParameter fields, caffe.proto:2-23
    loss, tops = loss_fn(embeddings, labels)                                # CUDA fp32 tensors [Q, D], [Q]
    loss.backward()                                                         # through npair_backward

torch is plumbing only (device memory, streams, the autograd graph); forward and backward are the library's kernels.  There
is no CPU path: CPU tensors raise.  `tops` = [loss, top1, top5, top10, feature_asum] as in the reference (.cu:388-401).

WHAT THE BACKWARD RETURNS.  npair_backward reproduces the reference's Backward_gpu (.cu:420-499), which is NOT the analytic
gradient of the loss it reports: at world = 1 it is exactly HALF of it (the 1/2 - 1/2 blend of .cu:492-497), and at
world > 1 the transposed term is additionally divided by the world size (.cu:474).  `true_gradient=True` (world = 1 only) multiplies
by 2 so that torch.autograd.gradcheck-style expectations hold; the default keeps the reference's values, which is what a net
trained with the reference layer sees.

The library context is stateful (S and the row records of the last forward): a second forward through the same module before the
backward of the first would silently change what that backward computes.  Every forward therefore stamps a generation number and
the backward refuses to run against a newer forward.
"""
from __future__ import annotations

import torch

from . import capi


class _NPairFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, feat, label, owner):
        layer = owner._context(feat)
        tops = layer.forward(feat, label)                     # blocks until the five scalars are on the host (as the reference)
        owner._generation += 1
        ctx.owner, ctx.layer, ctx.generation = owner, layer, owner._generation
        ctx.save_for_backward(feat, label)                    # the C ABI wants both unchanged until the backward is enqueued
        t = torch.tensor(tops, dtype=torch.float32, device=feat.device)
        ctx.mark_non_differentiable(t)
        return t[0].clone(), t

    @staticmethod
    def backward(ctx, grad_loss, _grad_tops):
        if ctx.generation != ctx.owner._generation or ctx.layer is not ctx.owner._ctx:
            raise RuntimeError("NPairLoss: another forward ran through this module after the one being differentiated; the library "
                               "context holds the newer batch.  Use one NPairLoss module per outstanding graph.")
        feat, _label = ctx.saved_tensors
        diff = torch.empty_like(feat)
        # the reference scales by top[0]->cpu_diff()[0] (.cu:435): a host scalar, hence the .item()
        ctx.layer.backward(float(grad_loss.item()), diff)
        if ctx.owner._true_gradient:
            diff.mul_(2.0)
        return diff, None, None


class NPairLoss(torch.nn.Module):
    """Module form; the library context is created on first use for the (rows, dims, device) it sees and re-created when
    they change.  Keyword arguments are the fields of capi.make_config (mining regions/methods, margins, SN, precision,
    normalize_input, ...)."""

    def __init__(self, world: int = 1, rank: int = 0, nccl_id: bytes | None = None, true_gradient: bool = False, _context_factory=None, **config):
        super().__init__()
        if true_gradient and world != 1:
            raise ValueError("true_gradient is defined for world = 1 (the reference's multi-rank blend is not a gradient of one loss)")
        self._config, self._world, self._rank, self._nccl_id = dict(config), world, rank, nccl_id
        self._factory = _context_factory or (lambda cfg, nid: capi.Context(cfg, nid))
        self._ctx, self._key = None, None
        self._generation = 0
        self._true_gradient = bool(true_gradient)

    def _context(self, feat):
        q, d = feat.shape[0], feat[0].numel()
        key = (q, d, feat.device.index)
        if key != self._key:
            old = self._ctx
            cfg = capi.make_config(q, d, world=self._world, rank=self._rank, device=feat.device.index or 0, **self._config)
            # the new context is created BEFORE the old one is closed: contexts made with the same NCCL id share one
            # communicator inside the library, which must stay referenced (a unique id can be consumed only once)
            self._ctx, self._key = self._factory(cfg, self._nccl_id), key
            if old is not None and hasattr(old, "close"):
                old.close()
        return self._ctx

    def forward(self, feat, label):
        if feat.dtype != torch.float32:
            raise TypeError("NPairLoss computes in fp32 like the reference (Dtype=float); cast the embeddings")
        feat2 = feat.reshape(feat.shape[0], -1).contiguous()
        label = label.to(torch.float32).contiguous()          # labels are stored as Dtype in the reference (bottom[1])
        loss, tops = _NPairFunction.apply(feat2, label, self)
        return loss, tops
