"""PyTorch autograd surface over the C ABI ("so the layer is usable outside Caffe").

    loss_fn = NPairLoss(margin_diff=-0.05, an_method=synth.HARD, ...)      # NPairLossParameter fields, caffe.proto:2-23
    loss, tops = loss_fn(embeddings, labels)                                # CUDA fp32 tensors [Q, D], [Q]
    loss.backward()                                                         # d loss / d embeddings through npair_backward

torch is plumbing only (device memory, streams, the autograd graph); forward and backward are the library's kernels.  There
is no CPU path: CPU tensors raise.  `tops` = [loss, top1, top5, top10, feature_asum] as in the reference (.cu:388-401)."""
from __future__ import annotations

import torch

from . import capi


class _NPairFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, feat, label, layer):
        tops = layer.forward(feat, label)                     # blocks until the five scalars are on the host (as the reference)
        ctx.layer = layer
        ctx.feat_shape = feat.shape
        ctx.feat_device = feat.device
        t = torch.tensor(tops, dtype=torch.float32, device=feat.device)
        ctx.mark_non_differentiable(t)
        return t[0].clone(), t

    @staticmethod
    def backward(ctx, grad_loss, _grad_tops):
        diff = torch.empty(ctx.feat_shape, dtype=torch.float32, device=ctx.feat_device)
        # the reference scales by top[0]->cpu_diff()[0] (.cu:435): a host scalar, hence the .item()
        ctx.layer.backward(float(grad_loss.item()), diff)
        return diff, None, None


class NPairLoss(torch.nn.Module):
    """Module form; the library context is created on first use for the (rows, dims, device) it sees and re-created when
    they change.  Keyword arguments are the fields of capi.make_config (mining regions/methods, margins, SN, precision)."""

    def __init__(self, world: int = 1, rank: int = 0, nccl_id: bytes | None = None, _context_factory=None, **config):
        super().__init__()
        self._config, self._world, self._rank, self._nccl_id = dict(config), world, rank, nccl_id
        self._factory = _context_factory or (lambda cfg, nid: capi.Context(cfg, nid))
        self._ctx, self._key = None, None

    def _context(self, feat):
        q, d = feat.shape[0], feat[0].numel()
        key = (q, d, feat.device.index)
        if key != self._key:
            if self._ctx is not None and hasattr(self._ctx, "close"):
                self._ctx.close()
            cfg = capi.make_config(q, d, world=self._world, rank=self._rank, device=feat.device.index or 0, **self._config)
            self._ctx, self._key = self._factory(cfg, self._nccl_id), key
        return self._ctx

    def forward(self, feat, label):
        if feat.dtype != torch.float32:
            raise TypeError("NPairLoss computes in fp32 like the reference (Dtype=float); cast the embeddings")
        feat2 = feat.reshape(feat.shape[0], -1).contiguous()
        label = label.to(torch.float32).contiguous()          # labels are stored as Dtype in the reference (bottom[1])
        loss, tops = _NPairFunction.apply(feat2, label, self._context(feat2))
        return loss, tops
