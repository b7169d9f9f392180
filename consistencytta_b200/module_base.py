"""Shared plumbing of the drop-in modules: reference-layout parameter trees and lazily packed kernel operands."""
import torch
from torch import nn


class _Node(nn.Module):
    """Anonymous container so that state_dict() keys reproduce the reference's dotted names."""


def register_tree(root, schema, make_tensor=None):
    """Creates nested submodules / nn.Parameters for every key of `schema` ({key: (shape, kind)})."""
    for key, (shape, _kind) in schema.items():
        parts = key.split(".")
        mod = root
        for p in parts[:-1]:
            if p not in mod._modules:
                mod.add_module(p, _Node())
            mod = mod._modules[p]
        t = make_tensor(key, shape) if make_tensor is not None else torch.zeros(shape)
        mod.register_parameter(parts[-1], nn.Parameter(t, requires_grad=False))


class PackedModule(nn.Module):
    """nn.Module whose kernels consume re-packed (16-bit, K-major, zero-padded) copies of the parameters.

    The packed operands are built lazily on the first CUDA forward and dropped whenever the parameters may have
    changed (load_state_dict, .to(), ._apply)."""

    def __init__(self):
        super().__init__()
        self._pk = None
        self._pk_device = None

    def _invalidate(self):
        self._pk = None
        self._pk_device = None

    def _apply(self, fn, *a, **k):
        self._invalidate()
        return super()._apply(fn, *a, **k)

    def load_state_dict(self, state_dict, strict=True, **kw):
        self._invalidate()
        return super().load_state_dict(state_dict, strict=strict, **kw)

    def packed(self):
        dev = next(self.parameters()).device
        if dev.type != "cuda":
            raise RuntimeError("%s runs on CUDA only (libctta kernels, no CPU fallback); call .to('cuda') first"
                               % type(self).__name__)
        if self._pk is None or self._pk_device != dev:
            with torch.no_grad():
                self._pk = self._pack({k: v.detach() for k, v in self.state_dict().items()}, dev)
            self._pk_device = dev
        return self._pk

    def _pack(self, sd, dev):  # pragma: no cover - abstract
        raise NotImplementedError
