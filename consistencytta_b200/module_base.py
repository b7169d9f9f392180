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


_PACK_CLOCK = [0]


def next_pack_version():
    """Process-wide monotonically increasing stamp: a consumer that baked packed-operand POINTERS into something
    long-lived (a captured CUDA graph) compares stamps to learn that the operands were re-packed (and freed)."""
    _PACK_CLOCK[0] += 1
    return _PACK_CLOCK[0]


class PackedModule(nn.Module):
    """nn.Module whose kernels consume re-packed (16-bit, K-major, zero-padded) copies of the parameters.

    The packed operands are built lazily on the first CUDA forward and dropped whenever the parameters may have
    changed: load_state_dict (its own or a PARENT module's — caught by a load_state_dict post-hook), .to(), .half(),
    ._apply.  `pack_version` changes every time that happens; SingleStepEngine keys its captured graphs on it.
    Parameters mutated in place (`p.data.copy_`) are invisible to nn.Module: call `_invalidate()` afterwards."""

    def __init__(self):
        super().__init__()
        self._pk = {}            # operand dtype -> packed operands (fp16 by default; bf16 when a stage was escalated)
        self._pk_device = None
        self.pack_version = next_pack_version()
        self.register_load_state_dict_post_hook(lambda module, incompatible_keys: module._invalidate())

    def _invalidate(self):
        self._pk = {}
        self._pk_device = None
        self.pack_version = next_pack_version()

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
        from . import ops
        dt = ops.OPERAND_DTYPE
        if self._pk_device != dev:
            self._pk = {}
            self._pk_device = dev
        if dt not in self._pk:   # entries are never replaced: captured graphs hold their pointers
            with torch.no_grad():
                self._pk[dt] = self._pack({k: v.detach() for k, v in self.state_dict().items()}, dev)
        return self._pk[dt]

    def to_prepacked(self, device):
        """Moves the module to `device` with the operands packed ON THE HOST first: the packing arithmetic (transposes,
        zero padding, 16-bit casts — about a thousand small tensor ops per model) runs on the CPU and every packed operand
        reaches the GPU as one host-to-device copy, so no kernel is launched before the first hot-path kernel (and the
        caching allocator sees no packing temporaries)."""
        device = torch.device(device)
        with torch.no_grad():
            cpu_sd = {k: v.detach().cpu() for k, v in self.state_dict().items()}
            pk = self._pack(cpu_sd, torch.device("cpu"))
        from . import ops
        self.to(device)            # parameters: plain H2D copies; invalidates any previous pack
        self._pk = {ops.OPERAND_DTYPE: tree_to(pk, device)}
        self._pk_device = next(self.parameters()).device
        return self

    def _pack(self, sd, dev):  # pragma: no cover - abstract
        raise NotImplementedError


def tree_to(obj, device):
    """Copies every tensor inside a packed-operand structure (dicts / lists / tuples / objects with tensor attributes)
    to `device`; everything else (ints, ctypes arrays, offsets) passes through."""
    if torch.is_tensor(obj):
        return obj.to(device)
    if isinstance(obj, dict):
        return {k: tree_to(v, device) for k, v in obj.items()}
    if isinstance(obj, tuple):
        return tuple(tree_to(v, device) for v in obj)
    if isinstance(obj, list):
        return [tree_to(v, device) for v in obj]
    if hasattr(obj, "__dict__") and not isinstance(obj, (type, nn.Module)):
        for k, v in list(vars(obj).items()):
            if torch.is_tensor(v) or isinstance(v, (dict, list, tuple)) or (hasattr(v, "__dict__") and not callable(v)):
                setattr(obj, k, tree_to(v, device))
        return obj
    return obj
