"""Host-side plumbing shared by the module surface: activation codes, fp32 weight views,
workspace cache, mask normalisation and the ctypes struct builders."""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch
import torch.nn as nn

from . import _lib as L


# ---------------------------------------------------------------------------------------------
class Swish(nn.Module):
    """speechbrain.nnet.activations.Swish: x * sigmoid(beta * x).  Used as an activation *class* argument
    (Conformer.py:117,404); the arithmetic runs inside the CUDA kernels (only beta == 1 is supported)."""

    def __init__(self, beta: float = 1.0):
        super().__init__()
        self.beta = beta

    def forward(self, x):
        raise RuntimeError("Swish is evaluated inside libsmx kernels; it is not a standalone op in this package")


def act_code(act: nn.Module) -> int:
    """Map an activation module instance to smx_act; loud failure for anything the kernels do not implement."""
    if isinstance(act, Swish) or type(act).__name__ == "Swish":
        if float(getattr(act, "beta", 1.0)) != 1.0:
            raise NotImplementedError("Swish with beta != 1 is not implemented in libsmx")
        return L.ACT_SWISH
    if isinstance(act, nn.SiLU):
        return L.ACT_SWISH
    if isinstance(act, nn.GELU):
        return L.ACT_GELU_TANH if getattr(act, "approximate", "none") == "tanh" else L.ACT_GELU
    if isinstance(act, nn.ReLU):
        return L.ACT_RELU
    if isinstance(act, nn.LeakyReLU):
        if abs(act.negative_slope - 0.01) > 1e-12:
            raise NotImplementedError("LeakyReLU with negative_slope != 0.01 is not implemented in libsmx")
        return L.ACT_LEAKY_RELU
    if isinstance(act, nn.Tanh):
        return L.ACT_TANH
    if isinstance(act, nn.Sigmoid):
        return L.ACT_SIGMOID
    if isinstance(act, nn.Identity):
        return L.ACT_IDENTITY
    raise NotImplementedError(f"activation {type(act).__name__} is not implemented in libsmx")


# ---------------------------------------------------------------------------------------------
def require_cuda(x: torch.Tensor, what: str) -> None:
    if not x.is_cuda:
        raise RuntimeError(
            f"{what}: input is on {x.device}; summarymixing_b200 runs on CUDA (sm_100a) only — there is no CPU path"
        )


def dtype_code(x: torch.Tensor) -> int:
    if x.dtype == torch.float32:
        return L.F32
    if x.dtype == torch.bfloat16:
        return L.BF16
    raise NotImplementedError(f"activation dtype {x.dtype} is not supported (float32 or bfloat16)")


def stream_ptr(device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


class WeightView:
    """fp32, contiguous, on-device views of a module's parameters, rebuilt only when a parameter changes.

    The library takes weights in the reference's own state_dict layouts as fp32 device pointers.  If the
    module was cast (.bfloat16()) or lives on another device, an fp32 copy is kept here instead.
    """

    def __init__(self):
        self._key = None
        self._keep = []
        self._struct = None

    @property
    def struct(self):
        return self._struct

    @struct.setter
    def struct(self, value):
        # the ctypes struct holds raw device pointers into the tensors of _keep (fp32 copies, packed operand images):
        # whoever holds the struct (an autograd node between forward and backward, a captured graph) holds them too
        if value is not None:
            value._keepalive = self._keep
        self._struct = value

    def stale(self, params, device) -> bool:
        key = (str(device),) + tuple((p.data_ptr(), p._version, p.dtype) for p in params)
        if key != self._key:
            self._key = key
            self._keep = []
            return True
        return False

    def ptr(self, p: Optional[torch.Tensor], device) -> Optional[int]:
        if p is None:
            return None
        t = p.detach()
        if t.dtype != torch.float32 or t.device != device or not t.is_contiguous():
            t = t.to(device=device, dtype=torch.float32).contiguous()
        self._keep.append(t)
        return t.data_ptr()


def invalidate_weights(module: nn.Module) -> int:
    """Drop the cached weight views / packed tensor-core images of `module` and its sub-modules, so that the next call rebuilds them
    from the parameters' current values.  The caches are keyed on the parameters' storage and version counters: every tracked
    in-place update (an optimizer step, load_state_dict, p.copy_()) is seen automatically, but a write that bypasses the version
    counter (p.data.mul_(...), an EMA or a weight clip through .data, raw pointer writes) is not — call this after such a write
    (and re-capture any GraphedForward / HostPipeline built on the module: a captured graph keeps the old images' addresses).
    Returns the number of caches dropped."""
    n = 0
    for m in module.modules():
        wv = getattr(m, "_wv", None)
        if isinstance(wv, WeightView):
            wv._key = None
            n += 1
    return n


_workspaces = {}


def workspace(device, nbytes: int) -> torch.Tensor:
    """Grow-only per-(device, stream) scratch buffer handed to the library (it allocates nothing itself)."""
    key = (device.index, torch.cuda.current_stream(device).cuda_stream)
    buf = _workspaces.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(max(nbytes, 1), dtype=torch.uint8, device=device)
        _workspaces[key] = buf
    return buf


def mask_u8(mask: Optional[torch.Tensor], B: int, T: int, device) -> Optional[torch.Tensor]:
    """(B,T) bool/float padding mask, 1/True = valid (TransformerASR.py:158-162) -> uint8 on device."""
    if mask is None:
        return None
    if tuple(mask.shape) != (B, T):
        raise RuntimeError(f"src_padding_mask must have shape ({B}, {T}), got {tuple(mask.shape)}")
    return (mask != 0).to(device=device, dtype=torch.uint8).contiguous()


def sum_mask_f32(sum_mask: Optional[torch.Tensor], T: int, device) -> Optional[torch.Tensor]:
    """(T,T) mask -> fp32 (the reference does sum_mask.float(), summary_mixing.py:188-189)."""
    if sum_mask is None:
        return None
    if tuple(sum_mask.shape) != (T, T):
        raise RuntimeError(f"sum_mask must have shape ({T}, {T}), got {tuple(sum_mask.shape)}")
    return sum_mask.to(device=device, dtype=torch.float32).contiguous()


def p_or_none(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def check_grad_mode(module: nn.Module) -> None:
    """The inference entry points refuse a training-mode call under autograd that the differentiable path did not take (nothing
    requires grad: no trainable parameter, input detached): training-mode behaviour (dropout, layerdrop) lives on the autograd path."""
    if module.training and torch.is_grad_enabled():
        raise NotImplementedError(
            "summarymixing_b200: this call is in training mode with autograd enabled, but neither the input nor any parameter requires "
            "grad, so it is not on the differentiable (training) path; call .eval() or wrap the call in torch.no_grad() for inference"
        )


def pack_tc(wv: WeightView, device, struct, bytes_fn, pack_fn) -> None:
    """Build the bf16 operand image for the tensor-core arm (once per weight set) and hang it on the
    weights struct.  Configurations the arm does not handle report 0 bytes and stay on the generic arm."""
    struct.packed = None
    nbytes = int(bytes_fn(C.byref(struct)))
    if nbytes == 0:
        return
    buf = torch.empty(nbytes + 1024, dtype=torch.uint8, device=device)
    ptr = (buf.data_ptr() + 1023) & ~1023
    with torch.cuda.device(device):
        L.check(pack_fn(C.byref(struct), ptr, nbytes, stream_ptr(device)))
    wv._keep.append(buf)
    struct.packed = ptr


def fill_linear(dst: L.Linear, wv: WeightView, device, w, b, in_dim, out_dim, n_split=1):
    dst.w = wv.ptr(w, device)
    dst.b = wv.ptr(b, device)
    dst.in_dim, dst.out_dim, dst.n_split = int(in_dim), int(out_dim), int(n_split)
