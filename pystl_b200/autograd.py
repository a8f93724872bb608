"""``torch.autograd.Function`` around the forward / backward entry points of the C ABI."""

from __future__ import annotations

import torch

from . import capi


class RobustnessFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, signal: torch.Tensor, sem, cf):  # signal [B, T, D]
        home = signal.device
        dev = signal.device if signal.is_cuda else torch.device("cuda", torch.cuda.current_device())
        sig = signal.detach().to(device=dev, dtype=torch.float32).contiguous()
        out = sem._forward_raw(cf, sig, None)
        ctx.save_for_backward(sig)
        ctx.sem, ctx.cf, ctx.home, ctx.in_dtype = sem, cf, home, signal.dtype
        return out

    @staticmethod
    def backward(ctx, grad_out: torch.Tensor):
        (sig,) = ctx.saved_tensors
        B, T, D = sig.shape
        g = grad_out.to(device=sig.device, dtype=torch.float32).contiguous()
        grad_signal = torch.empty_like(sig)
        stream = torch.cuda.current_stream(sig.device).cuda_stream
        with torch.cuda.device(sig.device):
            rc = capi.lib().pystl_b200_eval_backward(
                ctx.cf.ptr(), sig.data_ptr(), B, T, D, g.data_ptr(), 1.0, grad_signal.data_ptr(), stream
            )
        capi.raise_for(rc, capi.lib().pystl_b200_last_error())
        return grad_signal.to(device=ctx.home, dtype=ctx.in_dtype), None, None
