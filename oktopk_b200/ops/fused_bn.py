"""Fused (conv-bias +) BatchNorm2d + ReLU for channels_last fp32 activations (``csrc/bnrelu.cu``).

``bias_bn_relu(x, bn, conv_bias, relu=True)`` replaces ``relu(bn(x + conv_bias))`` in training mode on CUDA with two
small kernels forward and two backward instead of the seven stock ones (bias add, batch-norm, clamp, counter increment;
threshold backward, batch-norm backward, bias-gradient reduction).  The convolution is then called WITHOUT its bias:
a bias in front of a batch-norm cancels exactly (``BN(x + b) = BN(x)``, the batch mean absorbs it), so it only enters
the running-mean update, and its gradient is identically zero (returned as ``None``; autograd's own reduction over
``dy`` returns rounding noise for it).  Parameters, buffers and ``state_dict`` keys are the stock modules' ones.

Falls back to the stock ops whenever the fast path does not apply (CPU, eval mode, non-fp32, not channels_last,
channels not a multiple of 4, cumulative-average momentum).
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn.functional as F

from . import ext


def _fast_path_ok(x: torch.Tensor, bn: torch.nn.BatchNorm2d) -> bool:
    return (x.is_cuda and x.dtype == torch.float32 and x.dim() == 4 and bn.training and bn.affine
            and bn.momentum is not None and x.size(1) % 4 == 0
            and x.is_contiguous(memory_format=torch.channels_last) and ext.available()
            and not torch.is_autocast_enabled())


class _BiasBNReLU(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, gamma, beta, cbias, rmean, rvar, nbt, momentum, eps, relu):
        C = ext.require()
        N, Ch, H, W = x.shape
        M = N * H * W
        y = torch.empty_like(x)                       # preserves channels_last
        nblk = C.bn_num_blocks(M, Ch)
        partial = torch.empty(nblk * 2 * Ch, dtype=torch.float32, device=x.device)
        stats = torch.empty(2 * Ch, dtype=torch.float32, device=x.device)       # [mean | invstd]
        s = torch.cuda.current_stream().cuda_stream
        C.bn_forward(x.data_ptr(), y.data_ptr(), partial.data_ptr(), gamma.data_ptr(), beta.data_ptr(),
                     0 if cbias is None else cbias.data_ptr(), stats.data_ptr(), stats.data_ptr() + 4 * Ch,
                     0 if rmean is None else rmean.data_ptr(), 0 if rvar is None else rvar.data_ptr(),
                     0 if nbt is None else nbt.data_ptr(), float(momentum), float(eps), int(relu), M, Ch, s)
        ctx.save_for_backward(x, gamma, beta, stats)
        ctx.relu = bool(relu)
        return y

    @staticmethod
    def backward(ctx, dy):
        C = ext.require()
        x, gamma, beta, stats = ctx.saved_tensors
        N, Ch, H, W = x.shape
        M = N * H * W
        if not dy.is_contiguous(memory_format=torch.channels_last):
            dy = dy.contiguous(memory_format=torch.channels_last)
        dx = torch.empty_like(x)
        nblk = C.bn_num_blocks(M, Ch)
        partial = torch.empty(nblk * 2 * Ch, dtype=torch.float32, device=x.device)
        dgb = torch.empty(2 * Ch, dtype=torch.float32, device=x.device)         # [dgamma | dbeta]
        s = torch.cuda.current_stream().cuda_stream
        C.bn_backward(x.data_ptr(), dy.data_ptr(), dx.data_ptr(), partial.data_ptr(), gamma.data_ptr(), beta.data_ptr(),
                      stats.data_ptr(), stats.data_ptr() + 4 * Ch, dgb.data_ptr(), dgb.data_ptr() + 4 * Ch,
                      int(ctx.relu), M, Ch, s)
        # conv bias: the loss does not depend on it (see module docstring) -> no gradient
        return dx, dgb[:Ch], dgb[Ch:], None, None, None, None, None, None, None


def bias_bn_relu(x: torch.Tensor, bn: torch.nn.BatchNorm2d, conv_bias: Optional[torch.Tensor] = None,
                 relu: bool = True) -> torch.Tensor:
    """``relu(bn(x + conv_bias))`` (``relu=False``: without the ReLU)."""
    if _fast_path_ok(x, bn):
        track = bn.track_running_stats and bn.running_mean is not None
        return _BiasBNReLU.apply(x, bn.weight, bn.bias, conv_bias, bn.running_mean if track else None,
                                 bn.running_var if track else None, bn.num_batches_tracked if track else None,
                                 bn.momentum, bn.eps, relu)
    if conv_bias is not None:
        x = x + conv_bias.view(1, -1, 1, 1)
    y = bn(x)
    return F.relu(y) if relu else y


def conv_bn_relu(x: torch.Tensor, conv: torch.nn.Conv2d, bn: torch.nn.BatchNorm2d, relu: bool = True) -> torch.Tensor:
    """``relu(bn(conv(x)))`` with the convolution's bias folded into the fused batch-norm when the fast path applies."""
    if _fast_path_ok_pre(x, conv, bn):
        z = F.conv2d(x, conv.weight, None, conv.stride, conv.padding, conv.dilation, conv.groups)
        if _fast_path_ok(z, bn):
            return bias_bn_relu(z, bn, conv.bias, relu)
        if conv.bias is not None:
            z = z + conv.bias.view(1, -1, 1, 1)
        y = bn(z)
        return F.relu(y) if relu else y
    y = bn(conv(x))
    return F.relu(y) if relu else y


def _fast_path_ok_pre(x: torch.Tensor, conv: torch.nn.Conv2d, bn: torch.nn.BatchNorm2d) -> bool:
    return (x.is_cuda and x.dtype == torch.float32 and bn.training and bn.affine and bn.momentum is not None
            and conv.out_channels % 4 == 0 and conv.padding_mode == "zeros" and ext.available()
            and x.is_contiguous(memory_format=torch.channels_last) and not torch.is_autocast_enabled()
            and not isinstance(conv.padding, str))


class _MaxPool2x2(torch.autograd.Function):
    """2x2 / stride-2 max pooling on channels_last fp32 (``csrc/bnrelu.cu``): the backward pass writes every input position
    (no memset, no atomics: stride-2 windows do not overlap)."""

    @staticmethod
    def forward(ctx, x):
        C = ext.require()
        N, Ch, H, W = x.shape
        y = torch.empty((N, Ch, H // 2, W // 2), dtype=x.dtype, device=x.device, memory_format=torch.channels_last)
        arg = torch.empty(y.numel(), dtype=torch.uint8, device=x.device)
        C.maxpool2_fwd(x.data_ptr(), y.data_ptr(), arg.data_ptr(), N, H, W, Ch, torch.cuda.current_stream().cuda_stream)
        ctx.save_for_backward(arg)
        ctx.shape = (N, Ch, H, W)
        return y

    @staticmethod
    def backward(ctx, dy):
        C = ext.require()
        (arg,) = ctx.saved_tensors
        N, Ch, H, W = ctx.shape
        if not dy.is_contiguous(memory_format=torch.channels_last):
            dy = dy.contiguous(memory_format=torch.channels_last)
        dx = torch.empty((N, Ch, H, W), dtype=dy.dtype, device=dy.device, memory_format=torch.channels_last)
        C.maxpool2_bwd(dy.data_ptr(), arg.data_ptr(), dx.data_ptr(), N, H, W, Ch, torch.cuda.current_stream().cuda_stream)
        return dx


def max_pool_2x2(x: torch.Tensor, pool: torch.nn.MaxPool2d) -> torch.Tensor:
    ks = pool.kernel_size if isinstance(pool.kernel_size, tuple) else (pool.kernel_size,) * 2
    st = pool.stride if isinstance(pool.stride, tuple) else (pool.stride,) * 2
    pad = pool.padding if isinstance(pool.padding, tuple) else (pool.padding,) * 2
    dil = pool.dilation if isinstance(pool.dilation, tuple) else (pool.dilation,) * 2
    if (ks == (2, 2) and st == (2, 2) and pad == (0, 0) and dil == (1, 1) and not pool.return_indices
            and x.is_cuda and x.dtype == torch.float32 and x.dim() == 4 and x.size(1) % 4 == 0 and x.size(2) % 2 == 0
            and x.size(3) % 2 == 0 and x.is_contiguous(memory_format=torch.channels_last) and ext.available()
            and torch.is_grad_enabled() and not torch.is_autocast_enabled()):
        return _MaxPool2x2.apply(x)
    return pool(x)


def run_fused_sequential(seq: torch.nn.Sequential, x: torch.Tensor) -> torch.Tensor:
    """Run an ``nn.Sequential`` fusing every ``Conv2d -> BatchNorm2d [-> ReLU]`` run it contains."""
    mods = list(seq.children())
    i = 0
    while i < len(mods):
        m = mods[i]
        if isinstance(m, torch.nn.Conv2d) and i + 1 < len(mods) and isinstance(mods[i + 1], torch.nn.BatchNorm2d):
            relu = i + 2 < len(mods) and isinstance(mods[i + 2], torch.nn.ReLU)
            x = conv_bn_relu(x, m, mods[i + 1], relu)
            i += 3 if relu else 2
        elif isinstance(m, torch.nn.MaxPool2d):
            x = max_pool_2x2(x, m)
            i += 1
        elif isinstance(m, torch.nn.AvgPool2d) and m.kernel_size in (1, (1, 1)) and m.stride in (1, (1, 1)):
            i += 1                                     # 1x1 average pool: the identity (VGG/models/vgg.py:37), no kernel
        else:
            x = m(x)
            i += 1
    return x
