"""Forward multiply-accumulate counter (capability parity with ``BERT/ptflops/flops_counter.py:19-37``:
``get_model_complexity_info(model, input_res)``; the reference calls it once per stage at start-up,
``BERT/bert/main_bert.py:861``).

Hook-based: one forward pass with forward hooks on the leaf modules that do arithmetic.  Counts MACs
(1 MAC = 2 FLOPs) for Conv/Linear/LSTM/Embedding-free matmuls, element counts for norm / activation / pooling.
``count_flops(model, *inputs)`` takes real example inputs (any signature) instead of an input resolution tuple.
"""
from __future__ import annotations

from typing import Dict, Tuple

import torch
import torch.nn as nn


def _numel(x) -> int:
    return x.numel() if torch.is_tensor(x) else 0


def _conv_macs(m: nn.modules.conv._ConvNd, out: torch.Tensor) -> int:
    k = 1
    for s in m.kernel_size:
        k *= s
    per_out = k * (m.in_channels // m.groups)
    return out.numel() * per_out + (out.numel() if m.bias is not None else 0)


def _lstm_macs(m: nn.LSTM, inp, out) -> int:
    x = inp[0]
    if isinstance(x, nn.utils.rnn.PackedSequence):
        steps = int(x.data.size(0))                   # sum over batch of sequence lengths
    else:
        steps = int(x.numel() // x.size(-1))
    d = 2 if m.bidirectional else 1
    macs = 0
    for layer in range(m.num_layers):
        inp_size = m.input_size if layer == 0 else m.hidden_size * d
        macs += d * 4 * m.hidden_size * (inp_size + m.hidden_size) * steps
    return macs


def count_flops(model: nn.Module, *inputs, **kwargs) -> Tuple[int, int, Dict[str, int]]:
    """Returns ``(macs, params, per_module_type_macs)`` of one forward pass."""
    per_type: Dict[str, int] = {}
    handles = []

    def add(name: str, v: int) -> None:
        per_type[name] = per_type.get(name, 0) + int(v)

    def hook(m, inp, out):
        o = out[0] if isinstance(out, (tuple, list)) else out
        if isinstance(m, nn.modules.conv._ConvNd):
            add("conv", _conv_macs(m, o))
        elif isinstance(m, nn.Linear):
            add("linear", _numel(o) * m.in_features + (_numel(o) if m.bias is not None else 0))
        elif isinstance(m, nn.LSTM):
            add("lstm", _lstm_macs(m, inp, out))
        elif isinstance(m, (nn.modules.batchnorm._BatchNorm, nn.LayerNorm, nn.GroupNorm)):
            add("norm", 2 * _numel(o))
        elif isinstance(m, (nn.ReLU, nn.ReLU6, nn.Hardtanh, nn.GELU, nn.Tanh, nn.Sigmoid, nn.Softmax)):
            add("act", _numel(o))
        elif isinstance(m, (nn.MaxPool2d, nn.AvgPool2d, nn.AdaptiveAvgPool2d)):
            add("pool", _numel(inp[0]))

    for mod in model.modules():
        if len(list(mod.children())) == 0 or isinstance(mod, nn.LSTM):
            handles.append(mod.register_forward_hook(hook))
    was_training = model.training
    model.eval()
    try:
        with torch.no_grad():
            model(*inputs, **kwargs)
    finally:
        for h in handles:
            h.remove()
        model.train(was_training)
    # attention score/context matmuls are functional (no module): add them for our BERT layers analytically
    try:
        from ..models.bert import BertLayer
        layers = [m for m in model.modules() if isinstance(m, BertLayer)]
        if layers and inputs:
            B, S = inputs[0].shape[:2]
            H = layers[0].attention.h * layers[0].attention.dh
            if H:
                add("attention_matmul", len(layers) * 2 * B * S * S * H)
    except Exception:  # noqa: BLE001
        pass
    params = sum(p.numel() for p in model.parameters())
    return sum(per_type.values()), params, per_type


def get_model_complexity_info(model: nn.Module, input_res, print_per_layer_stat: bool = False, as_strings: bool = True,
                              input_constructor=None):
    """ptflops-compatible entry point."""
    if input_constructor is not None:
        kw = input_constructor(input_res)
        macs, params, per = count_flops(model, **kw)
    else:
        dev = next(model.parameters()).device
        macs, params, per = count_flops(model, torch.zeros((1,) + tuple(input_res), device=dev))
    if print_per_layer_stat:
        for k, v in sorted(per.items(), key=lambda kv: -kv[1]):
            print("%-18s %12.3f MMac" % (k, v / 1e6))
    if as_strings:
        return flops_to_string(macs), params_to_string(params)
    return macs, params


def flops_to_string(macs: float, precision: int = 2) -> str:
    for unit, div in (("GMac", 1e9), ("MMac", 1e6), ("KMac", 1e3)):
        if macs >= div:
            return "%.*f %s" % (precision, macs / div, unit)
    return "%d Mac" % macs


def params_to_string(n: int) -> str:
    if n >= 1e6:
        return "%.2f M" % (n / 1e6)
    if n >= 1e3:
        return "%.2f k" % (n / 1e3)
    return str(n)
