"""``StageRuntime``: the data-parallel branch of the reference's PipeDream-derived BERT runtime
(``BERT/runtime.py:55-1029``; live part: ``initialize`` DP branch ``:128-151``, ``run_forward`` ``:540-640``,
``run_backward`` ``:642-721``, ``run_training_loop_with_flushes`` ``:842-900``).

The reference instantiates *every* stage module on every rank and runs them back to back (pipeline parallelism is
commented out); a "model" is a list of ``(module, input_names, output_names)`` with the criterion last
(``BERT/bert/models/bert/depth=4/__init__.py:12-19``).  This class keeps that calling convention for users of the
reference API -- ``r.run_forward(); r.run_backward(); optimizer.step()`` with ``update_interval`` micro-batches
between flushes, optional activation recomputation (``--recompute_step``) and bf16 autocast (``--fp16`` in the
reference is apex O3) -- on top of plain autograd.  The B200 ``Trainer`` does not need it (it calls the fused
``BertForPreTraining`` directly); it exists for API parity and is exercised by the tests.
"""
from __future__ import annotations

import time
from typing import Callable, Dict, Iterable, List, Optional, Sequence, Tuple

import torch
from torch.utils.checkpoint import checkpoint as _checkpoint

BERT = "bert"
IMAGE_CLASSIFICATION = "image_classification"


class RuntimeStats:
    """Bytes sent/received per direction (all zero in data-parallel mode, ``BERT/runtime_utilities.py:4-27``)."""

    def __init__(self, forward: bool):
        self.forward = forward
        self.stats = {"compute_time": 0.0, "send_tensors": 0.0, "send_tensors_size": 0, "receive_tensors": 0.0,
                      "receive_tensors_size": 0}

    def reset_stats(self) -> None:
        for k in self.stats:
            self.stats[k] = 0.0 if isinstance(self.stats[k], float) else 0


class InputSource:
    """Endless micro-batch source keyed by tensor name (``BERT/bert/main_bert.py:616-639``)."""

    def __init__(self, loader: Iterable, names: Sequence[str] = ("input0", "input1", "input2", "target_lm", "target_nsp"),
                 device: Optional[torch.device] = None):
        self.loader, self.names, self.device = loader, tuple(names), device
        self.it = iter(loader)

    def get_inputs(self) -> Dict[str, torch.Tensor]:
        try:
            batch = next(self.it)
        except StopIteration:
            self.it = iter(self.loader)
            batch = next(self.it)
        out = {}
        for n, t in zip(self.names, batch):
            out[n] = t.to(self.device, non_blocking=True) if self.device is not None else t
        return out


class StageRuntime:
    def __init__(self, model: List[Tuple[torch.nn.Module, Sequence[str], Sequence[str]]], distributed_backend=None,
                 fp16: bool = False, loss_scale: float = 1.0, training_tensor_shapes=None, eval_tensor_shapes=None,
                 training_tensor_dtypes=None, inputs_module_destinations=None, target_tensor_names=("target_lm", "target_nsp"),
                 configuration_maps=None, master_addr=None, rank: int = 0, local_rank: int = 0, num_ranks_in_server: int = 1,
                 verbose_freq: int = 0, model_type: str = BERT, enable_recompute: bool = False,
                 device: Optional[torch.device] = None):
        self.model = model
        self.modules_with_dependencies = [(m, list(i), list(o)) for m, i, o in model]
        self.fp16, self.loss_scale = fp16, loss_scale
        self.target_tensor_names = tuple(target_tensor_names)
        self.rank, self.local_rank = rank, local_rank
        self.model_type = model_type
        self.enable_recompute = enable_recompute
        self.verbose_freq = verbose_freq
        self.device = device or torch.device("cuda" if torch.cuda.is_available() else "cpu")
        # data parallelism: one stage, all ranks (``conf_32nodes.json``: {"0": [0..31]})
        self.stage, self.num_stages = 0, 1
        self.comm_handler = None
        self.forward_stats, self.backward_stats = RuntimeStats(True), RuntimeStats(False)
        self.tensors: List[Dict[str, torch.Tensor]] = []
        self.loss = None
        self.forward_minibatch_id = self.backward_minibatch_id = 0
        self.input_source: Optional[InputSource] = None
        self.epoch_start_time = time.time()
        for m, _, _ in self.modules_with_dependencies:
            m.to(self.device)

    # ------------------------------------------------------------------ module access (optimizer construction)
    def modules(self) -> List[torch.nn.Module]:
        return [m for m, _, _ in self.modules_with_dependencies]

    def named_parameters(self):
        for i, m in enumerate(self.modules()):
            for n, p in m.named_parameters():
                yield "stages.%d.%s" % (i, n), p

    def parameters(self):
        for _, p in self.named_parameters():
            yield p

    def state_dict(self) -> Dict:
        return {"module%d" % i: m.state_dict() for i, m in enumerate(self.modules())}

    def load_state_dict(self, sd: Dict) -> None:
        for i, m in enumerate(self.modules()):
            m.load_state_dict(sd["module%d" % i])

    def train(self) -> None:
        self.tensors, self.forward_minibatch_id, self.backward_minibatch_id = [], 0, 0
        for m in self.modules():
            m.train()

    def eval(self) -> None:
        for m in self.modules():
            m.eval()

    def set_input_source(self, src: InputSource) -> None:
        self.input_source = src

    def set_loader(self, loader) -> None:
        self.input_source = InputSource(loader, device=self.device)

    # ------------------------------------------------------------------ forward / backward
    def run_forward(self, recompute_step: bool = False) -> torch.Tensor:
        assert self.input_source is not None, "call set_input_source()/set_loader() first"
        tensors = self.input_source.get_inputs()
        t0 = time.perf_counter()
        ctx = torch.autocast(self.device.type, dtype=torch.bfloat16) if self.fp16 else _Null()
        with ctx:
            for module, in_names, out_names in self.modules_with_dependencies[:-1]:
                args = [tensors[n] for n in in_names]
                if (self.enable_recompute or recompute_step) and any(torch.is_tensor(a) and a.requires_grad for a in args):
                    outs = _checkpoint(module, *args, use_reentrant=False)
                else:
                    outs = module(*args)
                if not isinstance(outs, (tuple, list)):
                    outs = (outs,)
                for n, o in zip(out_names, outs):
                    tensors[n] = o
            crit, in_names, out_names = self.modules_with_dependencies[-1]
            loss = crit(*[tensors[n] for n in in_names])
        tensors[out_names[0] if out_names else "loss"] = loss
        self.loss = loss
        self.tensors.append(tensors)
        self.forward_minibatch_id += 1
        self.forward_stats.stats["compute_time"] += time.perf_counter() - t0
        return loss

    def run_backward(self) -> None:
        tensors = self.tensors.pop(0)
        t0 = time.perf_counter()
        loss = tensors.get("loss", self.loss)
        torch.autograd.backward(loss * self.loss_scale if self.loss_scale != 1.0 else loss)
        self.backward_minibatch_id += 1
        self.backward_stats.stats["compute_time"] += time.perf_counter() - t0

    # ------------------------------------------------------------------ the loop (runtime.py:842-900)
    def run_training_loop_with_flushes(self, num_minibatches: int, optimizer, recompute_step: bool = False,
                                       update_interval: int = 1, log: Optional[Callable[[str], None]] = None) -> float:
        """``update_interval`` forward passes, the same number of backward passes, ``optimizer.step()``, ``zero_grad()``;
        communication is suppressed on all but the last micro-batch through ``optimizer.local``."""
        self.train()
        n_updates = num_minibatches // update_interval
        t_start = time.perf_counter()
        last = 0.0
        for u in range(n_updates):
            optimizer.zero_grad()
            for j in range(update_interval):
                if hasattr(optimizer, "local"):
                    optimizer.local = j < update_interval - 1
                loss = self.run_forward(recompute_step)
                self.run_backward()
            optimizer.step()
            if self.verbose_freq and (u + 1) % self.verbose_freq == 0:
                last = float(loss.detach())
                msg = "step %d/%d loss %.4f  %.3f s/update" % (u + 1, n_updates, last,
                                                               (time.perf_counter() - t_start) / (u + 1))
                (log or print)(msg)
        return (time.perf_counter() - t_start) / max(n_updates, 1)


class _Null:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


def bert_stage_model(config, depth: int = 4):
    """``models.bert<L>.depth=<N>.model(config, criterion)`` equivalent: stage modules wired by tensor names."""
    from ..models.bert import PretrainingCriterion, build_stages, extended_attention_mask

    class _First(torch.nn.Module):
        def __init__(self, st):
            super().__init__()
            self.st = st

        def forward(self, input_ids, token_type_ids, attention_mask):
            mask = extended_attention_mask(attention_mask)
            return self.st(input_ids, token_type_ids, mask), mask

    stages = build_stages(config, depth)
    model = [(_First(stages[0]), ["input0", "input1", "input2"], ["out0", "mask"])]
    prev = "out0"
    for i, st in enumerate(stages[1:-1]):
        model.append((st, [prev, "mask"], ["out%d" % (i + 1)]))
        prev = "out%d" % (i + 1)
    model.append((stages[-1], [prev, "mask"], ["scores", "nsp"]))
    model.append((PretrainingCriterion(config.vocab_size), ["scores", "nsp", "target_lm", "target_nsp"], ["loss"]))
    return model
