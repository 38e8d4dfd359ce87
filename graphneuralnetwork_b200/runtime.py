"""CUDA-graph runners for the launch-bound full-graph configs.

The reference's GCN / GAT / HAN training loops (`GCN/train_eval.py`, `GAT/train_eval.py`,
`HAN/train_eval.py`) run one full-graph step per epoch on the SAME tensors: features, adjacency
and labels never change between epochs.  On the Cora- and ACM-shaped graphs such a step is a few
dozen kernels of a few microseconds each, so the epoch time is launch overhead, not bandwidth.
`CapturedTrainStep` records one whole step — forward through the drop-in layers (the
libgnn_b200.so kernels and the torch matmuls alike), loss, backward, optimizer update — into a
CUDA graph once and replays it per epoch: one launch per epoch.

    step = CapturedTrainStep(model, lambda: F.cross_entropy(model(X, adj)[idx], y[idx]), optimizer)
    for epoch in range(200):
        loss = step()          # device scalar; .item() only when you want to look at it

Rules (checked where they can be): the closure must read only tensors that stay alive and are
updated in place; the optimizer must be built with `capturable=True` when it keeps device-side
step counters (Adam / AdamW); graph structures are planned during the warm-up steps so the
captured region holds no host synchronisation.
"""
from __future__ import annotations

from typing import Callable, Optional

import torch

from . import _lib


class CapturedTrainStep:
    def __init__(self, model: torch.nn.Module, loss_closure: Callable[[], torch.Tensor],
                 optimizer: Optional[torch.optim.Optimizer] = None, warmup: int = 3):
        if not torch.cuda.is_available():
            raise _lib.GnnError("CapturedTrainStep needs a CUDA device (the hot path has no CPU fallback)")
        if optimizer is not None:
            for group in optimizer.param_groups:
                if "capturable" in group and not group["capturable"]:
                    raise _lib.GnnError("build the optimizer with capturable=True: its step counter must live on the "
                                        "device to be replayed inside a CUDA graph")
        self.model, self.optimizer, self._closure = model, optimizer, loss_closure
        self.stream = torch.cuda.Stream()
        self.graph = torch.cuda.CUDAGraph()
        self.kernel_launches_per_replay = 0
        self.stream.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(self.stream):
            # warm-up on the side stream: allocator pools, cuBLAS workspaces, CSR conversion,
            # transposes and row plans of every adjacency (all cached on the CSRGraph afterwards)
            for _ in range(max(warmup, 1)):
                self._one_step()
            self.stream.synchronize()
            before = _lib.launch_count()
            self._zero_grad()
            with torch.cuda.graph(self.graph, stream=self.stream):
                self.loss = self._one_step(zero=False)
            self.kernel_launches_per_replay = _lib.launch_count() - before
        torch.cuda.current_stream().wait_stream(self.stream)

    def _zero_grad(self):
        if self.optimizer is not None:
            self.optimizer.zero_grad(set_to_none=True)
        else:
            for p in self.model.parameters():
                p.grad = None

    def _one_step(self, zero: bool = True) -> torch.Tensor:
        if zero:
            self._zero_grad()
        loss = self._closure()
        loss.backward()
        if self.optimizer is not None:
            self.optimizer.step()
        return loss.detach()

    def __call__(self) -> torch.Tensor:
        """Replay one step on the caller's current stream order; returns the (static) loss tensor."""
        self.stream.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(self.stream):
            self.graph.replay()
        torch.cuda.current_stream().wait_stream(self.stream)
        return self.loss


class CapturedForward:
    """Inference twin of `CapturedTrainStep`: `fn()` (a no-grad forward over static tensors)
    captured once, replayed per call; returns the static output tensor."""

    def __init__(self, fn: Callable[[], torch.Tensor], warmup: int = 2):
        if not torch.cuda.is_available():
            raise _lib.GnnError("CapturedForward needs a CUDA device (the hot path has no CPU fallback)")
        self.stream = torch.cuda.Stream()
        self.graph = torch.cuda.CUDAGraph()
        self.stream.wait_stream(torch.cuda.current_stream())
        with torch.no_grad(), torch.cuda.stream(self.stream):
            for _ in range(max(warmup, 1)):
                fn()
            self.stream.synchronize()
            before = _lib.launch_count()
            with torch.cuda.graph(self.graph, stream=self.stream):
                self.out = fn()
            self.kernel_launches_per_replay = _lib.launch_count() - before
        torch.cuda.current_stream().wait_stream(self.stream)

    def __call__(self) -> torch.Tensor:
        self.stream.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(self.stream):
            self.graph.replay()
        torch.cuda.current_stream().wait_stream(self.stream)
        return self.out
