"""CUDA-graph capture of one whole training step.

The reference's step loop (``src/networks.py:71-99``) issues, per minibatch, a few hundred tiny
kernels for the small models it is used on (README toy model, UCI MLPs): on a B200 those steps
are bound by launch latency, not by bytes.  ``GraphedTrainStep`` records forward + ELBO +
backward + optimizer step once into a ``torch.cuda.CUDAGraph`` (all of this repo's C-ABI calls are
plain stream-ordered launches with caller-owned buffers, so they capture as they are) and
replays it per minibatch: one launch per step.

Requirements (both checked by ``_check_optimizer``): the optimizer must be graph-capturable
(``capturable=True``, or ``whvi_b200.optim.FlatAdam``); if a learning-rate scheduler is to act on
a captured step (``train_model(cuda_graph=True)`` always steps one), the optimizer's ``lr`` must be
a CUDA tensor (PyTorch schedulers then update it in place) -- a float ``lr`` raises.
"""
from __future__ import annotations

import torch


def _check_optimizer(optimizer, scheduled: bool = False) -> None:
    """``scheduled``: a learning-rate scheduler will step between replays.  A Python-float ``lr`` is
    baked into the captured kernels' arguments, so the schedule would be silently ignored: the
    ``lr`` of every group must then be a CUDA tensor (schedulers update it in place)."""
    for group in optimizer.param_groups:
        if not group.get("capturable", False) and not getattr(optimizer, "whvi_graph_safe", False):
            raise RuntimeError("CUDA-graph training needs a capturable optimizer, e.g. "
                               "torch.optim.Adam(params, lr=torch.tensor(1e-3, device='cuda'), capturable=True)")
        if scheduled and not (torch.is_tensor(group["lr"]) and group["lr"].is_cuda):
            raise RuntimeError("CUDA-graph training with a learning-rate scheduler needs lr to be a CUDA tensor "
                               "(lr=torch.tensor(1e-3, device='cuda')): a float lr is frozen into the captured graph "
                               "and scheduler.step() would have no effect")


class GraphedTrainStep:
    """``step = GraphedTrainStep(model, optimizer, x, y, n); loss = step(x, y)``.

    ``x``/``y`` given at construction fix the minibatch shape; every call copies the new
    minibatch into the captured input buffers and replays the graph.  The returned loss (and
    ``model.current_kl`` / ``model.current_mnll``) are the captured output tensors, overwritten
    by the next replay.  ``zero_grad(set_to_none=True)`` semantics: gradients live in the
    graph's private pool and are recomputed (not accumulated) by every replay.
    """

    def __init__(self, model, optimizer, x, y, n: int, ignore_kl: bool = False, warmup: int = 3, scheduled: bool = False):
        if not (x.is_cuda and y.is_cuda):
            raise RuntimeError("GraphedTrainStep needs CUDA minibatches")
        _check_optimizer(optimizer, scheduled)
        self.model, self.optimizer = model, optimizer
        self.x, self.y = x.detach().clone(), y.detach().clone()
        # the warm-up steps below must not count as training: remember parameters and optimizer
        # state and put them back (in place -- the graph captures these very tensors)
        params = [p for group in optimizer.param_groups for p in group["params"]]
        saved_params = [p.detach().clone() for p in params]
        saved_state = {p: {k: (v.detach().clone() if torch.is_tensor(v) else v) for k, v in st.items()}
                       for p, st in optimizer.state.items()}
        # autograd graphs of earlier eager steps (kept alive by these attributes) would pin the
        # parameters' AccumulateGrad nodes to the stream they ran on -- the legacy default stream
        # cannot take part in a capture
        model.current_kl = model.current_mnll = 0.0
        side = torch.cuda.Stream(device=x.device)
        side.wait_stream(torch.cuda.current_stream(x.device))
        with torch.cuda.stream(side):  # warm-up off the capture stream: lazy inits, allocator, smem opt-ins
            for _ in range(max(warmup, 1)):
                optimizer.zero_grad(set_to_none=True)
                model.loss(self.x, self.y, n=n, ignore_kl=ignore_kl).backward()
                optimizer.step()
        torch.cuda.current_stream(x.device).wait_stream(side)
        with torch.no_grad():
            for p, saved in zip(params, saved_params):
                p.copy_(saved)
            for p, st in optimizer.state.items():  # lazily created state: zeros == never stepped
                old = saved_state.get(p, {})
                for k, v in st.items():
                    if torch.is_tensor(v):
                        v.copy_(old[k]) if k in old else v.zero_()
        optimizer.zero_grad(set_to_none=True)
        model.current_kl = model.current_mnll = 0.0
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.loss = model.loss(self.x, self.y, n=n, ignore_kl=ignore_kl)
            self.loss.backward()
            optimizer.step()
            if getattr(optimizer, "whvi_graph_safe", False):
                optimizer.zero_grad()  # FlatAdam: gradients persist as views of one buffer; the memset is part of the step
        self.kl, self.mnll = model.current_kl, model.current_mnll

    def __call__(self, x, y):
        self.x.copy_(x, non_blocking=True)
        self.y.copy_(y, non_blocking=True)
        self.graph.replay()
        self.model.current_kl, self.model.current_mnll = self.kl, self.mnll
        return self.loss


class GraphedStepCache:
    """One captured step per minibatch shape (a data loader's last batch is usually shorter)."""

    def __init__(self, model, optimizer, n: int, ignore_kl: bool = False, scheduled: bool = False):
        _check_optimizer(optimizer, scheduled)  # fail before the first minibatch, not inside the loop
        self.model, self.optimizer, self.n, self.ignore_kl, self.scheduled = model, optimizer, n, ignore_kl, scheduled
        self.steps: dict[tuple, GraphedTrainStep] = {}

    def __call__(self, x, y):
        key = (tuple(x.shape), tuple(y.shape), x.dtype, y.dtype)
        step = self.steps.get(key)
        if step is None:
            step = self.steps[key] = GraphedTrainStep(self.model, self.optimizer, x, y, self.n, self.ignore_kl,
                                                      scheduled=self.scheduled)
        return step(x, y)
