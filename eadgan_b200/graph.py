"""Whole-step CUDA graph: the training step (all phases, backward passes, Adam updates) is captured once
and replayed, so the ~600 kernel launches per step cost no host time (the
reference issues ~10^4 ATen dispatches per step, section 3.2).

    step = CelebAStep(...)
    gstep = GraphedStep(step, example_inputs)      # eager warm-up steps, then capture
    losses = gstep(imgs, z, code, labels)          # copies into the static inputs, replays

What makes the step capturable: no host synchronisation anywhere in it (the affine glue has a closed-form
inverse, eadgan_b200/affine.py); every kernel is launched on the caller's current stream; tensor maps are
kernel *arguments* (baked into the graph, valid because the graph's private memory pool keeps every
address); Adam's step count lives on the device while capturing (eadgan_adam_step_dev).
"""
from __future__ import annotations

import torch

from . import _lib, tc


class GraphedStep:
    def __init__(self, step, example_inputs, warmup=3):
        self.step = step
        self.static_in = [t.clone() for t in example_inputs]
        self.opts = list(step.optimizers())
        cur = torch.cuda.current_stream()
        side = torch.cuda.Stream()
        side.wait_stream(cur)
        with torch.cuda.stream(side):        # warm-up on a side stream, as torch's capture protocol asks
            for _ in range(max(1, warmup)):
                step(*self.static_in)
        cur.wait_stream(side)
        torch.cuda.synchronize()
        for o in self.opts:
            o.prepare_capture()
        self.graph = torch.cuda.CUDAGraph()
        k0 = _lib.lib().eadgan_kernel_launches()
        # packed-weight cache: a hit during capture would leave the pack kernel OUT of the graph (stale weights on
        # replay), and an entry filled during capture holds nothing until the first replay -> clear on both sides
        tc.invalidate_caches()
        with torch.cuda.graph(self.graph):
            self.static_out = step(*self.static_in)
        tc.invalidate_caches()
        self.kernels_per_replay = int(_lib.lib().eadgan_kernel_launches() - k0)
        self._first = True

    def __call__(self, *inputs):
        for dst, src in zip(self.static_in, inputs):
            if dst.data_ptr() != src.data_ptr():
                dst.copy_(src, non_blocking=True)
        if self._first:
            self._first = False          # the capture pass already advanced the python step mirrors once
        else:
            for o in self.opts:
                o.on_replay()
        self.graph.replay()
        _lib.bump_weights_epoch()        # the replay ran Adam steps: eager code must re-pack afterwards
        return self.static_out
