"""Fused multi-tensor Adam with torch.optim.Adam's constructor, param_groups, state and
``zero_grad`` / ``step`` semantics (celebA/EAD-GAN_celebA.py:211-217; op order of
torch/optim/adam.py::_single_tensor_adam).

Under data parallelism ``step()`` is also where this optimiser's gradients -- exactly
the parameter set the current phase owns -- are all-reduced (eadgan_b200.parallel).
"""
from __future__ import annotations

import ctypes as C
import math

import torch

from . import _lib as L
from ._lib import call, stream


class Adam(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0, amsgrad=False, **kw):
        if weight_decay != 0 or amsgrad:
            raise RuntimeError("eadgan_b200.Adam: weight_decay / amsgrad are not used by the reference and unsupported")
        for k in ("maximize", "capturable", "differentiable"):
            if kw.get(k):
                raise RuntimeError(f"eadgan_b200.Adam: {k}=True is unsupported")
        if not 0.0 <= lr or not 0.0 <= eps or not 0.0 <= betas[0] < 1.0 or not 0.0 <= betas[1] < 1.0:
            raise ValueError("Invalid Adam hyper-parameter")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=0, amsgrad=False))
        # data parallelism: when a DP state exists (parallel.init / init_from_env ran first, as under
        # ``torchrun -m eadgan_b200.run script.py``), every Adam all-reduces the gradients of its own parameter set
        # in step(); parallel.attach() does the same explicitly for optimisers built before init
        from . import parallel
        self._dp = parallel.get()
        self._step_dev = None       # device int64 step counter (CUDA-graph capture, eadgan_b200.graph)
        self._captured = None       # parameters stepped by the captured step() (their python "step" mirrors)

    # ---- CUDA-graph support (eadgan_b200.graph.GraphedStep) -----------------------------------------
    def prepare_capture(self):
        """call OUTSIDE capture, after at least one eager step: loads the device step counter with the
        current step count.  Inside capture, step() advances and reads it on the device."""
        steps = {st["step"] for st in self.state.values() if "step" in st}
        if len(steps) != 1:
            raise RuntimeError("eadgan_b200.Adam: graph capture needs one common step count over the parameters "
                               f"that receive gradients (found {sorted(steps)}); run an eager step first")
        dev = next(iter(self.state)).device
        if self._step_dev is None:
            self._step_dev = torch.zeros(1, device=dev, dtype=torch.int64)
        self._step_dev.fill_(steps.pop())
        self._captured = None

    def on_replay(self):
        """python-side mirror of the device counter: one more step has been (re)played"""
        for p in self._captured or ():
            self.state[p]["step"] += 1

    def zero_grad(self, set_to_none: bool = True):
        super().zero_grad(set_to_none=set_to_none)
        if self._dp is not None:
            self._dp.arm(self)

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        reduced = self._dp.reduce(self) if self._dp is not None else None  # {param: reduced grad}
        gscale = 1.0 if self._dp is None else 1.0 / self._dp.world_size
        capturing = torch.cuda.is_available() and torch.cuda.is_current_stream_capturing()
        serial = L.next_step_serial()
        if capturing:
            if self._step_dev is None:
                raise RuntimeError("eadgan_b200.Adam: call prepare_capture() before capturing step() in a CUDA graph")
            call("eadgan_adam_advance", C.c_void_p(self._step_dev.data_ptr()), stream())
            self._captured = []
        for group in self.param_groups:
            beta1, beta2 = group["betas"]
            ps, gs, ms, vs = [], [], [], []
            t = None
            for p in group["params"]:
                if p.grad is None:
                    continue
                g = p.grad if reduced is None else reduced[p]
                if not p.is_cuda:
                    raise RuntimeError("eadgan_b200.Adam: parameters must live on a CUDA device (no CPU fallback)")
                if g.dtype != torch.float32 or p.dtype != torch.float32:
                    raise RuntimeError("eadgan_b200.Adam: fp32 parameters and gradients only")
                st = self.state[p]
                if not st:
                    st["step"] = 0
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                st["step"] += 1
                if capturing:
                    self._captured.append(p)
                if t is None:
                    t = st["step"]
                elif t != st["step"]:
                    if capturing:
                        raise RuntimeError("eadgan_b200.Adam: parameters with different step counts cannot be captured")
                    self._launch(ps, gs, ms, vs, group, t, gscale)
                    ps, gs, ms, vs = [], [], [], []
                    t = st["step"]
                p._eadgan_stepped = serial   # its cached packs and prefetched spectral-norm results are stale now
                ps.append(p)
                gs.append(g if g.is_contiguous() else g.contiguous())
                ms.append(st["exp_avg"])
                vs.append(st["exp_avg_sq"])
            if ps:
                self._launch(ps, gs, ms, vs, group, t, gscale, self._step_dev if capturing else None)
        if not capturing and self._step_dev is not None:
            # an eager step between graph replays: the device counter the captured step() reads must advance too
            call("eadgan_adam_advance", C.c_void_p(self._step_dev.data_ptr()), stream())
        return loss

    @staticmethod
    def _launch(ps, gs, ms, vs, group, t, gscale, step_dev=None):
        beta1, beta2 = group["betas"]
        bc1 = 1.0 - beta1 ** t
        bc2 = 1.0 - beta2 ** t
        step_size = group["lr"] / bc1
        bc2_sqrt = math.sqrt(bc2)
        st = stream()
        for i0 in range(0, len(ps), L.ADAM_MAX_TENSORS):
            A = L.AdamTensors()
            chunk = range(i0, min(len(ps), i0 + L.ADAM_MAX_TENSORS))
            for j, i in enumerate(chunk):
                if not ps[i].is_contiguous():
                    raise RuntimeError("eadgan_b200.Adam: non-contiguous parameter")
                A.p[j] = ps[i].data_ptr()
                A.g[j] = gs[i].data_ptr()
                A.m[j] = ms[i].data_ptr()
                A.v[j] = vs[i].data_ptr()
                A.numel[j] = ps[i].numel()
            A.count = len(chunk)
            if step_dev is not None:
                call("eadgan_adam_step_dev", C.byref(A), float(beta1), float(beta2), float(group["eps"]),
                     float(group["lr"]), C.c_void_p(step_dev.data_ptr()), float(gscale), st)
            else:
                call("eadgan_adam_step", C.byref(A), float(beta1), float(beta2), float(group["eps"]),
                     float(step_size), float(bc2_sqrt), float(gscale), st)
