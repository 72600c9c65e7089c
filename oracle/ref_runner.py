"""TEST INFRASTRUCTURE ONLY -- never imported by the product package.

Runs the UNMODIFIED reference training scripts from /root/reference on CPU for a
fixed number of iterations on synthetic batches, and records what every
``optimizer.step()`` saw (gradients before, parameters after) plus the losses left
in the script's namespace.  This is how the golden vectors under tests/golden/ are
produced (see oracle/make_golden.py) and how oracle/torch_oracle.py (the stand-alone
restatement that travels to the GPU box) is pinned to the real reference.

How the script is executed (reference cites, relative to /root/reference):
  * the whole script source is parsed with ``ast``; only the statements that build
    the dataset / DataLoader are dropped (celebA/EAD-GAN_celebA.py:194-206,
    MNIST/EAD-GAN_rpqmnxy.py:234-246, dSprites/rp.py:236-246,
    colored_dSprites/rp_color.py:234-244) and a list of synthetic batches named
    ``dataloader`` is injected instead;
  * ``sample_image`` (PNG dumps, e.g. celebA/EAD-GAN_celebA.py:233-285) is replaced
    by a no-op and ``torch.save`` is stubbed: both are I/O side effects outside the
    hot path;
  * on a CPU-only host ``Tensor.cuda`` / ``Module.cuda`` are identity (the utils_*
    files call ``.cuda()`` unconditionally, e.g. celebA/utils_rpqxy.py:30,78);
  * everything else -- model classes, losses, optimisers, the loop body, the
    ``utils_*`` affine glue -- is the reference's own code, executed as is.

Nothing here runs on the GPU box (``/root/reference`` does not exist there).
"""
from __future__ import annotations

import ast
import contextlib
import os
import sys
import tempfile

import numpy as np
import torch

REF_ROOT = os.environ.get("EADGAN_REF", "/root/reference")

SCRIPTS = {
    "celeba": ("celebA", "EAD-GAN_celebA.py"),
    "mnist": ("MNIST", "EAD-GAN_rpqmnxy.py"),
    "dsprites": ("dSprites", "rp.py"),
    "colored": ("colored_dSprites", "rp_color.py"),
    "pxy": ("dSprites", "pxy.py"),
    "pxy_color": ("colored_dSprites", "pxy_color.py"),
}

_DATA_TARGETS = {"dataset", "dataloader", "dataset_zip", "x_train", "x_train_tensor", "transform"}


def available() -> bool:
    return os.path.isdir(REF_ROOT)


def _strip_data_statements(tree: ast.Module) -> ast.Module:
    body = []
    for node in tree.body:
        if isinstance(node, ast.Assign):
            names = {t.id for t in node.targets if isinstance(t, ast.Name)}
            if names & _DATA_TARGETS:
                continue
        if isinstance(node, ast.FunctionDef) and node.name == "sample_image":
            continue
        # os.makedirs("data/mnist") etc. are harmless inside the scratch cwd
        body.append(node)
    tree.body = body
    return tree


@contextlib.contextmanager
def _patched_torch(log):
    """identity .cuda(), no torch.save, record every Adam.step()."""
    saved = (torch.Tensor.cuda, torch.nn.Module.cuda, torch.save, torch.optim.Adam.step,
             torch.load)
    if not torch.cuda.is_available():
        torch.Tensor.cuda = lambda self, *a, **k: self
        torch.nn.Module.cuda = lambda self, *a, **k: self
    torch.save = lambda *a, **k: None
    real_step = saved[3]
    real_load = saved[4]

    def load(path, *a, **k):
        k.setdefault("weights_only", False)
        return real_load(path, *a, **k)

    def step(self, closure=None):
        params = [p for g in self.param_groups for p in g["params"]]
        entry = {
            "opt_id": id(self),
            "lr": self.param_groups[0]["lr"],
            "grads": [None if p.grad is None else p.grad.detach().clone() for p in params],
        }
        out = real_step(self, closure)
        entry["params_after"] = [p.detach().clone() for p in params]
        log.append(entry)
        return out

    torch.optim.Adam.step = step
    torch.load = load
    try:
        yield
    finally:
        (torch.Tensor.cuda, torch.nn.Module.cuda, torch.save, torch.optim.Adam.step,
         torch.load) = saved


def extract_defs(name: str, argv=()):
    """AST-extract only class/function definitions + ``opt`` of a reference script
    (SURVEY.md section 7.1).  Returns the exec namespace."""
    sub, fname = SCRIPTS[name]
    path = os.path.join(REF_ROOT, sub, fname)
    src = open(path).read()
    tree = ast.parse(src)
    keep = []
    for node in tree.body:
        if isinstance(node, (ast.Import, ast.ImportFrom, ast.ClassDef, ast.FunctionDef)):
            if isinstance(node, ast.ImportFrom) and node.module and node.module.startswith("utils_"):
                continue  # utils import handled by run_script only (needs artefacts)
            keep.append(node)
        elif isinstance(node, ast.Assign):
            names = {t.id for t in node.targets if isinstance(t, ast.Name)}
            if names & {"parser", "opt", "cuda"}:
                keep.append(node)
        elif isinstance(node, ast.Expr) and isinstance(node.value, ast.Call):
            f = node.value.func
            if isinstance(f, ast.Attribute) and f.attr == "add_argument":
                keep.append(node)
    tree.body = keep
    ns = {"__name__": "ref_defs_" + name}
    old_argv = sys.argv
    sys.argv = [fname, *argv]
    try:
        with contextlib.redirect_stdout(open(os.devnull, "w")):
            exec(compile(tree, path, "exec"), ns)
    finally:
        sys.argv = old_argv
    ns.setdefault("FloatTensor", torch.FloatTensor)
    ns.setdefault("LongTensor", torch.LongTensor)
    return ns


def run_script(name: str, batches, argv=(), seed=0, artefacts=None, pre_exec=None):
    """Execute reference script ``name`` for ``len(batches)`` iterations.

    batches  : list of what the script's DataLoader would yield
               (celebA/MNIST: (imgs, labels) tuples; dSprites: uint8 [B,64,64] tensors)
    artefacts: {filename: state_dict or callable(ns)->state_dict} written into the
               scratch cwd before the script starts (encoder_pxy_50000.pt, ...)
    Returns (namespace, step_log).
    """
    sub, fname = SCRIPTS[name]
    sdir = os.path.join(REF_ROOT, sub)
    path = os.path.join(sdir, fname)
    tree = _strip_data_statements(ast.parse(open(path).read()))
    code = compile(tree, path, "exec")
    log = []
    ns = {"__name__": "__ref_main__", "dataloader": list(batches),
          "sample_image": lambda *a, **k: None}
    old_argv, old_cwd, old_path = sys.argv, os.getcwd(), list(sys.path)
    for m in [m for m in sys.modules if m.startswith("utils_")]:
        del sys.modules[m]
    tmp = tempfile.mkdtemp(prefix="eadgan_ref_")
    try:
        os.chdir(tmp)
        sys.argv = [fname, "--n_epochs", "1", *argv]
        sys.path.insert(0, sdir)
        # artefacts are written with the *real* torch.save, before it is stubbed
        for fn, sd in (artefacts or {}).items():
            torch.save(sd, os.path.join(tmp, fn))
        with _patched_torch(log), contextlib.redirect_stdout(open(os.devnull, "w")):
            torch.manual_seed(seed)
            np.random.seed(seed)
            if pre_exec is not None:
                pre_exec(ns)
            exec(code, ns)
    finally:
        os.chdir(old_cwd)
        sys.argv = old_argv
        sys.path[:] = old_path
        for m in [m for m in sys.modules if m.startswith("utils_")]:
            del sys.modules[m]
    return ns, log


def run_approximator(n_iter: int, seed=0):
    """Execute MNIST/approximate_rpqmnxy.py (the script that pre-trains the affine approximator MLP, :111-153) for
    ``n_iter`` iterations: the only change is the hard-coded iteration count ``range(20001)`` (:118).  Returns
    (namespace, step_log) like run_script."""
    path = os.path.join(REF_ROOT, "MNIST", "approximate_rpqmnxy.py")
    tree = ast.parse(open(path).read())
    hits = 0
    for node in ast.walk(tree):
        if isinstance(node, ast.Constant) and node.value == 20001:
            node.value = int(n_iter)
            hits += 1
    assert hits == 1, "approximate_rpqmnxy.py: expected exactly one iteration-count literal"
    code = compile(tree, path, "exec")
    log = []
    ns = {"__name__": "__main__"}
    old_cwd = os.getcwd()
    tmp = tempfile.mkdtemp(prefix="eadgan_ref_")
    try:
        os.chdir(tmp)
        with _patched_torch(log), contextlib.redirect_stdout(open(os.devnull, "w")):
            torch.manual_seed(seed)
            np.random.seed(seed)
            exec(code, ns)
    finally:
        os.chdir(old_cwd)
    return ns, log
