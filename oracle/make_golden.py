"""TEST INFRASTRUCTURE ONLY.  Generates tests/golden/*.json by executing the UNMODIFIED reference
scripts (oracle/ref_runner.py) in this container:  python -m oracle.make_golden

The reference ships no tests or golden vectors, so these fixtures -- the
reference's own outputs on seeded synthetic inputs with seeded random-init weights -- are what
pins oracle/torch_oracle.py (and, through it, the CUDA path).  Each fixture stores the losses
and a fingerprint (sum, L2 norm, abs-max, 8 probes) of every gradient tensor seen by every
``optimizer.step()`` and of every parameter after it.
"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ref_runner as R  # noqa: E402
from oracle import torch_oracle as O  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")


def fingerprint_log(log):
    out = []
    for e in log:
        out.append({"lr": e["lr"],
                    "grads": [None if g is None else O.summarize(g) for g in e["grads"]],
                    "params_after": [O.summarize(p) for p in e["params_after"]]})
    return out


def make_celeba(B=4, seed=0):
    imgs = O.synth_celeba_images(B, seed)
    ns, log = R.run_script("celeba", [(imgs, torch.zeros(B, dtype=torch.long))], argv=["--batch_size", str(B)],
                           seed=seed)
    return {"config": "celeba", "batch": B, "seed": seed, "torch": torch.__version__,
            "source": "celebA/EAD-GAN_celebA.py executed by oracle/ref_runner.py",
            "losses": {"g_loss": ns["g_loss"].item(), "d_loss": ns["d_loss"].item(),
                       "info_loss": ns["info_loss"].item()},
            "phases": fingerprint_log(log)}


def make_dsprites(B=6, seed=0):
    imgs = O.synth_dsprites_images(B, seed)
    ns, log = R.run_script("dsprites", [imgs], argv=["--batch_size", str(B)], seed=seed,
                           artefacts={"encoder_pxy_50000.pt": O.dsprites_pxy_state(seed)})
    names = {"d_loss": "d_loss", "g_loss": "g_loss", "cat_loss": "cat_loss", "cont_loss": "cont_loss",
             "affine_loss": "affine_loss", "relative_cat_loss": "relative_cat_loss", "total": "info_affine_color_loss"}
    return {"config": "dsprites", "batch": B, "seed": seed, "torch": torch.__version__,
            "source": "dSprites/rp.py executed by oracle/ref_runner.py (encoder_pxy_50000.pt = seeded random-init stand-in)",
            "losses": {k: ns[v].item() for k, v in names.items()},
            "phases": fingerprint_log(log)}


def make_colored(B=6, seed=0):
    imgs = O.synth_dsprites_images(B, seed)
    ns, log = R.run_script("colored", [imgs], argv=["--batch_size", str(B)], seed=seed,
                           artefacts={"encoder_pxy_color_50000.pt": O.dsprites_pxy_state(seed, colored=True)})
    names = {"d_loss": "d_loss", "g_loss": "g_loss", "cat_loss": "cat_loss", "cont_loss": "cont_loss",
             "affine_loss": "affine_color_loss", "relative_cat_loss": "relative_cat_loss",
             "total": "info_affine_color_loss"}
    return {"config": "colored", "batch": B, "seed": seed, "torch": torch.__version__,
            "source": "colored_dSprites/rp_color.py executed by oracle/ref_runner.py "
                      "(encoder_pxy_color_50000.pt = seeded random-init stand-in)",
            "losses": {k: ns[v].item() for k, v in names.items()},
            "phases": fingerprint_log(log)}


def make_mnist(B=8, seed=0):
    imgs = O.synth_mnist_images(B, seed)
    ns, log = R.run_script("mnist", [(imgs, torch.zeros(B, dtype=torch.long))], argv=["--batch_size", str(B)],
                           seed=seed, artefacts={"rpqmnxy_approximator.pt": O.mnist_approximator_state(seed)})
    return {"config": "mnist", "batch": B, "seed": seed, "torch": torch.__version__,
            "source": "MNIST/EAD-GAN_rpqmnxy.py executed by oracle/ref_runner.py "
                      "(rpqmnxy_approximator.pt = seeded random-init stand-in)",
            "losses": {"g_loss": ns["g_loss"].item(), "d_loss": ns["d_loss"].item(),
                       "info_loss": ns["info_loss"].item()},
            "phases": fingerprint_log(log)}


def make_pxy(B=8, seed=0, colored=False):
    imgs = O.synth_dsprites_images(B, seed)
    ns, log = R.run_script("pxy_color" if colored else "pxy", [imgs], argv=["--batch_size", str(B)], seed=seed)
    return {"config": "pxy_color" if colored else "pxy", "batch": B, "seed": seed, "torch": torch.__version__,
            "source": ("colored_dSprites/pxy_color.py" if colored else "dSprites/pxy.py") + " executed by oracle/ref_runner.py",
            "losses": {"affine_loss": ns["affine_loss"].item()},
            "phases": fingerprint_log(log)}


def make_approximator(n_iter=3, seed=0):
    ns, log = R.run_approximator(n_iter, seed=seed)
    return {"config": "approximator", "batch": 128, "iterations": n_iter, "seed": seed, "torch": torch.__version__,
            "source": "MNIST/approximate_rpqmnxy.py executed by oracle/ref_runner.run_approximator (iteration count "
                      "20001 -> %d)" % n_iter,
            "losses": {"affine_loss": float(ns["affine_loss"])},        # the loss of the LAST iteration
            "phases": fingerprint_log(log)}


def main():
    os.makedirs(GOLDEN, exist_ok=True)
    torch.set_num_threads(8)
    for name, fn in (("celeba_b4_seed0", lambda: make_celeba(4, 0)), ("celeba_b6_seed3", lambda: make_celeba(6, 3)),
                     ("dsprites_b6_seed0", lambda: make_dsprites(6, 0)), ("dsprites_b8_seed2", lambda: make_dsprites(8, 2)),
                     ("colored_b6_seed0", lambda: make_colored(6, 0)), ("colored_b8_seed1", lambda: make_colored(8, 1)),
                     ("mnist_b8_seed0", lambda: make_mnist(8, 0)), ("mnist_b64_seed1", lambda: make_mnist(64, 1)),
                     ("pxy_b8_seed0", lambda: make_pxy(8, 0)), ("pxy_color_b8_seed0", lambda: make_pxy(8, 0, True)),
                     ("approximator_it3_seed0", lambda: make_approximator(3, 0))):
        if sys.argv[1:] and not any(name.startswith(a) for a in sys.argv[1:]):
            continue
        g = fn()
        with open(os.path.join(GOLDEN, name + ".json"), "w") as f:
            json.dump(g, f, indent=1)
        print(name, g["losses"])


if __name__ == "__main__":
    main()
