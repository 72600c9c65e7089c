"""Data-parallel parity check (run under torchrun, one process per GPU):
N-rank step on the global batch B  ==  1-device step on the same batch B  (SURVEY.md section 8e).

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 \
      tools/dp_check.py [B] [precision]

Every rank runs OUR step on its contiguous shard with gradient all-reduce + SyncBN; rank 0 then runs OUR
single-device step on the full batch (DP detached) and compares losses, every gradient of every phase (as
all-reduced and averaged), post-step weights and BatchNorm running statistics.  Exit code 1 on mismatch.
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from eadgan_b200 import functional as Fn, parallel  # noqa: E402
from eadgan_b200.steps.celeba import CelebAStep  # noqa: E402
from oracle.torch_oracle import sample_celeba, synth_celeba_images  # noqa: E402  (input generator only)

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
prec = sys.argv[2] if len(sys.argv) > 2 else "fp32"
graph_mode = len(sys.argv) > 3 and sys.argv[3] == "graph"
os.environ["EADGAN_PRECISION"] = prec
tol_g, tol_l = (2e-4, 1e-5) if prec == "fp32" else (1.5e-1, 5e-3)

dp = parallel.init_from_env()
rank, world = dist.get_rank(), dist.get_world_size()
dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", rank)))
imgs = synth_celeba_images(B, 0)
draws = sample_celeba(np.random.RandomState(0), B)
full = [imgs.to(dev), draws["z"].to(dev), draws["code"].to(dev), draws["labels"].to(dev)]
mine = [parallel.shard(t) for t in full]

step = CelebAStep(seed=0, device=dev)
parallel.attach(*step.optimizers())

if graph_mode:
    # the data-parallel step (bucketed NCCL all-reduce on the side stream + SyncBN all-reduces) captured in ONE
    # CUDA graph: 1 eager step + 2 replays on every rank  ==  3 eager single-device steps on the global batch
    from eadgan_b200.graph import GraphedStep
    imgs2 = synth_celeba_images(B, 1)
    draws2 = sample_celeba(np.random.RandomState(1), B)
    full2 = [imgs2.to(dev), draws2["z"].to(dev), draws2["code"].to(dev), draws2["labels"].to(dev)]
    mine2 = [parallel.shard(t) for t in full2]
    gs = GraphedStep(step, mine, warmup=1)
    outs = []
    for batch in (mine2, mine):
        o = gs(*batch)
        v = torch.stack([o["g_loss"], o["d_loss"], o["info_loss"]]).double()
        dist.all_reduce(v)
        outs.append(v / world)
    torch.cuda.synchronize()
    ok = True
    if rank == 0:
        Fn.set_allreduce(None, 1)
        ref = CelebAStep(seed=0, device=dev)
        ref(*full)
        tol = 5e-3 if prec == "fp32" else 3e-2   # Adam's lr*sign(g) noise after 1-2 updates, see above
        for i, batch in enumerate((full2, full)):
            o = ref(*batch)
            v = torch.stack([o["g_loss"], o["d_loss"], o["info_loss"]]).double()
            err = float(((outs[i] - v).abs() / v.abs().clamp_min(1.0)).max())
            print(f"[dp_check] graph replay {i}: dp {outs[i].tolist()} single {v.tolist()} err {err:.2e}")
            ok &= err <= tol
        werr = 0.0
        for (k, a), (_, b) in zip(step.G.state_dict().items(), ref.G.state_dict().items()):
            if a.is_floating_point() and "running" not in k:
                werr = max(werr, float((a - b).abs().max()))
        print(f"[dp_check] graph: max |w_dp - w_single| over G after 3 steps {werr:.2e} (lr 1e-3 / 2e-4)")
        ok &= werr <= 8e-3
        ok &= gs.kernels_per_replay > 100
        print("[dp_check] PASS" if ok else "[dp_check] FAIL")
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.broadcast(flag, 0)
    dist.barrier()
    code = 0 if int(flag.item()) else 1
    torch.cuda.synchronize()
    # The captured graph holds NCCL kernels of this communicator: tearing the process group down while the graph
    # is alive blocks in NCCL.  Leave without destroying the group (the process is about to exit anyway).
    sys.stdout.flush()
    sys.stderr.flush()
    os._exit(code)

rec = []
# record the REDUCED gradients: snapshot inside Adam.step via the dp.reduce() return value
red_log = []
orig_reduce = dp.reduce


def logging_reduce(opt):
    out = orig_reduce(opt)
    ps = [p for g in opt.param_groups for p in g["params"]]
    red_log.append([(out[p] / world).detach().clone() for p in ps])
    return out


dp.reduce = logging_reduce
losses = step(*mine, record=rec)
loss_vec = torch.stack([losses["g_loss"], losses["d_loss"], losses["info_loss"]]).double()
dist.all_reduce(loss_vec)
loss_vec /= world
torch.cuda.synchronize()

ok = True
if rank == 0:
    # single-device run of OUR implementation on the full batch
    Fn.set_allreduce(None, 1)
    ref = CelebAStep(seed=0, device=dev)
    rrec = []
    rl = ref(*full, record=rrec)
    rl = torch.stack([rl["g_loss"], rl["d_loss"], rl["info_loss"]]).double()
    lerr = float((loss_vec - rl).abs().max())
    print(f"[dp_check] world={world} B={B} {prec}: losses dp {loss_vec.tolist()} single {rl.tolist()} err {lerr:.2e}")
    ok &= lerr <= tol_l * 10
    for ph in range(3):
        worst, num, den2 = 0.0, 0.0, 0.0
        for a, b in zip(red_log[ph], rrec[ph]["grads"]):
            num += float((a - b).double().pow(2).sum())
            den2 += float(b.double().pow(2).sum())
            den = float(b.abs().max())
            if den < 1e-7:
                continue
            worst = max(worst, float((a - b).abs().max()) / den)
        l2 = (num / max(den2, 1e-300)) ** 0.5
        print(f"[dp_check] phase {ph}: worst tensor-normalised gradient error {worst:.2e}, L2-relative over all {l2:.2e}")
        # phase 0 starts from identical weights: tight.  Phases 1, 2 start from weights updated by Adam
        # (lr * sign(g) noise on near-zero gradients) and, in bf16, conv biases in front of a BatchNorm have a
        # mathematically zero gradient whose rounding-noise value is what "worst" then measures: bound the
        # L2-relative error over the whole gradient set there, and the worst tensor only loosely
        if ph == 0:
            ok &= worst <= tol_g
        else:
            ok &= l2 <= (2e-2 if prec == "fp32" else 5e-2) and worst <= max(tol_g, 5e-2) * 3
    rsd = ref.G.state_dict()
    for k, v in rsd.items():
        if "running" in k:
            # running_mean is compared on the scale of the channel standard deviations (the means themselves
            # are near zero, and the second G forward already runs on Adam-updated weights: sign(g) noise)
            scale = float(v.abs().max()) if "var" in k else float(rsd[k.replace("mean", "var")].sqrt().max())
            e = float((step.G.state_dict()[k] - v).abs().max()) / max(scale, 1e-12)
            ok &= e <= (1e-4 if prec == "fp32" else 2e-2)
            print(f"[dp_check] {k}: {e:.2e}")
    print("[dp_check] PASS" if ok else "[dp_check] FAIL")
flag = torch.tensor([1 if ok else 0], device=dev)
dist.broadcast(flag, 0)
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if int(flag.item()) else 1)
