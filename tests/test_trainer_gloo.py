"""CPU / gloo world_size-2 test of the N>1 host logic of omni3d_b200.train.FlatSGDTrainer (flat arena,
rank-0 broadcast, loss + gradient all-reduce, skip-together stabiliser, unused-parameter handling, LR
schedule).  The two CUDA kernels the trainer calls (fused SGD, finite scan) are replaced by their torch
definitions for this CPU test only — the kernels themselves are covered by tests/test_kernels_gpu.py."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _patch_kernels():
    from omni3d_b200 import kernels as Kx

    def sgd(p, g, mom, lr, momentum, wd, grad_scale=1.0, skip_flag=None):
        if skip_flag is not None and int(skip_flag) != 0:
            return
        d = g * grad_scale + wd * p
        mom.mul_(momentum).add_(d)
        p.sub_(lr * mom)

    def finite(g, flag):
        if not torch.isfinite(g).all():
            flag.fill_(1)
    Kx.sgd_momentum, Kx.grad_finite = sgd, finite


class Toy(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.a = torch.nn.Linear(8, 8)
        self.bn = torch.nn.BatchNorm1d(8)
        self.level3 = torch.nn.Module()
        self.level3.project = torch.nn.Linear(4, 4)        # never used -> "unused" arena region
        self.boom = False

    def forward(self, x):
        from omni3d_b200.train import LOSS_KEYS
        y = self.bn(self.a(x)).pow(2).mean()
        out = {k: y * (i + 1) / 55.0 for i, k in enumerate(LOSS_KEYS)}
        if self.boom:
            out[LOSS_KEYS[0]] = out[LOSS_KEYS[0]] * float("nan")
        return out


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    _patch_kernels()
    from omni3d_b200 import cubercnn as pc
    from omni3d_b200.train import FlatSGDTrainer, lr_at
    cfg = pc.get_cfg_defaults(pc.get_cfg())
    cfg.SOLVER.BASE_LR, cfg.SOLVER.WARMUP_ITERS, cfg.SOLVER.STEPS = 0.1, 2, (4,)
    cfg.MODEL.STABILIZE = 0.02
    torch.manual_seed(100 + rank)                      # ranks start with DIFFERENT weights
    model = Toy()
    tr = FlatSGDTrainer(cfg, model)
    p0 = tr.flat_p.clone()
    gathered = [torch.zeros_like(p0) for _ in range(world)]
    dist.all_gather(gathered, p0)
    assert all(torch.equal(gathered[0], g) for g in gathered), "rank-0 broadcast failed"
    torch.manual_seed(7 + rank)
    x = torch.randn(16, 8)
    tr.step(x)
    st = tr.status()
    # identical parameters on every rank after an averaged-gradient step, and they moved
    gathered = [torch.zeros_like(p0) for _ in range(world)]
    dist.all_gather(gathered, tr.flat_p.clone())
    assert all(torch.equal(gathered[0], g) for g in gathered)
    assert not torch.equal(tr.flat_p, p0)
    u0, u1 = tr.bounds["unused"]
    assert u1 > u0 and torch.equal(tr.flat_p[u0:u1], p0[u0:u1]), "unused parameters must not be updated"
    assert st["iterations_success"] == 1 and st["iterations_explode"] == 0
    # one rank produces a NaN loss -> ALL ranks skip together, parameters unchanged everywhere
    before = tr.flat_p.clone()
    model.boom = rank == 1
    tr.step(x)
    model.boom = False
    st = tr.status()
    assert torch.equal(tr.flat_p, before), "diverging step must be skipped on every rank"
    assert st["iterations_explode"] == 1
    # schedule: linear warm-up then step decay
    assert abs(lr_at(cfg, 0) - 0.1 * 0.001) < 1e-12 and abs(lr_at(cfg, 2) - 0.1) < 1e-12 and abs(lr_at(cfg, 4) - 0.01) < 1e-12
    q.put((rank, float(st["total_loss"] if st["total_loss"] == st["total_loss"] else -1.0)))
    dist.destroy_process_group()


def test_trainer_world2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=180)
        assert p.exitcode == 0
    got = sorted(q.get(timeout=5) for _ in range(2))
    assert [g[0] for g in got] == [0, 1]


def test_single_process_matches_torch_sgd():
    """flat-arena SGD == torch.optim.SGD with the reference's param-group rules (solver/build.py:6-69)."""
    _patch_kernels()
    from omni3d_b200 import cubercnn as pc
    from omni3d_b200.train import FlatSGDTrainer
    cfg = pc.get_cfg_defaults(pc.get_cfg())
    cfg.SOLVER.BASE_LR, cfg.SOLVER.WARMUP_ITERS = 0.05, 0
    torch.manual_seed(0)
    m1 = Toy(); m2 = Toy(); m2.load_state_dict(m1.state_dict())
    tr = FlatSGDTrainer(cfg, m1)
    groups = [{"params": [m2.a.weight], "weight_decay": 1e-4}, {"params": [m2.a.bias], "weight_decay": 1e-4},
              {"params": [m2.bn.weight, m2.bn.bias], "weight_decay": 0.0}]
    opt = torch.optim.SGD(groups, lr=0.05, momentum=0.9)
    x = torch.randn(16, 8)
    for _ in range(3):
        tr.step(x)
        opt.zero_grad(); sum(m2(x).values()).backward(); opt.step()
    for (n, a), (_, b) in zip(m1.named_parameters(), m2.named_parameters()):
        assert torch.allclose(a, b, atol=1e-6), n
