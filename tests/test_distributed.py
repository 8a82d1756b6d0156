"""World-size-2 tests of the multi-GPU host logic on the gloo backend (CPU): the only collectives of
the system are the PPO gradient / advantage-statistics all-reduces; env shards never talk."""
import json
import os
import subprocess
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tests.helpers import ROOT


def _init(rank, world, port):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)


def _worker_grad(rank, world, port, out):
    from solorl_b200.agents.ppo import FlatGradAllReduce, broadcast_parameters, global_mean_std
    _init(rank, world, port)
    torch.manual_seed(100 + rank)
    net = torch.nn.Sequential(torch.nn.Linear(5, 7), torch.nn.Tanh(), torch.nn.Linear(7, 3))
    broadcast_parameters(net)                       # every rank starts from rank 0's weights
    w0 = torch.cat([p.detach().reshape(-1) for p in net.parameters()])
    x = torch.randn(11, 5)
    net(x).pow(2).sum().backward()
    local = torch.cat([p.grad.reshape(-1) for p in net.parameters()]).clone()
    FlatGradAllReduce(net.parameters())()
    synced = torch.cat([p.grad.reshape(-1) for p in net.parameters()]).clone()
    adv = torch.randn(13 + rank, 4) * (1 + rank) + rank
    m, s = global_mean_std(adv)
    torch.save({"w0": w0, "local": local, "synced": synced, "adv": adv, "mean": m, "std": s}, f"{out}/r{rank}.pt")
    dist.destroy_process_group()


def test_flat_gradient_allreduce_and_global_advantage_stats(tmp_path):
    port = 29500 + os.getpid() % 400
    mp.spawn(_worker_grad, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    r = [torch.load(tmp_path / f"r{i}.pt") for i in range(2)]
    assert torch.equal(r[0]["w0"], r[1]["w0"])                                   # broadcast
    mean_grad = (r[0]["local"] + r[1]["local"]) / 2
    for i in range(2):
        assert torch.allclose(r[i]["synced"], mean_grad, atol=1e-6)            # gradient = mean over ranks
    alladv = torch.cat([r[0]["adv"].reshape(-1), r[1]["adv"].reshape(-1)])
    for i in range(2):                                                           # ppo.py:35-37 over ALL samples
        assert abs(float(r[i]["mean"]) - float(alladv.mean())) < 1e-5
        assert abs(float(r[i]["std"]) - float(alladv.std())) < 1e-5               # unbiased (N-1), like torch.std


def _worker_ppo(rank, world, port, out):
    """Two ranks with different data take identical PPO steps (same weights after the update)."""
    from solorl_b200.agents.policy import Policy
    from solorl_b200.agents.ppo import PPO, broadcast_parameters
    from solorl_b200.agents.storage import OPBuffer
    from solorl_b200.envs import Box
    import numpy as np
    _init(rank, world, port)
    torch.manual_seed(7 + rank)
    T, N, D, A = 6, 8, 10, 4
    ac = Policy((D,), Box(-np.ones(A), np.ones(A)), None, {"hidden_size": 16})
    broadcast_parameters(ac)
    agent = PPO(ac, 0.1, 2, 16, 0.5, 0.01, lr=1e-3, max_grad_norm=0.5)
    buf = OPBuffer(T, N, (D,), A, "cpu")
    buf.obs.normal_(); buf.actions.normal_(); buf.rewards.normal_(); buf.value_preds.normal_()
    buf.returns.normal_(); buf.action_log_probs.normal_().mul_(0.1).sub_(5.0)
    v, a, e = agent.update(buf)
    w = torch.cat([p.detach().reshape(-1) for p in ac.parameters()])
    torch.save({"w": w, "losses": (v, a, e)}, f"{out}/p{rank}.pt")
    dist.destroy_process_group()


def test_ppo_update_keeps_ranks_in_lockstep(tmp_path):
    port = 29900 + os.getpid() % 400
    mp.spawn(_worker_ppo, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    r = [torch.load(tmp_path / f"p{i}.pt") for i in range(2)]
    assert torch.allclose(r[0]["w"], r[1]["w"], atol=1e-6)
    assert all(torch.isfinite(torch.tensor(x["losses"])).all() for x in r)


def test_reference_arm_under_torchrun_prints_one_line():
    """bench.py --impl reference at N=2: rank 0 alone runs and prints, the other rank exits 0."""
    env = dict(os.environ)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", str(30300 + os.getpid() % 400),
           os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "2", "--warmup", "1"]
    p = subprocess.run(cmd, capture_output=True, text=True, cwd=ROOT, env=env, timeout=600)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [l for l in p.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["n_gpus"] == 2 and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["e2e"]["h2d_bytes_per_step"] == 0
    # torchrun exports OMP_NUM_THREADS=1 to every rank: the CPU arm must still use all host threads
    assert d["cpu_baseline"]["cores"] == len(os.sched_getaffinity(0))
