"""The N > 1 host logic on the CPU: two gloo ranks.  Envs shard across ranks and only two things
cross ranks -- the [2V + 1] advantage moments (sum, sum of squares, count) and the flattened
gradients -- so that the data-parallel update equals the single-process update on the concatenated
minibatch."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle.ppo_loss import normalize_advantages


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank: int, world: int, port: int, out_dir: str):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from rl_algo_impls_b200.envs import make_synthetic_env
        from rl_algo_impls_b200.policy import ActorCritic
        from rl_algo_impls_b200.ppo import PPO

        torch.manual_seed(0)  # replicas start from identical weights
        env = make_synthetic_env("CartPole-v1", 4, seed=100 + rank, pool=2)  # each rank owns its env slice
        policy = ActorCritic(env)
        algo = PPO(policy, torch.device("cpu"), None)
        # rank-specific local gradients: rank r contributes (r + 1) * ones
        params = list(policy.parameters())
        for p in params:
            p.grad = torch.full_like(p, float(rank + 1))
        algo._sync_grads(params)
        mean_grad = sum(range(1, world + 1)) / world
        ok_grads = all(torch.allclose(p.grad, torch.full_like(p, mean_grad)) for p in params)

        # advantage moments: local (sum, sumsq, count) all-reduced == moments of the concatenation
        g = torch.Generator().manual_seed(7)
        full = torch.randn(64, 3, generator=g, dtype=torch.float64) * 2 + 0.5
        local = full[rank::world]
        moments = torch.cat([local.sum(0), (local * local).sum(0), torch.tensor([float(local.shape[0])], dtype=torch.float64)])
        dist.all_reduce(moments)
        n = moments[-1]
        mean = moments[:3] / n
        std = torch.sqrt((moments[3:6] - moments[:3] * mean) / (n - 1))
        want = normalize_advantages(full.float())
        got = ((local.float() - mean.float()) / (std.float() + 1e-8))
        ok_norm = torch.allclose(got, want[rank::world], rtol=1e-5, atol=1e-6)

        obs0 = env.reset()[0]
        np.save(os.path.join(out_dir, f"obs{rank}.npy"), obs0)
        with open(os.path.join(out_dir, f"ok{rank}"), "w") as f:
            f.write(f"{int(ok_grads)}{int(ok_norm)}")
    finally:
        dist.destroy_process_group()


def _dp_problem():
    """A CartPole-sized PPO minibatch and a fresh MLP policy, identical in every process."""
    import torch.nn as nn

    g = torch.Generator().manual_seed(11)
    B = 64
    batch = dict(obs=torch.randn(B, 4, generator=g), actions=torch.randint(0, 2, (B,), generator=g),
                 old_logp=-torch.rand(B, generator=g), adv=torch.randn(B, generator=g) * 3 + 1,
                 old_values=torch.randn(B, generator=g), returns=torch.randn(B, generator=g))
    torch.manual_seed(5)
    net = nn.Sequential(nn.Linear(4, 16), nn.Tanh(), nn.Linear(16, 3))  # 2 logits + 1 value
    return batch, net


def _dp_loss(net, batch, rows, adv_normalised):
    """The reference's PPO loss (oracle/ppo_loss.py) on the rows `rows` of the minibatch."""
    from oracle.distributions import MaskedLogits
    from oracle.ppo_loss import ppo_loss

    out = net(batch["obs"][rows])
    dist_ = MaskedLogits(out[:, :2], None)
    parts = ppo_loss(dist_.log_prob(batch["actions"][rows]), dist_.entropy(), out[:, 2], batch["old_logp"][rows],
                     adv_normalised, batch["old_values"][rows], batch["returns"][rows], clip_range=0.2,
                     clip_range_vf=None, ent_coef=0.01, vf_coef=torch.tensor(0.5))
    return parts.loss


def _dp_worker(rank: int, world: int, port: int, out_dir: str):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from rl_algo_impls_b200.envs import make_synthetic_env
        from rl_algo_impls_b200.policy import ActorCritic
        from rl_algo_impls_b200.ppo import PPO

        batch, net = _dp_problem()
        rows = torch.arange(rank, 64, world)  # this rank's env slice of the global minibatch
        # global-minibatch advantage statistics from the all-reduced (sum, sum of squares, count)
        local = batch["adv"][rows].double()
        moments = torch.stack([local.sum(), (local * local).sum(), torch.tensor(float(len(rows)), dtype=torch.float64)])
        dist.all_reduce(moments)
        n = moments[2]
        mean = moments[0] / n
        std = torch.sqrt((moments[1] - moments[0] * mean) / (n - 1))
        adv = ((batch["adv"][rows] - mean.float()) / (std.float() + 1e-8))
        _dp_loss(net, batch, rows, adv).backward()
        algo = PPO(ActorCritic(make_synthetic_env("CartPole-v1", 2, seed=0, pool=1)), torch.device("cpu"), None)
        params = list(net.parameters())
        algo._sync_grads(params)  # the learner's own reduction: all-reduce of the flattened gradients, / R
        torch.save([p.grad.clone() for p in params], os.path.join(out_dir, f"grads{rank}.pt"))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_two_rank_update_equals_the_single_process_update_on_the_concatenated_minibatch(tmp_path):
    """local losses are means over B / R samples with GLOBAL advantage statistics, so the
    all-reduced gradient / R is the gradient of the reference's loss on the whole minibatch."""
    world = 2
    mp.spawn(_dp_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    batch, net = _dp_problem()
    rows = torch.arange(64)
    _dp_loss(net, batch, rows, normalize_advantages(batch["adv"])).backward()
    want = [p.grad for p in net.parameters()]
    for r in range(world):
        got = torch.load(tmp_path / f"grads{r}.pt")
        for g, w in zip(got, want):
            torch.testing.assert_close(g, w, rtol=2e-5, atol=1e-7)


@pytest.mark.timeout(180)
def test_two_rank_gradient_and_moment_reduction(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        assert open(tmp_path / f"ok{r}").read() == "11", f"rank {r} failed"
    # env slices differ across ranks (different seeds): no rank re-simulates another's envs
    assert not np.array_equal(np.load(tmp_path / "obs0.npy"), np.load(tmp_path / "obs1.npy"))


def test_bench_reference_arm_non_zero_ranks_exit_quietly():
    """bench.py --impl reference under torchrun: rank 0 alone runs and prints; the others exit 0."""
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, RANK="1", LOCAL_RANK="1", WORLD_SIZE="2")
    res = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--gpus", "2"],
                         capture_output=True, text=True, env=env, timeout=120)
    assert res.returncode == 0 and res.stdout.strip() == ""
