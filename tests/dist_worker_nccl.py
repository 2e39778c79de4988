"""Worker of tests/test_gpu_dist_nccl.py: launched as ``python -m torch.distributed.run --nproc-per-node R`` (one process
per GPU, NCCL).  Checks SURVEY.md section 8e: the R-rank data-parallel PPO update -- envs sharded, advantage moments and
gradients all-reduced -- equals the single-process update on the concatenation of the ranks' minibatches.

Every rank holds a C4-shaped learner (GridNet MicroRTS 16x16, the C4 trunk) with identical initial weights and its own
env slice of a synthetic rollout; the minibatch index lists are prescribed so that the single-process run can be fed
exactly the union of the ranks' minibatches.  Three runs per rank:
  A  data-parallel, eager update              B  data-parallel, CUDA-graph update (three captured segments with the
  C  single process (data_parallel=False) on     all-reduces between the replays)
     the concatenated rollout, batch R x B
Bars: all-reduced moments == moments of the concatenated minibatch (1e-14: f64 sums in a different order); gradient of
the first minibatch 1e-5 of each tensor's largest entry; final parameters (A vs C, and B vs A: cuDNN picks other
algorithms under capture, so the two data-parallel runs agree to rounding, not to the bit) RMS error <= 1 % of the RMS
update -- Adam's sign-like step amplifies rounding on near-zero gradients, while a missing or misplaced all-reduce
would show as O(100 %) -- and the last epoch's losses within 1e-3.
"""
import copy
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("nccl", device_id=dev)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False

    from rl_algo_impls_b200 import ops, spaces
    from rl_algo_impls_b200.configs import CONFIGS
    from rl_algo_impls_b200.policy import ActorCritic
    from rl_algo_impls_b200.ppo import PPO
    from rl_algo_impls_b200.rollout import VecRollout
    from tests.synth import MICRORTS_GATES, MICRORTS_NVEC, gae_inputs, gridnet_inputs

    T, N, HW, B, n_epochs = 8, 4, 256, 16, 2
    nvec, A = MICRORTS_NVEC, len(MICRORTS_NVEC)

    class Env:
        num_envs = N
        single_observation_space = spaces.Box(0, 1, (74, 16, 16), np.float32)
        action_plane_space = spaces.MultiDiscrete(nvec)
        single_action_space = spaces.MultiDiscrete(np.tile(np.asarray(nvec), HW))

    torch.manual_seed(0)
    policy0 = ActorCritic(Env(), **CONFIGS["C4"].policy).to(dev)

    def shard(r):
        rng = np.random.default_rng(500 + r)
        ro = gae_inputs(600 + r, T, N, 1, 0.1)
        g = gridnet_inputs(700 + r, T * N, HW, nvec, 0, 0.08)
        ro["obs"] = (rng.random((T, N, 74, 16, 16)) < 0.1).astype(np.float32)
        ro["actions"], ro["masks"] = g["actions"].reshape(T, N, HW, A).astype(np.uint8), g["mask"].reshape(T, N, HW, -1)
        with torch.no_grad():
            flat = lambda a: torch.as_tensor(a.reshape((-1,) + a.shape[2:])).to(dev)
            lp = policy0(flat(ro["obs"]), flat(ro["actions"]), flat(ro["masks"])).logp_a.cpu().numpy()
        ro["logprobs"] = (lp.reshape(T, N) + rng.standard_normal((T, N)).astype(np.float32) * 0.05).astype(np.float32)
        # minibatch index lists of this rank: n_epochs permutations of its T * N local rows
        gen = torch.Generator().manual_seed(800 + r)
        ro["idx"] = [torch.randperm(T * N, generator=gen) for _ in range(n_epochs)]
        return ro

    shards = [shard(r) for r in range(world)]  # every rank can build every shard (seeded): no need to gather them
    mine = shards[rank]

    class FixedIndexRollout(VecRollout):
        fixed = None

        def minibatch_indices(self, batch_size, shuffle=True):
            order = self.fixed.pop(0).to(self.device)
            return [order[i:i + batch_size] for i in range(0, order.numel(), batch_size)]

    class Gen:
        def __init__(self, ro, orders, n_envs):
            self.ro, self.orders, self.n_steps = ro, orders, T
            self.vec_env = type("E", (), {"num_envs": n_envs})()

        def rollout(self, gamma, gae_lambda):
            ro = self.ro
            r = FixedIndexRollout(dev, ro["next_episode_starts"], ro["next_values"], ro["obs"], ro["actions"], ro["rewards"],
                                  ro["episode_starts"], ro["values"], ro["logprobs"], ro["masks"], gamma, gae_lambda,
                                  subaction_mask=MICRORTS_GATES)
            r.fixed = [o.clone() for o in self.orders]
            return r

    hp = dict(batch_size=B, n_epochs=n_epochs, learning_rate=1e-3, clip_range=0.1, clip_range_vf=0.1, vf_coef=0.5,
              ent_coef=0.01, ppo2_vf_coef_halving=True, max_grad_norm=0.5, gamma=0.99, gae_lambda=0.95)

    def run(graphed, data_parallel, ro, orders, n_envs, batch):
        policy = copy.deepcopy(policy0)
        algo = PPO(policy, dev, None, **{**hp, "batch_size": batch})
        algo.cuda_graph_update, algo.data_parallel = graphed, data_parallel
        first = {}
        orig = algo._clip_and_step

        def spy(flat, w):
            if not first:
                first.update({n: (p.grad.detach().clone() / w) for n, p in policy.network.named_parameters()})
            return orig(flat, w)

        algo._clip_and_step = spy
        algo.learn_epoch(0, 1 << 30, Gen(ro, orders, n_envs), None)
        torch.cuda.synchronize()
        algo.graphed_used = bool(algo._update_graphs)
        return policy, first, algo

    pol_a, grads_a, algo_a = run(False, True, mine, mine["idx"], N, B)
    pol_b, _, algo_b = run(True, True, mine, mine["idx"], N, B)

    # ---- C: the concatenated problem, single process --------------------------------------------------------------
    cat = {}
    for k in ("rewards", "values", "episode_starts", "obs", "actions", "masks", "logprobs"):
        cat[k] = np.concatenate([s[k] for s in shards], axis=1)  # along the env axis: [T, R * N, ...]
    cat["next_episode_starts"] = np.concatenate([s["next_episode_starts"] for s in shards])
    cat["next_values"] = np.concatenate([s["next_values"] for s in shards])
    NT = world * N
    orders = []
    for e in range(n_epochs):
        per_rank = []
        for r, s in enumerate(shards):
            local = s["idx"][e]
            t, n = local // N, local % N
            per_rank.append((t * NT + r * N + n).reshape(-1, B))  # [minibatches, B] global rows of rank r
        orders.append(torch.cat(per_rank, dim=1).reshape(-1))      # minibatch k = union of the ranks' minibatches k
    pol_c, grads_c, algo_c = run(False, False, cat, orders, NT, world * B)

    report = {"rank": rank, "world": world}
    ok = True

    # moments: all-reduced local moments vs the moments of the concatenated minibatch
    h = algo_a._hyper(1, 1, 1.0, 1.0)
    adv_local = torch.randn(B, generator=torch.Generator().manual_seed(40 + rank)).to(dev)
    gathered = [torch.empty_like(adv_local) for _ in range(world)]
    dist.all_gather(gathered, adv_local)
    m_dp = algo_a._moments(adv_local, h)
    m_cat = ops.adv_moments(torch.cat(gathered).reshape(-1, 1), None, h.adv_mode, h.adv_weights)
    rel = ((m_dp - m_cat).abs() / m_cat.abs().clamp_min(1e-300)).max().item()
    report["moments_rel"] = rel
    ok &= rel <= 1e-14

    worst = 0.0
    for k, g in grads_c.items():
        e = ((grads_a[k].double() - g.double()).abs().max() / g.double().abs().max().clamp_min(1e-300)).item()
        worst = max(worst, e)
    report["first_gradient_rel"] = worst
    ok &= worst <= 1e-5

    def param_check(pa, pb, bar):
        worst = 0.0
        for (k, va), vb, v0 in zip(pa.network.state_dict().items(), pb.network.state_dict().values(),
                                   policy0.network.state_dict().values()):
            upd = (vb.double() - v0.double()).pow(2).mean().sqrt().item()
            err = (va.double() - vb.double()).pow(2).mean().sqrt().item()
            worst = max(worst, err / (upd + 1e-12))
        return worst, worst <= bar

    report["params_dp_vs_single"], good = param_check(pol_a, pol_c, 1e-2)
    ok &= good
    report["params_graphed_vs_eager"], good = param_check(pol_b, pol_a, 1e-2)
    ok &= good
    report["graphed_path_used"] = algo_b.graphed_used and not algo_a.graphed_used
    ok &= report["graphed_path_used"]
    sa, sb, sc = algo_a.last_train_stats, algo_b.last_train_stats, algo_c.last_train_stats
    # the data-parallel stats are this rank's local means; their average over the ranks is the global minibatch's
    def mean_over_ranks(x):
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t)
        return (t / world).item()
    worst = 0.0
    for k in ("loss", "pi_loss", "entropy_loss", "approx_kl"):
        for s_dp in (sa, sb):
            got, want = mean_over_ranks(getattr(s_dp, k)), getattr(sc, k)
            worst = max(worst, abs(got - want) / max(abs(want), 1e-2))
    report["stats_rel"] = worst
    ok &= worst <= 1e-3
    # replicas stay identical across ranks
    flat = torch.cat([p.detach().reshape(-1) for p in pol_a.parameters()])
    ref = flat.clone()
    dist.broadcast(ref, 0)
    report["replicas_identical"] = bool(torch.equal(flat, ref))
    ok &= report["replicas_identical"]
    report["ok"] = bool(ok)
    print("DIST_REPORT " + json.dumps(report), flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
