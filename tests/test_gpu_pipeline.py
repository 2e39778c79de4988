"""End-to-end: synthetic env -> SyncStepRolloutGenerator (HBM-resident buffers) -> VecRollout (GAE)
-> PPO.learn (gather, fused loss, optimizer) for all five BASELINE configs at reduced sizes, with a
host env (numpy contract, PCIe every step) and a device env."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

CONFIGS = {
    # name: (env, n_envs, n_steps, batch, policy kwargs, ppo kwargs)
    "C1": ("CartPole-v1", 8, 32, 256, {}, dict(n_epochs=3, gamma=0.98, gae_lambda=0.8, learning_rate=1e-3)),
    "C2": ("BreakoutNoFrameskip-v4", 8, 16, 64, {}, dict(n_epochs=2, clip_range=0.1, vf_coef=0.5, ent_coef=0.01)),
    "C3": ("HalfCheetah-v4", 64, 16, 256, dict(pi_hidden_sizes=[256, 256], v_hidden_sizes=[256, 256], log_std_init=-2),
           dict(n_epochs=2, clip_range=0.1, ent_coef=4e-4, vf_coef=0.581, max_grad_norm=0.8)),
    "C4": ("Microrts-16x16", 6, 16, 32, {}, dict(n_epochs=2, clip_range=0.1, clip_range_vf=0.1, ppo2_vf_coef_halving=True,
                                                 ent_coef=0.01, vf_coef=0.5)),
    "C5": ("LuxAI_S2-64x64", 4, 4, 8, dict(num_additional_critics=12),
           dict(n_epochs=2, gamma=[1.0] * 13, gae_lambda=[0.95] * 13, clip_range=0.1, ent_coef=0.01,
                vf_coef=[0.5] + [0.1] * 12, multi_reward_weights=[0.9] + [0.1 / 12] * 12, gradient_accumulation=True,
                autocast_loss=True)),
}


@pytest.mark.parametrize("cfg", list(CONFIGS))
@pytest.mark.parametrize("where", ["device_env", "host_env"])
def test_learn_runs(cuda, cfg, where):
    from rl_algo_impls_b200.envs import make_synthetic_env
    from rl_algo_impls_b200.policy import ActorCritic
    from rl_algo_impls_b200.ppo import PPO
    from rl_algo_impls_b200.rollout import SyncStepRolloutGenerator

    name, n_envs, n_steps, batch, pkw, akw = CONFIGS[cfg]
    torch.manual_seed(0)
    env = make_synthetic_env(name, n_envs, seed=1, device=cuda if where == "device_env" else None, pool=3)
    policy = ActorCritic(env, subaction_mask=env.spec.subaction_mask, **pkw).to(cuda)
    gen = SyncStepRolloutGenerator(policy, env, n_steps=n_steps, subaction_mask=env.spec.subaction_mask)
    algo = PPO(policy, cuda, None, batch_size=batch, **akw)
    seen = []

    class CB:
        def on_step(self, timesteps_elapsed, train_stats):
            seen.append(train_stats)
            return True

    before = [p.detach().clone() for p in policy.parameters()]
    algo.learn(2 * n_envs * n_steps, gen, callbacks=[CB()])
    assert len(seen) == 2
    for s in seen:
        assert np.isfinite(s.loss) and np.isfinite(s.pi_loss) and np.isfinite(s.entropy_loss) and np.isfinite(s.grad_norm)
        assert np.all(np.isfinite(np.asarray(s.v_loss))) and 0 <= s.clipped_frac <= 1
    assert any((a != b.detach()).any().item() for a, b in zip(before, policy.parameters())), "parameters did not move"
    assert algo.launches_last_epoch > 0
    # the rollout buffers live in HBM and were filled
    assert gen.obs.is_cuda and gen.logprobs.abs().sum().item() > 0


def test_step_and_value_numpy_contract(cuda):
    """policy.step / policy.value keep the reference's numpy contract (actor_critic.py:298-318)."""
    from rl_algo_impls_b200.envs import make_synthetic_env
    from rl_algo_impls_b200.policy import ActorCritic

    env = make_synthetic_env("Microrts-16x16", 3, seed=2)
    policy = ActorCritic(env, subaction_mask=env.spec.subaction_mask).to(cuda)
    obs, _ = env.reset()
    step = policy.step(obs, action_masks=env.get_action_mask())
    assert step.a.shape == (3, 256, 7) and step.a.dtype == np.int64
    assert step.v.shape == (3,) and step.logp_a.shape == (3,) and np.all(step.logp_a <= 0)
    assert policy.value(obs).shape == (3,)
    # sampled actions are valid wherever the head has a valid entry
    mask = env.get_action_mask()
    start = 0
    for h, n in enumerate(env.spec.nvec):
        m = mask[..., start:start + n]
        has = m.any(-1)
        chosen = np.take_along_axis(m, step.a[..., h][..., None], axis=-1)[..., 0]
        assert chosen[has].all(), f"head {h} sampled a masked action"
        start += n


def test_rollout_store_step(cuda):
    """K0: one launch writes every field's step slice into row (*step % T) of its [T, ...] buffer."""
    from rl_algo_impls_b200 import ops

    T, N = 5, 3
    g = torch.Generator().manual_seed(0)
    step = [torch.randn(N, 74, 16, 16, generator=g), torch.rand(N, generator=g) < 0.5,
            torch.randint(0, 6, (N, 256, 7), generator=g).to(torch.uint8), torch.randn(N, 13, generator=g),
            torch.randint(0, 9, (N,), generator=g)]
    bufs = [torch.zeros((T,) + tuple(t.shape), dtype=t.dtype, device=cuda) for t in step]
    counter = torch.tensor([T * 7 + 3], dtype=torch.int64, device=cuda)  # row 3
    ops.rollout_store_step([t.to(cuda) for t in step], bufs, counter)
    for t, b in zip(step, bufs):
        assert torch.equal(b[3].cpu(), t)
        assert b[:3].cpu().count_nonzero() == 0 and b[4:].cpu().count_nonzero() == 0


def test_rollout_store_step_with_carry_over(cuda):
    """K0 with carry: the step slice goes to its buffer row and is then overwritten by the next step's slice in the
    same launch (aligned 16 KB-chunked fields, an odd-sized one, a misaligned carry source, and a field without)."""
    from rl_algo_impls_b200 import ops

    T, N = 4, 3
    g = torch.Generator().manual_seed(1)
    shapes = [((N, 74, 16, 16), torch.float32), ((N, 256, 78), torch.bool), ((N, 7, 3), torch.uint8), ((N,), torch.float32)]
    mk = lambda shape, dt: (torch.randn(shape, generator=g) if dt == torch.float32 else
                            (torch.rand(shape, generator=g) < 0.5 if dt == torch.bool else
                             torch.randint(0, 255, shape, generator=g).to(dt)))
    cur = [mk(*sd) for sd in shapes]
    nxt = [mk(*sd) for sd in shapes]
    dev_cur = [t.to(cuda) for t in cur]
    pad = torch.zeros(nxt[2].numel() + 1, dtype=torch.uint8, device=cuda)
    pad[1:] = nxt[2].reshape(-1).to(cuda)  # a carry source that is not 16-byte aligned
    carry = [nxt[0].to(cuda), nxt[1].to(cuda), pad[1:].view(nxt[2].shape), None]
    bufs = [torch.zeros((T,) + tuple(t.shape), dtype=t.dtype, device=cuda) for t in cur]
    counter = torch.tensor([2], dtype=torch.int64, device=cuda)
    ops.rollout_store_step(dev_cur, bufs, counter, carry=carry)
    for k, (t, b) in enumerate(zip(cur, bufs)):
        assert torch.equal(b[2].cpu(), t), k
        assert b[:2].cpu().count_nonzero() == 0 and b[3:].cpu().count_nonzero() == 0
        assert torch.equal(dev_cur[k].cpu(), nxt[k] if carry[k] is not None else t), k
    with pytest.raises(ValueError):
        ops.rollout_store_step(dev_cur, bufs, counter, carry=[None, None, None, torch.zeros(N + 1, device=cuda)])


@pytest.mark.parametrize("N,C,H,W,Cp", [(24, 74, 16, 16, 80), (3, 75, 64, 64, 80), (5, 3, 5, 7, 8), (2, 8, 4, 4, 8)])
def test_rollout_store_step_fused_bookkeeping(cuda, N, C, H, W, Cp):
    """K0, fused form: the packed observation goes to its buffer row and is replaced by the env's raw [N, C, H, W]
    observation transposed into the packed layout; next_episode_starts <- terminations | truncations; a plain carry and a
    field without; the step counter advances by one per launch (several launches in a row, and inside a captured graph:
    the ticket returns to zero every time)."""
    from rl_algo_impls_b200 import ops

    T = 4
    g = torch.Generator(device=cuda).manual_seed(N + C)
    packed = torch.zeros((N, H, W, Cp), device=cuda)
    packed[..., :C] = torch.randn((N, H, W, C), device=cuda, generator=g)
    starts = torch.rand(N, device=cuda, generator=g) < 0.5
    mask = torch.rand((N, H * W, 5), device=cuda, generator=g) < 0.3
    values = torch.randn(N, device=cuda, generator=g)
    bufs = [torch.zeros((T,) + tuple(t.shape), dtype=t.dtype, device=cuda) for t in (packed, starts, mask, values)]
    counter = torch.tensor([T + 1], dtype=torch.int64, device=cuda)
    ticket = torch.zeros(1, dtype=torch.int32, device=cuda)
    cur = [packed, starts, mask, values]
    raw, term, trunc, nmask = (torch.empty((N, C, H, W), device=cuda), torch.empty(N, dtype=torch.bool, device=cuda),
                               torch.empty(N, dtype=torch.bool, device=cuda), torch.empty_like(mask))

    def launch():
        ops.rollout_store_step(cur, bufs, counter, carry=[raw, term, nmask, None], carry_or=[None, trunc, None, None],
                               pack=(0, N, C, H * W, Cp), advance_ticket=ticket)

    graph = None
    for it in range(5):
        before = [t.clone() for t in cur]
        raw.copy_(torch.randn((N, C, H, W), device=cuda, generator=g))
        term.copy_(torch.rand(N, device=cuda, generator=g) < 0.3), trunc.copy_(torch.rand(N, device=cuda, generator=g) < 0.3)
        nmask.copy_(torch.rand(mask.shape, device=cuda, generator=g) < 0.3)
        row = int(counter.item()) % T
        if it < 2:
            launch()
        else:  # replays of one captured launch
            if graph is None:
                side = torch.cuda.Stream()
                side.wait_stream(torch.cuda.current_stream())
                keep = [t.clone() for t in cur] + [b.clone() for b in bufs] + [counter.clone()]
                with torch.cuda.stream(side):
                    launch()
                torch.cuda.current_stream().wait_stream(side)
                for t, k in zip(cur + bufs + [counter], keep):  # undo the warm-up launch
                    t.copy_(k)
                graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(graph):
                    launch()
                for t, k in zip(cur + bufs + [counter], keep):  # capture does not run the launch; be explicit anyway
                    t.copy_(k)
            graph.replay()
        torch.cuda.synchronize()
        for k in range(4):
            assert torch.equal(bufs[k][row], before[k]), (it, k)
        want = torch.zeros_like(packed)
        want[..., :C] = raw.permute(0, 2, 3, 1)
        assert torch.equal(packed, want), it
        assert torch.equal(starts, term | trunc) and torch.equal(mask, nmask) and torch.equal(values, before[3])
        assert int(counter.item()) == T + 1 + it + 1 and int(ticket.item()) == 0


def test_rollout_store_step_advance_only(cuda):
    """No field at all: the launch still advances the counter."""
    from rl_algo_impls_b200 import ops

    counter = torch.tensor([41], dtype=torch.int64, device=cuda)
    ticket = torch.zeros(1, dtype=torch.int32, device=cuda)
    ops.rollout_store_step([], [], counter, advance_ticket=ticket)
    assert int(counter.item()) == 42 and int(ticket.item()) == 0


@pytest.mark.parametrize("cfg", ["C1", "C3", "C4", "C5"])
def test_graph_replayed_rollout_equals_eager(cuda, cfg):
    """The CUDA-graph path (one replay per env step) fills the rollout buffer with exactly what the
    eager path does: same policy, same env pool, same RNG seed -> identical tensors."""
    from rl_algo_impls_b200.actor import rng
    from rl_algo_impls_b200.envs import make_synthetic_env
    from rl_algo_impls_b200.policy import ActorCritic
    from rl_algo_impls_b200.rollout import SyncStepRolloutGenerator

    name, n_envs, n_steps, _, pkw, _ = CONFIGS[cfg]
    torch.manual_seed(0)
    ref_env = make_synthetic_env(name, n_envs, seed=1, device=cuda, pool=3)
    policy = ActorCritic(ref_env, subaction_mask=ref_env.spec.subaction_mask, **pkw).to(cuda)
    results = []
    for graph in (False, True):
        rng.reseed(4242)  # the seed is baked into captured launches: set it before the capture
        env = make_synthetic_env(name, n_envs, seed=1, device=cuda, pool=3)
        gen = SyncStepRolloutGenerator(policy, env, n_steps=n_steps, subaction_mask=env.spec.subaction_mask,
                                       cuda_graph=graph)
        if graph:  # capture (which steps the env a few times), then put env + carry-over back to the start
            gen._rollout(output_next_values=False)
            env.reset()
            gen._set_next_obs(env.reset()[0])
            gen.next_episode_starts.fill_(True)
            if gen.next_action_masks is not None:
                gen._upload_masks(env.get_action_mask())
            gen._rollouts_done = 0
        torch.manual_seed(5)
        r = gen.rollout(gamma=0.99 if policy.value_shape == () else [0.99] * 13, gae_lambda=0.95 if policy.value_shape == () else [0.95] * 13)
        acts = r.actions if isinstance(r.actions, dict) else {"a": r.actions}
        results.append(dict(obs=r.obs.clone(), values=r.values.clone(), logp=r.logprobs.clone(), rewards=r.rewards.clone(),
                            starts=r.episode_starts.clone(), adv=r.advantages.clone(),
                            **{k: v.clone() for k, v in acts.items()}))
    eager, graphed = results
    for k in eager:
        if cfg == "C3" and k in ("a", "logp", "adv", "values"):
            continue  # Gaussian sampling draws torch.randn: the captured generator state differs from eager
        assert torch.equal(eager[k], graphed[k]), k


@pytest.mark.parametrize("cfg", ["C4", "C5"])
@pytest.mark.parametrize("graph", [False, True])
def test_host_env_is_handed_the_sampled_actions_as_int64(cuda, cfg, graph):
    """Host env + GridNet: env.step receives int64 arrays (the reference's contract, actor_critic.py:315-318) that equal
    the uint8 actions stored in the rollout buffer, step for step; the array of step s is still intact while step s + 1
    runs (two pinned landing zones); observations the env hands back reach the buffer in the packed layout."""
    from rl_algo_impls_b200.envs import make_synthetic_env
    from rl_algo_impls_b200.policy import ActorCritic
    from rl_algo_impls_b200.rollout import SyncStepRolloutGenerator

    name, n_envs, n_steps, _, pkw, _ = CONFIGS[cfg]
    torch.manual_seed(0)
    env = make_synthetic_env(name, n_envs, seed=1, device=None, pool=3)
    seen, held, returned = [], [], []
    inner = env.step

    def step(actions):
        acts = actions if isinstance(actions, dict) else {"per_position": actions}
        for v in acts.values():
            assert isinstance(v, np.ndarray) and v.dtype == np.int64
        if held:  # the previous step's arrays were not overwritten by this step's download
            for k, v in held[-1][0].items():
                assert np.array_equal(v, held[-1][1][k]), k
        held.append((acts, {k: v.copy() for k, v in acts.items()}))
        seen.append({k: v.copy() for k, v in acts.items()})
        out = inner(actions)
        returned.append(np.array(out[0], copy=True))
        return out

    env.step = step
    policy = ActorCritic(env, subaction_mask=env.spec.subaction_mask, **pkw).to(cuda)
    gen = SyncStepRolloutGenerator(policy, env, n_steps=n_steps, subaction_mask=env.spec.subaction_mask, cuda_graph=graph)
    assert gen._wide_actions is not None
    for _ in range(2):  # the second rollout runs on replays when `graph`
        seen.clear(), returned.clear()
        r = gen.rollout(gamma=0.99 if policy.value_shape == () else [0.99] * 13,
                        gae_lambda=0.95 if policy.value_shape == () else [0.95] * 13)
        torch.cuda.synchronize()
        stored = r.actions if isinstance(r.actions, dict) else {"per_position": r.actions}
        assert len(seen) == n_steps
        for s in range(n_steps):
            for k, v in seen[s].items():
                assert np.array_equal(stored[k][s].cpu().numpy().astype(np.int64).reshape(v.shape), v), (s, k)
        # packed observations: row s + 1 of the buffer is what the env returned at step s
        for s in (0, n_steps - 2):
            packed, raw = gen.obs[s + 1].cpu(), torch.from_numpy(returned[s])
            C = raw.shape[1]
            assert torch.equal(packed[..., :C], raw.permute(0, 2, 3, 1).float()) and packed[..., C:].abs().sum() == 0


def test_gridnet_sampler_draws_from_the_masked_softmax(cuda):
    """K5: Gumbel-max over Philox noise is an exact draw from softmax(logits | mask); the returned
    log-prob is the oracle's log-prob of the drawn action.  20,000 i.i.d. draws of one 2-head cell."""
    from oracle.distributions import Gridnet
    from rl_algo_impls_b200 import ops

    B, nvec = 20000, (6, 49)
    g = torch.Generator().manual_seed(0)
    row = torch.randn(1, 1, sum(nvec), generator=g) * 1.5
    mrow = torch.rand(1, 1, sum(nvec), generator=g) < 0.6
    mrow[0, 0, 0] = mrow[0, 0, 6] = True
    logits, mask = row.expand(B, 1, -1).contiguous(), mrow.expand(B, 1, -1).contiguous()
    spec = ops.GridnetSpec(nvec)
    actions, _, logp = ops.gridnet_sample(spec, logits.to(cuda), mask.to(cuda), None, 1234, 0, torch.uint8)
    a = actions.cpu().long()
    dist = Gridnet(1, nvec, logits, mask)
    want_logp = dist.log_prob(a)
    assert torch.allclose(logp.cpu(), want_logp, rtol=1e-5, atol=2e-6)
    start = 0
    for h, n in enumerate(nvec):
        m = mrow[0, 0, start:start + n]
        p = torch.softmax(torch.where(m, row[0, 0, start:start + n], torch.tensor(float("-inf"))), -1)
        freq = torch.bincount(a[:, 0, h], minlength=n).float() / B
        assert (freq[~m] == 0).all(), "a masked entry was drawn"
        sigma = torch.sqrt(p * (1 - p) / B)
        assert ((freq - p).abs() <= 4.5 * sigma + 1e-4).all(), (h, freq, p)
        start += n


@pytest.mark.parametrize("cfg", ["C1", "C4"])
@pytest.mark.parametrize("frozen", ["backbone", "policy_head", "value_head"])
def test_freeze_flags_keep_the_frozen_group_fixed(cuda, cfg, frozen):
    """freeze_policy_head / freeze_value_head / freeze_backbone (ppo.py:275-278,392-397): only the other groups move,
    and every parameter trains again afterwards (the policy is unfrozen at the end of learn_epoch)."""
    from rl_algo_impls_b200.envs import make_synthetic_env
    from rl_algo_impls_b200.policy import ActorCritic
    from rl_algo_impls_b200.ppo import PPO
    from rl_algo_impls_b200.rollout import SyncStepRolloutGenerator

    name, n_envs, n_steps, batch, pkw, akw = CONFIGS[cfg]
    torch.manual_seed(0)
    env = make_synthetic_env(name, n_envs, seed=1, device=cuda, pool=3)
    policy = ActorCritic(env, subaction_mask=env.spec.subaction_mask, **pkw).to(cuda)
    gen = SyncStepRolloutGenerator(policy, env, n_steps=n_steps, subaction_mask=env.spec.subaction_mask)
    flags = {f"freeze_{k}": k == frozen for k in ("backbone", "policy_head", "value_head")}
    algo = PPO(policy, cuda, None, batch_size=batch, **akw, **flags)
    policy.freeze(flags["freeze_policy_head"], flags["freeze_value_head"], freeze_backbone=flags["freeze_backbone"])
    fixed = {n for n, p in policy.named_parameters() if not p.requires_grad}
    policy.unfreeze()
    assert fixed and len(fixed) < len(list(policy.parameters()))
    before = {n: p.detach().clone() for n, p in policy.named_parameters()}
    algo.learn_epoch(0, 1 << 30, gen, None)
    moved = {n for n, p in policy.named_parameters() if (p.detach() != before[n]).any().item()}
    assert moved and not (moved & fixed), f"frozen parameters moved: {sorted(moved & fixed)}"
    assert all(p.requires_grad for p in policy.parameters())


@pytest.mark.parametrize("cfg", ["C4", "C5"])
def test_packed_observation_layout_changes_no_number(cuda, cfg):
    """The rollout buffer stores observations in the trunk's layout (channels last, planes padded to a multiple of 8:
    ActorCritic.packed_obs_shape) so that the gather hands over trunk-ready rows.  Same seeds with the packing switched
    off (the trunk then packs every batch itself): identical rollout tensors, identical parameters after learn_epoch."""
    from rl_algo_impls_b200.actor import rng
    from rl_algo_impls_b200.envs import make_synthetic_env
    from rl_algo_impls_b200.policy import ActorCritic
    from rl_algo_impls_b200.ppo import PPO
    from rl_algo_impls_b200.rollout import SyncStepRolloutGenerator

    name, n_envs, n_steps, batch, pkw, akw = CONFIGS[cfg]
    finals = []
    for packed in (True, False):
        torch.manual_seed(0)
        rng.reseed(99)
        env = make_synthetic_env(name, n_envs, seed=1, device=cuda, pool=3)
        policy = ActorCritic(env, subaction_mask=env.spec.subaction_mask, **pkw).to(cuda)
        assert policy.packed_obs_shape is not None
        C, H, W = env.single_observation_space.shape
        if not packed:
            policy.network.packed_obs_shape = lambda: None
        gen = SyncStepRolloutGenerator(policy, env, n_steps=n_steps, subaction_mask=env.spec.subaction_mask)
        algo = PPO(policy, cuda, None, batch_size=batch, **akw)
        algo.learn_epoch(0, 1 << 30, gen, None)
        if packed:
            assert tuple(gen.obs.shape[2:]) == (H, W, (C + 7) // 8 * 8) and (gen.obs[..., C:] == 0).all()
            obs_nchw = gen.obs[..., :C].permute(0, 1, 4, 2, 3).contiguous()
        else:
            assert tuple(gen.obs.shape[2:]) == (C, H, W)
            obs_nchw = gen.obs.float()
        finals.append((obs_nchw, gen.logprobs.clone(), gen.values.clone(), [p.detach().clone() for p in policy.parameters()],
                       algo.last_train_stats))
    a, b = finals
    assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1]) and torch.equal(a[2], b[2])
    # the update itself runs cuDNN weight-gradient kernels whose summation order is not fixed from run to run:
    # the two runs agree to rounding, not to the bit
    for pa, pb in zip(a[3], b[3]):
        assert torch.allclose(pa, pb, rtol=1e-4, atol=1e-6)
    assert abs(a[4].loss - b[4].loss) <= 1e-4 * max(abs(b[4].loss), 1e-2)
    assert abs(a[4].grad_norm - b[4].grad_norm) <= 1e-3 * abs(b[4].grad_norm)
