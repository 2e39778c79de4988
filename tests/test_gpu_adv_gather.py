"""K2 / K3 parity: advantage moments + normalisation (ppo.py:307-318) and the minibatch row gather
(rollout.py:56-69), plus the VecRollout index stream (vec_rollout.py:166-175; bit-exact)."""
import numpy as np
import pytest
import torch

from oracle.ppo_loss import normalize_advantages
from oracle.rollout import minibatch_index_stream
from tests.parity import close

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("M,V,B", [(1024, 1, 256), (12288, 1, 3072), (4096, 13, 128), (500, 3, 500), (64, 2, 2)])
@pytest.mark.parametrize("mode", ["normalize", "standardize", "after_scaling", "none"])
def test_adv_normalize(cuda, M, V, B, mode):
    from rl_algo_impls_b200 import ops

    g = torch.Generator().manual_seed(M + V)
    adv = (torch.randn(M, V, generator=g) * 3 + 1) if V > 1 else (torch.randn(M, generator=g) * 3 + 1)
    idx = torch.randperm(M, generator=g)[:B]
    w = torch.linspace(0.2, 1.0, V) if V > 1 else None
    kw = dict(normalize_advantage=mode == "normalize", standardize_advantage=mode == "standardize",
              normalize_advantages_after_scaling=mode == "after_scaling")
    want = normalize_advantages(adv[idx], multi_reward_weights=w, **kw)
    m = {"normalize": ops.ADV_NORMALIZE, "standardize": ops.ADV_STANDARDIZE, "after_scaling": ops.ADV_AFTER_SCALING,
         "none": ops.ADV_NONE}[mode]
    adv_d = adv.to(cuda).reshape(M, V)
    got = ops.adv_normalize(adv_d, idx.to(cuda), m, w.tolist() if w is not None else None)
    close(got, want.reshape(got.shape), rtol=2e-5, what="normalised advantages")


def test_gather_rows_bit_exact(cuda):
    from rl_algo_impls_b200 import ops

    g = torch.Generator().manual_seed(0)
    M, B = 700, 300
    srcs = [
        torch.randint(0, 255, (M, 4, 84, 84), dtype=torch.uint8, generator=g),  # Atari obs (28,224 B rows)
        torch.randn(M, generator=g),  # logprobs
        torch.randint(0, 6, (M, 256, 7), generator=g),  # int64 per-cell actions
        torch.rand(M, 256, 78, generator=g) < 0.5,  # bool masks (19,968 B rows)
        torch.randn(M, 13, generator=g),  # multi-head values (52 B rows)
        torch.randn(M, 17, generator=g).double(),  # f64 obs
        torch.randint(0, 255, (M, 3, 5), dtype=torch.uint8, generator=g),  # 15 B rows: unaligned scalar path
        torch.randn(M, 75, 9, generator=g),  # 2,700 B rows: 4-byte aligned only
    ]
    idx = torch.randperm(M, generator=g)[:B]
    outs = ops.gather_rows([s.to(cuda) for s in srcs], idx.to(cuda))
    for s, o in zip(srcs, outs):
        assert o.dtype == s.dtype and o.shape == (B,) + s.shape[1:]
        assert torch.equal(o.cpu(), s[idx])


def test_vec_rollout_index_stream_and_batches(cuda):
    """Same seed -> the same randperm stream as the reference's VecRollout.minibatches, and every
    minibatch field equals the flat rollout indexed by it (short last minibatch kept)."""
    from rl_algo_impls_b200.rollout import VecRollout
    from oracle.gae import gae_advantages
    from tests.synth import gae_inputs

    T, N, bs = 16, 10, 48  # 160 steps -> 3 full minibatches + one of 16
    inp = gae_inputs(3, T, N, 1, 0.05)
    rng = np.random.default_rng(0)
    obs = rng.standard_normal((T, N, 4), dtype=np.float32)
    actions = rng.integers(0, 2, size=(T, N))
    logprobs = rng.standard_normal((T, N), dtype=np.float32)
    r = VecRollout(cuda, inp["next_episode_starts"], inp["next_values"], obs, actions, inp["rewards"],
                   inp["episode_starts"], inp["values"], logprobs, None, gamma=0.99, gae_lambda=0.95)
    adv = gae_advantages(gamma=0.99, gae_lambda=0.95, **inp)
    assert r.total_steps == T * N and r.num_minibatches(bs) == 4
    np.testing.assert_array_equal(r.y_true, (adv + inp["values"]).reshape(-1))
    np.testing.assert_array_equal(r.y_pred, inp["values"].reshape(-1))
    torch.manual_seed(77)
    want_stream = minibatch_index_stream(T * N, bs, shuffle=True)
    torch.manual_seed(77)
    got = list(r.minibatches(bs, shuffle=True))
    assert [len(b) for b in got] == [48, 48, 48, 16]
    for mb, idx in zip(got, want_stream):
        i = idx.numpy()
        np.testing.assert_array_equal(mb.obs.cpu().numpy(), obs.reshape(-1, 4)[i])
        np.testing.assert_array_equal(mb.actions.cpu().numpy(), actions.reshape(-1)[i])
        np.testing.assert_array_equal(mb.logprobs.cpu().numpy(), logprobs.reshape(-1)[i])
        np.testing.assert_array_equal(mb.advantages.cpu().numpy(), adv.reshape(-1)[i])
        np.testing.assert_array_equal(mb.values.cpu().numpy(), inp["values"].reshape(-1)[i])
    got2 = list(r.minibatches(bs, shuffle=False))
    np.testing.assert_array_equal(got2[0].logprobs.cpu().numpy(), logprobs.reshape(-1)[:48])


@pytest.mark.parametrize("kind,HW", [("microrts", 64), ("lux", 64), ("microrts_ungated", 64), ("lux_ungated", 100),
                                     ("lux", 4096), ("microrts", 1000)])
def test_num_actions_matches_the_reference_rule(cuda, kind, HW):
    """a3: Batch.num_actions (rollout/rollout.py:130-180) -- one launch over the whole rollout
    (b200rl_gridnet_num_actions): (cell, plane) pairs with a valid action under the value-dependent gating, or cells
    with any valid action when no subaction mask is configured, plus log(#valid pick cells).  Exact counts; maps of
    more than one 256-cell chunk combine through atomics."""
    from oracle.distributions import gates_from_subaction_mask
    from oracle.rollout import num_actions
    from rl_algo_impls_b200 import spaces
    from rl_algo_impls_b200.rollout import VecRollout
    from tests.synth import LUX_GATES, LUX_NVEC, MICRORTS_GATES, MICRORTS_NVEC, gae_inputs, gridnet_inputs

    T, N = 3, 4
    nvec, gates, n_pick = (MICRORTS_NVEC, MICRORTS_GATES, 0) if kind.startswith("microrts") else (LUX_NVEC, LUX_GATES, 1)
    if kind.endswith("ungated"):
        gates = None
    g = gridnet_inputs(5, T * N, HW, nvec, n_pick, 0.3)
    acts = g["actions"].reshape(T, N, HW, len(nvec))
    mask = g["mask"].reshape(T, N, HW, -1)
    actions, masks = acts, mask
    if n_pick:
        actions = {"per_position": acts, "pick_position": g["pick_actions"].reshape(T, N, n_pick)}
        masks = {"per_position": mask, "pick_position": g["pick_mask"].reshape(T, N, n_pick, HW)}
    want = num_actions(actions, masks, gates_from_subaction_mask(gates), np.asarray(nvec))
    inp = gae_inputs(1, T, N, 1, 0.1)
    r = VecRollout(cuda, inp["next_episode_starts"], inp["next_values"], np.zeros((T, N, 2), np.float32), actions,
                   inp["rewards"], inp["episode_starts"], inp["values"], np.zeros((T, N), np.float32), masks, 0.99, 0.95,
                   subaction_mask=gates, action_plane_space=spaces.MultiDiscrete(nvec), include_num_actions=True)
    got = r.batch().num_actions.cpu().numpy().reshape(T, N)
    if n_pick:
        np.testing.assert_allclose(got, want, rtol=1e-6)
    else:
        np.testing.assert_array_equal(got, want)


def test_gather_full_size_round_trip(cuda):
    """The whole C4 rollout (12,288 rows: 75,776 B observations, 19,968 B masks, 1,792 B actions, four scalars)
    shuffled by one gather and unshuffled by a second: bit-identical to the source; every minibatch slice of the
    shuffled copy equals the rows its indices name."""
    from rl_algo_impls_b200 import ops

    M = 12288
    g = torch.Generator(device=cuda).manual_seed(1)
    srcs = [torch.rand((M, 74, 16, 16), device=cuda, generator=g),
            torch.rand((M, 256, 78), device=cuda, generator=g) < 0.5,
            torch.randint(0, 49, (M, 256, 7), device=cuda, generator=g, dtype=torch.uint8),
            torch.randn(M, device=cuda, generator=g), torch.randn(M, device=cuda, generator=g),
            torch.randn(M, device=cuda, generator=g), torch.randn(M, device=cuda, generator=g)]
    perm = torch.randperm(M, device=cuda, generator=g)
    inverse = torch.empty_like(perm)
    inverse[perm] = torch.arange(M, device=cuda)
    shuffled = ops.gather_rows(srcs, perm)
    restored = ops.gather_rows(shuffled, inverse)
    for s, r in zip(srcs, restored):
        assert torch.equal(s, r)
    mb = perm[3 * 3072:4 * 3072]
    for s, sh, o in zip(srcs, shuffled, ops.gather_rows(srcs, mb)):
        assert torch.equal(o, sh[3 * 3072:4 * 3072]) and torch.equal(o[:5], s[mb[:5]])


def test_gather_consumer_on_stream_without_sync(cuda):
    """The narrow-row grid is a programmatic dependent of the wide-row grid.  Whatever follows on the stream (the
    trunk reading the gathered observations) must see EVERY wide row: a consumer kernel that reads the last wide
    rows immediately, on the same stream with no synchronisation, inside a captured graph at the C4 minibatch size,
    replayed with fresh indices -- a wide grid still running when the consumer starts would leave stale rows."""
    from rl_algo_impls_b200 import ops

    M, B = 12288, 3072
    g = torch.Generator(device=cuda).manual_seed(5)
    obs = torch.rand((M, 74, 16, 16), device=cuda, generator=g)      # 75,776 B rows: wide path
    masks = torch.rand((M, 256, 78), device=cuda, generator=g) < 0.5  # 19,968 B rows: wide path
    logp = torch.randn(M, device=cuda, generator=g)                   # narrow path
    idx = torch.zeros(B, dtype=torch.int64, device=cuda)
    tail = 64  # the consumer reads the LAST gathered rows: the wide grid's last CTAs write them
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(2):
            o, m, lp = ops.gather_rows([obs, masks, logp], idx)
            _ = o[-tail:].clone()
    torch.cuda.current_stream().wait_stream(side)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        o, m, lp = ops.gather_rows([obs, masks, logp], idx)
        got_obs, got_mask, got_lp = o[-tail:].clone(), m[-tail:].clone(), lp.clone()  # consumers, same stream
        o.zero_(), m.zero_()  # stale rows of an earlier replay can never pass for fresh ones
    for rep in range(12):
        perm = torch.randperm(M, device=cuda, generator=g)[:B]
        idx.copy_(perm)
        graph.replay()
        torch.cuda.synchronize()
        assert torch.equal(got_obs, obs[perm[-tail:]]), f"replay {rep}: consumer saw a partially gathered obs"
        assert torch.equal(got_mask, masks[perm[-tail:]]), f"replay {rep}: consumer saw partially gathered masks"
        assert torch.equal(got_lp, logp[perm])
