"""What stands in for compute-sanitizer (closed on the B200 pool: profiles/r02/sanitizer/README.md): canaries around
every output buffer of the streaming kernels and bit-for-bit repeat-determinism of the kernels that mix
asynchronous-proxy (bulk / TMA) traffic with generic stores."""
import numpy as np
import pytest
import torch

from tests.synth import LUX_GATES, LUX_NVEC, MICRORTS_GATES, MICRORTS_NVEC, gridnet_inputs, ppo_inputs, to_torch

pytestmark = pytest.mark.gpu

PAD = 4096  # canary bytes on either side
PATTERN = 0xA5


class Guarded:
    """A tensor that lives `lead` bytes into a larger pattern-filled byte buffer."""

    def __init__(self, shape, dtype, device, lead=PAD):
        n = int(np.prod(shape)) * torch.empty((), dtype=dtype).element_size()
        self.raw = torch.full((lead + n + PAD,), PATTERN, dtype=torch.uint8, device=device)
        self.lead, self.n = lead, n
        self.t = self.raw[lead:lead + n].view(dtype).reshape(shape)

    def check(self, what):
        assert (self.raw[:self.lead] == PATTERN).all(), f"{what}: bytes BEFORE the buffer were written"
        assert (self.raw[self.lead + self.n:] == PATTERN).all(), f"{what}: bytes AFTER the buffer were written"


def _grid_args(cuda, B, HW, nvec, gates, n_pick, seed):
    from rl_algo_impls_b200 import ops

    inp = to_torch(gridnet_inputs(seed, B, HW, nvec, n_pick, 0.08), cuda)
    pp = to_torch(ppo_inputs(seed, B, 1), cuda)
    spec = ops.GridnetSpec.from_subaction_mask(nvec, gates, n_pick)
    return inp, pp, spec


@pytest.mark.parametrize("nvec,gates,n_pick,HW", [(MICRORTS_NVEC, MICRORTS_GATES, 0, 256), (LUX_NVEC, LUX_GATES, 1, 1024),
                                                  ((3, 5), None, 1, 100)])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("lead", [PAD, PAD + 4, PAD + 2])  # 16-byte aligned, 4-byte aligned, 2-byte aligned (bf16 only)
def test_fused_gridnet_loss_stays_inside_its_buffers(cuda, nvec, gates, n_pick, HW, dtype, lead):
    import ctypes as C

    from rl_algo_impls_b200 import _lib, ops

    if lead % 4 and dtype == torch.float32:
        pytest.skip("f32 rows are 4-byte aligned by construction")
    B = 7
    inp, pp, spec = _grid_args(cuda, B, HW, nvec, gates, n_pick, 5)
    logits = inp["logits"].to(dtype)
    old_logp = torch.full((B,), -20.0, device=cuda)
    h = ops.PpoHyper(clip_range=0.1, clip_range_vf=0.1, ent_coef=0.01, vf_coef=[0.5], vf_halving=True)
    g = ops._GridCall(spec, logits, inp["mask"], inp["pick_mask"], inp["actions"], inp["pick_actions"])
    call = ops.PpoCall(h, old_logp, pp["adv"], pp["old_values"], pp["returns"], pp["new_values"],
                       workspace_bytes=_lib.lib().b200rl_ppo_gridnet_workspace_bytes(B, HW, n_pick, 1))
    dlogits = Guarded(tuple(logits.shape), dtype, cuda, lead)
    dvalues = Guarded((B,), torch.float32, cuda)
    stats = Guarded((6 + 2,), torch.float32, cuda)
    call.args.dvalues, call.args.stats_out = dvalues.t.data_ptr(), stats.t.data_ptr()
    ws = Guarded((call.workspace.numel(),), torch.uint8, cuda)
    rc = _lib.lib().b200rl_ppo_gridnet_loss(C.byref(g.desc), logits.data_ptr(), g.mask.data_ptr(), ops._ptr(g.pick_mask),
                                            inp["actions"].data_ptr(), ops._ptr(inp["pick_actions"]), C.byref(call.args),
                                            dlogits.t.data_ptr(), None, None, ws.t.data_ptr(), ws.t.numel(),
                                            torch.cuda.current_stream().cuda_stream)
    _lib.check(rc, "b200rl_ppo_gridnet_loss")
    torch.cuda.synchronize()
    for name, buf in (("dlogits", dlogits), ("dvalues", dvalues), ("stats", stats), ("workspace", ws)):
        buf.check(name)
    # and the result is the regular one
    want = ops.ppo_gridnet_loss(h, spec, logits, inp["mask"], inp["pick_mask"], inp["actions"], inp["pick_actions"], old_logp,
                                pp["adv"], pp["old_values"], pp["returns"], pp["new_values"])
    assert torch.equal(dlogits.t, want.grads[0]) and torch.equal(stats.t, want.stats)


def test_gather_and_store_step_stay_inside_their_buffers(cuda):
    import ctypes as C

    from rl_algo_impls_b200 import _lib

    g = torch.Generator(device=cuda).manual_seed(3)
    M, B = 300, 77
    shapes = [((M, 74, 16, 16), torch.float32), ((M, 256, 78), torch.uint8), ((M, 3, 5), torch.uint8), ((M,), torch.float32),
              ((M, 13), torch.float32), ((M, 1, 9), torch.float64)]
    srcs = [(torch.rand(s, device=cuda, generator=g) * 200).to(d) for s, d in shapes]
    idx = torch.randperm(M, device=cuda, generator=g)[:B]
    for lead in (PAD, PAD + 1, PAD + 4):
        outs = [Guarded((B,) + s[1:], d, cuda, lead if d == torch.uint8 else PAD) for s, d in shapes]
        n = len(srcs)
        sa, da, rb = (C.c_void_p * n)(), (C.c_void_p * n)(), (C.c_int64 * n)()
        for k, (s, o) in enumerate(zip(srcs, outs)):
            sa[k], da[k], rb[k] = s.data_ptr(), o.t.data_ptr(), (s.numel() // M) * s.element_size()
        _lib.check(_lib.lib().b200rl_gather_rows(sa, da, rb, n, idx.data_ptr(), B, M, torch.cuda.current_stream().cuda_stream),
                   "b200rl_gather_rows")
        torch.cuda.synchronize()
        for k, (s, o) in enumerate(zip(srcs, outs)):
            o.check(f"gather dst {k} (lead {lead})")
            assert torch.equal(o.t, s[idx])
    # K0: one step's slices into row (step % T) of guarded [T, N, ...] buffers
    T, N = 5, 6
    step = torch.tensor([13], dtype=torch.int64, device=cuda)
    fields = [((N, 74, 16, 16), torch.float32), ((N, 256, 78), torch.uint8), ((N,), torch.float32), ((N, 3), torch.uint8)]
    slices = [(torch.rand(s, device=cuda, generator=g) * 200).to(d) for s, d in fields]
    bufs = [Guarded((T,) + s, d, cuda) for s, d in fields]
    for b in bufs:
        b.t.zero_()
    n = len(fields)
    sa, da, sb = (C.c_void_p * n)(), (C.c_void_p * n)(), (C.c_int64 * n)()
    for k, (s, b) in enumerate(zip(slices, bufs)):
        sa[k], da[k], sb[k] = s.data_ptr(), b.t.data_ptr(), s.numel() * s.element_size()
    _lib.check(_lib.lib().b200rl_rollout_store_step(sa, da, sb, n, step.data_ptr(), T, torch.cuda.current_stream().cuda_stream),
               "b200rl_rollout_store_step")
    torch.cuda.synchronize()
    for k, (s, b) in enumerate(zip(slices, bufs)):
        b.check(f"store_step buffer {k}")
        assert torch.equal(b.t[13 % T], s) and (b.t[[0, 1, 2, 4]] == 0).all()


@pytest.mark.parametrize("nvec,gates,n_pick,HW,B", [(MICRORTS_NVEC, MICRORTS_GATES, 0, 256, 512), (LUX_NVEC, LUX_GATES, 1, 4096, 48)])
def test_async_proxy_kernels_repeat_bit_for_bit(cuda, nvec, gates, n_pick, HW, B):
    """The fused loss zero-fills dlogits through the async proxy (cp.async.bulk) and then overwrites the unit cells' rows
    with generic stores; the order is only as good as its wait_group + fence.proxy.async.  Forty back-to-back launches
    into the SAME output buffer (pre-poisoned each time), with a second stream hammering HBM, must reproduce the first
    result bit for bit -- a zero fill landing late would leave zero rows where gradients belong, sporadically."""
    from rl_algo_impls_b200 import ops

    inp, pp, spec = _grid_args(cuda, B, HW, nvec, gates, n_pick, 17)
    old_logp = torch.full((B,), -25.0, device=cuda)
    V = 1
    h = ops.PpoHyper(clip_range=0.1, clip_range_vf=None, ent_coef=0.01, vf_coef=[0.5] * V)
    run = lambda: ops.ppo_gridnet_loss(h, spec, inp["logits"], inp["mask"], inp["pick_mask"], inp["actions"], inp["pick_actions"],
                                       old_logp, pp["adv"], pp["old_values"], pp["returns"], pp["new_values"])
    first = run()
    ref_grad, ref_stats = first.grads[0].clone(), first.stats.clone()
    assert ref_grad.abs().sum() > 0
    noise = torch.empty(64 << 20, dtype=torch.float32, device=cuda)
    side = torch.cuda.Stream()
    for rep in range(40):
        with torch.cuda.stream(side):
            noise.add_(1.0)
        out = run()
        assert torch.equal(out.grads[0], ref_grad), f"repeat {rep}: dlogits differ"
        assert torch.equal(out.stats, ref_stats), f"repeat {rep}: stats differ"
    side.synchronize()
    # the sampler: same seed / offset -> same draws
    a0, p0, l0 = ops.gridnet_sample(spec, inp["logits"], inp["mask"], inp["pick_mask"], 1234, 7)
    for rep in range(10):
        a, p, l = ops.gridnet_sample(spec, inp["logits"], inp["mask"], inp["pick_mask"], 1234, 7)
        assert torch.equal(a, a0) and torch.equal(l, l0) and (p is None or torch.equal(p, p0))
