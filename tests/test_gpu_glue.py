"""K8 parity: bias + max-pool + ReLU and bias + ReLU on channels-last maps against the PyTorch sequence they replace
between the convolutions of the GridNet encoder / decoder (shared/encoder/gridnet_encoder.py:26-51,
shared/actor/gridnet_decoder.py:36-53).  Forward: bit-exact.  Backward: the same terms summed in another order
(input gradient: <= 4 terms per element; bias gradient: a column sum) -> 1e-6 / 1e-5 relative."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

from tests.parity import close  # noqa: E402


def _map(shape, cuda, seed, channels_last=True, sparse=True):
    g = torch.Generator(device=cuda).manual_seed(seed)
    y = torch.randn(shape, device=cuda, generator=g)
    if sparse:  # ties: many equal entries inside a window, as after a ReLU / on one-hot planes
        y = torch.where(torch.rand(shape, device=cuda, generator=g) < 0.5, torch.zeros_like(y), y)
    return y.contiguous(memory_format=torch.channels_last) if channels_last else y


POOL_SHAPES = [
    # N, C, H, W, kernel, stride, padding
    (24, 32, 16, 16, 3, 2, 1),    # the C4 rollout step, encoder level 1
    (24, 256, 2, 2, 3, 2, 1),     # ... level 4 (every window hangs over the border)
    (7, 64, 8, 8, 3, 2, 1),
    (3, 6, 9, 11, 3, 2, 1),       # C % 4 != 0: scalar path; odd extents
    (5, 8, 10, 7, 2, 2, 0),       # non-overlapping windows, rows / columns left over
    (2, 12, 12, 12, 5, 3, 2),     # general window (runtime kernel size)
    (4, 16, 1, 1, 3, 2, 1),       # a 1x1 map
    (0, 16, 4, 4, 3, 2, 1),       # empty batch
]


@pytest.mark.parametrize("N,C,H,W,k,s,p", POOL_SHAPES)
@pytest.mark.parametrize("relu", [True, False])
@pytest.mark.parametrize("with_bias", [True, False])
def test_bias_pool_relu_matches_torch(cuda, N, C, H, W, k, s, p, relu, with_bias):
    from rl_algo_impls_b200 import ops

    y = _map((N, C, H, W), cuda, 1).requires_grad_(True)
    bias = (torch.randn(C, device=cuda) * 0.5).requires_grad_(True) if with_bias else None
    y2 = y.detach().clone().requires_grad_(True)
    b2 = bias.detach().clone().requires_grad_(True) if with_bias else None

    got = ops.bias_pool_relu(y, bias, k, s, p, relu)
    want = F.max_pool2d(y2 + b2[None, :, None, None] if with_bias else y2, k, s, p)
    if relu:
        want = F.relu(want)
    assert got.shape == want.shape and got.is_contiguous(memory_format=torch.channels_last)
    assert torch.equal(got, want)  # bit-exact forward
    if N == 0:
        return
    dout = _map(tuple(want.shape), cuda, 2, sparse=False)
    got.backward(dout)
    want.backward(dout)
    close(y.grad, y2.grad, rtol=1e-6, what="dx")
    # the gradient lands on exactly the positions torch routes it to (same arg-max rule, ties included)
    assert torch.equal(y.grad != 0, y2.grad != 0)
    if with_bias:
        close(bias.grad, b2.grad, rtol=1e-5, what="dbias")


def test_bias_pool_relu_no_grad_and_nan(cuda):
    from rl_algo_impls_b200 import ops

    y = _map((6, 32, 16, 16), cuda, 3)
    y[0, 0, 3, 3] = float("nan")
    y[1, 1, :, :] = float("-inf")
    bias = torch.randn(32, device=cuda)
    with torch.no_grad():
        got = ops.bias_pool_relu(y, bias)
        want = F.relu(F.max_pool2d(y + bias[None, :, None, None], 3, 2, 1))
    assert torch.equal(torch.isnan(got), torch.isnan(want)) and bool(torch.isnan(got).any())
    assert torch.equal(torch.nan_to_num(got, nan=7.0), torch.nan_to_num(want, nan=7.0))


def test_bias_pool_relu_accepts_nchw_strides(cuda):
    """A contiguous (NCHW) map is re-laid out, not misread."""
    from rl_algo_impls_b200 import ops

    y = _map((3, 8, 6, 6), cuda, 4, channels_last=False)
    bias = torch.randn(8, device=cuda)
    assert torch.equal(ops.bias_pool_relu(y, bias), F.relu(F.max_pool2d(y + bias[None, :, None, None], 3, 2, 1)))


@pytest.mark.parametrize("N,C,H,W", [(24, 128, 2, 2), (24, 32, 8, 8), (5, 6, 3, 7), (3072, 32, 8, 8), (256, 80, 16, 16),
                                     (0, 8, 2, 2)])
@pytest.mark.parametrize("relu", [True, False])
@pytest.mark.parametrize("inplace", [False, True])
def test_bias_relu_matches_torch(cuda, N, C, H, W, relu, inplace):
    """relu=False is the logit head's plain bias; inplace: the input is a non-leaf (as a convolution's output is) and
    is overwritten."""
    from rl_algo_impls_b200 import ops

    y = _map((N, C, H, W), cuda, 5).requires_grad_(True)
    bias = (torch.randn(C, device=cuda) * 0.5).requires_grad_(True)
    y2, b2 = y.detach().clone().requires_grad_(True), bias.detach().clone().requires_grad_(True)
    src = (y * 1.0).contiguous(memory_format=torch.channels_last) if inplace else y
    got = ops.bias_relu(src, bias, relu=relu)
    assert (got.data_ptr() == src.data_ptr()) == (inplace or N == 0)
    want = y2 + b2[None, :, None, None]
    if relu:
        want = F.relu(want)
    assert torch.equal(got, want) and got.is_contiguous(memory_format=torch.channels_last)
    if N == 0:
        return
    dout = _map(tuple(want.shape), cuda, 6, sparse=False)
    got.backward(dout)
    want.backward(dout)
    assert torch.equal(y.grad, y2.grad)  # a mask (or the identity): exact
    close(bias.grad, b2.grad, rtol=1e-5, what="dbias")


def test_bias_gradient_is_deterministic_and_large_batch(cuda):
    """The C4 minibatch at encoder level 1 (3072 x 32 x 16 x 16): two runs give the same bits; against float64 sums."""
    from rl_algo_impls_b200 import ops

    y = _map((3072, 32, 16, 16), cuda, 7)
    bias = torch.randn(32, device=cuda) * 0.1
    dout = _map((3072, 32, 8, 8), cuda, 8, sparse=False)
    grads = []
    for _ in range(2):
        yy, bb = y.clone().requires_grad_(True), bias.clone().requires_grad_(True)
        ops.bias_pool_relu(yy, bb).backward(dout)
        grads.append((yy.grad, bb.grad))
    assert torch.equal(grads[0][0], grads[1][0]) and torch.equal(grads[0][1], grads[1][1])
    y64, b64 = y.double().requires_grad_(True), bias.double().requires_grad_(True)
    F.relu(F.max_pool2d(y64 + b64[None, :, None, None], 3, 2, 1)).backward(dout.double())
    close(grads[0][1], b64.grad, rtol=1e-5, what="dbias vs float64")
    close(grads[0][0], y64.grad, rtol=1e-6, what="dx vs float64")


@pytest.mark.parametrize("tf32", [True, False])
def test_gridnet_trunk_fused_glue_equals_torch_modules(cuda, monkeypatch, tf32):
    """The MicroRTS trunk with the fused glue against the same module running its PyTorch layers, eager and inside a
    captured graph.  With cuDNN's default TF32 convolutions (what the learner runs) the outputs are bit-identical and
    the parameter gradients agree to TF32 noise: a last-bit difference in a convolution's input (the re-ordered
    <= 4-term sums of the pooling backward) can round to another TF32 value.  With full-precision convolutions the
    library may pick other convolution kernels for the two module graphs (outputs to 1e-6), and the gradients agree
    to 2e-5."""
    from rl_algo_impls_b200.policy.networks import GridEncoderDecoderActorCritic

    monkeypatch.setattr(torch.backends.cudnn, "allow_tf32", tf32)
    torch.manual_seed(0)
    net = GridEncoderDecoderActorCritic(74, (16, 16), 78).to(cuda)
    obs = (torch.rand((48, 74, 16, 16), device=cuda) < 0.1).float()
    outs, grads = [], []
    for fused in (True, False):
        net.fused_glue = fused
        net.zero_grad(set_to_none=True)
        out = net(obs)
        (out.pi.square().mean() + out.values.square().mean()).backward()
        outs.append((out.pi.detach().clone(), out.values.detach().clone()))
        grads.append({n: p.grad.clone() for n, p in net.named_parameters()})
    if tf32:
        assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])
    else:
        close(outs[0][0], outs[1][0], rtol=1e-6, what="logits")
        close(outs[0][1], outs[1][1], rtol=1e-6, what="values")
    for n in grads[0]:
        close(grads[0][n], grads[1][n], rtol=5e-4 if tf32 else 2e-5, what=n)
    # no-grad evaluation (the rollout step), captured
    net.eval()
    net.fused_glue = True
    static = obs[:24].clone()
    with torch.no_grad():
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(2):
                net(static)
        torch.cuda.current_stream().wait_stream(side)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            captured = net(static)
        static.copy_(obs[24:48])
        graph.replay()
        net.fused_glue = False
        want = net(obs[24:48])
    if tf32:
        assert torch.equal(captured.pi, want.pi) and torch.equal(captured.values, want.values)
    else:
        close(captured.pi, want.pi, rtol=1e-6, what="captured logits")
        close(captured.values, want.values, rtol=1e-6, what="captured values")


# ---- K9: squeeze U-net glue (float32 / bfloat16) --------------------------------------------------------------------
def _bias_act_ref(y, bias, act):
    v = y + bias.to(y.dtype)[None, :, None, None]
    return F.gelu(v) if act == "gelu" else (F.relu(v) if act == "relu" else v)


@pytest.mark.parametrize("N,C,H,W", [(4, 128, 16, 16), (3, 24, 5, 7), (2, 6, 3, 3), (0, 8, 2, 2)])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("act", ["gelu", "relu", "none"])
def test_bias_act_matches_torch(cuda, N, C, H, W, dtype, act):
    """Forward bit-exact against the PyTorch chain in the map's dtype (bias rounded as autocast rounds it, the sum rounded
    before the activation); backward to the dtype's precision, bias gradient against the float32 column sums."""
    from rl_algo_impls_b200 import ops

    y = _map((N, C, H, W), cuda, 11, sparse=False).to(dtype).requires_grad_(True)
    bias = (torch.randn(C, device=cuda) * 0.5).requires_grad_(True)
    y2, b2 = y.detach().clone().requires_grad_(True), bias.detach().clone().requires_grad_(True)
    got = ops.bias_act(y, bias, act)
    want = _bias_act_ref(y2, b2, act)
    assert got.dtype == dtype and torch.equal(got, want)
    if N == 0:
        return
    dout = _map(tuple(want.shape), cuda, 12, sparse=False).to(dtype)
    got.backward(dout)
    want.backward(dout)
    tol = 1e-6 if dtype == torch.float32 else 2 ** -7
    close(y.grad, y2.grad, rtol=tol, what="dx")
    close(bias.grad, b2.grad, rtol=1e-5 if dtype == torch.float32 else 2e-2, what="dbias")
    # the bias gradient is the float32 column sum of the input gradient
    close(bias.grad, y.grad.float().sum((0, 2, 3)), rtol=1e-5, what="dbias vs column sums")


class _TorchSEBlock(torch.nn.Module):
    """The PyTorch modules the fused tail replaces (networks._SEResBlock with FUSED_GLUE off)."""

    def __init__(self, c):
        super().__init__()
        from rl_algo_impls_b200.policy import networks

        self.block = networks._SEResBlock(c)

    def forward(self, x):
        from rl_algo_impls_b200.policy import networks

        keep, networks.FUSED_GLUE = networks.FUSED_GLUE, False
        try:
            return self.block(x)
        finally:
            networks.FUSED_GLUE = keep


@pytest.mark.parametrize("N,C,H,W", [(6, 128, 16, 16), (3, 32, 8, 8), (2, 48, 5, 3)])
@pytest.mark.parametrize("autocast", [False, True])
def test_se_residual_block_fused_equals_torch_modules(cuda, monkeypatch, N, C, H, W, autocast):
    """One SE-residual block, fused (bias_act + se_tail) against its PyTorch modules on the same weights: outputs and every
    gradient.  float32: full-precision convolutions, outputs 1e-6 (the squeeze mean is summed in another order), gradients
    1e-5.  bfloat16 autocast: bf16 bars (the fused path rounds at the same points, sums in float32)."""
    from rl_algo_impls_b200.policy import networks

    monkeypatch.setattr(torch.backends.cudnn, "allow_tf32", False)
    monkeypatch.setattr(networks, "FUSED_GLUE", True)
    torch.manual_seed(3)
    ref = _TorchSEBlock(C).to(cuda).to(memory_format=torch.channels_last)
    for p in ref.parameters():  # non-trivial biases and gates
        if p.dim() == 1:
            torch.nn.init.normal_(p, std=0.3)
    x0 = _map((N, C, H, W), cuda, 21, sparse=False)
    dout = _map((N, C, H, W), cuda, 22, sparse=False)
    res = []
    for fused in (True, False):
        x = x0.clone().requires_grad_(True)
        ref.zero_grad(set_to_none=True)
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
            xin = F.gelu(x * 1.0) if not autocast else F.gelu((x * 1.0).to(torch.bfloat16))  # a map of the block's dtype
            out = ref.block(xin) if fused else ref(xin)
        out.backward(dout.to(out.dtype))
        res.append((out.detach().float(), x.grad.clone(), {n: p.grad.clone() for n, p in ref.named_parameters()}))
    (o1, gx1, gp1), (o0, gx0, gp0) = res
    if autocast:
        close(o1, o0, rtol=2 ** -7, what="out (bf16)")
        close(gx1, gx0, rtol=3e-2, what="dx (bf16)")
        for n in gp0:
            close(gp1[n], gp0[n], rtol=5e-2, what=n + " (bf16)")
    else:
        close(o1, o0, rtol=1e-6, what="out")
        close(gx1, gx0, rtol=1e-5, what="dx")
        for n in gp0:
            close(gp1[n], gp0[n], rtol=2e-5, what=n)


def test_squeeze_unet_fused_glue_equals_torch_modules(cuda, monkeypatch):
    """The whole squeeze U-net (a small Lux-shaped one), fused against PyTorch modules: float32, full-precision
    convolutions; logits / values 1e-5, parameter gradients 1e-4 of their scale."""
    from rl_algo_impls_b200.policy import networks

    monkeypatch.setattr(torch.backends.cudnn, "allow_tf32", False)
    torch.manual_seed(0)
    net = networks.SqueezeUnetActorCritic(75, 29, channels_per_level=[32, 32, 32], strides_per_level=[4, 4],
                                          deconv_strides_per_level=[[2, 2], [2, 2]], encoder_residual_blocks_per_level=[2, 1, 1],
                                          decoder_residual_blocks_per_level=[1, 2], critic_channels=32,
                                          critic_activations=["identity"] * 3, shared_critic_head=True, obs_hw=(32, 32)).to(cuda)
    obs = (torch.rand((5, 75, 32, 32), device=cuda) < 0.1).float()
    res = []
    for fused in (True, False):
        monkeypatch.setattr(networks, "FUSED_GLUE", fused)
        net.zero_grad(set_to_none=True)
        out = net(obs)
        (out.pi.square().mean() + out.values.square().mean()).backward()
        res.append((out.pi.detach().clone(), out.values.detach().clone(), {n: p.grad.clone() for n, p in net.named_parameters()}))
    close(res[0][0], res[1][0], rtol=1e-5, what="logits")
    close(res[0][1], res[1][1], rtol=1e-5, what="values")
    for n in res[0][2]:
        close(res[0][2][n], res[1][2][n], rtol=1e-4, what=n)


def test_k9_at_the_c5_minibatch_size(cuda, monkeypatch):
    """One C5 per-GPU minibatch at encoder level 0 (128 x 128 x 64 x 64, bfloat16 autocast, 134 MB per map): bias + GELU
    bit-exact against the PyTorch chain; the SE-residual block fused vs its PyTorch modules to bf16 precision; the
    float32 reductions (bias / gate gradients) deterministic across runs."""
    from rl_algo_impls_b200 import ops
    from rl_algo_impls_b200.policy import networks

    monkeypatch.setattr(networks, "FUSED_GLUE", True)
    N, C, H, W = 128, 128, 64, 64
    y = _map((N, C, H, W), cuda, 31, sparse=False).to(torch.bfloat16)
    bias = torch.randn(C, device=cuda) * 0.5
    assert torch.equal(ops.bias_act(y, bias, "gelu"), F.gelu(y + bias.to(torch.bfloat16)[None, :, None, None]))
    torch.manual_seed(5)
    ref = _TorchSEBlock(C).to(cuda).to(memory_format=torch.channels_last)
    x0 = F.gelu(_map((N, C, H, W), cuda, 32, sparse=False)).to(torch.bfloat16)
    dout = _map((N, C, H, W), cuda, 33, sparse=False).to(torch.bfloat16)
    runs = []
    for mode in ("fused", "fused", "torch"):
        x = x0.clone().requires_grad_(True)
        ref.zero_grad(set_to_none=True)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            out = ref.block(x) if mode == "fused" else ref(x)
        out.backward(dout)
        runs.append((out.detach(), x.grad.detach(), {n: p.grad.clone() for n, p in ref.named_parameters()}))
        del out
    (o1, g1, p1), (o2, g2, p2), (o0, g0, p0) = runs
    # deterministic where only these kernels (and forward convolutions) are upstream: the output, the second
    # convolution's bias gradient, the gate's linears (cuDNN's dgrad / wgrad make no such promise for the rest)
    assert torch.equal(o1, o2)
    for n in p1:
        if n.endswith("residual.2.bias") or ".fc." in n:
            assert torch.equal(p1[n], p2[n]), n
    close(o1.float(), o0.float(), rtol=2 ** -7, what="out")
    close(g1.float(), g0.float(), rtol=3e-2, what="dx")
    for n in p0:
        close(p1[n], p0[n], rtol=5e-2, what=n)
