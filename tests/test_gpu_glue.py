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
