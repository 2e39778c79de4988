"""Policy / value trunks.  These stay on PyTorch's own CUDA ops (BASELINE.json north_star): the
data path only needs modules that turn observations into head outputs (logits or a Gaussian mean,
and values) so that the fused loss kernels have something real to differentiate through.

Architectures follow the families the reference configs name:
MLP actor-critic (CartPole / HalfCheetah, shared/policy/actor_critic_network/connected_trio.py),
NatureCNN (Atari, shared/encoder/nature_cnn.py), a conv encoder + transposed-conv decoder for
MicroRTS GridNet (shared/encoder/gridnet_encoder.py + shared/actor/gridnet_decoder.py) and the
squeeze U-net with SE-residual blocks and several critic outputs for Lux (actor_critic_network/squeeze_unet.py).
"""
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn as nn

_ACTIVATIONS = {"tanh": nn.Tanh, "relu": nn.ReLU, "gelu": nn.GELU, "identity": nn.Identity}


def ortho_(layer: nn.Module, std: float = float(np.sqrt(2)), bias: float = 0.0) -> nn.Module:
    """Orthogonal weights / constant bias (shared/module/utils.py:36-44 initialisation scheme)."""
    nn.init.orthogonal_(layer.weight, std)
    if layer.bias is not None:
        nn.init.constant_(layer.bias, bias)
    return layer


def mlp(sizes: Sequence[int], activation: str, out_std: float = float(np.sqrt(2)), final_activation: bool = False):
    layers: List[nn.Module] = []
    for i in range(len(sizes) - 1):
        last = i == len(sizes) - 2
        layers.append(ortho_(nn.Linear(sizes[i], sizes[i + 1]), out_std if last else float(np.sqrt(2))))
        if not last or final_activation:
            layers.append(_ACTIVATIONS[activation]())
    return nn.Sequential(*layers)


class HeadOutputs:
    """Raw head outputs of one forward pass: what the fused loss kernels consume."""

    __slots__ = ("pi", "log_std", "values")

    def __init__(self, pi: torch.Tensor, values: torch.Tensor, log_std: Optional[torch.Tensor] = None):
        self.pi = pi  # logits [B, n] / [B, HW, S'] or the Gaussian mean [B, D]
        self.values = values  # [B] or [B, V]
        self.log_std = log_std


class MlpActorCritic(nn.Module):
    """Separate tanh MLPs for policy and value over a flat observation."""

    def __init__(self, obs_dim: int, pi_out: int, pi_hidden=(64, 64), v_hidden=(64, 64), activation="tanh",
                 gaussian: bool = False, log_std_init: float = -0.5):
        super().__init__()
        self.pi = mlp([obs_dim, *pi_hidden, pi_out], activation, out_std=0.01)
        self.v = mlp([obs_dim, *v_hidden, 1], activation, out_std=1.0)
        self.log_std = nn.Parameter(torch.full((pi_out,), float(log_std_init))) if gaussian else None
        # freeze_* (separate_actor_critic.py:115-126): the output layers are the heads, the hidden layers the backbone
        self.policy_head_modules = [f"pi.{len(self.pi) - 1}"] + (["log_std"] if gaussian else [])
        self.value_head_modules = [f"v.{len(self.v) - 1}"]

    def forward(self, obs: torch.Tensor) -> HeadOutputs:
        x = obs.float().reshape(obs.shape[0], -1)
        return HeadOutputs(self.pi(x), self.v(x).squeeze(-1), self.log_std)


class NatureCnnActorCritic(nn.Module):
    """Conv 8x8/4 -> 4x4/2 -> 3x3/1 -> 512 features; linear policy and value heads (uint8 frames / 255)."""

    def __init__(self, in_channels: int, n_actions: int, hw: Tuple[int, int] = (84, 84), flatten_dim: int = 512):
        super().__init__()
        self.cnn = nn.Sequential(
            ortho_(nn.Conv2d(in_channels, 32, 8, stride=4)), nn.ReLU(),
            ortho_(nn.Conv2d(32, 64, 4, stride=2)), nn.ReLU(),
            ortho_(nn.Conv2d(64, 64, 3, stride=1)), nn.ReLU(), nn.Flatten(),
        )  # fmt: skip
        with torch.no_grad():
            n = self.cnn(torch.zeros(1, in_channels, *hw)).shape[1]
        self.fc = nn.Sequential(ortho_(nn.Linear(n, flatten_dim)), nn.ReLU())
        self.pi = ortho_(nn.Linear(flatten_dim, n_actions), 0.01)
        self.v = ortho_(nn.Linear(flatten_dim, 1), 1.0)
        self.policy_head_modules, self.value_head_modules = ["pi"], ["v"]  # freeze_*: cnn + fc are the backbone
        self.to(memory_format=torch.channels_last)

    def forward(self, obs: torch.Tensor) -> HeadOutputs:
        x = self.fc(self.cnn((obs.float() / 255.0).contiguous(memory_format=torch.channels_last)))
        return HeadOutputs(self.pi(x), self.v(x).squeeze(-1))


class GridEncoderDecoderActorCritic(nn.Module):
    """MicroRTS GridNet: 4 x (conv3x3 + maxpool/2) encoder, 4 x transposed-conv decoder emitting
    [B, H, W, S] logits, and an MLP critic on the encoded map."""

    def __init__(self, in_channels: int, map_hw: Tuple[int, int], n_logits: int, v_hidden=(128,), n_values: int = 1):
        super().__init__()
        chans = [in_channels, 32, 64, 128, 256]
        enc: List[nn.Module] = []
        for i in range(4):
            enc += [ortho_(nn.Conv2d(chans[i], chans[i + 1], 3, padding=1)), nn.MaxPool2d(3, stride=2, padding=1), nn.ReLU()]
        self.encoder = nn.Sequential(*enc)
        dchans = [256, 128, 64, 32, n_logits]
        dec: List[nn.Module] = []
        for i in range(4):
            last = i == 3
            dec.append(ortho_(nn.ConvTranspose2d(dchans[i], dchans[i + 1], 3, stride=2, padding=1, output_padding=1),
                              0.01 if last else float(np.sqrt(2))))
            if not last:
                dec.append(nn.ReLU())
        self.decoder = nn.Sequential(*dec)
        with torch.no_grad():
            feat = self.encoder(torch.zeros(1, in_channels, *map_hw))
        self.critic = nn.Sequential(nn.Flatten(), mlp([int(np.prod(feat.shape[1:])), *v_hidden, n_values], "relu", 1.0))
        self.n_values = n_values
        self.policy_head_modules, self.value_head_modules = ["decoder"], ["critic"]  # freeze_* (backbone_actor_critic.py:254-265)
        # NHWC in memory: cuDNN's native layout on sm_100 (no nchw<->nhwc transposes around every conv),
        # and the decoder output is then physically [B, H, W, S] -- the layout the fused loss kernel
        # reads and writes -- so permute(0, 2, 3, 1) is a free view instead of a 245 MB copy per minibatch.
        self.to(memory_format=torch.channels_last)

    def forward(self, obs: torch.Tensor) -> HeadOutputs:
        z = self.encoder(obs.float().contiguous(memory_format=torch.channels_last))
        logits = self.decoder(z).permute(0, 2, 3, 1)  # [B, H, W, S]
        v = self.critic(z)
        return HeadOutputs(logits, v.squeeze(-1) if self.n_values == 1 else v)


class _SqueezeExcite(nn.Module):
    """Channel gate: global mean -> C/16 bottleneck (GELU) -> sigmoid scale (double_cone.py:18-47)."""

    def __init__(self, c: int, reduction: int = 16):
        super().__init__()
        self.fc = nn.Sequential(nn.Linear(c, c // reduction, bias=False), nn.GELU(), nn.Linear(c // reduction, c, bias=False),
                                nn.Sigmoid())

    def forward(self, x):
        return x * self.fc(x.mean(dim=(2, 3)))[:, :, None, None]


class _SEResBlock(nn.Module):
    """gelu(x + SE(conv3x3(gelu(conv3x3(x))))) (double_cone.py:50-86, normalization=None)."""

    def __init__(self, c: int):
        super().__init__()
        self.residual = nn.Sequential(nn.Conv2d(c, c, 3, padding=1), nn.GELU(), nn.Conv2d(c, c, 3, padding=1), _SqueezeExcite(c))
        self.act = nn.GELU()

    def forward(self, x):
        return self.act(x + self.residual(x))


def _stride_list(s) -> List[int]:
    return [int(v) for v in s] if isinstance(s, (list, tuple)) else [int(s)]


class SqueezeUnetActorCritic(nn.Module):
    """The Lux / MicroRTS "squeeze U-net" (actor_critic_network/squeeze_unet.py:20-196 backbone +
    backbone_actor_critic.py:94-187 heads), same layer list and therefore the same parameter count (4,719,274 at the
    Lux 64x64 YAML entry): a 3x3 stem and SE-residual blocks per level, strided convs down (kernel = stride), chains of
    transposed convs up (``deconv_strides_per_level``; a level's encoder output is ADDED to what comes up from below),
    a 3x3 conv actor head emitting [B, H, W, S'] logits, and critic head(s) of strided convs -> global average pool ->
    two linears with a per-head output activation (``shared_critic_head``: one head with ``n_values`` outputs)."""

    def __init__(self, in_channels: int, n_logits: int, channels_per_level=(64, 128, 256), strides_per_level=None,
                 deconv_strides_per_level=None, encoder_residual_blocks_per_level=None,
                 decoder_residual_blocks_per_level=None, critic_channels: int = 64,
                 critic_activations: Sequence[str] = ("identity",), shared_critic_head: bool = False,
                 increment_kernel_size_on_down_conv: bool = False, obs_range: float = 1.0):
        super().__init__()
        ch = [int(c) for c in channels_per_level]
        L = len(ch)
        strides = list(strides_per_level) if strides_per_level is not None else [2] * (L - 1)
        up_strides = list(deconv_strides_per_level) if deconv_strides_per_level is not None else strides
        enc_blocks = list(encoder_residual_blocks_per_level) if encoder_residual_blocks_per_level is not None else [1] * L
        dec_blocks = list(decoder_residual_blocks_per_level) if decoder_residual_blocks_per_level is not None else enc_blocks[:-1]
        assert len(strides) == L - 1 and len(enc_blocks) == L and len(dec_blocks) == L - 1
        self.obs_range = float(obs_range)

        def down(cin: int, cout: int, stride) -> List[nn.Module]:
            layers: List[nn.Module] = []
            for i, s in enumerate(_stride_list(stride)):
                k, pad = (s + 1, 1) if increment_kernel_size_on_down_conv and s % 2 == 0 else (s, 0)
                layers += [nn.Conv2d(cin if i == 0 else cout, cout, k, stride=s, padding=pad), nn.GELU()]
            return layers

        def up(cin: int, cout: int, stride) -> List[nn.Module]:
            layers: List[nn.Module] = []
            for i, s in enumerate(_stride_list(stride)):
                layers += [nn.ConvTranspose2d(cin if i == 0 else cout, cout, s, stride=s), nn.GELU()]
            return layers

        self.encoders = nn.ModuleList([nn.Sequential(nn.Conv2d(in_channels, ch[0], 3, padding=1), nn.GELU(),
                                                     *[_SEResBlock(ch[0]) for _ in range(enc_blocks[0])])])
        for lvl in range(1, L):
            self.encoders.append(nn.Sequential(*down(ch[lvl - 1], ch[lvl], strides[lvl - 1]),
                                               *[_SEResBlock(ch[lvl]) for _ in range(enc_blocks[lvl])]))
        # decoders, deepest first: level L-1 only goes up; a middle level runs its blocks, then goes up; level 0 only blocks
        self.decoders = nn.ModuleList([nn.Sequential(*up(ch[-1], ch[-2], up_strides[-1]))])
        for lvl in range(L - 2, 0, -1):
            # squeeze_unet.py:153-158 pairs level lvl with decoder_residual_blocks_per_level[lvl - 1] (its zip runs over
            # the list without its LAST entry, which is therefore never used); kept, or parameter counts would differ
            self.decoders.append(nn.Sequential(*[_SEResBlock(ch[lvl]) for _ in range(dec_blocks[lvl - 1])],
                                               *up(ch[lvl], ch[lvl - 1], up_strides[lvl - 1])))
        self.decoders.append(nn.Sequential(*[_SEResBlock(ch[0]) for _ in range(dec_blocks[0])]))
        self.actor = ortho_(nn.Conv2d(ch[0], n_logits, 3, padding=1), 0.01)

        flat_strides = [s for st in strides for s in _stride_list(st)]

        def critic(n_out: int) -> nn.Sequential:
            layers: List[nn.Module] = []
            cin = ch[0]
            for s in flat_strides:
                k = max(3, s)
                layers += [nn.Conv2d(cin, critic_channels, k, stride=s, padding=1 if k % 2 else 0), nn.GELU()]
                cin = critic_channels
            layers += [nn.AdaptiveAvgPool2d(1), nn.Flatten(), ortho_(nn.Linear(critic_channels, critic_channels)), nn.GELU(),
                       ortho_(nn.Linear(critic_channels, n_out), 1.0)]
            return nn.Sequential(*layers)

        self.critic_activations = [a for a in critic_activations]
        self.n_values = len(self.critic_activations)
        self.shared_critic_head = bool(shared_critic_head)
        self.critics = nn.ModuleList([critic(self.n_values)] if shared_critic_head else [critic(1) for _ in range(self.n_values)])
        self.policy_head_modules, self.value_head_modules = ["actor"], ["critics"]  # freeze_* (backbone_actor_critic.py:254-265)
        self.to(memory_format=torch.channels_last)  # see GridEncoderDecoderActorCritic

    def _values(self, x: torch.Tensor) -> torch.Tensor:
        v = self.critics[0](x) if self.shared_critic_head else torch.cat([c(x) for c in self.critics], dim=1)
        if any(a != "identity" for a in self.critic_activations):  # ChannelwiseActivation: one activation per head
            v = torch.stack([_apply_activation(a, v[:, i]) for i, a in enumerate(self.critic_activations)], dim=1)
        return v.squeeze(-1) if self.n_values == 1 else v

    def forward(self, obs: torch.Tensor) -> HeadOutputs:
        x = obs.float().contiguous(memory_format=torch.channels_last)
        if self.obs_range != 1.0:
            x = x / self.obs_range  # backbone_actor_critic.py:189-192
        skips = []
        for enc in self.encoders:
            x = enc(x)
            skips.append(x)
        x = self.decoders[0](skips[-1])
        for skip, dec in zip(reversed(skips[:-1]), list(self.decoders)[1:]):
            x = dec(skip + x)
        logits = self.actor(x).permute(0, 2, 3, 1)  # [B, H, W, S']
        return HeadOutputs(logits, self._values(x))


def _apply_activation(name: str, x: torch.Tensor) -> torch.Tensor:
    if name == "identity":
        return x
    if name == "tanh":
        return torch.tanh(x)
    if name == "relu":
        return torch.relu(x)
    if name == "sigmoid":
        return torch.sigmoid(x)
    raise NotImplementedError(f"critic output activation {name!r}")
