"""Policy / value trunks.  These stay on PyTorch's own CUDA ops (BASELINE.json north_star): the
data path only needs modules that turn observations into head outputs (logits or a Gaussian mean,
and values) so that the fused loss kernels have something real to differentiate through.

Architectures follow the families the reference configs name:
MLP actor-critic (CartPole / HalfCheetah, shared/policy/actor_critic_network/connected_trio.py),
NatureCNN (Atari, shared/encoder/nature_cnn.py), a conv encoder + transposed-conv decoder for
MicroRTS GridNet (shared/encoder/gridnet_encoder.py + shared/actor/gridnet_decoder.py) and a
U-shaped residual backbone with several critic heads for Lux (actor_critic_network/squeeze_unet.py).
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


class _ResBlock(nn.Module):
    def __init__(self, c: int):
        super().__init__()
        self.a, self.b = nn.Conv2d(c, c, 3, padding=1), nn.Conv2d(c, c, 3, padding=1)
        self.act = nn.GELU()

    def forward(self, x):
        return x + self.b(self.act(self.a(self.act(x))))


class UShapedActorCritic(nn.Module):
    """Lux-style U-shaped residual backbone: stride-4 down levels with residual blocks, transposed
    convs back up with skip connections, a 3x3 conv actor head emitting [B, H, W, S'] logits and
    `n_values` critic heads (conv + global average pool + linear) on the backbone output."""

    def __init__(self, in_channels: int, n_logits: int, n_values: int = 1, channels=(64, 128, 256),
                 strides=(4, 4), blocks=(1, 1, 1), critic_channels: int = 128):
        super().__init__()
        self.stem = nn.Conv2d(in_channels, channels[0], 3, padding=1)
        self.enc = nn.ModuleList([nn.Sequential(*[_ResBlock(c) for _ in range(n)]) for c, n in zip(channels, blocks)])
        self.down = nn.ModuleList([nn.Conv2d(channels[i], channels[i + 1], strides[i], stride=strides[i])
                                   for i in range(len(strides))])
        self.up = nn.ModuleList([nn.ConvTranspose2d(channels[i + 1], channels[i], strides[i], stride=strides[i])
                                 for i in range(len(strides))])
        self.dec = nn.ModuleList([nn.Sequential(*[_ResBlock(c) for _ in range(n)])
                                  for c, n in zip(channels[:-1], blocks[:-1])])
        self.actor = ortho_(nn.Conv2d(channels[0], n_logits, 3, padding=1), 0.01)
        self.critic_conv = nn.Sequential(nn.Conv2d(channels[0], critic_channels, 3, stride=2, padding=1), nn.GELU())
        self.critic_out = ortho_(nn.Linear(critic_channels, n_values), 1.0)
        self.n_values = n_values
        self.policy_head_modules, self.value_head_modules = ["actor"], ["critic_conv", "critic_out"]  # freeze_* (unet.py:221-245)
        self.to(memory_format=torch.channels_last)  # see GridEncoderDecoderActorCritic

    def forward(self, obs: torch.Tensor) -> HeadOutputs:
        x = self.stem(obs.float().contiguous(memory_format=torch.channels_last))
        skips = []
        for i, enc in enumerate(self.enc):
            x = enc(x)
            if i < len(self.down):
                skips.append(x)
                x = self.down[i](x)
        for i in reversed(range(len(self.up))):
            x = self.dec[i](self.up[i](x) + skips[i])
        logits = self.actor(x).permute(0, 2, 3, 1)  # [B, H, W, S']
        v = self.critic_out(self.critic_conv(x).mean(dim=(2, 3)))
        return HeadOutputs(logits, v.squeeze(-1) if self.n_values == 1 else v)
