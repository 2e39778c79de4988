"""Policy / value trunks.  Their convolutions and linears stay on PyTorch's own CUDA ops (cuDNN / cuBLAS; BASELINE.json
north_star): the data path only needs modules that turn observations into head outputs (logits or a Gaussian mean,
and values) so that the fused loss kernels have something real to differentiate through.  What runs BETWEEN the
convolutions of the two grid trunks on the channels-last maps -- bias adds, max-pool, ReLU / GELU, the squeeze-excite
mean / gate / residual sum -- goes through the fused K8 / K9 ops (``FUSED_GLUE``; bit-identical forward, same modules,
parameters and state dict).

Architectures follow the families the reference configs name:
MLP actor-critic (CartPole / HalfCheetah, shared/policy/actor_critic_network/connected_trio.py),
NatureCNN (Atari, shared/encoder/nature_cnn.py), a conv encoder + transposed-conv decoder for
MicroRTS GridNet (shared/encoder/gridnet_encoder.py + shared/actor/gridnet_decoder.py) and the
squeeze U-net with SE-residual blocks and several critic outputs for Lux (actor_critic_network/squeeze_unet.py).
"""
import os
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import ops

# The launches between the trunks' convolutions run as fused kernels (ops.bias_pool_relu / bias_relu for the MicroRTS
# encoder-decoder, ops.bias_act / se_tail for the squeeze U-net) when the maps are float32 / bfloat16 CUDA tensors.
# False (or B200RL_FUSED_GLUE=0 in the environment): the PyTorch modules themselves -- parity tests compare the two.
FUSED_GLUE = os.environ.get("B200RL_FUSED_GLUE", "1") != "0"

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


def _pad8(n: int) -> int:
    return (n + 7) // 8 * 8


class _PaddedEnds(nn.Module):
    """Channel padding at the two ends of a grid trunk, so that the convolution library never has to repair a layout.

    cuDNN's NHWC tensor-op kernels want channel counts that are multiples of 8 (16 bytes of bf16, 32 of f32); with 74 /
    75 observation planes and 78 / 29 logit planes it pads both tensors itself, on every call (`nhwcAddPaddingKernel`,
    `nchwToFoldedNhwcKernel`: 11 % of a C4 step's GPU time in round 1).  Here the padding is part of the data layout:

    * observations arrive PACKED -- [B, H, W, Cp] with the planes beyond C zero (``pack_observations``; the rollout
      buffer stores them like this, so the minibatch gather hands over trunk-ready rows) -- and the first convolution
      runs on its weight zero-padded to Cp input channels;
    * the head that emits the logits runs on its weight / bias zero-padded to Lp output channels, so the logits are
      physically [B, H, W, Lp]: rows of 80 (32) elements, 16-byte aligned, handed to the loss kernels as they are
      (b200rl.h ``logits_ld``); columns L.. are exact zeros forward and receive zero gradient.

    Both paddings are mathematically inert (zero weights meet zero inputs / produce zero outputs) and leave the
    parameters, their count and the state dict untouched.  During a no-grad evaluation pass (the rollout) the padded
    weights come from buffers refreshed when the module enters eval mode; with gradients on they are padded on the fly."""

    def _init_padding(self, in_channels: int, n_logits: int, obs_hw: Tuple[int, int], first: nn.Module, head: nn.Module) -> None:
        self.in_channels, self.n_logits, self.obs_hw = int(in_channels), int(n_logits), (int(obs_hw[0]), int(obs_hw[1]))
        self.cin_pad, self.logit_pad = _pad8(in_channels), _pad8(n_logits)
        self._first, self._head = [first], [head]  # lists: not registered twice as submodules
        self.register_buffer("_first_w", None, persistent=False)
        self.register_buffer("_head_w", None, persistent=False)
        self.register_buffer("_head_b", None, persistent=False)
        self._pad_versions = None

    def _padded_now(self):
        first, head = self._first[0], self._head[0]
        w0 = F.pad(first.weight, (0, 0, 0, 0, 0, self.cin_pad - self.in_channels))       # [out, in -> Cp, k, k]
        if isinstance(head, nn.ConvTranspose2d):
            wh = F.pad(head.weight, (0, 0, 0, 0, 0, self.logit_pad - self.n_logits))      # [in, out -> Lp, k, k]
        else:
            wh = F.pad(head.weight, (0, 0, 0, 0, 0, 0, 0, self.logit_pad - self.n_logits))  # [out -> Lp, in, k, k]
        bh = F.pad(head.bias, (0, self.logit_pad - self.n_logits))
        return (w0.contiguous(memory_format=torch.channels_last), wh.contiguous(memory_format=torch.channels_last), bh)

    def _versions(self):
        first, head = self._first[0], self._head[0]
        return (first.weight._version, head.weight._version, head.bias._version, first.weight.data_ptr(), head.weight.data_ptr())

    @torch.no_grad()
    def refresh_padded_weights(self) -> None:
        w0, wh, bh = self._padded_now()
        if self._first_w is None or self._first_w.shape != w0.shape or self._first_w.device != w0.device:
            self._first_w, self._head_w, self._head_b = w0.clone(), wh.clone(), bh.clone()
        else:  # in place: a captured rollout step keeps reading the same addresses
            self._first_w.copy_(w0), self._head_w.copy_(wh), self._head_b.copy_(bh)
        self._pad_versions = self._versions()

    def train(self, mode: bool = True):
        super().train(mode)
        if not mode and self._first[0].weight.is_cuda:
            self.refresh_padded_weights()
        return self

    def _padded_weights(self):
        if self.training or torch.is_grad_enabled():
            return self._padded_now()
        if self._first_w is None or self._pad_versions != self._versions():
            self.refresh_padded_weights()
        return self._first_w, self._head_w, self._head_b

    def packed_obs_shape(self) -> Optional[Tuple[int, int, int]]:
        """(H, W, Cp), the shape of one packed observation; None when it could not be told from a raw (C, H, W) one."""
        H, W = self.obs_hw
        packed = (H, W, self.cin_pad)
        return None if packed == (self.in_channels, H, W) else packed

    def pack_observations(self, obs: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """[B, C, H, W] (any real dtype) -> [B, H, W, Cp] float32, planes C.. zero.  `out` (already zero beyond C) is
        written in place."""
        B, C, H, W = obs.shape
        if out is None:
            out = torch.zeros((B, H, W, self.cin_pad), dtype=torch.float32, device=obs.device)
        out[..., :C].copy_(obs.permute(0, 2, 3, 1))
        return out

    def _packed_input(self, obs: torch.Tensor) -> torch.Tensor:
        """The trunk's input as a channels-last [B, Cp, H, W] view, from packed or raw observations."""
        packed_shape = self.packed_obs_shape()
        if packed_shape is not None and tuple(obs.shape[1:]) == packed_shape:
            packed = obs if obs.dtype == torch.float32 else obs.float()
        else:
            packed = self.pack_observations(obs)
        return packed.permute(0, 3, 1, 2)


class GridEncoderDecoderActorCritic(_PaddedEnds):
    """MicroRTS GridNet: 4 x (conv3x3 + maxpool/2) encoder, 4 x transposed-conv decoder emitting
    [B, H, W, S] logits (physically [B, H, W, pad8(S)], see _PaddedEnds), and an MLP critic on the encoded map."""

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
        # and the decoder output is then physically [B, H, W, pad8(S)] -- the layout the fused loss kernel
        # reads and writes -- so permute(0, 2, 3, 1) is a free view instead of a 245 MB copy per minibatch.
        self.to(memory_format=torch.channels_last)
        self._init_padding(in_channels, n_logits, map_hw, self.encoder[0], self.decoder[-1])
        self.fused_glue = FUSED_GLUE  # False: the PyTorch modules between the convolutions

    def forward(self, obs: torch.Tensor) -> HeadOutputs:
        w0, wh, bh = self._padded_weights()
        x_in = self._packed_input(obs)
        if self.fused_glue and ops.nhwc_glue_supported(x_in):
            # float32 on the device: the convolutions run without their bias, and what sits between them -- bias +
            # max-pool + ReLU in the encoder, bias + ReLU in the decoder -- is one launch each (ops.bias_pool_relu /
            # ops.bias_relu: same arithmetic, bit-identical forward) instead of three / two PyTorch launches
            z = ops.bias_pool_relu(F.conv2d(x_in, w0, None, padding=1), self.encoder[0].bias)
            for i in (3, 6, 9):
                conv = self.encoder[i]
                z = ops.bias_pool_relu(F.conv2d(z, conv.weight, None, padding=1), conv.bias)
            y = z
            for i in (0, 2, 4):
                up = self.decoder[i]
                y = ops.bias_relu(F.conv_transpose2d(y, up.weight, None, stride=2, padding=1, output_padding=1), up.bias)
            # the logit head's bias: one vectorised pass over the [B, H, W, Lp] logits in place (PyTorch: a strided
            # elementwise add forward, a full reduction of the logit gradient backward -- 280 us per C4 minibatch)
            logits = ops.bias_relu(F.conv_transpose2d(y, wh, None, stride=2, padding=1, output_padding=1), bh, relu=False)
        else:
            x = F.conv2d(x_in, w0, self.encoder[0].bias, padding=1)
            z = self.encoder[1:](x)
            y = self.decoder[:-1](z)
            logits = F.conv_transpose2d(y, wh, bh, stride=2, padding=1, output_padding=1)
        logits = logits.permute(0, 2, 3, 1)  # [B, H, W, Lp]
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
        if FUSED_GLUE and ops.act_glue_supported(x):
            # the two convolutions run without their bias; bias + GELU after the first and everything after the
            # second -- bias, squeeze mean, gate, residual sum, output GELU -- are ops.bias_act / ops.se_tail
            conv1, _, conv2, se = self.residual
            y1 = F.conv2d(x, conv1.weight, None, padding=1)
            if y1.dtype == x.dtype:  # (an autocast region fed a float32 map: PyTorch's own promotion rules apply)
                y2 = F.conv2d(ops.bias_act(y1, conv1.bias, "gelu"), conv2.weight, None, padding=1)
                return ops.se_tail(x, y2, conv2.bias, se.fc[0].weight, se.fc[2].weight)
        return self.act(x + self.residual(x))


def _run_glued(seq, x: torch.Tensor) -> torch.Tensor:
    """``seq(x)`` for a Sequential of the squeeze U-net, with every (transposed) convolution -> GELU pair as a bias-free
    convolution + ops.bias_act (one launch for the bias add and the activation, forward and backward)."""
    mods = list(seq)
    i = 0
    while i < len(mods):
        m = mods[i]
        if (FUSED_GLUE and i + 1 < len(mods) and isinstance(mods[i + 1], nn.GELU) and getattr(m, "bias", None) is not None
                and isinstance(m, (nn.Conv2d, nn.ConvTranspose2d)) and x.is_cuda):
            if isinstance(m, nn.Conv2d):
                y = F.conv2d(x, m.weight, None, m.stride, m.padding, m.dilation, m.groups)
            else:
                y = F.conv_transpose2d(x, m.weight, None, m.stride, m.padding, m.output_padding, m.groups, m.dilation)
            if ops.act_glue_supported(y):
                x = ops.bias_act(y, m.bias, "gelu")
            else:
                x = mods[i + 1](y + m.bias.to(y.dtype)[None, :, None, None])
            i += 2
            continue
        x = m(x)
        i += 1
    return x


def _stride_list(s) -> List[int]:
    return [int(v) for v in s] if isinstance(s, (list, tuple)) else [int(s)]


class SqueezeUnetActorCritic(_PaddedEnds):
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
                 increment_kernel_size_on_down_conv: bool = False, obs_range: float = 1.0,
                 obs_hw: Tuple[int, int] = (0, 0)):
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
        self._init_padding(in_channels, n_logits, obs_hw, self.encoders[0][0], self.actor)

    def _values(self, x: torch.Tensor) -> torch.Tensor:
        v = (_run_glued(self.critics[0], x) if self.shared_critic_head
             else torch.cat([_run_glued(c, x) for c in self.critics], dim=1))
        if any(a != "identity" for a in self.critic_activations):  # ChannelwiseActivation: one activation per head
            v = torch.stack([_apply_activation(a, v[:, i]) for i, a in enumerate(self.critic_activations)], dim=1)
        return v.squeeze(-1) if self.n_values == 1 else v

    def forward(self, obs: torch.Tensor) -> HeadOutputs:
        w0, wh, bh = self._padded_weights()
        x = self._packed_input(obs)
        if self.obs_range != 1.0:
            x = x / self.obs_range  # backbone_actor_critic.py:189-192
        stem = self.encoders[0]
        y = F.conv2d(x, w0, None, padding=1) if FUSED_GLUE and x.is_cuda else None
        if y is not None and ops.act_glue_supported(y):
            x = _run_glued(stem[2:], ops.bias_act(y, stem[0].bias, "gelu"))
        else:
            x = stem[1:](F.conv2d(x, w0, stem[0].bias, padding=1))
        skips = [x]
        for enc in list(self.encoders)[1:]:
            x = _run_glued(enc, x)
            skips.append(x)
        x = _run_glued(self.decoders[0], skips[-1])
        for skip, dec in zip(reversed(skips[:-1]), list(self.decoders)[1:]):
            x = _run_glued(dec, skip + x)
        logits = F.conv2d(x, wh, bh, padding=1).permute(0, 2, 3, 1)  # [B, H, W, Lp]
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
