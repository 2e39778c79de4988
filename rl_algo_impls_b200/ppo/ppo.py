"""PPO learner over the fused B200 kernels.

Mirrors ``rl_algo_impls/ppo/ppo.py:106-447``: same constructor keywords (every one is a YAML key of
the reference's hyperparams files), same mutable attributes (callbacks ``setattr`` them between
epochs, hyperparam_transitions.py:19-43), same ``learn`` / ``learn_epoch`` / ``optimizer_step``
and the same ``TrainStats`` handed to ``callback.on_step``.

What differs is the inside of the minibatch loop (ppo.py:290-409).  The reference runs ~60-80
eager torch ops per action head forward, twice that backward, ~20 loss ops and 6-8 host syncs per
minibatch; here one minibatch is

    K3 gather  ->  trunk forward (PyTorch)  ->  K2 moments  ->  K4 fused loss fwd+bwd (one launch:
    dlogits, dvalues, stats)  ->  torch.autograd.backward through the trunk  ->  optimizer step

with the statistics left on the device and read back once per ``learn_epoch``.  Under
``torch.distributed`` (one process per GPU, envs sharded) the advantage moments and the gradients
are all-reduced so that the update equals the single-process update on the concatenated minibatch.
"""
from dataclasses import asdict, dataclass
from time import perf_counter
from typing import Dict, List, NamedTuple, Optional, Sequence, Tuple, TypeVar, Union

import numpy as np
import torch
import torch.distributed as dist
import torch.nn as nn
from torch.optim import Adam

from .. import ops
from ..algorithm import Algorithm, update_learning_rate

NumOrList = Union[float, int, Sequence[float]]


def num_or_array(x):
    """shared/tensor_utils.py: a list becomes a float64 ndarray, a number stays a number."""
    return np.array(x, dtype=np.float64) if isinstance(x, (list, tuple, np.ndarray)) else x


class TrainStepStats(NamedTuple):
    loss: float
    pi_loss: float
    v_loss: np.ndarray
    entropy_loss: float
    approx_kl: float
    clipped_frac: float
    val_clipped_frac: np.ndarray
    additional_losses: Dict[str, float]


@dataclass
class TrainStats:
    loss: float
    pi_loss: float
    v_loss: Union[float, np.ndarray]
    entropy_loss: float
    approx_kl: float
    clipped_frac: float
    val_clipped_frac: Union[float, np.ndarray]
    additional_losses: Dict[str, float]
    explained_var: float
    grad_norm: float

    def __init__(self, step_stats: List[TrainStepStats], explained_var: float, grad_norms: List[float]) -> None:
        self.loss = np.mean([s.loss for s in step_stats]).item()
        self.pi_loss = np.mean([s.pi_loss for s in step_stats]).item()
        self.v_loss = np.mean([s.v_loss for s in step_stats], axis=0)
        self.entropy_loss = np.mean([s.entropy_loss for s in step_stats]).item()
        self.approx_kl = np.mean([s.approx_kl for s in step_stats]).item()
        self.clipped_frac = np.mean([s.clipped_frac for s in step_stats]).item()
        self.val_clipped_frac = np.mean([s.val_clipped_frac for s in step_stats], axis=0)
        self.additional_losses = {
            k: np.mean([s.additional_losses[k] for s in step_stats]).item() for k in step_stats[0].additional_losses
        }
        self.explained_var = explained_var
        self.grad_norm = np.mean(grad_norms).item()

    def write_to_tensorboard(self, tb_writer) -> None:
        for name, value in asdict(self).items():
            if isinstance(value, np.ndarray):
                for idx, v in enumerate(value.flatten()):
                    tb_writer.add_scalar(f"losses/{name}_{idx}", v)
            elif isinstance(value, dict):
                for k, v in value.items():
                    tb_writer.add_scalar(f"losses/{k}", v)
            else:
                tb_writer.add_scalar(f"losses/{name}", value)


PPOSelf = TypeVar("PPOSelf", bound="PPO")


def _world() -> int:
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1



class _FlatGrads:
    """Every trainable parameter's ``.grad`` as a view into ONE flat float32 buffer (same strides as the parameter, so
    channels-last conv weights keep channels-last gradients and the optimizer's multi-tensor path stays on).  The
    backward pass accumulates into the views; the data-parallel all-reduce, the norm and the clip then each touch one
    contiguous buffer in place -- no flatten / unflatten copies and no per-parameter launches (ppo.py:441-447 does
    ``clip_grad_norm_`` + ``zero_grad(set_to_none=True)``; here "none" is a zeroed buffer)."""

    def __init__(self, params: List[nn.Parameter]):
        self.params = params
        self.key = tuple(id(p) for p in params)
        n = sum(p.numel() for p in params)
        self.flat = torch.zeros(n, dtype=params[0].dtype, device=params[0].device)
        self.views: List[torch.Tensor] = []
        off = 0
        for p in params:
            dense = p.is_contiguous() or (p.dim() == 4 and p.is_contiguous(memory_format=torch.channels_last))
            v = (torch.as_strided(self.flat, tuple(p.shape), tuple(p.stride()), off) if dense
                 else self.flat[off:off + p.numel()].view(p.shape))
            self.views.append(v)
            off += p.numel()

    def attach(self) -> None:
        for p, v in zip(self.params, self.views):
            if p.grad is not v:
                p.grad = v

    def zero(self) -> None:
        self.flat.zero_()


class _GraphedUpdate:
    """One minibatch update -- K3 gather, K2 moments, trunk forward, K4 fused loss, trunk backward and
    (unless gradients accumulate) clip + Adam -- captured once per (batch size, hyperparameters, buffer
    addresses) and replayed per minibatch.  Collectives stay outside the captured segments: at world
    size R > 1 the update is three segments with the moments all-reduce and the gradient all-reduce
    between them; at R == 1 (or under gradient accumulation, where a minibatch needs no collective at all)
    it is a single graph."""

    def __init__(self, algo: "PPO", batch, B: int, h: ops.PpoHyper, accumulate: bool, flat: _FlatGrads):
        self.algo, self.B, self.accumulate, self.flat = algo, B, accumulate, flat
        self.world = algo._dp_world()
        dev = algo.device
        self.idx = torch.zeros(B, dtype=torch.int64, device=dev)
        self.params = flat.params
        self.pool = torch.cuda.graph_pool_handle()
        self.graphs: List[torch.cuda.CUDAGraph] = []
        self.kernels_per_replay = 0
        self.stats: Optional[torch.Tensor] = None
        self.grad_norm: Optional[torch.Tensor] = None
        self._mb = self._moments = None
        self.collectives = self.world > 1 and not accumulate

        def seg_a():
            self._mb = batch[self.idx]
            self._moments = algo._moments_local(self._mb.advantages, h)

        def seg_b():
            if not accumulate:
                flat.zero()
            self.stats, _ = algo._minibatch(self._mb, h, None, moments=self._moments)

        def seg_c():
            self.grad_norm = algo._clip_and_step(flat, self.world)

        self.segments = [seg_a, seg_b] + ([] if accumulate else [seg_c])
        self._capture()

    def _between(self, k: int) -> None:
        """Collectives between segment k and k + 1 (eager, on the same stream)."""
        if not self.collectives:
            return
        if k == 0 and self._moments is not None:
            dist.all_reduce(self._moments)
        if k == 1:
            dist.all_reduce(self.flat.flat)

    def _run_eager(self) -> None:
        for k, seg in enumerate(self.segments):
            seg()
            self._between(k)

    def _capture(self) -> None:
        algo, opt = self.algo, self.algo.optimizer
        # the warm-up runs really execute (and, without accumulation, really step the optimizer):
        # snapshot parameters and Adam state, restore them in place afterwards
        saved_p = [p.detach().clone() for p in self.params]
        had_state = {id(p): (p in opt.state and len(opt.state[p]) > 0) for p in self.params}
        saved_s = {id(p): {k: v.detach().clone() for k, v in opt.state[p].items() if torch.is_tensor(v)}
                   for p in self.params if had_state[id(p)]}
        self.flat.attach()
        self.flat.zero()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(2):
                self._run_eager()
        torch.cuda.current_stream().wait_stream(side)
        before = ops.LAUNCHES
        if not self.collectives:
            g = torch.cuda.CUDAGraph()
            with ops.no_gc_during_capture(), torch.cuda.graph(g, pool=self.pool):
                for seg in self.segments:
                    seg()
            self.graphs = [g]
        else:
            for k, seg in enumerate(self.segments):
                g = torch.cuda.CUDAGraph()
                with ops.no_gc_during_capture(), torch.cuda.graph(g, pool=self.pool):
                    seg()
                self.graphs.append(g)
        self.kernels_per_replay = ops.LAUNCHES - before
        with torch.no_grad():
            for p, sp in zip(self.params, saved_p):
                p.copy_(sp)
                st = opt.state.get(p, {})
                if had_state[id(p)]:
                    for k, v in saved_s[id(p)].items():
                        st[k].copy_(v)
                else:
                    for v in st.values():
                        if torch.is_tensor(v):
                            v.zero_()
            self.flat.zero()

    def attach(self) -> None:
        """The captured backward adds into the flat gradient buffer: make sure its views are the parameters' .grad."""
        self.flat.attach()

    def run(self, idx: torch.Tensor):
        self.idx.copy_(idx, non_blocking=True)
        if len(self.graphs) == 1:
            self.graphs[0].replay()
        else:
            for k, g in enumerate(self.graphs):
                g.replay()
                self._between(k)
        ops.LAUNCHES += self.kernels_per_replay
        return self.stats.clone(), (None if self.accumulate else self.grad_norm.clone())


class PPO(Algorithm):
    def __init__(
        self,
        policy,
        device: torch.device,
        tb_writer=None,
        learning_rate: float = 3e-4,
        batch_size: int = 64,
        n_epochs: int = 10,
        gamma: NumOrList = 0.99,
        gae_lambda: NumOrList = 0.95,
        clip_range: float = 0.2,
        clip_range_vf: Optional[float] = None,
        normalize_advantage: bool = True,
        standardize_advantage: bool = False,
        ent_coef: float = 0.0,
        vf_coef: NumOrList = 0.5,
        ppo2_vf_coef_halving: bool = False,
        max_grad_norm: float = 0.5,
        multi_reward_weights: Optional[List[float]] = None,
        gradient_accumulation: bool = False,
        kl_cutoff: Optional[float] = None,
        freeze_policy_head: bool = False,
        freeze_value_head: bool = False,
        freeze_backbone: bool = False,
        switch_range: Optional[int] = None,
        guide_probability: Optional[float] = None,
        normalize_advantages_after_scaling: bool = False,
        autocast_loss: bool = False,
        vf_loss_fn: str = "mse_loss",
        vf_weights: Optional[List[float]] = None,
        teacher_kl_loss_coef: Optional[float] = None,
        teacher_kl_loss_fn=None,
        teacher_loss_importance_sampling: bool = True,
    ) -> None:
        on_cuda = torch.device(device).type == "cuda"
        # Adam(eps=1e-7) as ppo.py:146.  On CUDA the step counter and the learning rate live on the device
        # (capturable=True, tensor lr) so that the whole update can be replayed from a CUDA graph while
        # callbacks keep changing the learning rate between epochs.
        lr = torch.tensor(float(learning_rate), dtype=torch.float32, device=device) if on_cuda else learning_rate
        super().__init__(policy, device, tb_writer, learning_rate,
                         Adam(policy.parameters(), lr=lr, eps=1e-7, capturable=on_cuda))
        self.cuda_graph_update = on_cuda  # replay the minibatch update from CUDA graphs when possible
        self.persistent_dlogits = True  # the fused GridNet loss keeps its dlogits buffer across minibatches (ops.ppo_gridnet_loss)
        self.flat_gradients = on_cuda  # gradients live in one flat buffer (see _FlatGrads) unless freeze_* is active
        self._flat: Optional[_FlatGrads] = None
        self._params_broadcast = False  # data-parallel replicas take rank 0's weights before the first update
        # One process per GPU with torch.distributed initialised = data-parallel learner (envs sharded across the
        # ranks, moments + gradients all-reduced).  False keeps this learner local to its process even then.
        self.data_parallel = True
        self._update_graphs: Dict[tuple, "_GraphedUpdate"] = {}
        self._captures_in_a_row = 0
        self.policy = policy
        self.gamma = num_or_array(gamma)
        self.gae_lambda = num_or_array(gae_lambda)
        self.max_grad_norm = max_grad_norm
        self.clip_range = clip_range
        self.clip_range_vf = clip_range_vf
        self.normalize_advantage = normalize_advantage
        self.standardize_advantage = standardize_advantage
        assert not (normalize_advantage and standardize_advantage), "Cannot both normalize and standardize advantage"
        self.ent_coef = ent_coef
        self.vf_coef = num_or_array(vf_coef)
        self.vf_weights = np.array(vf_weights) if vf_weights is not None else None
        self.ppo2_vf_coef_halving = ppo2_vf_coef_halving
        self.batch_size = batch_size
        self.n_epochs = n_epochs
        self.multi_reward_weights = np.array(multi_reward_weights) if multi_reward_weights else None
        self.gradient_accumulation = gradient_accumulation
        self.kl_cutoff = kl_cutoff
        self.freeze_policy_head = freeze_policy_head
        self.freeze_value_head = freeze_value_head
        self.freeze_backbone = freeze_backbone
        self.switch_range = switch_range
        self.guide_probability = guide_probability
        self.normalize_advantages_after_scaling = normalize_advantages_after_scaling
        self.autocast_loss = autocast_loss
        if vf_loss_fn not in ops.VF_LOSSES:
            raise NotImplementedError(f"vf_loss_fn={vf_loss_fn!r}: the fused kernels implement {sorted(ops.VF_LOSSES)}")
        self.vf_loss_fn = vf_loss_fn
        self.teacher_kl_loss_coef = teacher_kl_loss_coef
        self.teacher_kl_loss_fn = teacher_kl_loss_fn
        self.teacher_loss_importance_sampling = teacher_loss_importance_sampling
        self.last_train_stats: Optional[TrainStats] = None
        self.launches_last_epoch = 0  # libb200rl kernels launched by the last learn_epoch (bench.py reports it)
        self.d2h_bytes_last_epoch = 0
        self.profile_stages = False  # bench.py: CUDA events around rollout / update of the next learn_epoch
        self.stage_ms: Optional[Dict[str, float]] = None
        self._allreduce_events: List[tuple] = []

    # ---------------------------------------------------------------------------------------------
    def learn(self: PPOSelf, train_timesteps: int, rollout_generator, callbacks: Optional[List] = None,
              total_timesteps: Optional[int] = None, start_timesteps: int = 0) -> PPOSelf:
        if total_timesteps is None:
            total_timesteps = train_timesteps
        assert start_timesteps + train_timesteps <= total_timesteps
        timesteps_elapsed = start_timesteps
        while timesteps_elapsed < start_timesteps + train_timesteps:
            timesteps_elapsed, should_continue = self.learn_epoch(timesteps_elapsed, total_timesteps,
                                                                  rollout_generator, callbacks)
            if not should_continue:
                break
        return self

    # ---------------------------------------------------------------------------------------------
    def _hyper(self, V: int, adv_v: int, loss_scale: float, pi_coef: float) -> ops.PpoHyper:
        """Host scalars of this epoch, re-read every learn_epoch (callbacks may have changed them)."""
        vf = np.asarray(self.vf_coef, dtype=np.float64).reshape(-1)
        if vf.size == 1:  # a scalar vf_coef applies to every value head
            vf = np.repeat(vf, V)
        if self.vf_weights is not None:
            # v_loss @ vf_weights -> scalar value loss (ppo.py:344-345): fold the weights into vf_coef
            assert np.ndim(self.vf_coef) == 0 or np.size(self.vf_coef) == 1, "vf_weights needs a scalar vf_coef"
            vf = float(np.asarray(self.vf_coef).reshape(-1)[0]) * np.asarray(self.vf_weights, dtype=np.float64)
        assert vf.size == V, f"vf_coef has {vf.size} entries for {V} value heads"
        if self.normalize_advantages_after_scaling:
            mode = ops.ADV_AFTER_SCALING
        elif self.normalize_advantage:
            mode = ops.ADV_NORMALIZE
        elif self.standardize_advantage:
            mode = ops.ADV_STANDARDIZE
        else:
            mode = ops.ADV_NONE
        w = None if self.multi_reward_weights is None else [float(x) for x in self.multi_reward_weights]
        if adv_v > 1 and w is None:
            raise ValueError("multi-head advantages need multi_reward_weights (ppo.py:317-318)")
        return ops.PpoHyper(clip_range=float(self.clip_range), clip_range_vf=self.clip_range_vf,
                            ent_coef=float(self.ent_coef), vf_coef=[float(x) for x in vf],
                            vf_halving=bool(self.ppo2_vf_coef_halving), pi_coef=pi_coef, loss_scale=loss_scale,
                            adv_mode=mode, adv_weights=w,
                            teacher_kl_coef=float(self.teacher_kl_loss_coef or 0.0),
                            teacher_unbiased=bool(getattr(self.teacher_kl_loss_fn, "unbiased", True)),
                            teacher_importance=bool(self.teacher_loss_importance_sampling),
                            vf_loss=ops.VF_LOSSES[self.vf_loss_fn])

    def _dp_world(self) -> int:
        return _world() if self.data_parallel else 1

    def _moments_local(self, adv: torch.Tensor, h: ops.PpoHyper) -> Optional[torch.Tensor]:
        if h.adv_mode == ops.ADV_NONE:
            return None
        return ops.adv_moments(adv.reshape(adv.shape[0], -1), None, h.adv_mode, h.adv_weights)

    def _moments(self, adv: torch.Tensor, h: ops.PpoHyper) -> Optional[torch.Tensor]:
        """Per-minibatch advantage statistics (ppo.py:307-318).  Data-parallel ranks stepping per minibatch form ONE
        global minibatch out of their local ones: (sum, sumsq, count) are additive, so one all-reduce gives its exact
        statistics.  Under gradient accumulation the minibatches are the reference's own contiguous index ranges
        (shuffle=False) and every rank walks its env shard's ranges, so each local minibatch IS a reference minibatch
        and is normalised by its own statistics, as there: no collective per minibatch."""
        moments = self._moments_local(adv, h)
        if moments is not None and self._dp_world() > 1 and not self.gradient_accumulation:
            dist.all_reduce(moments)
        return moments

    _UNSET = object()

    def _minibatch(self, mb, h: ops.PpoHyper, pi_coef_state: Optional[torch.Tensor], moments=_UNSET) -> Tuple[torch.Tensor, int]:
        """Forward + fused loss + backward of one minibatch.  Returns the device stats vector."""
        obs, old_logp, actions, masks, _, old_values, adv, returns, _additional = (
            mb.obs, mb.logprobs, mb.actions, mb.action_masks, mb.num_actions, mb.values, mb.advantages, mb.returns,
            mb.additional)
        policy = self.policy
        B = obs.shape[0]
        teacher_logp = None
        if self.teacher_kl_loss_coef:
            assert self.teacher_kl_loss_fn is not None, "teacher_kl_loss_coef needs teacher_kl_loss_fn"
            teacher_logp = _additional["teacher_logprobs"]
        if moments is PPO._UNSET:
            moments = self._moments(adv, h)
        kind = getattr(policy, "kind", None)
        fused = kind is not None and hasattr(policy, "head_outputs") and self.kl_cutoff is None
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=bool(self.autocast_loss)):
            if fused:
                out = policy.head_outputs(obs)
            else:
                logp_a, entropy, new_values = policy(obs, actions, action_masks=masks)
        if fused:
            values = out.values
            v32 = values.detach().float().contiguous()
            if kind == "gridnet":
                cells = actions["per_position"] if isinstance(actions, dict) else actions
                pick = actions.get("pick_position") if isinstance(actions, dict) else None
                cmask = masks["per_position"] if isinstance(masks, dict) else masks
                pmask = masks.get("pick_position") if isinstance(masks, dict) else None
                logits = out.pi.detach()
                logits = logits.reshape(B, policy.map_size, logits.shape[-1])
                if not logits.is_contiguous():
                    logits = logits.contiguous()
                # inplace: d loss / d logits lives in a buffer kept across minibatches -- only the rows the previous
                # minibatch wrote are cleared (the gradient is zero outside the unit cells); consumed right below
                res = ops.ppo_gridnet_loss(h, policy.spec, logits, cmask, pmask, cells, pick, old_logp, adv,
                                           old_values, returns, v32, moments=moments, teacher_logp=teacher_logp,
                                           inplace=self.persistent_dlogits)
                grads = [res.grads[0].reshape(out.pi.shape)]
                roots = [out.pi]
            elif kind == "categorical":
                logits = out.pi.detach().float().contiguous()
                res = ops.ppo_categorical_loss(h, logits, masks, actions, old_logp, adv, old_values, returns, v32,
                                               moments=moments, teacher_logp=teacher_logp)
                grads, roots = [res.grads[0].to(out.pi.dtype)], [out.pi]
            else:
                mu = out.pi.detach().float().contiguous()
                res = ops.ppo_gaussian_loss(h, mu, out.log_std.detach().float().contiguous(),
                                            actions.float().contiguous(), old_logp, adv, old_values, returns, v32,
                                            moments=moments, teacher_logp=teacher_logp)
                grads, roots = [res.grads[0].to(out.pi.dtype), res.grads[1]], [out.pi, out.log_std]
            roots.append(values)
            grads.append(res.dvalues.reshape(values.shape).to(values.dtype))
            _backward(roots, grads)
            return res.stats, 1
        # distribution-level path: any policy with forward(obs, actions, masks) -> (logp, entropy, v)
        res = ops.ppo_scalar_loss(h, logp_a.detach().float().contiguous(), entropy.detach().float().contiguous(),
                                  old_logp, adv, old_values, returns, new_values.detach().float().contiguous(),
                                  moments=moments, kl_cutoff=self.kl_cutoff, pi_coef_state=pi_coef_state,
                                  teacher_logp=teacher_logp)
        _backward(
            [logp_a, entropy, new_values],
            [res.grads[0].to(logp_a.dtype).reshape(logp_a.shape), res.grads[1].to(entropy.dtype).reshape(entropy.shape),
             res.dvalues.reshape(new_values.shape).to(new_values.dtype)])
        return res.stats, 1

    def _graphed_update(self, r, h: ops.PpoHyper) -> Optional["_GraphedUpdate"]:
        """The captured update for this rollout, or None when the eager path must run: a policy without
        raw head outputs, the device-side KL cut-off (a data-dependent branch between forward and
        backward), a ragged last minibatch, or buffers / hyperparameters that change on every epoch
        (then re-capturing would cost more than it saves)."""
        if not self.cuda_graph_update or not hasattr(r, "minibatch_indices") or not hasattr(r, "batch"):
            return None
        kind = getattr(self.policy, "kind", None)
        if kind is None or not hasattr(self.policy, "head_outputs") or self.kl_cutoff is not None:
            return None
        if self.freeze_policy_head or self.freeze_value_head or self.freeze_backbone:
            return None  # requires_grad flips between epochs: a captured backward would keep the old graph
        if r.total_steps % self.batch_size != 0:
            return None
        batch = r.batch() if callable(r.batch) else r.batch  # VecRollout builds it lazily, TrajectoryRollout holds it
        _, tensors = batch._flat()
        key = (self.batch_size, bool(self.gradient_accumulation), bool(self.autocast_loss), float(self.max_grad_norm),
               h.clip_range, h.clip_range_vf, h.ent_coef, tuple(h.vf_coef), h.vf_halving, h.loss_scale, h.adv_mode,
               tuple(h.adv_weights) if h.adv_weights is not None else None, h.vf_loss, h.teacher_kl_coef, h.teacher_unbiased,
               h.teacher_importance, tuple(t.data_ptr() for t in tensors),
               tuple(id(p) for p in self.policy.parameters() if p.requires_grad))
        g = self._update_graphs.get(key)
        if g is not None:
            self._captures_in_a_row = 0
            return g
        if self._captures_in_a_row >= 3:  # the key keeps changing (fresh buffers / scheduled hyperparameters)
            return None
        self._captures_in_a_row += 1
        if len(self._update_graphs) >= 4:
            self._update_graphs.pop(next(iter(self._update_graphs)))
        flat = self._flat_grads()
        if flat is None:
            return None
        g = _GraphedUpdate(self, batch, self.batch_size, h, bool(self.gradient_accumulation), flat)
        self._update_graphs[key] = g
        return g

    def learn_epoch(self, timesteps_elapsed: int, total_timesteps: int, rollout_generator,
                    callbacks: Optional[List] = None) -> Tuple[int, bool]:
        start_time = perf_counter()
        launches0 = ops.LAUNCHES
        update_learning_rate(self.optimizer, self.learning_rate)
        if self.switch_range is not None:
            assert hasattr(rollout_generator, "switch_range")
            setattr(rollout_generator, "switch_range", self.switch_range)
        if self.guide_probability is not None:
            assert hasattr(rollout_generator, "guide_probability")
            setattr(rollout_generator, "guide_probability", self.guide_probability)
        if self.tb_writer is not None:
            self._log_chart_scalars(timesteps_elapsed)

        if self.profile_stages:
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
            ev[0].record()
            self._allreduce_events = []
        r = rollout_generator.rollout(gamma=self.gamma, gae_lambda=self.gae_lambda)
        if self.teacher_kl_loss_fn is not None:  # ppo.py:259-262
            r.add_to_batch(self.teacher_kl_loss_fn.add_to_batch, rollout_generator.vec_env.num_envs)
        if self.profile_stages:
            ev[1].record()
        timesteps_elapsed += r.total_steps

        V = int(getattr(r, "value_heads", 1))
        n_mb = r.num_minibatches(self.batch_size)
        loss_scale = 1.0 / n_mb if self.gradient_accumulation else 1.0
        h = self._hyper(V, V, loss_scale, 1.0)
        pi_coef_state = torch.ones(1, dtype=torch.float32, device=self.device) if self.kl_cutoff is not None else None
        if self.freeze_policy_head or self.freeze_value_head or self.freeze_backbone:
            self.policy.freeze(self.freeze_policy_head, self.freeze_value_head, self.freeze_backbone)

        step_stats: List[torch.Tensor] = []
        grad_norms: List[torch.Tensor] = []
        self._broadcast_parameters_once()
        graphed = self._graphed_update(r, h)
        flat = graphed.flat if graphed is not None else self._flat_grads()
        if flat is None:
            self.optimizer.zero_grad(set_to_none=True)  # a flat buffer of an earlier epoch may still be attached
        else:
            flat.attach()
        for _ in range(self.n_epochs):
            step_stats.clear()  # only the last epoch's stats are reported (ppo.py:287-289)
            grad_norms.clear()
            if flat is not None and self.gradient_accumulation:
                flat.zero()
            if graphed is not None:
                for idx in r.minibatch_indices(self.batch_size, shuffle=not self.gradient_accumulation):
                    stats, gn = graphed.run(idx)
                    step_stats.append(stats)
                    if gn is not None:
                        grad_norms.append(gn)
                if self.gradient_accumulation:
                    grad_norms.append(self.optimizer_step_device())
                continue
            for mb in r.minibatches(self.batch_size, shuffle=not self.gradient_accumulation):
                self.policy.reset_noise(self.batch_size)
                if flat is not None and not self.gradient_accumulation:
                    flat.zero()
                stats, _ = self._minibatch(mb, h, pi_coef_state)
                step_stats.append(stats)
                if not self.gradient_accumulation:
                    grad_norms.append(self.optimizer_step_device())
            if self.gradient_accumulation:
                grad_norms.append(self.optimizer_step_device())
        if self.freeze_policy_head or self.freeze_value_head or self.freeze_backbone:
            self.policy.unfreeze()

        if self.profile_stages:
            ev[2].record()
            torch.cuda.synchronize()
            self.stage_ms = {"rollout_and_gae": ev[0].elapsed_time(ev[1]), "update": ev[1].elapsed_time(ev[2])}
            if self._allreduce_events:  # data-parallel: the in-place gradient all-reduce(s) of this learn_epoch
                times = [a.elapsed_time(b) for a, b in self._allreduce_events]
                self.stage_ms["grad_allreduce"] = sum(times)
                self.stage_ms["grad_allreduce_calls"] = len(times)
                self.stage_ms["grad_allreduce_bytes"] = self._flat.flat.numel() * self._flat.flat.element_size()
        # one device -> host read for the whole epoch
        exv = r.explained_variance() if hasattr(r, "explained_variance") else None
        packed = torch.stack(step_stats).double()
        norms = torch.stack(grad_norms).double().reshape(-1)
        tail = torch.cat([norms, exv.reshape(1).double()]) if exv is not None else norms
        host = torch.cat([packed.reshape(-1), tail]).cpu().numpy()
        self.d2h_bytes_last_epoch = host.nbytes
        self.launches_last_epoch = ops.LAUNCHES - launches0
        S = packed.shape[1]
        rows = host[: packed.numel()].reshape(-1, S)
        gn = host[packed.numel(): packed.numel() + norms.numel()]
        if exv is not None:
            explained_var = float(host[-1])
        else:
            var_y = np.var(r.y_true).item()
            explained_var = np.nan if var_y == 0 else 1 - np.var(r.y_true - r.y_pred).item() / var_y
        if self.vf_weights is not None:
            # ppo.py:344-345 contracts the per-head value losses with vf_weights BEFORE the batch mean, so the reported
            # v_loss is their weighted sum (a scalar) and, without value clipping, so is the zero clip fraction (:394)
            vw = np.asarray(self.vf_weights, dtype=np.float64).reshape(-1)
            v_loss_of = lambda x: np.asarray(np.dot(x[5:5 + V], vw), dtype=np.float64)
            vclip_of = lambda x: _vec(x[5 + V:5 + 2 * V], V) if self.clip_range_vf is not None else np.zeros(())
        else:
            v_loss_of = lambda x: _vec(x[5:5 + V], V)
            vclip_of = lambda x: _vec(x[5 + V:5 + 2 * V], V)
        steps = [
            TrainStepStats(float(x[0]), float(x[1]), v_loss_of(x), float(x[2]), float(x[3]), float(x[4]), vclip_of(x),
                           {"teacher_kl_loss": float(x[5 + 2 * V])} if self.teacher_kl_loss_coef else {})
            for x in rows
        ]
        train_stats = TrainStats(steps, explained_var, [float(g) for g in gn])
        self.last_train_stats = train_stats
        rollout_steps = r.total_steps
        if self.tb_writer is not None:
            train_stats.write_to_tensorboard(self.tb_writer)
            self.tb_writer.add_scalar("train/steps_per_second", rollout_steps / (perf_counter() - start_time))
            if hasattr(self.tb_writer, "on_steps"):
                self.tb_writer.on_steps(rollout_steps)
        if callbacks:
            if not all(c.on_step(timesteps_elapsed=rollout_steps, train_stats=train_stats) for c in callbacks):
                return timesteps_elapsed, False
        return timesteps_elapsed, True

    # ---------------------------------------------------------------------------------------------
    def _flat_grads(self) -> Optional[_FlatGrads]:
        """The flat gradient buffer of the trainable parameters, or None when the per-parameter path must run: a CPU
        policy, ``flat_gradients`` off, or freeze_* in force (a frozen parameter must keep ``grad is None`` so that
        Adam skips it, exactly like the reference's zero_grad(set_to_none=True))."""
        if not self.flat_gradients or self.freeze_policy_head or self.freeze_value_head or self.freeze_backbone:
            if self._flat is not None:
                for p in self._flat.params:
                    p.grad = None
                self._flat = None
            return None
        params = [p for p in self.policy.parameters() if p.requires_grad]
        if not params or not params[0].is_cuda or any(p.dtype != params[0].dtype for p in params):
            return None
        if self._flat is None or self._flat.key != tuple(id(p) for p in params):
            self._flat = _FlatGrads(params)
        return self._flat

    def _broadcast_parameters_once(self) -> None:
        """Data-parallel replicas (one process per GPU) must start from the same weights: rank 0's."""
        if self._params_broadcast or self._dp_world() == 1:
            return
        for t in list(self.policy.parameters()) + list(self.policy.buffers()):
            dist.broadcast(t.data, 0)
        self._params_broadcast = True

    def _sync_grads(self, params: List[nn.Parameter]) -> None:
        """Data-parallel ranks (envs sharded), per-parameter path: one all-reduce of the flattened gradients, then / R,
        issued before the clip so that clipping sees the global gradient (ppo.py:441-447)."""
        world = self._dp_world()
        if world == 1:
            return
        grads = [p.grad for p in params]
        flat = torch._utils._flatten_dense_tensors(grads)
        dist.all_reduce(flat)
        flat.div_(world)
        for g, f in zip(grads, torch._utils._unflatten_dense_tensors(flat, grads)):
            g.copy_(f)

    def _clip_and_step(self, flat: _FlatGrads, world: int) -> torch.Tensor:
        """clip_grad_norm_ + Adam.step over the flat buffer, which holds the SUM of the ranks' gradients (already
        all-reduced): the mean's norm is |sum| / R, and the / R is folded into the clip coefficient -- one norm and one
        scale over one contiguous buffer (torch.nn.utils.clip_grad_norm_: coef = min(1, max_norm / (norm + 1e-6)))."""
        norm = torch.linalg.vector_norm(flat.flat)
        if world > 1:
            norm = norm / world
        coef = torch.clamp(self.max_grad_norm / (norm + 1e-6), max=1.0)
        flat.flat.mul_(coef / world if world > 1 else coef)
        self.optimizer.step()
        return norm.detach()

    def optimizer_step_device(self) -> torch.Tensor:
        flat = self._flat
        if flat is not None and all(p.grad is v for p, v in zip(flat.params, flat.views)):
            world = self._dp_world()
            if world > 1:  # in place, no flatten / unflatten copies; before the clip, so clipping sees the global gradient
                if self.profile_stages:
                    ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
                    ev[0].record()
                    dist.all_reduce(flat.flat)
                    ev[1].record()
                    self._allreduce_events.append(ev)
                else:
                    dist.all_reduce(flat.flat)
            return self._clip_and_step(flat, world)
        params = [p for p in self.policy.parameters() if p.grad is not None]
        self._sync_grads(params)
        grad_norm = nn.utils.clip_grad_norm_(params, self.max_grad_norm)
        self.optimizer.step()
        self.optimizer.zero_grad(set_to_none=True)
        return grad_norm.detach()

    def optimizer_step(self) -> float:
        return self.optimizer_step_device().item()

    def _log_chart_scalars(self, timesteps_elapsed: int) -> None:
        scalars = {"learning_rate": self.optimizer.param_groups[0]["lr"], "ent_coef": self.ent_coef,
                   "pi_clip": self.clip_range, "gamma": self.gamma, "gae_lambda": self.gae_lambda,
                   "vf_coef": self.vf_coef}
        if self.clip_range_vf is not None:
            scalars["v_clip"] = self.clip_range_vf
        if self.multi_reward_weights is not None:
            scalars["reward_weights"] = self.multi_reward_weights
        for name, value in scalars.items():
            if isinstance(value, np.ndarray):
                for i, v in enumerate(value.reshape(-1)):
                    self.tb_writer.add_scalar(f"charts/{name}_{i}", float(v), timesteps_elapsed)
            else:
                self.tb_writer.add_scalar(f"charts/{name}", float(value), timesteps_elapsed)


def _backward(roots, grads) -> None:
    """autograd.backward through the outputs that still lead to a trainable parameter (freeze_* can cut some off)."""
    live = [(r, g) for r, g in zip(roots, grads) if r.requires_grad]
    if live:
        torch.autograd.backward([r for r, _ in live], [g for _, g in live])


def _vec(x: np.ndarray, V: int):
    return np.asarray(x, dtype=np.float64) if V > 1 else np.asarray(x[0], dtype=np.float64)
