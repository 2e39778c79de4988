"""SyncStepRolloutGenerator: the rollout buffer and step loop, resident in HBM.

Mirrors ``rl_algo_impls/rollout/sync_step_rollout.py:14-294`` (constructor keywords, ``prepare``,
``rollout(gamma, gae_lambda) -> VecRollout``, the masked-reset options).  What changes is where
the buffers live: the reference keeps host numpy ``[T, N, ...]`` arrays (:98-131) and round-trips
every policy step through H2D / D2H copies (actor_critic.py:306-318); here the arrays are CUDA
tensors allocated once, ``policy.step_device`` samples and evaluates log-probs in one launch and
its outputs are written straight into row ``s`` of the buffers.  A host ``VectorEnv`` (numpy in,
numpy out -- the reference's contract) still works: its observations / masks are uploaded through
pinned staging buffers and the sampled actions are downloaded for ``env.step``; a device env
(CUDA tensors in / out) never touches PCIe.
"""
from typing import Dict, Optional

import numpy as np
import torch

from .. import _lib
from .rollout import RolloutGenerator
from .vec_rollout import VecRollout

_TORCH_DTYPES = {"float32": torch.float32, "float64": torch.float32, "uint8": torch.uint8, "int64": torch.int64,
                 "bool": torch.bool, "int32": torch.int32, "float16": torch.float16}


def _torch_dtype(np_dtype) -> torch.dtype:
    return _TORCH_DTYPES[np.dtype(np_dtype).name]


class _Uploader:
    """numpy -> CUDA.  Large arrays that keep coming from the same host buffer (a vec env that reuses its output
    arrays, or cycles through a pool of them) are page-locked IN PLACE with cudaHostRegister the second time
    their buffer is seen, so the H2D copy is one DMA straight out of the env's memory; everything else goes
    through a reusable pinned staging buffer (one per distinct field), which also does dtype conversion."""

    REGISTER_MIN_BYTES = 256 << 10   # below this a staging memcpy is cheaper than bookkeeping
    REGISTER_MAX_BYTES = 8 << 30     # total page-locked in place

    def __init__(self, device: torch.device):
        self.device = device
        self._pinned: Dict[str, torch.Tensor] = {}
        self._done: Dict[str, torch.cuda.Event] = {}
        self._seen: Dict[tuple, int] = {}        # (root pointer, nbytes) -> sightings before registration
        self._registered: Dict[tuple, bool] = {}  # (root pointer, nbytes) -> cudaHostRegister succeeded
        self._registered_bytes = 0
        self._roots: list = []  # the page-locked arrays, kept alive while they are registered
        self.bytes = 0  # host -> device bytes moved so far
        self._fast: Dict[tuple, tuple] = {}     # (field, host pointer) -> (nbytes, device pointer, shape, dtype): a plain DMA
        self._batches: Dict[tuple, tuple] = {}  # host pointers of one step's DMAs -> prebuilt argument arrays

    def _page_locked(self, a: np.ndarray) -> bool:
        """True when `a` lives in a host buffer this uploader has page-locked (registering it on its second sighting)."""
        root = a
        while isinstance(root.base, np.ndarray):
            root = root.base
        if root.base is not None or not root.flags.c_contiguous:  # memory owned by something else (mmap, torch, ...)
            return False
        key = (root.ctypes.data, root.nbytes)
        state = self._registered.get(key)
        if state is not None:
            return state
        self._seen[key] = self._seen.get(key, 0) + 1
        if self._seen[key] < 2 or self._registered_bytes + root.nbytes > self.REGISTER_MAX_BYTES:
            if len(self._seen) > 4096:  # an env that allocates fresh arrays every step: stop tracking
                self._seen.clear()
            return False
        ok = _lib.lib().b200rl_host_register(root.ctypes.data, root.nbytes) == 0  # a refusal leaves no CUDA error behind
        self._registered[key] = ok
        if ok:
            self._registered_bytes += root.nbytes
            self._roots.append(root)
        return ok

    def close(self) -> None:
        """Release the page locks (before the arrays can be freed: a stale registration would make a later
        registration of recycled memory fail)."""
        for (ptr, _), ok in self._registered.items():
            if ok:
                _lib.lib().b200rl_host_unregister(ptr)
        self._registered.clear()
        self._roots = []
        self._fast.clear(), self._batches.clear()

    def __del__(self):
        try:
            self.close()
        except Exception:  # interpreter shutdown: the library / driver may already be gone
            pass

    def upload_step(self, items) -> None:
        """One env step's uploads, ``items = [(field, host array or tensor, device tensor)]``.  Fields that arrive in a
        page-locked buffer seen before (an env that reuses its output arrays, or cycles through a pool) are enqueued by
        ONE C call (b200rl_h2d_batch) with cached arguments: ~4 us of host time for the step instead of ~15 us per
        field through tensor wrappers; the rest takes the per-field path."""
        import ctypes as C

        ptrs = []
        for name, src, dst in items:
            entry = None
            if type(src) is np.ndarray and src.flags.c_contiguous:
                ptr = src.__array_interface__["data"][0]
                entry = self._fast.get((name, ptr))
                if entry is not None and (entry[1] != dst.data_ptr() or entry[2] != src.shape or entry[3] != src.dtype):
                    entry = None  # the buffer at this address is another array now
                if entry is None and src.nbytes >= self.REGISTER_MIN_BYTES and src.shape == tuple(dst.shape) \
                        and torch.from_numpy(src).dtype == dst.dtype and dst.is_contiguous() and self._registered.get(self._root_key(src)):
                    entry = self._fast[(name, ptr)] = (src.nbytes, dst.data_ptr(), src.shape, src.dtype)
                    if len(self._fast) > 4096:
                        self._fast.clear(), self._batches.clear()
            if entry is None:
                self(name, src, dst)
            else:
                ptrs.append((ptr, entry))
        if not ptrs:
            return
        key = tuple((p, e[1]) for p, e in ptrs)
        batch = self._batches.get(key)
        if batch is None:
            n = len(ptrs)
            dst_arr, src_arr, nb_arr = (C.c_void_p * n)(), (C.c_void_p * n)(), (C.c_int64 * n)()
            for i, (p, e) in enumerate(ptrs):
                dst_arr[i], src_arr[i], nb_arr[i] = e[1], p, e[0]
            batch = self._batches[key] = (n, dst_arr, src_arr, nb_arr, sum(e[0] for _, e in ptrs))
        rc = _lib.lib().b200rl_h2d_batch(batch[0], batch[1], batch[2], batch[3], torch.cuda.current_stream().cuda_stream)
        _lib.check(rc, "b200rl_h2d_batch")
        self.bytes += batch[4]

    @staticmethod
    def _root_key(a: np.ndarray) -> tuple:
        root = a
        while isinstance(root.base, np.ndarray):
            root = root.base
        return (root.ctypes.data, root.nbytes)

    def __call__(self, name: str, src, dst: torch.Tensor) -> None:
        if isinstance(src, torch.Tensor):
            dst.copy_(src, non_blocking=True)
            return
        a = np.ascontiguousarray(src)
        if a.nbytes >= self.REGISTER_MIN_BYTES and a.shape == tuple(dst.shape):
            t = torch.from_numpy(a)
            if t.dtype == dst.dtype and self._page_locked(a):
                # DMA out of the env's own buffer.  The caller must not overwrite it before the copy has run:
                # the step loops read the sampled actions back (a stream synchronisation) before the next env.step.
                self.bytes += a.nbytes
                dst.copy_(t, non_blocking=True)
                return
        stage = self._pinned.get(name)
        if stage is None or stage.shape != a.shape or stage.dtype != dst.dtype:
            stage = torch.empty(a.shape, dtype=dst.dtype, pin_memory=True)
            self._pinned[name] = stage
            self._done[name] = torch.cuda.Event()
        else:
            self._done[name].synchronize()  # the previous upload out of this staging buffer has landed
        self.bytes += stage.numel() * stage.element_size()
        stage.copy_(torch.from_numpy(a))  # host-side convert (e.g. float64 obs -> float32) + copy
        dst.copy_(stage, non_blocking=True)
        self._done[name].record()


class SyncStepRolloutGenerator(RolloutGenerator):
    def __init__(
        self,
        policy,
        vec_env,
        n_steps: int = 2048,
        sde_sample_freq: int = -1,
        scale_advantage_by_values_accuracy: bool = False,
        full_batch_off_accelerator: bool = False,
        include_logp: bool = True,
        subaction_mask: Optional[Dict[int, Dict[int, int]]] = None,
        num_envs_reset_every_rollout: int = 0,
        rolling_num_envs_reset_every_rollout: int = 0,
        random_num_envs_reset_every_rollout: int = 0,
        prepare_steps: int = 0,
        rolling_num_envs_reset_every_prepare_step: int = 0,
        cuda_graph: bool = True,
    ) -> None:
        super().__init__(policy, vec_env)
        self.cuda_graph = cuda_graph
        self.include_num_actions = False  # A2C's scale_loss_by_num_actions turns the Batch.num_actions field on
        self.n_steps = int(n_steps)
        self.sde_sample_freq = sde_sample_freq
        self.scale_advantage_by_values_accuracy = scale_advantage_by_values_accuracy
        self.full_batch_off_accelerator = full_batch_off_accelerator
        self.include_logp = include_logp
        self.subaction_mask = subaction_mask
        for name, v in (("num_envs_reset_every_rollout", num_envs_reset_every_rollout),
                        ("rolling_num_envs_reset_every_rollout", rolling_num_envs_reset_every_rollout),
                        ("random_num_envs_reset_every_rollout", random_num_envs_reset_every_rollout),
                        ("rolling_num_envs_reset_every_prepare_step", rolling_num_envs_reset_every_prepare_step)):
            assert v % 2 == 0, f"{name} must be even, got {v}"
        self.num_envs_reset_every_rollout = num_envs_reset_every_rollout
        self.rolling_num_envs_reset_every_rollout = rolling_num_envs_reset_every_rollout
        self.random_num_envs_reset_every_rollout = random_num_envs_reset_every_rollout
        self.prepare_steps = prepare_steps
        self.rolling_num_envs_reset_every_prepare_step = rolling_num_envs_reset_every_prepare_step
        N = vec_env.num_envs
        assert N > (num_envs_reset_every_rollout + rolling_num_envs_reset_every_rollout
                    + random_num_envs_reset_every_rollout), "more envs reset per rollout than envs"
        self.rolling_mask_idx = 0
        self.rolling_reset_indexes = np.random.permutation(N // 2)

        self.device = torch.device(policy.device)
        if self.device.type != "cuda":
            raise RuntimeError("SyncStepRolloutGenerator keeps the rollout in HBM: the policy must be on a CUDA device")
        self._upload = _Uploader(self.device)
        # env steps taken so far, on the device: row index of the buffer write (mod n_steps) and RNG
        # offset of the sampling kernel, both read inside captured launches
        self.step_count = torch.zeros(1, dtype=torch.int64, device=self.device)
        self._k0_ticket = torch.zeros(1, dtype=torch.int32, device=self.device)  # K0's "last CTA advances the step" counter
        self._rollouts_done = 0
        self._graph: Optional[torch.cuda.CUDAGraph] = None
        self._graph_outputs = None
        self._kernels_per_replay = 0
        self.d2h_bytes = 0  # device -> host bytes (sampled actions handed to a host env)
        self._host_actions = None  # pinned landing zone of the sampled actions (host env)
        self._wide_actions: Optional[torch.Tensor] = None  # GridNet + host env: int64 per-cell actions, written by K5
        self._host_wide = None   # ... and their two pinned landing zones (alternating)
        self._host_parity = 0
        self._host_rewards: Optional[torch.Tensor] = None  # pinned [T, N(, V)]: a host env's rewards, uploaded per rollout
        self._host_starts: Optional[torch.Tensor] = None   # pinned [T + 1, N]: episode-start flags, row T carries over
        self.get_action_mask = getattr(vec_env, "get_action_mask", None)

        T = self.n_steps
        dev = self.device
        obs_space = vec_env.single_observation_space
        value_shape = tuple(policy.value_shape)
        act_shape = policy.action_shape
        # Observations are stored in the layout the trunk consumes without a copy when the policy names one
        # (ActorCritic.packed_obs_shape: channels last, planes padded to a multiple of 8): packed once per env step,
        # so that the minibatch gather hands the learner trunk-ready rows.  Otherwise the env's own layout / dtype.
        self._packed = getattr(policy, "packed_obs_shape", None) if len(tuple(obs_space.shape)) == 3 else None
        obs_shape = tuple(self._packed) if self._packed else tuple(obs_space.shape)
        obs_dtype = torch.float32 if self._packed else _torch_dtype(obs_space.dtype)
        self._raw_obs = (torch.zeros((N,) + tuple(obs_space.shape), dtype=_torch_dtype(obs_space.dtype), device=dev)
                         if self._packed else None)  # landing zone of a host env's upload, packed from there
        # host env: _raw_obs is the source of truth and every policy step starts by packing it (inside the captured
        # step: no launch of its own between the upload and the replay)
        self._pack_in_step = bool(self._packed) and getattr(vec_env, "device", None) is None
        self.obs = torch.zeros((T, N) + obs_shape, dtype=obs_dtype, device=dev)
        self.rewards = torch.zeros((T, N) + value_shape, dtype=torch.float32, device=dev)
        self.episode_starts = torch.zeros((T, N), dtype=torch.bool, device=dev)
        self.values = torch.zeros((T, N) + value_shape, dtype=torch.float32, device=dev)
        self.logprobs = torch.zeros((T, N), dtype=torch.float32, device=dev) if include_logp else None
        # GAE outputs live in fixed buffers too: a captured update graph gathers from the same addresses every rollout
        self.advantages = torch.zeros((T, N) + value_shape, dtype=torch.float32, device=dev)
        self.returns = torch.zeros((T, N) + value_shape, dtype=torch.float32, device=dev)
        self.next_episode_starts = torch.ones((N,), dtype=torch.bool, device=dev)
        self.next_obs = torch.zeros((N,) + obs_shape, dtype=self.obs.dtype, device=dev)

        kind = getattr(policy, "kind", None)
        if isinstance(act_shape, dict):
            self.actions = {
                k: torch.zeros((T, N) + tuple(s), dtype=torch.uint8 if k == "per_position" else torch.int64, device=dev)
                for k, s in act_shape.items()
            }
        else:
            adt = torch.uint8 if kind == "gridnet" else (torch.float32 if kind == "gaussian" else torch.int64)
            self.actions = torch.zeros((T, N) + tuple(act_shape), dtype=adt, device=dev)

        if (kind == "gridnet" and getattr(vec_env, "device", None) is None
                and getattr(policy, "supports_wide_actions", False)):
            cells = act_shape["per_position"] if isinstance(act_shape, dict) else act_shape
            self._wide_actions = torch.zeros((N,) + tuple(cells), dtype=torch.int64, device=dev)

        first_obs, _ = vec_env.reset()
        self._set_next_obs(first_obs)
        self.action_masks = None
        self.next_action_masks = None
        if self.get_action_mask is not None:
            m = self.get_action_mask()
            if m is not None:
                if isinstance(m, dict):
                    self.action_masks = {k: torch.zeros((T,) + tuple(v.shape), dtype=torch.bool, device=dev) for k, v in m.items()}
                    self.next_action_masks = {k: torch.zeros(tuple(v.shape), dtype=torch.bool, device=dev) for k, v in m.items()}
                else:
                    self.action_masks = torch.zeros((T,) + tuple(m.shape), dtype=torch.bool, device=dev)
                    self.next_action_masks = torch.zeros(tuple(m.shape), dtype=torch.bool, device=dev)
                self._upload_masks(m)

    # -- helpers -------------------------------------------------------------------------------
    def _set_next_obs(self, obs) -> None:
        """The env's observation for the next step -> self.next_obs (uploaded if it is a host array, packed into the
        trunk's layout if the policy names one)."""
        if not self._packed:
            self._upload("obs", obs, self.next_obs)
            return
        if not isinstance(obs, torch.Tensor):
            self._upload("obs", obs, self._raw_obs)
            obs = self._raw_obs
            if self._pack_in_step:  # the captured policy step packs _raw_obs itself (its first launch)
                return
        elif self._pack_in_step:
            self._raw_obs.copy_(obs)
            return
        self.policy.pack_observations(obs, out=self.next_obs)

    def _pack_raw(self) -> None:
        """Host env, packed observations: self.next_obs <- pack(self._raw_obs), the landing zone of the env's upload.
        Idempotent; the first launch of every policy step, and called before anything else reads next_obs."""
        if self._pack_in_step:
            self.policy.pack_observations(self._raw_obs, out=self.next_obs)

    def _upload_env_outputs(self, next_obs, masks) -> None:
        """A host env's observation and masks for the next step, enqueued together (one C call when they come out of
        page-locked buffers)."""
        if isinstance(next_obs, torch.Tensor) or (self._packed and not self._pack_in_step):
            self._set_next_obs(next_obs)
            items = []
        else:
            items = [("obs", next_obs, self._raw_obs if self._packed else self.next_obs)]
        if masks is not None:
            if isinstance(masks, dict):
                items += [("mask_" + k, v, self.next_action_masks[k]) for k, v in masks.items()]
            else:
                items.append(("mask", masks, self.next_action_masks))
        self._upload.upload_step(items)

    def _upload_masks(self, m) -> None:
        if isinstance(m, dict):
            for k, v in m.items():
                self._upload("mask_" + k, v, self.next_action_masks[k])
        else:
            self._upload("mask", m, self.next_action_masks)

    def _download_wide(self, a) -> None:
        """Start the int64 per-cell actions (and the pick actions) towards the pinned buffer of this step's parity.  The
        arrays handed to the env are views of it: valid until the policy step after the next one."""
        self._host_parity ^= 1
        if self._host_wide is None:
            pin = lambda t: torch.empty(tuple(t.shape), dtype=t.dtype, pin_memory=True)
            self._host_wide = [{"per_position": pin(self._wide_actions),
                                **({"pick_position": pin(a["pick_position"])} if isinstance(a, dict) else {})}
                               for _ in range(2)]
        buf = self._host_wide[self._host_parity]
        buf["per_position"].copy_(self._wide_actions, non_blocking=True)
        if isinstance(a, dict):
            buf["pick_position"].copy_(a["pick_position"], non_blocking=True)

    def _env_actions(self, a, landed: bool = False):
        """What vec_env.step receives: CUDA tensors for a device env, numpy (int64 / f32) for a host env.
        ``landed``: the step already copied `a` into the pinned ``_host_actions`` (wait for it, read it there)."""
        if getattr(self.vec_env, "device", None) is not None:
            return a
        from ..policy.actor_critic import clamp_actions

        if landed and self._wide_actions is not None:
            torch.cuda.current_stream().synchronize()
            buf = self._host_wide[self._host_parity]
            self.d2h_bytes += sum(t.numel() * t.element_size() for t in buf.values())
            if isinstance(a, dict):
                return {k: t.numpy().reshape(tuple(a[k].shape)) for k, t in buf.items()}
            return buf["per_position"].numpy().reshape(tuple(a.shape))

        if landed:
            torch.cuda.current_stream().synchronize()
            fetch = lambda k, t: (self._host_actions[k] if k is not None else self._host_actions).numpy()
        else:
            fetch = lambda k, t: t.cpu().numpy()
        if isinstance(a, dict):
            self.d2h_bytes += sum(t.numel() * t.element_size() for t in a.values())
            return {k: fetch(k, t).astype(np.int64) for k, t in a.items()}
        self.d2h_bytes += a.numel() * a.element_size()
        a_np = fetch(None, a)
        a_np = a_np.astype(np.int64) if a_np.dtype == np.uint8 else (a_np.copy() if landed else a_np)
        return clamp_actions(a_np, self.vec_env.single_action_space, getattr(self.policy, "squash_output", False))

    def prepare(self) -> None:
        if not self.prepare_steps:
            return
        for _ in range(0, self.prepare_steps, self.n_steps):
            self._rollout(output_next_values=False)
            self._reset_envs(0, self.rolling_num_envs_reset_every_prepare_step, 0)

    # -- the step loop (sync_step_rollout.py:181-216) ------------------------------------------------
    def _fields(self, a, v, logp, rewards=None, with_starts: bool = True):
        """(this step's slices, their [T, ...] buffers) in one list pair for the K0 store."""
        src, dst = [self.next_obs], [self.obs]
        if with_starts:
            src.append(self.next_episode_starts), dst.append(self.episode_starts)
        if self.action_masks is not None:
            if isinstance(self.action_masks, dict):
                for k, buf in self.action_masks.items():
                    src.append(self.next_action_masks[k]), dst.append(buf)
            else:
                src.append(self.next_action_masks), dst.append(self.action_masks)
        if isinstance(self.actions, dict):
            for k, buf in self.actions.items():
                src.append(a[k].contiguous()), dst.append(buf)
        else:
            src.append(a.contiguous()), dst.append(self.actions)
        src.append(v.contiguous()), dst.append(self.values)
        if self.logprobs is not None:
            src.append(logp.contiguous()), dst.append(self.logprobs)
        if rewards is not None:
            src.append(rewards.contiguous()), dst.append(self.rewards)
        return src, dst

    def _policy_step(self):
        """Host-env half step: sample + evaluate on the current next_obs / masks, write the pre-env fields of
        this step into row (step_count % T) of the buffers and start the sampled actions towards pinned host
        memory.  Static addresses only: graph-capturable (the D2H copy becomes a memcpy node).  Rewards and
        episode-start flags never take part: a host env hands them over on the host, where they are collected
        in pinned [T, N] arrays and uploaded once per rollout."""
        from .. import ops

        self._pack_raw()
        if self._wide_actions is not None:
            # GridNet: the sampling kernel also writes the per-cell actions as int64, the dtype the env is handed; they are
            # downloaded after the replay into alternating pinned buffers (_download_wide): no cast and no copy on the host
            a, v, logp = self.policy.step_device(self.next_obs, self.next_action_masks, offset_dev=self.step_count,
                                                 wide_out=self._wide_actions)
            src, dst = self._fields(a, v, logp, with_starts=False)
            ops.rollout_store_step(src, dst, self.step_count, advance_ticket=self._k0_ticket)  # ... and step_count += 1
            return a
        a, v, logp = self.policy.step_device(self.next_obs, self.next_action_masks, offset_dev=self.step_count)
        src, dst = self._fields(a, v, logp, with_starts=False)
        ops.rollout_store_step(src, dst, self.step_count, advance_ticket=self._k0_ticket)  # ... and step_count += 1
        if self._host_actions is None:
            pin = lambda t: torch.empty(tuple(t.shape), dtype=t.dtype, pin_memory=True)
            self._host_actions = {k: pin(t) for k, t in a.items()} if isinstance(a, dict) else pin(a)
        if isinstance(a, dict):
            for k, t in a.items():
                self._host_actions[k].copy_(t, non_blocking=True)
        else:
            self._host_actions.copy_(a, non_blocking=True)
        return a

    def _device_env_step(self):
        """One whole env step with a device env: policy, env, buffer write, carry-over.  No host
        synchronisation and no host-dependent address: one CUDA-graph replay per env step."""
        from .. import ops

        a, v, logp = self.policy.step_device(self.next_obs, self.next_action_masks, offset_dev=self.step_count)
        next_obs, rewards, terminations, truncations, _ = self.vec_env.step(a)
        src, dst = self._fields(a, v, logp, rewards.reshape(self.rewards.shape[1:]))
        # Everything the reference does after env.step (sync_step_rollout.py:202-212) rides on the same K0 launch: the
        # env's outputs for the next step replace next_obs / next masks (a raw float32 [N, C, H, W] observation is
        # transposed into the packed layout on the way), next_episode_starts becomes terminations | truncations, and
        # the step counter advances -- no launch of its own for any of them.
        same = lambda new, cur: (isinstance(new, torch.Tensor) and new.is_cuda and new.is_contiguous()
                                 and new.data_ptr() != cur.data_ptr())
        fits = lambda new, cur: same(new, cur) and new.dtype == cur.dtype and new.shape == cur.shape
        carry, carry_or, pack = {}, {}, None
        if fits(next_obs, self.next_obs):
            carry[id(self.next_obs)] = next_obs
        elif (self._packed and same(next_obs, self.next_obs) and next_obs.dtype == torch.float32 and next_obs.dim() == 4
              and self.next_obs.shape[-1] <= 128
              and tuple(self.next_obs.shape) == (next_obs.shape[0], next_obs.shape[2], next_obs.shape[3], self.next_obs.shape[-1])):
            N, Cc, H, W = next_obs.shape
            carry[id(self.next_obs)] = next_obs
            pack = (0, N, Cc, H * W, self.next_obs.shape[-1])  # _fields puts the observation first
        masks = self.get_action_mask() if self.next_action_masks is not None else None
        pairs = []
        if masks is not None:
            pairs = ([(self.next_action_masks[k], masks[k]) for k in self.next_action_masks]
                     if isinstance(self.next_action_masks, dict) else [(self.next_action_masks, masks)])
            carry.update({id(cur): new for cur, new in pairs if fits(new, cur)})
        starts = self.next_episode_starts
        if fits(terminations, starts) and fits(truncations, starts):
            carry[id(starts)], carry_or[id(starts)] = terminations, truncations
        ops.rollout_store_step(src, dst, self.step_count, carry=[carry.get(id(t)) for t in src],
                               carry_or=[carry_or.get(id(t)) for t in src], pack=pack, advance_ticket=self._k0_ticket)
        if id(self.next_obs) not in carry:
            self._set_next_obs(next_obs)
        if id(starts) not in carry:
            torch.logical_or(terminations, truncations, out=starts)
        for cur, new in pairs:
            if id(cur) not in carry:
                cur.copy_(new)

    def _capture(self, fn):
        """Warm up on a side stream, then capture `fn` once (torch.cuda.graph)."""
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(3):
                out = fn()
        torch.cuda.current_stream().wait_stream(side)
        from .. import ops

        graph = torch.cuda.CUDAGraph()
        before = ops.LAUNCHES
        with ops.no_gc_during_capture(), torch.cuda.graph(graph):
            out = fn()
        self._kernels_per_replay = ops.LAUNCHES - before  # libb200rl kernels inside one replay
        return graph, out

    def _rollout(self, output_next_values: bool) -> Optional[torch.Tensor]:
        from .. import ops

        self.policy.eval()
        self.policy.reset_noise()
        device_env = getattr(self.vec_env, "device", None) is not None
        T = self.n_steps
        if self.cuda_graph and self._graph is None:
            with torch.no_grad():
                if device_env:
                    self._graph, _ = self._capture(self._device_env_step)
                else:
                    self._graph, self._graph_outputs = self._capture(self._policy_step)
        self.step_count.fill_(self._rollouts_done * T)  # row 0 of the buffer, fresh RNG offsets
        self._rollouts_done += 1
        if not device_env:
            if self._host_rewards is None:
                self._host_rewards = torch.zeros(tuple(self.rewards.shape), dtype=torch.float32, pin_memory=True)
                self._host_starts = torch.ones((T + 1, self.rewards.shape[1]), dtype=torch.bool, pin_memory=True)
            else:
                torch.cuda.current_stream().synchronize()  # last rollout's upload out of these arrays has run
                self._host_starts[0].copy_(self._host_starts[T])
            host_rewards, host_starts = self._host_rewards.numpy(), self._host_starts.numpy()
        for s in range(T):
            if self.sde_sample_freq > 0 and s > 0 and s % self.sde_sample_freq == 0:
                self.policy.reset_noise()
            if device_env:
                if self._graph is not None:
                    self._graph.replay()
                    ops.LAUNCHES += self._kernels_per_replay
                else:
                    with torch.no_grad():
                        self._device_env_step()
                continue
            # host env: the policy half runs on the device (captured), the env half on the host
            if self._graph is not None:
                self._graph.replay()
                ops.LAUNCHES += self._kernels_per_replay
                a = self._graph_outputs
            else:
                with torch.no_grad():
                    a = self._policy_step()
            if self._wide_actions is not None:
                self._download_wide(a)
            next_obs, rewards, terminations, truncations, _ = self.vec_env.step(self._env_actions(a, landed=True))
            masks = (self.get_action_mask() if self.get_action_mask is not None and self.next_action_masks is not None
                     else None)
            self._upload_env_outputs(next_obs, masks)
            host_rewards[s] = np.asarray(rewards, dtype=np.float32).reshape(host_rewards.shape[1:])
            np.logical_or(terminations, truncations, out=host_starts[s + 1])
        if not device_env:  # one upload per rollout for the scalars the env produced on the host
            self._upload.bytes += self._host_rewards.numel() * 4 + self._host_starts.numel()
            self.rewards.copy_(self._host_rewards, non_blocking=True)
            self.episode_starts.copy_(self._host_starts[:T], non_blocking=True)
            self.next_episode_starts.copy_(self._host_starts[T], non_blocking=True)
        self._pack_raw()
        next_values = self.policy.value_device(self.next_obs) if output_next_values else None
        self.policy.train()
        return next_values

    def rollout(self, gamma, gae_lambda) -> VecRollout:
        next_values = self._rollout(output_next_values=True)
        assert next_values is not None
        self._reset_envs(self.num_envs_reset_every_rollout, self.rolling_num_envs_reset_every_rollout,
                         self.random_num_envs_reset_every_rollout)
        return VecRollout(
            device=self.device,
            next_episode_starts=self.next_episode_starts,
            next_values=next_values,
            obs=self.obs,
            actions=self.actions,
            rewards=self.rewards,
            episode_starts=self.episode_starts,
            values=self.values,
            logprobs=self.logprobs,
            action_masks=self.action_masks,
            gamma=gamma,
            gae_lambda=gae_lambda,
            scale_advantage_by_values_accuracy=self.scale_advantage_by_values_accuracy,
            full_batch_off_accelerator=self.full_batch_off_accelerator,
            subaction_mask=self.subaction_mask,
            action_plane_space=getattr(self.vec_env, "action_plane_space", None),
            out_advantages=self.advantages,
            out_returns=self.returns,
            include_num_actions=self.include_num_actions,
        )

    # -- masked resets (sync_step_rollout.py:218-278) ------------------------------------------------
    def _reset_envs(self, num_envs_reset: int, rolling_num_envs_reset: int, random_num_envs_reset: int) -> None:
        assert bool(num_envs_reset) + bool(rolling_num_envs_reset) + bool(random_num_envs_reset) <= 1, \
            "Only one of num_envs_reset, rolling_num_envs_reset, random_num_envs_reset can be set"
        N = self.vec_env.num_envs
        reset = np.zeros(N, dtype=np.bool_)
        if num_envs_reset > 0:
            reset[-num_envs_reset:] = True
        if rolling_num_envs_reset > 0:
            pairs = len(self.rolling_reset_indexes)
            end_idx = (self.rolling_mask_idx + rolling_num_envs_reset // 2) % pairs
            if end_idx < self.rolling_mask_idx:
                chosen = np.concatenate((self.rolling_reset_indexes[self.rolling_mask_idx:],
                                         self.rolling_reset_indexes[:end_idx]))
                self.rolling_reset_indexes = np.random.permutation(N // 2)
            else:
                chosen = self.rolling_reset_indexes[self.rolling_mask_idx:end_idx]
            pair_mask = np.zeros(N // 2, dtype=np.bool_)
            pair_mask[chosen] = True
            reset[pair_mask.repeat(2)] = True
            self.rolling_mask_idx = end_idx
        if random_num_envs_reset > 0:
            pair_mask = np.zeros(N // 2, dtype=np.bool_)
            pair_mask[np.random.choice(N // 2, random_num_envs_reset // 2, replace=False)] = True
            reset[pair_mask.repeat(2)] = True
        assert reset.sum() == num_envs_reset + rolling_num_envs_reset + random_num_envs_reset
        if not reset.any():
            return
        next_obs, action_mask, _ = self.vec_env.masked_reset(reset)
        rows = torch.from_numpy(np.nonzero(reset)[0]).to(self.device)
        fresh = torch.as_tensor(next_obs).to(self.device)
        if self._pack_in_step:
            self._raw_obs[rows] = fresh.to(self._raw_obs.dtype)
            self._pack_raw()
        else:
            self.next_obs[rows] = self.policy.pack_observations(fresh) if self._packed else fresh.to(self.next_obs.dtype)
        if self.next_action_masks is not None and action_mask is not None:
            if isinstance(self.next_action_masks, dict):
                for k, dst in self.next_action_masks.items():
                    dst[rows] = torch.as_tensor(action_mask[k]).to(self.device, dtype=torch.bool)
            else:
                self.next_action_masks[rows] = torch.as_tensor(action_mask).to(self.device, dtype=torch.bool)
