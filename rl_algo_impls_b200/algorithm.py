"""Base class of the learners (PPO, A2C).

Keeps the surface the reference's runner relies on (``rl_algo_impls/shared/algorithm.py:19-60``):
the five constructor fields, an abstract ``learn``, and ``save`` / ``load`` of the optimizer state
under ``<dir>/optimizer.pt``.  ``load`` takes a state dict written by either implementation: the reference's has
``capturable=False``, a Python-float learning rate and CPU step counters, which ``load`` converts back to what this
optimizer was built with (device step counters, the SAME device learning-rate tensor captured graphs point at).
"""
import logging
from abc import ABC, abstractmethod
from pathlib import Path
from typing import List, Optional, TypeVar

import torch
from torch.optim import Optimizer

OPTIMIZER_FILENAME = "optimizer.pt"
AlgorithmSelf = TypeVar("AlgorithmSelf", bound="Algorithm")
_log = logging.getLogger(__name__)


class Algorithm(ABC):
    policy: torch.nn.Module
    device: torch.device
    learning_rate: float
    optimizer: Optimizer

    @abstractmethod
    def __init__(self, policy, device: torch.device, tb_writer, learning_rate: float, optimizer: Optimizer,
                 **kwargs) -> None:
        super().__init__()
        self.policy, self.device, self.tb_writer = policy, device, tb_writer
        self.learning_rate, self.optimizer = learning_rate, optimizer

    @abstractmethod
    def learn(self: AlgorithmSelf, train_timesteps: int, rollout_generator, callbacks: Optional[List] = None,
              total_timesteps: Optional[int] = None, start_timesteps: int = 0) -> AlgorithmSelf:
        """Train for ``train_timesteps`` env steps drawn from ``rollout_generator``."""

    def save(self, path: str) -> None:
        torch.save(self.optimizer.state_dict(), Path(path) / OPTIMIZER_FILENAME)

    def load(self, path: str) -> None:
        state_file = Path(path) / OPTIMIZER_FILENAME
        if not state_file.exists():
            _log.info("no optimizer state at %s: the optimizer keeps its current state", state_file)
            return
        self.load_optimizer_state(torch.load(state_file, map_location=self.device))

    def load_optimizer_state(self, state: dict) -> None:
        opt = self.optimizer
        before = [(g.get("capturable"), g.get("lr")) for g in opt.param_groups]
        opt.load_state_dict(state)
        for group, (capturable, lr) in zip(opt.param_groups, before):
            if isinstance(lr, torch.Tensor):  # keep the tensor object: captured update graphs read it
                loaded = group["lr"]
                lr.fill_(float(loaded))
                group["lr"] = lr
            if capturable is not None:
                group["capturable"] = capturable
            if capturable:
                for p in group["params"]:
                    st = opt.state.get(p)
                    if st and "step" in st:
                        step = st["step"]
                        st["step"] = (step.to(device=p.device, dtype=torch.float32) if isinstance(step, torch.Tensor)
                                      else torch.tensor(float(step), dtype=torch.float32, device=p.device))
        graphs = getattr(self, "_update_graphs", None)
        if graphs:  # captured updates hold the old state tensors
            graphs.clear()


def update_learning_rate(optimizer: Optimizer, learning_rate: float) -> None:
    """Apply the (possibly scheduled) learning rate to every parameter group (shared/schedule.py:64-66).
    A capturable optimizer keeps its rate in a device tensor: fill it instead of rebinding."""
    for group in optimizer.param_groups:
        current = group["lr"]
        if isinstance(current, torch.Tensor):
            current.fill_(float(learning_rate))
        else:
            group["lr"] = learning_rate
