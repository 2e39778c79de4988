"""Algorithm ABC (mirrors rl_algo_impls/shared/algorithm.py:19-60: constructor fields, learn(), and
optimizer-state save / load to ``optimizer.pt``)."""
import logging
import os
from abc import ABC, abstractmethod
from typing import List, Optional, TypeVar

import torch
from torch.optim import Optimizer

OPTIMIZER_FILENAME = "optimizer.pt"
AlgorithmSelf = TypeVar("AlgorithmSelf", bound="Algorithm")


class Algorithm(ABC):
    @abstractmethod
    def __init__(self, policy, device: torch.device, tb_writer, learning_rate: float, optimizer: Optimizer,
                 **kwargs) -> None:
        super().__init__()
        self.policy = policy
        self.device = device
        self.tb_writer = tb_writer
        self.learning_rate = learning_rate
        self.optimizer = optimizer

    @abstractmethod
    def learn(self: AlgorithmSelf, train_timesteps: int, rollout_generator, callbacks: Optional[List] = None,
              total_timesteps: Optional[int] = None, start_timesteps: int = 0) -> AlgorithmSelf: ...

    def save(self, path: str) -> None:
        torch.save(self.optimizer.state_dict(), os.path.join(path, OPTIMIZER_FILENAME))

    def load(self, path: str) -> None:
        optimizer_path = os.path.join(path, OPTIMIZER_FILENAME)
        if os.path.exists(optimizer_path):
            self.optimizer.load_state_dict(torch.load(optimizer_path, map_location=self.device))
        else:
            logging.info(f"Optimizer state not found at {optimizer_path}. Not overwriting optimizer state.")


def update_learning_rate(optimizer: Optimizer, learning_rate: float) -> None:
    """shared/schedule.py:64-66"""
    for group in optimizer.param_groups:
        if isinstance(group["lr"], torch.Tensor):  # capturable optimizer: the value lives on the device
            group["lr"].fill_(float(learning_rate))
        else:
            group["lr"] = learning_rate
