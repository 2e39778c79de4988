from .actor_critic import ACForward, ActorCritic, Step, clamp_actions
from .networks import HeadOutputs

__all__ = ["ACForward", "ActorCritic", "Step", "clamp_actions", "HeadOutputs"]
