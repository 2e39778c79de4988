from .synthetic import SPECS, EnvSpec, SyntheticVecEnv, make_synthetic_env

__all__ = ["SPECS", "EnvSpec", "SyntheticVecEnv", "make_synthetic_env"]
