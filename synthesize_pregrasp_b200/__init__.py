"""pregrasp_b200: B200-native (sm_100a) implementation of the MPPI rollout hot path of synthesize_pregrasp."""
from .envspec import ENV_NAMES, EnvSpec, make_spec  # noqa: F401

__all__ = ["ENV_NAMES", "EnvSpec", "make_spec"]
