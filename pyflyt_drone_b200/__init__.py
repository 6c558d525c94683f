"""pyflyt_drone_b200 -- B200-native batched fixed-wing simulator behind the reference's env/VecEnv surface.

Host side is Python; the env step is hand-written sm_100a CUDA in libfwsim.so (C ABI: include/fwsim.h).
"""
from .config import EnvConfig, make_config, waypoints_v3, waypoint_objlock, physics_only  # noqa: F401

__all__ = ["EnvConfig", "make_config", "waypoints_v3", "waypoint_objlock", "physics_only"]
