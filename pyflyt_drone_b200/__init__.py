"""pyflyt_drone_b200 -- B200-native batched fixed-wing simulator behind the reference's env/VecEnv surface.

Host side is Python; the env step is hand-written sm_100a CUDA in libfwsim.so (C ABI: include/fwsim.h).
"""
from .config import EnvConfig, from_gym_kwargs, make_config, waypoints_v3, waypoint_objlock, physics_only, lowlevel, objlock_duck  # noqa: F401

__all__ = ["EnvConfig", "make_config", "from_gym_kwargs", "waypoints_v3", "waypoint_objlock", "physics_only", "lowlevel", "objlock_duck",
           "FixedwingVecEnv"]


def __getattr__(name):
    # FixedwingVecEnv loads libfwsim.so on construction; importing it lazily keeps `import pyflyt_drone_b200`
    # usable for config handling on a machine without the built library.
    if name == "FixedwingVecEnv":
        from .vec_env import FixedwingVecEnv
        return FixedwingVecEnv
    raise AttributeError(name)
