"""The stable-baselines3 ``VecEnv`` / ``VecEnvWrapper`` base classes, or stand-ins of the same public shape.

``FixedwingVecEnv`` derives from ``VecEnv`` so that SB3's own code (``VecNormalize``, ``PPO.collect_rollouts``,
``evaluate_policy``, ``isinstance(env, VecEnv)`` in the reference's callbacks,
/root/reference/train/train_Fixedwing_Waypoints_v3.py:175-176) accepts it.  stable_baselines3 is not installed in
this image: when it is importable its classes are used, otherwise the two small classes below keep the same
attribute and method names (``num_envs``, ``observation_space``, ``action_space``, ``render_mode``, ``reset_infos``,
``step() = step_async + step_wait``, ``unwrapped``, ``getattr_depth_check``; a wrapper that forwards to ``venv``) so
the host-side tests can exercise the seam with an SB3-shaped wrapper.
"""
from __future__ import annotations

from typing import Any

try:  # pragma: no cover - exercised only where stable_baselines3 exists
    from stable_baselines3.common.vec_env.base_vec_env import VecEnv, VecEnvWrapper  # type: ignore
    HAVE_SB3 = True
except Exception:  # ModuleNotFoundError in this image
    HAVE_SB3 = False

    class VecEnv:  # type: ignore[no-redef]
        """Stand-in for ``stable_baselines3.common.vec_env.VecEnv`` (abstract batched environment)."""

        def __init__(self, num_envs: int, observation_space, action_space):
            self.num_envs = int(num_envs)
            self.observation_space = observation_space
            self.action_space = action_space
            self.reset_infos: list[dict[str, Any]] = [{} for _ in range(self.num_envs)]
            self._seeds: list[int | None] = [None for _ in range(self.num_envs)]
            self._options: list[dict[str, Any]] = [{} for _ in range(self.num_envs)]
            self.render_mode = None

        def _reset_seeds(self) -> None:
            self._seeds = [None for _ in range(self.num_envs)]

        def _reset_options(self) -> None:
            self._options = [{} for _ in range(self.num_envs)]

        def reset(self):
            raise NotImplementedError

        def step_async(self, actions) -> None:
            raise NotImplementedError

        def step_wait(self):
            raise NotImplementedError

        def close(self) -> None:
            raise NotImplementedError

        def step(self, actions):
            self.step_async(actions)
            return self.step_wait()

        def get_images(self):
            raise NotImplementedError

        def render(self, mode: str | None = None):
            return None

        def set_options(self, options=None) -> None:
            if options is None:
                options = {}
            self._options = [dict(options) for _ in range(self.num_envs)] if isinstance(options, dict) else list(options)

        @property
        def unwrapped(self) -> "VecEnv":
            return self.venv.unwrapped if isinstance(self, VecEnvWrapper) else self

        def getattr_depth_check(self, name: str, already_found: bool):
            return f"{type(self).__module__}.{type(self).__name__}" if hasattr(self, name) and already_found else None

        def _get_indices(self, indices):
            if indices is None:
                return range(self.num_envs)
            if isinstance(indices, int):
                return [indices]
            return indices

    class VecEnvWrapper(VecEnv):  # type: ignore[no-redef]
        """Stand-in for ``stable_baselines3.common.vec_env.VecEnvWrapper``: forwards everything to ``venv``."""

        def __init__(self, venv: VecEnv, observation_space=None, action_space=None):
            self.venv = venv
            super().__init__(venv.num_envs, observation_space or venv.observation_space, action_space or venv.action_space)
            self.class_attributes = dict(vars(type(self)))

        def step_async(self, actions) -> None:
            self.venv.step_async(actions)

        def reset(self):
            raise NotImplementedError

        def step_wait(self):
            raise NotImplementedError

        def seed(self, seed=None):
            return self.venv.seed(seed)

        def set_options(self, options=None) -> None:
            return self.venv.set_options(options)

        def close(self) -> None:
            return self.venv.close()

        def get_attr(self, attr_name: str, indices=None):
            return self.venv.get_attr(attr_name, indices)

        def set_attr(self, attr_name: str, value, indices=None) -> None:
            return self.venv.set_attr(attr_name, value, indices)

        def env_method(self, method_name: str, *args, indices=None, **kwargs):
            return self.venv.env_method(method_name, *args, indices=indices, **kwargs)

        def env_is_wrapped(self, wrapper_class, indices=None):
            return self.venv.env_is_wrapped(wrapper_class, indices=indices)

        def __getattr__(self, name: str):
            if name == "venv":
                raise AttributeError(name)
            return getattr(self.venv, name)
