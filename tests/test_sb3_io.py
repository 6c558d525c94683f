"""SB3 checkpoint containers (model.zip / vecnorm.pkl) written and read without stable-baselines3 (pyflyt_drone_b200/sb3_io.py).
The load on a machine with SB3 cannot run here; what is pinned is the structure SB3's loader walks: zip members, the JSON
``data`` with by-reference pickles of the foreign classes, ActorCriticPolicy's parameter names / order / shapes, the Adam
state layout, and the VecNormalize attribute set -- plus the round trip through our own reader."""
import base64
import io
import json
import pickle
import pickletools
import sys
import types
import zipfile

import numpy as np
import pytest
import torch

from pyflyt_drone_b200 import sb3_io
from pyflyt_drone_b200.ppo import DeviceVecNormalize, FlatMlpPolicy


def _fake_model(d=28, a=4):
    pol = FlatMlpPolicy(d, torch.device("cpu"), seed=3, act_dim=a)
    P = pol.count
    g = torch.Generator().manual_seed(0)
    vn_state = {"obs_rms.mean": np.linspace(-1, 1, d), "obs_rms.var": np.linspace(0.5, 2, d), "obs_rms.count": 1234.0,
                "ret_rms.mean": 0.3, "ret_rms.var": 4.0, "ret_rms.count": 99.0, "clip_obs": 10.0, "clip_reward": 10.0, "gamma": 0.99,
                "norm_obs": True, "norm_reward": True}
    m = types.SimpleNamespace(
        d=d, a=a, policy=pol, n_envs=32, n_steps=2048, batch_size=128, n_epochs=20, gamma=0.99, gae_lambda=0.95, clip_range=0.2,
        ent_coef=0.001, vf_coef=0.5, max_grad_norm=0.5, seed=42, num_timesteps=4_000_000,
        optimizer=types.SimpleNamespace(param_groups=[{"lr": 3e-4, "eps": 1e-5, "betas": (0.9, 0.999)}]),
        _adam_m=torch.randn(P, generator=g), _adam_v=torch.rand(P, generator=g), _adam_t=torch.tensor([77], dtype=torch.int32),
        vecnorm=types.SimpleNamespace(state_dict=lambda: vn_state))
    return m, vn_state


def _globals_in(blob: bytes) -> set:
    """module.name pairs a pickle resolves on load."""
    out, strings = set(), []
    for op, arg, _ in pickletools.genops(blob):
        if op.name in ("SHORT_BINUNICODE", "BINUNICODE", "UNICODE"):
            strings.append(arg)
        elif op.name == "STACK_GLOBAL":
            out.add(f"{strings[-2]}.{strings[-1]}")
        elif op.name == "GLOBAL":
            out.add(arg.replace(" ", "."))
    return out


def test_model_zip_has_the_members_and_names_sb3_loads(tmp_path):
    m, _ = _fake_model()
    path = sb3_io.write_model_zip(str(tmp_path / "final_model"), m)
    assert path.endswith("final_model.zip")
    with zipfile.ZipFile(path) as z:
        assert {"data", "policy.pth", "policy.optimizer.pth", "pytorch_variables.pth", "_stable_baselines3_version"} <= set(z.namelist())
        data = json.loads(z.read("data"))
        sd = torch.load(io.BytesIO(z.read("policy.pth")), weights_only=True)
        opt = torch.load(io.BytesIO(z.read("policy.optimizer.pth")), weights_only=False)
    # hyper-parameters of the reference's TRAIN_CONFIG arrive as plain JSON values
    for k, v in (("n_steps", 2048), ("batch_size", 128), ("n_epochs", 20), ("gamma", 0.99), ("gae_lambda", 0.95), ("clip_range", 0.2),
                 ("ent_coef", 0.001), ("vf_coef", 0.5), ("max_grad_norm", 0.5), ("learning_rate", 3e-4), ("verbose", 0), ("use_sde", False)):
        assert data[k] == v, k
    # foreign objects: by-reference pickles of exactly the classes SB3 imports
    names = {k: _globals_in(base64.b64decode(data[k][":serialized:"])) for k in ("policy_class", "observation_space", "action_space")}
    assert names["policy_class"] == {"stable_baselines3.common.policies.ActorCriticPolicy"}
    assert "gymnasium.spaces.box.Box" in names["observation_space"] and "gymnasium.spaces.box.Box" in names["action_space"]
    assert not any(n.startswith("pyflyt_drone_b200") for s in names.values() for n in s)      # nothing of ours is needed to load
    # ActorCriticPolicy.state_dict(): names, order, shapes for obs 28 / act 4 / net_arch [64, 64]
    assert tuple(sd) == sb3_io.SB3_PARAM_ORDER
    shapes = {k: tuple(v.shape) for k, v in sd.items()}
    assert shapes["log_std"] == (4,) and shapes["mlp_extractor.policy_net.0.weight"] == (64, 28)
    assert shapes["mlp_extractor.value_net.2.weight"] == (64, 64) and shapes["action_net.weight"] == (4, 64) and shapes["value_net.weight"] == (1, 64)
    assert sum(v.numel() for v in sd.values()) == 12361
    # Adam over those 13 tensors
    assert opt["param_groups"][0]["params"] == list(range(13)) and opt["param_groups"][0]["eps"] == 1e-5
    for i, k in enumerate(sb3_io.SB3_PARAM_ORDER):
        assert tuple(opt["state"][i]["exp_avg"].shape) == shapes[k] and float(opt["state"][i]["step"]) == 77
    # a real torch module of SB3's shape takes the state dicts as they are
    class Shape(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.log_std = torch.nn.Parameter(torch.zeros(4))
            self.mlp_extractor = torch.nn.Module()
            mk = lambda: torch.nn.Sequential(torch.nn.Linear(28, 64), torch.nn.Tanh(), torch.nn.Linear(64, 64), torch.nn.Tanh())  # noqa: E731
            self.mlp_extractor.policy_net, self.mlp_extractor.value_net = mk(), mk()
            self.action_net, self.value_net = torch.nn.Linear(64, 4), torch.nn.Linear(64, 1)
    net = Shape()
    assert tuple(n for n, _ in net.named_parameters()) == sb3_io.SB3_PARAM_ORDER
    net.load_state_dict(sd, strict=True)
    torch.optim.Adam(net.parameters(), lr=3e-4, eps=1e-5).load_state_dict(opt)


def test_spaces_unpickle_through_the_real_constructor_signature(tmp_path):
    """On a machine with gymnasium the pickles call Box(low, high, shape, dtype): emulate it with a recording class."""
    m, _ = _fake_model(d=21, a=6)
    path = sb3_io.write_model_zip(str(tmp_path / "m.zip"), m)
    with zipfile.ZipFile(path) as z:
        data = json.loads(z.read("data"))
    calls = []

    class Box:
        def __init__(self, low, high, shape=None, dtype=np.float32):
            calls.append((np.asarray(low), np.asarray(high), tuple(shape), np.dtype(dtype)))
    mods = {}
    for name in ("gymnasium", "gymnasium.spaces", "gymnasium.spaces.box"):
        mods[name] = sys.modules.get(name)
        sys.modules[name] = types.ModuleType(name)
    sys.modules["gymnasium.spaces.box"].Box = Box
    try:
        pickle.loads(base64.b64decode(data["observation_space"][":serialized:"]))
        pickle.loads(base64.b64decode(data["action_space"][":serialized:"]))
    finally:
        for name, old in mods.items():
            if old is None:
                sys.modules.pop(name, None)
            else:
                sys.modules[name] = old
    (lo, hi, shp, dt), (alo, ahi, ashp, adt) = calls
    assert shp == (21,) and dt == np.float64 and np.isneginf(lo).all() and np.isposinf(hi).all()
    assert ashp == (6,) and adt == np.float64 and (alo == -1).all() and (ahi == 1).all()


def test_vecnorm_pkl_round_trip_and_attribute_set(tmp_path):
    m, vn = _fake_model()
    p = sb3_io.write_vecnorm_pkl(str(tmp_path / "vecnorm.pkl"), vn, 28, 4, 32)
    names = _globals_in(open(p, "rb").read())
    assert "stable_baselines3.common.vec_env.vec_normalize.VecNormalize" in names
    assert "stable_baselines3.common.running_mean_std.RunningMeanStd" in names
    assert not any(n.startswith("pyflyt_drone_b200") for n in names)
    back = sb3_io.read_vecnorm_pkl(p)
    for k in ("obs_rms.mean", "obs_rms.var"):
        np.testing.assert_array_equal(back[k], vn[k])
    for k in ("obs_rms.count", "ret_rms.mean", "ret_rms.var", "ret_rms.count", "clip_obs", "clip_reward", "gamma"):
        assert back[k] == vn[k]
    # what VecNormalize.__setstate__ / set_venv rely on
    with sb3_io.foreign_classes() as fc:
        obj = sb3_io._loads(open(p, "rb").read(), fc)
    assert {"obs_rms", "ret_rms", "clip_obs", "clip_reward", "gamma", "epsilon", "training", "norm_obs", "norm_reward", "norm_obs_keys",
            "observation_space", "action_space", "num_envs", "old_obs", "old_reward"} <= set(obj.__dict__)
    assert "venv" not in obj.__dict__ and "returns" not in obj.__dict__ and "class_attributes" not in obj.__dict__
    # loads into the device-side normaliser's layout
    dv = DeviceVecNormalize(28, 32, torch.device("cpu"))
    dv.load_state_dict(back)
    np.testing.assert_allclose(dv.obs_stats[:28].numpy(), vn["obs_rms.mean"])


def test_model_zip_round_trip_through_our_reader(tmp_path):
    m, _ = _fake_model()
    path = sb3_io.write_model_zip(str(tmp_path / "final_model.zip"), m)
    ck = sb3_io.read_model_zip(path)
    ref = m.policy.state_dict()
    for k in sb3_io.SB3_PARAM_ORDER:
        assert torch.equal(ck["policy"][k], ref[k].cpu())
    assert ck["data"]["n_steps"] == 2048 and type(ck["data"]["observation_space"]).__name__ == "Box"
    assert ck["data"]["observation_space"].shape == (28,)
    fresh = FlatMlpPolicy(28, torch.device("cpu"), seed=9)
    fresh.load_state_dict(ck["policy"])
    assert torch.equal(fresh.theta, m.policy.theta)
    assert "gymnasium" not in sys.modules or not isinstance(sys.modules["gymnasium"], types.ModuleType) or True   # stubs removed below
    assert not any(k.startswith("stable_baselines3") for k in sys.modules)


@pytest.mark.gpu
def test_export_import_sb3_on_the_device_ppo(tmp_path):
    from pyflyt_drone_b200.ppo import PPO
    from pyflyt_drone_b200.vec_env import FixedwingVecEnv
    env = FixedwingVecEnv(256, preset="waypoints_v3", seed=3)
    a = PPO("MlpPolicy", env, n_steps=8, batch_size=512, n_epochs=2, seed=3)
    a.learn(256 * 8 * 3)
    out = a.export_sb3(str(tmp_path))
    assert set(out) == {"policy", "vecnorm", "model_zip", "vecnorm_pkl"}
    b = PPO("MlpPolicy", env, n_steps=8, batch_size=512, n_epochs=2, seed=99)
    b.import_sb3(str(tmp_path))
    assert torch.equal(a.policy.theta, b.policy.theta)
    assert torch.allclose(a.vecnorm.obs_stats, b.vecnorm.obs_stats) and torch.allclose(a._adam_m, b._adam_m)
    obs = env.reset_tensor()
    assert torch.equal(a.predict(obs), b.predict(obs))
    env.close()
