"""CPU tests of the PPO host logic, including the world_size-2 gloo paths (gradient all-reduce, statistics sync)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def test_flat_policy_layout_and_sb3_names():
    from pyflyt_drone_b200.ppo import FlatMlpPolicy
    pol = FlatMlpPolicy(28, torch.device("cpu"), seed=1)
    assert pol.count == 12361      # 6,276 + 6,081 + 4
    sd = pol.state_dict()
    assert set(sd) == {"mlp_extractor.policy_net.0.weight", "mlp_extractor.policy_net.0.bias",
                       "mlp_extractor.policy_net.2.weight", "mlp_extractor.policy_net.2.bias",
                       "mlp_extractor.value_net.0.weight", "mlp_extractor.value_net.0.bias",
                       "mlp_extractor.value_net.2.weight", "mlp_extractor.value_net.2.bias",
                       "action_net.weight", "action_net.bias", "value_net.weight", "value_net.bias", "log_std"}
    w = sd["mlp_extractor.policy_net.2.weight"]
    assert torch.allclose(w @ w.t(), 2 * torch.eye(64), atol=1e-4)          # orthogonal, gain sqrt(2)
    assert torch.allclose(sd["action_net.weight"] @ sd["action_net.weight"].t(), 1e-4 * torch.eye(4), atol=1e-6)
    assert float(sd["log_std"].abs().max()) == 0.0 and float(sd["value_net.bias"].abs().max()) == 0.0
    pol2 = FlatMlpPolicy(28, torch.device("cpu"), seed=2)
    pol2.load_state_dict(sd)
    assert torch.equal(pol2.theta, pol.theta)
    # evaluate_actions is the diagonal-Gaussian log-prob / entropy of SB3
    obs, act = torch.randn(5, 28), torch.randn(5, 4)
    v, lp, ent = pol.evaluate_actions(obs, act)
    mean, _ = pol.towers(obs)
    ref = torch.distributions.Normal(mean, torch.ones(4)).log_prob(act).sum(-1)
    assert torch.allclose(lp, ref, atol=1e-5)
    assert torch.allclose(ent, torch.distributions.Normal(mean, torch.ones(4)).entropy().sum(-1), atol=1e-5)


def test_policy_state_dict_loads_into_an_sb3_shaped_module(tmp_path):
    """the exported policy.pth must load, strict, into a module with stable_baselines3's
    ActorCriticPolicy structure (mlp_extractor.policy_net / value_net = Sequential(Linear, Tanh, Linear, Tanh),
    action_net, value_net, log_std) and give the same outputs; the VecNormalize moments round-trip through the .npz."""
    import torch.nn as nn
    from pyflyt_drone_b200.ppo import FlatMlpPolicy, export_vecnorm_npz, import_vecnorm_npz

    class Extractor(nn.Module):
        def __init__(self, d):
            super().__init__()
            self.policy_net = nn.Sequential(nn.Linear(d, 64), nn.Tanh(), nn.Linear(64, 64), nn.Tanh())
            self.value_net = nn.Sequential(nn.Linear(d, 64), nn.Tanh(), nn.Linear(64, 64), nn.Tanh())

    class Sb3Shaped(nn.Module):
        def __init__(self, d):
            super().__init__()
            self.mlp_extractor = Extractor(d)
            self.action_net = nn.Linear(64, 4)
            self.value_net = nn.Linear(64, 1)
            self.log_std = nn.Parameter(torch.zeros(4))

    pol = FlatMlpPolicy(28, torch.device("cpu"), seed=3)
    with torch.no_grad():
        pol.theta.add_(0.05 * torch.randn(pol.count, generator=torch.Generator().manual_seed(0)))
    path = tmp_path / "policy.pth"
    torch.save({k: v.cpu() for k, v in pol.state_dict().items()}, path)
    m = Sb3Shaped(28)
    m.load_state_dict(torch.load(path, weights_only=True), strict=True)
    obs = torch.randn(7, 28)
    mean, val = pol.towers(obs)
    assert torch.allclose(m.action_net(m.mlp_extractor.policy_net(obs)), mean, atol=1e-6)
    assert torch.allclose(m.value_net(m.mlp_extractor.value_net(obs)).squeeze(-1), val.reshape(-1), atol=1e-6)
    assert torch.equal(m.log_std.detach(), pol.view("log_std").detach())
    # and back: a state_dict coming from such a module restores the flat vector
    pol2 = FlatMlpPolicy(28, torch.device("cpu"), seed=9)
    pol2.load_state_dict(m.state_dict())
    assert torch.equal(pol2.theta, pol.theta)
    sd = {"obs_rms.mean": np.arange(28.0), "obs_rms.var": np.arange(28.0) + 1, "obs_rms.count": 1234.0, "ret_rms.mean": 0.5,
          "ret_rms.var": 2.0, "ret_rms.count": 99.0, "clip_obs": 10.0, "clip_reward": 10.0, "gamma": 0.99,
          "norm_obs": True, "norm_reward": True}
    export_vecnorm_npz(sd, str(tmp_path / "vecnorm.npz"))
    back = import_vecnorm_npz(str(tmp_path / "vecnorm.npz"))
    assert np.array_equal(back["obs_rms.mean"], sd["obs_rms.mean"]) and back["ret_rms.var"] == 2.0 and back["gamma"] == 0.99


def test_moment_merge_is_chan():
    from pyflyt_drone_b200.ppo import DeviceVecNormalize
    rng = np.random.default_rng(0)
    a, b = rng.normal(3, 2, (500, 4)), rng.normal(-1, 5, (300, 4))
    base = torch.tensor(np.concatenate([a.mean(0), a.var(0), [500.0]]))
    acc = torch.tensor(np.concatenate([b.sum(0), (b * b).sum(0), [300.0]]))
    out = DeviceVecNormalize._merge(base, acc, 4).numpy()
    both = np.concatenate([a, b])
    assert np.allclose(out[:4], both.mean(0)) and np.allclose(out[4:8], both.var(0)) and out[8] == 800


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from pyflyt_drone_b200.ppo import DeviceVecNormalize, FlatMlpPolicy
    dev = torch.device("cpu")
    # statistics sync: each rank accumulated its own batch; after the sync both hold the pooled moments
    rng = np.random.default_rng(10 + rank)
    x = rng.normal(rank * 3.0, 1.0 + rank, (400, 28))
    vn = DeviceVecNormalize(28, 8, dev)
    vn.obs_accum[:28] = torch.tensor(x.sum(0)); vn.obs_accum[28:56] = torch.tensor((x * x).sum(0)); vn.obs_accum[56] = 400
    r = rng.normal(0, 2 + rank, 400)
    vn.ret_accum[:] = torch.tensor([r.sum(), (r * r).sum(), 400.0])
    vn.sync_across_ranks()
    # gradient all-reduce: mean of per-rank gradients == gradient of the pooled batch (equal shard sizes)
    pol = FlatMlpPolicy(28, dev, seed=0)
    dist.broadcast(pol.theta.data, src=0)
    g = torch.Generator().manual_seed(5)
    obs, act = torch.randn(64, 28, generator=g), torch.randn(64, 4, generator=g)
    sl = slice(rank * 32, (rank + 1) * 32)
    v, lp, _ = pol.evaluate_actions(obs[sl], act[sl])
    (lp.mean() + v.pow(2).mean()).backward()
    dist.all_reduce(pol.theta.grad); pol.theta.grad.div_(world)
    if rank == 0:
        pol_full = FlatMlpPolicy(28, dev, seed=0)
        v, lp, _ = pol_full.evaluate_actions(obs, act)
        (lp.mean() + v.pow(2).mean()).backward()
        torch.save({"obs_stats": vn.obs_stats, "ret_stats": vn.ret_stats, "grad": pol.theta.grad,
                    "grad_full": pol_full.theta.grad}, out)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_gradient_allreduce_and_stats_sync(tmp_path):
    out = str(tmp_path / "r0.pt")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    res = torch.load(out, weights_only=False)
    assert torch.allclose(res["grad"], res["grad_full"], atol=1e-6)
    xs = [np.random.default_rng(10 + r).normal(r * 3.0, 1.0 + r, (400, 28)) for r in range(2)]
    pooled = np.concatenate(xs)
    st = res["obs_stats"].numpy()
    assert np.allclose(st[:28], pooled.mean(0), atol=1e-3) and np.allclose(st[28:56], pooled.var(0), rtol=1e-3)
    assert st[56] == pytest.approx(800 + 1e-4)


def test_six_channel_policy_layout():
    """train/train_lowlevel_cmd.py trains MlpPolicy on FixedwingLowLevelEnv: Box(21) observations, Box(6) actions."""
    from pyflyt_drone_b200.ppo import FlatMlpPolicy
    pol = FlatMlpPolicy(21, torch.device("cpu"), seed=1, act_dim=6)
    assert pol.count == (64 * 21 + 64 + 64 * 64 + 64 + 6 * 64 + 6) + (64 * 21 + 64 + 64 * 64 + 64 + 64 + 1) + 6
    sd = pol.state_dict()
    assert sd["action_net.weight"].shape == (6, 64) and sd["log_std"].shape == (6,) and sd["value_net.weight"].shape == (1, 64)
    obs, act = torch.randn(5, 21), torch.randn(5, 6)
    v, lp, ent = pol.evaluate_actions(obs, act)
    mean, _ = pol.towers(obs)
    assert torch.allclose(lp, torch.distributions.Normal(mean, torch.ones(6)).log_prob(act).sum(-1), atol=1e-5)
    with pytest.raises(ValueError):
        FlatMlpPolicy(21, torch.device("cpu"), act_dim=5)
