"""CPU PPO baseline -- TEST / BENCH INFRASTRUCTURE ONLY (same rules as the rest of oracle/).

What the reference runs around its env on the host (train/train_Fixedwing_Waypoints_v3.py:251-337,
train/train_Fixedwing_Waypoints_ObjLock.py:306-365): ``VecNormalize(SubprocVecEnv(...))`` + stable_baselines3
``PPO("MlpPolicy", ...)`` on ``device="cpu"``.  stable_baselines3 is not installable in this image, so this file restates
the algorithm [UP-RECALL u10 of SURVEY.md] with plain fp32 PyTorch on the CPU over the fp64 oracle VecEnv: separate pi/vf
towers [64, 64] tanh, orthogonal init, state-independent log_std, VecNormalize(obs + reward, clip 10, RunningMeanStd
eps 1e-4), GAE with time-limit bootstrap, per-minibatch advantage normalisation, clipped surrogate + vf_coef * MSE -
ent_coef * entropy, clip_grad_norm_, Adam(eps 1e-5).  It is the CPU arm of bench.py's ``ppo`` object (kind "port") and
the reference the device PPO's learning behaviour can be compared with; nothing in pyflyt_drone_b200/ imports it.
"""
from __future__ import annotations

import math
import time

import numpy as np
import torch

from . import fw_oracle as fo


class RunningMeanStd:
    def __init__(self, shape=(), eps: float = 1e-4):
        self.mean, self.var, self.count = np.zeros(shape, np.float64), np.ones(shape, np.float64), eps

    def update(self, x: np.ndarray) -> None:
        bm, bv, bn = x.mean(axis=0), x.var(axis=0), x.shape[0]
        delta, tot = bm - self.mean, self.count + bn
        self.mean = self.mean + delta * bn / tot
        self.var = (self.var * self.count + bv * bn + delta * delta * self.count * bn / tot) / tot
        self.count = tot


class CpuPPO:
    """One process, ``nthreads`` host threads for the env step (the stand-in for SubprocVecEnv's worker processes) and
    torch's intra-op threads for the policy."""

    def __init__(self, cfg: dict, n_envs: int, n_steps: int, batch_size: int, n_epochs: int, lr: float = 3e-4,
                 gamma: float = 0.99, gae_lambda: float = 0.95, clip_range: float = 0.2, ent_coef: float = 0.001,
                 vf_coef: float = 0.5, max_grad_norm: float = 0.5, seed: int = 42, nthreads: int = 1):
        torch.manual_seed(seed)
        self.env = fo.OracleVecEnv(cfg, n_envs, seed=seed, nthreads=nthreads)
        self.n, self.T, self.bs, self.epochs = int(n_envs), int(n_steps), int(batch_size), int(n_epochs)
        self.gamma, self.lam, self.clip, self.ent, self.vfc, self.mgn = gamma, gae_lambda, clip_range, ent_coef, vf_coef, max_grad_norm
        d, a = self.env.obs_dim, self.env.act_dim
        self.d, self.a = d, a

        def tower(out, gain):
            layers = [torch.nn.Linear(d, 64), torch.nn.Tanh(), torch.nn.Linear(64, 64), torch.nn.Tanh(), torch.nn.Linear(64, out)]
            for m, g in zip((layers[0], layers[2], layers[4]), (math.sqrt(2), math.sqrt(2), gain)):
                torch.nn.init.orthogonal_(m.weight, gain=g)
                torch.nn.init.zeros_(m.bias)
            return torch.nn.Sequential(*layers)

        self.pi, self.vf = tower(a, 0.01), tower(1, 1.0)
        self.log_std = torch.nn.Parameter(torch.zeros(a))
        self.params = list(self.pi.parameters()) + list(self.vf.parameters()) + [self.log_std]
        self.opt = torch.optim.Adam(self.params, lr=lr, eps=1e-5)
        self.obs_rms, self.ret_rms = RunningMeanStd((d,)), RunningMeanStd(())
        self.ret = np.zeros(self.n)
        self.obs = None
        self.rollout_s = self.update_s = 0.0
        self.samples = 0

    def _norm(self, o: np.ndarray) -> np.ndarray:
        return np.clip((o - self.obs_rms.mean) / np.sqrt(self.obs_rms.var + 1e-8), -10.0, 10.0)

    def _logp(self, mean, act):
        z = (act - mean) * torch.exp(-self.log_std)
        return (-0.5 * z * z - self.log_std - 0.5 * math.log(2 * math.pi)).sum(-1)

    def iteration(self) -> None:
        n, T, d, a = self.n, self.T, self.d, self.a
        t0 = time.perf_counter()
        if self.obs is None:
            self.obs = self.env.reset().copy()
            self.obs_rms.update(self.obs)
        B = dict(obs=np.zeros((T, n, d), np.float32), act=np.zeros((T, n, a), np.float32), rew=np.zeros((T, n), np.float32),
                 done=np.zeros((T, n), np.float32), val=np.zeros((T, n), np.float32), logp=np.zeros((T, n), np.float32))
        with torch.no_grad():
            for t in range(T):
                on = torch.from_numpy(self._norm(self.obs).astype(np.float32))
                mean, val = self.pi(on), self.vf(on).squeeze(-1)
                act = mean + torch.exp(self.log_std) * torch.randn_like(mean)
                B["obs"][t], B["act"][t], B["val"][t], B["logp"][t] = on.numpy(), act.numpy(), val.numpy(), self._logp(mean, act).numpy()
                obs, rew, flags, term = self.env.step(np.clip(act.numpy(), -1.0, 1.0))
                done = (flags & 3) != 0
                self.obs_rms.update(obs)
                self.ret = self.ret * self.gamma + rew
                self.ret_rms.update(self.ret)
                r = np.clip(rew / np.sqrt(self.ret_rms.var + 1e-8), -10.0, 10.0)
                self.ret[done] = 0.0
                tl = ((flags & 2) != 0) & ((flags & 1) == 0)           # TimeLimit.truncated: bootstrap with V(terminal obs)
                if tl.any():
                    tv = self.vf(torch.from_numpy(self._norm(term[tl]).astype(np.float32))).squeeze(-1).numpy()
                    r[tl] += self.gamma * tv
                B["rew"][t], B["done"][t] = r, done
                self.obs = obs.copy()
            last = self.vf(torch.from_numpy(self._norm(self.obs).astype(np.float32))).squeeze(-1).numpy()
        adv = np.zeros((T, n), np.float32)
        gae = np.zeros(n, np.float32)
        for t in reversed(range(T)):
            nv = last if t == T - 1 else B["val"][t + 1]
            nonterm = 1.0 - B["done"][t]
            delta = B["rew"][t] + self.gamma * nv * nonterm - B["val"][t]
            gae = delta + self.gamma * self.lam * nonterm * gae
            adv[t] = gae
        ret = adv + B["val"]
        t1 = time.perf_counter()
        total = T * n
        obs_t = torch.from_numpy(B["obs"].reshape(total, d)); act_t = torch.from_numpy(B["act"].reshape(total, a))
        lp_t = torch.from_numpy(B["logp"].reshape(total)); adv_t = torch.from_numpy(adv.reshape(total))
        ret_t = torch.from_numpy(ret.reshape(total))
        bs = min(self.bs, total)
        for _ in range(self.epochs):
            perm = torch.randperm(total)
            for s in range(0, total, bs):
                idx = perm[s:s + bs]
                ad = adv_t[idx]
                if len(idx) > 1:
                    ad = (ad - ad.mean()) / (ad.std() + 1e-8)
                mean, val = self.pi(obs_t[idx]), self.vf(obs_t[idx]).squeeze(-1)
                ratio = torch.exp(self._logp(mean, act_t[idx]) - lp_t[idx])
                pl = -torch.min(ad * ratio, ad * torch.clamp(ratio, 1 - self.clip, 1 + self.clip)).mean()
                vl = torch.nn.functional.mse_loss(ret_t[idx], val)
                el = -(0.5 + 0.5 * math.log(2 * math.pi) + self.log_std).sum()
                loss = pl + self.ent * el + self.vfc * vl
                self.opt.zero_grad(set_to_none=True)
                loss.backward()
                torch.nn.utils.clip_grad_norm_(self.params, self.mgn)
                self.opt.step()
        t2 = time.perf_counter()
        self.rollout_s += t1 - t0
        self.update_s += t2 - t1
        self.samples += total
