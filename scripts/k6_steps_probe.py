"""Phase times of the single-CTA optimizer steps (needs a -DUT_PROFILE build: FWSIM_LIB=build_ab/lib_prof.so)."""
import sys
sys.path.insert(0, '/root/repo')
import torch
from pyflyt_drone_b200 import _lib
from pyflyt_drone_b200.ppo import PPO, _p, _stream
from pyflyt_drone_b200.vec_env import FixedwingVecEnv
env = FixedwingVecEnv(4096, preset="waypoints_v3", seed=1)
m = PPO("MlpPolicy", env, n_steps=128, batch_size=128, n_epochs=1, seed=1)
m.collect_rollouts()
perm = torch.randperm(4096 * 128, device="cuda")
stats = torch.zeros(64, device="cuda")
b = m.buf
def run(steps):
    _lib.check(m.lib.ppo_minibatch_steps_a(_p(m.policy.theta.data), m.d, m.a, _p(b["obs"]), _p(b["act"]), _p(b["logp"]), _p(b["adv"]),
               _p(b["ret"]), _p(perm), 128, steps, 0.2, 0.001, 0.5, _p(m._adam_m), _p(m._adam_v), 3e-4, 0.9, 0.999, 1e-5, 0.5,
               _p(m._adam_t), _p(m._grad_norm), _p(m._grad), _p(stats), _stream()))
run(64); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); run(256); e1.record(); torch.cuda.synchronize()
print(f"256 steps: {e0.elapsed_time(e1) * 1000 / 256:.1f} us per step")
print("last step phases (ns): stage+advstats, tiles, readout+sums, norm, adam:", [int(x) for x in stats[8:13].tolist()])
print("last tile (ns): gather wait+sync, M1, E1, M2, E2, loss+E3, M3/M4, E4:", [int(x) for x in stats[16:24].tolist()])
env.close()
