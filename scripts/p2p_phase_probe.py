"""Phase times of the single-CTA optimizer steps with the peer-memory gradient exchange (needs a -DUT_PROFILE build):
    FWSIM_LIB=build_ab/lib_prof.so torchrun --nproc-per-node 2 scripts/p2p_phase_probe.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from pyflyt_drone_b200 import _lib
from pyflyt_drone_b200.ppo import PPO, _p, _stream
from pyflyt_drone_b200.vec_env import FixedwingVecEnv
rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
env = FixedwingVecEnv(4096, preset="waypoints_v3", device=local, seed=1, env_id0=rank * 4096)
m = PPO("MlpPolicy", env, n_steps=128, batch_size=128, n_epochs=1, seed=1)
assert m._p2p is not None
m.collect_rollouts()
perm = torch.randperm(4096 * 128, device="cuda")
stats = torch.zeros(64, device="cuda")
b, pp = m.buf, m._p2p
def run(steps):
    _lib.check(m.lib.ppo_minibatch_steps_p2p_a(_p(m.policy.theta.data), m.d, m.a, _p(b["obs"]), _p(b["act"]), _p(b["logp"]), _p(b["adv"]),
               _p(b["ret"]), _p(perm), 128, steps, 0.2, 0.001, 0.5, _p(m._adam_m), _p(m._adam_v), 3e-4, 0.9, 0.999, 1e-5, 0.5,
               _p(m._adam_t), _p(m._grad_norm), _p(m._grad), _p(stats), world, pp["rank"], pp["peers"], _p(pp["seq"]), _stream()))
run(64); torch.cuda.synchronize(); dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); run(256); e1.record(); torch.cuda.synchronize()
if rank == 0:
    print(f"world {world}: {e0.elapsed_time(e1) * 1000 / 256:.1f} us per step")
    print("last step phases (ns): stage+advstats, tiles, readout+sums, exchange+norm, adam:", [int(x) for x in stats[8:13].tolist()], flush=True)
dist.barrier()
os._exit(0)
