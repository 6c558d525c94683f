import sys
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import torch
from pyflyt_drone_b200.ppo import PPO
from pyflyt_drone_b200.vec_env import FixedwingVecEnv
from test_ppo_gpu import _torch_minibatch_grad
env = FixedwingVecEnv(1000, preset="waypoints_v3", seed=5)
m = PPO("MlpPolicy", env, n_steps=16, batch_size=4000, n_epochs=2, seed=5, ent_coef=0.001)
with torch.no_grad():
    m.policy.theta.add_(0.05 * torch.randn(m.policy.count, device=m.device, generator=m._gen))
m.collect_rollouts()
with torch.no_grad():
    m.policy.theta.add_(0.02 * torch.randn(m.policy.count, device=m.device, generator=m._gen))
for batch in (128, 1000):
    idx = torch.randperm(16000, device=m.device, generator=m._gen)[:batch]
    ref, pl, vl, kl = _torch_minibatch_grad(m, idx)
    got = torch.zeros_like(ref); stats = torch.zeros(8, device=m.device)
    m._minibatch_grad_kernel(idx, got, stats)
    torch.cuda.synchronize()
    print("batch", batch, "stats", stats.tolist(), "ref pl vl kl", pl, vl, kl)
    for name, (a, b, shp) in m.policy.slices.items():
        r, g = ref[a:b], got[a:b]
        rel = float((g - r).norm() / (r.norm() + 1e-12)); cos = float(torch.dot(g, r) / (g.norm() * r.norm() + 1e-20))
        print(f"  {name:20s} |ref| {float(r.norm()):.4e} |got| {float(g.norm()):.4e} rel {rel:.3e} cos {cos:.5f}")
    # transposed hypothesis for W2 / W1
    for name in ("pi.2.weight", "vf.2.weight"):
        a, b, shp = m.policy.slices[name]
        r, g = ref[a:b].view(shp), got[a:b].view(shp)
        print("   ", name, "vs transposed: rel", float((g.t() - r).norm() / r.norm()))
env.close()
