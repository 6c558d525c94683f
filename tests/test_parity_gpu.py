"""GPU parity: the sm_100a kernels (through the C ABI of libfwsim.so) against the fp64 oracle.

Tolerance (BASELINE.json north_star): per-step state within 1e-4 relative (fp32 vs fp64, single step from an
identical injected state), reward and termination/truncation/info flags matching exactly on those steps,
trajectory divergence over a 1 s horizon (30 agent steps = 240 substeps) reported.  "Relative" is taken
norm-wise per observation group with a floor of 1.0 on the reference norm, i.e.
|x_gpu - x_ref|_inf <= 1e-4 * max(|x_ref|_inf, 1).
"""
import os

import numpy as np
import pytest

import pyflyt_drone_b200 as fw
from pyflyt_drone_b200.config import FLAG_TERM, FLAG_TRUNC

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

pytestmark = pytest.mark.gpu

RTOL = 1e-4
GROUPS_EULER = {"ang_vel": slice(0, 3), "ang_pos": slice(3, 6), "lin_vel": slice(6, 9), "lin_pos": slice(9, 12),
                "action": slice(12, 16), "aux": slice(16, 22), "delta0": slice(22, 25), "delta1": slice(25, 28)}


def group_err(got, ref, groups=GROUPS_EULER):
    out = {}
    for k, sl in groups.items():
        d = np.abs(got[:, sl] - ref[:, sl]).max(axis=1)
        scale = np.maximum(np.abs(ref[:, sl]).max(axis=1), 1.0)
        out[k] = float((d / scale).max()) if len(d) else 0.0
    return out


def wrap_angles(obs):
    o = obs.copy()
    return o


@pytest.fixture(scope="module")
def fo(oracle_mod):
    return oracle_mod


def make_pair(fo, n, cfg, seed=7, env_id0=0):
    from pyflyt_drone_b200.vec_env import FixedwingVecEnv
    env = FixedwingVecEnv(n, config=cfg, seed=seed, env_id0=env_id0)
    orc = fo.OracleVecEnv(cfg.as_dict(), n, seed=seed, env_id0=env_id0)
    return env, orc


def angle_safe(og, oc):
    """yaw/roll wrap at +-pi: compare angles modulo 2 pi."""
    og = og.copy()
    d = og[:, 3:6] - oc[:, 3:6]
    og[:, 3:6] = oc[:, 3:6] + (d + np.pi) % (2 * np.pi) - np.pi
    return og


def test_reset_parity(fo):
    cfg = fw.waypoints_v3()
    env, orc = make_pair(fo, 257, cfg)
    og, oc = env.reset(), orc.reset()
    errs = group_err(og, oc)
    assert max(errs.values()) < RTOL, errs
    sg, sc = env.get_state(), orc.get_state()
    assert np.abs(sg["targets"] - sc["targets"]).max() < 1e-4 * 90
    assert np.array_equal(sg["physics_steps"], sc["physics_steps"]) and np.array_equal(sg["episode"], sc["episode"])
    env.close()


@pytest.mark.parametrize("variant", ["sparse_euler", "dense_quat", "noise", "wind", "signs"])
def test_single_step_parity_from_injected_state(fo, variant):
    kw = dict(noise_ratio=0.0)
    groups = GROUPS_EULER
    if variant == "dense_quat":
        kw.update(sparse_reward=0, angle_repr=1)
        groups = {"ang_vel": slice(0, 3), "quat": slice(3, 7), "lin_vel": slice(7, 10), "lin_pos": slice(10, 13),
                  "action": slice(13, 17), "aux": slice(17, 23), "delta0": slice(23, 26), "delta1": slice(26, 29)}
    elif variant == "noise":
        kw.update(noise_ratio=0.02, sparse_reward=0)
    elif variant == "signs":
        kw.update(ail_left_sign=-1.0, ail_right_sign=1.0, pitch_sign=-1.0, freestream_3d=0, cd90_degrees=0)
    wind = None
    if variant == "wind":
        wind = dict(enabled=True, mode="gust_sine", wind_enu_mps_range=[[-5, 5], [-5, 5], [-0.5, 0.5]],
                    gust_amp_enu_mps_range=[[0, 3], [0, 3], [0, 0.3]], gust_freq_hz=0.2,
                    randomize_on_reset=True, randomize_gust_phase=True)
    cfg = fw.waypoints_v3(wind=wind, **kw)
    if variant == "wind":
        cfg = cfg.replace(wind_start_substep=0)     # env-hook flavour: wind acts during the warm-up too
    N = 512
    env, orc = make_pair(fo, N, cfg)
    env.reset(); orc.reset()
    rng = np.random.default_rng(1)
    worst = {}
    n_done = 0
    for k in range(40):
        # free-run the oracle, inject its state, step both once
        env.set_state(orc.get_state())
        a = rng.uniform(-1, 1, (N, 4)).astype(np.float32)
        og, rg, fg, tg = env.step_arrays(a)
        og, rg, fg, tg = og.copy(), rg.copy(), fg.copy().astype(np.int32), tg.copy()
        oc, rc, fc, tc = orc.step(a.astype(np.float64))
        assert np.array_equal(fg, fc), (k, np.nonzero(fg != fc))
        done = (fc & (FLAG_TERM | FLAG_TRUNC)) != 0
        n_done += int(done.sum())
        if variant != "dense_quat":
            og, tg = angle_safe(og, oc), angle_safe(tg, tc)
        e = group_err(og, oc, groups)
        if done.any():
            et = group_err(tg[done], tc[done], groups)
            e = {g: max(e[g], et[g]) for g in e}
        for g, v in e.items():
            worst[g] = max(worst.get(g, 0.0), v)
        assert np.abs(rg - rc).max() <= 1e-4 * max(1.0, np.abs(rc).max()), k
        sparse_like = np.isin(np.round(rc, 6), (-0.1, 100.0, -100.0))
        assert np.abs(rg - rc)[sparse_like].max(initial=0.0) < 1e-6
    print(f"\n[{variant}] single-step worst norm-wise relative error per group: "
          + ", ".join(f"{g}={v:.2e}" for g, v in worst.items()) + f"; done events {n_done}")
    assert max(worst.values()) < RTOL, worst
    env.close()


def test_one_second_free_run_divergence_is_reported_and_small(fo):
    cfg = fw.waypoints_v3(noise_ratio=0.0)
    N = 1024
    env, orc = make_pair(fo, N, cfg)
    env.reset(); orc.reset()
    rng = np.random.default_rng(2)
    alive = np.ones(N, bool)
    div = []
    for k in range(30):
        a = rng.uniform(-1, 1, (N, 4)).astype(np.float32)
        og, rg, fg, _ = env.step_arrays(a)
        oc, rc, fc, _ = orc.step(a.astype(np.float64))
        same = fg.astype(np.int32) == fc
        alive &= same & ((fc & 3) == 0)
        if alive.any():
            div.append(float(np.abs(og[alive][:, 9:12] - oc[alive][:, 9:12]).max()))
    print(f"\n1 s free-run (30 agent steps, 240 substeps, {int(alive.sum())}/{N} envs alive in both): "
          f"max position divergence {div[-1]:.3e} m; per-step trace {['%.1e' % d for d in div[::5]]}")
    assert alive.mean() > 0.9
    assert div[-1] < 5e-3


def test_ragged_batch_sizes_and_terminal_observation(fo):
    cfg = fw.waypoints_v3(noise_ratio=0.0)
    for n in (1, 31, 33, 65, 100):
        env, orc = make_pair(fo, n, cfg, seed=3)
        og, oc = env.reset(), orc.reset()
        assert og.shape == (n, 28) and max(group_err(og, oc).values()) < RTOL
        st = orc.get_state()
        st["pos"][:, 2] = np.linspace(0.03, 0.5, n)    # most of them hit the ground this step
        orc.set_state(st); env.set_state(st)
        a = np.zeros((n, 4), np.float32)
        obs, rew, dones, infos = env.step(a)
        oc, rc, fc, tc = orc.step(a.astype(np.float64))
        assert np.array_equal(env._h_flags.astype(np.int32), fc)
        assert dones.sum() >= 1
        for i in np.nonzero(dones)[0]:
            assert infos[i]["collision"] and not infos[i]["TimeLimit.truncated"]
            assert np.abs(infos[i]["terminal_observation"] - tc[i]).max() < 1e-3
        # auto-reset rows hold the next episode's first observation
        assert max(group_err(angle_safe(obs, oc), oc).values()) < RTOL
        env.close()


def test_random_action_lane_matches_oracle_rollout(fo):
    for preset in ("physics_only", "waypoints_v3"):
        cfg = fw.make_config(preset, noise_ratio=0.02)
        N = 300
        env, orc = make_pair(fo, N, cfg, seed=5, env_id0=1000)
        env.reset() if cfg.task else None
        orc.reset()
        env.step_random(8)
        orc.rollout_random(8)
        sg, sc = env.get_state(), orc.get_state()
        assert np.array_equal(sg["episode"], sc["episode"]) and np.array_equal(sg["step_count"], sc["step_count"])
        assert np.abs(sg["pos"] - sc["pos"]).max() < 2e-3
        assert np.abs(sg["quat"] - sc["quat"]).max() < 1e-3
        env.close()


def test_tensor_lane_equals_host_lane(fo):
    import torch
    from pyflyt_drone_b200.vec_env import FixedwingVecEnv
    cfg = fw.waypoints_v3()
    a_env, b_env = FixedwingVecEnv(200, config=cfg, seed=1), FixedwingVecEnv(200, config=cfg, seed=1)
    oa = a_env.reset()
    ob = b_env.reset_tensor().cpu().numpy()
    assert np.array_equal(oa, ob)
    rng = np.random.default_rng(0)
    for _ in range(5):
        a = rng.uniform(-1, 1, (200, 4)).astype(np.float32)
        o1, r1, f1, _ = a_env.step_arrays(a)
        o2, r2, f2 = b_env.step_tensor(torch.from_numpy(a).cuda())
        torch.cuda.synchronize()
        assert np.array_equal(o1, o2.cpu().numpy()) and np.array_equal(r1, r2.cpu().numpy())
        assert np.array_equal(f1, f2.cpu().numpy())
    a_env.close(); b_env.close()


@pytest.mark.parametrize("preset", ["waypoints_v3", "waypoint_objlock", "lowlevel", "objlock_duck"])
def test_sharding_is_split_invariant(fo, preset):
    """Global env ids key the RNG: two half batches reproduce one full batch bit for bit (multi-GPU contract) -- targets,
    wind, obstacles, duck positions, camera frames and vision history included."""
    from pyflyt_drone_b200.vec_env import FixedwingVecEnv
    cfg = fw.make_config(preset) if preset != "objlock_duck" else fw.make_config(preset, num_obstacles=10)
    full = FixedwingVecEnv(128, config=cfg, seed=2)
    lo, hi = FixedwingVecEnv(64, config=cfg, seed=2, env_id0=0), FixedwingVecEnv(64, config=cfg, seed=2, env_id0=64)
    assert np.array_equal(full.reset(), np.concatenate([lo.reset(), hi.reset()]))
    full.step_random(40); lo.step_random(40); hi.step_random(40)
    sf, sl, sh = full.get_state(), lo.get_state(), hi.get_state()
    keys = ["pos", "quat", "episode", "targets", "wind", "act"]
    keys += [k for k in ("duck", "obst", "ol_f", "ol_i", "vis_hist") if k in sf]
    for k in keys:
        assert np.array_equal(sf[k], np.concatenate([sl[k], sh[k]])), k
    for e in (full, lo, hi):
        e.close()


def test_fused_and_graph_rollouts_are_bitwise_identical_to_single_step_launches():
    """steps_per_launch > 1 keeps the state in registers between env-steps, and the CUDA-graph path replays
    the same launches: both must reproduce the one-launch-per-step trajectory bit for bit."""
    import torch
    from pyflyt_drone_b200.vec_env import FixedwingVecEnv
    for preset in ("physics_only", "waypoints_v3", "waypoint_objlock", "lowlevel", "objlock_duck"):
        cfg = fw.make_config(preset)
        ref = [FixedwingVecEnv(1000, config=cfg, seed=8, env_id0=k * 1000) for k in range(3)]
        fused = [FixedwingVecEnv(1000, config=cfg, seed=8, env_id0=k * 1000) for k in range(3)]
        graph = [FixedwingVecEnv(1000, config=cfg, seed=8, env_id0=k * 1000) for k in range(3)]
        for e in ref:
            e.step_random(48)
        FixedwingVecEnv.rollout_random(fused, n_launches=3 * 6, steps_per_launch=8, use_graph=False)
        FixedwingVecEnv.rollout_random(graph, n_launches=3 * 48 + 0, steps_per_launch=1, use_graph=True)
        torch.cuda.synchronize()
        for k in range(3):
            a, b, c = ref[k].get_state(), fused[k].get_state(), graph[k].get_state()
            keys = ["pos", "quat", "vel", "omega", "act", "episode", "step_count", "physics_steps", "targets"]
            keys += [x for x in ("duck", "obst", "ol_f", "ol_i", "vis_hist") if x in a]
            for key in keys:
                assert np.array_equal(a[key], b[key]), (preset, "fused", key)
                assert np.array_equal(a[key], c[key]), (preset, "graph", key)
        sa, sb = ref[0].episode_stats(), fused[0].episode_stats()
        assert sa["episodes"] == sb["episodes"] and (sa["episodes"] > 0 or preset in ("lowlevel", "objlock_duck"))
        for e in ref + fused + graph:
            e.close()


def test_full_size_invariants_65536():
    """BASELINE config 2 size: properties that do not need the oracle."""
    from pyflyt_drone_b200.vec_env import FixedwingVecEnv
    env = FixedwingVecEnv(65536, preset="waypoints_v3", seed=0)
    env.reset()
    total_done = 0
    for s in range(0, 120, 20):
        rew, flags = env.step_random(20, with_outputs=True)
    import torch
    torch.cuda.synchronize()
    st = env.get_state()
    q = np.linalg.norm(st["quat"], axis=1)
    assert np.abs(q - 1).max() < 1e-5
    for k in ("pos", "vel", "omega", "act"):
        assert np.isfinite(st[k]).all()
    assert np.abs(st["vel"]).max() <= 100.0 and np.abs(st["omega"]).max() <= 100.0   # Bullet's clamp
    assert np.linalg.norm(st["pos"], axis=1).max() <= 100.0 + 100 * 8 / 240 + 1e-3    # dome + one step
    assert np.abs(st["act"][:, :5]).max() <= 1.0 + 1e-6 and st["act"][:, 5].min() >= -0.1
    stats = env.episode_stats()
    assert stats["episodes"] == st["episode"].astype(np.int64).sum()
    assert stats["collisions"] + stats["out_of_bounds"] >= stats["episodes"] * 0.99
    env.close()


def test_standard_layout_kernels_match_the_generic_kernels():
    """The kernels specialised for the reference aircraft's layout (forward units +x, lift +z / +y, xz-symmetric
    inertia) skip structural zeros only.  Single agent steps from identical injected states agree with the generic
    kernels to 2e-5 (norm-wise, floor 1; fp32 re-association), flags and sparse rewards exactly; free-running for 40
    steps the two stay together except where the dynamics amplify rounding (median 1e-6, 99th percentile 1e-3)."""
    from pyflyt_drone_b200.vec_env import FixedwingVecEnv
    N = 2048
    rng = np.random.default_rng(3)
    std = FixedwingVecEnv(N, config=fw.waypoints_v3(noise_ratio=0.02), seed=5)
    gen = FixedwingVecEnv(N, config=fw.waypoints_v3(noise_ratio=0.02, force_generic_kernel=1), seed=5)
    std.reset(); gen.reset()
    keys = ("pos", "vel", "omega", "quat", "act")

    def rel(a, b):
        return np.abs(a - b).max(axis=1) / np.maximum(np.abs(b).max(axis=1), 1.0)

    for t in range(40):                                   # free run, same actions
        a = rng.uniform(-1, 1, (N, 4)).astype(np.float32)
        _, _, f1, _ = std.step_arrays(a)
        f1 = f1.copy()
        _, _, f0, _ = gen.step_arrays(a)
        assert (f1 != f0).mean() < 0.002, t
    s1, s0 = std.get_state(), gen.get_state()
    same_ep = s1["episode"] == s0["episode"]
    for k in keys:
        d = rel(s1[k][same_ep], s0[k][same_ep])
        assert np.median(d) < 1e-6 and np.quantile(d, 0.99) < 1e-3, (k, np.median(d), np.quantile(d, 0.99))
    for t in range(10):                                   # single steps from identical states
        std.set_state(gen.get_state())
        a = rng.uniform(-1, 1, (N, 4)).astype(np.float32)
        o1, r1, f1, _ = std.step_arrays(a)
        o1, r1, f1 = o1.copy(), r1.copy(), f1.copy()
        o0, r0, f0, _ = gen.step_arrays(a)
        assert np.array_equal(f1, f0), t
        assert np.abs(r1 - r0).max() < 1e-6, t            # sparse reward: -0.1 / +-100
        s1, s0 = std.get_state(), gen.get_state()
        for k in keys:
            assert rel(s1[k], s0[k]).max() < 2e-5, (t, k, rel(s1[k], s0[k]).max())
    std.close(); gen.close()


def test_page_locked_caller_actions_are_read_in_place():
    """fw_step_host stages pageable action arrays chunk by chunk, but reads a page-locked caller buffer (torch
    pin_memory, cudaHostAlloc) directly from the kernel: both routes must give bitwise identical steps."""
    import torch
    from pyflyt_drone_b200.vec_env import FixedwingVecEnv
    N = 20000                                             # above the chunking threshold of the host lane
    a_env = FixedwingVecEnv(N, preset="waypoints_v3", seed=8)
    b_env = FixedwingVecEnv(N, preset="waypoints_v3", seed=8)
    a_env.reset(); b_env.reset()
    rng = np.random.default_rng(0)
    pinned = torch.empty((N, 4), dtype=torch.float32, pin_memory=True)
    for _ in range(5):
        a = rng.uniform(-1, 1, (N, 4)).astype(np.float32)
        pinned.numpy()[:] = a
        o1, r1, f1, _ = a_env.step_arrays(a)
        o2, r2, f2, _ = b_env.step_arrays(pinned.numpy())
        assert np.array_equal(o1, o2) and np.array_equal(r1, r2) and np.array_equal(f1, f2)
    # an unaligned view of pinned memory falls back to the staging copy
    wide = torch.empty((N * 4 + 1,), dtype=torch.float32, pin_memory=True)
    off = wide.numpy()[1:].reshape(N, 4)
    a = rng.uniform(-1, 1, (N, 4)).astype(np.float32)
    off[:] = a
    o1, r1, f1, _ = a_env.step_arrays(a)
    o2, r2, f2, _ = b_env.step_arrays(off)
    assert np.array_equal(o1, o2) and np.array_equal(r1, r2) and np.array_equal(f1, f2)
    a_env.close(); b_env.close()


def test_host_lane_routes_agree(tmp_path):
    """The default host lane lets the kernel write observations straight into pinned host memory; FWSIM_HOST_ZEROCOPY_OBS=0
    (read once per process) restores the device-buffer + copy-engine route, and FWSIM_HOST_CHUNKS changes the chunking.
    All of them must return bitwise identical steps, terminal observations included."""
    import subprocess
    import sys
    from pyflyt_drone_b200.vec_env import FixedwingVecEnv
    script = (
        "import sys, numpy as np\n"
        f"sys.path.insert(0, {str(ROOT)!r})\n"
        "from pyflyt_drone_b200.vec_env import FixedwingVecEnv\n"
        "env = FixedwingVecEnv(20000, preset='waypoints_v3', seed=8)\n"
        "env.reset(); rng = np.random.default_rng(0); out = {}\n"
        "for k in range(30):\n"
        "    a = rng.uniform(-1, 1, (20000, 4)).astype(np.float32); a[:, 1] = -1.0\n"
        "    o, r, f, t = env.step_arrays(a)\n"
        "    out[f'o{k}'], out[f'r{k}'], out[f'f{k}'] = o.copy(), r.copy(), f.copy()\n"
        "    done = (f & 3) != 0\n"
        "    out[f't{k}'] = t[done].copy()\n"
        "np.savez(sys.argv[1], **out)\n")
    results = []
    for name, extra in (("default", {}), ("dma", {"FWSIM_HOST_ZEROCOPY_OBS": "0"}),
                        ("dma4", {"FWSIM_HOST_ZEROCOPY_OBS": "0", "FWSIM_HOST_CHUNKS": "4"}), ("one", {"FWSIM_HOST_CHUNKS": "1"})):
        path = tmp_path / f"{name}.npz"
        env = dict(os.environ, **extra)
        subprocess.run([sys.executable, "-c", script, str(path)], check=True, env=env, timeout=300)
        results.append(np.load(path))
    ref = results[0]
    assert sum(len(ref[f"t{k}"]) for k in range(30)) > 0, "the scripted dive must end some episodes"
    for other in results[1:]:
        for key in ref.files:
            assert np.array_equal(ref[key], other[key]), key


def test_config0_single_env_10k_random_action_steps(fo):
    """BASELINE configs[0] / SURVEY 8(d) config 1: ONE Fixedwing-Waypoints env (FlattenWaypointEnv, euler, 8 targets, reach
    4 m, sparse reward, motor noise on), 10,100 steps with actions np.random.default_rng(0).uniform(-1, 1, (10100, 4)),
    re-reset on done.  The fp64 oracle free-runs that trajectory; before every step its state is injected into the
    one-env CUDA batch, both step with the same action, and observation / reward / flags (and the terminal observation and
    the post-reset observation of finished episodes) are compared at the single-step tolerance."""
    cfg = fw.waypoints_v3()
    env, orc = make_pair(fo, 1, cfg, seed=0)
    og, oc = env.reset(), orc.reset()
    assert max(group_err(og, oc).values()) < RTOL
    acts = np.random.default_rng(0).uniform(-1, 1, (10100, 4))
    rows, episodes, worst_rew = [], 0, 0.0
    for t in range(10100):
        env.set_state(orc.get_state())
        a = acts[t:t + 1]
        og, rg, fg, tg = env.step_arrays(a.astype(np.float32))
        oc, rc, fc, tc = orc.step(a)
        assert int(fg[0]) == int(fc[0]), (t, int(fg[0]), int(fc[0]))
        worst_rew = max(worst_rew, float(abs(rg[0] - rc[0]) / max(1.0, abs(rc[0]))))
        e = max(group_err(angle_safe(og.astype(np.float64), oc), oc).values())
        if fc[0] & (FLAG_TERM | FLAG_TRUNC):
            episodes += 1
            e = max(e, max(group_err(angle_safe(tg.astype(np.float64), tc), tc).values()))
        rows.append(e)
    rows = np.array(rows)
    flips = int((rows >= RTOL).sum())
    print(f"\n[configs[0]] 10,100 single-env steps, {episodes} episodes: median step error {np.median(rows):.1e}, "
          f"99.9th percentile {np.quantile(rows, 0.999):.1e}, steps past {RTOL:g}: {flips}, worst {rows.max():.1e}; "
          f"worst reward error {worst_rew:.1e}")
    assert episodes >= 20
    assert worst_rew <= 1e-4
    assert flips <= 5 and rows.max() < 5e-3          # stall-boundary branch flips (see tests/test_duck_gpu.py), if any
    env.close()


def test_sweep_end_points_invariants():
    """BASELINE configs[4] sweeps 1K..1M envs per GPU: oracle-free properties at both ends (1,024 and 1,048,576 envs),
    and the two sizes agree on the envs they share (the RNG is keyed by the global env id, not by the batch size)."""
    import torch
    from pyflyt_drone_b200.vec_env import FixedwingVecEnv
    small = FixedwingVecEnv(1024, preset="physics_only", seed=3)
    big = FixedwingVecEnv(1 << 20, preset="physics_only", seed=3)
    small.step_random(60); big.step_random(60)
    torch.cuda.synchronize()
    ss, sb = small.get_state(), big.get_state()
    for k in ("pos", "quat", "vel", "omega", "act", "episode", "physics_steps"):
        assert np.array_equal(ss[k], sb[k][:1024]), k
    assert np.abs(np.linalg.norm(sb["quat"], axis=1) - 1).max() < 1e-5
    for k in ("pos", "vel", "omega", "act"):
        assert np.isfinite(sb[k]).all(), k
    assert np.abs(sb["vel"]).max() <= 100.0 and np.abs(sb["omega"]).max() <= 100.0
    assert np.linalg.norm(sb["pos"], axis=1).max() <= 100.0 + 100 * 8 / 240 + 1e-3
    st = big.episode_stats()
    assert st["episodes"] == sb["episode"].astype(np.int64).sum() and st["episodes"] > 0
    small.close(); big.close()
