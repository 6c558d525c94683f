"""Memory-safety and race checks that do not need compute-sanitizer (closed on this GPU pool: profiles/r2_sanitizer.md).

  * guard bands: every output buffer of a step / reset is a window inside a larger sentinel-filled allocation; a kernel
    that writes one element before or after its window (partial last warp, bulk-store row count, odd batch, pair lanes)
    trips the sentinel check;
  * determinism: the warp-cooperative camera rasterises into a shared depth row with atomicMin and hands data between
    lanes through shared memory -- a missing __syncwarp shows up as run-to-run differences, so the same step from the same
    state must reproduce bit for bit, twice.
"""
import numpy as np
import pytest
import torch

import pyflyt_drone_b200 as fw
from pyflyt_drone_b200 import _lib
from pyflyt_drone_b200.vec_env import FixedwingVecEnv
import ctypes as C

pytestmark = pytest.mark.gpu
PAD = 4096


def _guarded(n_elems, dtype, sentinel):
    full = torch.full((n_elems + 2 * PAD,), sentinel, dtype=dtype, device="cuda")
    return full, full[PAD:PAD + n_elems]


@pytest.mark.parametrize("preset,kw", [("waypoints_v3", {}), ("waypoints_v3", {"packed_pairs": 1}), ("lowlevel", {}),
                                       ("lowlevel", {"packed_pairs": 1}), ("waypoint_objlock", {}), ("objlock_duck", {})])
def test_outputs_stay_inside_their_windows(preset, kw):
    n = 1000 + 37                                       # partial last CTA, partial last warp, odd
    env = FixedwingVecEnv(n, config=fw.make_config(preset, **kw), seed=9)
    D, A = env.obs_dim, env.act_dim
    fo, obs = _guarded(n * D, torch.float32, 7777.0)
    ft, term = _guarded(n * D, torch.float32, 7777.0)
    fr, rew = _guarded(n, torch.float32, 7777.0)
    ff, flags = _guarded(n, torch.uint8, 0xAB)
    st = C.c_void_p(int(torch.cuda.current_stream().cuda_stream))
    p = lambda t: C.c_void_p(t.data_ptr())           # noqa: E731
    _lib.check(env.lib.fw_reset(env._h, None, p(obs), st))
    act = (torch.rand((n, A), device="cuda") * 2 - 1).contiguous()
    for _ in range(6):
        _lib.check(env.lib.fw_step(env._h, p(act), p(obs), p(rew), p(flags), p(term), st))
    mask = (torch.arange(n, device="cuda") % 3 == 0).to(torch.uint8)
    _lib.check(env.lib.fw_reset(env._h, p(mask), p(obs), st))
    torch.cuda.synchronize()
    for name, full, fill in (("obs", fo, 7777.0), ("term_obs", ft, 7777.0), ("rew", fr, 7777.0), ("flags", ff, 0xAB)):
        head, tail = full[:PAD], full[-PAD:]
        assert bool((head == fill).all()) and bool((tail == fill).all()), f"{preset}: {name} written outside its window"
    assert bool(torch.isfinite(obs).all()) and bool((obs != 7777.0).all()), "an observation element was never written"
    assert bool((rew != 7777.0).all()) and bool((flags != 0xAB).all())
    env.close()


@pytest.mark.parametrize("preset", ["waypoint_objlock", "objlock_duck", "waypoints_v3"])
def test_step_is_bitwise_reproducible_from_the_same_state(preset):
    n = 2048
    cfg = fw.make_config(preset)
    a = FixedwingVecEnv(n, config=cfg, seed=4)
    a.reset()
    rng = np.random.default_rng(2)
    for _ in range(30):                                  # spread the fleet out: obstacles in view, frames due
        a.step_arrays(rng.uniform(-1, 1, (n, 4)).astype(np.float32))
    st = a.get_state()
    act = rng.uniform(-1, 1, (n, 4)).astype(np.float32)
    outs = []
    for _ in range(3):
        a.set_state(st)
        rows = []
        for k in range(6):                               # cam_interval 12 substeps: at least three frames per env
            o, r, f, t = a.step_arrays(act)
            rows.append((o.copy(), r.copy(), f.copy()))
        outs.append(rows)
    for rows in outs[1:]:
        for (o0, r0, f0), (o1, r1, f1) in zip(outs[0], rows):
            assert np.array_equal(f0, f1) and np.array_equal(r0.view(np.uint32), r1.view(np.uint32))
            assert np.array_equal(o0.view(np.uint32), o1.view(np.uint32))
    a.close()
