"""CPU checks of the drop-in boundary: libfwsim.so loads without a GPU, exports every symbol the headers
declare, agrees with the ctypes mirror on struct sizes, and fails loudly (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

import pyflyt_drone_b200 as fw
from pyflyt_drone_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared(header):
    src = open(os.path.join(ROOT, "include", header)).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b((?:fw|ppo)_[a-z0-9_]+)\s*\(", src)))


def test_library_loads_and_exports_every_declared_symbol():
    lib = _lib.load()
    for header in sorted(os.listdir(os.path.join(ROOT, "include"))):
        names = _declared(header)
        assert names, header
        for n in names:
            assert hasattr(lib, n), f"{header}: {n} not exported by libfwsim.so"


def test_abi_version_and_struct_size():
    lib = _lib.load()
    assert lib.fw_abi_version() == 11 == _lib.ABI_VERSION
    assert lib.fw_config_size() == C.sizeof(fw.config.FwConfigC)


def test_bound_symbol_table_covers_header():
    bound = {s[0] for s in _lib.SYMBOLS}
    assert set(_declared("fwsim.h")) <= bound


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    lib = _lib.load()
    h = C.c_void_p()
    rc = lib.fw_create(C.byref(fw.waypoints_v3().to_c()), 4, 0, 0, 0, C.byref(h))
    assert rc == -2 and b"no CPU fallback" in lib.fw_last_error()
    from pyflyt_drone_b200.vec_env import FixedwingVecEnv
    with pytest.raises(_lib.FwError):
        FixedwingVecEnv(4)


def test_null_arguments_are_rejected_without_a_device():
    lib = _lib.load()
    assert lib.fw_create(None, 4, 0, 0, 0, None) == -1
    assert lib.fw_step(None, None, None, None, None, None, None) == -1
    assert lib.fw_destroy(None) == 0


def test_round2_entry_points_validate_arguments_without_a_device():
    """The entry points added in round 2 reject bad arguments before touching CUDA (the same error contract as the rest)."""
    lib = _lib.load()
    assert lib.fw_render(None, 0, 16, 16, None, None, None, None) == -1
    assert lib.fw_set_obs_accumulator(None, None) == -1
    assert lib.fw_observe_host(None, None) == -1
    assert lib.ppo_moments_finalize(None, 64, 10, 28, None, None, None) == -1
    assert lib.ppo_peer_alloc(None) == -1 and lib.ppo_peer_export(None, None) == -1 and lib.ppo_peer_import(None, None) == -1
    assert lib.ppo_peer_free(None) == 0 and lib.ppo_peer_close(None) == 0
    assert lib.ppo_peer_bytes() > 2 * 8 * 12361 * 8                      # two parities x eight ranks x the gradient as 8-byte pairs
    one = C.c_void_p(16)                                                 # any non-null address: never dereferenced on these paths
    args = [one, 28, 4, one, one, one, one, one, one, 128, 1, 0.2, 0.0, 0.5, one, one, 3e-4, 0.9, 0.999, 1e-5, 0.5, one, one]
    # world / rank out of range, missing peer table
    assert lib.ppo_window_update_p2p_a(*args, one, one, one, 9, 0, None, None, None) == -1
    assert lib.ppo_window_update_p2p_a(*args, one, one, one, 2, 2, None, None, None) == -1
    assert lib.ppo_window_update_p2p_a(*args, one, one, one, 2, 0, None, None, None) == -1
    assert lib.ppo_minibatch_steps_p2p_a(*args, one, one, 9, 0, None, None, None) == -1
    assert lib.ppo_minibatch_steps_p2p_a(*args, one, one, 2, 0, None, None, None) == -1
    assert b"peer" in lib.fw_last_error() or b"world" in lib.fw_last_error()
    # window sizes and minibatch sizes outside the documented ranges
    bad = list(args); bad[10] = 17
    assert lib.ppo_window_update_a(*bad, one, one, one, None) == -1
    bad = list(args); bad[9] = 5000
    assert lib.ppo_minibatch_steps_a(*bad, one, one, None) == -1
