"""CPU-only checks of the drop-in boundary: the C-ABI library loads, exports every symbol include/gpd.h declares,
and fails loudly (no CPU fallback) when there is no CUDA device.  No compute calls."""
import ctypes as C
import os
import re

import pytest

import gpd_b200  # noqa: F401
from gpd_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "gpd.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(gpd_[a-z0-9_]+)\s*\(", hdr)))


def test_library_exports_every_declared_symbol():
    if not os.path.exists(_lib.LIB_PATH):
        pytest.skip("libgpd_b200.so not built (run __graft_entry__.build())")
    L = C.CDLL(_lib.LIB_PATH)
    declared = _declared_symbols()
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(L, name), f"{name} declared in include/gpd.h but not exported"
    assert sorted(_lib.SYMBOLS) == declared
    L.gpd_version.restype = C.c_int
    assert L.gpd_version() == 200


def test_struct_layouts_match_header_sizes():
    # gpd_drone_params: 2 int32 + 43 doubles; gpd_pid_params: 36 doubles
    assert C.sizeof(_lib.DroneParamsC) == 8 + 8 * (3 + 6 + 2 + 3 + 3 + 3 + 3 + 8 + 12)
    assert C.sizeof(_lib.PidParamsC) == 8 * (18 + 4 + 12 + 2)
    assert C.sizeof(_lib.ConfigC) % 8 == 0


def test_no_cpu_fallback():
    if not os.path.exists(_lib.LIB_PATH):
        pytest.skip("libgpd_b200.so not built")
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from gpd_b200.envs import HoverAviary
    with pytest.raises(_lib.GpdError) as ei:
        HoverAviary(num_envs=2)
    assert ei.value.code in (-2, -3)


def test_missing_library_fails_loudly(tmp_path):
    with pytest.raises(_lib.GpdLibraryError):
        _lib.load(str(tmp_path / "nope.so"))


def test_product_never_touches_the_oracle():
    """The oracle is test infrastructure: nothing under the package may import, link or mention it."""
    pkg = os.path.join(ROOT, "gym-pybullet-drones-routing_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".inl", ".sh")):
                txt = open(os.path.join(dp, f)).read().lower()
                assert "oracle" not in txt, f"{os.path.join(dp, f)} mentions the oracle"


def test_create_validates_arguments_before_touching_the_device():
    """gpd_create rejects bad configurations with GPD_ERR_INVALID and a message (never exit(), SURVEY 8b "Errors") — the
    checks run before the first CUDA call, so they are observable without a GPU.  A valid configuration then fails on the
    missing device with a CUDA error, not with a CPU fallback."""
    import numpy as np
    from gpd_b200._lib import GpdError
    from gpd_b200.params import load_drone_params
    from gpd_b200.sim import BatchedSim
    from gpd_b200.utils.enums import DroneModel
    dp = load_drone_params(DroneModel.CF2X)
    tgt = np.array([[0.0, 0.0, 1.0]])
    base = dict(num_envs=4, num_drones=1, env_kind="hover", action_type="rpm", target_pos=tgt)
    cases = [
        (dict(num_drones=300, env_kind="multihover", target_pos=np.zeros((300, 3))), "num_drones must be in [1, 256]"),
        (dict(num_envs=0), "num_envs must be >= 1"),
        (dict(threads_per_block=48), "threads_per_block"),
        (dict(num_drones=2, target_pos=np.zeros((2, 3))), "HoverAviary is single-drone"),
        (dict(action_type="ctrl_rpm"), "GPD_ENV_CTRL"),
        (dict(target_pos=None), "target_pos is required"),
        (dict(physics_flags=64), "bad physics_flags"),
    ]
    for over, msg in cases:
        kw = dict(base, **over)
        E, N = kw.pop("num_envs"), kw.pop("num_drones")
        with pytest.raises(GpdError) as ei:
            BatchedSim(dp, E, N, **kw)
        assert msg in str(ei.value), (over, str(ei.value))
        assert ei.value.args and "-1" in str(ei.value)          # GPD_ERR_INVALID
    with pytest.raises(ValueError, match="not divisible"):      # the reference's own ValueError (BaseAviary.py:79-80)
        BatchedSim(dp, 4, 1, env_kind="hover", action_type="rpm", target_pos=tgt, pyb_freq=240, ctrl_freq=35)
    import torch
    if not torch.cuda.is_available():
        with pytest.raises(Exception) as ei:
            BatchedSim(dp, 4, 1, env_kind="hover", action_type="rpm", target_pos=tgt)
        assert "CUDA" in str(ei.value) or "cuda" in str(ei.value) or "device" in str(ei.value)
