"""Pins the CPU oracle (oracle/gpd_oracle.c) to the golden vectors produced by the REFERENCE's own
Python (oracle/gen_golden.py).  CPU-only; the CUDA parity tests (-m gpu) then compare against both."""
import numpy as np
import pytest

from helpers import (S_ANGV, S_POS, S_QUAT, S_RATES, S_RPM, S_RPY, S_VEL, angle_err, case_setup, constants,
                     load_golden, make_oracle, quat_err, rel_err, traj_cases)
from gpd_b200.params import default_pid_params, load_drone_params
from gpd_b200.utils.enums import DroneModel
from oracle import oracle as orc

# float32 sub-expressions of ActionType.VEL go through BLAS sdot in the reference (BaseRLAviary.py:209-210),
# whose rounding is library-dependent: that one case is held to float32-level agreement only.
TOL = {}      # every trajectory at 1e-9 (the float32 norm of the VEL map is reproduced bit for bit)


@pytest.mark.parametrize("name", traj_cases())
def test_oracle_replays_reference_trajectory(name):
    g = load_golden(name)
    sim = make_oracle(case_setup(g))
    tol = TOL.get(name, 1e-9)
    acts = g["actions"]
    ck = {int(t): i for i, t in enumerate(g["ckpt_idx"])}
    osteps = {int(t): i for i, t in enumerate(g["obs_steps"])}
    assert np.max(np.abs(sim.obs[0].astype(np.float64) - g["obs0"])) <= 1e-7 * (tol / 1e-9)
    for t in range(acts.shape[0]):
        obs, r, te, tr = sim.step(acts[t][None])
        if t in ck:
            i = ck[t]
            ref = g["ckpt_state"][i]
            st = np.concatenate([sim.state20[0], sim.rpy_rates[0]], axis=1)
            assert rel_err(st[:, S_POS], ref[:, S_POS]) <= tol, (t, "pos")
            assert rel_err(st[:, S_VEL], ref[:, S_VEL]) <= tol, (t, "vel")
            assert rel_err(st[:, S_RATES], ref[:, S_RATES]) <= tol, (t, "rpy_rates")
            assert rel_err(st[:, S_ANGV], ref[:, S_ANGV]) <= tol, (t, "ang_v")
            assert quat_err(st[:, S_QUAT], ref[:, S_QUAT]) <= tol, (t, "quat")
            assert angle_err(st[:, S_RPY], ref[:, S_RPY]) <= tol * 10, (t, "rpy")
            assert rel_err(st[:, S_RPM], ref[:, S_RPM]) <= tol, (t, "rpm")
            assert abs(r[0] - g["ckpt_reward"][i]) <= tol * max(abs(g["ckpt_reward"][i]), 1e-3) * 10, (t, "reward")
            assert bool(te[0]) == bool(g["ckpt_terminated"][i]), (t, "terminated")
            assert bool(tr[0]) == bool(g["ckpt_truncated"][i]), (t, "truncated")
            assert sim.step_counter[0] == g["ckpt_counter"][i]
        if t in osteps:
            want = g["obs_rows"][osteps[t]]
            got = obs[0].astype(np.float64)
            scale = np.maximum(np.abs(want), 1.0)
            assert np.max(np.abs(got - want) / scale) <= max(1e-9, tol * 100), (t, "obs")


def test_survey_kat_a():
    """SURVEY Appendix C KAT-A: HoverAviary(CF2X, DYN, 240/30), 10x step([.1,-.2,.3,-.4] float32)."""
    sim = orc.OracleSim(load_drone_params(DroneModel.CF2X), 1)
    for _ in range(10):
        _, r, _, _ = sim.step(np.array([[[0.1, -0.2, 0.3, -0.4]]], np.float32))
    want = [0.026834250867889462, -0.0027901269813013, 0.10824690587224504, 0.009376228767352113, 0.1457999945654991,
            -0.21301506953160254, 0.9660636770271664, -0.045928938382823935, 0.289735452173094, -0.4407521661985643,
            0.3097696934057264, -0.05001598322836299, -0.04326093757468569, 0.22829154794195858, 1.6884957225745667,
            -2.585088485549571, 14540.771260427357, 14323.745029647385, 14685.45541428067, 14179.060875794072]
    np.testing.assert_allclose(sim.state20[0, 0], want, rtol=1e-12, atol=1e-15)
    np.testing.assert_allclose(sim.rpy_rates[0, 0], [0.24610303394348734, 1.7432907009137282, -2.5467995882925005], rtol=1e-12)
    assert abs(r[0] - 1.3664613008404878) < 1e-12
    assert sim.step_counter[0] == 80


@pytest.mark.parametrize("model", ["cf2x", "cf2p"])
def test_oracle_pid_teacher_forced(model):
    """DSLPIDControl.computeControl call log (DSLPIDControl.py:82-259): identical inputs, stateful."""
    g = load_golden(f"pid_calls_{model}.npz")
    pid = orc.make_pid(default_pid_params(DroneModel(model)))
    st = np.zeros(9)
    ins, outs, sa = g["inputs"], g["outputs"], g["state_after"]
    for t in range(ins.shape[0]):
        if t == int(g["reset_at"]):
            st[:] = 0
        x = ins[t]
        rpm, pos_e, yaw_e = orc.pid_compute(pid, float(g["dt"]), x[0:3], x[3:7], x[7:10], x[10:13], x[13:16],
                                            x[16:19], x[19:22], st)
        assert rel_err(rpm, outs[t, 0:4]) <= 1e-12, t
        assert np.max(np.abs(pos_e - outs[t, 4:7])) <= 1e-15, t
        assert abs(((yaw_e - outs[t, 7]) + np.pi) % (2 * np.pi) - np.pi) <= 1e-12, t
        np.testing.assert_allclose(st, sa[t], rtol=1e-12, atol=1e-13)


def test_oracle_pid_katb():
    g = load_golden("pid_katb.npz")
    for model in ("cf2x", "cf2p"):
        pid = orc.make_pid(default_pid_params(DroneModel(model)))
        st = np.zeros(9)
        for k in range(3):
            rpm, pos_e, yaw_e = orc.pid_compute(pid, 1 / 48, [.01, -.02, .98], g["quat"], [.03, .01, -.02], [.02, 0, 1],
                                                [0, 0, .3], [.01, 0, 0], None, st)
            np.testing.assert_allclose(np.hstack([rpm, pos_e, yaw_e]), g[model][k], rtol=1e-12, atol=1e-15)
        np.testing.assert_allclose(st, g[model + "_state"], rtol=1e-12, atol=1e-16)


@pytest.mark.parametrize("model", ["cf2x", "cf2p", "racer"])
def test_oracle_force_models(model):
    """_groundEffect/_drag/_downwash applyExternalForce arguments (BaseAviary.py:715-811)."""
    g = load_golden("forces.npz")
    d = orc.make_drone(load_drone_params(DroneModel(model)))
    inp = g[model + "_inputs"]
    for t in range(inp.shape[0]):
        pos_all = inp[t][:, 0:3]
        for i in range(inp.shape[1]):
            x = inp[t, i]
            pos, quat, rpy, vel, rpm = x[0:3], x[3:7], x[7:10], x[10:13], x[13:17]
            ge, ok = orc.ground_effect(d, rpm, pos, quat, rpy)
            assert rel_err(ge, g[model + "_gnd"][t, i], floor=1e-12) <= 1e-12
            assert ok == bool(g[model + "_gnd_applied"][t, i])
            dr = orc.drag(d, rpm, quat, vel)
            assert rel_err(dr, g[model + "_drag_body"][t, i], floor=1e-12) <= 1e-12
            dw = orc.downwash(d, pos_all, i)
            want = g[model + "_dw"][t, i]
            assert abs(dw - want) <= 1e-12 * max(abs(want), 1e-12)


def test_oracle_composite_katd():
    """SURVEY Appendix C KAT-D (build-defined DYN+GND+DRAG composite, reference force values)."""
    g = load_golden("composite_katd.npz")
    sim = orc.OracleSim(load_drone_params(DroneModel.CF2X), 1, num_drones=2, env_kind="multihover",
                        physics_flags=orc.PHY_GND | orc.PHY_DRAG)
    for _ in range(10):
        sim.step(np.array([[[0.1, -0.2, 0.3, -0.4], [0, 0, 0, 0]]], np.float32))
    st = np.concatenate([sim.state20[0], sim.rpy_rates[0]], axis=1)
    np.testing.assert_allclose(st, g["state"], rtol=1e-11, atol=1e-14)
    np.testing.assert_allclose(st[0, 0:3], [0.02726966, -0.00277237, 0.12380728], atol=1e-8)


def test_oracle_reset_quirks():
    """Ring survives reset() (BaseRLAviary.py:153-154); truncation clock fires before the counter advances
    (BaseAviary.py:379-382): first truncated step is 241 (30 Hz) / 385 (48 Hz) -> 242 / 386 steps per episode."""
    g = load_golden("reset_quirks.npz")
    sim = orc.OracleSim(load_drone_params(DroneModel.CF2X), 1)
    acts = g["actions"]
    for t in range(int(g["reset_after"])):
        sim.step(acts[t][None])
    rows = [sim.reset().copy()[0]]
    for t in range(int(g["reset_after"]), acts.shape[0]):
        rows.append(sim.step(acts[t][None])[0].copy()[0])
    np.testing.assert_allclose(np.array(rows, np.float64), g["obs_after_reset_then_steps"], rtol=1e-7, atol=1e-9)
    for freq in (30, 48):
        sim = orc.OracleSim(load_drone_params(DroneModel.CF2X), 1, ctrl_freq=freq)
        first = None
        for t in range(400):
            _, _, _, tr = sim.step(np.zeros((1, 1, 4), np.float32))
            if tr[0] and first is None:
                first = t
        assert first == int(g[f"first_truncated_step_{freq}"])
        assert first == {30: 241, 48: 385}[freq]


def test_oracle_threads_agree():
    rng = np.random.default_rng(0)
    a = rng.uniform(-1, 1, (5, 64, 1, 4)).astype(np.float32)
    s1 = orc.OracleSim(load_drone_params(DroneModel.CF2X), 64)
    s4 = orc.OracleSim(load_drone_params(DroneModel.CF2X), 64)
    for t in range(5):
        s1.step(a[t], nthreads=1)
        s4.step(a[t], nthreads=4)
    assert np.array_equal(s1.state20, s4.state20) and np.array_equal(s1.obs, s4.obs)


@pytest.mark.parametrize("model", ["cf2x", "cf2p"])
def test_oracle_pidpy_config0(model):
    """BASELINE configs[0]: the examples/pid.py loop on Physics.DYN (closed loop, first 40 ctrl steps)."""
    g = load_golden("pidpy_dyn.npz")
    xyz, rpy, wps = g[model + "_init_xyz"], g[model + "_init_rpy"], g[model + "_waypoints"]
    n = xyz.shape[0]
    dm = DroneModel(model)
    kw = dict(model=dm, env_kind="ctrl", action_type="ctrl_rpm", num_drones=n, pyb_freq=240, ctrl_freq=48,
              physics_flags=0, init_xyz=xyz, init_rpy=rpy)
    sim = make_oracle(kw)
    pid = orc.make_pid(default_pid_params(dm))
    pst = np.zeros((n, 9)); action = np.zeros((1, n, 4)); wp = g[model + "_wp0"].copy()
    ref_obs, ref_act = g[model + "_obs"], g[model + "_actions"]
    for t in range(ref_obs.shape[0]):
        obs, _, _, _ = sim.step(action)
        for j in range(n):
            s = obs[0, j]
            action[0, j], _, _ = orc.pid_compute(pid, 1 / 48, s[0:3], s[3:7], s[10:13], [wps[wp[j], 0], wps[wp[j], 1], xyz[j, 2]],
                                                 rpy[j], None, None, pst[j])
        wp = np.where(wp < wps.shape[0] - 1, wp + 1, 0)
        tol = 1e-9 if t < 20 else 1e-6          # chaotic closed loop: 1-ulp differences grow (SURVEY finding 6)
        assert rel_err(obs[0][:, 0:3], ref_obs[t][:, 0:3]) <= tol, t
        assert rel_err(action[0], ref_act[t]) <= tol * 10, t
