#!/usr/bin/env python
"""Generate tests/golden/*.npz|json by running the REFERENCE's own unmodified Python
(/root/reference/gym_pybullet_drones) under the stand-ins in oracle/refshim.

TEST INFRASTRUCTURE. Runs only in the build container (the GPU box has no /root/reference);
the fixtures it writes are committed, together with this script.

    python oracle/gen_golden.py [--out tests/golden] [--ref /root/reference]

What is recorded (all float64 unless noted; NumPy version stored in every file):
  constants.json        _parseURDFParameters + derived constants + DSLPIDControl gains, per model
  traj_roundtrip_*.npz  the same protocol with the stand-in's quaternion read-back round trip ON (what real Bullet does)
  traj_*.npz            action-replay trajectories (incl. VelocityAviary): float32 (RL) / float64 (Ctrl) action sequence,
                        state20 + rpy_rates + reward/terminated/truncated at checkpoints,
                        obs rows at a few steps
  pid_calls_*.npz       teacher-forced DSLPIDControl.computeControl call logs
  forces.npz            _groundEffect/_drag/_downwash applyExternalForce arguments on random states
  composite_*.npz       DYN-form composite (build-defined injection of the reference's own force
                        values; NOT a reference mode — SURVEY §8a)
  reset_quirks.npz      ring/controller survival across reset(), truncation clock
  logger.npz            utils/Logger.py arrays and CSV texts for a small random log
  pidpy_dyn.npz         BASELINE configs[0]: the examples/pid.py loop on Physics.DYN, first 40 ctrl steps
"""
import argparse
import contextlib
import io
import json
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def _load_reference(ref_root):
    warnings.filterwarnings("ignore")
    sys.path[:0] = [os.path.join(HERE, "refshim"), ref_root]
    with contextlib.redirect_stdout(io.StringIO()):
        import pybullet  # noqa: F401  (the stand-in)
        from gym_pybullet_drones.control.DSLPIDControl import DSLPIDControl
        from gym_pybullet_drones.envs.CtrlAviary import CtrlAviary
        from gym_pybullet_drones.envs.HoverAviary import HoverAviary
        from gym_pybullet_drones.envs.MultiHoverAviary import MultiHoverAviary
        from gym_pybullet_drones.envs.VelocityAviary import VelocityAviary
        from gym_pybullet_drones.utils.enums import ActionType, DroneModel, ObservationType, Physics
    return dict(DSLPIDControl=DSLPIDControl, CtrlAviary=CtrlAviary, HoverAviary=HoverAviary,
                MultiHoverAviary=MultiHoverAviary, VelocityAviary=VelocityAviary, ActionType=ActionType, DroneModel=DroneModel,
                ObservationType=ObservationType, Physics=Physics)


@contextlib.contextmanager
def quiet():
    with contextlib.redirect_stdout(io.StringIO()):
        yield


META = dict(numpy=np.__version__, standin="oracle/refshim/pybullet.py pass-through (ROUNDTRIP=False)")


def full_state(env):
    """(N,23): state20 + rpy_rates for every drone."""
    return np.array([np.hstack([env._getDroneStateVector(i), env.rpy_rates[i]]) for i in range(env.NUM_DRONES)])


def gen_constants(R, out):
    res = {"meta": META, "models": {}}
    for m in R["DroneModel"]:
        with quiet():
            env = R["CtrlAviary"](drone_model=m, physics=R["Physics"].DYN)
        d = {k: float(getattr(env, k)) for k in
             ["M", "L", "THRUST2WEIGHT_RATIO", "KF", "KM", "COLLISION_H", "COLLISION_R", "COLLISION_Z_OFFSET",
              "MAX_SPEED_KMH", "GND_EFF_COEFF", "PROP_RADIUS", "DW_COEFF_1", "DW_COEFF_2", "DW_COEFF_3",
              "G", "GRAVITY", "HOVER_RPM", "MAX_RPM", "MAX_THRUST", "MAX_XY_TORQUE", "MAX_Z_TORQUE", "GND_EFF_H_CLIP",
              "CTRL_TIMESTEP", "PYB_TIMESTEP"]}
        d["J"] = env.J.tolist()
        d["J_INV"] = env.J_INV.tolist()
        d["DRAG_COEFF"] = env.DRAG_COEFF.tolist()
        d["INIT_XYZS_3"] = None
        with quiet():
            env3 = R["CtrlAviary"](drone_model=m, num_drones=3, physics=R["Physics"].DYN)
        d["INIT_XYZS_3"] = np.asarray(env3.INIT_XYZS).tolist()
        if m.value in ("cf2x", "cf2p"):
            with quiet():
                c = R["DSLPIDControl"](drone_model=m)
            d["pid"] = {k: np.asarray(getattr(c, k)).tolist() for k in
                        ["P_COEFF_FOR", "I_COEFF_FOR", "D_COEFF_FOR", "P_COEFF_TOR", "I_COEFF_TOR", "D_COEFF_TOR",
                         "PWM2RPM_SCALE", "PWM2RPM_CONST", "MIN_PWM", "MAX_PWM", "MIXER_MATRIX", "GRAVITY", "KF"]}
        res["models"][m.value] = d
    with quiet():
        h = R["HoverAviary"](physics=R["Physics"].DYN)
        mh = R["MultiHoverAviary"](num_drones=2, physics=R["Physics"].DYN)
        hv = R["HoverAviary"](physics=R["Physics"].DYN, act=R["ActionType"].VEL)
    res["hover"] = dict(TARGET_POS=np.asarray(h.TARGET_POS).tolist(), EPISODE_LEN_SEC=h.EPISODE_LEN_SEC,
                        ACTION_BUFFER_SIZE=h.ACTION_BUFFER_SIZE, obs_shape=list(h.observation_space.shape),
                        act_shape=list(h.action_space.shape), INIT_XYZS=np.asarray(h.INIT_XYZS).tolist(),
                        SPEED_LIMIT=float(hv.SPEED_LIMIT))
    res["multihover2"] = dict(TARGET_POS=np.asarray(mh.TARGET_POS).tolist(), INIT_XYZS=np.asarray(mh.INIT_XYZS).tolist(),
                              obs_shape=list(mh.observation_space.shape))
    json.dump(res, open(os.path.join(out, "constants.json"), "w"), indent=1)


def action_stream(kind, rng, steps, n, a):
    if kind == "uniform":
        return rng.uniform(-1, 1, size=(steps, n, a)).astype(np.float32)
    if kind == "nearhover":
        return (0.05 * rng.standard_normal(size=(steps, n, a))).astype(np.float32)
    if kind == "symmetric":
        s = (0.3 * rng.standard_normal(size=(steps, n, 1))).astype(np.float32)
        return np.repeat(s, a, axis=2)
    raise ValueError(kind)


def replay(env, actions, ckpt_every, full_first, obs_steps):
    """Open-loop replay; returns dict of checkpoint arrays."""
    steps = actions.shape[0]
    idx, st, rew, term, trunc, cnt = [], [], [], [], [], []
    obs_rec = {}
    with quiet():
        obs0, _ = env.reset()
    for t in range(steps):
        with quiet():
            obs, r, te, tr, _ = env.step(actions[t])
        if t < full_first or (t + 1) % ckpt_every == 0 or t == steps - 1:
            idx.append(t)
            st.append(full_state(env))
            rew.append(float(r)); term.append(bool(te)); trunc.append(bool(tr)); cnt.append(int(env.step_counter))
        if t in obs_steps:
            obs_rec[t] = np.asarray(obs, dtype=np.float64)
    return dict(ckpt_idx=np.array(idx, np.int32), ckpt_state=np.array(st), ckpt_reward=np.array(rew),
                ckpt_terminated=np.array(term, np.uint8), ckpt_truncated=np.array(trunc, np.uint8),
                ckpt_counter=np.array(cnt, np.int32), obs0=np.asarray(obs0, np.float64),
                obs_steps=np.array(sorted(obs_rec), np.int32),
                obs_rows=np.array([obs_rec[k] for k in sorted(obs_rec)]))


def gen_traj(R, out):
    P, A, DM = R["Physics"], R["ActionType"], R["DroneModel"]
    cases = [
        # name, env ctor, kwargs, action kind, steps, A
        ("hover_cf2x_30_uniform", "HoverAviary", dict(drone_model=DM.CF2X, ctrl_freq=30), "uniform", 1000),
        ("hover_cf2x_30_nearhover", "HoverAviary", dict(drone_model=DM.CF2X, ctrl_freq=30), "nearhover", 1000),
        ("hover_cf2x_30_symmetric", "HoverAviary", dict(drone_model=DM.CF2X, ctrl_freq=30), "symmetric", 1000),
        ("hover_cf2x_48_uniform", "HoverAviary", dict(drone_model=DM.CF2X, ctrl_freq=48), "uniform", 1000),
        ("hover_cf2p_30_uniform", "HoverAviary", dict(drone_model=DM.CF2P, ctrl_freq=30), "uniform", 1000),
        ("hover_cf2p_48_nearhover", "HoverAviary", dict(drone_model=DM.CF2P, ctrl_freq=48), "nearhover", 1000),
        ("hover_racer_30_uniform", "HoverAviary", dict(drone_model=DM.RACE, ctrl_freq=30), "uniform", 1000),
        ("hover_cf2x_240_nearhover", "HoverAviary", dict(drone_model=DM.CF2X, ctrl_freq=240), "nearhover", 300),
        ("hover1d_cf2x_30_nearhover", "HoverAviary", dict(drone_model=DM.CF2X, ctrl_freq=30, act=A.ONE_D_RPM), "nearhover", 1000),
        ("multihover2_cf2x_30_uniform", "MultiHoverAviary", dict(drone_model=DM.CF2X, num_drones=2, ctrl_freq=30), "uniform", 1000),
        ("multihover3_cf2p_30_nearhover", "MultiHoverAviary", dict(drone_model=DM.CF2P, num_drones=3, ctrl_freq=30), "nearhover", 400),
        # closed-loop in-env controllers are chaotic (SURVEY finding 6): short horizons only
        ("hoverpid_cf2x_48", "HoverAviary", dict(drone_model=DM.CF2X, ctrl_freq=48, act=A.PID), "uniform", 24),
        ("hoverpid_cf2p_48", "HoverAviary", dict(drone_model=DM.CF2P, ctrl_freq=48, act=A.PID), "uniform", 24),
        ("hovervel_cf2p_48", "HoverAviary", dict(drone_model=DM.CF2P, ctrl_freq=48, act=A.VEL), "uniform", 24),
        ("hover1dpid_cf2p_48", "HoverAviary", dict(drone_model=DM.CF2P, ctrl_freq=48, act=A.ONE_D_PID), "uniform", 24),
    ]
    for ci, (name, ctor, kw, kind, steps) in enumerate(cases):
        rng = np.random.default_rng(1000 + ci)
        with quiet():
            env = R[ctor](physics=P.DYN, **kw)
        n, a = env.action_space.shape
        acts = action_stream(kind, rng, steps, n, a)
        short = steps <= 24
        rec = replay(env, acts, ckpt_every=1 if short else 10, full_first=steps if short else 20,
                     obs_steps={0, 1, 5, 16, 30, steps - 1} if not short else set(range(steps)))
        np.savez_compressed(os.path.join(out, f"traj_{name}.npz"), actions=acts, kind=kind, env=ctor,
                            model=kw["drone_model"].value, ctrl_freq=kw["ctrl_freq"], pyb_freq=240,
                            num_drones=n, act_type=kw.get("act", A.RPM).value, numpy=np.__version__, **rec)
    # --- CtrlAviary: float64 raw RPM actions, custom initial poses (pid.py / downwash.py shape) ---
    for ci, (name, model, n, freq, steps) in enumerate([("ctrl3_cf2x_48", DM.CF2X, 3, 48, 600),
                                                        ("ctrl2_racer_240", DM.RACE, 2, 240, 300)]):
        rng = np.random.default_rng(2000 + ci)
        xyz = rng.uniform([-1, -1, 0.2], [1, 1, 1.5], size=(n, 3))
        rpy = rng.uniform(-0.2, 0.2, size=(n, 3))
        with quiet():
            env = R["CtrlAviary"](drone_model=model, num_drones=n, initial_xyzs=xyz, initial_rpys=rpy, physics=P.DYN,
                                  pyb_freq=240, ctrl_freq=freq)
        acts = env.HOVER_RPM * (1 + 0.02 * rng.uniform(-1, 1, size=(steps, n, 4)))
        acts[5] = -100.0                 # exercises the [0, MAX_RPM] clip (CtrlAviary.py:140)
        acts[6] = 1e6
        rec = replay(env, acts, ckpt_every=10, full_first=20, obs_steps={0, 5, 6, steps - 1})
        np.savez_compressed(os.path.join(out, f"traj_{name}.npz"), actions=acts, kind="ctrl", env="CtrlAviary",
                            model=model.value, ctrl_freq=freq, pyb_freq=240, num_drones=n, act_type="ctrl_rpm",
                            init_xyz=xyz, init_rpy=rpy, numpy=np.__version__, **rec)


def gen_roundtrip(R, out):
    """The same replay protocol with the stand-in emulating what real Bullet does on every pose read-back: the base quaternion
    goes through a rotation-matrix round trip (unit norm, canonical sign).  The kernels never renormalise
    (BaseAviary.py:888 does not either): these files pin that the difference stays far below 1e-9 and only the quaternion's
    sign can differ (compared up to sign).  Real pybullet itself is not installable here: residual, stated in DESIGN.md."""
    import pybullet
    P, DM = R["Physics"], R["DroneModel"]
    pybullet.ROUNDTRIP = True
    try:
        for ci, (name, ctor, kw, kind, steps) in enumerate([
                ("roundtrip_hover_cf2x_30_uniform", "HoverAviary", dict(drone_model=DM.CF2X, ctrl_freq=30), "uniform", 1000),
                ("roundtrip_multihover2_cf2p_30_uniform", "MultiHoverAviary", dict(drone_model=DM.CF2P, num_drones=2, ctrl_freq=30),
                 "uniform", 600)]):
            rng = np.random.default_rng(3000 + ci)
            with quiet():
                env = R[ctor](physics=P.DYN, **kw)
            n, a = env.action_space.shape
            acts = action_stream(kind, rng, steps, n, a)
            rec = replay(env, acts, ckpt_every=10, full_first=20, obs_steps={0, 1, 5, 16, 30, steps - 1})
            np.savez_compressed(os.path.join(out, f"traj_{name}.npz"), actions=acts, kind=kind, env=ctor,
                                model=kw["drone_model"].value, ctrl_freq=kw["ctrl_freq"], pyb_freq=240, num_drones=n,
                                act_type="rpm", numpy=np.__version__, standin="ROUNDTRIP=True", **rec)
    finally:
        pybullet.ROUNDTRIP = False


def gen_velocity(R, out):
    """VelocityAviary (envs/VelocityAviary.py) on Physics.DYN: float64 velocity commands, closed loop -> short horizon."""
    P, DM = R["Physics"], R["DroneModel"]
    rng = np.random.default_rng(2100)
    n, steps = 2, 24
    xyz = rng.uniform([-1, -1, 0.3], [1, 1, 1.2], size=(n, 3))
    rpy = np.zeros((n, 3)); rpy[:, 2] = rng.uniform(-.5, .5, n)
    with quiet():
        env = R["VelocityAviary"](drone_model=DM.CF2P, num_drones=n, initial_xyzs=xyz, initial_rpys=rpy, physics=P.DYN,
                                  pyb_freq=240, ctrl_freq=48)
    acts = np.concatenate([rng.uniform(-1, 1, size=(steps, n, 3)), rng.uniform(0, 1, size=(steps, n, 1))], axis=2)
    acts[3, 0, :3] = 0.0            # zero direction -> zero unit vector branch (VelocityAviary.py:151-154)
    rec = replay(env, acts, ckpt_every=1, full_first=steps, obs_steps=set(range(steps)))
    np.savez_compressed(os.path.join(out, "traj_velocity2_cf2p_48.npz"), actions=acts, kind="ctrl_vel", env="VelocityAviary",
                        model="cf2p", ctrl_freq=48, pyb_freq=240, num_drones=n, act_type="ctrl_vel", init_xyz=xyz,
                        init_rpy=rpy, numpy=np.__version__, **rec)


def gen_pidpy(R, out):
    """BASELINE configs[0]: examples/pid.py wired exactly as the script (examples/pid.py:65-77,101-151) on Physics.DYN:
    CtrlAviary + one DSLPIDControl per drone tracking the circle trajectory, 240 Hz sim / 48 Hz ctrl.  The closed loop is
    chaotic (SURVEY finding 6; CF2X flips within ~14 ctrl steps, finding 5), so only the first 40 ctrl steps are kept."""
    P, DM = R["Physics"], R["DroneModel"]
    res = {}
    for model, n in ((DM.CF2X, 1), (DM.CF2P, 3)):
        ctrl_hz, H, H_STEP, RAD = 48, .1, .05, .3
        INIT_XYZS = np.array([[RAD * np.cos((i / 6) * 2 * np.pi + np.pi / 2), RAD * np.sin((i / 6) * 2 * np.pi + np.pi / 2) - RAD,
                               H + i * H_STEP] for i in range(n)])
        INIT_RPYS = np.array([[0, 0, i * (np.pi / 2) / n] for i in range(n)])
        NUM_WP = ctrl_hz * 10
        TARGET_POS = np.zeros((NUM_WP, 3))
        for i in range(NUM_WP):
            TARGET_POS[i, :] = (RAD * np.cos((i / NUM_WP) * (2 * np.pi) + np.pi / 2) + INIT_XYZS[0, 0],
                                RAD * np.sin((i / NUM_WP) * (2 * np.pi) + np.pi / 2) - RAD + INIT_XYZS[0, 1], 0)
        wp = np.array([int((i * NUM_WP / 6) % NUM_WP) for i in range(n)])
        with quiet():
            env = R["CtrlAviary"](drone_model=model, num_drones=n, initial_xyzs=INIT_XYZS, initial_rpys=INIT_RPYS,
                                  physics=P.DYN, neighbourhood_radius=10, pyb_freq=240, ctrl_freq=ctrl_hz, gui=False,
                                  record=False, obstacles=False, user_debug_gui=False)
            ctrl = [R["DSLPIDControl"](drone_model=model) for _ in range(n)]
        action = np.zeros((n, 4))
        T = 40
        obs_log, act_log = [], []
        for i in range(T):
            with quiet():
                obs, *_ = env.step(action)
                for j in range(n):
                    action[j, :], _, _ = ctrl[j].computeControlFromState(
                        control_timestep=env.CTRL_TIMESTEP, state=obs[j],
                        target_pos=np.hstack([TARGET_POS[wp[j], 0:2], INIT_XYZS[j, 2]]), target_rpy=INIT_RPYS[j, :])
            for j in range(n):
                wp[j] = wp[j] + 1 if wp[j] < (NUM_WP - 1) else 0
            obs_log.append(np.array(obs)); act_log.append(action.copy())
        k = model.value
        res[k + "_init_xyz"], res[k + "_init_rpy"], res[k + "_waypoints"] = INIT_XYZS, INIT_RPYS, TARGET_POS
        res[k + "_wp0"] = np.array([int((i * NUM_WP / 6) % NUM_WP) for i in range(n)], np.int32)
        res[k + "_obs"], res[k + "_actions"] = np.array(obs_log), np.array(act_log)
    np.savez_compressed(os.path.join(out, "pidpy_dyn.npz"), numpy=np.__version__, **res)


def gen_logger(R, out):
    """Reference utils/Logger.py on a small random log: arrays + the text of a few CSV files."""
    import glob
    import tempfile
    with quiet():
        from gym_pybullet_drones.utils.Logger import Logger
    rng = np.random.default_rng(7000)
    n, T, hz = 2, 6, 48
    states = rng.uniform(-1, 1, size=(T, n, 20))
    states[..., 16:20] = rng.uniform(9000, 20000, size=(T, n, 4))
    controls = rng.uniform(-1, 1, size=(T, n, 12))
    with tempfile.TemporaryDirectory() as tmp:
        lg = Logger(logging_freq_hz=hz, output_folder=tmp, num_drones=n)
        for t in range(T):
            for j in range(n):
                lg.log(drone=j, timestamp=t / hz, state=states[t, j], control=controls[t, j])
        lg.save()
        lg.save_as_csv("kat")
        npy = glob.glob(os.path.join(tmp, "save-flight-*.npy"))[0]
        z = np.load(npy)
        csv_dir = [d for d in glob.glob(os.path.join(tmp, "save-flight-kat-*")) if os.path.isdir(d)][0]
        names = sorted(os.listdir(csv_dir))
        texts = {k: open(os.path.join(csv_dir, k)).read() for k in ["z0.csv", "rr1.csv", "pwm2-0.csv", "wy1.csv", "ya0.csv"]}
        np.savez_compressed(os.path.join(out, "logger.npz"), in_states=states, in_controls=controls, hz=hz,
                            timestamps=z["timestamps"], states=z["states"], controls=z["controls"],
                            csv_names=np.array(names), csv_keys=np.array(list(texts)), csv_texts=np.array(list(texts.values())),
                            numpy=np.__version__)


def gen_pid(R, out):
    import pybullet as p
    for mi, m in enumerate([R["DroneModel"].CF2X, R["DroneModel"].CF2P]):
        rng = np.random.default_rng(3000 + mi)
        with quiet():
            c = R["DSLPIDControl"](drone_model=m)
        ncall = 400
        ins = np.zeros((ncall, 3 + 4 + 3 + 3 + 3 + 3 + 3))
        outs = np.zeros((ncall, 4 + 3 + 1))
        states = np.zeros((ncall, 9))
        dt = 1 / 48
        for t in range(ncall):
            if t == 200:            # second half: large attitudes / far targets (clips, near-gimbal targets)
                c.reset()
            big = t >= 200
            pos = rng.uniform(-1, 1, 3) * (3 if big else 1)
            rpy = rng.uniform(-1, 1, 3) * (1.4 if big else 0.3)
            quat = np.array(p.getQuaternionFromEuler(rpy))
            if big and t % 7 == 0:
                quat = -quat * 1.0001          # non-unit / negative-w quaternions
            vel = rng.uniform(-1, 1, 3) * (2 if big else 0.3)
            tpos = pos + rng.uniform(-1, 1, 3) * (2 if big else 0.2)
            trpy = np.array([0, 0, rng.uniform(-3, 3) if big else rng.uniform(-.5, .5)])
            tvel = rng.uniform(-.3, .3, 3) if t % 3 == 0 else np.zeros(3)
            trates = rng.uniform(-.3, .3, 3) if t % 5 == 0 else np.zeros(3)
            with quiet():
                rpm, pos_e, yaw_e = c.computeControl(dt, pos, quat, vel, np.zeros(3), tpos, trpy, tvel, trates)
            ins[t] = np.hstack([pos, quat, vel, tpos, trpy, tvel, trates])
            outs[t] = np.hstack([rpm, pos_e, yaw_e])
            states[t] = np.hstack([c.integral_pos_e, c.integral_rpy_e, c.last_rpy])
        np.savez_compressed(os.path.join(out, f"pid_calls_{m.value}.npz"), inputs=ins, outputs=outs, state_after=states,
                            dt=dt, reset_at=200, model=m.value, numpy=np.__version__)
    # KAT-B of SURVEY Appendix C (three identical calls)
    kat = {}
    for m in [R["DroneModel"].CF2X, R["DroneModel"].CF2P]:
        with quiet():
            c = R["DSLPIDControl"](drone_model=m)
        quat = np.array(p.getQuaternionFromEuler([.002, -.001, .3]))
        rows = []
        for _ in range(3):
            with quiet():
                rpm, pos_e, yaw_e = c.computeControl(1 / 48, np.array([.01, -.02, .98]), quat, np.array([.03, .01, -.02]),
                                                     np.zeros(3), np.array([.02, 0, 1]), np.array([0, 0, .3]),
                                                     np.array([.01, 0, 0]))
            rows.append(np.hstack([rpm, pos_e, yaw_e]))
        kat[m.value] = np.array(rows)
        kat[m.value + "_state"] = np.hstack([c.integral_pos_e, c.integral_rpy_e, c.last_rpy])
    np.savez_compressed(os.path.join(out, "pid_katb.npz"), quat=quat, numpy=np.__version__, **kat)


def gen_forces(R, out):
    import pybullet as p
    P, DM = R["Physics"], R["DroneModel"]
    rec = {}
    for mi, m in enumerate([DM.CF2X, DM.CF2P, DM.RACE]):
        rng = np.random.default_rng(4000 + mi)
        n, trials = 6, 40
        gnd, gnd_ok, drag, dw, inp = [], [], [], [], []
        for t in range(trials):
            xyz = rng.uniform([-1, -1, 0.01], [1, 1, 1.2], size=(n, 3))
            if t % 4 == 0:
                xyz[:, 2] = rng.uniform(0.0, 0.08, size=n)          # inside the height clip
            if t % 5 == 0:
                xyz[1, :2] = xyz[0, :2] + rng.uniform(-.05, .05, 2)  # stacked pair -> strong downwash
                xyz[1, 2] = xyz[0, 2] + rng.uniform(0.1, 1.0)
            rpy = rng.uniform(-1, 1, size=(n, 3)) * (0.3 if t % 3 else 1.7)
            with quiet():
                env = R["CtrlAviary"](drone_model=m, num_drones=n, initial_xyzs=xyz, initial_rpys=rpy,
                                      physics=P.PYB_GND_DRAG_DW, pyb_freq=240, ctrl_freq=48)
            rpm = env.HOVER_RPM * (1 + 0.3 * rng.uniform(-1, 1, size=(n, 4)))
            vel = rng.uniform(-2, 2, size=(n, 3))
            env.vel[:] = vel
            g_t, ok_t, d_t, w_t = [], [], [], []
            for i in range(n):
                p.pop_recorded(env.CLIENT)
                env._groundEffect(rpm[i], i)
                f, _ = p.pop_recorded(env.CLIENT)
                # recompute the values even when the attitude gate (BaseAviary.py:742) suppressed the call
                ls = p.getLinkStates(env.DRONE_IDS[i], linkIndices=[0, 1, 2, 3, 4], physicsClientId=env.CLIENT)
                h = np.clip(np.array([ls[k][0][2] for k in range(4)]), env.GND_EFF_H_CLIP, np.inf)
                vals = np.array(rpm[i] ** 2) * env.KF * env.GND_EFF_COEFF * (env.PROP_RADIUS / (4 * h)) ** 2
                if f:
                    assert [x[1] for x in f] == [0, 1, 2, 3]
                    assert np.array_equal(np.array([x[2][2] for x in f]), vals)
                g_t.append(vals); ok_t.append(len(f) == 4)
                env._drag(rpm[i], i)
                f, _ = p.pop_recorded(env.CLIENT)
                assert len(f) == 1 and f[0][1] == 4
                d_t.append(np.array(f[0][2]))
                env._downwash(i)
                f, _ = p.pop_recorded(env.CLIENT)
                w_t.append(sum(x[2][2] for x in f) if f else 0.0)
            inp.append(np.hstack([xyz, np.array([p.getQuaternionFromEuler(r) for r in rpy]), env.rpy, vel, rpm]))
            gnd.append(g_t); gnd_ok.append(ok_t); drag.append(d_t); dw.append(w_t)
        rec[m.value + "_inputs"] = np.array(inp)          # (trials, n, 3+4+3+3+4): pos, quat, rpy, vel, rpm
        rec[m.value + "_gnd"] = np.array(gnd)
        rec[m.value + "_gnd_applied"] = np.array(gnd_ok, np.uint8)
        rec[m.value + "_drag_body"] = np.array(drag)
        rec[m.value + "_dw"] = np.array(dw)
    np.savez_compressed(os.path.join(out, "forces.npz"), numpy=np.__version__, **rec)


def make_composite_class(R, base_name, flags):
    """Test-only subclass (reference files untouched): DYN substep with the reference's OWN
    _groundEffect/_drag/_downwash outputs injected (ground effect adds to the rotor forces
    before thrust/torques; drag and downwash are CoM LINK-frame forces -> R·f added to the world force)."""
    import pybullet as p
    Base = R[base_name]
    DM = R["DroneModel"]

    class Composite(Base):
        def _dynamics(self, rpm, nth_drone):
            gnd = np.zeros(4)
            body = np.zeros(3)
            p.pop_recorded(self.CLIENT)
            if flags & 1:
                self._groundEffect(rpm, nth_drone)
                f, _ = p.pop_recorded(self.CLIENT)
                for x in f:
                    gnd[x[1]] += x[2][2]
            if flags & 2:
                self._drag(self.last_clipped_action[nth_drone, :], nth_drone)
                f, _ = p.pop_recorded(self.CLIENT)
                for x in f:
                    body += np.array(x[2])
            if flags & 4:
                self._downwash(nth_drone)
                f, _ = p.pop_recorded(self.CLIENT)
                for x in f:
                    body += np.array(x[2])
            i = nth_drone
            Rm = np.array(p.getMatrixFromQuaternion(self.quat[i, :])).reshape(3, 3)
            f = np.array(rpm ** 2) * self.KF + gnd
            Fw = np.dot(Rm, np.array([0, 0, np.sum(f)])) - np.array([0, 0, self.GRAVITY])
            if flags & 6:
                Fw = Fw + np.dot(Rm, body)
            zt = np.array(rpm ** 2) * self.KM
            if self.DRONE_MODEL == DM.RACE:
                zt = -zt
            tz = (-zt[0] + zt[1] - zt[2] + zt[3])
            if self.DRONE_MODEL == DM.CF2P:
                tx = (f[1] - f[3]) * self.L
                ty = (-f[0] + f[2]) * self.L
            else:
                tx = (f[0] + f[1] - f[2] - f[3]) * (self.L / np.sqrt(2))
                ty = (- f[0] + f[1] + f[2] - f[3]) * (self.L / np.sqrt(2))
            w = self.rpy_rates[i, :]
            tau = np.array([tx, ty, tz]) - np.cross(w, np.dot(self.J, w))
            wdot = np.dot(self.J_INV, tau)
            acc = Fw / self.M
            vel = self.vel[i, :] + self.PYB_TIMESTEP * acc
            w = w + self.PYB_TIMESTEP * wdot
            pos = self.pos[i, :] + self.PYB_TIMESTEP * vel
            quat = self._integrateQ(self.quat[i, :], w, self.PYB_TIMESTEP)
            p.resetBasePositionAndOrientation(self.DRONE_IDS[i], pos, quat, physicsClientId=self.CLIENT)
            p.resetBaseVelocity(self.DRONE_IDS[i], vel, np.dot(Rm, w), physicsClientId=self.CLIENT)
            self.rpy_rates[i, :] = w

    return Composite


def gen_composite(R, out):
    P, DM = R["Physics"], R["DroneModel"]
    # C3 shape: MultiHoverAviary N=2, DYN+GND+DRAG, float32 U(-1,1) actions
    rng = np.random.default_rng(5000)
    Cls = make_composite_class(R, "MultiHoverAviary", 3)
    with quiet():
        env = Cls(drone_model=DM.CF2X, num_drones=2, physics=P.DYN, ctrl_freq=30)
    acts = action_stream("uniform", rng, 400, 2, 4)
    acts[:60] *= 0.2       # stay near the ground first so the ground effect matters
    rec = replay(env, acts, ckpt_every=10, full_first=20, obs_steps={0, 5, 399})
    np.savez_compressed(os.path.join(out, "composite_multihover2_gnd_drag.npz"), actions=acts, env="MultiHoverAviary",
                        model="cf2x", ctrl_freq=30, pyb_freq=240, num_drones=2, act_type="rpm", flags=3,
                        numpy=np.__version__, **rec)
    # KAT-D of SURVEY Appendix C
    with quiet():
        env = Cls(drone_model=DM.CF2X, num_drones=2, physics=P.DYN, ctrl_freq=30)
        env.reset()
        for _ in range(10):
            env.step(np.array([[0.1, -0.2, 0.3, -0.4], [0, 0, 0, 0]], dtype=np.float32))
    katd = full_state(env)
    # C4 shape (small): CtrlAviary N=8, DYN+DW and DYN+GND+DRAG+DW, float64 RPM around hover
    for name, flags in [("ctrl8_dw", 4), ("ctrl8_gnd_drag_dw", 7)]:
        rng = np.random.default_rng(5100 + flags)
        n = 8
        xyz = np.hstack([rng.uniform(-.3, .3, size=(n, 2)), rng.uniform(0.05, 1.5, size=(n, 1))])
        rpy = rng.uniform(-0.2, 0.2, size=(n, 3))
        Cls = make_composite_class(R, "CtrlAviary", flags)
        with quiet():
            env = Cls(drone_model=DM.CF2X, num_drones=n, initial_xyzs=xyz, initial_rpys=rpy, physics=P.DYN,
                      pyb_freq=240, ctrl_freq=48)
        acts = env.HOVER_RPM * (1 + 0.02 * rng.uniform(-1, 1, size=(200, n, 4)))
        rec = replay(env, acts, ckpt_every=10, full_first=20, obs_steps={0, 199})
        np.savez_compressed(os.path.join(out, f"composite_{name}.npz"), actions=acts, env="CtrlAviary", model="cf2x",
                            ctrl_freq=48, pyb_freq=240, num_drones=n, act_type="ctrl_rpm", flags=flags,
                            init_xyz=xyz, init_rpy=rpy, numpy=np.__version__, **rec)
    np.savez_compressed(os.path.join(out, "composite_katd.npz"), state=katd, numpy=np.__version__)


def gen_reset_quirks(R, out):
    P, DM = R["Physics"], R["DroneModel"]
    rng = np.random.default_rng(6000)
    with quiet():
        env = R["HoverAviary"](drone_model=DM.CF2X, physics=P.DYN, ctrl_freq=30)
        env.reset()
    acts = action_stream("uniform", rng, 12, 1, 4)
    rows = []
    with quiet():
        for t in range(7):
            obs, *_ = env.step(acts[t])
        obs_reset, _ = env.reset()                 # ring survives reset (BaseRLAviary.py:153-154)
        rows.append(np.asarray(obs_reset, np.float64))
        for t in range(7, 12):
            obs, *_ = env.step(acts[t])
            rows.append(np.asarray(obs, np.float64))
    # truncation clock: symmetric zero action -> pure vertical motion, only the time limit can fire
    flags = {}
    for freq in (30, 48):
        with quiet():
            env = R["HoverAviary"](drone_model=DM.CF2X, physics=P.DYN, ctrl_freq=freq)
            env.reset()
            tr_seq = []
            # hover exactly: action 0 -> rpm = HOVER_RPM; z stays ~0.1125
            for t in range(400):
                _, _, te, tr, _ = env.step(np.zeros((1, 4), np.float32))
                tr_seq.append(bool(tr))
        flags[f"first_truncated_step_{freq}"] = int(np.argmax(tr_seq))
    np.savez_compressed(os.path.join(out, "reset_quirks.npz"), actions=acts, obs_after_reset_then_steps=np.array(rows),
                        reset_after=7, numpy=np.__version__, **flags)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(os.path.dirname(HERE), "tests", "golden"))
    ap.add_argument("--ref", default="/root/reference")
    ap.add_argument("--only", default="", help="comma-separated subset of: constants,traj,roundtrip,velocity,pid,forces,composite,reset,logger,pidpy")
    a = ap.parse_args()
    os.makedirs(a.out, exist_ok=True)
    R = _load_reference(a.ref)
    gens = dict(constants=gen_constants, traj=gen_traj, roundtrip=gen_roundtrip, velocity=gen_velocity, pid=gen_pid, forces=gen_forces,
                composite=gen_composite, reset=gen_reset_quirks, logger=gen_logger, pidpy=gen_pidpy)
    for name, fn in gens.items():
        if not a.only or name in a.only.split(","):
            fn(R, a.out)
    total = sum(os.path.getsize(os.path.join(a.out, f)) for f in os.listdir(a.out))
    print(f"wrote {len(os.listdir(a.out))} files, {total / 1e6:.2f} MB to {a.out}")


if __name__ == "__main__":
    main()
