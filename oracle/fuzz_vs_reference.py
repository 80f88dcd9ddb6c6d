#!/usr/bin/env python
"""TEST INFRASTRUCTURE.  Differential fuzzing of the C oracle against the UNMODIFIED reference Python (run under the
stand-ins of oracle/refshim, like oracle/gen_golden.py): random (env class, drone model, drones, ctrl frequency, action
type, initial poses, DYN-form force models), open-loop action replay, every ctrl step compared.  Needs the reference tree (this container only):

    python oracle/fuzz_vs_reference.py --ref /root/reference --seeds 40        # prints one JSON line
"""
import argparse
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
import importlib.util  # noqa: E402

_spec = importlib.util.spec_from_file_location("gen_golden", os.path.join(HERE, "gen_golden.py"))
G = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(G)


def one_case(R, seed):
    import gpd_b200  # noqa: F401
    from gpd_b200.params import default_pid_params, load_drone_params
    from gpd_b200.utils.enums import DroneModel as DMh
    from oracle import oracle as orc
    rng = np.random.default_rng(7000 + seed)
    P, A, DM = R["Physics"], R["ActionType"], R["DroneModel"]
    kind = ["hover", "multihover", "ctrl"][rng.integers(3)]
    model = ["cf2x", "cf2p", "racer"][rng.integers(3)]
    n = 1 if kind == "hover" else int(rng.integers(1, 5))
    if kind == "multihover" and n == 1:
        n = 2
    freq = int([30, 48, 240, 60, 80][rng.integers(5)])
    if kind == "ctrl":
        act = "ctrl_rpm"
    else:
        act = ["rpm", "one_d_rpm", "pid", "one_d_pid"][rng.integers(4)]
    if act in ("pid", "one_d_pid") and model == "racer":
        model = "cf2x"
    closed_loop = act in ("pid", "one_d_pid")
    steps = 12 if closed_loop else int(rng.integers(20, 60))
    xyz = rng.uniform([-1, -1, 0.1], [1, 1, 1.5], size=(n, 3))
    rpy = rng.uniform(-0.3, 0.3, size=(n, 3))
    dm = DM(model)
    # every third case: the DYN-form composite (the reference's own _groundEffect/_drag/_downwash values injected)
    flags = 0
    if seed % 3 == 0:
        flags = int(rng.integers(1, 8)) if n > 1 else int(rng.integers(1, 4))
        xyz[:, 2] = 0.05 + 0.13 * rng.permutation(n) + rng.uniform(0, 0.02, n)       # near the ground, distinct heights
    C = {k: (G.make_composite_class(R, k, flags) if flags else R[k]) for k in ("CtrlAviary", "HoverAviary", "MultiHoverAviary")}
    R = dict(R, **C)
    with G.quiet():
        if kind == "ctrl":
            env = R["CtrlAviary"](drone_model=dm, num_drones=n, initial_xyzs=xyz, initial_rpys=rpy, physics=P.DYN, pyb_freq=240,
                                  ctrl_freq=freq)
        elif kind == "hover":
            env = R["HoverAviary"](drone_model=dm, initial_xyzs=xyz, initial_rpys=rpy, physics=P.DYN, pyb_freq=240, ctrl_freq=freq,
                                   act=A(act))
        else:
            env = R["MultiHoverAviary"](drone_model=dm, num_drones=n, initial_xyzs=xyz, initial_rpys=rpy, physics=P.DYN,
                                        pyb_freq=240, ctrl_freq=freq, act=A(act))
    dp = load_drone_params(DMh(model))
    target = None if kind == "ctrl" else np.asarray(env.TARGET_POS, np.float64).reshape(-1, 3)
    pid = default_pid_params(DMh.CF2X) if closed_loop else None          # in-env controllers are CF2X (BaseRLAviary.py:75)
    sim = orc.OracleSim(dp, 1, num_drones=n, env_kind=kind, action_type=act, pyb_freq=240, ctrl_freq=freq, physics_flags=flags,
                        pid_params=pid, init_xyz=xyz[None], init_rpy=rpy[None], target_pos=target)
    na, a = (n, 4) if kind == "ctrl" else env.action_space.shape
    if kind == "ctrl":
        acts = env.HOVER_RPM * (1 + 0.05 * rng.uniform(-1, 1, size=(steps, n, 4)))
    else:
        acts = (rng.uniform(-1, 1, size=(steps, na, a)) * (1.0 if rng.random() < .5 else 0.1)).astype(np.float32)
    with G.quiet():
        obs0, _ = env.reset()
    worst = float(np.max(np.abs(np.asarray(obs0, np.float64) - sim.obs[0])))
    flags_ok = True
    for t in range(steps):
        with G.quiet():
            o, r, te, tr, _ = env.step(acts[t])
        o2, r2, te2, tr2 = sim.step(acts[t][None])
        ref = G.full_state(env)
        mine = np.concatenate([sim.state20[0], sim.rpy_rates[0]], axis=-1)
        for sl in (slice(0, 3), slice(10, 13), slice(13, 16), slice(16, 20), slice(20, 23)):
            d = np.linalg.norm(ref[:, sl] - mine[:, sl], axis=-1) / np.maximum(np.linalg.norm(ref[:, sl], axis=-1), 1e-3)
            worst = max(worst, float(d.max()))
        q = np.minimum(np.linalg.norm(ref[:, 3:7] - mine[:, 3:7], axis=-1), np.linalg.norm(ref[:, 3:7] + mine[:, 3:7], axis=-1))
        worst = max(worst, float(q.max()), abs(float(r) - float(r2[0])) / max(abs(float(r)), 1e-3))
        worst = max(worst, float(np.max(np.abs(np.asarray(o, np.float64) - o2[0]) / np.maximum(np.abs(o2[0]), 1.0))) * 1e-3)
        flags_ok &= bool(te) == bool(te2[0]) and bool(tr) == bool(tr2[0]) and int(env.step_counter) == int(sim.step_counter[0])
    return dict(seed=seed, kind=kind, model=model, n=n, freq=freq, act=act, flags=flags, steps=steps, worst=worst, flags_ok=flags_ok)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default="/root/reference")
    ap.add_argument("--seeds", type=int, default=40)
    a = ap.parse_args()
    R = G._load_reference(a.ref)
    res = [one_case(R, s) for s in range(a.seeds)]
    bad = [r for r in res if not r["flags_ok"] or r["worst"] > (1e-6 if r["act"] in ("pid", "one_d_pid") else 1e-10)]
    print(json.dumps(dict(cases=len(res), worst=max(r["worst"] for r in res), failures=bad,
                          closed_loop_worst=max([r["worst"] for r in res if r["act"] in ("pid", "one_d_pid")] or [0.0]),
                          open_loop_worst=max([r["worst"] for r in res if r["act"] not in ("pid", "one_d_pid")] or [0.0]))))
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
