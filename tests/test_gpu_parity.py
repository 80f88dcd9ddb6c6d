"""-m gpu: the CUDA path (through the C ABI, include/gpd.h) against the golden vectors generated from the
reference and against the CPU oracle on seeded inputs.  FP64: <= 1e-9 relative over up to 1,000 ctrl steps
(action replay); FP32: <= 1e-4 relative on position/velocity over 100 ctrl steps (BASELINE.json north_star)."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")

from helpers import (S_ANGV, S_POS, S_QUAT, S_RATES, S_RPM, S_RPY, S_VEL, angle_err, case_setup, default_targets,
                     load_golden, make_oracle, quat_err, rel_err, traj_cases)
from gpd_b200.params import default_pid_params, load_drone_params
from gpd_b200.utils.enums import DroneModel, Physics

pytestmark = pytest.mark.gpu

TOL64 = {}     # every trajectory at 1e-9 (the float32 BLAS sdot of the reference's VEL mapping is reproduced bit for bit)


def make_sim(kw, num_envs=1, precision="f64", auto_reset=False, tpb=0):
    from gpd_b200.sim import BatchedSim
    dp = load_drone_params(kw["model"])
    n = kw["num_drones"]
    target = default_targets(kw)
    return BatchedSim(dp, num_envs, n, env_kind=kw["env_kind"], action_type=kw["action_type"], pyb_freq=kw["pyb_freq"],
                      ctrl_freq=kw["ctrl_freq"], physics_flags=kw["physics_flags"], precision=precision,
                      auto_reset=auto_reset, target_pos=target, init_xyz=kw["init_xyz"], init_rpy=kw["init_rpy"],
                      threads_per_block=tpb)


def state_np(sim):
    st, rr, ps, cnt = sim.get_state()
    return (np.concatenate([st.double().cpu().numpy(), rr.double().cpu().numpy()], axis=-1), ps.double().cpu().numpy(),
            cnt.cpu().numpy())


@pytest.mark.parametrize("name", traj_cases())
def test_cuda_f64_replays_reference_trajectory(name):
    g = load_golden(name)
    kw = case_setup(g)
    E = 3                                             # same actions in 3 envs: exercises the batch axis + block tail
    sim = make_sim(kw, num_envs=E)
    tol = TOL64.get(name, 1e-9)
    acts = g["actions"]
    ck = {int(t): i for i, t in enumerate(g["ckpt_idx"])}
    osteps = {int(t): i for i, t in enumerate(g["obs_steps"])}
    obs = sim.reset()
    assert np.max(np.abs(obs.double().cpu().numpy()[1] - g["obs0"])) <= 1e-7 * (tol / 1e-9)
    adt = torch.float64 if kw["env_kind"] == "ctrl" else torch.float32
    if kw["action_type"] == "ctrl_vel":
        from gpd_b200.envs import VelocityAviary          # façade smoke: same library path, reference class name
        ve = VelocityAviary(drone_model=kw["model"], num_drones=kw["num_drones"], initial_xyzs=kw["init_xyz"],
                            initial_rpys=kw["init_rpy"], pyb_freq=240, ctrl_freq=kw["ctrl_freq"], num_envs=2, precision="f64")
        assert ve.action_space.shape == (kw["num_drones"], 4) and abs(ve.SPEED_LIMIT - 0.25) < 1e-12
        o, r, te, tr, _ = ve.step(torch.as_tensor(acts[0][None].repeat(2, 0)).cuda())
        assert o.shape == (2, kw["num_drones"], 20) and float(r[0]) == -1.0 and not bool(te[0])
        ve.close()
    dev_acts = torch.as_tensor(np.repeat(acts[:, None], E, axis=1), dtype=adt).cuda()
    for t in range(acts.shape[0]):
        obs, rew, term, trunc = sim.step(dev_acts[t])
        if t in ck:
            i = ck[t]
            ref = g["ckpt_state"][i]
            st_all, _, cnt = state_np(sim)
            for e in (0, E - 1):
                st = st_all[e]
                assert rel_err(st[:, S_POS], ref[:, S_POS]) <= tol, (t, "pos")
                assert rel_err(st[:, S_VEL], ref[:, S_VEL]) <= tol, (t, "vel")
                assert rel_err(st[:, S_RATES], ref[:, S_RATES]) <= tol, (t, "rpy_rates")
                assert rel_err(st[:, S_ANGV], ref[:, S_ANGV]) <= tol, (t, "ang_v")
                assert quat_err(st[:, S_QUAT], ref[:, S_QUAT]) <= tol, (t, "quat")
                assert angle_err(st[:, S_RPY], ref[:, S_RPY]) <= tol * 10, (t, "rpy")
                assert rel_err(st[:, S_RPM], ref[:, S_RPM]) <= tol, (t, "rpm")
            r = rew.cpu().numpy()
            assert np.all(np.abs(r - g["ckpt_reward"][i]) <= tol * max(abs(g["ckpt_reward"][i]), 1e-3) * 10), (t, "reward")
            assert np.all(term.cpu().numpy().astype(bool) == bool(g["ckpt_terminated"][i])), (t, "terminated")
            assert np.all(trunc.cpu().numpy().astype(bool) == bool(g["ckpt_truncated"][i])), (t, "truncated")
            assert np.all(cnt == g["ckpt_counter"][i])
        if t in osteps:
            want = g["obs_rows"][osteps[t]]
            got = obs.double().cpu().numpy()
            scale = np.maximum(np.abs(want), 1.0)
            for e in (0, E - 1):
                assert np.max(np.abs(got[e] - want) / scale) <= max(1e-9, tol * 100), (t, "obs")
    sim.close()


def _random_init(rng, E, N):
    xyz = rng.uniform([-1, -1, 0.2], [1, 1, 1.5], size=(E, N, 3))
    rpy = rng.uniform(-0.2, 0.2, size=(E, N, 3))
    return xyz, rpy


@pytest.mark.parametrize("model,env_kind,N,freq,flags,steps", [
    (DroneModel.CF2X, "hover", 1, 30, 0, 1000),
    (DroneModel.CF2P, "hover", 1, 48, 0, 300),
    (DroneModel.CF2X, "multihover", 2, 30, 3, 300),      # BASELINE config 3 shape: DYN+GND+DRAG, FP64
    (DroneModel.RACE, "multihover", 3, 30, 0, 200),
])
def test_cuda_f64_vs_oracle_random_inits(model, env_kind, N, freq, flags, steps):
    """1,000-step action replay from randomised per-env initial poses, E=257 (not a multiple of the block)."""
    rng = np.random.default_rng(7)
    E = 257
    xyz, rpy = _random_init(rng, E, N)
    if N > 1:
        xyz[..., 2] = rng.uniform(0.03, 0.4, size=(E, N))           # near the ground: ground effect active
    kw = dict(model=model, env_kind=env_kind, action_type="rpm", num_drones=N, pyb_freq=240, ctrl_freq=freq,
              physics_flags=flags, init_xyz=xyz, init_rpy=rpy)
    ref = make_oracle(kw, num_envs=E)
    sim = make_sim(kw, num_envs=E)
    sim.reset()
    scale = 0.05 if env_kind == "hover" else 0.3
    for t in range(steps):
        a = (scale * rng.standard_normal(size=(E, N, 4))).astype(np.float32)
        obs, rew, term, trunc = sim.step(torch.from_numpy(a).cuda())
        o_ref, r_ref, te_ref, tr_ref = ref.step(a, nthreads=4)
        if t % 50 == 49 or t == steps - 1:
            st, _, cnt = state_np(sim)
            rs = np.concatenate([ref.state20, ref.rpy_rates], axis=-1)
            for sl, nm in ((S_POS, "pos"), (S_VEL, "vel"), (S_RATES, "rates"), (S_ANGV, "ang_v")):
                assert rel_err(st[..., sl], rs[..., sl]) <= 1e-9, (t, nm)
            assert quat_err(st[..., S_QUAT], rs[..., S_QUAT]) <= 1e-9
            assert np.array_equal(cnt, ref.step_counter)
            assert np.max(np.abs(rew.cpu().numpy() - r_ref) / np.maximum(np.abs(r_ref), 1e-3)) <= 1e-8
            assert np.array_equal(term.cpu().numpy(), te_ref) and np.array_equal(trunc.cpu().numpy(), tr_ref)
            assert np.max(np.abs(obs.cpu().numpy().astype(np.float64) - o_ref) / np.maximum(np.abs(o_ref), 1.0)) <= 1e-6
    sim.close()


@pytest.mark.parametrize("model,env_kind,N,freq,scale", [
    (DroneModel.CF2X, "hover", 1, 30, 0.05), (DroneModel.CF2X, "hover", 1, 30, 1.0),
    (DroneModel.CF2P, "hover", 1, 48, 0.05), (DroneModel.CF2X, "multihover", 2, 30, 1.0),
    (DroneModel.RACE, "hover", 1, 30, 0.3),
])
def test_cuda_f32_within_1e4_over_100_steps(model, env_kind, N, freq, scale):
    """FP32 throughput mode: <= 1e-4 relative on position and velocity over 100 ctrl steps vs the FP64 oracle.
    Error metric: ||dx||_2 / max(||x_ref||_2, 1e-3) per drone."""
    rng = np.random.default_rng(11)
    E = 512
    kw = dict(model=model, env_kind=env_kind, action_type="rpm", num_drones=N, pyb_freq=240, ctrl_freq=freq,
              physics_flags=0, init_xyz=None, init_rpy=None)
    ref = make_oracle(kw, num_envs=E)
    sim = make_sim(kw, num_envs=E, precision="f32")
    sim.reset()
    worst = {"pos": 0.0, "vel": 0.0}
    for t in range(100):
        a = rng.uniform(-scale, scale, size=(E, N, 4)).astype(np.float32)
        sim.step(torch.from_numpy(a).cuda())
        ref.step(a, nthreads=4)
        if t % 10 == 9:
            st, _, _ = state_np(sim)
            worst["pos"] = max(worst["pos"], rel_err(st[..., S_POS], ref.state20[..., S_POS]))
            worst["vel"] = max(worst["vel"], rel_err(st[..., S_VEL], ref.state20[..., S_VEL]))
    assert worst["pos"] <= 1e-4 and worst["vel"] <= 1e-4, worst
    sim.close()


@pytest.mark.parametrize("model", ["cf2x", "cf2p"])
@pytest.mark.parametrize("precision,tol", [("f64", 1e-12), ("f32", 2e-4)])
def test_cuda_pid_teacher_forced(model, precision, tol):
    """Batched DSLPIDControl.computeControl against the reference's call log, identical inputs each call."""
    from gpd_b200.control.DSLPIDControl import DSLPIDControl
    g = load_golden(f"pid_calls_{model}.npz")
    ins, outs, sa = g["inputs"], g["outputs"], g["state_after"]
    c = DSLPIDControl(DroneModel(model), num=2, precision=precision)
    rdt = torch.float64 if precision == "f64" else torch.float32
    for t in range(ins.shape[0]):
        if t == int(g["reset_at"]):
            c.reset()
        # teacher forcing: controller state from the reference before the call
        prev = np.zeros(9) if t in (0, int(g["reset_at"])) else sa[t - 1]
        c.state[:] = torch.as_tensor(prev, dtype=rdt).cuda()
        x = ins[t]
        rpm, pos_e, yaw_e = c.computeControl(float(g["dt"]), x[0:3], x[3:7], x[7:10], np.zeros(3), x[10:13], x[13:16],
                                             x[16:19], x[19:22])
        rpm, pos_e, yaw_e = rpm.double().cpu().numpy(), pos_e.double().cpu().numpy(), yaw_e.double().cpu().numpy()
        for row in (0, 1):
            assert rel_err(rpm[row], outs[t, 0:4]) <= tol, (t, rpm[row], outs[t, 0:4])
            assert np.max(np.abs(pos_e[row] - outs[t, 4:7])) <= (1e-15 if precision == "f64" else 1e-6)
            assert abs(((yaw_e[row] - outs[t, 7]) + np.pi) % (2 * np.pi) - np.pi) <= (1e-12 if precision == "f64" else 1e-5)
        st = c.state.double().cpu().numpy()[0]
        np.testing.assert_allclose(st, sa[t], rtol=1e-12 if precision == "f64" else 1e-4,
                                   atol=1e-13 if precision == "f64" else 1e-5)


@pytest.mark.parametrize("model", ["cf2x", "cf2p", "racer"])
def test_cuda_force_models(model):
    """gpd_force_* against the recorded applyExternalForce arguments of the reference (<= 1e-12 relative)."""
    import ctypes as C
    from gpd_b200 import _lib
    L = _lib.load()
    g = load_golden("forces.npz")
    d = _lib.drone_params_c(load_drone_params(DroneModel(model)))
    inp = g[model + "_inputs"]
    T, N = inp.shape[0], inp.shape[1]
    flat = torch.as_tensor(inp.reshape(T * N, 17)).cuda()
    pos, quat, vel, rpm = (flat[:, 0:3].contiguous(), flat[:, 3:7].contiguous(), flat[:, 10:13].contiguous(),
                           flat[:, 13:17].contiguous())
    p = lambda t: C.c_void_p(t.data_ptr())
    ge = torch.empty((T * N, 4), dtype=torch.float64, device="cuda")
    ok = torch.empty((T * N,), dtype=torch.uint8, device="cuda")
    _lib.check(L.gpd_force_ground_effect(0, 1, C.byref(d), T * N, p(rpm), p(pos), p(quat), p(ge), p(ok), None))
    dr = torch.empty((T * N, 3), dtype=torch.float64, device="cuda")
    _lib.check(L.gpd_force_drag(0, 1, C.byref(d), T * N, p(rpm), p(quat), p(vel), p(dr), None))
    dw = torch.empty((T, N), dtype=torch.float64, device="cuda")
    posT = pos.reshape(T, N, 3).contiguous()
    _lib.check(L.gpd_force_downwash(0, 1, C.byref(d), T, N, p(posT), p(dw), None))
    torch.cuda.synchronize()
    assert rel_err(ge.cpu().numpy(), g[model + "_gnd"].reshape(T * N, 4), floor=1e-12) <= 1e-12
    assert np.array_equal(ok.cpu().numpy(), g[model + "_gnd_applied"].reshape(-1))
    assert rel_err(dr.cpu().numpy(), g[model + "_drag_body"].reshape(T * N, 3), floor=1e-12) <= 1e-12
    want = g[model + "_dw"]
    assert np.max(np.abs(dw.cpu().numpy() - want) / np.maximum(np.abs(want), 1e-12)) <= 1e-11


def test_cuda_reset_quirks_and_env_api():
    """Ring survives reset (BaseRLAviary.py:153-154); truncation clock (BaseAviary.py:379-382); env façade API."""
    from gpd_b200.envs import HoverAviary
    g = load_golden("reset_quirks.npz")
    env = HoverAviary(num_envs=2, precision="f64")
    assert env.observation_space.shape == (1, 72) and env.action_space.shape == (1, 4)
    assert env.ACTION_BUFFER_SIZE == 15 and env.PYB_STEPS_PER_CTRL == 8 and env.EPISODE_LEN_SEC == 8
    acts = torch.as_tensor(g["actions"]).float().cuda()
    env.reset()
    for t in range(int(g["reset_after"])):
        env.step(acts[t].expand(2, 1, 4).contiguous())
    rows = [env.reset()[0][1].double().cpu().numpy()]
    for t in range(int(g["reset_after"]), acts.shape[0]):
        rows.append(env.step(acts[t].expand(2, 1, 4).contiguous())[0][1].double().cpu().numpy())
    np.testing.assert_allclose(np.array(rows), g["obs_after_reset_then_steps"], rtol=1e-7, atol=1e-9)
    env.close()
    for freq in (30, 48):
        env = HoverAviary(num_envs=1, ctrl_freq=freq, precision="f64")
        env.reset()
        first = None
        z = torch.zeros((1, 1, 4), dtype=torch.float32, device="cuda")
        for t in range(400):
            _, _, te, tr, info = env.step(z)
            if bool(tr[0]) and first is None:
                first = t
        assert first == int(g[f"first_truncated_step_{freq}"]) == {30: 241, 48: 385}[freq]
        assert info == {"answer": 42}
        env.close()
    with pytest.raises(ValueError):
        HoverAviary(pyb_freq=240, ctrl_freq=7)


def test_cuda_auto_reset_terminal_obs_and_stats():
    """SB3-style auto-reset inside the kernel: done envs restart from the initial pose, the terminal kinematic
    observation is kept, Monitor-style episode statistics agree with a host-side recount (oracle-driven)."""
    rng = np.random.default_rng(3)
    E = 300
    kw = dict(model=DroneModel.CF2X, env_kind="hover", action_type="rpm", num_drones=1, pyb_freq=240, ctrl_freq=30,
              physics_flags=0, init_xyz=None, init_rpy=None)
    sim = make_sim(kw, num_envs=E, precision="f64", auto_reset=True)
    ref = make_oracle(kw, num_envs=E)
    sim.reset()
    ep_ret = np.zeros(E); ep_len = np.zeros(E, int)
    n_ep = 0; sum_ret = 0.0; sum_len = 0; n_term = 0; rmin, rmax = np.inf, -np.inf
    for t in range(60):
        a = rng.uniform(-1, 1, size=(E, 1, 4)).astype(np.float32)
        obs, rew, term, trunc = sim.step(torch.from_numpy(a).cuda())
        o_ref, r_ref, te_ref, tr_ref = ref.step(a)
        o_ref = o_ref.copy()
        done = (te_ref | tr_ref).astype(bool)
        assert np.array_equal((term | trunc).cpu().numpy().astype(bool), done)
        tk = sim.terminal_kin.cpu().numpy()
        assert np.max(np.abs(tk[done][:, 0] - o_ref[done][:, 0, :12])) <= 1e-6 if done.any() else True
        ep_ret += r_ref; ep_len += 1
        for e in np.nonzero(done)[0]:
            n_ep += 1; sum_ret += ep_ret[e]; sum_len += ep_len[e]; n_term += int(te_ref[e])
            rmin, rmax = min(rmin, ep_ret[e]), max(rmax, ep_ret[e])
            ep_ret[e] = 0; ep_len[e] = 0
        if done.any():
            o_reset = ref.reset(done.astype(np.uint8))
            o_ref[done] = o_reset[done]
        assert np.max(np.abs(obs.cpu().numpy() - o_ref)) <= 1e-6
    stats = sim.episode_stats()
    assert stats[0] == n_ep and n_ep > 0
    assert abs(stats[1] - sum_ret) <= 1e-4 * max(1.0, abs(sum_ret))
    assert stats[2] == sum_len and stats[6] == 60 * E and stats[7] == n_term
    assert abs(stats[4] - rmin) <= 1e-5 * max(1, abs(rmin)) and abs(stats[5] - rmax) <= 1e-5 * max(1, abs(rmax))
    sim.close()


def test_cuda_get_set_state_roundtrip_and_host_path():
    """gpd_get_state/gpd_set_state checkpoint a run bit-exactly; gpd_step_host equals gpd_step."""
    rng = np.random.default_rng(5)
    E = 130
    kw = dict(model=DroneModel.CF2P, env_kind="hover", action_type="pid", num_drones=1, pyb_freq=240, ctrl_freq=48,
              physics_flags=0, init_xyz=None, init_rpy=None)
    a = rng.uniform(-1, 1, size=(12, E, 1, 3)).astype(np.float32)
    s1 = make_sim(kw, num_envs=E)
    s1.reset()
    for t in range(6):
        s1.step(torch.from_numpy(a[t]).cuda())
    st, rr, ps, cnt = s1.get_state()
    s2 = make_sim(kw, num_envs=E)
    s2.set_state(st, rr, ps, cnt)
    s2.reset(torch.zeros(E, dtype=torch.uint8))            # no env reset: only primes obs with a zero ring
    for t in range(6, 12):
        o1, r1, _, _ = s1.step(torch.from_numpy(a[t]).cuda())
        o2, r2, _, _ = s2.step(torch.from_numpy(a[t]).cuda())
    x1, x2 = s1.get_state(), s2.get_state()
    for u, v in zip(x1, x2):
        assert torch.equal(u, v)
    assert torch.equal(o1[..., :12], o2[..., :12]) and torch.equal(r1, r2)
    # host path
    s3 = make_sim(kw, num_envs=E)
    s4 = make_sim(kw, num_envs=E)
    s3.reset(); s4.reset_host()
    for t in range(5):
        od, rd, td, trd = s3.step(torch.from_numpy(a[t]).cuda())
        oh, rh, th, trh, _ = s4.step_host(a[t])
    assert np.array_equal(od.cpu().numpy(), oh) and np.array_equal(rd.cpu().numpy(), rh)
    assert np.array_equal(td.cpu().numpy(), th) and np.array_equal(trd.cpu().numpy(), trh)
    for s in (s1, s2, s3, s4):
        s.close()


def test_cuda_rollout_pid_matches_stepwise_loop():
    """gpd_rollout_pid (examples/pid.py loop in one launch) == gpd_step + gpd_pid_compute called step by step,
    and both track the oracle over a short horizon (closed loop is chaotic beyond ~20 steps, SURVEY finding 6)."""
    from gpd_b200.control.DSLPIDControl import DSLPIDControl
    from oracle import oracle as orc
    E, N, steps = 70, 2, 16
    model = DroneModel.CF2P
    dp = load_drone_params(model)
    rng = np.random.default_rng(9)
    xyz = rng.uniform([-.3, -.3, 0.1], [.3, .3, 0.6], size=(E, N, 3))
    rpy = np.zeros((E, N, 3)); rpy[..., 2] = rng.uniform(-.5, .5, size=(E, N))
    kw = dict(model=model, env_kind="ctrl", action_type="ctrl_rpm", num_drones=N, pyb_freq=240, ctrl_freq=48,
              physics_flags=0, init_xyz=xyz, init_rpy=rpy)
    n_wp = 48
    wps = np.stack([.3 * np.cos(np.arange(n_wp) / n_wp * 2 * np.pi), .3 * np.sin(np.arange(n_wp) / n_wp * 2 * np.pi),
                    np.zeros(n_wp)], axis=1)
    from gpd_b200.sim import BatchedSim
    def mk():
        return BatchedSim(dp, E, N, env_kind="ctrl", action_type="ctrl_rpm", pyb_freq=240, ctrl_freq=48,
                          precision="f64", pid=default_pid_params(model), init_xyz=xyz, init_rpy=rpy)
    # (a) one launch
    sa = mk()
    wp_a = torch.as_tensor(rng.integers(0, n_wp, size=(E, N)), dtype=torch.int32).cuda()
    wp0 = wp_a.clone()
    act_a = torch.zeros((E, N, 4), dtype=torch.float64, device="cuda")
    sa.rollout_pid(steps, torch.as_tensor(wps).cuda(), wp_a, act_a)
    # (b) stepwise through the public pieces
    sb = mk()
    ctrl = DSLPIDControl(model, num=E * N, precision="f64")
    act_b = torch.zeros((E, N, 4), dtype=torch.float64, device="cuda")
    wp_b = wp0.clone().long()
    wps_t = torch.as_tensor(wps).cuda()
    init_z = torch.as_tensor(xyz[..., 2]).cuda()
    trpy = torch.as_tensor(rpy).cuda().reshape(-1, 3)
    # (c) oracle
    ref = make_oracle(kw, num_envs=E)
    pid_o = orc.make_pid(default_pid_params(model))
    pst = np.zeros((E, N, 9)); act_o = np.zeros((E, N, 4)); wp_o = wp0.cpu().numpy().copy()
    for t in range(steps):
        obs, _, _, _ = sb.step(act_b)
        tp = torch.cat([wps_t[wp_b.reshape(-1)][:, 0:2], init_z.reshape(-1, 1)], dim=1)
        rpm, _, _ = ctrl.computeControlFromState(1 / 48, obs.reshape(-1, 20), tp, trpy)
        act_b = rpm.reshape(E, N, 4).clone()
        wp_b = torch.where(wp_b < n_wp - 1, wp_b + 1, torch.zeros_like(wp_b))
        ref.step(act_o)
        for e in range(E):
            for i in range(N):
                s = ref.state20[e, i]
                r, _, _ = orc.pid_compute(pid_o, 1 / 48, s[0:3], s[3:7], s[10:13],
                                          [wps[wp_o[e, i], 0], wps[wp_o[e, i], 1], xyz[e, i, 2]], rpy[e, i], None, None, pst[e, i])
                act_o[e, i] = r
        wp_o = np.where(wp_o < n_wp - 1, wp_o + 1, 0)
    xa, xb = sa.get_state(), sb.get_state()
    assert torch.equal(wp_a.long(), wp_b)
    assert rel_err(xa[0][..., 0:3].cpu().numpy(), xb[0][..., 0:3].cpu().numpy()) <= 1e-12
    assert rel_err(act_a.cpu().numpy(), act_b.cpu().numpy()) <= 1e-10
    assert rel_err(xa[0][..., 0:3].cpu().numpy(), ref.state20[..., 0:3]) <= 1e-7
    assert rel_err(act_a.cpu().numpy(), act_o) <= 1e-6
    sa.close(); sb.close()


def test_cuda_vec_env_protocol():
    """SB3 VecEnv contract over the batched env: numpy in/out, auto-reset, terminal_observation, TimeLimit.truncated,
    Monitor-style episode infos — checked against the oracle stepped with the same actions."""
    from gpd_b200.envs import HoverAviary
    from gpd_b200.vec_env import GpdVecEnv
    rng = np.random.default_rng(21)
    E = 96
    venv = GpdVecEnv(HoverAviary, E, precision="f64")
    assert venv.num_envs == E and venv.observation_space.shape == (1, 72) and venv.action_space.shape == (1, 4)
    kw = dict(model=DroneModel.CF2X, env_kind="hover", action_type="rpm", num_drones=1, pyb_freq=240, ctrl_freq=30,
              physics_flags=0, init_xyz=None, init_rpy=None)
    ref = make_oracle(kw, num_envs=E)
    obs = venv.reset()
    assert isinstance(obs, np.ndarray) and obs.shape == (E, 1, 72) and obs.dtype == np.float32
    ep_r = np.zeros(E); ep_l = np.zeros(E, int)
    seen_done = 0
    for t in range(40):
        a = rng.uniform(-1, 1, size=(E, 1, 4)).astype(np.float32)
        obs, rew, dones, infos = venv.step(a)
        o_ref, r_ref, te_ref, tr_ref = ref.step(a)
        o_ref = o_ref.copy()
        d_ref = (te_ref | tr_ref).astype(bool)
        assert np.array_equal(dones, d_ref) and len(infos) == E
        np.testing.assert_allclose(rew, r_ref, rtol=1e-9, atol=1e-12)
        ep_r += r_ref; ep_l += 1
        for e in np.nonzero(d_ref)[0]:
            info = infos[int(e)]
            np.testing.assert_allclose(info["terminal_observation"], o_ref[e], rtol=1e-6, atol=1e-6)
            assert info["TimeLimit.truncated"] == bool(tr_ref[e] and not te_ref[e])
            assert abs(info["episode"]["r"] - ep_r[e]) <= 1e-9 * max(1, abs(ep_r[e])) and info["episode"]["l"] == ep_l[e]
            ep_r[e] = 0; ep_l[e] = 0
            seen_done += 1
        not_done = np.nonzero(~d_ref)[0]
        if len(not_done):
            assert "terminal_observation" not in infos[int(not_done[0])]
        if d_ref.any():
            o_ref[d_ref] = ref.reset(d_ref.astype(np.uint8))[d_ref]
        np.testing.assert_allclose(obs, o_ref, rtol=1e-6, atol=1e-6)
    assert seen_done > 0
    assert venv.env_is_wrapped(object) == [False] * E and venv.get_attr("CTRL_FREQ")[0] == 30
    o, r, te, tr = venv.step_tensor(torch.zeros((E, 1, 4), device="cuda"))
    assert o.is_cuda and o.shape == (E, 1, 72) and te.dtype == torch.bool
    venv.close()


@pytest.mark.parametrize("desc,kw_over,E,steps", [
    ("single env, single drone", dict(), 1, 40),
    ("S=1, B=120: history box wider than a TMA box -> register copy", dict(ctrl_freq=240), 33, 30),
    ("B=1: ring holds only the newest action", dict(pyb_freq=6, ctrl_freq=3), 65, 12),
    ("B=2", dict(pyb_freq=20, ctrl_freq=4), 65, 12),
    ("3 drones, E not a multiple of the envs per block", dict(env_kind="multihover", num_drones=3), 43, 30),
    ("ONE_D_RPM (A=1, scalar rows)", dict(action_type="one_d_rpm"), 70, 30),
    ("DYN+DRAG single drone", dict(physics_flags=2), 70, 30),
    ("DYN+GND+DRAG+DW, 5 drones", dict(env_kind="multihover", num_drones=5, physics_flags=7), 21, 30),
    ("240 drones per RL env: FP64 block without the DMA warp", dict(env_kind="multihover", num_drones=240), 3, 10),
])
def test_cuda_f64_edge_shapes_vs_oracle(desc, kw_over, E, steps):
    """Ragged / extreme shapes: every code path of the step kernel (TMA and register history copy, all block tails)
    against the oracle in FP64."""
    rng = np.random.default_rng(abs(hash(desc)) % 2**31)
    kw = dict(model=DroneModel.CF2X, env_kind="hover", action_type="rpm", num_drones=1, pyb_freq=240, ctrl_freq=30,
              physics_flags=0, init_xyz=None, init_rpy=None)
    kw.update(kw_over)
    N = kw["num_drones"]
    A = {"rpm": 4, "one_d_rpm": 1}[kw["action_type"]]
    if N > 1:
        xyz = np.stack([rng.uniform(-.5, .5, (E, N)), rng.uniform(-.5, .5, (E, N)), rng.uniform(0.03, 1.0, (E, N))], -1)
        kw["init_xyz"], kw["init_rpy"] = xyz, rng.uniform(-.2, .2, (E, N, 3))
    ref = make_oracle(kw, num_envs=E)
    sim = make_sim(kw, num_envs=E)
    obs0 = sim.reset()
    assert np.max(np.abs(obs0.cpu().numpy() - ref.obs)) <= 1e-6
    for t in range(steps):
        a = (0.3 * rng.standard_normal(size=(E, N, A))).astype(np.float32)
        obs, rew, term, trunc = sim.step(torch.from_numpy(a).cuda())
        o_ref, r_ref, te_ref, tr_ref = ref.step(a)
        st, _, cnt = state_np(sim)
        rs = np.concatenate([ref.state20, ref.rpy_rates], axis=-1)
        for sl in (S_POS, S_VEL, S_RATES, S_ANGV, S_RPM):
            assert rel_err(st[..., sl], rs[..., sl]) <= 1e-9, (desc, t, sl)
        assert np.max(np.abs(obs.cpu().numpy().astype(np.float64) - o_ref) / np.maximum(np.abs(o_ref), 1.0)) <= 1e-6, (desc, t)
        assert np.array_equal(term.cpu().numpy(), te_ref) and np.array_equal(trunc.cpu().numpy(), tr_ref)
        assert np.array_equal(cnt, ref.step_counter)
    sim.close()


def test_cuda_ctrl_max_drones_per_env():
    """N = GPD_MAX_DRONES_PER_ENV (256): one block per env; CtrlAviary state20 observation, DYN+DW, vs the oracle."""
    rng = np.random.default_rng(77)
    E, N = 3, 256
    xyz = np.concatenate([rng.uniform(-3, 3, (E, N, 2)), rng.uniform(0.2, 3, (E, N, 1))], -1)
    kw = dict(model=DroneModel.CF2X, env_kind="ctrl", action_type="ctrl_rpm", num_drones=N, pyb_freq=240, ctrl_freq=48,
              physics_flags=4, init_xyz=xyz, init_rpy=np.zeros((E, N, 3)))
    ref = make_oracle(kw, num_envs=E)
    sim = make_sim(kw, num_envs=E)
    sim.reset()
    hover = load_drone_params(DroneModel.CF2X).HOVER_RPM
    for t in range(6):
        a = hover * (1 + 0.02 * rng.uniform(-1, 1, (E, N, 4)))
        obs, _, _, _ = sim.step(torch.from_numpy(a).cuda())
        o_ref, _, _, _ = ref.step(a)
        assert rel_err(obs.cpu().numpy()[..., 0:3], o_ref[..., 0:3]) <= 1e-9
        assert rel_err(obs.cpu().numpy()[..., 10:13], o_ref[..., 10:13]) <= 1e-8
    with pytest.raises(Exception):
        make_sim(dict(kw, num_drones=257, init_xyz=None, init_rpy=None), num_envs=1)
    sim.close()


def test_cuda_masked_reset_with_per_env_poses():
    """reset(mask) touches only the masked envs, uses each env's own initial pose, keeps the ring (obs history)."""
    rng = np.random.default_rng(31)
    E = 150
    xyz, rpy = _random_init(rng, E, 1)
    kw = dict(model=DroneModel.CF2P, env_kind="hover", action_type="rpm", num_drones=1, pyb_freq=240, ctrl_freq=30,
              physics_flags=0, init_xyz=xyz, init_rpy=rpy)
    ref = make_oracle(kw, num_envs=E)
    sim = make_sim(kw, num_envs=E)
    sim.reset()
    for t in range(20):
        a = rng.uniform(-1, 1, (E, 1, 4)).astype(np.float32)
        sim.step(torch.from_numpy(a).cuda()); ref.step(a)
        if t in (5, 11):
            mask = rng.uniform(size=E) < 0.3
            o = sim.reset(torch.from_numpy(mask.astype(np.uint8)))
            o_ref = ref.reset(mask.astype(np.uint8))
            assert np.max(np.abs(o.cpu().numpy() - o_ref)) <= 1e-6
            st, _, cnt = state_np(sim)
            assert np.all(cnt[mask] == 0) and np.all(cnt[~mask] > 0)
            assert rel_err(st[mask][:, 0, 0:3], xyz[mask][:, 0]) <= 1e-15
    st, _, _ = state_np(sim)
    assert rel_err(st[..., S_POS], ref.state20[..., S_POS]) <= 1e-9
    sim.close()


def test_cuda_graphed_rollout_matches_eager_stepping():
    """policy + env.step captured in one CUDA graph == the same loop stepped eagerly (bit-exact)."""
    from gpd_b200.envs import HoverAviary
    from gpd_b200.rollout import GraphedRollout
    torch.manual_seed(0)
    E, T = 200, 6
    W1 = (0.05 * torch.randn(72, 32)).cuda()
    W2 = (0.5 * torch.randn(32, 4)).cuda()

    def policy(obs):
        return torch.tanh(torch.tanh(obs.reshape(obs.shape[0], -1) @ W1) @ W2).reshape(obs.shape[0], 1, 4)

    env_g = HoverAviary(num_envs=E, auto_reset=True, precision="f32")
    env_e = HoverAviary(num_envs=E, auto_reset=True, precision="f32")
    env_e.reset()
    for _ in range(2 * T):                     # the collector's two warm-up rollouts advance the env: mirror them
        o = env_e._sim.obs
        env_e._sim.step(policy(o))
    ro = GraphedRollout(env_g, policy, T)      # capture itself does not execute
    for rep in range(2):
        obs, act, rew, term, trunc = ro.run()
        torch.cuda.synchronize()
        for t in range(T):
            o = env_e._sim.obs
            assert torch.equal(obs[t], o)
            a = policy(o)
            o2, r2, te2, tr2 = env_e._sim.step(a)
            assert torch.equal(act[t], a) and torch.equal(rew[t], r2)
            assert torch.equal(term[t], te2.view(torch.bool)) and torch.equal(trunc[t], tr2.view(torch.bool))
        assert torch.equal(obs[T], env_e._sim.obs)
        # exported state (ang_v / last_clipped_action are re-derived from the latest observation in this lean FP32 sim)
        assert all(torch.equal(x, y) for x, y in zip(env_g._sim.get_state(), env_e._sim.get_state()))
    del ro, obs, act, rew, term, trunc          # the trajectory buffers go away: the env must not depend on them
    torch.cuda.empty_cache()
    junk = torch.full((E * 73 * 8,), 7.0, device="cuda")
    assert all(torch.equal(x, y) for x, y in zip(env_g._sim.get_state(), env_e._sim.get_state()))
    a = policy(env_e._sim.obs)
    assert all(torch.equal(x, y) for x, y in zip(env_g._sim.step(a), env_e._sim.step(a)))
    del junk
    with pytest.raises(ValueError):
        GraphedRollout(env_g, policy, 3)
    env_g.close(); env_e.close()


@pytest.mark.parametrize("desc,kw_over,E,steps,tol", [
    ("DYN+GND+DRAG+DW 3 drones (MUFU downwash, predicate ground-effect gate)",
     dict(env_kind="multihover", num_drones=3, physics_flags=7), 64, 8, 1e-3),
    ("Ctrl 3 drones DYN+DW", dict(env_kind="ctrl", action_type="ctrl_rpm", num_drones=3, physics_flags=4, ctrl_freq=48), 64, 12, 1e-3),
    ("ActionType.PID 48 Hz (TMA ring load + shared-memory shift)", dict(action_type="pid", ctrl_freq=48, model=DroneModel.CF2P), 200, 12, 2e-3),
    ("ActionType.ONE_D_PID 48 Hz", dict(action_type="one_d_pid", ctrl_freq=48, model=DroneModel.CF2P), 200, 12, 2e-3),
    ("ActionType.VEL 48 Hz", dict(action_type="vel", ctrl_freq=48, model=DroneModel.CF2P), 200, 12, 2e-3),
    ("VelocityAviary", dict(env_kind="ctrl", action_type="ctrl_vel", num_drones=2, ctrl_freq=48, model=DroneModel.CF2P), 100, 12, 2e-3),
    ("ONE_D_RPM 48 Hz (A=1, 16-byte rows)", dict(action_type="one_d_rpm", ctrl_freq=48), 200, 60, 1e-4),
])
def test_cuda_f32_other_paths_vs_oracle(desc, kw_over, E, steps, tol):
    """FP32 throughput mode on every non-headline path (force models with fast math, controllers, A<4 rows) against the
    FP64 oracle.  Closed-loop controllers amplify rounding (SURVEY finding 6): short horizons, looser tolerance."""
    rng = np.random.default_rng(abs(hash(desc)) % 2**31)
    kw = dict(model=DroneModel.CF2X, env_kind="hover", action_type="rpm", num_drones=1, pyb_freq=240, ctrl_freq=30,
              physics_flags=0, init_xyz=None, init_rpy=None)
    kw.update(kw_over)
    N, act = kw["num_drones"], kw["action_type"]
    A = {"rpm": 4, "one_d_rpm": 1, "pid": 3, "one_d_pid": 1, "vel": 4, "ctrl_rpm": 4, "ctrl_vel": 4}[act]
    if N > 1:
        # heights 0.2 / 0.4 / 0.55 m: pair gaps 0.15-0.35 m stay away from both singularities of the reference's downwash
        # (alpha ~ 1/delta_z^2 at delta_z -> 0+, BaseAviary.py:802; beta = 0.16*delta_z - 0.11 -> 0 at 0.6875 m, :803-804),
        # where any rounding difference is amplified without bound and an FP32 comparison is meaningless
        z = np.array([0.2, 0.4, 0.55, 0.75, 0.9, 1.05])[None, :N] + rng.uniform(0, 0.01, (E, N))
        perm = np.argsort(rng.uniform(size=(E, N)), axis=1)
        z = np.take_along_axis(z, perm, axis=1)
        xyz = np.stack([rng.uniform(-.15, .15, (E, N)), rng.uniform(-.15, .15, (E, N)), z], -1)
        kw["init_xyz"], kw["init_rpy"] = xyz, rng.uniform(-.1, .1, (E, N, 3))
    ref = make_oracle(kw, num_envs=E)
    sim = make_sim(kw, num_envs=E, precision="f32")
    sim.reset()
    hover = load_drone_params(kw["model"]).HOVER_RPM
    worst = 0.0
    for t in range(steps):
        if act == "ctrl_rpm":
            a = hover * (1 + 0.02 * rng.uniform(-1, 1, (E, N, A)))
            dev_a = torch.from_numpy(a.astype(np.float32)).cuda()
            a = a.astype(np.float32).astype(np.float64)            # the oracle sees exactly the FP32 command
        elif act == "ctrl_vel":
            a = np.concatenate([rng.uniform(-1, 1, (E, N, 3)), rng.uniform(0, 1, (E, N, 1))], -1).astype(np.float32)
            dev_a = torch.from_numpy(a).cuda()
            a = a.astype(np.float64)
        else:
            a = (0.3 * rng.uniform(-1, 1, (E, N, A))).astype(np.float32)
            dev_a = torch.from_numpy(a).cuda()
        obs, rew, term, trunc = sim.step(dev_a)
        o_ref, r_ref, _, _ = ref.step(a)
        st, _, _ = state_np(sim)
        worst = max(worst, rel_err(st[..., S_POS], ref.state20[..., S_POS]), rel_err(st[..., S_VEL], ref.state20[..., S_VEL], floor=1e-2))
        if kw["env_kind"] != "ctrl":          # action ring part of the observation is bit-exact in every mode
            assert np.array_equal(obs.cpu().numpy()[..., 12:], o_ref[..., 12:]), (desc, t)
    assert worst <= tol, (desc, worst)
    sim.close()


def test_cuda_f32_force_models_vs_f64_kernels():
    """FP32 fast-math force kernels (MUFU reciprocals / ex2) against the FP64 kernels (which are pinned to the reference's
    recorded forces) on pairs kept away from the downwash singularities."""
    import ctypes as C
    from gpd_b200 import _lib
    L = _lib.load()
    rng = np.random.default_rng(123)
    d = _lib.drone_params_c(load_drone_params(DroneModel.CF2X))
    E, N = 4000, 2
    pos = np.zeros((E, N, 3))
    pos[:, 1, 0:2] = rng.uniform(-.3, .3, (E, 2))
    dz = np.where(rng.uniform(size=E) < 0.5, rng.uniform(0.05, 0.55, E), rng.uniform(0.85, 2.5, E))   # |beta| >= 0.022
    pos[:, 1, 2] = dz
    pos[:, 0, :] += rng.uniform(-1, 1, (E, 3))
    pos[:, 1, :] += pos[:, 0, :]
    out = {}
    for prec, dt in ((1, torch.float64), (0, torch.float32)):
        p_t = torch.as_tensor(pos, dtype=dt).cuda().contiguous()
        o = torch.empty((E, N), dtype=dt, device="cuda")
        _lib.check(L.gpd_force_downwash(0, prec, C.byref(d), E, N, C.c_void_p(p_t.data_ptr()), C.c_void_p(o.data_ptr()), None))
        out[prec] = o.double().cpu().numpy()
    ref, got = out[1][:, 0], out[0][:, 0]                 # drone 0 sits below drone 1
    assert np.all(out[1][:, 1] == 0) and np.all(out[0][:, 1] == 0)
    big = np.abs(ref) > 1e-6
    assert big.sum() > 500
    # inputs are rounded to FP32 first: d(force)/d(delta_z) ~ force * (2/dz + dxy^2*0.16/beta^3) amplifies that rounding
    assert np.max(np.abs(got[big] - ref[big]) / np.abs(ref[big])) <= 5e-4
    assert np.max(np.abs(got[~big] - ref[~big])) <= 1e-8
    # ground effect and drag
    n = 3000
    rpm = rng.uniform(9000, 20000, (n, 4)); p3 = rng.uniform([-1, -1, 0.01], [1, 1, 1.0], (n, 3)); vel = rng.uniform(-2, 2, (n, 3))
    q = rng.standard_normal((n, 4)); q /= np.linalg.norm(q, axis=1, keepdims=True)
    res = {}
    for prec, dt in ((1, torch.float64), (0, torch.float32)):
        t = lambda a: torch.as_tensor(a, dtype=dt).cuda().contiguous()
        r_, p_, q_, v_ = t(rpm), t(p3), t(q), t(vel)
        ge = torch.empty((n, 4), dtype=dt, device="cuda"); ok = torch.empty(n, dtype=torch.uint8, device="cuda")
        dr = torch.empty((n, 3), dtype=dt, device="cuda")
        pp = lambda x: C.c_void_p(x.data_ptr())
        _lib.check(L.gpd_force_ground_effect(0, prec, C.byref(d), n, pp(r_), pp(p_), pp(q_), pp(ge), pp(ok), None))
        _lib.check(L.gpd_force_drag(0, prec, C.byref(d), n, pp(r_), pp(q_), pp(v_), pp(dr), None))
        res[prec] = (ge.double().cpu().numpy(), ok.cpu().numpy(), dr.double().cpu().numpy())
    assert rel_err(res[0][0], res[1][0], floor=1e-9) <= 1e-4
    assert rel_err(res[0][2], res[1][2], floor=1e-9) <= 1e-4
    assert np.mean(res[0][1] == res[1][1]) > 0.995          # gate differs only within rounding of the +-pi/2 boundary


def test_cuda_full_size_properties_65536_envs():
    """BASELINE config size (65,536 HoverAviary envs, FP32): size-independent properties.
    (1) determinism; (2) env-permutation equivariance (envs are independent: row e depends only on env e's inputs);
    (3) a 256-env sample agrees with the FP64 oracle within the FP32 tolerance; (4) the observation ring shifts by exactly
    one slot per step and its newest slot is the action; (5) statistics count every env-step."""
    E, T = 65536, 24
    g = torch.Generator(device="cuda"); g.manual_seed(5)
    acts = [(torch.rand((E, 1, 4), generator=g, device="cuda") * 2 - 1) * 0.3 for _ in range(T)]
    perm = torch.randperm(E, generator=g, device="cuda")
    rng = np.random.default_rng(5)
    xyz, rpy = _random_init(rng, E, 1)
    kw = dict(model=DroneModel.CF2X, env_kind="hover", action_type="rpm", num_drones=1, pyb_freq=240, ctrl_freq=30,
              physics_flags=0, init_xyz=xyz, init_rpy=rpy)
    p = perm.cpu().numpy()
    kw_p = dict(kw, init_xyz=xyz[p], init_rpy=rpy[p])
    a_sim, b_sim, p_sim = (make_sim(kw, E, "f32", auto_reset=True), make_sim(kw, E, "f32", auto_reset=True),
                           make_sim(kw_p, E, "f32", auto_reset=True))
    sample = np.sort(rng.choice(E, 256, replace=False))
    ref = make_oracle(dict(kw, init_xyz=xyz[sample], init_rpy=rpy[sample]), num_envs=256)
    for s in (a_sim, b_sim, p_sim):
        s.reset()
    prev = a_sim.obs.clone()
    worst = 0.0
    alive = np.ones(256, bool)
    for t in range(T):
        oa, ra, ta, tra = a_sim.step(acts[t])
        ob, rb, tb, trb = b_sim.step(acts[t])
        op, rp, tp, trp = p_sim.step(acts[t][perm].contiguous())
        assert torch.equal(oa, ob) and torch.equal(ra, rb) and torch.equal(ta, tb) and torch.equal(tra, trb)      # (1)
        assert torch.equal(oa[perm], op) and torch.equal(ra[perm], rp) and torch.equal(tra[perm], trp)            # (2)
        assert torch.equal(oa[..., 12:68], prev[..., 16:72]) and torch.equal(oa[..., 68:72], acts[t])            # (4)
        prev = oa.clone()
        o_ref, r_ref, te_ref, tr_ref = ref.step(acts[t][sample].cpu().numpy())
        done = (te_ref | tr_ref).astype(bool)
        flags = (ta | tra).cpu().numpy().astype(bool)[sample]
        alive &= ~(flags != done)            # an FP32 flag flip right at a threshold ends the comparison for that env
        st = a_sim.get_state()[0].double().cpu().numpy()[sample]
        chk = alive & ~done
        if chk.any():
            worst = max(worst, rel_err(st[chk][:, 0, S_POS], ref.state20[chk][:, 0, S_POS]))
        if done.any():
            ref.reset(done.astype(np.uint8))
        alive &= ~done                       # after a reset the FP32/FP64 episodes restart in lockstep only if flags agreed
        alive |= done & (flags == done)
    assert alive.mean() > 0.9 and worst <= 1e-4, (alive.mean(), worst)                                            # (3)
    stats = a_sim.episode_stats()
    assert stats[6] == E * T and stats[0] > 0                                                                     # (5)
    for s in (a_sim, b_sim, p_sim):
        s.close()


def test_cuda_checkpoint_restore_and_helpers():
    """env.checkpoint()/restore() resume a HoverAviary run bit-exactly, ring included; batched helper methods."""
    from gpd_b200.control.DSLPIDControl import DSLPIDControl
    from gpd_b200.envs import HoverAviary
    g = torch.Generator(device="cuda"); g.manual_seed(1)
    E = 500
    acts = [(torch.rand((E, 1, 4), generator=g, device="cuda") * 2 - 1) * 0.2 for _ in range(12)]
    a = HoverAviary(num_envs=E, precision="f32")
    a.reset()
    for t in range(6):
        a.step(acts[t])
    ck = a.checkpoint()
    b = HoverAviary(num_envs=E, precision="f32")
    b.reset()
    b.restore(ck)
    for t in range(6, 12):
        oa, ra, _, _, _ = a.step(acts[t])
        ob, rb, _, _, _ = b.step(acts[t])
        assert torch.equal(oa, ob) and torch.equal(ra, rb)
    cur = torch.tensor([[0., 0., 0.], [0., 0., 0.]], device="cuda")
    dst = torch.tensor([[0.3, 0., 0.4], [3., 0., 4.]], device="cuda")
    nxt = a._calculateNextStep(cur, dst, 1)
    assert torch.allclose(nxt, torch.tensor([[0.3, 0., 0.4], [0.6, 0., 0.8]], device="cuda"))
    c = DSLPIDControl(DroneModel.CF2X, num=1)
    pwm = c._one23DInterface(torch.tensor([[0.3]], dtype=torch.float64))
    want = min(max((np.sqrt(0.3 / (c.KF * 4)) - 4070.3) / 0.2685, 20000), 65535)
    assert pwm.shape == (1, 4) and abs(float(pwm[0, 0]) - want) < 1e-6
    a.close(); b.close()


@pytest.mark.parametrize("model", ["cf2x", "cf2p"])
def test_cuda_pidpy_config0_rollout(model):
    """BASELINE configs[0] (examples/pid.py on Physics.DYN): the in-kernel loop gpd_rollout_pid, one launch per ctrl step
    so every step can be compared with the reference's recorded observations and actions (FP64)."""
    from gpd_b200.sim import BatchedSim
    g = load_golden("pidpy_dyn.npz")
    xyz, rpy, wps = g[model + "_init_xyz"], g[model + "_init_rpy"], g[model + "_waypoints"]
    n, E = xyz.shape[0], 5
    dm = DroneModel(model)
    sim = BatchedSim(load_drone_params(dm), E, n, env_kind="ctrl", action_type="ctrl_rpm", pyb_freq=240, ctrl_freq=48,
                     precision="f64", pid=default_pid_params(dm), init_xyz=xyz, init_rpy=rpy)
    wp = torch.as_tensor(np.tile(g[model + "_wp0"][None], (E, 1)), dtype=torch.int32).cuda().contiguous()
    act = torch.zeros((E, n, 4), dtype=torch.float64, device="cuda")
    wps_t = torch.as_tensor(wps).cuda()
    ref_obs, ref_act = g[model + "_obs"], g[model + "_actions"]
    for t in range(ref_obs.shape[0]):
        sim.rollout_pid(1, wps_t, wp, act)
        st = sim.get_state()[0].cpu().numpy()
        tol = 1e-9 if t < 20 else 1e-6
        for e in (0, E - 1):
            assert rel_err(st[e][:, 0:3], ref_obs[t][:, 0:3]) <= tol, t
            assert quat_err(st[e][:, 3:7], ref_obs[t][:, 3:7]) <= tol, t
            assert rel_err(act.cpu().numpy()[e], ref_act[t]) <= tol * 10, t
    # and the whole thing in ONE launch reproduces the stepwise result
    sim2 = BatchedSim(load_drone_params(dm), E, n, env_kind="ctrl", action_type="ctrl_rpm", pyb_freq=240, ctrl_freq=48,
                      precision="f64", pid=default_pid_params(dm), init_xyz=xyz, init_rpy=rpy)
    wp2 = torch.as_tensor(np.tile(g[model + "_wp0"][None], (E, 1)), dtype=torch.int32).cuda().contiguous()
    act2 = torch.zeros((E, n, 4), dtype=torch.float64, device="cuda")
    sim2.rollout_pid(ref_obs.shape[0], wps_t, wp2, act2)
    assert torch.equal(sim2.get_state()[0], sim.get_state()[0]) and torch.equal(act2, act) and torch.equal(wp2, wp)
    sim.close(); sim2.close()


def test_cuda_example_scripts_run():
    """Batched counterparts of the reference's examples (its own tests/test_examples.py only checks they run)."""
    from gpd_b200.examples import downwash, pid
    e1 = pid.run(num_envs=64, num_drones=3, duration_sec=4, fused=False)
    e2 = pid.run(num_envs=64, num_drones=3, duration_sec=4, fused=True)
    assert e1 < 0.1 and e2 < 0.1                      # CF2P tracks the circle on Physics.DYN (SURVEY finding 5)
    z_dw = downwash.run(num_envs=16, duration_sec=6)
    z_no = downwash.run(num_envs=16, duration_sec=6, physics=Physics.DYN)
    assert z_dw < z_no - 1e-4                         # the wake pushes the lower drone down


def test_cuda_count_nonfinite():
    kw = dict(model=DroneModel.CF2X, env_kind="hover", action_type="rpm", num_drones=1, pyb_freq=240, ctrl_freq=30,
              physics_flags=0, init_xyz=None, init_rpy=None)
    sim = make_sim(kw, num_envs=1000, precision="f32")
    sim.reset()
    assert sim.count_nonfinite() == 0
    a = torch.zeros((1000, 1, 4), device="cuda")
    a[7] = float("nan"); a[500, 0, 2] = float("inf")
    sim.step(a)
    assert sim.count_nonfinite() == 2
    sim.close()


@pytest.mark.parametrize("precision", ["f64", "f32"])
@pytest.mark.parametrize("N,flags,act", [(1, 0, "rpm"), (1, 3, "rpm"), (3, 7, "rpm"), (1, 0, "pid")])
def test_cuda_block_size_does_not_change_results(precision, N, flags, act):
    """The block size (threads_per_block, or the occupancy-aware default of gpd_create) is a pure scheduling choice:
    every layout must give bit-identical observations, rewards, flags and state, with auto-reset on."""
    rng = np.random.default_rng(5)
    E = 517
    xyz = np.stack([rng.uniform(-.5, .5, (E, N)), rng.uniform(-.5, .5, (E, N)), rng.uniform(0.05, 1.0, (E, N))], -1)
    kw = dict(model=DroneModel.CF2X, env_kind="hover" if N == 1 else "multihover", action_type=act, num_drones=N, pyb_freq=240,
              ctrl_freq=48 if act == "pid" else 30, physics_flags=flags, init_xyz=xyz, init_rpy=rng.uniform(-.2, .2, (E, N, 3)))
    A = 3 if act == "pid" else 4
    acts = [torch.from_numpy(rng.uniform(-1, 1, (E, N, A)).astype(np.float32)).cuda() for _ in range(25)]
    outs = []
    for tpb in (0, 32, 96, 128, 224, 256):
        sim = make_sim(kw, num_envs=E, precision=precision, auto_reset=True, tpb=tpb)
        sim.reset()
        rec = []
        for a in acts:
            o, r, te, tr = sim.step(a)
            rec.append((o.clone(), r.clone(), te.clone(), tr.clone()))
        st, rr, ps, cnt = sim.get_state()
        outs.append((rec, st.clone(), rr.clone(), cnt.clone(), sim.episode_stats()))
        sim.close()
    ref = outs[0]
    for o in outs[1:]:
        for (a0, a1, a2, a3), (b0, b1, b2, b3) in zip(ref[0], o[0]):
            assert torch.equal(a0, b0) and torch.equal(a1, b1) and torch.equal(a2, b2) and torch.equal(a3, b3)
        assert torch.equal(ref[1], o[1]) and torch.equal(ref[2], o[2]) and torch.equal(ref[3], o[3])
        assert ref[4][0] == o[4][0] and ref[4][2] == o[4][2] and ref[4][6] == o[4][6]      # episodes, lengths, env-steps
        assert abs(ref[4][1] - o[4][1]) <= 1e-6 * max(1.0, abs(ref[4][1]))                # sum of returns (FP32 partial sums)


@pytest.mark.parametrize("name", ["c3_multihover2_gnd_drag_f64", "c4_ctrl64_dw_f64", "c4_ctrl64_dw_f32", "c5_hover_pid_f32"])
def test_cuda_full_size_properties_other_configs(name):
    """BASELINE.json configs 3-5 at their full sizes, through size-independent properties: (1) determinism,
    (2) env-permutation equivariance, (3) a sample of envs against the FP64 oracle fed the same actions."""
    rng = np.random.default_rng(11)
    g = torch.Generator(device="cuda"); g.manual_seed(11)
    hover = load_drone_params(DroneModel.CF2X).HOVER_RPM
    if name.startswith("c3"):
        E, N, A, T, prec, ns, tol = 32768, 2, 4, 12, "f64", 128, 1e-9
        xyz, rpy = _random_init(rng, E, N)
        kw = dict(model=DroneModel.CF2X, env_kind="multihover", action_type="rpm", num_drones=N, pyb_freq=240, ctrl_freq=30,
                  physics_flags=3, init_xyz=xyz, init_rpy=rpy)
        acts = [(torch.rand((E, N, A), generator=g, device="cuda") * 2 - 1) for _ in range(T)]
    elif name.startswith("c4"):
        prec = name[-3:]
        # FP32 tolerance: the 64-drone downwash field is stiff (dF/dz = 2F/dz at dz = 0.045 m), rounding differences grow ~10x per
        # 5 ctrl steps; measured 7.6e-4
        E, N, A, T, ns, tol = 4096, 64, 4, 5, 6, (1e-9 if prec == "f64" else 2e-3)
        # FP64 uses the BASELINE box; FP32 gives every drone of an env its own height, 0.045 m apart in a random order: the
        # downwash model is singular at dz -> 0+ (alpha ~ 1/dz^2) and amplifies any rounding difference there
        z = (rng.uniform(0.2, 3, (E, N, 1)) if prec == "f64"
             else 0.2 + 0.045 * np.argsort(rng.random((E, N)), axis=1)[..., None].astype(np.float64))
        xyz = np.concatenate([rng.uniform(-2, 2, (E, N, 2)), z], -1)
        rpy = np.zeros((E, N, 3))
        kw = dict(model=DroneModel.CF2X, env_kind="ctrl", action_type="ctrl_rpm", num_drones=N, pyb_freq=240, ctrl_freq=48,
                  physics_flags=4, init_xyz=xyz, init_rpy=rpy)
        dt = torch.float64 if prec == "f64" else torch.float32
        acts = [(hover * (1 + 0.02 * (torch.rand((E, N, A), generator=g, device="cuda", dtype=torch.float64) * 2 - 1))).to(dt)
                for _ in range(T)]
    else:
        E, N, A, T, prec, ns, tol = 2097152, 1, 3, 6, "f32", 256, 1e-4
        xyz, rpy = _random_init(rng, E, N)
        kw = dict(model=DroneModel.CF2X, env_kind="hover", action_type="pid", num_drones=N, pyb_freq=240, ctrl_freq=48,
                  physics_flags=0, init_xyz=xyz, init_rpy=rpy)
        acts = [(torch.rand((E, N, A), generator=g, device="cuda") * 2 - 1) for _ in range(T)]
    perm = torch.randperm(E, generator=g, device="cuda")
    p = perm.cpu().numpy()
    sims = [make_sim(kw, E, prec), make_sim(kw, E, prec), make_sim(dict(kw, init_xyz=xyz[p], init_rpy=rpy[p]), E, prec)]
    sample = np.sort(rng.choice(E, ns, replace=False))
    ref = make_oracle(dict(kw, init_xyz=xyz[sample], init_rpy=rpy[sample]), num_envs=ns)
    for s in sims:
        s.reset()
    worst = 0.0
    for t in range(T):
        oa, ra, ta, tra = sims[0].step(acts[t])
        ob, rb, tb, trb = sims[1].step(acts[t])
        op, rp, tp, trp = sims[2].step(acts[t][perm].contiguous())
        assert torch.equal(oa, ob) and torch.equal(ra, rb) and torch.equal(ta, tb) and torch.equal(tra, trb)      # (1)
        assert torch.equal(oa[perm], op) and torch.equal(ra[perm], rp) and torch.equal(tra[perm], trp)            # (2)
        ref.step(acts[t][sample].cpu().numpy())
        st = sims[0].get_state()[0].double().cpu().numpy()[sample]
        worst = max(worst, rel_err(st[..., S_POS], ref.state20[..., S_POS]), rel_err(st[..., S_VEL], ref.state20[..., S_VEL]))
    assert worst <= tol, worst                                                                                    # (3)
    assert sims[0].count_nonfinite() == 0
    for s in sims:
        s.close()


@pytest.mark.parametrize("extra", [[], ["--streams", "2"], ["--no-graph"], ["--steps", "37"]])
def test_cuda_bench_contract_line(extra):
    """bench.py (product arm) on a small workload: ONE JSON line with the agreed keys, exactly K launches, a live roofline."""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, os.path.join(root, "bench.py"), "--steps", "64", "--warmup", "4", "--envs", "4096", "--sets", "2",
           "--e2e-steps", "3", "--cpu-seconds", "0.5", "--no-others", "--ref-kind", "port"] + extra
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=root)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip().startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    K = 37 if "--steps" in extra else 64
    for k in ["metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "roofline", "cpu_baseline"]:
        assert k in d, k
    assert d["metric"] == "drone-substeps/sec" and d["steps"] == K and d["gpu_launches"] == K and d["n_gpus"] == 1
    assert d["value"] > 0 and d["vs_baseline"] is None and d["dtype"] == "f32" and "workload" in d["config"]
    assert d["episode_stats"]["env_steps"] >= 4096 * K              # every timed launch stepped every env (warm-up on top)
    r = d["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    assert abs(r["achieved"] - 646 * 4096 / (d["ms_per_step"] * 1e-3) / 1e9) < 1e-6 * r["achieved"]
    # the device sends back only what it computed: 12 kin floats + reward + 2 flags per env (the ring is the host's own data)
    assert d["e2e"]["h2d_bytes_per_step"] == 4096 * 16 and d["e2e"]["d2h_bytes_per_step"] == 4096 * (12 * 4 + 6)
    assert d["e2e"]["host_obs_equals_device_obs"] is True
    assert len(d["trials_ms"]) == 5 and d["rank_ms"]["max"] >= d["rank_ms"]["min"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["value"] > 0
    assert d["measurement"]["streams"] == (2 if "--streams" in extra else 1) and "l2" in d["config"]
    if not extra or "--steps" in extra:                             # informational multi-stream figure beside the headline
        assert d["async_pools"]["streams"] == 2 and d["async_pools"]["ms_per_step"] > 0
    else:
        assert d["async_pools"] is None


def test_cuda_async_env_pools_match_sequential_stepping():
    """AsyncEnvPools: pools stepping on their own streams give exactly the results of stepping them one after the other."""
    from gpd_b200.envs import HoverAviary
    from gpd_b200.pool import AsyncEnvPools
    E, T, P = 3000, 12, 3
    g = torch.Generator(device="cuda"); g.manual_seed(3)
    acts = [[torch.rand((E, 1, 4), generator=g, device="cuda") * 2 - 1 for _ in range(P)] for _ in range(T)]
    seq = [HoverAviary(num_envs=E, precision="f32", auto_reset=True) for _ in range(P)]
    pools = AsyncEnvPools([HoverAviary(num_envs=E, precision="f32", auto_reset=True) for _ in range(P)])
    for e in seq:
        e.reset()
    obs0 = pools.reset()
    assert all(torch.equal(o, e._sim.obs) for o, e in zip(obs0, seq))
    for t in range(T):
        ref = [tuple(x.clone() for x in e._sim.step(a)) for e, a in zip(seq, acts[t])]
        if t % 2:
            out = pools.step_all(acts[t])
        else:                                  # interleaved use: launch all, consume in reverse order
            for j in range(P):
                pools.step_async(j, acts[t][j])
            out = [None] * P
            for j in reversed(range(P)):
                out[j] = pools.wait(j)
        for r, o in zip(ref, out):
            assert all(torch.equal(a, b) for a, b in zip(r, o))
    with pytest.raises(RuntimeError):
        pools.wait(0)
    pools.close()
    for e in seq:
        e.close()


@pytest.mark.parametrize("act,N", [("rpm", 1), ("one_d_rpm", 1), ("rpm", 3)])
def test_cuda_f32_lean_state_export_matches_explicit_aux_arrays(act, N, monkeypatch):
    """Lean FP32 KIN sims do not write ang_v / last_clipped_action per step: gpd_get_state re-derives them from the
    observation row.  A twin sim built with GPD_AUX_ALWAYS=1 (arrays written every step) must export bit-identical
    state through steps, auto-resets, masked resets, set_state and the host path."""
    rng = np.random.default_rng(9)
    E = 700
    xyz, rpy = _random_init(rng, E, N)
    kw = dict(model=DroneModel.CF2X, env_kind="hover" if N == 1 else "multihover", action_type=act, num_drones=N, pyb_freq=240,
              ctrl_freq=30, physics_flags=0, init_xyz=xyz, init_rpy=rpy)
    A = 4 if act == "rpm" else 1
    lean = make_sim(kw, E, "f32", auto_reset=True)
    monkeypatch.setenv("GPD_AUX_ALWAYS", "1")
    twin = make_sim(kw, E, "f32", auto_reset=True)
    monkeypatch.delenv("GPD_AUX_ALWAYS")

    def same():
        a, b = lean.get_state(), twin.get_state()
        for x, y in zip(a, b):
            assert torch.equal(x, y)
        return a
    lean.reset(); twin.reset()
    same()
    nonzero = False
    for t in range(30):
        a = torch.from_numpy(rng.uniform(-1, 1, (E, N, A)).astype(np.float32)).cuda()
        if t == 17:                            # host-buffer path in the middle of the run
            ol = lean.step_host(a.cpu().numpy()); ot = twin.step_host(a.cpu().numpy())
            assert all(np.array_equal(x, y) for x, y in zip(ol[:4], ot[:4]))
        else:
            ol, ot = lean.step(a), twin.step(a)
            assert all(torch.equal(x, y) for x, y in zip(ol, ot))
        st = same()
        nonzero |= bool((st[0][..., 16:20] != 0).any()) and bool((st[0][..., 13:16] != 0).any())
        if t == 9:                             # masked reset: untouched envs keep ang_v / rpm, reset ones read zero
            mask = torch.from_numpy((rng.random(E) < 0.3).astype(np.uint8)).cuda()
            assert torch.equal(lean.reset(mask), twin.reset(mask))
            st = same()
            m = mask.bool().cpu()
            assert (st[0].cpu()[m][..., 13:20] == 0).all() and (st[0].cpu()[~m][..., 16:20] != 0).any()
        if t == 20:                            # set_state makes the arrays authoritative until the next step
            s20, rr, ps, cnt = [x.clone() for x in st]
            s20[..., 13:20] = torch.rand_like(s20[..., 13:20])
            lean.set_state(s20, rr, ps, cnt); twin.set_state(s20, rr, ps, cnt)
            assert torch.equal(same()[0][..., 13:20], s20[..., 13:20])
    assert nonzero
    lean.close(); twin.close()


def test_cuda_graphed_pool_rollout_matches_single_pool_rollouts():
    """GraphedPoolRollout (one graph, one branch per pool) == independent GraphedRollouts of the same pools, bit-exact."""
    from gpd_b200.envs import HoverAviary
    from gpd_b200.rollout import GraphedPoolRollout, GraphedRollout
    torch.manual_seed(1)
    E, T, P = 300, 4, 3
    W1 = (0.05 * torch.randn(72, 16)).cuda()
    W2 = (0.5 * torch.randn(16, 4)).cuda()

    def policy(obs):
        return torch.tanh(torch.tanh(obs.reshape(obs.shape[0], -1) @ W1) @ W2).reshape(obs.shape[0], 1, 4)

    rng = np.random.default_rng(2)
    inits = [_random_init(rng, E, 1) for _ in range(P)]
    mk = lambda j: HoverAviary(num_envs=E, auto_reset=True, precision="f32", initial_xyzs=inits[j][0], initial_rpys=inits[j][1])
    pool = GraphedPoolRollout([mk(j) for j in range(P)], policy, T)
    singles = [GraphedRollout(mk(j), policy, T) for j in range(P)]
    for rep in range(2):
        obs, act, rew, term, trunc = pool.run()
        torch.cuda.synchronize()
        for j in range(P):
            o1, a1, r1, te1, tr1 = singles[j].run()
            torch.cuda.synchronize()
            assert torch.equal(obs[j], o1) and torch.equal(act[j], a1) and torch.equal(rew[j], r1)
            assert torch.equal(term[j], te1) and torch.equal(trunc[j], tr1)
    assert not torch.equal(obs[0], obs[1])          # the pools really are different simulations
    with pytest.raises(ValueError):
        GraphedPoolRollout([], policy, T)


def _fuzz_config(seed):
    rng = np.random.default_rng(1000 + seed)
    env_kind = rng.choice(["hover", "multihover", "ctrl"])
    N = 1 if env_kind == "hover" else int(rng.choice([1, 2, 3, 5, 9, 33]))
    if env_kind == "multihover" and N == 1:
        N = 2
    freq = [(240, 30), (240, 48), (240, 240), (120, 30), (96, 48), (60, 20)][rng.integers(6)]
    model = [DroneModel.CF2X, DroneModel.CF2P, DroneModel.RACE][rng.integers(3)]
    if env_kind == "ctrl":
        act = rng.choice(["ctrl_rpm", "ctrl_vel"])
    else:
        act = rng.choice(["rpm", "one_d_rpm", "pid", "vel", "one_d_pid"])
    if act in ("pid", "vel", "one_d_pid", "ctrl_vel") and model == DroneModel.RACE:
        model = DroneModel.CF2X                                   # no controller for the racer (BaseRLAviary.py:77-78)
    flags = int(rng.integers(8)) if N > 1 else int(rng.integers(4))   # downwash needs neighbours
    E = int(rng.choice([1, 7, 63, 64, 65, 130, 257]))
    return rng, dict(model=model, env_kind=env_kind, action_type=act, num_drones=N, pyb_freq=freq[0], ctrl_freq=freq[1],
                     physics_flags=flags), E, int(rng.choice([0, 32, 64, 96, 160, 256]))


import os as _os


@pytest.mark.parametrize("seed", range(int(_os.environ.get("GPD_FUZZ_SEEDS", "24"))))     # GPD_FUZZ_SEEDS=400 for a campaign
def test_cuda_f64_fuzz_vs_oracle(seed):
    """Differential fuzzing: random (model, env, N, frequencies, action type, force models, ragged E, block size) against
    the oracle, FP64, 12 ctrl steps from random poses; everything the step returns plus the exported state."""
    rng, kw, E, tpb = _fuzz_config(seed)
    N = kw["num_drones"]
    A = {"rpm": 4, "one_d_rpm": 1, "pid": 3, "vel": 4, "one_d_pid": 1, "ctrl_rpm": 4, "ctrl_vel": 4}[kw["action_type"]]
    xyz = np.stack([rng.uniform(-1, 1, (E, N)), rng.uniform(-1, 1, (E, N)), rng.uniform(0.05, 1.5, (E, N))], -1)
    if kw["physics_flags"] & 4:                                   # keep clear of the downwash singularities
        xyz[..., 2] = 0.2 + 0.11 * np.argsort(rng.random((E, N)), axis=1)
    kw.update(init_xyz=xyz, init_rpy=rng.uniform(-.2, .2, (E, N, 3)))
    hover = load_drone_params(kw["model"]).HOVER_RPM
    ref = make_oracle(kw, num_envs=E)
    sim = make_sim(kw, num_envs=E, tpb=tpb)
    obs0 = sim.reset()
    assert np.max(np.abs(obs0.double().cpu().numpy() - ref.obs)) <= 1e-6, kw
    for t in range(12):
        if kw["action_type"] == "ctrl_rpm":
            a = hover * (1 + 0.05 * rng.uniform(-1, 1, (E, N, A)))
            at = torch.from_numpy(a).cuda()
        elif kw["action_type"] == "ctrl_vel":
            a = rng.uniform(-1, 1, (E, N, A))
            at = torch.from_numpy(a).cuda()
        else:
            a = (0.3 * rng.standard_normal((E, N, A))).astype(np.float32)
            at = torch.from_numpy(a).cuda()
        obs, rew, term, trunc = sim.step(at)
        o_ref, r_ref, te_ref, tr_ref = ref.step(a)
        st, _, cnt = state_np(sim)
        rs = np.concatenate([ref.state20, ref.rpy_rates], axis=-1)
        tol = 1e-9
        for sl, nm in ((S_POS, "pos"), (S_VEL, "vel"), (S_RATES, "rates"), (S_ANGV, "ang_v"), (S_RPM, "rpm")):
            assert rel_err(st[..., sl], rs[..., sl]) <= tol, (kw, E, tpb, t, nm)
        assert quat_err(st[..., S_QUAT], rs[..., S_QUAT]) <= tol, (kw, t)
        assert np.array_equal(cnt, ref.step_counter)
        assert np.max(np.abs(obs.double().cpu().numpy() - o_ref) / np.maximum(np.abs(o_ref), 1.0)) <= 1e-6, (kw, t)
        assert np.max(np.abs(rew.double().cpu().numpy() - r_ref) / np.maximum(np.abs(r_ref), 1e-3)) <= 10 * tol
        assert np.array_equal(term.cpu().numpy(), te_ref) and np.array_equal(trunc.cpu().numpy(), tr_ref), (kw, t)
    sim.close()


@pytest.mark.parametrize("seed", range(int(_os.environ.get("GPD_FUZZ_SEEDS", "16"))))
def test_cuda_f32_fuzz_vs_oracle(seed):
    """FP32 throughput mode under the same random configurations (RPM-type actions, no downwash: the well-conditioned
    paths), 8 ctrl steps: position / velocity within 1e-4 of the FP64 oracle, the exported rpm exact to FP32 rounding."""
    rng, kw, E, tpb = _fuzz_config(seed)
    kw["action_type"] = "ctrl_rpm" if kw["env_kind"] == "ctrl" else ("rpm" if seed % 3 else "one_d_rpm")
    kw["physics_flags"] &= 3
    N = kw["num_drones"]
    A = 1 if kw["action_type"] == "one_d_rpm" else 4
    xyz = np.stack([rng.uniform(-1, 1, (E, N)), rng.uniform(-1, 1, (E, N)), rng.uniform(0.05, 1.5, (E, N))], -1)
    kw.update(init_xyz=xyz, init_rpy=rng.uniform(-.2, .2, (E, N, 3)))
    hover = load_drone_params(kw["model"]).HOVER_RPM
    ref = make_oracle(kw, num_envs=E)
    sim = make_sim(kw, num_envs=E, precision="f32", tpb=tpb)
    sim.reset()
    for t in range(8):
        if kw["action_type"] == "ctrl_rpm":
            a32 = (hover * (1 + 0.05 * rng.uniform(-1, 1, (E, N, A)))).astype(np.float32)
            a = a32.astype(np.float64)                             # the oracle sees exactly the float32 command
        else:
            a32 = a = (0.3 * rng.standard_normal((E, N, A))).astype(np.float32)
        sim.step(torch.from_numpy(a32).cuda())
        ref.step(a)
        st, _, cnt = state_np(sim)
        rs = np.concatenate([ref.state20, ref.rpy_rates], axis=-1)
        assert rel_err(st[..., S_POS], rs[..., S_POS]) <= 1e-4 and rel_err(st[..., S_VEL], rs[..., S_VEL]) <= 1e-4, (kw, E, tpb, t)
        assert rel_err(st[..., S_RPM], rs[..., S_RPM]) <= 1e-6 and rel_err(st[..., S_ANGV], rs[..., S_ANGV]) <= 2e-3, (kw, t)
        assert np.array_equal(cnt, ref.step_counter)
    sim.close()


@pytest.mark.parametrize("seed", range(int(_os.environ.get("GPD_FUZZ_SEEDS", "12"))))
def test_cuda_f64_fuzz_auto_reset_vs_oracle(seed):
    """The bench path's in-kernel auto-reset under random RL configurations: the oracle is reset by hand wherever it reports
    an episode end; observations, flags, exported state and Monitor-style statistics must keep agreeing for 40 steps."""
    rng, kw, E, tpb = _fuzz_config(seed + 5000)
    if kw["env_kind"] == "ctrl":
        kw["env_kind"], kw["num_drones"] = "hover", 1
    kw["action_type"] = ["rpm", "one_d_rpm", "pid", "one_d_pid"][seed % 4]
    if kw["action_type"] in ("pid", "one_d_pid") and kw["model"] == DroneModel.RACE:
        kw["model"] = DroneModel.CF2P
    N = kw["num_drones"]
    kw["physics_flags"] &= 3
    A = {"rpm": 4, "one_d_rpm": 1, "pid": 3, "one_d_pid": 1}[kw["action_type"]]
    xyz = np.stack([rng.uniform(-1, 1, (E, N)), rng.uniform(-1, 1, (E, N)), rng.uniform(0.05, 1.5, (E, N))], -1)
    kw.update(init_xyz=xyz, init_rpy=rng.uniform(-.2, .2, (E, N, 3)))
    ref = make_oracle(kw, num_envs=E)
    sim = make_sim(kw, num_envs=E, auto_reset=True, tpb=tpb)
    sim.reset()
    n_ep, sum_len, ep_len = 0, 0, np.zeros(E, int)
    for t in range(40):
        a = rng.uniform(-1, 1, (E, N, A)).astype(np.float32)
        obs, rew, term, trunc = sim.step(torch.from_numpy(a).cuda())
        o_ref, r_ref, te_ref, tr_ref = ref.step(a)
        o_ref = o_ref.copy()
        assert np.array_equal(term.cpu().numpy(), te_ref) and np.array_equal(trunc.cpu().numpy(), tr_ref), (kw, E, t)
        assert np.max(np.abs(rew.cpu().numpy() - r_ref) / np.maximum(np.abs(r_ref), 1e-3)) <= 1e-8
        done = (te_ref | tr_ref).astype(bool)
        ep_len += 1
        n_ep += int(done.sum()); sum_len += int(ep_len[done].sum()); ep_len[done] = 0
        if done.any():
            o_ref[done] = ref.reset(done.astype(np.uint8))[done]
        assert np.max(np.abs(obs.double().cpu().numpy() - o_ref) / np.maximum(np.abs(o_ref), 1.0)) <= 1e-6, (kw, E, t)
        st, _, cnt = state_np(sim)
        rs = np.concatenate([ref.state20, ref.rpy_rates], axis=-1)
        tol = 1e-6 if kw["action_type"] in ("pid", "one_d_pid") else 1e-9     # closed loops amplify libm-level differences
        for sl in (S_POS, S_VEL, S_RATES, S_ANGV, S_RPM):
            assert rel_err(st[..., sl], rs[..., sl]) <= tol, (kw, E, t, sl)
        assert np.array_equal(cnt, ref.step_counter)
    stats = sim.episode_stats()
    assert stats[0] == n_ep and stats[2] == sum_len and stats[6] == 40 * E
    sim.close()
