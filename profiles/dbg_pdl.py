import os, sys, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import gpd_b200
from test_gpu_round2 import make_sim, _kw
from gpd_b200.params import load_drone_params
from gpd_b200.utils.enums import DroneModel

def run(kw, E, prec, ar, env, reps=20, glen=16):
    for k, v in env.items(): os.environ[k] = v
    sim = make_sim(kw, E, prec, auto_reset=ar)
    for k in env: del os.environ[k]
    rng = np.random.default_rng(5)
    A, N = sim.A, sim.N
    if kw["env_kind"] == "ctrl":
        hov = load_drone_params(kw["model"]).HOVER_RPM
        acts = [torch.from_numpy((hov * (1 + 0.05 * rng.uniform(-1, 1, (E, N, A)))).astype(np.float32)).cuda() for _ in range(4)]
    else:
        acts = [torch.from_numpy(rng.uniform(-1, 1, (E, N, A)).astype(np.float32)).cuda() for _ in range(4)]
    sim.reset()
    for k in range(4): sim.step(acts[k])
    torch.cuda.synchronize()
    side = torch.cuda.Stream(); g = torch.cuda.CUDAGraph()
    with torch.cuda.stream(side):
        with torch.cuda.graph(g, stream=side):
            for k in range(glen): sim.step(acts[k % 4])
    for _ in range(reps): g.replay()
    torch.cuda.synchronize()
    st = sim.get_state()
    out = [x.clone() for x in st] + [sim.obs.clone()]
    sim.close()
    return out

kw = _kw("ctrl", "ctrl_rpm", 2, 48, int(sys.argv[1]) if len(sys.argv) > 1 else 4)
E = 5000
ser = {"GPD_TILE_DEP": "0", "GPD_PDL": "0"}
for reps in (1, 3, 20):
    a = run(kw, E, "f32", False, {}, reps); b = run(kw, E, "f32", False, {}, reps)
    c = run(kw, E, "f32", False, ser, reps); d = run(kw, E, "f32", False, ser, reps)
    def diff(x, y): return [int((~((u == v) | (torch.isnan(u) & torch.isnan(v)))).sum()) if u.dtype.is_floating_point else int((u != v).sum()) for u, v in zip(x, y)]
    print("reps", reps, "pdl-vs-pdl", diff(a, b), "serial-vs-serial", diff(c, d), "pdl-vs-serial", diff(a, c), "nan count", int(torch.isnan(a[0]).sum()), flush=True)
