import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gpd_b200  # noqa
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
for _ in range(2 * T):
    o = env_e._sim.obs
    env_e._sim.step(policy(o))
ro = GraphedRollout(env_g, policy, T)
for rep in range(2):
    obs, act, rew, term, trunc = ro.run()
    torch.cuda.synchronize()
    for t in range(T):
        o = env_e._sim.obs
        if not torch.equal(obs[t], o):
            bad = (obs[t] != o).nonzero()
            print("rep", rep, "t", t, "obs differ at", bad[:8].tolist(), "n", len(bad), obs[t][tuple(bad[0])].item(), o[tuple(bad[0])].item())
        a = policy(o)
        o2, r2, te2, tr2 = env_e._sim.step(a)
        if not torch.equal(act[t], a): print("rep", rep, "t", t, "act differ", (act[t] != a).sum().item())
        if not torch.equal(rew[t], r2): print("rep", rep, "t", t, "rew differ", (rew[t] != r2).sum().item())
    if not torch.equal(obs[T], env_e._sim.obs):
        bad = (obs[T] != env_e._sim.obs).nonzero()
        print("rep", rep, "final obs differ", bad[:8].tolist(), len(bad))
print("done")
