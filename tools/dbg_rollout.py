import sys, numpy as np, torch
sys.path.insert(0,'/root/repo')
from tests.helpers import make_config, obs_diff, flag_slots
from solorl_b200.envs import SoloVecEnv
from oracle.oracle import OracleEnv
rng = np.random.default_rng(25)
n, nref = 64, 6
cfg = make_config("solo8", task="stand", control="pd", H=0, episode_length=12, solver_residual_threshold=0.0)
env = SoloVecEnv(cfg, n, device="cuda:0", seed=3)
ors = [OracleEnv(env.model, env.params, seed=3, env_id=i) for i in range(nref)]
obs = env.reset().cpu().numpy()
for i,o in enumerate(ors): o.reset()
fl = flag_slots(ors[0].d0, 8, 1)
for t in range(30):
    a = rng.uniform(-1.2, 1.2, size=(n, env.sim.act_dim)).astype(np.float32)
    ob, rw, dn, infos = env.step(torch.from_numpy(a).cuda())
    ob = ob.cpu().numpy()
    line=[]
    for i,o in enumerate(ors):
        oo, r, d, info = o.step(a[i].astype(np.float64), auto_reset=True)
        df = obs_diff(oo, ob[i], o.d0); df[fl]=0
        rel = df/np.maximum(1,np.abs(oo))
        k=int(rel.argmax())
        line.append("%.1e@%d%s"%(rel.max(),k,'*' if d else ''))
    print(t, ' '.join(line))
