"""Find the worst single-substep sample against the oracle and print its anatomy."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from tests.helpers import make_config, stance_states
from solorl_b200.abi import params_from_config
from solorl_b200.model import SoloModel
from solorl_b200.sim import SoloSim
from oracle.oracle import OracleEnv

robot, seed = sys.argv[1], int(sys.argv[2])
rng = np.random.default_rng(seed)
n = 128
m = SoloModel.resolve(robot); p = params_from_config(make_config(robot), m)
sim = SoloSim(m, p, n, device=0)
nj = sim.nj
cur = stance_states(rng, n, nj)
target = cur[:, 13:13 + nj].copy()
o = OracleEnv(m, p)
worst = (0, None)
for t in range(40):
    tau = np.clip(3.0 * (target - cur[:, 13:13 + nj]) - 0.05 * cur[:, 13 + nj:], -3, 3) + rng.normal(size=(n, nj)) * 0.3
    tau = tau.astype(np.float32).astype(np.float64)
    sim.set_state(torch.from_numpy(cur.astype(np.float32)).cuda())
    sim.substep(torch.from_numpy(tau.astype(np.float32)).cuda())
    nxt = sim.get_state().cpu().numpy().astype(np.float64)
    con = sim.get_contacts().cpu().numpy()
    work = sim.get_work_counters().cpu().numpy()
    for i in range(0, n, 2):
        o.set_state(cur[i]); o.substep(tau[i])
        ref = o.get_state()
        e = (np.abs(ref - nxt[i]) / np.maximum(1.0, np.abs(ref)))
        if e.max() > worst[0]:
            worst = (e.max(), dict(t=t, i=i, err=e.copy(), ref=ref.copy(), got=nxt[i].copy(), oc=o.get_contacts().copy(),
                                   gc=con[i].copy(), its=o.last_solver_iters, work=work[i].copy(), s0=cur[i].copy(), tau=tau[i].copy(),
                                   feet=o.foot_positions().copy()))
    cur = nxt
w = worst[1]
np.set_printoptions(precision=6, suppress=False, linewidth=200)
print("worst", worst[0], "t", w["t"], "env", w["i"], "oracle iters", w["its"], "gpu work (nc, nc*sweeps)", w["work"])
print("err by slot", np.argsort(-w["err"])[:6], w["err"][np.argsort(-w["err"])[:6]])
print("ref ", w["ref"]); print("got ", w["got"])
print("oracle contacts (flag, has, force)\n", w["oc"]); print("gpu contacts\n", w["gc"])
print("foot z after step (oracle)", w["feet"][:, 2] - m.foot_radius)
np.save("gpurun_out/worst_state.npy", np.concatenate([w["s0"], w["tau"]]))
