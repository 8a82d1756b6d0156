"""Error statistics of one contact substep against the oracle on stance states (both robots)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from tests.helpers import make_config, stance_states
from solorl_b200.abi import params_from_config
from solorl_b200.model import SoloModel
from solorl_b200.sim import SoloSim
from oracle.oracle import OracleEnv

for robot in ("solo8", "solo12"):
    for seed in (24, 25, 26):
        rng = np.random.default_rng(seed)
        n = 128
        m = SoloModel.resolve(robot); p = params_from_config(make_config(robot), m)
        sim = SoloSim(m, p, n, device=0)
        nj = sim.nj
        cur = stance_states(rng, n, nj)
        target = cur[:, 13:13 + nj].copy()
        o = OracleEnv(m, p)
        errs = []; its = []
        for t in range(40):
            tau = np.clip(3.0 * (target - cur[:, 13:13 + nj]) - 0.05 * cur[:, 13 + nj:], -3, 3) + rng.normal(size=(n, nj)) * 0.3
            tau = tau.astype(np.float32).astype(np.float64)
            sim.set_state(torch.from_numpy(cur.astype(np.float32)).cuda())
            sim.substep(torch.from_numpy(tau.astype(np.float32)).cuda())
            nxt = sim.get_state().cpu().numpy().astype(np.float64)
            for i in range(0, n, 2):
                o.set_state(cur[i]); o.substep(tau[i])
                ref = o.get_state()
                errs.append((np.abs(ref - nxt[i]) / np.maximum(1.0, np.abs(ref))).max()); its.append(o.last_solver_iters)
            cur = nxt
        errs = np.array(errs); its = np.array(its)
        print(f"{robot} seed {seed}: max {errs.max():.3e} p99 {np.percentile(errs,99):.3e} median {np.median(errs):.3e} "
              f"frac>1e-4 {(errs>1e-4).mean():.4f}; at-cap frac {(its>=50).mean():.3f}, max err among converged {errs[its<50].max():.3e}")
        sim.close()
