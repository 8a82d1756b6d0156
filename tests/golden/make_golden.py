"""Generate golden vectors from the REFERENCE's own importable code.

Run in the builder container only (needs /root/reference; it does not exist on the GPU box):
    python tests/golden/make_golden.py [/root/reference]
Imports, unmodified:
  controllers/PD.py      -> pd_cases.npz     (PD law)
  agents/ppo/storage.py  -> gae_cases.npz    (OPBuffer.compute_returns)
  agents/ppo/policy.py   -> policy_shapes.json (parameter counts of the reference Policy)
PyBullet / gym are not installed, so no physics vectors can be produced (parity unpinned).
"""
import importlib.util
import json
import os
import sys
import types

import numpy as np
import torch

ref = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
here = os.path.dirname(os.path.abspath(__file__))


def load(path, name):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


# ---- PD law ---------------------------------------------------------------------------
PD = load(os.path.join(ref, "controllers", "PD.py"), "ref_PD").PD
rng = np.random.default_rng(20261018)
n = 64
q_ref = rng.uniform(-12, 12, size=(n, 12))
q = rng.uniform(-3, 3, size=(n, 12))
qd = rng.normal(size=(n, 12)) * 20
kp = rng.uniform(0, 8, size=n)
kd = rng.uniform(0, 0.5, size=n)
out = np.stack([PD(q_ref[i], q[i], qd[i], kp[i], kd[i], 3) for i in range(n)])
# hand-checked vector quoted in SURVEY §8c
assert np.allclose(PD(np.array([1., -1.]), np.zeros(2), np.ones(2), 5, .2, 3), [3, -3])
np.savez(os.path.join(here, "pd_cases.npz"), q_ref=q_ref, q=q, qd=qd, kp=kp, kd=kd, torque_limit=3.0, out=out)

# ---- GAE / returns ----------------------------------------------------------------------
storage = load(os.path.join(ref, "agents", "ppo", "storage.py"), "ref_storage")
cases = {}
for ci, (T, N, gamma, lam) in enumerate(((5, 3, 0.99, 0.95), (63, 17, 0.99, 0.95), (400, 8, 0.995, 0.9), (1, 1, 0.9, 1.0))):
    g = torch.Generator().manual_seed(100 + ci)
    buf = storage.OPBuffer(T, N, (4,), 2, torch.device("cpu"))
    buf.rewards.copy_(torch.randn(T, N, 1, generator=g))
    buf.value_preds.copy_(torch.randn(T + 1, N, 1, generator=g))
    buf.masks.copy_((torch.rand(T + 1, N, 1, generator=g) > 0.1).float())
    next_value = torch.randn(N, 1, generator=g)
    buf.compute_returns(next_value, True, gamma, lam)
    ret_gae = buf.returns.clone()
    buf.compute_returns(next_value, False, gamma, lam)
    ret_disc = buf.returns.clone()
    cases[f"c{ci}_rewards"] = buf.rewards.squeeze(-1).numpy()
    v = buf.value_preds.squeeze(-1).numpy().copy()
    cases[f"c{ci}_values"] = v                     # values[T] = next_value after the GAE call
    cases[f"c{ci}_masks"] = buf.masks.squeeze(-1).numpy()
    cases[f"c{ci}_next_value"] = next_value.squeeze(-1).numpy()
    cases[f"c{ci}_ret_gae"] = ret_gae.squeeze(-1).numpy()
    cases[f"c{ci}_ret_disc"] = ret_disc.squeeze(-1).numpy()
    cases[f"c{ci}_meta"] = np.array([T, N, gamma, lam])
np.savez(os.path.join(here, "gae_cases.npz"), **cases)

# ---- Policy shape -------------------------------------------------------------------------
pkg = types.ModuleType("soloRL"); pkg.__path__ = [ref]
sys.modules["soloRL"] = pkg
agents = types.ModuleType("soloRL.agents"); agents.__path__ = [os.path.join(ref, "agents")]
sys.modules["soloRL.agents"] = agents
gym = types.ModuleType("gym"); sys.modules["gym"] = gym
load(os.path.join(ref, "agents", "utils.py"), "soloRL.agents.utils")
policy = load(os.path.join(ref, "agents", "ppo", "policy.py"), "ref_policy")


class Box:
    def __init__(self, n): self.shape = (n,)


shapes = {}
for obs_dim, act_dim in ((76, 12), (60, 8), (84, 12), (30, 8)):
    torch.manual_seed(1)
    pol = policy.Policy((obs_dim,), Box(act_dim), None, {"hidden_size": 64})
    shapes[f"{obs_dim}x{act_dim}"] = {
        "n_params": int(sum(p.numel() for p in pol.parameters())),
        "state_dict": {k: list(v.shape) for k, v in pol.state_dict().items()},
    }
with open(os.path.join(here, "policy_shapes.json"), "w") as f:
    json.dump(shapes, f, indent=1)
print("golden written:", {k: v["n_params"] for k, v in shapes.items()})

# ---- Policy forward + one PPO update (reference agents/ppo/policy.py, ppo.py, storage.py) --------
# Full-batch mini-batches (mini_batch_size = T*N) make the update independent of the sampler's order.
ref_ppo = load(os.path.join(ref, "agents", "ppo", "ppo.py"), "ref_ppo")
torch.manual_seed(5)
T, N, D, A = 6, 8, 76, 12
pol = policy.Policy((D,), Box(A), None, {"hidden_size": 64})
with torch.no_grad():
    pol.pi_dist.logstd.copy_(torch.randn(A) * 0.3)
init_sd = {k: v.clone() for k, v in pol.state_dict().items()}
gen = torch.Generator().manual_seed(6)
buf = storage.OPBuffer(T, N, (D,), A, torch.device("cpu"))
buf.obs.copy_(torch.randn(T + 1, N, D, generator=gen))
buf.actions.copy_(torch.randn(T, N, A, generator=gen))
with torch.no_grad():
    v, lp, ent = pol.evaluate_actions(buf.obs[:-1].reshape(-1, D), buf.actions.reshape(-1, A))
buf.value_preds[:-1].copy_(v.reshape(T, N, 1) + 0.05 * torch.randn(T, N, 1, generator=gen))
buf.action_log_probs.copy_(lp.reshape(T, N, 1) + 0.05 * torch.randn(T, N, 1, generator=gen))
buf.returns.copy_(torch.randn(T + 1, N, 1, generator=gen))
agent = ref_ppo.PPO(pol, 0.1, 3, T * N, 0.5, 0.01, lr=1e-3, l2_coef=0.0, max_grad_norm=0.5)
losses = agent.update(buf)
blob = {"meta": np.array([T, N, D, A]), "losses": np.array(losses, dtype=np.float64),
        "fwd_value": v.numpy(), "fwd_logp": lp.numpy(), "fwd_entropy": np.array(float(ent)),
        "obs": buf.obs.numpy(), "actions": buf.actions.numpy(), "value_preds": buf.value_preds.numpy(),
        "action_log_probs": buf.action_log_probs.numpy(), "returns": buf.returns.numpy()}
for k, t in init_sd.items():
    blob["init/" + k] = t.numpy()
for k, t in pol.state_dict().items():
    blob["final/" + k] = t.detach().numpy()
np.savez(os.path.join(here, "ppo_update.npz"), **blob)
print("ppo_update golden: losses", losses)
