import sys, torch
sys.path.insert(0,'/root/repo')
from solorl_b200.agents.policy import Policy
from solorl_b200.agents.storage import OPBuffer
from solorl_b200.agents.train import EpisodeTracker, Rollout
from solorl_b200.envs import make_vec_envs
cfg = {"model_urdf": "solo12", "mode": "headless", "episode_length": 400, "frame_skip": 4, "control": "torque", "task": "walk", "num_history_stack": 1}
n, T = 4096, 8
envs = make_vec_envs(cfg, n, seed=2)
ac = Policy(envs.observation_space.shape, envs.action_space, None, {"hidden_size": 64}).cuda()
buf = OPBuffer(T, n, envs.observation_space.shape, 12, "cuda")
buf.obs[0].copy_(envs.reset())
ro = Rollout(envs, ac, buf, EpisodeTracker(torch.device("cuda")), T, use_graph=False)
ro(); torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    ro(); torch.cuda.synchronize()
ev = [e for e in prof.key_averages() if e.device_type is not None]
rows = sorted(((e.key, e.count, e.device_time_total) for e in prof.key_averages()), key=lambda r: -r[2])
tot = sum(r[2] for r in rows); cnt = sum(r[1] for r in rows)
print("kernels per step %.1f, device us per step %.1f" % (cnt / T, tot / T))
for k, c, t in rows[:22]:
    print("%5.1f/step %7.1f us/step  %s" % (c / T, t / T, k[:90]))
