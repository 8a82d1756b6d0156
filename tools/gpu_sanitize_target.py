"""Small workload touching every kernel and every path of the step (ragged batch, resets, limit rows, host buffers,
actuator, feet, GAE, accumulate) (written as a compute-sanitizer target; the sanitizer is closed on this GPU pool, so the out-of-bounds check that
runs is tests/test_gpu_parity.py::test_outputs_stay_inside_their_buffers, with canaries around every output)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from solorl_b200.envs import SoloVecEnv
from solorl_b200.gait import ActuatorSim
from solorl_b200.sim import gae

for variant in ("latency", "throughput"):
    os.environ["SOLO_STEP_VARIANT"] = variant
    for robot, task, ctl, H in (("solo12", "pointgoal", "torque", 2), ("solo8", "stand", "vpd", 0)):
        cfg = {"model_urdf": robot, "mode": "headless", "episode_length": 5, "frame_skip": 4, "control": ctl,
               "task": task, "num_history_stack": H}
        n = 37
        env = SoloVecEnv(cfg, n, device="cuda:0", seed=1)
        env.reset()
        s = env.sim.get_state()
        s[::3, 13] = 10.3                                    # joint-limit rows in some envs
        env.sim.set_state(s)
        acc = torch.tensor([0.0] * 10 + [float("inf"), float("-inf"), 0.0], dtype=torch.float64, device="cuda")
        for t in range(8):
            a = torch.rand(n, env.sim.act_dim, device="cuda") * 2 - 1
            o, r, d, info = env.step(a)
            env.sim.accumulate_episode_stats(d, acc)
        ho = torch.empty(n, env.sim.d).pin_memory(); hr = torch.empty(n).pin_memory(); hd = torch.empty(n).pin_memory()
        ha = torch.rand(n, env.sim.act_dim).pin_memory()
        env.sim.step_host(ha.numpy(), ho.numpy(), hr.numpy(), hd.numpy())
        env.sim.step_host(np.random.rand(n, env.sim.act_dim).astype(np.float32), np.empty((n, env.sim.d), np.float32),
                          np.empty(n, np.float32), np.empty(n, np.float32))
        env.sim.reset(torch.arange(n, device="cuda") % 2 == 0)
        env.get_observation(); env.sim.get_contacts(); env.sim.get_work_counters(); env.sim.get_feet()
        env.sim.forward_dynamics(env.sim.get_state(), torch.zeros(n, env.sim.nj, device="cuda"))
        env.sim.action_to_torque(torch.zeros(n, env.sim.act_dim, device="cuda"))
        env.sim.substep(torch.zeros(n, env.sim.nj, device="cuda"))
        if task == "pointgoal":
            env.sim.set_goals(torch.ones(n, 2, device="cuda"))
        env.close()
rob = ActuatorSim(13, solo12=True)
rob.sim.actuator_step(torch.zeros(13, 5, 12, device="cuda"), 3)
T, N = 7, 33
ret = torch.zeros(T + 1, N, device="cuda")
gae(torch.randn(T, N, device="cuda"), torch.randn(T + 1, N, device="cuda"), torch.ones(T + 1, N, device="cuda"), ret, 0.99, 0.95)
torch.cuda.synchronize()
print("sanitize target done")
