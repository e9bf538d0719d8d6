import sys, time, numpy as np, torch
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ranslice_b200 import create_batched_env
N = 16384
env = create_batched_env(1, 0, N, L1_level=False); env.reset()
rng = np.random.default_rng(0)
acts = [torch.from_numpy(rng.integers(60, 200, (N, 1)).astype(np.int32)).cuda() for _ in range(8)]
out = None
for i in range(200): out = env.step_device(acts[i % 8], out)
torch.cuda.synchronize(); t0 = time.perf_counter()
for i in range(20): out = env.step_device(acts[i % 8], out)
torch.cuda.synchronize(); dt = time.perf_counter() - t0
print("L1_level=False scenario_0: %d envs, %.3f ms/step, %.3f M env-steps/s, mean UEs per env %.1f, flagged envs %d" % (N, 1e3 * dt / 20, N * 20 / dt / 1e6, env.n_ues().mean(), int((out["flags"].cpu().numpy() != 0).sum())))
