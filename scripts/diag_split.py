import sys, numpy as np, torch
sys.path.insert(0, '.')
from babyai_b200 import BabyAIVecEnv
level, n = sys.argv[1], int(sys.argv[2])
env = BabyAIVecEnv(level, n, seeds=np.arange(n, dtype=np.uint64) + 100)
env.reset()
acts = torch.randint(0, 7, (64, n), device='cuda', dtype=torch.int8)
for t in range(200): env.step(acts[t % 64])
ks, kg = [], []
for t in range(60):
    a, b = env.step_timed(acts[t % 64])
    ks.append(a); kg.append(b)
print(level, n, 'k_step %.1f us  k_gen(one step worth, isolated) %.1f us' % (1e3 * np.mean(ks[10:]), 1e3 * np.mean(kg[10:])), env.counters())
