import sys, numpy as np, torch
sys.path.insert(0, '.')
from babyai_b200 import BabyAIVecEnv
n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
seeds = np.array([100 + i for i in range(n)], dtype=np.uint64)
a = BabyAIVecEnv('GoToLocal', n, seeds=seeds)
b = BabyAIVecEnv('GoToLocal', n, seeds=seeds)
oa = a.reset().clone(); ob = b.reset().clone()
torch.cuda.synchronize()
bad = (oa != ob).reshape(n, -1).any(1).nonzero().flatten().cpu().numpy()
print('n', n, 'differing envs', len(bad), bad[:40])
if len(bad):
    print('mod 2368:', sorted(set((bad % 2368).tolist()))[:40])
    print('div 2368:', sorted(set((bad // 2368).tolist()))[:40])
    for i in bad[:5]:
        print(i, a.state(int(i))[1], b.state(int(i))[1])
        print(a.missions([int(i)]), b.missions([int(i)]))
# third pool to see which one is "right" vs oracle
sys.path.insert(0, 'oracle')
import oracle as orc
o = orc.OraclePool('GoToLocal', n, seeds)
oo = torch.as_tensor(o.reset())
ba = (oa.cpu() != oo).reshape(n, -1).any(1).nonzero().flatten().numpy()
bb = (ob.cpu() != oo).reshape(n, -1).any(1).nonzero().flatten().numpy()
print('a vs oracle', len(ba), ba[:20], 'b vs oracle', len(bb), bb[:20])
