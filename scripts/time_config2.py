import sys, torch
sys.path.insert(0, '.')
import bench
r = bench.secondary_config2(torch.device('cuda', 0))
print({k: (round(v['loop_ms'], 4), round(v['setup_ms'], 1), round(v['admm_iterations_mean'], 4)) for k, v in r['variants'].items()})
