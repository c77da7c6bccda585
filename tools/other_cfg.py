"""cfg3 / cfg4 forward throughput (bench.other_configs) printed compactly."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench

d = bench.other_configs(torch.device("cuda", 0))
for k, v in d.items():
    print(k, {kk: (round(vv, 3) if isinstance(vv, float) else vv) for kk, vv in v.items() if kk in ("ms_per_step", "frames_per_s", "error")})
