import sys, os
sys.path.insert(0, "/root/repo")
import torch
from tools.gemm_sweep import bench
impl = int(sys.argv[1])
print(bench(4608, 12288, 3072, impl, iters=5))
