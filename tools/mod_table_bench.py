#!/usr/bin/env python
"""Time of rt_model_build_modulation_table on the full-size models (run under gpurun)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from reptext_b200 import config, models, _lib

dt, dev = torch.bfloat16, "cuda"
tr = models.FluxTransformer2DModel.random_init(config.FLUX_DEV, seed=100, dtype=dt, device=dev)
cn = models.FluxControlNetModel.random_init(config.REPTEXT_CONTROLNET, seed=101, dtype=dt, device=dev)
po = torch.randn(2, 768, device=dev).to(dt)
g = torch.tensor([3.5], device=dev)
for v in (0,):
  for K, B in ((28, 1), (28, 2)):
      ts = torch.linspace(1.0, 0.03, K, device=dev).to(dt)[:, None]
      for name, net in (("transformer", tr), ("controlnet", cn)):
          for rep in range(3):
              torch.cuda.synchronize()
              t0 = time.perf_counter()
              e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
              e0.record()
              n0 = _lib.launch_count()
              net.build_modulation_table(ts, g, po[:B])
              e1.record()
              torch.cuda.synchronize()
              print(f"{name} steps={K} batch={B} rep={rep}: device {e0.elapsed_time(e1):8.2f} ms, wall {(time.perf_counter() - t0) * 1e3:8.2f} ms, "
                    f"{_lib.launch_count() - n0} launches", flush=True)
