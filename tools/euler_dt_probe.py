#!/usr/bin/env python
"""What does torch do in ``sample.float() + dt * model_output`` (FlowMatchEulerDiscreteScheduler.step) when
``model_output`` is bf16 on CUDA and ``dt`` is a 0-dim fp32 tensor (a) on the DEVICE (diffusers 0.36 keeps
``scheduler.sigmas`` on the device) or (b) on the CPU (EulerDiscrete-style ``sigmas.to("cpu")``)?  Candidates:
  A  bf16( f32(x) + f32( bf16( f32(bf16(dt)) * f32(v) ) ) )      dt rounded to bf16, product rounded to bf16
  B  bf16( f32(x) + f32( bf16( dt * f32(v) ) ) )                 dt kept fp32, product rounded to bf16
  C  bf16( f32(x) + dt * f32(v) )                                nothing rounded before the add
"""
import torch

torch.manual_seed(0)
x = torch.randn(1, 4096, 64, device="cuda").to(torch.bfloat16)
v = torch.randn(1, 4096, 64, device="cuda").to(torch.bfloat16)
sig = torch.tensor([0.9271, 0.8913], dtype=torch.float32)
for where in ("cuda", "cpu"):
    s = sig.to(where)
    dt = s[1] - s[0]
    got = (x.to(torch.float32) + dt * v).to(torch.bfloat16)
    prod = dt * v
    d = float(dt)
    dbf = float(torch.tensor(d).to(torch.bfloat16))
    A = (x.float() + (dbf * v.float()).to(torch.bfloat16).float()).to(torch.bfloat16)
    Bc = (x.float() + (d * v.float()).to(torch.bfloat16).float()).to(torch.bfloat16)
    C = (x.float() + d * v.float()).to(torch.bfloat16)
    print(f"dt on {where}: product dtype {prod.dtype}; equals A {torch.equal(got, A)}  B {torch.equal(got, Bc)}  "
          f"C {torch.equal(got, C)}; mismatches A/B/C {(got != A).sum().item()}/{(got != Bc).sum().item()}/{(got != C).sum().item()}")
