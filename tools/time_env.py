"""Quick device timing of the fused env step (CUDA events on the launching stream)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "wiki-grx-gym_b200")]
import torch
from grx_b200.config import make_cfg
from grx_b200.env import GRXVecEnv

N = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 200
cfg = make_cfg("GR1T1", N, sys.argv[3] if len(sys.argv) > 3 else "heightfield")
env = GRXVecEnv(cfg)
env.reset()
a = 0.3 * torch.randn(steps, N, 10, device="cuda")
for i in range(20):
    env.step(a[i])
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(steps):
    env.step(a[i])
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / steps
print(f"N={N} {ms*1e3:.1f} us/step  {N/ms*1e3:.3e} env-steps/s  resets/step={float(env.reset_buf.float().mean())*N:.1f}")
