"""Phase stamps + in-stream time of the chained forward kernel (csrc/grx_mlp_chain.cuh).  Usage: time_chain.py [M] [store]"""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "wiki-grx-gym_b200")]
import torch
from grx_b200 import _lib as L
from grx_b200.config import make_train_cfg
from grx_b200.ppo import PPO, ActorCriticMLP

M = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
tc = make_train_cfg()
torch.manual_seed(1)
ac = ActorCriticMLP(39, 168, 10, **tc["policy"])
alg = PPO(ac, device="cuda:0", **tc["algorithm"])
alg.init_storage(M, 4)
lib = alg.lib
obs, cobs = torch.randn(M, 39, device="cuda"), torch.randn(M, 168, device="cuda")
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for fused in (1, 0):
    lib.grx_ppo_debug_fused(fused)
    for _ in range(5):
        alg.step = 0
        alg.act(obs, cobs)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(50):
        alg.step = 0
        alg.act(obs, cobs)
    e1.record()
    torch.cuda.synchronize()
    print(f"M={M} fused={fused}: PPO.act {e0.elapsed_time(e1) / 50 * 1e3:.1f} us per call (2 staging copies + hidden layers + heads)")
    if fused:
        stamps = (C.c_uint64 * 16)()
        alg.step = 0
        alg.act(obs, cobs)
        L.check(lib.grx_gemm_debug_stamps(stamps))
        st = [int(x) for x in stamps]
        names = ["entry", "setup", "mma:X landed", "mma:L0(0) issued", "mma:first A box", "mma:L1(3) issued", "mma:L2 issued", "epi:acc0[0] full",
                 "epi:chunk0 boxed", "epi:chunk3 boxed", "epi:acc1 full", "epi:acc1 boxed", "epi:acc2 full", "epi:done"]
        print("chain kernel CTA 0, us after entry: " + ", ".join(f"{n} {(x - st[0]) / 1e3:.2f}" for n, x in zip(names[1:], st[1:14])))
