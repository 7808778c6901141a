"""In-stream time of one PPO.act call (staging, hidden layers, heads + sampling + storage) with the layer-pipelined launches off / on.
Usage: time_act.py [M]"""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "wiki-grx-gym_b200")]
import torch
from grx_b200.config import make_train_cfg
from grx_b200.ppo import PPO, ActorCriticMLP

M = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
tc = make_train_cfg()
torch.manual_seed(1)
ac = ActorCriticMLP(39, 168, 10, **tc["policy"])
alg = PPO(ac, device="cuda:0", **tc["algorithm"])
alg.init_storage(M, 4)
lib = alg.lib
lib.grx_ppo_debug_pipe.argtypes = [C.c_int32]
obs, cobs = torch.randn(M, 39, device="cuda"), torch.randn(M, 168, device="cuda")
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for pipe in (0, 1, 0, 1):
    lib.grx_ppo_debug_pipe(pipe)
    for _ in range(10):
        alg.step = 0
        alg.act(obs, cobs)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(200):
        alg.step = 0
        alg.act(obs, cobs)
    e1.record()
    torch.cuda.synchronize()
    print(f"M={M} forward chain pipelined={pipe}: PPO.act {e0.elapsed_time(e1) / 200 * 1e3:.1f} us per call")
