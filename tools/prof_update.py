"""Small driver for ncu: a rollout of T steps at N envs with random observations, then `nmb` minibatches of the PPO update
(explicit grads/apply, no graph) at the registered sizes.  Usage: prof_update.py [N] [T] [nmb] [use_tc]"""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "wiki-grx-gym_b200")]
import torch
from grx_b200 import _lib as L
from grx_b200.config import make_train_cfg
from grx_b200.ppo import PPO, ActorCriticMLP

N = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
T = int(sys.argv[2]) if len(sys.argv) > 2 else 64
nmb = int(sys.argv[3]) if len(sys.argv) > 3 else 3
use_tc = int(sys.argv[4]) if len(sys.argv) > 4 else 1
tc = make_train_cfg()
torch.manual_seed(1)
O, P, A = (int(x) for x in os.environ.get("GRX_PROF_DIMS", "39,168,10").split(","))   # e.g. 105,234,32 = the full-body task
ac = ActorCriticMLP(O, P, A, **tc["policy"])
alg = PPO(ac, device="cuda:0", use_tensor_cores=use_tc, **tc["algorithm"])
alg.init_storage(N, T)
obs, cobs = torch.randn(N, O, device="cuda"), torch.randn(N, P, device="cuda")
for s in range(T):
    alg.act(obs, cobs)
    alg.process_env_step(torch.randn(N, device="cuda") * 0.1, torch.rand(N, device="cuda") < 0.01, {})
alg.compute_returns(cobs)
alg.draw_indices()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for rep in range(2):
    e0.record()
    for mb in range(nmb):
        L.check(alg.lib.grx_ppo_minibatch_grads(alg._h, C.c_void_p(alg._indices.data_ptr()), mb, alg._stream()))
        L.check(alg.lib.grx_ppo_minibatch_apply(alg._h, alg._stream()))
    e1.record()
    torch.cuda.synchronize()
if os.environ.get("GRX_PPO_TIMING"):
    out, cnt = (C.c_float * 7)(), C.c_int32()
    L.check(alg.lib.grx_ppo_debug_timing(alg._h, out, C.byref(cnt)))
    names = ["memset+gather", "actor fwd", "critic fwd", "heads", "actor bwd", "critic bwd", "apply"]
    print("phase us/minibatch (stepwise, events):", {n: round(v, 1) for n, v in zip(names, out)}, "over", cnt.value)
print(f"N={N} T={T} use_tc={use_tc}: {e0.elapsed_time(e1) / nmb * 1e3:.1f} us per minibatch (B={alg.mini_batch_size})", alg.minibatch_stats())
stamps = (C.c_uint64 * 16)()
L.check(alg.lib.grx_gemm_debug_stamps(stamps))
st = [int(x) for x in stamps]
print("apply_kernel block 0 (us after entry): reduced %s  barrier %s  scalars %s  adam %s" % tuple(round((x - st[0]) / 1e3, 2) for x in st[1:5]))
# whole update through the CUDA graph (what the runner uses): no per-launch CPU cost
alg.update()
torch.cuda.synchronize()
e0.record()
alg.update()
e1.record()
torch.cuda.synchronize()
nmb_total = alg.num_learning_epochs * alg.num_mini_batches
print(f"graph replay: {e0.elapsed_time(e1) / nmb_total * 1e3:.1f} us per minibatch over {nmb_total} minibatches")
