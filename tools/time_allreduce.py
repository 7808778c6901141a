"""2+ ranks (torchrun): phase stamps of the in-kernel NVLink gradient all-reduce and of the apply kernel.  Usage:
python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/time_allreduce.py"""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "wiki-grx-gym_b200")]
import torch, torch.distributed as dist
from grx_b200 import _lib as L
from grx_b200.config import make_train_cfg
from grx_b200.ppo import PPO, ActorCriticMLP

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = f"cuda:{local}"
os.environ.setdefault("NCCL_DEBUG", "WARN")
dist.init_process_group("nccl", device_id=torch.device(dev))
tc = make_train_cfg()
torch.manual_seed(1)
N, T = 4096, 64
ac = ActorCriticMLP(39, 168, 10, **tc["policy"])
alg = PPO(ac, device=dev, world_size=world, env_id_offset=rank * N, **tc["algorithm"])
alg.init_storage(N, T)
obs, cobs = torch.randn(N, 39, device=dev), torch.randn(N, 168, device=dev)
for s in range(T):
    alg.act(obs, cobs)
    alg.process_env_step(torch.randn(N, device=dev) * 0.1, torch.rand(N, device=dev) < 0.01, {})
alg.compute_returns(cobs)
alg.draw_indices()
lib = alg.lib
idx = C.c_void_p(alg._indices.data_ptr())
for mb in range(6):
    L.check(lib.grx_ppo_minibatch_grads(alg._h, idx, mb, alg._stream()))
    dist.barrier()
    torch.cuda.synchronize()
    L.check(lib.grx_ppo_minibatch_apply_comm(alg._h, alg._stream()))
    torch.cuda.synchronize()
st = (C.c_uint64 * 32)()
L.check(lib.grx_debug_stamps32(st))
s = [int(x) for x in st]
t0 = s[16]
fmt = lambda xs: " ".join(f"{(x - t0) / 1e3:7.2f}" for x in xs)
print(f"rank {rank}: phase 0 [entry ready reduced done] {fmt(s[16:20])} | phase 1 {fmt(s[20:24])} | apply [entry waited barrier scalars adam] {fmt(s[0:5])}  (us after phase-0 entry; ranks were aligned by a host barrier)", flush=True)
# graph update timing
for rep in range(2):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    dist.barrier(); torch.cuda.synchronize()
    e0.record(); alg.update(); e1.record(); torch.cuda.synchronize()
print(f"rank {rank}: graph update {e0.elapsed_time(e1) / 200 * 1e3:.1f} us per minibatch", flush=True)
alg.check_comm(wait=True)
dist.destroy_process_group()
