"""Trajectory effect of TF32 on a whole PPO.update at the registered width: the wide rsl_rl fixture (tests/golden/ppo_wide.npz: 8 optimiser
steps) replayed on the CPU with every hidden-layer GEMM operand rounded to TF32 (fp32 accumulate), compared with the fp32 golden result —
the same two numbers tests/test_ppo_gpu.py::_wide_check_final reports for the CUDA tensor-core path."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "wiki-grx-gym_b200"), os.path.join(ROOT, "tests")]
import numpy as np, torch
from golden_util import wide_inputs
from grx_b200.config import make_train_cfg
from oracle import ppo_autograd as pa

MODE = sys.argv[1] if len(sys.argv) > 1 else "round"      # round | trunc | fp32


def tf32(x):
    i = x.contiguous().view(torch.int32)
    if MODE == "trunc":
        return (i & ~0x1FFF).view(torch.float32)
    return ((i + 0x0FFF + ((i >> 13) & 1)) & ~0x1FFF).view(torch.float32)


class TF32Linear(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, b):
        ctx.save_for_backward(x, w)
        return tf32(x) @ tf32(w).t() + b

    @staticmethod
    def backward(ctx, g):
        x, w = ctx.saved_tensors
        return tf32(g) @ tf32(w), tf32(g).t() @ tf32(x), g.sum(0)


class Lin(torch.nn.Linear):
    def forward(self, x):
        return TF32Linear.apply(x, self.weight, self.bias)


fx = dict(np.load(os.path.join(ROOT, "tests", "golden", "ppo_wide.npz")))
N, T, nmb, nep, O, P, A = [int(v) for v in fx["meta/dims"]]
d, _ = wide_inputs(int(fx["meta/seed"]), N, T, O, P, A)
tc = make_train_cfg()["algorithm"]
torch.manual_seed(int(fx["meta/seed"]))
ac = pa.ActorCritic(O, P, A)
init = {k: v.detach().clone() for k, v in ac.state_dict().items()}
if MODE != "fp32":
    for net in (ac.actor, ac.critic):
        for i in (0, 2, 4):                                   # hidden layers on the tensor cores; the output heads stay fp32 (fused heads kernels)
            old = net.model[i]
            new = Lin(old.in_features, old.out_features)
            new.weight, new.bias = old.weight, old.bias
            net.model[i] = new
step = pa.PPOStep(ac, clip=tc["clip_param"], vcoef=tc["value_loss_coef"], ecoef=tc["entropy_coef"], lr=float(fx["meta/lr0"]), lr_min=tc["learning_rate_min"],
                  lr_max=tc["learning_rate_max"], desired_kl=tc["desired_kl"], max_grad_norm=tc["max_grad_norm"])
# rollout quantities from the fp32 golden (the comparison isolates the update)
with torch.no_grad():
    mu = torch.stack([ac.actor(d["obs"][t]) for t in range(T)])
act = mu + 0.2 * d["eps"]
flat = lambda x: x.flatten(0, 1)
st = dict(obs=flat(d["obs"]), critic_obs=flat(d["critic_obs"]), actions=flat(act), values=flat(torch.from_numpy(fx["storage/values"])),
          advantages=flat(torch.from_numpy(fx["storage/advantages"])), returns=flat(torch.from_numpy(fx["storage/returns"])),
          old_log_prob=flat(torch.from_numpy(fx["storage/actions_log_prob"])), old_mu=flat(mu), old_sigma=torch.full((N * T, A), 0.2))
B = (N * T) // nmb
for ep in range(nep):
    for mb in range(nmb):
        sel = d["indices"][mb * B:(mb + 1) * B]
        step.minibatch({k: v[sel] for k, v in st.items()})
print("mode", MODE, "kl/lr", [(round(k, 5), round(l, 7)) for k, l in step.kl_log])
worst, worst_m = 0.0, 0.0
names = [n for n, _ in ac.named_parameters()]
osd = step.opt.state_dict()["state"]
for i, n in enumerate(names):
    w0, ref, ours = init[n].numpy(), fx["final/" + n], ac.state_dict()[n].detach().numpy()
    rel = np.linalg.norm((ours - w0) - (ref - w0)) / np.linalg.norm(ref - w0)
    m8, rm = osd[i]["exp_avg"].flatten()[::8].numpy(), fx["adam_m8/" + n]
    em = np.linalg.norm(m8 - rm) / np.linalg.norm(rm)
    if n != "std":
        worst, worst_m = max(worst, rel), max(worst_m, em)
print(f"worst per-tensor relative update error {worst:.4f}; worst Adam first-moment relative error {worst_m:.4f}")
