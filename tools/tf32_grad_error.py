"""How much of the TF32-path gradient error is inherent to TF32?  CPU emulation: one PPO minibatch at the registered width with every GEMM
operand rounded to a 10-bit mantissa (round-to-nearest-even, what tcgen05 kind::tf32 consumes), fp32 accumulate — vs the same minibatch
in plain fp32.  Prints the norm-wise relative error per gradient tensor (tests/test_ppo_gpu.py::_wide_check_final cites the result)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "wiki-grx-gym_b200"), os.path.join(ROOT, "tests")]
import numpy as np, torch
from oracle import ppo_oracle as po


def tf32(x):
    i = x.contiguous().view(torch.int32)
    r = (i + 0x0FFF + ((i >> 13) & 1)) & ~0x1FFF      # round to nearest even on the 13 dropped bits
    return r.view(torch.float32)


class TF32MM(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a, b):
        ctx.save_for_backward(a, b)
        return tf32(a) @ tf32(b)

    @staticmethod
    def backward(ctx, g):
        a, b = ctx.saved_tensors
        return tf32(g) @ tf32(b).t(), tf32(a).t() @ tf32(g)


def run(p, b, emulate):
    pa = {k: v.clone().requires_grad_(True) for k, v in p.items()}
    mm = TF32MM.apply if emulate else (lambda x, w: x @ w)

    def mlp(net, x, head_fp32=True):
        for i in range(4):
            w, bias = pa[f"{net}.model.{2 * i}.weight"], pa[f"{net}.model.{2 * i}.bias"]
            x = (mm(x, w.t()) if i < 3 else x @ w.t()) + bias          # the output heads run in fp32 (fused heads kernels)
            if i < 3:
                x = torch.nn.functional.elu(x)
        return x
    mu, v = mlp("actor", b["obs"]), mlp("critic", b["critic_obs"])
    dist = torch.distributions.Normal(mu, mu * 0.0 + pa["std"])
    lp = dist.log_prob(b["actions"]).sum(-1)
    ratio = torch.exp(lp - b["old_log_prob"].squeeze(-1)); A_ = b["advantages"].squeeze(-1)
    sl = torch.max(-A_ * ratio, -A_ * torch.clamp(ratio, 0.8, 1.2)).mean()
    vc = b["values"] + (v - b["values"]).clamp(-0.2, 0.2)
    vl = torch.max((v - b["returns"]).pow(2), (vc - b["returns"]).pow(2)).mean()
    (sl + vl - 0.01 * dist.entropy().sum(-1).mean()).backward()
    return {k: t.grad for k, t in pa.items()}


torch.manual_seed(0)
O, P, A, M = 39, 168, 10, 1024
p = po.init_params(O, P, A)
b = dict(obs=torch.randn(M, O), critic_obs=torch.randn(M, P), values=torch.randn(M, 1) * 0.1, advantages=torch.randn(M, 1), returns=torch.randn(M, 1) * 0.2,
         old_mu=torch.zeros(M, A), old_sigma=torch.full((M, A), 0.2))
mu0 = po.mlp_forward(p, "actor", b["obs"])
b["actions"] = mu0 + 0.2 * torch.randn(M, A)
b["old_log_prob"] = po.log_prob(b["actions"], mu0 + 0.01 * torch.randn(M, A), b["old_sigma"]).unsqueeze(1)
g32, gtf = run(p, b, False), run(p, b, True)
for k in g32:
    e = float((gtf[k] - g32[k]).norm() / g32[k].norm())
    emax = float((gtf[k] - g32[k]).abs().max() / g32[k].abs().max())
    print(f"{k:28s} norm-wise rel err {e:.4f}   max err / max |g| {emax:.4f}")
