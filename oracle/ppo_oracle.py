"""TEST INFRASTRUCTURE — CPU restatement (torch fp32, explicit forward AND hand-derived backward, no
autograd) of the reference's rsl_rl PPO for the ActorCriticMLP policy.  Checker for the CUDA PPO kernels
and the `port` CPU baseline.  Pinned against the reference itself by tests/test_ppo_oracle.py with
tests/golden/ppo_*.npz produced by the UNMODIFIED rsl_rl (oracle/ref_harness/gen_ppo_golden.py).

References (under /root/reference/rsl_rl/rsl_rl/):
  MLP  modules/mlp.py:7-42                 ACM  modules/actor_critic_mlp.py:10-231
  BS   storage/base_storage.py:120-141     RS   storage/rollout_storage.py:63-112
  PPO  algorithms/ppo.py:144-321
Parameter dict keys follow the reference state_dict: 'std', 'actor.model.{0,2,4,6}.{weight,bias}',
'critic.model.{0,2,4,6}.{weight,bias}'.
"""
from __future__ import annotations

import math

import torch

LOG_SQRT_2PI = math.log(math.sqrt(2 * math.pi))


def layer_keys(net, n_layers=4):
    return [(f"{net}.model.{2 * i}.weight", f"{net}.model.{2 * i}.bias") for i in range(n_layers)]


def init_params(num_obs, num_pri_obs, num_actions, actor_hidden=(512, 256, 128), critic_hidden=(512, 256, 128),
                init_noise_std=0.2, generator=None):
    """torch.nn.Linear default init (kaiming_uniform(a=sqrt 5) == U(+-1/sqrt(fan_in)) for weight and bias), MLP:26-31."""
    p = {"std": init_noise_std * torch.ones(num_actions)}                               # ACM:79-82
    for net, dims in (("actor", [num_obs, *actor_hidden, num_actions]), ("critic", [num_pri_obs, *critic_hidden, 1])):
        for i, (wk, bk) in enumerate(layer_keys(net, len(dims) - 1)):
            bound = 1.0 / math.sqrt(dims[i])
            p[wk] = (torch.rand(dims[i + 1], dims[i], generator=generator) * 2 - 1) * bound
            p[bk] = (torch.rand(dims[i + 1], generator=generator) * 2 - 1) * bound
    return p


def elu(x):
    return torch.where(x > 0, x, torch.expm1(x))                                        # nn.ELU(alpha=1), utils.py:240-241


def mlp_forward(p, net, x, keep=False):
    """MLP:40-41.  Returns output (and the hidden activations when keep=True)."""
    keys = layer_keys(net, sum(1 for k in p if k.startswith(net + ".") and k.endswith("weight")))
    hs = [x]
    for i, (wk, bk) in enumerate(keys):
        x = x @ p[wk].t() + p[bk]
        if i < len(keys) - 1:
            x = elu(x)
            hs.append(x)
    return (x, hs) if keep else x


def log_prob(actions, mu, sigma):                                                       # ACM:196-207 (Normal.log_prob summed)
    var = sigma ** 2
    return (-((actions - mu) ** 2) / (2 * var) - torch.log(sigma) - LOG_SQRT_2PI).sum(dim=-1)


def entropy(sigma_row, n_rows):                                                         # ACM:160-163
    return (0.5 + 0.5 * math.log(2 * math.pi) + torch.log(sigma_row)).sum().expand(n_rows)


def act(p, obs, critic_obs, eps):
    """PPO.act (PPO:144-171) with the Normal sample written as mu + sigma * eps."""
    mu = mlp_forward(p, "actor", obs)
    sigma = mu * 0.0 + p["std"]                                                         # ACM:179-181
    actions = mu + sigma * eps
    values = mlp_forward(p, "critic", critic_obs)
    return dict(actions=actions, values=values, actions_log_prob=log_prob(actions, mu, sigma), action_mean=mu,
                action_sigma=sigma)


def process_rewards(rewards, values, time_outs, gamma):                                 # PPO:186-191
    return rewards + gamma * torch.squeeze(values * time_outs.unsqueeze(1).to(values.dtype), 1)


def compute_returns(rewards, dones, values, last_values, gamma, lam):
    """BS:120-141.  rewards/values [T,N,1], dones [T,N,1] uint8/bool.  Returns (returns, normalised advantages)."""
    T = rewards.shape[0]
    returns = torch.zeros_like(rewards)
    adv = 0
    for step in reversed(range(T)):
        nv = last_values if step == T - 1 else values[step + 1]
        nt = 1.0 - dones[step].float()
        delta = rewards[step] + nt * gamma * nv - values[step]
        adv = delta + nt * gamma * lam * adv
        returns[step] = adv + values[step]
    a = returns - values
    return returns, (a - a.mean()) / (a.std() + 1e-8)


def minibatch_loss_and_grads(p, b, clip_param=0.2, value_loss_coef=1.0, entropy_coef=0.01, use_clipped_value_loss=True):
    """One minibatch of PPO.update (PPO:244-295): forward, losses, KL, and the analytic gradient of
    loss = surrogate + c_v * value_loss - c_e * entropy.mean() w.r.t. every parameter.
    b: dict(obs, critic_obs, actions, values, advantages, returns, old_log_prob, old_mu, old_sigma)."""
    M = b["obs"].shape[0]
    mu, ha = mlp_forward(p, "actor", b["obs"], keep=True)
    v, hc = mlp_forward(p, "critic", b["critic_obs"], keep=True)
    std = p["std"]
    sigma = mu * 0.0 + std
    lp = log_prob(b["actions"], mu, sigma)
    ent = entropy(std, M)
    osig, omu = b["old_sigma"], b["old_mu"]
    kl = torch.sum(torch.log(sigma / osig + 1.0e-5) + (osig ** 2 + (omu - mu) ** 2) / (2.0 * sigma ** 2) - 0.5, dim=-1)  # PPO:257-261
    kl_mean = kl.mean()
    A = b["advantages"].squeeze(-1)
    ratio = torch.exp(lp - b["old_log_prob"].squeeze(-1))
    s1 = -A * ratio
    s2 = -A * torch.clamp(ratio, 1.0 - clip_param, 1.0 + clip_param)
    surrogate_loss = torch.max(s1, s2).mean()                                           # PPO:271-277
    R, V0 = b["returns"], b["values"]
    if use_clipped_value_loss:                                                          # PPO:280-285
        vc = V0 + (v - V0).clamp(-clip_param, clip_param)
        l1, l2 = (v - R).pow(2), (vc - R).pow(2)
        value_loss = torch.max(l1, l2).mean()
    else:
        value_loss = (R - v).pow(2).mean()
    loss = surrogate_loss + value_loss_coef * value_loss - entropy_coef * ent.mean()
    # ---- backward (hand-derived; torch.max sends the gradient to the first argument on ties)
    use1 = s1 >= s2
    in_clip = (ratio >= 1.0 - clip_param) & (ratio <= 1.0 + clip_param)
    dratio = torch.where(use1, -A, torch.where(in_clip, -A, torch.zeros_like(A))) / M
    dlp = dratio * ratio
    diff = b["actions"] - mu
    dmu = dlp.unsqueeze(1) * diff / sigma ** 2
    dstd = (dlp.unsqueeze(1) * (diff ** 2 / sigma ** 3 - 1.0 / sigma)).sum(0) - entropy_coef * (1.0 / std)
    if use_clipped_value_loss:
        usel1 = l1 >= l2
        inc = ((v - V0) >= -clip_param) & ((v - V0) <= clip_param)
        dv = torch.where(usel1, 2 * (v - R), torch.where(inc, 2 * (vc - R), torch.zeros_like(v))) * (value_loss_coef / M)
    else:
        dv = -2 * (R - v) * (value_loss_coef / M)
    g = {"std": dstd}
    for net, hs, dy in (("actor", ha, dmu), ("critic", hc, dv)):
        keys = layer_keys(net, len(hs))
        for i in reversed(range(len(keys))):
            wk, bk = keys[i]
            g[wk] = dy.t() @ hs[i]
            g[bk] = dy.sum(0)
            if i > 0:
                dh = dy @ p[wk]
                dy = dh * torch.where(hs[i] > 0, torch.ones_like(hs[i]), hs[i] + 1.0)   # ELU'(z) = 1 (z>0) else e^z = h + 1
    stats = dict(loss=loss, surrogate_loss=surrogate_loss, value_loss=value_loss, kl_mean=kl_mean, entropy=ent.mean())
    return stats, g


def update_learning_rate(lr, kl_mean, desired_kl, lr_min, lr_max):                     # PPO:207-213
    if kl_mean > desired_kl * 2.0:
        return max(lr_min, lr / 1.5)
    if kl_mean < desired_kl / 2.0 and kl_mean > 0.0:
        return min(lr_max, lr * 1.5)
    return lr


def clip_grad_norm_(g, max_norm):                                                       # nn.utils.clip_grad_norm_ (PPO:304)
    total = torch.sqrt(sum((x.double() ** 2).sum() for x in g.values())).float()
    coef = torch.clamp(max_norm / (total + 1e-6), max=1.0)
    for k in g:
        g[k] = g[k] * coef
    return total


def adam_step(p, g, state, lr, beta1=0.9, beta2=0.999, eps=1e-8):
    """torch.optim.Adam (PPO:81, defaults; weight_decay 0).  state: dict(step=int, m={}, v={})."""
    state["step"] += 1
    t = state["step"]
    bc1, bc2 = 1 - beta1 ** t, 1 - beta2 ** t
    for k in p:
        m = state["m"].setdefault(k, torch.zeros_like(p[k]))
        v = state["v"].setdefault(k, torch.zeros_like(p[k]))
        m.mul_(beta1).add_(g[k], alpha=1 - beta1)
        v.mul_(beta2).addcmul_(g[k], g[k], value=1 - beta2)
        denom = (v.sqrt() / math.sqrt(bc2)).add_(eps)
        p[k] = p[k] - (lr / bc1) * (m / denom)


def ppo_update(p, adam, storage, indices, cfg, lr):
    """PPO.update (PPO:215-321) over a rollout.  storage: flat [T*N, .] tensors; indices: the single permutation
    reused for all epochs (RS:75).  Returns (mean_value_loss, mean_surrogate_loss, lr, per-minibatch log)."""
    nmb, nep = cfg["num_mini_batches"], cfg["num_learning_epochs"]
    mbs = indices.numel() // nmb
    mvl = msl = 0.0
    log = []
    for ep in range(nep):
        for i in range(nmb):
            idx = indices[i * mbs:(i + 1) * mbs]
            b = {k: v[idx] for k, v in storage.items()}
            stats, g = minibatch_loss_and_grads(p, b, cfg["clip_param"], cfg["value_loss_coef"], cfg["entropy_coef"],
                                                cfg["use_clipped_value_loss"])
            if cfg.get("desired_kl") is not None and cfg["schedule"] == "adaptive":
                lr = update_learning_rate(lr, float(stats["kl_mean"]), cfg["desired_kl"], cfg["learning_rate_min"],
                                          cfg["learning_rate_max"])
            log.append(dict(kl=float(stats["kl_mean"]), lr=lr, loss=float(stats["loss"]),
                            value_loss=float(stats["value_loss"]), surrogate_loss=float(stats["surrogate_loss"])))
            if torch.isnan(stats["loss"]):                                              # PPO:297-299
                continue
            gnorm = clip_grad_norm_(g, cfg["max_grad_norm"])
            log[-1]["grad_norm"] = float(gnorm)
            adam_step(p, g, adam, lr)
            mvl += float(stats["value_loss"])
            msl += float(stats["surrogate_loss"])
    n = nep * nmb
    return mvl / n, msl / n, lr, log
