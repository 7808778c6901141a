"""TEST INFRASTRUCTURE — the reference's PPO step restated the way rsl_rl itself runs it: torch.nn modules, autograd backward,
torch.optim.Adam (rsl_rl/rsl_rl/algorithms/ppo.py:81, 215-321; modules/actor_critic_mlp.py:160-231; modules/mlp.py:7-42).
Used ONLY as the CPU baseline arm of bench.py (`--impl reference` / `cpu_baseline`): the hand-derived backward of
oracle/ppo_oracle.py is a checker, not a fair speed reference (autograd + fused Adam are what the reference pays on a CPU).
Pinned to the unmodified rsl_rl by tests/test_ppo_oracle.py::test_autograd_port_matches_rsl_rl (golden fixtures).
"""
from __future__ import annotations

import torch
from torch import nn
from torch.distributions import Normal


class MLP(nn.Module):                                                                  # mlp.py:7-42
    def __init__(self, n_in, n_out, hidden):
        super().__init__()
        dims = [n_in, *hidden, n_out]
        layers = []
        for i in range(len(dims) - 1):
            layers.append(nn.Linear(dims[i], dims[i + 1]))
            if i < len(dims) - 2:
                layers.append(nn.ELU())
        self.model = nn.Sequential(*layers)

    def forward(self, x):
        return self.model(x)


class ActorCritic(nn.Module):                                                          # actor_critic_mlp.py:10-231
    def __init__(self, num_obs, num_pri_obs, num_actions, actor_hidden=(512, 256, 128), critic_hidden=(512, 256, 128), init_noise_std=0.2):
        super().__init__()
        self.actor, self.critic = MLP(num_obs, num_actions, actor_hidden), MLP(num_pri_obs, 1, critic_hidden)
        self.std = nn.Parameter(init_noise_std * torch.ones(num_actions))
        self.distribution = None

    def update_distribution(self, obs):                                                # ACM:176-181
        mean = self.actor(obs)
        self.distribution = Normal(mean, mean * 0.0 + self.std)

    def act(self, obs):
        self.update_distribution(obs)
        return self.distribution.sample()


class PPOStep:
    """act() for the rollout and one minibatch of PPO.update() (ppo.py:244-305) incl. the adaptive-KL learning rate."""

    def __init__(self, ac, clip=0.2, vcoef=1.0, ecoef=0.01, lr=1e-4, lr_min=1e-5, lr_max=1e-3, desired_kl=0.03, max_grad_norm=1.0):
        self.ac, self.clip, self.vcoef, self.ecoef = ac, clip, vcoef, ecoef
        self.lr, self.lr_min, self.lr_max, self.desired_kl, self.max_grad_norm = lr, lr_min, lr_max, desired_kl, max_grad_norm
        self.opt = torch.optim.Adam(ac.parameters(), lr=lr)                            # ppo.py:81
        self.kl_log = []

    @torch.no_grad()
    def act(self, obs, critic_obs):                                                    # ppo.py:144-171
        a = self.ac.act(obs)
        v = self.ac.critic(critic_obs)
        lp = self.ac.distribution.log_prob(a).sum(dim=-1)
        return a, v, lp, self.ac.distribution.mean, self.ac.distribution.stddev

    def minibatch(self, b):
        ac = self.ac
        ac.update_distribution(b["obs"])
        lp = ac.distribution.log_prob(b["actions"]).sum(dim=-1)
        value = ac.critic(b["critic_obs"])
        mu, sigma = ac.distribution.mean, ac.distribution.stddev
        entropy = ac.distribution.entropy().sum(dim=-1)
        with torch.inference_mode():                                                   # ppo.py:253-268
            kl = torch.sum(torch.log(sigma / b["old_sigma"] + 1.0e-5)
                           + (torch.square(b["old_sigma"]) + torch.square(b["old_mu"] - mu)) / (2.0 * torch.square(sigma)) - 0.5, axis=-1)
            kl_mean = torch.mean(kl)
            if kl_mean > self.desired_kl * 2.0:
                self.lr = max(self.lr_min, self.lr / 1.5)
            elif self.desired_kl / 2.0 > kl_mean > 0.0:
                self.lr = min(self.lr_max, self.lr * 1.5)
            for g in self.opt.param_groups:
                g["lr"] = self.lr
            self.kl_log.append((float(kl_mean), self.lr))
        adv = torch.squeeze(b["advantages"])
        ratio = torch.exp(lp - torch.squeeze(b["old_log_prob"]))                      # ppo.py:271-277
        surrogate = torch.max(-adv * ratio, -adv * torch.clamp(ratio, 1.0 - self.clip, 1.0 + self.clip)).mean()
        vclip = b["values"] + (value - b["values"]).clamp(-self.clip, self.clip)      # ppo.py:280-285
        value_loss = torch.max((value - b["returns"]).pow(2), (vclip - b["returns"]).pow(2)).mean()
        loss = surrogate + self.vcoef * value_loss - self.ecoef * entropy.mean()
        if torch.isnan(loss):                                                          # ppo.py:297-299
            return float(value_loss.detach()), float(surrogate.detach())
        self.opt.zero_grad()
        loss.backward()
        nn.utils.clip_grad_norm_(ac.parameters(), self.max_grad_norm)                 # ppo.py:304
        self.opt.step()
        return float(value_loss.detach()), float(surrogate.detach())
