"""Pin the PPO oracle (hand-derived backward, no autograd) to the UNMODIFIED reference rsl_rl
(golden fixtures from oracle/ref_harness/gen_ppo_golden.py)."""
import os

import numpy as np
import pytest
import torch

from grx_b200.config import make_train_cfg
from oracle import ppo_oracle as po

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _load(name):
    fx = dict(np.load(os.path.join(GOLDEN, f"ppo_{name}.npz")))
    t = lambda k: torch.from_numpy(fx[k])
    return fx, t


@pytest.mark.parametrize("name", ["small", "small_hot"])
def test_rollout_and_gae_match_rsl_rl(name):
    fx, t = _load(name)
    N, T, nmb, nep, O, P, A = [int(v) for v in fx["meta/dims"]]
    p = {k[len("init/"):]: t(k) for k in fx if k.startswith("init/")}
    cfg = make_train_cfg()["algorithm"]
    vals, rews = [], []
    for s in range(T):
        out = po.act(p, t("roll/obs")[s], t("roll/critic_obs")[s], t("roll/eps")[s])
        np.testing.assert_allclose(out["actions"].numpy(), fx["storage/actions"][s], rtol=1e-6, atol=1e-6)
        np.testing.assert_allclose(out["values"].numpy(), fx["storage/values"][s], rtol=1e-6, atol=1e-6)
        np.testing.assert_allclose(out["actions_log_prob"].numpy(), fx["storage/actions_log_prob"][s, :, 0], rtol=1e-5, atol=1e-5)
        np.testing.assert_allclose(out["action_mean"].numpy(), fx["storage/mu"][s], rtol=1e-6, atol=1e-6)
        r = po.process_rewards(t("roll/rewards")[s], out["values"], t("roll/time_outs")[s], cfg["gamma"])
        np.testing.assert_allclose(r.numpy(), fx["storage/rewards"][s, :, 0], rtol=1e-6, atol=1e-7)
        vals.append(out["values"]); rews.append(r.unsqueeze(1))
    last_v = po.mlp_forward(p, "critic", t("roll/last_critic_obs"))
    ret, adv = po.compute_returns(torch.stack(rews), t("roll/dones").unsqueeze(-1), torch.stack(vals), last_v, cfg["gamma"], cfg["lam"])
    np.testing.assert_allclose(ret.numpy(), fx["storage/returns"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(adv.numpy(), fx["storage/advantages"], rtol=1e-4, atol=1e-5)


@pytest.mark.parametrize("name", ["small", "small_hot"])
def test_update_matches_rsl_rl(name):
    """8x25-style minibatch loop: KL -> adaptive LR sequence, clipped surrogate/value loss, grad clip, Adam."""
    fx, t = _load(name)
    N, T, nmb, nep, O, P, A = [int(v) for v in fx["meta/dims"]]
    p = {k[len("init/"):]: t(k).clone() for k in fx if k.startswith("init/")}
    cfg = dict(make_train_cfg()["algorithm"], num_mini_batches=nmb, num_learning_epochs=nep)
    flat = lambda k: t(k).flatten(0, 1)
    storage = dict(obs=flat("roll/obs"), critic_obs=flat("roll/critic_obs"), actions=flat("storage/actions"),
                   values=flat("storage/values"), advantages=flat("storage/advantages"), returns=flat("storage/returns"),
                   old_log_prob=flat("storage/actions_log_prob"), old_mu=flat("storage/mu"), old_sigma=flat("storage/sigma"))
    adam = dict(step=0, m={}, v={})
    mvl, msl, lr, log = po.ppo_update(p, adam, storage, t("update/indices"), cfg, float(fx["meta/lr0"]))
    ref = fx["update/kl_lr"]
    assert len(log) == len(ref) == nmb * nep
    np.testing.assert_allclose([l["kl"] for l in log], ref[:, 0], rtol=2e-3, atol=2e-6)
    np.testing.assert_allclose([l["lr"] for l in log], ref[:, 1], rtol=1e-12)
    np.testing.assert_allclose([mvl, msl], fx["update/mean_losses"], rtol=1e-4, atol=1e-6)
    assert adam["step"] == int(fx["adam_step"])
    for k in p:
        np.testing.assert_allclose(p[k].numpy(), fx["final/" + k], rtol=2e-4, atol=2e-6, err_msg=k)
        np.testing.assert_allclose(adam["m"][k].numpy(), fx["adam_m/" + k], rtol=2e-3, atol=1e-7, err_msg="m " + k)


def test_hand_backward_equals_autograd():
    """The hand-derived gradient == torch.autograd of the same loss at the registered task's width."""
    torch.manual_seed(0)
    O, P, A, M = 39, 168, 10, 257
    p = po.init_params(O, P, A)
    b = dict(obs=torch.randn(M, O), critic_obs=torch.randn(M, P), actions=torch.randn(M, A) * 0.3,
             values=torch.randn(M, 1) * 0.1, advantages=torch.randn(M, 1), returns=torch.randn(M, 1) * 0.2,
             old_log_prob=torch.randn(M, 1) * 0.1 + 5, old_mu=torch.randn(M, A) * 0.1, old_sigma=torch.full((M, A), 0.2))
    b["old_log_prob"] = po.log_prob(b["actions"], po.mlp_forward(p, "actor", b["obs"]) + 0.02 * torch.randn(M, A), b["old_sigma"]).unsqueeze(1)
    stats, g = po.minibatch_loss_and_grads(p, b)
    pa = {k: v.clone().requires_grad_(True) for k, v in p.items()}
    mu = po.mlp_forward(pa, "actor", b["obs"]); v = po.mlp_forward(pa, "critic", b["critic_obs"])
    dist = torch.distributions.Normal(mu, mu * 0.0 + pa["std"])
    lp = dist.log_prob(b["actions"]).sum(-1)
    ratio = torch.exp(lp - b["old_log_prob"].squeeze(-1)); A_ = b["advantages"].squeeze(-1)
    sl = torch.max(-A_ * ratio, -A_ * torch.clamp(ratio, 0.8, 1.2)).mean()
    vc = b["values"] + (v - b["values"]).clamp(-0.2, 0.2)
    vl = torch.max((v - b["returns"]).pow(2), (vc - b["returns"]).pow(2)).mean()
    loss = sl + 1.0 * vl - 0.01 * dist.entropy().sum(-1).mean()
    loss.backward()
    assert abs(float(loss) - float(stats["loss"])) < 1e-5
    for k in p:
        np.testing.assert_allclose(g[k].numpy(), pa[k].grad.numpy(), rtol=2e-3, atol=2e-6, err_msg=k)


@pytest.mark.parametrize("name", ["small", "small_hot"])
def test_autograd_port_matches_rsl_rl(name):
    """oracle/ppo_autograd.py (the CPU baseline arm of bench.py: nn.Module + autograd + torch.optim.Adam, as rsl_rl runs) reproduces
    the unmodified rsl_rl update: KL / LR sequence and final weights."""
    from oracle import ppo_autograd as pa
    fx, t = _load(name)
    N, T, nmb, nep, O, P, A = [int(v) for v in fx["meta/dims"]]
    hidden = [int(v) for v in fx["meta/hidden"]]
    ac = pa.ActorCritic(O, P, A, hidden, hidden)
    ac.load_state_dict({k[len("init/"):]: t(k) for k in fx if k.startswith("init/")})
    cfg = make_train_cfg()["algorithm"]
    step = pa.PPOStep(ac, clip=cfg["clip_param"], vcoef=cfg["value_loss_coef"], ecoef=cfg["entropy_coef"], lr=float(fx["meta/lr0"]),
                      lr_min=cfg["learning_rate_min"], lr_max=cfg["learning_rate_max"], desired_kl=cfg["desired_kl"], max_grad_norm=cfg["max_grad_norm"])
    flat = lambda k: t(k).flatten(0, 1)
    st = dict(obs=flat("roll/obs"), critic_obs=flat("roll/critic_obs"), actions=flat("storage/actions"), values=flat("storage/values"),
              advantages=flat("storage/advantages"), returns=flat("storage/returns"), old_log_prob=flat("storage/actions_log_prob"),
              old_mu=flat("storage/mu"), old_sigma=flat("storage/sigma"))
    idx, B = t("update/indices"), (N * T) // nmb
    for ep in range(nep):
        for mb in range(nmb):
            sel = idx[mb * B:(mb + 1) * B]
            step.minibatch({k: v[sel] for k, v in st.items()})
    ref = fx["update/kl_lr"]
    np.testing.assert_allclose([k for k, _ in step.kl_log], ref[:, 0], rtol=1e-4, atol=1e-7)
    np.testing.assert_allclose([l for _, l in step.kl_log], ref[:, 1], rtol=1e-12)
    for k, v in ac.state_dict().items():
        np.testing.assert_allclose(v.numpy(), fx["final/" + k], rtol=1e-4, atol=1e-6, err_msg=k)
