"""TEST INFRASTRUCTURE — golden PPO fixtures from the UNMODIFIED reference rsl_rl (build container only).
python -m oracle.ref_harness.gen_ppo_golden

One fixture = one rollout + one PPO.update() at reduced width (hidden 64/32/16, N=48, T=6) so it stays small;
the arithmetic path is the registered task's (ActorCriticMLP + RolloutStorage + PPO, adaptive-KL schedule)."""
import contextlib
import io
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, "..", ".."))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "wiki-grx-gym_b200"))
from oracle.ref_harness import stub  # noqa: E402
sys.path.insert(0, os.path.join(ROOT, "tests"))
from golden_util import wide_inputs  # noqa: E402  (shared with tests/test_ppo_gpu.py: the inputs are regenerated, not stored)


def gen_wide():
    """rsl_rl at the REGISTERED width (512/256/128): one rollout (N=256, T=16) + PPO.update() with 4 minibatches x 2 epochs.  Stored:
    KL/LR log, mean losses, returns / advantages, final weights, strided Adam moments.  Initial weights and inputs are regenerated
    from their seeds by the test (tests/test_ppo_gpu.py::test_wide_*)."""
    with contextlib.redirect_stdout(io.StringIO()):
        from rsl_rl.algorithms import PPO
        from rsl_rl.modules import ActorCriticMLP
    from grx_b200.config import make_train_cfg
    from torch.distributions import Normal
    tc = make_train_cfg("GR1T1")
    seed, N, T, nmb, nep, lr = 13, 256, 16, 4, 2, 1e-4
    O, P, A = 39, 168, 10
    torch.manual_seed(seed)
    with contextlib.redirect_stdout(io.StringIO()):
        ac = ActorCriticMLP(O, P, A, **tc["policy"])
        alg = PPO(actor_critic=ac, device="cpu", **dict(tc["algorithm"], num_mini_batches=nmb, num_learning_epochs=nep, learning_rate=lr))
        alg.init_storage(N, T)
    d, digest = wide_inputs(seed, N, T, O, P, A)
    out = {"meta/dims": np.array([N, T, nmb, nep, O, P, A]), "meta/hidden": np.array(tc["policy"]["actor_hidden_dims"]), "meta/lr0": np.array(lr),
           "meta/seed": np.array(seed), "meta/inputs_sha256": np.array(digest)}
    cur = {}
    orig_sample = Normal.sample
    Normal.sample = lambda self, sample_shape=torch.Size(): self.mean + self.stddev * cur["eps"]
    with torch.inference_mode():
        for t in range(T):
            cur["eps"] = d["eps"][t]
            alg.act(d["obs"][t], d["critic_obs"][t])
            alg.process_env_step(d["rewards"][t].clone(), d["dones"][t], {"time_outs": d["time_outs"][t]})
        alg.compute_returns(d["last_critic_obs"])
    Normal.sample = orig_sample
    st = alg.storage
    for k in ("values", "actions_log_prob", "returns", "advantages"):
        out["storage/" + k] = getattr(st, k).detach().clone().numpy()
    orig_randperm = torch.randperm
    torch.randperm = lambda n, **kw: d["indices"][:n]
    klog = []
    orig_ulr = alg.update_learning_rate

    def ulr(kl_mean):
        orig_ulr(kl_mean)
        klog.append((float(kl_mean), alg.learning_rate))
    alg.update_learning_rate = ulr
    mvl, msl = alg.update()
    torch.randperm = orig_randperm
    out["update/kl_lr"] = np.array(klog, np.float64)
    out["update/mean_losses"] = np.array([mvl, msl], np.float64)
    for k, v in ac.state_dict().items():
        out["final/" + k] = v.detach().clone().numpy()
    names = [n for n, _ in ac.named_parameters()]
    osd = alg.optimizer.state_dict()["state"]
    for i, n in enumerate(names):   # every 8th element: enough to pin the moments, a quarter of the bytes
        out["adam_m8/" + n] = osd[i]["exp_avg"].flatten()[::8].numpy()
        out["adam_v8/" + n] = osd[i]["exp_avg_sq"].flatten()[::8].numpy()
    out["adam_step"] = np.array(float(osd[0]["step"]))
    path = os.path.join(ROOT, "tests", "golden", "ppo_wide.npz")
    np.savez_compressed(path, **out)
    print("wide kl/lr:", klog, "losses", mvl, msl, f"{os.path.getsize(path) / 1e3:.0f} kB")


def main():
    stub.install()
    gen_wide()
    with contextlib.redirect_stdout(io.StringIO()):
        from rsl_rl.algorithms import PPO
        from rsl_rl.modules import ActorCriticMLP
    from grx_b200.config import make_train_cfg
    tc = make_train_cfg("GR1T1")
    for name, seed, N, T, nmb, nep, hidden, lr in [("small", 11, 48, 6, 4, 3, [64, 32, 16], 1e-4),
                                                    ("small_hot", 12, 40, 5, 3, 4, [32, 32, 16], 1e-3)]:
        torch.manual_seed(seed)
        O, P, A = 39, 168, 10
        pol = dict(tc["policy"], actor_hidden_dims=hidden, critic_hidden_dims=hidden)
        alg_cfg = dict(tc["algorithm"], num_mini_batches=nmb, num_learning_epochs=nep, learning_rate=lr)
        with contextlib.redirect_stdout(io.StringIO()):
            ac = ActorCriticMLP(O, P, A, **pol)
            alg = PPO(actor_critic=ac, device="cpu", **alg_cfg)
            alg.init_storage(N, T)
        out = {"meta/dims": np.array([N, T, nmb, nep, O, P, A]), "meta/hidden": np.array(hidden), "meta/lr0": np.array(lr)}
        for k, v in ac.state_dict().items():
            out["init/" + k] = v.detach().clone().numpy()
        g = torch.Generator().manual_seed(seed + 100)
        eps_all = torch.randn(T, N, A, generator=g)
        from torch.distributions import Normal
        cur = {}
        orig_sample = Normal.sample
        Normal.sample = lambda self, sample_shape=torch.Size(): self.mean + self.stddev * cur["eps"]
        obs = torch.randn(N, O, generator=g)
        cobs = torch.randn(N, P, generator=g)
        roll = {k: [] for k in ("obs", "critic_obs", "rewards", "dones", "time_outs")}
        with torch.inference_mode():
            for t in range(T):
                cur["eps"] = eps_all[t]
                roll["obs"].append(obs.clone()); roll["critic_obs"].append(cobs.clone())
                alg.act(obs, cobs)
                rew = torch.randn(N, generator=g) * 0.1
                dones = torch.rand(N, generator=g) < 0.15
                touts = dones & (torch.rand(N, generator=g) < 0.5)
                roll["rewards"].append(rew.clone()); roll["dones"].append(dones.clone()); roll["time_outs"].append(touts.clone())
                alg.process_env_step(rew, dones, {"time_outs": touts})
                obs = torch.randn(N, O, generator=g); cobs = torch.randn(N, P, generator=g)
            alg.compute_returns(cobs)
        Normal.sample = orig_sample
        out["roll/eps"] = eps_all.numpy()
        out["roll/last_critic_obs"] = cobs.numpy()
        for k, v in roll.items():
            out["roll/" + k] = torch.stack(v).numpy()
        st = alg.storage
        for k in ("actions", "values", "actions_log_prob", "mu", "sigma", "rewards", "returns", "advantages"):
            out["storage/" + k] = getattr(st, k).detach().clone().numpy()
        perm = {}
        orig_randperm = torch.randperm

        def randperm(n, **kw):
            perm["idx"] = orig_randperm(n, generator=g)
            return perm["idx"]
        torch.randperm = randperm
        klog = []
        orig_ulr = alg.update_learning_rate

        def ulr(kl_mean):
            orig_ulr(kl_mean)
            klog.append((float(kl_mean), alg.learning_rate))
        alg.update_learning_rate = ulr
        mvl, msl = alg.update()
        torch.randperm = orig_randperm
        out["update/indices"] = perm["idx"].numpy()
        out["update/kl_lr"] = np.array(klog, np.float64)
        out["update/mean_losses"] = np.array([mvl, msl], np.float64)
        for k, v in ac.state_dict().items():
            out["final/" + k] = v.detach().clone().numpy()
        names = [n for n, _ in ac.named_parameters()]
        osd = alg.optimizer.state_dict()["state"]
        for i, n in enumerate(names):
            out["adam_m/" + n] = osd[i]["exp_avg"].numpy()
            out["adam_v/" + n] = osd[i]["exp_avg_sq"].numpy()
        out["adam_step"] = np.array(float(osd[0]["step"]))
        path = os.path.join(ROOT, "tests", "golden", f"ppo_{name}.npz")
        np.savez_compressed(path, **out)
        print(name, "kl/lr first,last:", klog[0], klog[-1], "losses", mvl, msl, f"{os.path.getsize(path) / 1e3:.0f} kB")


if __name__ == "__main__":
    main()
