"""OnPolicyRunner with the reference's surface (rsl_rl/rsl_rl/runners/on_policy_runner.py:16-345): same constructor
(env, train_cfg dict, log_dir, device), ``learn`` / ``save`` / ``load`` / ``get_inference_policy`` and the same checkpoint
dict (``model_state_dict``, ``optimizer_state_dict``, ``iter``, ``infos``), driving the CUDA env and PPO kernels.

The rollout loop issues, per policy step, the policy-forward kernels, ONE env kernel and one storage kernel on the current
stream and never synchronises with the host; timing uses CUDA events (the reference's time.time() around asynchronous
launches, on_policy_runner.py:146-196, does not measure the device).
"""
from __future__ import annotations

import os
import statistics
import time
from collections import deque

import torch

from .ppo import PPO, ActorCriticMLP


class OnPolicyRunner:
    def __init__(self, env, train_cfg, log_dir=None, device="cuda:0", world_size=1, process_group=None):
        self.cfg = train_cfg["runner"]
        self.algorithm_cfg = dict(train_cfg["algorithm"])
        self.policy_cfg = dict(train_cfg["policy"])
        self.device = device
        self.env = env
        critic_in = env.num_pri_obs if env.num_pri_obs is not None else env.num_obs       # on_policy_runner.py:71-76
        actor_critic = ActorCriticMLP(env.num_obs, critic_in, env.num_actions, **self.policy_cfg)
        # the action-noise stream is keyed by (task seed, GLOBAL env id, step): sharded ranks draw what one big env would
        rng_kw = dict(seed=int(getattr(getattr(env, "cfg", None), "seed", 1)), env_id_offset=int(getattr(env, "env_id_offset", 0)))
        rng_kw.update({k: self.algorithm_cfg.pop(k) for k in ("seed", "env_id_offset") if k in self.algorithm_cfg})
        self.algorithm = PPO(actor_critic=actor_critic, device=device, world_size=world_size, process_group=process_group,
                             **rng_kw, **self.algorithm_cfg)
        self.alg = self.algorithm
        self.num_steps_per_env = self.cfg["num_steps_per_env"]
        self.save_interval = self.cfg["save_interval"]
        self.algorithm.init_storage(env.num_envs, self.num_steps_per_env)
        self.world_size = world_size
        self.env.reset()                                                                  # on_policy_runner.py:105
        self.log_dir = log_dir
        self.writer = None
        self.tot_timesteps = 0
        self.tot_time = 0
        self.current_learning_iteration = 0
        self.last_timing = {}

    def learn(self, num_learning_iterations, init_at_random_ep_len=False):                # on_policy_runner.py:115-207
        env, alg = self.env, self.algorithm
        if self.log_dir is not None and self.writer is None:
            try:
                from torch.utils.tensorboard import SummaryWriter
                self.writer = SummaryWriter(log_dir=self.log_dir, flush_secs=10)
            except Exception:
                self.writer = None
        if init_at_random_ep_len:
            env.episode_length_buf = torch.randint_like(env.episode_length_buf, high=int(env.max_episode_length))
        obs = env.get_observations()
        pri = env.get_privileged_observations()
        critic_obs = pri if pri is not None else obs
        ep_infos = []
        rewbuffer, lenbuffer = deque(maxlen=100), deque(maxlen=100)
        log = self.log_dir is not None
        if log:
            cur_reward_sum = torch.zeros(env.num_envs, dtype=torch.float, device=self.device)
            cur_episode_length = torch.zeros(env.num_envs, dtype=torch.float, device=self.device)
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        tot_iter = self.current_learning_iteration + num_learning_iterations
        for it in range(self.current_learning_iteration, tot_iter):
            ev[0].record()
            for _ in range(self.num_steps_per_env):
                actions = alg.act(obs, critic_obs)
                obs, pri, rewards, dones, infos = env.step(actions)
                critic_obs = pri if pri is not None else obs
                alg.process_env_step(rewards, dones, infos)
                if log:   # bookkeeping exactly as on_policy_runner.py:171-181 (host syncs; only when logging)
                    if "episode" in infos:
                        ep_infos.append(infos["episode"])
                    cur_reward_sum += rewards
                    cur_episode_length += 1
                    new_ids = (dones > 0).nonzero(as_tuple=False)
                    rewbuffer.extend(cur_reward_sum[new_ids][:, 0].cpu().numpy().tolist())
                    lenbuffer.extend(cur_episode_length[new_ids][:, 0].cpu().numpy().tolist())
                    cur_reward_sum[new_ids] = 0
                    cur_episode_length[new_ids] = 0
            ev[1].record()
            alg.compute_returns(critic_obs)
            mean_value_loss, mean_surrogate_loss = alg.update()
            alg.clear_storage()
            ev[2].record()
            if log or it == tot_iter - 1:
                ev[2].synchronize()
                collection_time, learn_time = ev[0].elapsed_time(ev[1]) / 1e3, ev[1].elapsed_time(ev[2]) / 1e3
                self.last_timing = dict(collection_time=collection_time, learn_time=learn_time)
            if log:
                self.log(locals())
                if it % self.save_interval == 0:
                    self.save(os.path.join(self.log_dir, "model_{}.pt".format(it)))
            ep_infos.clear()
        self.current_learning_iteration += num_learning_iterations
        if log:
            self.save(os.path.join(self.log_dir, "model_{}.pt".format(self.current_learning_iteration)))

    def log(self, locs, width=80, pad=35):                                                # on_policy_runner.py:209-295 (same scalar tags)
        n_total = self.num_steps_per_env * self.env.num_envs * self.world_size
        self.tot_timesteps += n_total
        self.tot_time += locs["collection_time"] + locs["learn_time"]
        iteration_time = locs["collection_time"] + locs["learn_time"]
        fps = int(n_total / iteration_time)
        ep_string = ""
        scalars = {}
        if locs["ep_infos"]:
            for key in locs["ep_infos"][0]:
                vals = [ep[key] if isinstance(ep[key], torch.Tensor) else torch.tensor(ep[key]) for ep in locs["ep_infos"]]
                value = torch.mean(torch.stack([v.float().reshape(()) for v in vals]))
                scalars["Episode/" + key] = float(value)
                ep_string += f"""{f'Mean episode {key}:':>{pad}} {float(value):.4f}\n"""
        mean_std = self.algorithm.actor_critic.std.mean()
        scalars.update({"Loss/value_function": float(locs["mean_value_loss"]), "Loss/surrogate": float(locs["mean_surrogate_loss"]),
                        "Loss/learning_rate": self.algorithm.learning_rate, "Policy/mean_noise_std": float(mean_std),
                        "Perf/total_fps": fps, "Perf/collection time": locs["collection_time"], "Perf/learning_time": locs["learn_time"]})
        for i, s in enumerate(self.algorithm.actor_critic.std.tolist()):
            scalars[f"Policy/noise_std_{i}"] = s
        if len(locs["rewbuffer"]) > 0:
            scalars["Train/mean_reward"] = statistics.mean(locs["rewbuffer"])
            scalars["Train/mean_episode_length"] = statistics.mean(locs["lenbuffer"])
        if self.writer is not None:
            for k, v in scalars.items():
                self.writer.add_scalar(k, v, locs["it"])
        head = f" \033[1m Learning iteration {locs['it']}/{locs['tot_iter']} \033[0m "
        print(f"""{'#' * width}\n{head.center(width, ' ')}\n\n"""
              f"""{'Computation:':>{pad}} {fps:.0f} steps/s (collection: {locs['collection_time']:.3f}s, learning {locs['learn_time']:.3f}s)\n"""
              f"""{'Value function loss:':>{pad}} {float(locs['mean_value_loss']):.4f}\n"""
              f"""{'Surrogate loss:':>{pad}} {float(locs['mean_surrogate_loss']):.4f}\n"""
              f"""{'Mean action noise std:':>{pad}} {float(mean_std):.2f}\n""" + ep_string)
        self.last_scalars = scalars

    def save(self, path, infos=None):                                                     # on_policy_runner.py:297-309
        self.algorithm.check_comm(wait=True)   # never checkpoint replicas that stopped stepping because a peer went missing
        torch.save({"model_state_dict": {k: v.detach().contiguous().cpu().clone() for k, v in self.algorithm.actor_critic.state_dict().items()},
                    "optimizer_state_dict": self.algorithm.optimizer_state_dict(),
                    "iter": self.current_learning_iteration, "infos": infos}, path)

    def load(self, path, load_optimizer=True):                                            # on_policy_runner.py:311-331
        loaded = torch.load(path, map_location="cpu", weights_only=False)
        self.algorithm.actor_critic.load_state_dict(loaded["model_state_dict"])
        if load_optimizer:
            self.algorithm.load_optimizer_state_dict(loaded["optimizer_state_dict"])
        self.current_learning_iteration = loaded["iter"]
        return loaded["infos"]

    def get_inference_policy(self, device=None):                                          # on_policy_runner.py:333-345
        return self.algorithm.actor_critic.act_inference
