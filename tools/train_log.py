#!/usr/bin/env python
"""Training evidence: run OnPolicyRunner.learn for K iterations at the BASELINE size and keep the reference's own scalars
(on_policy_runner.py:233-261: Train/mean_reward, Train/mean_episode_length, Episode/terrain_level, losses, noise std, fps).

  python tools/train_log.py --robot GR1T1 --mesh heightfield --envs 4096 --iters 300 --out gpurun_out/train_hf.jsonl

One JSON line every `--every` iterations + a final summary line comparing the first and the last tenth of the run."""
import argparse
import contextlib
import io
import json
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "wiki-grx-gym_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--robot", default="GR1T1")
    ap.add_argument("--mesh", default="heightfield")
    ap.add_argument("--envs", type=int, default=4096)
    ap.add_argument("--iters", type=int, default=300)
    ap.add_argument("--every", type=int, default=10)
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--out", default=None)
    ap.add_argument("--full-body", action="store_true", help="the unregistered full-body 32-DOF task (generic-topology kernels, self-collision)")
    ap.add_argument("--no-self-collision", action="store_true", help="full body: robot self-collision off (cfg.asset.self_collisions = 1)")
    args = ap.parse_args()
    import torch
    from grx_b200.config import make_cfg, make_full_body_cfg, make_train_cfg
    from grx_b200.env import GRXVecEnv
    from grx_b200.runner import OnPolicyRunner
    torch.manual_seed(args.seed)
    cfg = (make_full_body_cfg if args.full_body else make_cfg)(args.robot, args.envs, args.mesh)
    cfg.seed = args.seed
    if args.no_self_collision:
        cfg.asset.self_collisions = 1   # legged_robot_config.py:121: 1 = disabled
    env = GRXVecEnv(cfg, sim_device="cuda:0")
    tc = make_train_cfg()
    tc["runner"]["save_interval"] = 10 ** 9
    tmp = tempfile.mkdtemp(prefix="grx_train_")
    runner = OnPolicyRunner(env, tc, log_dir=tmp, device="cuda:0")
    rows = []
    out = open(args.out, "w") if args.out else None
    orig_log = runner.log
    t0 = time.time()

    def log(locs, **kw):
        with contextlib.redirect_stdout(io.StringIO()):
            orig_log(locs, **kw)
        s = runner.last_scalars
        it = locs["it"]
        row = {"it": it, "mean_reward": s.get("Train/mean_reward"), "mean_episode_length": s.get("Train/mean_episode_length"),
               "terrain_level": s.get("Episode/terrain_level"), "value_loss": s["Loss/value_function"], "surrogate": s["Loss/surrogate"],
               "lr": s["Loss/learning_rate"], "noise_std": s["Policy/mean_noise_std"], "fps": s["Perf/total_fps"],
               "rew_cmd_diff_lin_vel_x": s.get("Episode/rew_cmd_diff_lin_vel_x"), "wall_s": round(time.time() - t0, 1)}
        rows.append(row)
        if it % args.every == 0 or it == args.iters - 1:
            line = json.dumps(row)
            print(line, flush=True)
            if out:
                out.write(line + "\n"); out.flush()
    runner.log = log
    runner.learn(args.iters, init_at_random_ep_len=True)
    k = max(1, len(rows) // 10)

    def avg(key, part):
        v = [r[key] for r in part if r[key] is not None]
        return sum(v) / len(v) if v else None
    first, last = rows[k:2 * k], rows[-k:]      # skip the very first tenth's start-up (random episode lengths)
    summary = {"summary": True, "robot": args.robot, "mesh": args.mesh, "envs": args.envs, "iters": args.iters,
               "mean_reward_first": avg("mean_reward", first), "mean_reward_last": avg("mean_reward", last),
               "mean_episode_length_first": avg("mean_episode_length", first), "mean_episode_length_last": avg("mean_episode_length", last),
               "terrain_level_first": avg("terrain_level", first), "terrain_level_last": avg("terrain_level", last),
               "noise_std_last": rows[-1]["noise_std"], "fps_last": rows[-1]["fps"], "wall_s": round(time.time() - t0, 1)}
    line = json.dumps(summary)
    print(line, flush=True)
    if out:
        out.write(line + "\n"); out.close()


if __name__ == "__main__":
    main()
