#!/usr/bin/env python
"""bench.py — env-steps/sec of GRx rough-terrain PPO (BASELINE.json metric) on N B200s.

One "step" = one PPO iteration of the reference's OnPolicyRunner.learn loop (on_policy_runner.py:145-207):
64 policy steps x `envs/GPU` robots (policy forward -> fused env kernel -> storage), GAE, and the 8x25-minibatch PPO
update.  value = N_total * 64 * K / (time of K iterations) == the reference's Perf/total_fps (on_policy_runner.py:235).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--config 2|3|5|6]  # our arm (CUDA, no fallback)
  python bench.py --impl reference ...                                    # the reference's CPU path: the oracle port on host cores

--config selects the BASELINE.json workload (default 2 == the configuration the metric is quoted on; #4 is #2 at --gpus 8):
  2: GR1T1, rough heightfield + curriculum, 4096 envs/GPU        3: GR1T2 + full domain randomisation, heightfield, 8192 envs/GPU
  5: GR1T1, trimesh terrain + curriculum, 4096 envs/GPU (quoted at --gpus 4)
  6: (not a BASELINE config) the unregistered FULL-BODY 32-DOF GR1T1 with robot self-collision on the generic-topology
     kernels, heightfield + curriculum, 4096 envs/GPU
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "wiki-grx-gym_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

T_STEPS = 64
CONFIGS = {2: dict(robot="GR1T1", mesh="heightfield", envs=4096,
                   name="#2/#4 GR1T1 lower-limb (registered task), rough heightfield 10x20 tiles + curriculum + domain randomisation"),
           3: dict(robot="GR1T2", mesh="heightfield", envs=8192,
                   name="#3 GR1T2 lower-limb, rough heightfield 10x20 tiles + curriculum + full domain randomisation"),
           5: dict(robot="GR1T1", mesh="trimesh", envs=4096,
                   name="#5 GR1T1 lower-limb, trimesh terrain (slope_treshold 0.75) + curriculum + domain randomisation"),
           6: dict(robot="GR1T1", mesh="heightfield", envs=4096, full_body=True,
                   name="(extra, SURVEY 8-f3) GR1T1 FULL-BODY 32-DOF (gr1t1_config.py GR1T1Cfg, obs 105 / pri 234 / 32 actions) + robot self-collision, "
                        "generic-topology kernels, rough heightfield 10x20 tiles + curriculum + domain randomisation")}
METRIC, UNIT = "env-steps/sec GR1T1 rough-terrain PPO @4096 envs/GPU", "env-steps/s"
ENV_BYTES_PER_STEP = 1794          # 530 B read + 1264 B written per env-step by the fused env kernel
FWD_FLOP = 870144                  # per transition, actor + critic forward


def shapes(c):
    """(D, O, P, algorithmic env bytes per env-step, forward FLOP per transition) of config c: SURVEY.md §8(d)'s formulas — reads
    4(7D + 17) + 4R + 86, writes 4(18 + 6D + O + P) + 4R + 28 with R = 24 reward terms; FLOP = 2 x MACs of actor O-512-256-128-D + critic P-512-256-128-1."""
    D = 32 if CONFIGS[c].get("full_body") else 10
    O, P, R = 9 + 3 * D, 9 + 3 * D + 8 + 121, 24
    env_bytes = 4 * (7 * D + 17) + 4 * R + 86 + 4 * (18 + 6 * D + O + P) + 4 * R + 28
    flop = 2 * (O * 512 + 512 * 256 + 256 * 128 + 128 * D + P * 512 + 512 * 256 + 256 * 128 + 128 * 1)
    return D, O, P, env_bytes, flop


def make_task_cfg(c, n):
    from grx_b200.config import make_cfg, make_full_body_cfg
    cf = CONFIGS[c]
    return (make_full_body_cfg if cf.get("full_body") else make_cfg)(cf["robot"], n, cf["mesh"])


def workload(c):
    cf = CONFIGS[c]
    return (f"{cf['name']}, {cf['envs']} envs/GPU x {T_STEPS} steps/iteration, PPO 8 epochs x 25 minibatches of "
            f"{cf['envs'] * T_STEPS // 25}, live policy actions")


def metric_for(c):
    if c == 2:
        return METRIC
    if CONFIGS[c].get("full_body"):
        return f"env-steps/sec {CONFIGS[c]['robot']} full-body 32-DOF {CONFIGS[c]['mesh']} PPO @{CONFIGS[c]['envs']} envs/GPU (extra config #{c}, SURVEY 8-f3)"
    return f"env-steps/sec {CONFIGS[c]['robot']} {CONFIGS[c]['mesh']} PPO @{CONFIGS[c]['envs']} envs/GPU (BASELINE config #{c})"


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            d = json.load(f)
        return d["hbm_gbs"], d["bf16_tflops_sustained"], "measured"
    except Exception:
        return 6650.0, 1400.0, "fallback"


def measured_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum per launch, from the committed ncu --set full capture of the CURRENT kernels
    (profiles/traffic.json, written by tools/ncu_report.py from the .ncu-rep); None when no capture has been summarised."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            return json.load(f)
    except Exception:
        return {}


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md recipe).  The sampler is started BEFORE the warm-up
    iterations (nvidia-smi needs a few hundred ms to deliver its first line; the warm-up runs the same workload) and every line is stamped;
    the median is taken over the samples that fall inside [begin(), end()], or — when the timed region is shorter than a few sampling
    periods — over all samples under load since the start of the warm-up (`window` says which)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc, self.t0, self.t1 = index, [], None, None, None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "50"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.monotonic(), [x.strip() for x in line.split(",")]))

    def begin(self):
        self.t0 = time.monotonic()

    def end(self):
        self.t1 = time.monotonic()

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.12)
        self.proc.terminate()
        inside = [r for ts, r in self.rows if self.t0 is not None and self.t0 <= ts <= (self.t1 or ts) + 0.06]
        window = "timed region"
        if len(inside) < 3:
            inside, window = [r for _, r in self.rows], "warm-up + timed region (timed region shorter than 3 sampling periods)"
        sm, mx, reasons = [], None, set()
        for r in inside:
            try:
                sm.append(float(r[1])); mx = float(r[2])
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm), "window": window}


# ---------------------------------------------------------------------------------------------------------
# CPU arm: the reference's Python path restated (oracle/): EnvOracle (reference task arithmetic, torch CPU) over the C physics
# oracle (OpenMP), rsl_rl's PPO as nn.Module + autograd + torch.optim.Adam (oracle/ppo_autograd.py), all host cores.
# One reference "step" = 1/8 of a PPO iteration at the SAME config and env count: 8 policy steps (policy forward, env step, storage)
# + 25 minibatches of the full minibatch size (one of the 8 epochs) — the same 64:200 ratio as the whole iteration, so
# env-steps/s of the sample == env-steps/s of whole iterations; K >= 8 steps cover at least one whole iteration of work.
# ---------------------------------------------------------------------------------------------------------
class CpuArm:
    SUB = 8                       # a reference step = 1/SUB of an iteration

    def __init__(self, config):
        import numpy as np
        import torch
        from grx_b200.config import make_cfg, make_train_cfg
        from grx_b200.robot import sample_domain_rand, task_tables
        from grx_b200.terrain import Terrain
        from grx_b200.urdf import builtin_model
        from grx_b200 import rng_layout as RL
        from oracle import ppo_autograd as pa
        from oracle.env_oracle import EnvOracle
        from oracle.phys import PhysOracle
        self.torch, self.RL = torch, RL
        cf = CONFIGS[config]
        self.cores = os.cpu_count() or 1
        torch.set_num_threads(self.cores)
        os.environ.setdefault("OMP_NUM_THREADS", str(self.cores))
        N = self.N = cf["envs"]
        cfg = make_task_cfg(config, N)
        model = builtin_model(cf["robot"] + ("_full" if cf.get("full_body") else ""))
        tables = task_tables(model, cfg)
        D, O, P, _, _ = shapes(config)
        self.K = RL.layout(D).K
        np.random.seed(1)
        ter = Terrain(cfg.terrain, N)
        rng = np.random.default_rng(1)
        levels = rng.integers(0, cfg.terrain.max_init_terrain_level + 1, N)
        types = np.floor(np.arange(N) / (N / cfg.terrain.num_cols)).astype(np.int64)
        consts = sample_domain_rand(model, cfg, N, rng)
        consts.update(env_origins=ter.env_origins[levels, types], terrain_origins=ter.env_origins, terrain_levels=levels, terrain_types=types)
        terrain = dict(heights=ter.heightsamples, hscale=cfg.terrain.horizontal_scale, vscale=cfg.terrain.vertical_scale,
                       border=float(cfg.terrain.border_size), friction=1.0, restitution=0.0)
        if cf["mesh"] == "trimesh":   # structured trimesh: the vertex shifts of the reference's conversion (top-surface contact, as the CUDA arm)
            from grx_b200.terrain import heightfield_to_trimesh
            from oracle.phys import moves_from_vertices
            verts, _ = heightfield_to_trimesh(ter.heightsamples, cfg.terrain.horizontal_scale, cfg.terrain.vertical_scale, cfg.terrain.slope_treshold)
            terrain["moves"] = moves_from_vertices(verts, ter.heightsamples.shape[0], ter.heightsamples.shape[1], cfg.terrain.horizontal_scale)
        ctl, sim = dict(tables), {}
        if cf.get("full_body"):   # robot self-collision, the same candidate pairs and contact budget as the CUDA arm
            from grx_b200.robot import self_collision_pairs
            ctl["self_pairs"], sim = self_collision_pairs(model, tables), dict(max_self_contacts=4)
        phys = PhysOracle(model, ctl, terrain, dtype=np.float32, sim=sim)
        self.env = EnvOracle(cfg, tables, consts, phys, terrain)
        self.env.root_states[:, :3] = torch.as_tensor(consts["env_origins"], dtype=torch.float32) + torch.tensor([0.0, 0.0, 0.95])
        self.env.dof_pos[:] = self.env.default_dof_pos
        self.g = torch.Generator().manual_seed(0)
        tc = make_train_cfg()["algorithm"]
        torch.manual_seed(1)
        self.ac = pa.ActorCritic(O, P, D)
        self.ppo = pa.PPOStep(self.ac, clip=tc["clip_param"], vcoef=tc["value_loss_coef"], ecoef=tc["entropy_coef"], lr=tc["learning_rate"],
                              lr_min=tc["learning_rate_min"], lr_max=tc["learning_rate_max"], desired_kl=tc["desired_kl"], max_grad_norm=tc["max_grad_norm"])
        self.M = N * T_STEPS // 25
        self.n_pol, self.n_mb = T_STEPS // self.SUB, 200 // self.SUB
        self.obs = torch.zeros(N, O)
        self.cobs = torch.zeros(N, P)
        self.store = dict(obs=torch.zeros(self.n_pol, N, O), critic_obs=torch.zeros(self.n_pol, N, P), actions=torch.zeros(self.n_pol, N, D),
                          values=torch.zeros(self.n_pol, N, 1), lp=torch.zeros(self.n_pol, N, 1), mu=torch.zeros(self.n_pol, N, D),
                          sigma=torch.zeros(self.n_pol, N, D), rew=torch.zeros(self.n_pol, N, 1))

    def step(self):
        """8 policy steps x N robots + 25 minibatches of N*64/25 rows; returns seconds."""
        torch, st = self.torch, self.store
        t0 = time.perf_counter()
        for s in range(self.n_pol):
            a, v, lp, mu, sg = self.ppo.act(self.obs, self.cobs)
            st["obs"][s], st["critic_obs"][s], st["actions"][s], st["values"][s], st["lp"][s, :, 0], st["mu"][s], st["sigma"][s] = self.obs, self.cobs, a, v, lp, mu, sg
            obs, pri, rew, reset, _ = self.env.step(a, torch.rand(self.N, self.K, generator=self.g), 5.0)
            self.obs, self.cobs = obs.clone(), pri.clone()
            st["rew"][s, :, 0] = rew
        # minibatch rows: the sample's own transitions, re-drawn with replacement up to the full minibatch size M
        flat = {k: v.flatten(0, 1) for k, v in st.items()}
        adv = flat["rew"] - flat["values"]
        adv = (adv - adv.mean()) / (adv.std() + 1e-8)
        ret = flat["rew"] + 0.99 * flat["values"]
        n = flat["obs"].shape[0]
        for _ in range(self.n_mb):
            sel = torch.randint(0, n, (self.M,), generator=self.g)
            self.ppo.minibatch(dict(obs=flat["obs"][sel], critic_obs=flat["critic_obs"][sel], actions=flat["actions"][sel], values=flat["values"][sel],
                                    advantages=adv[sel], returns=ret[sel], old_log_prob=flat["lp"][sel], old_mu=flat["mu"][sel], old_sigma=flat["sigma"][sel]))
        return time.perf_counter() - t0

    def describe(self, k):
        return (f"{k} steps x (8 policy steps x {self.N} robots [reference task arithmetic in torch CPU over the C physics oracle, OpenMP] + "
                f"25 PPO minibatches x {self.M} rows [nn.Module + autograd + torch.optim.Adam as rsl_rl]) = {k / self.SUB:.2f} whole iterations "
                f"of the same config; every step really timed, none projected")


def cpu_baseline(config, budget_s=25.0):
    arm = CpuArm(config)
    arm.step()                                    # warm-up (allocator, thread pools)
    ts = []
    while sum(ts) < budget_s and len(ts) < 16:
        ts.append(arm.step())
    per = sum(ts) / len(ts)
    return dict(value=arm.N * (T_STEPS // arm.SUB) / per, unit=UNIT, cores=arm.cores, kind="port", sample=arm.describe(len(ts)))


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    t_wall = time.perf_counter()
    arm = CpuArm(args.config)
    for _ in range(min(args.warmup, 2)):          # thread pools / allocator; more warm-up buys nothing on a CPU
        arm.step()
    ts = [arm.step() for _ in range(args.steps)]
    total = sum(ts)
    steps_per_sample = arm.N * (T_STEPS // arm.SUB)
    v = steps_per_sample * args.steps / total
    cb = dict(value=v, unit=UNIT, cores=arm.cores, kind="port", sample=arm.describe(args.steps))
    line = {"impl": "reference", "metric": metric_for(args.config), "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": total / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload(args.config), "envs_total": CONFIGS[args.config]["envs"],
                       "reference_step": "1/8 PPO iteration (8 of 64 policy steps + 25 of 200 minibatches): same env-steps/s as whole iterations",
                       "note": "CPU port of the reference Python (Isaac Gym binaries cannot run here: CPython <= 3.8, closed-source PhysX); "
                               "physics = our C spec on all host cores"},
            "cpu_baseline": cb, "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0, "wall_s": time.perf_counter() - t_wall}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------------
# CUDA arm
# ---------------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    from grx_b200 import _lib as L
    from grx_b200.config import make_cfg, make_train_cfg
    from grx_b200.env import GRXVecEnv
    from grx_b200.runner import OnPolicyRunner
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = f"cuda:{local}"
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("NCCL_DEBUG", "WARN")   # keep stdout to the one JSON line (NCCL prints its version banner there otherwise)
        dist.init_process_group("nccl", device_id=torch.device(dev))
    cf = CONFIGS[args.config]
    n_per_gpu = cf["envs"]
    n_total = n_per_gpu * world
    torch.manual_seed(1)
    cfg = make_task_cfg(args.config, n_total)
    env = GRXVecEnv(cfg, sim_device=dev, rank=rank, world_size=world)
    tc = make_train_cfg()
    runner = OnPolicyRunner(env, tc, log_dir=None, device=dev, world_size=world)
    alg = runner.algorithm
    lib = L.lib()
    runner.learn(1, init_at_random_ep_len=True)       # first iteration also builds the CUDA graph of the update

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # instrument the env launch with CUDA events on the launching stream (for the roofline of the env kernel)
    ev_pairs = []
    orig_step = env.step

    def timed_step(actions):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        out = orig_step(actions)
        b.record()
        ev_pairs.append((a, b))
        return out
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    for _ in range(args.warmup):
        runner.learn(1)
    barrier()
    env.step = timed_step
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    phase = []
    barrier()
    launches0 = int(lib.grx_debug_launch_count())
    clocks.begin()
    e0.record()
    for _ in range(args.steps):
        runner.learn(1)
        phase.append(dict(runner.last_timing))
    e1.record()
    barrier()
    clocks.end()
    launches = int(lib.grx_debug_launch_count()) - launches0
    ms = e0.elapsed_time(e1)
    clk = clocks.stop() if rank == 0 else None
    env.step = orig_step
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t)
    env_ms = sum(a.elapsed_time(b) for a, b in ev_pairs) / len(ev_pairs)
    coll = sum(p["collection_time"] for p in phase) / len(phase) * 1e3
    learn = sum(p["learn_time"] for p in phase) / len(phase) * 1e3
    value = n_total * T_STEPS * args.steps / (ms / 1e3)

    # ---- replica health after the timed region: no peer-flag time-out, and every rank holds bit-identical parameters
    alg.check_comm(wait=True)
    comm_error = int(alg.ctl[17:18].view(torch.int32))
    replicas_identical = True
    if world > 1:
        pi = alg.params.view(torch.int32).to(torch.int64)
        sig = torch.stack([pi.sum(), (pi * torch.arange(1, pi.numel() + 1, device=dev)).sum()])   # two checksums of the bit patterns
        sigs = [torch.zeros_like(sig) for _ in range(world)]
        dist.all_gather(sigs, sig)
        replicas_identical = all(bool(torch.equal(s, sigs[0])) for s in sigs)
        ce = torch.tensor([comm_error], device=dev)
        dist.all_reduce(ce, op=dist.ReduceOp.MAX)
        comm_error = int(ce)

    # ---- e2e: the same iteration through the public API with HOST buffers inside the timed region: the minibatch permutation comes
    # from pinned host memory every iteration, per policy step the rewards / dones go back to pinned host memory (what the reference
    # runner's bookkeeping reads, on_policy_runner.py:171-181), and the two mean losses are read at the end of the iteration.
    nidx = alg._indices.numel()
    h_idx = torch.randperm(nidx).pin_memory()
    h_rew = torch.empty(T_STEPS, env.num_envs, dtype=torch.float32).pin_memory()
    h_done = torch.empty(T_STEPS, env.num_envs, dtype=torch.uint8).pin_memory()
    obs, cobs = env.get_observations(), env.get_privileged_observations()
    k_e2e = max(2, min(args.steps, 5))
    barrier()
    t0 = time.perf_counter()
    for _ in range(k_e2e):
        for s in range(T_STEPS):
            actions = alg.act(obs, cobs)
            obs, cobs, rew, dones, infos = env.step(actions)
            alg.process_env_step(rew, dones, infos)
            h_rew[s].copy_(rew, non_blocking=True)
            h_done[s].copy_(dones.view(torch.uint8), non_blocking=True)
        alg.compute_returns(cobs)
        d_idx = h_idx.to(dev, non_blocking=True)
        mvl, msl = alg.update(indices=d_idx)
        alg.clear_storage()
        losses = (float(mvl), float(msl))           # D2H read of the iteration's result (syncs)
    barrier()
    e2e_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_s], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t)
    e2e_val = n_total * T_STEPS * k_e2e / e2e_s

    # ---- second end-to-end figure: the env alone through the host-buffer C entry grx_env_step_host (H2D actions, one fused step,
    # D2H obs + privileged obs + rewards + resets, stream sync) — what a non-torch host of the reference's VecEnv.step() would call
    import ctypes as C
    import numpy as np
    n_host = 20
    h_act = torch.zeros(env.num_envs, env.num_actions).pin_memory()
    h_o, h_p = torch.empty(env.num_envs, env.num_obs).pin_memory(), torch.empty(env.num_envs, env.num_pri_obs).pin_memory()
    h_r, h_d = torch.empty(env.num_envs).pin_memory(), torch.empty(env.num_envs, dtype=torch.uint8).pin_memory()
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    barrier()
    t0 = time.perf_counter()
    for i in range(n_host):
        L.check(lib.grx_env_step_host(env._h, C.c_void_p(h_act.data_ptr()), C.c_float(5.0), 0, C.c_uint64(10 ** 6 + i), C.c_void_p(h_o.data_ptr()),
                                      C.c_void_p(h_p.data_ptr()), C.c_void_p(h_r.data_ptr()), C.c_void_p(h_d.data_ptr()), st))
    host_s = time.perf_counter() - t0
    env_host = {"value": env.num_envs * world * n_host / host_s, "unit": "env-steps/s (env only, host buffers, grx_env_step_host)",
                "h2d_bytes_per_step": env.num_envs * env.num_actions * 4,
                "d2h_bytes_per_step": env.num_envs * ((env.num_obs + env.num_pri_obs + 1) * 4 + 1), "us_per_step": host_s / n_host * 1e6}

    if rank == 0:
        hbm, tf_sus, which = peaks()
        traffic = measured_traffic()
        _, _, _, env_bytes, fwd_flop = shapes(args.config)
        env_kernel = "envg_step_kernel (generic topology)" if CONFIGS[args.config].get("full_body") else "env_step_kernel"
        env_achieved = env_bytes * env.num_envs / (env_ms * 1e-3) / 1e9
        # dense layers of the update only: 8 epochs x (forward + backward ~ 3 x forward) per transition; the rollout's policy forward runs
        # in the collection phase and is in neither the numerator nor the time
        ppo_flops = env.num_envs * T_STEPS * fwd_flop * 8 * 3
        ppo_tf = ppo_flops / (learn * 1e-3) / 1e12
        dominant_env = coll >= learn
        roof_env = {"kernel": env_kernel, "bound": "hbm", "achieved": env_achieved, "peak": hbm, "unit": "GB/s", "frac": env_achieved / hbm,
                    "traffic": traffic.get("env_step_kernel") if env_kernel == "env_step_kernel" else None, "peak_source": which, "us_per_launch": env_ms * 1e3,
                    "note": f"{env_bytes} algorithmic B per env-step x {env.num_envs} robots per launch.  The kernel is issue/latency-bound "
                            "(10 substeps of articulated dynamics per launch, ~55k warp-instructions per env-step), not bandwidth-bound: the "
                            "HBM fraction is reported because BASELINE asks for it, the meaningful figure is us_per_launch"}
        roof_ppo = {"kernel": "PPO update dense layers (tcgen05 kind::tf32, 200 minibatches)", "bound": "tensor", "achieved": ppo_tf, "peak": tf_sus,
                    "unit": "TFLOP/s", "frac": ppo_tf / tf_sus, "traffic": traffic.get("ppo_update_per_minibatch") if args.config != 6 else None, "peak_source": which,
                    "note": f"numerator = 8 epochs x 3 x {fwd_flop} FLOP per transition (update only); time = compute_returns + update (learn_ms). "
                            "Peak = measured bf16 sustained (MEASURED_PEAKS.json); the layers run kind::tf32, whose tensor peak is half of it. "
                            "traffic = DRAM bytes of the kernels of ONE minibatch from the ncu --set full capture under profiles/"}
        cb = cpu_baseline(args.config) if world == 1 else None
        line = {"metric": metric_for(args.config), "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "tf32+f32",
                "data": "synthetic",
                "config": {"workload": workload(args.config), "baseline_config": args.config,
                           "dtype_note": "dense layers tcgen05 kind::tf32 (fp32 storage, fp32 accumulate); env kernel, heads, losses, GAE, Adam fp32",
                           "envs_total": n_total, "parallelism": f"dp{world} (env shards by index; per minibatch one all-reduce of 436893 floats over NVLink peer memory, "
                                                          "fused with the gradient-norm reduction inside the update's CUDA graph)",
                           "l2": f"per-iteration working set (rollout storage {env.num_envs * T_STEPS * 961 / 1e6:.0f} MB/GPU) exceeds the 126 MB L2; no explicit flush",
                           "collection_ms": coll, "learn_ms": learn},
                "roofline": roof_env if dominant_env else roof_ppo, "roofline_env": roof_env, "roofline_ppo": roof_ppo,
                "cpu_baseline": cb, "clocks": clk,
                "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": nidx * 8, "d2h_bytes_per_step": T_STEPS * env.num_envs * 5 + 8,
                        "iterations": k_e2e, "last_losses": losses, "env_step_host": env_host},
                "comm_error": comm_error, "replicas_identical": replicas_identical,
                "gpu_launches": launches}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--config", type=int, default=2, choices=sorted(CONFIGS))
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
        return
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.gpus > 1 and world == 1:   # convenience: relaunch under torchrun
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}", "--master-addr", "127.0.0.1",
               "--master-port", "29511", os.path.abspath(__file__), "--gpus", str(args.gpus), "--steps", str(args.steps), "--warmup", str(args.warmup),
               "--config", str(args.config)]
        sys.exit(subprocess.call(cmd))
    run_ours(args)


if __name__ == "__main__":
    main()
