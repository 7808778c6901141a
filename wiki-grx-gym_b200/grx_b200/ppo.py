"""PPO + ActorCriticMLP over the sm_100a kernels of libgrx_b200.so, behind the reference's Python surface:
``rsl_rl.algorithms.PPO`` (rsl_rl/rsl_rl/algorithms/ppo.py:10-333) and ``rsl_rl.modules.ActorCriticMLP``
(modules/actor_critic_mlp.py:10-231).  All tensors are zero-copy views of device memory owned by the library;
no autograd, no torch.optim — forward, loss, backward, clip and Adam are hand-written CUDA (csrc/grx_ppo.cu).

No CPU fallback: constructing these objects without the CUDA library / a CUDA device raises.
"""
from __future__ import annotations

import ctypes as C
import os
from collections import OrderedDict

import numpy as np
import torch

from . import _lib as L
from .env import _DevArray


def _layer_keys(net):
    return [(f"{net}.model.{2 * i}.weight", f"{net}.model.{2 * i}.bias") for i in range(4)]


class _NetView:
    """`actor` / `critic` attribute of the policy (what helpers.export_policy_as_jit and play.py touch): callable forward."""

    def __init__(self, owner, name):
        self._owner, self._name = owner, name

    def __call__(self, x):
        if self._name != "actor":
            raise NotImplementedError("critic forward is fused into PPO.act / compute_returns")
        return self._owner.act_inference(x)

    def train(self, mode=True):
        return self

    def eval(self):
        return self

    def state_dict(self):
        return OrderedDict((k[len(self._name) + 1:], v) for k, v in self._owner.state_dict().items() if k.startswith(self._name + "."))

    def to_module(self):
        """A plain ``torch.nn.Module`` with the reference MLP's structure (``model = Sequential(Linear, ELU, ..., Linear)``,
        rsl_rl/modules/mlp.py:22-42) holding a CPU copy of the current weights — what deployment code expects."""
        return _TorchMLP(self.state_dict())

    def __deepcopy__(self, memo):
        # helpers.export_policy_as_jit does copy.deepcopy(actor_critic.actor).to('cpu') and torch.jit.script()s the result
        # (legged_gym/utils/helpers.py:188-201): hand it a real nn.Module
        return self.to_module()


class _TorchMLP(torch.nn.Module):
    def __init__(self, sd):
        super().__init__()
        keys = sorted({int(k.split(".")[1]) for k in sd})
        layers = []
        for n, i in enumerate(keys):
            w = sd[f"model.{i}.weight"]
            lin = torch.nn.Linear(w.shape[1], w.shape[0])
            with torch.no_grad():
                lin.weight.copy_(w.detach().cpu())
                lin.bias.copy_(sd[f"model.{i}.bias"].detach().cpu())
            layers.append(lin)
            if n < len(keys) - 1:
                layers.append(torch.nn.ELU())
        self.model = torch.nn.Sequential(*layers)

    def forward(self, x):
        return self.model(x)


class ActorCriticMLP:
    """Parameters of actor 39->512->256->128->10 / critic 168->512->256->128->1 (ELU) + per-action ``std`` as views into
    one flat device vector laid out in the reference's state_dict order (SURVEY.md §5 checkpoint row)."""

    def __init__(self, actor_num_input, critic_num_input, actor_num_output, actor_hidden_dims=(512, 256, 128),
                 critic_hidden_dims=(512, 256, 128), activation="elu", fixed_std=False, init_noise_std=1.0,
                 set_std=True, set_noise_std=1.0, actor_output_activation=None, critic_output_activation=None, **kwargs):
        # positional order == actor_critic_mlp.py:12-25
        if activation != "elu" or len(actor_hidden_dims) != 3 or len(critic_hidden_dims) != 3:
            raise L.GrxError("the CUDA policy implements the registered architecture: 3 hidden layers, ELU")
        if fixed_std:
            raise L.GrxError("fixed_std=True is not implemented (registered tasks use fixed_std=False, gr1t1_config.py:345)")
        if actor_output_activation is not None or critic_output_activation is not None:
            raise L.GrxError("output activations are not implemented (registered tasks use None)")
        if kwargs:   # actor_critic_mlp.py:45-47 prints and ignores
            import warnings
            warnings.warn("ActorCritic.__init__ got unexpected arguments, which will be ignored: " + str(list(kwargs.keys())))
        self.num_actor_input, self.num_critic_input, self.num_actor_output = actor_num_input, critic_num_input, actor_num_output
        self.actor_hidden_dims, self.critic_hidden_dims = list(actor_hidden_dims), list(critic_hidden_dims)
        self.init_noise_std, self.set_std, self.set_noise_std = init_noise_std, set_std, set_noise_std
        # same generator call order as the reference constructor (actor layers, critic layers): nn.Linear default init
        host = OrderedDict()
        host["std"] = init_noise_std * torch.ones(actor_num_output)                     # ACM:79-82
        for net, dims in (("actor", [actor_num_input, *actor_hidden_dims, actor_num_output]),
                          ("critic", [critic_num_input, *critic_hidden_dims, 1])):
            for i, (wk, bk) in enumerate(_layer_keys(net)):
                lin = torch.nn.Linear(dims[i], dims[i + 1])                              # mlp.py:26-31
                host[wk], host[bk] = lin.weight.detach().clone(), lin.bias.detach().clone()
        self._host_init = host
        self._views = None          # bound by PPO (needs the library object)
        self.actor, self.critic = _NetView(self, "actor"), _NetView(self, "critic")
        self.distribution = None

    # ---- bound once PPO created the device buffers
    def _bind(self, ppo, flat):
        self._ppo = ppo
        # flat layout of csrc/grx_ppo.cu: reference state_dict order; every tensor starts 16-byte aligned and weight matrices are
        # stored with their rows padded to a multiple of 4 floats (TMA / vector loads) — the views hide the padding
        self._slices, off = OrderedDict(), 0
        for k, v in self._host_init.items():
            off = (off + 3) // 4 * 4
            if v.dim() == 2:
                ld = (v.shape[1] + 3) // 4 * 4
                self._slices[k] = (off, v.shape[0], v.shape[1], ld)
                off += v.shape[0] * ld
            else:
                self._slices[k] = (off, v.numel())
                off += v.numel()
        assert (off + 3) // 4 * 4 == flat.numel(), (off, flat.numel())
        self._views = OrderedDict((k, self.view_of(flat, k)) for k in self._host_init)
        self.load_state_dict(self._host_init, set_std=False)
        self._host_init = None

    def view_of(self, flat, key):
        """The tensor `key` inside a flat vector laid out like the parameter vector (params, grads, Adam moments)."""
        sl = self._slices[key]
        if len(sl) == 2:
            return flat[sl[0]:sl[0] + sl[1]]
        off, rows, cols, ld = sl
        return flat[off:off + rows * ld].view(rows, ld)[:, :cols]

    @property
    def std(self):
        return self._views["std"]

    def to(self, device):
        return self

    def train(self, mode=True):
        return self

    def eval(self):
        return self

    def parameters(self):
        return list(self._views.values())

    def state_dict(self):
        return OrderedDict((k, v) for k, v in self._views.items())

    def load_state_dict(self, state_dict, strict=True, set_std=None):
        """actor_critic_mlp.py:116-134: after loading, ``std`` is overwritten with ``set_noise_std`` unless set_std is False."""
        if strict and set(state_dict.keys()) != set(self._views.keys()):
            raise KeyError(f"state_dict keys mismatch: {sorted(set(state_dict) ^ set(self._views))}")
        for k, v in state_dict.items():
            if k in self._views:
                self._views[k].copy_(torch.as_tensor(v).to(self._views[k].device, torch.float32))
        if self.set_std if set_std is None else set_std:
            self._views["std"].fill_(self.set_noise_std)
        return True

    def act_inference(self, observations):                                              # ACM:209-217
        return self._ppo.act_inference(observations)


class PPO:
    """Same constructor arguments and methods as rsl_rl's PPO (ppo.py:24-110)."""

    def __init__(self, actor_critic, num_learning_epochs=1, num_mini_batches=1, clip_param=0.2, gamma=0.998, lam=0.95,
                 value_loss_coef=1.0, entropy_coef=0.0, learning_rate=1e-3, learning_rate_min=1e-5, learning_rate_max=1e-2,
                 max_grad_norm=1.0, use_clipped_value_loss=True, schedule="fixed", desired_kl=0.01, device="cuda:0",
                 storage_class="RolloutStorage", use_tensor_cores=1, world_size=1, process_group=None, seed=1, env_id_offset=0,
                 comm_timeout_ms=0, weight_decay=0.0, **kwargs):
        """seed / env_id_offset: the task seed and the global index of this rank's env 0 — the action-noise stream of ``act`` is keyed by
        (seed, global env id, step), so W ranks draw exactly what one GPU holding all envs would."""
        if not torch.cuda.is_available():
            raise L.GrxError("PPO needs a CUDA device (no CPU fallback)")
        if weight_decay:
            raise L.GrxError("weight_decay != 0 is not implemented by the fused Adam kernel (the registered tasks use 0, ppo.py:81)")
        if kwargs:   # ppo.py:98-102 prints the unexpected arguments and ignores them; do the same, loudly
            import warnings
            warnings.warn("PPO.__init__ got unexpected arguments, which will be ignored: " + str(list(kwargs.keys())))
        self.seed, self.env_id_offset, self.comm_timeout_ms = int(seed), int(env_id_offset), int(comm_timeout_ms)
        self.lib = L.lib()
        self.device = torch.device(device)
        self.actor_critic = actor_critic
        self.num_learning_epochs, self.num_mini_batches = num_learning_epochs, num_mini_batches
        self.clip_param, self.gamma, self.lam = clip_param, gamma, lam
        self.value_loss_coef, self.entropy_coef = value_loss_coef, entropy_coef
        self.learning_rate_init = learning_rate
        self.learning_rate_min, self.learning_rate_max, self.max_grad_norm = learning_rate_min, learning_rate_max, max_grad_norm
        self.use_clipped_value_loss, self.schedule, self.desired_kl = use_clipped_value_loss, schedule, desired_kl
        self.use_tensor_cores, self.world_size, self.process_group = use_tensor_cores, world_size, process_group
        self.symmetry_coef = 0                                                          # ppo.py:96 (never enabled in this fork)
        self._h = None
        self.storage = None
        self.step = 0
        self._act_counter = 0

    # ------------------------------------------------------------------ construction
    def init_storage(self, num_envs, num_transitions_per_env, *unused):                 # ppo.py:112-134
        ac = self.actor_critic
        c = L.PPOCfg()
        c.num_envs, c.num_steps = num_envs, num_transitions_per_env
        c.num_obs, c.num_pri_obs, c.num_actions = ac.num_actor_input, ac.num_critic_input, ac.num_actor_output
        for i in range(3):
            c.actor_hidden[i], c.critic_hidden[i] = ac.actor_hidden_dims[i], ac.critic_hidden_dims[i]
        c.num_learning_epochs, c.num_mini_batches = self.num_learning_epochs, self.num_mini_batches
        c.clip_param, c.gamma, c.lam = self.clip_param, self.gamma, self.lam
        c.value_loss_coef, c.entropy_coef = self.value_loss_coef, self.entropy_coef
        c.learning_rate, c.learning_rate_min, c.learning_rate_max = self.learning_rate_init, self.learning_rate_min, self.learning_rate_max
        c.desired_kl = self.desired_kl if self.desired_kl is not None else 0.0
        c.max_grad_norm = self.max_grad_norm
        c.adaptive_schedule = int(self.desired_kl is not None and self.schedule == "adaptive")   # ppo.py:253
        c.use_clipped_value_loss = int(self.use_clipped_value_loss)
        c.init_noise_std, c.use_tensor_cores, c.world_size = ac.init_noise_std, int(self.use_tensor_cores), self.world_size
        c.seed, c.env_id_offset, c.comm_timeout_ms = self.seed, self.env_id_offset, self.comm_timeout_ms
        self._cfg = c
        self._h = C.c_void_p()
        dev = self.device.index if self.device.index is not None else torch.cuda.current_device()
        L.check(self.lib.grx_ppo_create(C.byref(c), dev, C.byref(self._h)))
        self.num_envs, self.num_steps = num_envs, num_transitions_per_env
        v = self._view
        self.params, self.grads, self.reduce_buf = v("params"), v("grads"), v("reduce_buf")
        self.adam_m, self.adam_v = v("adam_m"), v("adam_v")
        self.ctl = v("ctl")
        self.mb_log = v("mb_log")       # [epochs * minibatches, 4]: (mean KL, lr, loss, grad norm) per minibatch of the last update
        self.storage = _Storage({k: v(k) for k in ("obs", "critic_obs", "actions", "values", "rewards", "dones",
                                                   "actions_log_prob", "mu", "sigma", "returns", "advantages")})
        self._moments = v("adv_moments").view(torch.float64)
        ac._bind(self, self.params)
        self._actions = torch.zeros(num_envs, c.num_actions, device=self.device)
        self.mini_batch_size = (num_envs * num_transitions_per_env) // self.num_mini_batches
        self._indices = torch.zeros(self.num_mini_batches * self.mini_batch_size, dtype=torch.int64, device=self.device)
        self._comm = False
        self._err_host = torch.zeros(1, dtype=torch.int32).pin_memory()
        self._err_event = None
        # world_size > 1 with an initialised process group: we own the collectives (weights broadcast now; per minibatch the in-graph
        # NVLink all-reduce, or NCCL with GRX_COMM=nccl).  Without one the caller drives them through the stepwise entries.
        self._dist = self.world_size > 1 and torch.distributed.is_available() and torch.distributed.is_initialized()
        if self._dist:
            self.broadcast_parameters()
            if os.environ.get("GRX_COMM", "nvlink") == "nvlink":
                self._open_comm()

    def _open_comm(self):
        """Map every peer's gradient block (cudaIpc over NVLink) so the per-minibatch all-reduce is one of OUR kernels inside the
        update's CUDA graph instead of an NCCL call from the host loop.  GRX_COMM=nccl keeps the NCCL path.  If any rank cannot
        export / map the blocks (no peer access, IPC disabled in the container) every rank falls back to the NCCL path together."""
        import torch.distributed as dist
        import warnings
        rank = dist.get_rank(self.process_group)
        h = (C.c_char * 64)()
        ok = self.lib.grx_ppo_comm_handle(self._h, h) == 0
        err = "" if ok else self.lib.grx_last_error().decode()
        gathered = [None] * self.world_size
        dist.all_gather_object(gathered, (ok, bytes(h.raw)), group=self.process_group)
        if all(g[0] for g in gathered):
            if self.lib.grx_ppo_comm_open(self._h, rank, self.world_size, b"".join(g[1] for g in gathered)) != 0:
                ok, err = False, self.lib.grx_last_error().decode()
        else:
            ok = False
        flag = torch.tensor([1 if ok else 0], device=self.device, dtype=torch.int32)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=self.process_group)    # also the barrier: every rank has mapped its peers
        self._comm = bool(int(flag) == 1)
        if not self._comm:
            if ok:   # some other rank failed: drop our mappings' use (the C side keeps them until destroy; the graph path is simply not taken)
                err = "a peer rank could not map the gradient blocks"
            warnings.warn(f"grx_b200: NVLink peer-memory all-reduce unavailable ({err}); using the NCCL all-reduce path")

    def _view(self, name):
        b = L.Buffer()
        L.check(self.lib.grx_ppo_get_buffer(self._h, name.encode(), C.byref(b)))
        if b.dtype == L.DT_U64:   # u64 bit patterns of doubles -> int32 pairs, re-viewed as float64 by the caller
            b.dtype = L.DT_I32
            b.dims[0] *= 2
        return torch.as_tensor(_DevArray(b, self), device=self.device)

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def close(self):
        if self._h is not None and self._h.value:
            self.lib.grx_ppo_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def test_mode(self):
        pass

    def train_mode(self):
        pass

    @property
    def learning_rate(self):
        return float(self.ctl[0])          # host read (sync) — logging only

    @learning_rate.setter
    def learning_rate(self, v):
        self.ctl[0] = float(v)

    @property
    def adam_step(self):
        return int(self.ctl[3:4].view(torch.int32))

    # ------------------------------------------------------------------ rollout
    def act(self, obs, critic_obs, eps=None):                                           # ppo.py:144-171
        t = self.step
        self._act_counter += 1
        L.check(self.lib.grx_ppo_act(self._h, C.c_void_p(obs.data_ptr()), C.c_void_p(critic_obs.data_ptr()),
                                     C.c_void_p(eps.data_ptr()) if eps is not None else None, t, C.c_void_p(self._actions.data_ptr()),
                                     C.c_uint64(self._act_counter), self._stream()))
        return self._actions

    def process_env_step(self, rewards, dones, infos):                                  # ppo.py:177-196
        to = infos.get("time_outs") if isinstance(infos, dict) else None
        d8 = dones if dones.dtype == torch.uint8 else dones.view(torch.uint8) if dones.dtype == torch.bool else dones.to(torch.uint8)
        t8 = None
        if to is not None:
            t8 = to if to.dtype == torch.uint8 else to.view(torch.uint8) if to.dtype == torch.bool else to.to(torch.uint8)
        L.check(self.lib.grx_ppo_process_env_step(self._h, C.c_void_p(rewards.data_ptr()), C.c_void_p(d8.data_ptr()),
                                                  C.c_void_p(t8.data_ptr()) if t8 is not None else None, self.step, self._stream()))
        self.step += 1

    def compute_returns(self, last_critic_obs):                                         # ppo.py:198-205
        p = C.c_void_p(last_critic_obs.data_ptr())
        if self.world_size == 1:
            L.check(self.lib.grx_ppo_compute_returns(self._h, p, self._stream()))
        elif not self._dist:
            raise L.GrxError("world_size > 1 without torch.distributed: call grx_ppo_compute_returns_local, all-reduce `adv_moments`, "
                             "then grx_ppo_normalize_advantages yourself")
        else:   # global advantage statistics: one 3-double all-reduce per iteration
            L.check(self.lib.grx_ppo_compute_returns_local(self._h, p, self._stream()))
            torch.distributed.all_reduce(self._moments[:3], group=self.process_group)
            L.check(self.lib.grx_ppo_normalize_advantages(self._h, self._stream()))

    def clear_storage(self):                                                            # base_storage.py:115-116
        self.step = 0

    # ------------------------------------------------------------------ update
    def draw_indices(self):
        """rollout_storage.py:71-75: ONE permutation of mini_batches*mini_batch_size reused for every epoch."""
        n = self.num_mini_batches * self.mini_batch_size
        self._indices.copy_(torch.randperm(n, device=self.device))
        return self._indices

    def check_comm(self, wait=False):
        """Raise if a peer-flag / grid-barrier wait of the in-graph NVLink all-reduce timed out (ctl.comm_error, sticky: the device
        skips every optimiser step from then on, so the replicas never diverge silently).  The flag travels to pinned host memory
        asynchronously after every update; without ``wait`` only an already finished copy is inspected (no host sync)."""
        if self._err_event is None:
            return
        if wait:
            self._err_event.synchronize()
        if self._err_event.query() and int(self._err_host[0]) != 0:
            code = int(self._err_host[0])
            raise L.GrxError(f"grx_b200: multi-GPU gradient all-reduce failed (comm_error={code}: "
                             + ("a peer rank did not arrive within comm_timeout_ms" if code == 1 else "apply grid barrier timed out")
                             + "); optimiser steps were skipped from that minibatch on — restart from the last checkpoint")

    def update(self, indices=None):                                                     # ppo.py:215-321
        self.check_comm()
        if indices is None:
            indices = self.draw_indices()
        else:
            self._indices.copy_(indices.to(self.device, torch.int64))
        idx = C.c_void_p(self._indices.data_ptr())
        if self.world_size == 1 or self._comm:
            L.check(self.lib.grx_ppo_update(self._h, idx, self._stream()))
        elif not self._dist:
            raise L.GrxError("world_size > 1 without torch.distributed: drive grx_ppo_minibatch_grads / all-reduce(reduce_buf) / "
                             "grx_ppo_minibatch_apply yourself")
        else:
            self.ctl[11:14].zero_()
            for _ in range(self.num_learning_epochs):
                for mb in range(self.num_mini_batches):
                    L.check(self.lib.grx_ppo_minibatch_grads(self._h, idx, mb, self._stream()))
                    torch.distributed.all_reduce(self.reduce_buf, group=self.process_group)   # ONE all-reduce per minibatch
                    L.check(self.lib.grx_ppo_minibatch_apply(self._h, self._stream()))
        if self.world_size > 1:   # comm_error -> pinned host memory, inspected by the next check_comm() (no sync here)
            self._err_host.copy_(self.ctl[17:18].view(torch.int32), non_blocking=True)
            self._err_event = torch.cuda.Event()
            self._err_event.record()
        n = self.num_learning_epochs * self.num_mini_batches
        self.last_losses = self.ctl[11:13] / n                                          # device tensor; ppo.py:314-316
        return _LazyFloat(self.last_losses, 0), _LazyFloat(self.last_losses, 1)

    def minibatch_stats(self):
        c = self.ctl.cpu()
        return dict(lr=float(c[0]), skip=int(c[2:3].view(torch.int32)), step=int(c[3:4].view(torch.int32)), kl=float(c[6]),
                    loss=float(c[7]), value_loss=float(c[8]), surrogate_loss=float(c[9]), grad_norm=float(c[10]),
                    comm_error=int(c[17:18].view(torch.int32)),   # != 0: a peer flag / grid barrier wait timed out (results invalid)
                    chain_error=int(c[21:22].view(torch.int32)))  # != 0: a barrier wait inside a chained-layer kernel timed out (protocol bug)

    def act_inference(self, obs):
        obs = obs.to(self.device, torch.float32).contiguous()
        out = torch.empty(obs.shape[0], self._cfg.num_actions, device=self.device)
        L.check(self.lib.grx_ppo_act_inference(self._h, C.c_void_p(obs.data_ptr()), obs.shape[0], C.c_void_p(out.data_ptr()), self._stream()))
        return out

    def broadcast_parameters(self, src=0):
        torch.distributed.broadcast(self.params, src=src, group=self.process_group)

    # ---- optimizer state in torch.optim.Adam's state_dict layout (on_policy_runner.py:297-331)
    def optimizer_state_dict(self):
        state = {}
        step = float(self.adam_step)
        ac = self.actor_critic
        for i, k in enumerate(ac.state_dict()):
            state[i] = {"step": torch.tensor(step), "exp_avg": ac.view_of(self.adam_m, k).clone(),
                        "exp_avg_sq": ac.view_of(self.adam_v, k).clone()}
        group = {"lr": self.learning_rate, "betas": (0.9, 0.999), "eps": 1e-8, "weight_decay": 0, "amsgrad": False,
                 "maximize": False, "foreach": None, "capturable": False, "differentiable": False, "fused": None,
                 "params": list(range(len(state)))}
        # extra top-level key (torch.optim.Optimizer.load_state_dict reads only "state" / "param_groups"): counters of our RNG streams
        return {"state": state, "param_groups": [group], "grx_b200": {"act_counter": int(self._act_counter)}}

    def load_optimizer_state_dict(self, sd):
        ac = self.actor_critic
        for i, k in enumerate(ac.state_dict()):
            if i in sd["state"]:
                ac.view_of(self.adam_m, k).copy_(sd["state"][i]["exp_avg"].to(self.device))
                ac.view_of(self.adam_v, k).copy_(sd["state"][i]["exp_avg_sq"].to(self.device))
                step = int(float(sd["state"][i]["step"]))
        if sd["state"]:
            self.ctl[3:4].view(torch.int32).fill_(step)
            self.ctl[24:28].view(torch.float64).copy_(torch.tensor([0.9 ** step, 0.999 ** step], dtype=torch.float64))   # running beta^step
        self.learning_rate = sd["param_groups"][0]["lr"]
        self._act_counter = int(sd.get("grx_b200", {}).get("act_counter", self._act_counter))


class _Storage:
    """[T, N, .] rollout tensors under the reference's attribute names (base_storage.py:38-78)."""

    def __init__(self, views):
        self.__dict__.update(views)
        self.observations, self.critic_observations = views["obs"], views["critic_obs"]


class _LazyFloat:
    """update() returns the two mean losses; converting them to Python floats would force a device sync every iteration,
    so the conversion is deferred until somebody formats / logs them."""

    def __init__(self, t, i):
        self._t, self._i = t, i

    def __float__(self):
        return float(self._t[self._i])

    def __format__(self, spec):
        return format(float(self), spec)

    def __repr__(self):
        return repr(float(self))
