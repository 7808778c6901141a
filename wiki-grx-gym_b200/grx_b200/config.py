"""Effective configuration of the registered GR1T1 / GR1T2 tasks.

The registered tasks are the lower-limb variants (envs/__init__.py:40-55).  The
attribute names and nesting follow the reference's config classes
(legged_robot_config.py:34-294, gr1t1_config.py:10-345,
gr1t1_lower_limb_config.py:9-116) so that ``GRXVecEnv`` accepts either one of
these objects or the reference's own ``GR1T1LowerLimbCfg()`` instance unchanged.
Values are restated from those files, not imported.
"""
from __future__ import annotations

import copy
import math

import numpy as np

E = math.e  # the reference scales its sigmas with torch.e


class _NS:
    """Attribute bag; nested like the reference's class-in-class configs."""

    def __init__(self, **kw):
        self.__dict__.update(kw)

    def __repr__(self):
        return "NS(" + ", ".join(f"{k}={v!r}" for k, v in self.__dict__.items()) + ")"

    def to_dict(self):
        return {k: (v.to_dict() if isinstance(v, _NS) else v) for k, v in self.__dict__.items()}


_DEG30 = math.radians(30.0)
_ACT_MAX = np.array([0.79, 0.7, 0.7, 1.92, 0.52, 0.09, 0.7, 0.7, 1.92, 0.52])
_ACT_MIN = np.array([-0.09, -0.7, -1.75, -0.09, -1.05, -0.79, -0.7, -1.75, -0.09, -1.05])


def _stiffness():
    # gr1t1_lower_limb_config.py:20-28
    return {"hip_roll": 48 / _DEG30, "hip_yaw": 66 / _DEG30, "hip_pitch": 130 / _DEG30,
            "knee_pitch": 130 / _DEG30, "ankle_pitch": 15 / _DEG30}


def make_cfg(robot="GR1T1", num_envs=4096, mesh_type="plane"):
    """Config of task ``robot`` ('GR1T1' or 'GR1T2'; both lower-limb)."""
    assert robot in ("GR1T1", "GR1T2")
    st = _stiffness()
    dm = {k: v / 10 * 0.5 for k, v in st.items()}  # gr1t1_lower_limb_config.py:29-35
    d15, d30 = math.radians(15.0), math.radians(30.0)
    dja = {}
    for side in ("left", "right"):
        dja.update({f"{side}_hip_roll_joint": 0.0, f"{side}_hip_yaw_joint": 0.0, f"{side}_hip_pitch_joint": -d15,
                    f"{side}_knee_pitch_joint": d30, f"{side}_ankle_pitch_joint": -d15, f"{side}_ankle_roll_joint": 0.0})
    cfg = _NS(
        robot=robot,
        seed=1,
        sim=_NS(dt=0.002, substeps=1, gravity=[0.0, 0.0, -9.81], up_axis=1,
                physx=_NS(num_threads=10, solver_type=1, num_position_iterations=4, num_velocity_iterations=0,
                          contact_offset=0.01, rest_offset=0.0, bounce_threshold_velocity=0.5,
                          max_depenetration_velocity=1.0, max_gpu_contact_pairs=2 ** 23,
                          default_buffer_size_multiplier=5, contact_collection=2)),
        env=_NS(num_envs=num_envs, episode_length_s=20, num_obs=39, num_pri_obs=168, num_actions=10,
                env_spacing=3.0, send_timeouts=True),
        terrain=_NS(mesh_type=mesh_type, horizontal_scale=0.1, vertical_scale=0.005, border_size=25,
                    curriculum=True, num_rows=10, num_cols=20, max_init_terrain_level=9,
                    static_friction=1.0, dynamic_friction=1.0, restitution=0.0, measure_heights=True,
                    measured_points_x=[-0.5, -0.4, -0.3, -0.2, -0.1, 0.0, 0.1, 0.2, 0.3, 0.4, 0.5],
                    measured_points_y=[-0.5, -0.4, -0.3, -0.2, -0.1, 0.0, 0.1, 0.2, 0.3, 0.4, 0.5],
                    selected=False, terrain_kwargs=None, terrain_proportions=[0.1, 0.1, 0.35, 0.25, 0.2],
                    slope_treshold=0.75, terrain_length=8.0, terrain_width=8.0),
        asset=_NS(name=robot, torso_name="torso", forehead_name="head_pitch", imu_name="imu", waist_name="waist",
                  upper_arm_name="upper_arm", lower_arm_name="lower_arm", hand_name="hand",
                  thigh_name="thigh", shank_name="shank", foot_name="foot_roll", sole_name="sole",
                  ankle_name="ankle", penalize_contacts_on=[],
                  terminate_after_contacts_on=["imu", "torso", "head_pitch", "waist", "upper_arm", "lower_arm", "hand"],
                  self_collisions=0),
        init_state=_NS(pos=[0.0, 0.0, 0.95], rot=[0.0, 0.0, 0.0, 1.0], lin_vel=[0.0, 0.0, 0.0], ang_vel=[0.0, 0.0, 0.0],
                       default_joint_angles=dja),
        commands=_NS(curriculum=False, num_commands=3, resampling_command_interval_s=10.0, heading_command=False,
                     ranges=_NS(lin_vel_x=[-1.0, 1.0], lin_vel_y=[-0.5, 0.5], ang_vel_yaw=[-1.0, 1.0], heading=[-3.14, 3.14])),
        control=_NS(control_type="P", stiffness=st, damping=dm, action_scale=1.0, decimation=10),
        domain_rand=_NS(randomize_friction=True, friction_range=[0.1, 1.0], randomize_restitution=True,
                        restitution_range=[0.0, 0.5], randomize_base_mass=True, multiply_base_mass_range=[0.9, 1.1],
                        randomize_base_com=True, add_base_com_range_x=[-0.1, 0.1], add_base_com_range_y=[-0.1, 0.1],
                        add_base_com_range_z=[-0.1, 0.1], randomize_motor_strength=True,
                        multiply_motor_strength=[0.9, 1.1], push_robots=True, push_interval_s=10.0,
                        max_push_vel_xy=0.5, randomize_init_dof_pos=True, randomize_init_base_velocity=True),
        rewards=_NS(
            only_positive_rewards=False, base_height_target=0.85, swing_feet_height_target=0.10,
            feet_stumble_ratio=5.0, feet_air_time_target=0.5, feet_land_time_max=1.0, tracking_sigma=1.0,
            soft_dof_pos_limit=0.95, soft_dof_vel_limit=0.95, soft_torque_limit=0.95, max_contact_force=500.0,
            sigma_collision=-1.0 * E, sigma_stand_still=-1.0 * E,
            sigma_cmd_diff_lin_vel_x=-1.0 * E * (1.0 / 0.50), sigma_cmd_diff_lin_vel_y=-1.0 * E * (1.0 / 1.00),
            sigma_cmd_diff_lin_vel_z=-1.0 * E, sigma_cmd_diff_ang_vel_yaw=-1.0 * E * (1.0 / 3.00),
            sigma_cmd_diff_base_height=-10.0 * E, sigma_cmd_diff_base_orient=-20.0,
            sigma_cmd_diff_torso_orient=-20.0, sigma_action_diff=-0.1, sigma_action_diff_diff=-1.0,
            sigma_dof_acc_new=-0.001 * E, sigma_dof_tor_new=-0.01 * E, sigma_dof_tor_ankle_feet_lift_up=-1.0,
            sigma_pose_offset=-0.1, sigma_limits_dof_pos=-1.0, sigma_limits_dof_vel=-10.0,
            sigma_limits_dof_tor=-0.1, sigma_feet_speed_xy_close_to_ground=-10.0, sigma_feet_air_time=-1.0,
            sigma_feet_air_height=-200.0, sigma_feet_air_force=-0.05, sigma_feet_land_time=-1.0,
            sigma_feet_stumble=-1.0,
            scales=_NS(termination=-0.0, collision=-0.0, stand_still=1.0, cmd_diff_lin_vel_x=1.0,
                       cmd_diff_lin_vel_y=0.5, cmd_diff_ang_vel_yaw=0.75, cmd_diff_lin_vel_z=0.25,
                       cmd_diff_base_height=0.5, cmd_diff_base_orient=0.25, cmd_diff_torso_orient=0.5,
                       action_diff=-5.0, action_diff_diff=-1.0, dof_acc_new=-0.25, dof_tor_new=-0.05,
                       dof_tor_ankle_feet_lift_up=-0.5, pose_offset=1.0, limits_dof_pos=-10.0,
                       limits_dof_vel=-5.0, limits_dof_tor=-1.0, feet_speed_xy_close_to_ground=0.5,
                       feet_speed_z_close_to_height_target=0.0, feet_air_time=2.0, feet_air_height=1.5,
                       feet_air_force=1.0, feet_land_time=-1.0, on_the_air=-10.0, feet_stumble=-0.2)),
        noise=_NS(add_noise=True, noise_level=1.0,
                  noise_scales=_NS(action=0.0, lin_vel=0.10, ang_vel=0.05, gravity=0.03, dof_pos=0.04, dof_vel=0.20,
                                   height_measurements=0.05)),
        normalization=_NS(obs_scales=_NS(action=1.0, lin_vel=1.0, ang_vel=1.0, gravity=1.0, dof_pos=1.0, dof_vel=1.0,
                                         height_measurements=5.0),
                          clip_observations=100.0,
                          clip_actions_max=_ACT_MAX + _DEG30, clip_actions_min=_ACT_MIN - _DEG30),
    )
    return cfg


def make_full_body_cfg(robot="GR1T1", num_envs=4096, mesh_type="plane"):
    """The reference's UNREGISTERED full-body configuration (``GR1T1Cfg`` / ``GR1T2Cfg``: gr1t1_config.py:10-307, gr1t2_config.py:7-14): 32 DOF /
    32 actions, PD gains and default angles for legs, waist, head and arms, the full-body termination list, action limits.  As upstream, its
    reward scales are all zero except the inherited ``termination = 0`` (gr1t1_config.py:252-254) and its ``num_obs = 121`` is stale: the
    observation layout of gr1t1.py:281-313 gives 9 + 3 * 32 = 105, which is what this object declares."""
    cfg = make_cfg(robot, num_envs, mesh_type)
    cfg.robot = robot + "_full"
    d15, d30 = math.radians(15.0), math.radians(30.0)
    dja = {}
    for side, sgn in (("left", 1.0), ("right", -1.0)):
        dja.update({f"{side}_hip_roll_joint": 0.0, f"{side}_hip_yaw_joint": 0.0, f"{side}_hip_pitch_joint": -d15, f"{side}_knee_pitch_joint": d30,
                    f"{side}_ankle_pitch_joint": -d15, f"{side}_ankle_roll_joint": 0.0,
                    f"{side}_shoulder_pitch_joint": 0.0, f"{side}_shoulder_roll_joint": 0.2 * sgn, f"{side}_shoulder_yaw_joint": 0.0,
                    f"{side}_elbow_pitch_joint": -0.3, f"{side}_wrist_yaw_joint": 0.0, f"{side}_wrist_roll_joint": 0.0, f"{side}_wrist_pitch_joint": 0.0})
    for j in ("waist_yaw", "waist_pitch", "waist_roll", "head_yaw", "head_pitch", "head_roll"):
        dja[j + "_joint"] = 0.0
    cfg.init_state.default_joint_angles = dja                                           # gr1t1_config.py:93-136
    cfg.control.stiffness = {"hip_roll": 251.625, "hip_yaw": 362.5214, "hip_pitch": 200, "knee_pitch": 200, "ankle_pitch": 10.9805, "ankle_roll": 0.25,
                             "waist_yaw": 362.5214, "waist_pitch": 362.5214, "waist_roll": 362.5214, "head_yaw": 10.0, "head_pitch": 10.0, "head_roll": 10.0,
                             "shoulder_pitch": 92.85, "shoulder_roll": 92.85, "shoulder_yaw": 112.06, "elbow_pitch": 112.06,
                             "wrist_yaw": 10.0, "wrist_roll": 10.0, "wrist_pitch": 10.0}  # gr1t1_config.py:158-168
    cfg.control.damping = {"hip_roll": 14.72, "hip_yaw": 10.0833, "hip_pitch": 11, "knee_pitch": 11, "ankle_pitch": 0.5991, "ankle_roll": 0.01,
                           "waist_yaw": 10.0833, "waist_pitch": 10.0833, "waist_roll": 10.0833, "head_yaw": 1.0, "head_pitch": 1.0, "head_roll": 1.0,
                           "shoulder_pitch": 2.575, "shoulder_roll": 2.575, "shoulder_yaw": 3.1, "elbow_pitch": 3.1,
                           "wrist_yaw": 1.0, "wrist_roll": 1.0, "wrist_pitch": 1.0}      # gr1t1_config.py:169-178
    cfg.asset.terminate_after_contacts_on = ["imu", "torso", "head_pitch", "waist", "upper_arm", "lower_arm", "hand"]   # gr1t1_config.py:78-86
    cfg.env.num_actions, cfg.env.num_obs = 32, 105
    cfg.env.num_pri_obs = 105 + 8 + len(cfg.terrain.measured_points_x) * len(cfg.terrain.measured_points_y)
    amax = np.array([0.79, 0.7, 0.7, 1.92, 0.52, 0.44, 0.09, 0.7, 0.7, 1.92, 0.52, 0.44, 1.05, 1.22, 0.7, 2.71, 0.35, 0.35,
                     1.92, 3.27, 2.97, 2.27, 2.97, 0.61, 0.61, 1.92, 0.57, 2.97, 2.27, 2.97, 0.61, 0.61])
    amin = np.array([-0.09, -0.7, -1.75, -0.09, -1.05, -0.44, -0.79, -0.7, -1.75, -0.09, -1.05, -0.44, -1.05, -0.52, -0.7, -2.71, -0.35, -0.52,
                     -2.79, -0.57, -2.97, -2.27, -2.97, -0.61, -0.61, -2.79, -3.27, -2.97, -2.27, -2.97, -0.61, -0.61])   # gr1t1_config.py:282-299
    cfg.normalization.clip_actions_max = amax + (np.abs(amax) + np.abs(amin)) * 0.01
    cfg.normalization.clip_actions_min = amin - (np.abs(amax) + np.abs(amin)) * 0.01
    return cfg


def full_body_tables(model):
    """robot.task_tables for a full-body model (urdf.builtin_model('GR1T1_full' / 'GR1T2_full'))."""
    from .robot import task_tables
    return task_tables(model, make_full_body_cfg("GR1T2" if "GR1T2" in str(model.get("name", "")) else "GR1T1", 4, "plane"))


def make_train_cfg(robot="GR1T1"):
    """PPO / runner config of the registered tasks as the dict ``OnPolicyRunner`` takes
    (gr1t1_config.py:310-345, gr1t1_lower_limb_config.py:107-116, legged_robot_config.py:254-294)."""
    return {
        "seed": 1,
        "runner_class_name": "OnPolicyRunner",
        "runner": {"algorithm_class_name": "PPO", "policy_class_name": "ActorCriticMLP", "experiment_name": "GR1T1",
                   "num_steps_per_env": 64, "run_name": robot.lower() + "_lower_limb", "max_iterations": 1000,
                   "save_interval": 100, "resume": False, "load_run": -1, "checkpoint": -1, "resume_path": None},
        "algorithm": {"value_loss_coef": 1.0, "use_clipped_value_loss": True, "clip_param": 0.2, "entropy_coef": 0.01,
                      "num_learning_epochs": 8, "num_mini_batches": 25, "learning_rate": 1.0e-4,
                      "learning_rate_min": 1.0e-5, "learning_rate_max": 1.0e-3, "schedule": "adaptive", "gamma": 0.99,
                      "lam": 0.95, "desired_kl": 0.03, "max_grad_norm": 1.0, "storage_class": "RolloutStorage"},
        "policy": {"init_noise_std": 0.2, "actor_hidden_dims": [512, 256, 128], "critic_hidden_dims": [512, 256, 128],
                   "activation": "elu", "actor_output_activation": None, "critic_output_activation": None,
                   "fixed_std": False},
    }


def class_to_dict(obj):
    """Same contract as legged_gym.utils.helpers.class_to_dict (helpers.py:42-57): keys in ``dir()`` (alphabetical)
    order — this fixes the reward evaluation / summation order."""
    if not hasattr(obj, "__dict__"):
        return obj
    out = {}
    for key in dir(obj):
        if key.startswith("_") or key == "to_dict":
            continue
        val = getattr(obj, key)
        if callable(val) and not hasattr(val, "__dict__"):
            continue
        if isinstance(val, list):
            out[key] = [class_to_dict(v) for v in val]
        else:
            out[key] = class_to_dict(val)
    return out


def clone_cfg(cfg):
    return copy.deepcopy(cfg)
