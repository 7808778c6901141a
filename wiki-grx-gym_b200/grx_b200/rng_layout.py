"""Layout of the per-env, per-step table of uniform draws U[N, K] consumed by one env step.

In *parity mode* the host supplies U (drawn by torch in any order it likes); in *fast mode* the kernel
fills the same slots from its own counter-based Philox stream keyed by (seed, global env id, step).
Each slot corresponds to one generator call site of the reference:
  NOISE      torch.rand_like(obs_buf)                      legged_robot.py:481
  RESET_DOF  torch_rand_float(0.5, 1.5, (n, num_dof))      legged_robot.py:726-729
  RESET_XY   torch_rand_float(-1, 1, (n, 2))               legged_robot.py:755-758
  RESET_YAW  torch_rand_float(-2pi, 2pi, (n, 1))           legged_robot.py:762-765
  RESET_VEL  torch_rand_float(-0.5, 0.5, (n, 6))           legged_robot.py:774-777
  CMD_TIME   _resample_commands from the ep_len % 500 == 0 path    legged_robot.py:317-318, 656-677
  CMD_RESET  _resample_commands from reset_idx                      legged_robot.py:402
  PUSH       torch_rand_float(-v, v, (N, 2))               legged_robot.py:790-793
  CURRICULUM torch.randint_like(levels, max_level) -> floor(u * max_level)   legged_robot.py:822
"""
NOISE = 0
NOISE_N = 39
RESET_DOF = 39
RESET_XY = 49
RESET_YAW = 51
RESET_VEL = 52
CMD_TIME = 58
CMD_RESET = 61
PUSH = 64
CURRICULUM = 66
K = 68  # padded to a multiple of 4 floats


class Layout:
    """The same slot table for a robot with `nd` actuated DOF (observation width 9 + 3 nd): K = 28 + 4 nd, a multiple of 4 floats.
    nd = 10 reproduces the module-level constants of the registered lower-limb tasks."""

    def __init__(self, nd):
        self.nd = nd
        self.NOISE, self.NOISE_N = 0, 9 + 3 * nd
        self.RESET_DOF = self.NOISE_N
        self.RESET_XY = self.RESET_DOF + nd
        self.RESET_YAW = self.RESET_XY + 2
        self.RESET_VEL = self.RESET_YAW + 1
        self.CMD_TIME = self.RESET_VEL + 6
        self.CMD_RESET = self.CMD_TIME + 3
        self.PUSH = self.CMD_RESET + 3
        self.CURRICULUM = self.PUSH + 2
        self.K = self.CURRICULUM + 2


def layout(nd):
    return Layout(nd)


_L10 = Layout(10)
assert (_L10.NOISE_N, _L10.RESET_DOF, _L10.RESET_XY, _L10.RESET_YAW, _L10.RESET_VEL, _L10.CMD_TIME, _L10.CMD_RESET, _L10.PUSH, _L10.CURRICULUM, _L10.K) == \
    (NOISE_N, RESET_DOF, RESET_XY, RESET_YAW, RESET_VEL, CMD_TIME, CMD_RESET, PUSH, CURRICULUM, K)
