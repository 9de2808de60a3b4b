"""Shims that let the UNMODIFIED reference classes run in the build container (TEST INFRASTRUCTURE).

Used only by ``tests/golden/make_golden.py`` (and the pinning tests that skip when /root/reference
is absent).  Nothing here travels into the product and nothing here is read on the GPU box.

``install(reference_root)`` registers stub modules for the packages the reference imports but this
image lacks:

* ``isaacgym`` - FakeGym: ``gymapi`` / ``gymtorch`` / ``gymutil`` + ``torch_utils`` (=
  ``oracle.isaac_torch_utils``).  The physics step is replaced by a frame provider: state frames
  pushed with ``FakeGym.push_frame`` become visible at ``refresh_*`` (at ``simulate`` for tasks that
  never refresh); ``set_*_tensor_indexed`` / force setters are recorded in ``gym.log``.
* ``gym.spaces`` (Space / Box), ``matplotlib.pyplot``; ``np.Inf`` for numpy 2.
* the ``agents`` package is registered as a bare namespace so that ``agents/__init__.py`` (which pulls
  every algorithm in) is skipped, and ``agents/tasks`` resolves first to a temp directory holding
  patched copies of ``ten_ant.py`` / ``one_ant.py`` (bool-minus-int fix of SURVEY finding 5 and the four
  in-jit prints removed).  TorchScript needs real files, hence copies - they live under the system
  temp dir, never in this repository.
"""
import os
import re
import sys
import tempfile
import types

import numpy as np
import torch

from .. import isaac_torch_utils as itu

ANT_BODY_NAMES = ["torso", "front_left_leg", "front_left_foot", "front_right_leg", "front_right_foot",
                  "left_back_leg", "left_back_foot", "right_back_leg", "right_back_foot"]


class _Vec3:
    def __init__(self, x=0.0, y=0.0, z=0.0):
        self.x, self.y, self.z = float(x), float(y), float(z)


class _Quat:
    def __init__(self, x=0.0, y=0.0, z=0.0, w=1.0):
        self.x, self.y, self.z, self.w = x, y, z, w


class _Transform:
    def __init__(self):
        self.p = _Vec3()
        self.r = _Quat()


class _Bag:
    """attribute bag used for PlaneParams / AssetOptions / SimParams / CameraProperties"""

    def __init__(self, **kw):
        self.__dict__.update(kw)

    def __getattr__(self, name):  # nested attribute bags on demand (sim_params.gravity.x = ...)
        if name.startswith("__"):
            raise AttributeError(name)
        b = _Bag()
        self.__dict__[name] = b
        return b


class _Asset:
    def __init__(self, kind):
        self.kind = kind
        self.n_sensors = 0


class FakeGym:
    def __init__(self):
        self.actors = []          # (kind, Transform) in creation order = sim-domain index
        self.sensors = 0
        self.frames = []          # list of dicts root/dof/sensor
        self.cursor = -1
        self.log = []             # (name, tensor clone)
        self.visible_at_simulate = False
        self.sim_params = None

    # --- frame provider ------------------------------------------------------------------
    def push_frame(self, root, dof=None, sensor=None):
        self.frames.append(dict(root=root, dof=dof, sensor=sensor))

    def _cur(self):
        return self.frames[self.cursor] if 0 <= self.cursor < len(self.frames) else None

    def simulate(self, sim):
        self.cursor += 1
        if self.visible_at_simulate and self._cur() is not None:
            self.root.copy_(self._cur()["root"])

    def refresh_actor_root_state_tensor(self, sim):
        f = self._cur()
        if f is not None:
            self.root.copy_(f["root"])

    def refresh_dof_state_tensor(self, sim):
        f = self._cur()
        if f is not None and f["dof"] is not None:
            self.dof.copy_(f["dof"])

    def refresh_force_sensor_tensor(self, sim):
        f = self._cur()
        if f is not None and f["sensor"] is not None:
            self.sensor.copy_(f["sensor"].view_as(self.sensor))

    # --- construction --------------------------------------------------------------------
    def create_sim(self, compute_device, graphics_device, physics_engine, sim_params):
        self.sim_params = sim_params
        return object()

    def add_ground(self, sim, params):
        pass

    def load_asset(self, sim, root, file, options=None):
        return _Asset("ant" if "ant" in file else "ingenuity")

    def create_box(self, sim, x, y, z, options=None):
        return _Asset("box")

    def get_asset_dof_count(self, asset):
        return {"ant": 8, "ingenuity": 4, "box": 0}[asset.kind]

    def get_asset_rigid_body_count(self, asset):
        return {"ant": 9, "ingenuity": 6, "box": 1}[asset.kind]

    def get_asset_rigid_body_name(self, asset, i):
        return ANT_BODY_NAMES[i]

    def find_asset_rigid_body_index(self, asset, name):
        return ANT_BODY_NAMES.index(name)

    def get_asset_actuator_properties(self, asset):
        return [_Bag(motor_effort=15.0) for _ in range(8)]

    def create_asset_force_sensor(self, asset, body_idx, pose):
        asset.n_sensors += 1

    def create_env(self, sim, lower, upper, num_per_row):
        return object()

    def create_actor(self, env, asset, pose, name, group=0, filt=0, seg=0):
        self.actors.append((asset, pose))
        return len(self.actors) - 1

    def get_actor_index(self, env, handle, domain):
        return handle

    def get_actor_rigid_shape_properties(self, env, handle):
        return [_Bag()]

    def set_actor_rigid_shape_properties(self, env, handle, props):
        pass

    def set_rigid_body_color(self, *a):
        pass

    def get_actor_dof_properties(self, env, handle):
        from ..task_oracle import ANT_DOF_RANGE_DEG
        asset = self.actors[handle][0]
        n = self.get_asset_dof_count(asset)
        props = {"lower": np.zeros(n, np.float32), "upper": np.zeros(n, np.float32),
                 "stiffness": np.zeros(n, np.float32), "damping": np.zeros(n, np.float32)}
        if asset.kind == "ant":
            props["lower"] = np.array([np.float32(np.radians(a)) for a, _ in ANT_DOF_RANGE_DEG], np.float32)
            props["upper"] = np.array([np.float32(np.radians(b)) for _, b in ANT_DOF_RANGE_DEG], np.float32)
        return props

    def set_actor_dof_properties(self, env, handle, props):
        pass

    def find_actor_rigid_body_handle(self, env, handle, name):
        return ANT_BODY_NAMES.index(name)

    def prepare_sim(self, sim):
        rows, ndof, nsens = [], 0, 0
        for asset, pose in self.actors:
            rows.append([pose.p.x, pose.p.y, pose.p.z, pose.r.x, pose.r.y, pose.r.z, pose.r.w, 0, 0, 0, 0, 0, 0])
            ndof += self.get_asset_dof_count(asset)
            nsens += asset.n_sensors
        # every actor of an asset shares the asset object, so n_sensors is per asset (4 feet)
        nsens = sum(4 for asset, _ in self.actors if asset.kind == "ant")
        self.root = torch.tensor(rows, dtype=torch.float32)
        self.dof = torch.zeros(ndof, 2)
        self.sensor = torch.zeros(max(nsens, 1), 6)

    def acquire_actor_root_state_tensor(self, sim):
        return self.root

    def acquire_dof_state_tensor(self, sim):
        return self.dof

    def acquire_force_sensor_tensor(self, sim):
        return self.sensor

    def get_sim_dof_count(self, sim):
        return self.dof.shape[0]

    def get_sim_params(self, sim):
        return self.sim_params

    def get_frame_count(self, sim):
        return self.cursor

    def fetch_results(self, sim, flag):
        pass

    # --- recorded side effects -----------------------------------------------------------------
    def set_actor_root_state_tensor_indexed(self, sim, states, idx, n):
        self.log.append(("root_indexed", idx.clone(), int(n)))

    def set_dof_state_tensor_indexed(self, sim, states, idx, n):
        self.log.append(("dof_indexed", idx.clone(), int(n), states.clone()))

    def set_dof_actuation_force_tensor(self, sim, forces):
        self.log.append(("dof_forces", forces.clone()))

    def apply_rigid_body_force_tensors(self, sim, forces, torques, space):
        self.log.append(("body_forces", forces.clone()))


_GYM = None


def current_gym() -> FakeGym:
    return _GYM


def new_gym() -> FakeGym:
    global _GYM
    _GYM = FakeGym()
    return _GYM


def _stub_modules():
    if not hasattr(np, "Inf"):
        np.Inf = np.inf
    gymapi = types.ModuleType("isaacgym.gymapi")
    gymapi.acquire_gym = lambda: _GYM
    gymapi.Vec3 = _Vec3
    gymapi.Quat = _Quat
    gymapi.Transform = _Transform
    for n in ("PlaneParams", "AssetOptions", "SimParams", "CameraProperties"):
        setattr(gymapi, n, _Bag)
    for i, n in enumerate(("DOMAIN_SIM", "MESH_VISUAL", "UP_AXIS_Z", "DOF_MODE_NONE", "SIM_PHYSX", "SIM_FLEX",
                           "LOCAL_SPACE", "KEY_ESCAPE", "KEY_V")):
        setattr(gymapi, n, i)
    gymtorch = types.ModuleType("isaacgym.gymtorch")
    gymtorch.wrap_tensor = lambda t: t
    gymtorch.unwrap_tensor = lambda t: t
    gymutil = types.ModuleType("isaacgym.gymutil")
    for n in ("get_property_setter_map", "get_property_getter_map", "get_default_setter_args",
              "apply_random_samples", "check_buckets", "generate_random_samples"):
        setattr(gymutil, n, lambda *a, **k: None)
    isaacgym = types.ModuleType("isaacgym")
    isaacgym.gymapi, isaacgym.gymtorch, isaacgym.gymutil, isaacgym.torch_utils = gymapi, gymtorch, gymutil, itu
    sys.modules.update({"isaacgym": isaacgym, "isaacgym.gymapi": gymapi, "isaacgym.gymtorch": gymtorch,
                        "isaacgym.gymutil": gymutil, "isaacgym.torch_utils": itu})

    class Space:
        def __init__(self, shape=None, dtype=None):
            self.shape = shape

    class Box(Space):
        def __init__(self, low, high, shape=None, dtype=np.float32):
            if shape is None:
                shape = np.shape(low)
            super().__init__(tuple(shape), dtype)
            self.low, self.high = low, high

    gym = types.ModuleType("gym")
    spaces = types.ModuleType("gym.spaces")
    spaces.Space, spaces.Box = Space, Box
    gym.spaces = spaces
    sys.modules.setdefault("gym", gym)
    sys.modules.setdefault("gym.spaces", spaces)

    mpl = types.ModuleType("matplotlib")
    plt = types.ModuleType("matplotlib.pyplot")
    def _plt_attr(name):
        if name.startswith("__"):               # (inspect.getmodule walks sys.modules and asks for __file__)
            raise AttributeError(name)
        return lambda *a, **k: None
    plt.__getattr__ = _plt_attr
    mpl.pyplot = plt
    sys.modules.setdefault("matplotlib", mpl)
    sys.modules.setdefault("matplotlib.pyplot", plt)


def _patched_task_dir(reference_root):
    """Write patched copies of the two task files that cannot run as shipped (see module docstring)."""
    out = os.path.join(tempfile.gettempdir(), "mmb_refshim_%d" % os.getuid(), "tasks")
    os.makedirs(out, exist_ok=True)
    for name in ("ten_ant.py", "one_ant.py"):
        src = open(os.path.join(reference_root, "agents", "tasks", name)).read()
        src, n = re.subn(r"abs\((ant_push(?:_\d+)?) - 1\)", r"abs(\1.long() - 1)", src)
        assert n in (1, 10), (name, n)
        if name == "ten_ant.py":
            lines = src.split("\n")
            start = next(i for i, l in enumerate(lines) if l.startswith("def compute_ant_observations"))
            kept, dropped = [], 0
            for i, l in enumerate(lines):
                if i > start and dropped < 4 and l.strip().startswith("print("):
                    dropped += 1
                    continue
                kept.append(l)
            assert dropped == 4
            src = "\n".join(kept)
        with open(os.path.join(out, name), "w") as f:
            f.write(src)
    return out


def install(reference_root="/root/reference"):
    """Make ``import agents.tasks.ten_ant`` etc. resolve to the reference under the shims."""
    if not os.path.isdir(os.path.join(reference_root, "agents")):
        raise FileNotFoundError(reference_root)
    _stub_modules()
    if "agents" in sys.modules and getattr(sys.modules["agents"], "_mmb_shim", False):
        return
    patched = _patched_task_dir(reference_root)

    def ns(name, paths):
        m = types.ModuleType(name)
        m.__path__ = paths
        m._mmb_shim = True
        sys.modules[name] = m
        return m

    ag = os.path.join(reference_root, "agents")
    ns("agents", [ag])
    ns("agents.tasks", [patched, os.path.join(ag, "tasks")])
    ns("agents.tasks.agent_base", [os.path.join(ag, "tasks", "agent_base")])
    ns("agents.utils", [os.path.join(ag, "utils")])
    ns("agents.algorithms", [os.path.join(ag, "algorithms")])
    ns("agents.algorithms.rl", [os.path.join(ag, "algorithms", "rl")])
    ns("agents.algorithms.rl.ppo", [os.path.join(ag, "algorithms", "rl", "ppo")])
    ns("agents.algorithms.marl", [os.path.join(ag, "algorithms", "marl")])
    ns("agents.algorithms.marl.utils", [os.path.join(ag, "algorithms", "marl", "utils")])
    ns("agents.algorithms.utils", [os.path.join(ag, "algorithms", "utils")])


def make_cfg(num_envs, env_name):
    """The cfg dict the task constructors read (cfg/TenAnt.yaml et al.)."""
    return {
        "env": {"numEnvs": num_envs, "env_name": env_name, "envSpacing": 40, "episodeLength": 1000,
                "enableDebugVis": False, "powerScale": 1.0, "headingWeight": 0.5, "upWeight": 0.1,
                "actionsCost": 0.005, "energyCost": 0.05, "dofVelocityScale": 0.2, "contactForceScale": 0.1,
                "jointsAtLimitCost": 0.1, "deathCost": -2.0, "terminationHeight": 0.31,
                "plane": {"staticFriction": 1.0, "dynamicFriction": 1.0, "restitution": 0.0},
                "asset": {"assetFileName": "mjcf/open_ai_assets/ant/nv_ant.xml"}},
        "sim": {"dt": 0.0166},
        "task": {"randomize": False, "randomization_params": {}},
        "seed": 1,
    }


def make_task(kind, num_envs, is_multi_agent=False):
    """Construct the reference task class ('TenAnt' | 'OneAnt' | 'MultiIngenuity') on CPU under FakeGym."""
    import importlib
    import io
    import contextlib
    g = new_gym()
    g.visible_at_simulate = (kind == "MultiIngenuity")
    mod = {"TenAnt": "ten_ant", "OneAnt": "one_ant", "MultiIngenuity": "multi_ingenuity"}[kind]
    m = importlib.import_module("agents.tasks." + mod)
    sim_params = _Bag(dt=0.0166)
    with contextlib.redirect_stdout(io.StringIO()):
        task = getattr(m, kind)(make_cfg(num_envs, mod), sim_params, 0, "cpu", 0, True, is_multi_agent)
    return task, g
