"""lab_course scene constants in NED, transcribed from the reference's single source of truth
``uav_ac/simulation/models/lab_course.xml`` (MuJoCo ENU -> NED is diag(1,-1,-1),
``uav_ac/simulation/mujoco_sim.py:10``).  Values match what ``MujocoSimulation`` exposes as
``mission_waypoints`` / ``obstacles`` / ``goal_position`` (reference tests
``tests/unit/simulation/test_mujoco_sim.py:10,32,40-50,246``).
"""
import numpy as np

# start (lab_course.xml:98), waypoint_00..06 (:78-84), goal (:96)
LAB_COURSE_WAYPOINTS = np.array([
    [1.0, 7.0, -0.021], [1.0, 7.0, -1.3], [4.0, 7.0, -1.3], [7.5, 4.0, -3.0], [11.0, 7.0, -3.5],
    [14.0, 10.0, -2.5], [17.0, 10.0, -3.2], [20.5, 7.0, -1.4], [23.0, 7.0, -2.0]])
# obstacle_00..03 (lab_course.xml:37,53,54,67) as [xmin xmax ymin ymax zmin zmax] (mujoco_sim.py:282-300)
LAB_COURSE_OBSTACLES = np.array([
    [3.7, 4.3, 4.0, 10.0, -3.4, -2.8], [10.7, 11.3, 4.0, 10.0, -2.2, 0.0],
    [13.3, 14.7, 6.3, 7.7, -6.0, 0.0], [20.2, 20.8, 4.0, 10.0, -3.3, -2.7]])
LAB_COURSE_START = LAB_COURSE_WAYPOINTS[0].copy()
LAB_COURSE_GOAL = LAB_COURSE_WAYPOINTS[-1].copy()
PLANNING_BOUNDS = np.array([[0.0, 0.0, -6.0], [24.0, 14.0, 0.0]])   # lab_course.xml:8


# ---------------------------------------------------------------------------------------------------------
# Scene reader without MuJoCo: the fields MujocoSimulation.__init__ extracts from an MJCF
# file (mujoco_sim.py:51-91, 258-347), read with the standard-library XML parser.  Only what the batched path
# needs is understood: top-level world geoms / sites without nested frames or rotations, one free body
# "quadrotor".  Anything else raises ValueError with the reference's wording where one exists.
import xml.etree.ElementTree as _ET
from dataclasses import dataclass as _dataclass

ENU_TO_NED = np.diag([1.0, -1.0, -1.0])          # mujoco_sim.py:10


@_dataclass
class Scene:
    """What the batched path needs from a scene, all in NED (mujoco_to_ned_state frames, mujoco_sim.py:20-45)."""
    timestep: float
    gravity: float
    mass: float
    inertia: np.ndarray
    arm_length: float
    rotor_spins: np.ndarray            # site `user` values of rotor_0..3 (+1 / -1)
    force_coefficient: float
    drag_to_thrust: float
    thrust_limits: np.ndarray
    motor_time_constants: np.ndarray
    flight_limits: np.ndarray
    planning_bounds: np.ndarray        # (2, 3): lower, upper
    start_position: np.ndarray
    goal_position: np.ndarray
    mission_waypoints: np.ndarray      # (n + 2, 3): start, waypoint_00.., goal
    obstacles: np.ndarray              # (n_obs, 6) [xmin xmax ymin ymax zmin zmax]

    def quad_kwargs(self) -> dict:
        """Keyword arguments of ``Quad(...)`` exactly as ``_create_quad`` passes them (mujoco_sim.py:268-279)."""
        return dict(g=self.gravity, dt=self.timestep, mass=self.mass, inertia=self.inertia, arm_length=self.arm_length,
                    force_coefficient=self.force_coefficient, drag_to_thrust=self.drag_to_thrust, thrust_limits=self.thrust_limits,
                    motor_time_constants=self.motor_time_constants, flight_limits=self.flight_limits)


def _floats(text, size, name):
    try:
        v = np.array([float(x) for x in str(text).split()], dtype=float)
    except ValueError:
        v = np.empty(0)
    if v.shape != (size,) or not np.all(np.isfinite(v)):
        raise ValueError(f"MuJoCo {name} must contain {size} finite values")
    return v


def load_scene(xml_path) -> Scene:
    """Parse an MJCF scene the way ``MujocoSimulation`` reads it, without MuJoCo."""
    root = _ET.parse(str(xml_path)).getroot()
    option = root.find("option")
    if option is None:
        raise ValueError("MuJoCo scene is missing required element 'option'")
    timestep = float(option.get("timestep", "0.002"))
    gravity = float(np.linalg.norm(_floats(option.get("gravity", "0 0 -9.81"), 3, "gravity")))
    if gravity == 0:
        raise ValueError("MuJoCo gravity must be non-zero")                                   # mujoco_sim.py:265-266

    numerics = {n.get("name"): n.get("data") for n in root.iter("numeric")}

    def numeric(name, size):                                                                  # mujoco_sim.py:328-334
        if name not in numerics:
            raise ValueError(f"MuJoCo scene is missing required element '{name}'")
        v = np.array([float(x) for x in numerics[name].split()])
        if v.size != size:
            raise ValueError(f"MuJoCo numeric '{name}' must contain {size} values")
        return v

    world = root.find("worldbody")
    if world is None:
        raise ValueError("MuJoCo scene is missing required element 'worldbody'")
    body = next((b for b in world.iter("body") if b.get("name") == "quadrotor"), None)
    if body is None:
        raise ValueError("MuJoCo scene is missing required element 'quadrotor'")
    inertial = body.find("inertial")
    if inertial is None:
        raise ValueError("MuJoCo scene is missing required element 'inertial'")
    start = ENU_TO_NED @ _floats(body.get("pos", "0 0 0"), 3, "start position")
    rotor_pos, spins = [], []
    for k in range(4):
        site = next((s for s in body.iter("site") if s.get("name") == f"rotor_{k}"), None)
        if site is None:
            raise ValueError(f"MuJoCo scene is missing required element 'rotor_{k}'")
        rotor_pos.append(_floats(site.get("pos"), 3, f"rotor_{k} position"))
        spins.append(float(site.get("user", "0").split()[0]))
    arms = np.abs(np.array(rotor_pos)[:, :2])
    if not np.allclose(arms, arms[0, 0]):
        raise ValueError("MuJoCo rotor sites must use a symmetric X configuration")           # mujoco_sim.py:261-262

    sites = {s.get("name"): s for s in world.findall("site")}
    if "goal" not in sites:
        raise ValueError("MuJoCo scene is missing required element 'goal'")
    goal = ENU_TO_NED @ _floats(sites["goal"].get("pos"), 3, "goal position")
    names = sorted(n for n in sites if n is not None and n.startswith("waypoint_"))
    if names != [f"waypoint_{i:02d}" for i in range(len(names))]:
        raise ValueError("MuJoCo mission waypoints must be consecutively numbered from waypoint_00")   # :317-318
    if not names:
        raise ValueError("MuJoCo scene must define at least one mandatory waypoint")                   # :319-320
    mandatory = np.array([ENU_TO_NED @ _floats(sites[n].get("pos"), 3, n) for n in names])

    obstacles = []
    for geom in world.findall("geom"):                                                        # mujoco_sim.py:282-300
        name = geom.get("name")
        if name is None or not name.startswith("obstacle_"):
            continue
        if geom.get("type") != "box":
            raise ValueError(f"MuJoCo planning obstacle '{name}' must be an axis-aligned box")
        if any(geom.get(a) is not None for a in ("quat", "euler", "axisangle", "xyaxes", "zaxis")):
            raise ValueError(f"MuJoCo planning obstacle '{name}' must be axis-aligned")
        c = ENU_TO_NED @ _floats(geom.get("pos", "0 0 0"), 3, f"{name} position")
        h = _floats(geom.get("size"), 3, f"{name} size")
        obstacles.append([c[0] - h[0], c[0] + h[0], c[1] - h[1], c[1] + h[1], c[2] - h[2], c[2] + h[2]])
    obstacles = np.asarray(obstacles, dtype=float).reshape(-1, 6)
    for box in obstacles:                                                                     # mujoco_sim.py:90-91
        if box[0] <= start[0] <= box[1] and box[2] <= start[1] <= box[3] and box[4] <= start[2] <= box[5]:
            raise ValueError("MuJoCo start position is inside a planning obstacle")

    b = numeric("planning_bounds", 6)
    return Scene(timestep=timestep, gravity=gravity, mass=float(inertial.get("mass")),
                 inertia=_floats(inertial.get("diaginertia"), 3, "diaginertia"), arm_length=float(arms[0, 0]), rotor_spins=np.array(spins),
                 force_coefficient=float(numeric("rotor_force_coefficient", 1)[0]), drag_to_thrust=float(numeric("rotor_drag_to_thrust", 1)[0]),
                 thrust_limits=numeric("rotor_thrust_limits", 2), motor_time_constants=numeric("motor_time_constants", 2),
                 flight_limits=numeric("flight_limits", 5), planning_bounds=np.array([b[:3], b[3:]]), start_position=start, goal_position=goal,
                 mission_waypoints=np.vstack((start, mandatory, goal)), obstacles=obstacles)


def mujoco_to_ned_state(position, quaternion, velocity) -> np.ndarray:
    """MuJoCo ENU/FLU free-joint state(s) -> the NED/FRD state vector(s) X of the batched path (mujoco_sim.py:20-45),
    for users who replay or seed rollouts from a MuJoCo simulation.  Accepts (3,), (4,), (6,) or batches (B, .)."""
    p, q, v = (np.asarray(a, dtype=float) for a in (position, quaternion, velocity))
    if p.shape[-1] != 3 or q.shape[-1] != 4 or v.shape[-1] != 6 or not all(np.all(np.isfinite(a)) for a in (p, q, v)):
        raise ValueError("MuJoCo position / quaternion / velocity must contain 3 / 4 / 6 finite values")
    n = np.linalg.norm(q, axis=-1, keepdims=True)
    if np.any(n == 0):
        raise ValueError("MuJoCo quaternion cannot be zero")
    flip = np.array([1.0, -1.0, -1.0])
    return np.concatenate((p * flip, q / n * np.array([1.0, 1.0, -1.0, -1.0]), v[..., :3] * flip, v[..., 3:] * flip), axis=-1)
