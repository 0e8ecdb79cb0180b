"""Pins the C twin of the oracle (oracle/oracle_c.c) on the reference-generated goldens and on the NumPy oracle."""
import numpy as np
import pytest

from oracle import c_port, flight_np, minsnap_np
from helpers import GOAL, normwise


def test_c_planner_matches_reference_goldens(golden):
    g = golden["planning"]
    for tag in ("v2", "v3"):
        for name, wp in (("takeoff", g["waypoints"][:2]), ("course", g["waypoints"][1:])):
            c, T = c_port.solve_coeffs(wp, float(tag[1]))
            assert normwise(c, g[f"{tag}_{name}_coeffs_solve"]) < 1e-11
            np.testing.assert_allclose(T, g[f"{tag}_{name}_times"], rtol=1e-15)
        tab = c_port.mission_table(g["waypoints"], float(tag[1]), 0.01)
        assert tab.shape == g[f"{tag}_table"].shape
        ref = g[f"{tag}_table"]                                         # golden table = reference lstsq branch (~1e-9 relative noise in c)
        np.testing.assert_allclose(tab[:, :9], ref[:, :9], rtol=0, atol=1e-7)
        np.testing.assert_allclose(tab[:, 9], ref[:, 9], rtol=0, atol=1e-5)   # a row on the 1e-3 speed threshold may flip its validity
        np.testing.assert_array_equal(tab[:, 10], ref[:, 10])
    c, T, bad = c_port.solve_batch(g["c2_waypoints"], g["c2_velocity"], threads=2)
    assert bad == 0 and max(normwise(c[i], g["c2_coeffs_solve"][i]) for i in range(64)) < 1e-10
    np.testing.assert_allclose(c_port.sample_table(g["c2_coeffs_solve"][3], g["c2_times"][3], 0.01), g["c2_table3"], rtol=0, atol=1e-7)
    with pytest.raises(np.linalg.LinAlgError):
        c_port.solve_coeffs(np.array([[0.0, 0, 0], [1, 0, 0], [1, 0, 0]]), 1.0)


@pytest.mark.parametrize("v", [2, 3])
def test_c_closed_loop_matches_reference_objects(golden, v):
    g, ref = golden["planning"], golden[f"closed_loop_v{v}"]
    out = c_port.closed_loop(flight_np.Vehicle(), g[f"v{v}_table"], g["waypoints"][0], obstacles=g["obstacles"], goal=GOAL, log_stride=10)
    np.testing.assert_allclose(out["log"], ref["X"], rtol=0, atol=1e-9)
    np.testing.assert_allclose(out["omega"], ref["omega"][-1], rtol=0, atol=1e-9)
    for k in ("final_dist", "mean_err", "rmse", "max_err"):
        assert out[k] == pytest.approx(float(ref[k]), abs=1e-9)
    assert out["collision"] == bool(ref["collision"]) and out["periods"] == len(ref["errors"])


def test_c_closed_loop_variants_and_batch(golden):
    g, var = golden["planning"], golden["closed_loop_variants"]
    tab, start = g["v3_table"], g["waypoints"][0]
    vehs, refs = [], []
    for j in range(3):
        vehs.append(flight_np.Vehicle().perturbed(var[f"mc{j}_gain_scale"], float(var[f"mc{j}_mass_scale"]), var[f"mc{j}_inertia_scale"]))
        refs.append(var[f"mc{j}_X"][-1])
    m, X = c_port.closed_loop_batch(vehs, tab, start, obstacles=g["obstacles"], goal=GOAL, threads=3)
    np.testing.assert_allclose(X, np.array(refs), rtol=0, atol=1e-9)
    for j in range(3):
        assert m[j, 0] == pytest.approx(float(var[f"mc{j}_final_dist"]), abs=1e-9) and m[j, 3] == pytest.approx(float(var[f"mc{j}_mean_err"]), abs=1e-9)
    out = c_port.closed_loop(flight_np.Vehicle(), tab, start, obstacles=var["hit_obstacles"], goal=GOAL)
    assert out["collision"] and out["first_collision_tick"] == 4086
    out = c_port.closed_loop(flight_np.Vehicle(), tab, start, goal=GOAL, wind=var["wind_force"])
    np.testing.assert_allclose(out["X"], var["wind_X"][-1], rtol=0, atol=1e-9)
    out = c_port.closed_loop(flight_np.Vehicle(), tab, start, goal=GOAL, thrust_frame_lag=0)
    np.testing.assert_allclose(out["X"], var["nolag_X"][-1], rtol=0, atol=1e-9)
    # and the NumPy oracle on a perturbed vehicle the goldens do not contain
    rng = np.random.default_rng(5)
    veh = flight_np.Vehicle().perturbed(rng.uniform(0.8, 1.2, 11), 1.07, rng.uniform(0.9, 1.1, 3))
    a = flight_np.closed_loop(veh, tab[:300], start, goal=GOAL)
    b = c_port.closed_loop(veh, tab[:300], start, goal=GOAL)
    np.testing.assert_allclose(b["X"], a["X"], rtol=0, atol=1e-10)
