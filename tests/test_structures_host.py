"""CPU: the host-side hydraulics of anuga_core_b200/structures.py against the reference's own functions and
operator objects, live, on random inputs (skipped when the scratch build of the reference is absent; the
golden evolve cases cover the same code through the oracle and on the GPU)."""
import numpy as np
import pytest

import anuga_core_b200 as ab
from anuga_core_b200 import structures as S
from oracle import pyref

pytestmark = pytest.mark.skipif(not pyref.available(), reason="python reference not built (oracle/build_pyref.py)")


def test_rating_functions_equal_reference_on_random_inputs():
    anuga = pyref.import_anuga()
    from anuga.structures.boyd_box_operator import boyd_box_function, total_energy, smooth_discharge
    from anuga.structures.boyd_pipe_operator import boyd_pipe_function
    rng = np.random.default_rng(5)
    cases_seen = set()
    for _ in range(3000):
        width, depth = rng.uniform(0.3, 4.0), rng.uniform(0.2, 3.0)
        blockage = float(rng.choice([0.0, rng.uniform(0, 0.9), rng.uniform(0.9, 1.0), 1.0]))
        barrels = float(rng.integers(1, 4))
        length = rng.uniform(1.0, 40.0)
        drive = rng.uniform(0.011, 4.0)
        delta = drive * rng.uniform(0.0, 1.5)
        tail = rng.uniform(0.0, 4.0)
        loss, manning = rng.uniform(0.0, 3.0), rng.uniform(0.01, 0.05)
        a = S.boyd_box_function(width, depth, blockage, barrels, width, length, drive, delta, tail, loss, manning)
        b = boyd_box_function(width, depth, blockage, barrels, width, length, drive, delta, tail, loss, manning)
        assert tuple(a[:4]) == tuple(b[:4]), (a, b)
        cases_seen.add(b[4])
        a = S.boyd_pipe_function(drive, width, blockage, barrels, length, drive, delta, tail, loss, manning)
        b = boyd_pipe_function(drive, width, blockage, barrels, length, drive, delta, tail, loss, manning)
        assert tuple(a[:4]) == tuple(b[:4]), (a, b)
        cases_seen.add(b[4])
        sm, de, dt, ts = rng.uniform(-1, 1), rng.uniform(-1, 1), float(rng.choice([0.0, rng.uniform(0, 0.2)])), rng.uniform(0, 3)
        assert S.total_energy(sm, de, dt, ts, True) == total_energy(sm, de, dt, ts, True)
        assert S.total_energy(sm, de, dt, ts, False) == total_energy(sm, de, dt, ts, False)
        sq, q, area, tt = rng.uniform(-2, 2), rng.uniform(0, 3), float(rng.choice([0.0, rng.uniform(0.1, 3)])), rng.uniform(0, 1)
        assert tuple(S.smooth_discharge(sm, sq, q, area, tt, True)) == tuple(smooth_discharge(sm, sq, q, area, tt, True))
    assert len(cases_seen) >= 8          # weir / orifice / submerged / full / part-full / blocked branches


def test_weir_orifice_trapezoid_function_equals_reference_on_random_inputs():
    anuga = pyref.import_anuga()
    from anuga.structures.weir_orifice_trapezoid_operator import weir_orifice_trapezoid_function
    rng = np.random.default_rng(9)
    cases_seen = set()
    for _ in range(2000):
        width, depth = rng.uniform(0.3, 4.0), rng.uniform(0.2, 3.0)
        blockage = float(rng.choice([0.0, rng.uniform(0, 0.9), 1.0]))
        barrels = float(rng.integers(1, 4))
        z1, z2 = float(rng.choice([0.0, rng.uniform(0, 2)])), float(rng.uniform(0, 2))
        length = rng.uniform(1.0, 40.0)
        drive = rng.uniform(0.011, 4.0)
        delta = drive * rng.uniform(0.01, 1.5)
        tail = rng.uniform(0.0, 4.0)
        loss, manning = rng.uniform(0.0, 3.0), rng.uniform(0.01, 0.05)
        args = (width, depth, blockage, barrels, z1, z2, width, length, drive, delta, tail, loss, manning)
        a = S.weir_orifice_trapezoid_function(*args)
        b = weir_orifice_trapezoid_function(*args)
        assert tuple(a[:4]) == tuple(b[:4]), (args, a, b)
        cases_seen.add(b[4])
    assert len(cases_seen) >= 5


class _HostArrays:
    """stands in for the device: the gather / scatter entry points on plain arrays"""

    def __init__(self, d):
        q = d.quantities
        self.a = [q[k].centroid_values for k in ("stage", "xmomentum", "ymomentum", "elevation")]

    def gather_centroids(self, ids):
        return np.stack([x[ids] for x in self.a], axis=1)

    def scatter_centroids(self, ids, v):
        for k in range(3):
            self.a[k][ids] = v[:, k]

    def set_time(self, t):
        pass


def _pair(seed):
    anuga = pyref.import_anuga()
    rng = np.random.default_rng(seed)
    coef = rng.uniform(-1, 1, size=6)
    level = rng.uniform(0.1, 1.0, size=2)

    def build(A):
        d = A.rectangular_cross_domain(14, 8, len1=14.0, len2=8.0)
        d.set_flow_algorithm("DE1")
        d.set_store(False)
        d.set_quantity("elevation", lambda x, y: 0.6 / (1.0 + (x - 7.0) ** 2) + 0.02 * coef[0] * y)
        d.set_quantity("stage", lambda x, y: (x < 7.0) * level[0] + (x >= 7.0) * level[1] + 0.01 * coef[1] * x,
                       location="centroids")
        d.set_quantity("xmomentum", lambda x, y: 0.1 * coef[2] + 0.01 * coef[3] * y, location="centroids")
        d.set_quantity("ymomentum", lambda x, y: 0.1 * coef[4] + 0.01 * coef[5] * x, location="centroids")
        return d
    return anuga, build, rng


@pytest.mark.parametrize("seed", range(8))
def test_operators_equal_reference_objects_on_random_states(seed):
    """Inlet_operator, Boyd_box_operator, Boyd_pipe_operator: one application with a random state,
    parameters and timestep - same arrays and statistics as the reference's objects"""
    anuga, build, rng = _pair(seed)
    ref, mine = build(anuga), build(ab)
    dt = float(rng.uniform(0.01, 0.3))
    Qin = float(rng.uniform(-3.0, 3.0))
    kw_box = dict(losses=float(rng.uniform(0, 2)), width=float(rng.uniform(0.8, 2.0)), height=float(rng.uniform(0.3, 1.0)),
                  barrels=float(rng.integers(1, 3)), blockage=float(rng.choice([0.0, 0.3])),
                  end_points=[[4.3, 5.3], [9.7, 5.3]], apron=0.55, enquiry_gap=0.4,
                  smoothing_timescale=float(rng.choice([0.0, 1.0])),
                  use_momentum_jet=bool(rng.integers(0, 2)), use_velocity_head=bool(rng.integers(0, 2)))
    kw_pipe = dict(losses=[0.5, 0.7], diameter=float(rng.uniform(0.4, 1.2)), barrels=1.0, blockage=float(rng.choice([0.0, 0.95])),
                   end_points=[[4.3, 2.3], [9.7, 2.3]], apron=0.55, enquiry_gap=0.45,
                   use_momentum_jet=bool(rng.integers(0, 2)), use_velocity_head=True)
    kw_weir = dict(losses={"in": 0.5, "out": 1.0}, width=float(rng.uniform(0.8, 2.0)), height=float(rng.uniform(0.3, 1.0)),
                   z1=float(rng.uniform(0, 1.5)), z2=float(rng.uniform(0, 1.5)), barrels=1.0,
                   end_points=[[4.3, 6.9], [9.7, 6.9]], apron=0.45, enquiry_gap=0.35,
                   smoothing_timescale=float(rng.choice([0.0, 0.5])), use_momentum_jet=bool(rng.integers(0, 2)),
                   use_velocity_head=bool(rng.integers(0, 2)))
    ops = {}
    for A, d in ((anuga, ref), (ab, mine)):
        c = d.centroid_coordinates
        ids = np.flatnonzero((c[:, 0] > 1.1) & (c[:, 0] < 3.2) & (c[:, 1] > 5.8))
        ops[A] = [A.Inlet_operator(d, A.Region(d, indices=ids), Q=Qin),
                  A.Boyd_box_operator(d, **kw_box), A.Boyd_pipe_operator(d, **kw_pipe),
                  A.Weir_orifice_trapezoid_operator(d, **kw_weir)]
    mine._dev = _HostArrays(mine)
    for d in (ref, mine):
        d.timestep = dt
        d.yieldstep = 1.0
    for a, b in zip(ops[anuga], ops[ab]):
        a()
        b()
    for k in ("stage", "xmomentum", "ymomentum"):
        assert np.array_equal(mine.quantities[k].centroid_values, ref.quantities[k].centroid_values), k
    for a, b in zip(ops[anuga][1:], ops[ab][1:]):
        for k in ("discharge", "velocity", "outlet_depth", "accumulated_flow", "delta_total_energy", "driving_energy",
                  "smooth_Q", "smooth_delta_total_energy", "case"):
            assert getattr(a, k) == getattr(b, k), k
    assert ops[anuga][0].applied_Q == ops[ab][0].applied_Q


class _HostArraysWithBed(_HostArrays):
    def scatter_bed(self, ids, v):
        self.a[3][ids] = v


@pytest.mark.parametrize("seed", range(6))
def test_host_side_operators_equal_reference_objects(seed):
    """Rate_operator with rate(x, y, t), Set_stage_operator, Set_quantity_operator, Set_elevation_operator and the
    one-off Set_stage / Set_elevation: one application on a random state == the reference's objects"""
    anuga, build, rng = _pair(100 + seed)
    ref, mine = build(anuga), build(ab)
    dt = float(rng.uniform(0.01, 0.3))
    k1, k2, k3 = rng.uniform(0.5, 2.0, size=3)
    poly = [[2.2, 1.3], [9.1, 1.9], [8.4, 6.6], [3.3, 5.2]]
    ops = {}
    for A, d in ((anuga, ref), (ab, mine)):
        A.Set_stage(d, stage=0.55 * k1, center=[10.5, 4.0], radius=1.7)()
        A.Set_elevation(d, elevation=lambda x, y: 0.1 * k2 + 0.01 * x, polygon=poly)()
        ops[A] = [A.Rate_operator(d, rate=lambda x, y, t: 0.03 * k1 / (1.0 + (x - 3.0 - t) ** 2)),
                  A.Rate_operator(d, rate=lambda x, y, t: -0.5 * k2 * (1.0 + 0.1 * y), factor=0.7, polygon=poly),
                  A.Set_stage_operator(d, stage=lambda t: 0.4 * k3 + t, center=[4.0, 4.0], radius=1.3),
                  A.Set_quantity_operator(d, "xmomentum", value=lambda x, y: 0.01 * k3 * x, polygon=poly),
                  A.Set_elevation_operator(d, elevation=lambda x, y, t: 0.05 * k1 * y + t, center=[11.0, 2.0], radius=1.5)]
    for k in ("stage", "elevation"):      # the one-off setters already agree
        assert np.array_equal(mine.quantities[k].centroid_values, ref.quantities[k].centroid_values), k
    for d in (ref, mine):
        d.timestep = dt
        d.yieldstep = 1.0
        d.set_time(0.37)
    mine._dev = _HostArraysWithBed(mine)
    added = 0.0
    for a, b in zip(ops[anuga], ops[ab]):
        a()
        added += b() or 0.0
    for k in ("stage", "xmomentum", "ymomentum", "elevation"):
        assert np.array_equal(mine.quantities[k].centroid_values, ref.quantities[k].centroid_values), k
    assert abs(added - ref.fractional_step_volume_integral) <= 1e-12 * max(1.0, abs(added))


def test_force_constant_inlet_elevations_matches_reference():
    anuga, build, rng = _pair(42)
    ref, mine = build(anuga), build(ab)
    for A, d in ((anuga, ref), (ab, mine)):
        A.Structure_operator(d, end_points=[[4.3, 5.3], [9.7, 5.3]], width=1.4, height=0.8, apron=0.7,
                             enquiry_gap=0.3, force_constant_inlet_elevations=True)
    a, b = mine.quantities["elevation"].centroid_values, ref.quantities["elevation"].centroid_values
    assert np.array_equal(a, b)
    assert len(np.unique(a)) < len(np.unique(build(ab).quantities["elevation"].centroid_values))


def test_inlet_from_line_or_polygon_matches_reference():
    anuga, build, rng = _pair(77)
    ref, mine = build(anuga), build(ab)
    for shape in ([[2.3, 1.2], [2.9, 6.6]], [[9.1, 1.1], [12.2, 1.4], [11.8, 3.9], [9.4, 3.2]]):
        a = ab.Inlet_operator(mine, shape, Q=1.0).inlet.triangle_indices
        b = anuga.Inlet_operator(ref, shape, Q=1.0).inlet.triangle_indices
        assert np.array_equal(a, np.asarray(b, dtype=np.int64)) and len(a) > 3


@pytest.mark.parametrize("seed", range(8))
def test_internal_boundary_operator_equals_reference_object(seed):
    """Internal_boundary_operator (explicit and implicit discharge, with and without smoothing / velocity head)
    driven by a weir-like rating and by pumping_station_function: three applications == the reference's object
    (structures/internal_boundary_operator.py, internal_boundary_functions.py:392-477)"""
    anuga, build, rng = _pair(200 + seed)
    from anuga.structures.internal_boundary_functions import pumping_station_function as ref_pump
    ref, mine = build(anuga), build(ab)
    crest = float(rng.uniform(0.02, 0.12))
    k = float(rng.uniform(0.5, 2.0))

    def weir(hw, tw):
        head = max(hw, tw) - crest
        if head <= 0.0:
            return 0.0
        sub = max(min(hw, tw) - crest, 0.0) / head
        q = k * head ** 1.5 * (1.0 - sub ** 1.5) ** 0.385
        return q if hw >= tw else -q
    common = dict(width=float(rng.uniform(0.8, 2.0)), height=1.0, apron=0.55, enquiry_gap=0.4, verbose=False,
                  force_constant_inlet_elevations=bool(rng.integers(0, 2)))
    kw_weir = dict(compute_discharge_implicitly=bool(seed % 2), smoothing_timescale=float(rng.choice([0.0, 1.0])),
                   use_velocity_head=bool(rng.integers(0, 2)), zero_outflow_momentum=bool(rng.integers(0, 2)))
    ops = {}
    for A, d, pump in ((anuga, ref, ref_pump), (ab, mine, ab.pumping_station_function)):
        P = pump(d, pump_capacity=0.8, hw_to_start_pumping=0.3, hw_to_stop_pumping=0.1, initial_pump_rate=0.1,
                 pump_rate_of_increase=0.7, pump_rate_of_decrease=0.9, verbose=False)
        ops[A] = [A.Internal_boundary_operator(d, weir, end_points=[[4.3, 5.3], [9.7, 5.3]], **kw_weir, **common),
                  A.Internal_boundary_operator(d, P, end_points=[[4.3, 2.3], [9.7, 2.3]],
                                               compute_discharge_implicitly=False, **common)]
    for key in ("stage", "elevation"):
        assert np.array_equal(mine.quantities[key].centroid_values, ref.quantities[key].centroid_values), key
    mine._dev = _HostArrays(mine)
    t = 0.0
    for step in range(3):
        dt = float(rng.uniform(0.01, 0.3))
        t += dt
        for d in (ref, mine):
            d.timestep = dt
            d.yieldstep = 1.0
            d.set_time(t)
        for a, b in zip(ops[anuga], ops[ab]):
            a()
            b()
        for key in ("stage", "xmomentum", "ymomentum"):
            assert np.array_equal(mine.quantities[key].centroid_values, ref.quantities[key].centroid_values), (step, key)
        for a, b in zip(ops[anuga], ops[ab]):
            for key in ("discharge", "accumulated_flow", "delta_total_energy", "driving_energy", "smooth_Q",
                        "smooth_delta_total_energy", "case"):
                assert getattr(a, key) == getattr(b, key), (step, key, getattr(a, key), getattr(b, key))
    assert ops[ab][1].accumulated_flow != 0.0      # (the weir moves water in 7 of the 8 seeds)


def test_operator_reports_follow_the_reference(capsys):
    """statistics / timestepping_statistics of the operators and the domain-level print_operator_* loops
    (structure_operator.py:498-640, inlet_operator.py:186-233, rate_operators.py:444-481, 598-601,
    generic_domain.py:2320-2326): same figures as the reference's objects after one application"""
    anuga, build, rng = _pair(300)
    ref, mine = build(anuga), build(ab)
    ops = {}
    for A, d in ((anuga, ref), (ab, mine)):
        c = d.centroid_coordinates
        ids = np.flatnonzero((c[:, 0] > 1.1) & (c[:, 0] < 3.2) & (c[:, 1] > 5.8))
        ops[A] = [A.Rate_operator(d, rate=lambda t: 0.01 * (1.0 + t), factor=0.5, radius=2.0, center=[5.0, 3.0]),
                  A.Inlet_operator(d, A.Region(d, indices=ids), Q=1.7, label="feed"),
                  A.Boyd_box_operator(d, losses=1.5, width=1.2, height=0.8, end_points=[[4.3, 5.3], [9.7, 5.3]],
                                      apron=0.55, enquiry_gap=0.4, label="box", verbose=False)]
    mine._dev = _HostArrays(mine)
    for d in (ref, mine):
        d.timestep = 0.11
        d.yieldstep = 1.0
    keep = ref.quantities["stage"].centroid_values.copy()
    ops[anuga][0]()          # (fills the reference's report; on the device this operator is a kernel, not a call)
    ref.quantities["stage"].centroid_values[:] = keep
    for a, b in zip(ops[anuga][1:], ops[ab][1:]):
        a()
        b()
    ops[ab][2].refresh()     # the reference's getters read the live arrays: re-read ours from the "device"
    rate_r, rate_m = ops[anuga][0], ops[ab][0]
    assert np.isclose(rate_r.get_Q(), rate_m.get_Q(), rtol=1e-14) and rate_r.get_factor() == rate_m.get_factor()
    # "label: Min rate = .., Max rate = .., Total Q = .." - the figures after the label agree
    tail = lambda s: s.split(":", 1)[1]
    assert tail(rate_r.timestepping_statistics()) == tail(rate_m.timestepping_statistics())
    inl_r, inl_m = ops[anuga][1], ops[ab][1]
    assert inl_r.get_Q() == inl_m.get_Q() == 1.7 and inl_r.get_applied_Q() == inl_m.get_applied_Q()
    assert inl_r.get_total_applied_volume() == inl_m.get_total_applied_volume()
    body = lambda s: s.split("\n", 3)[3]             # (after the header that carries the label's counter)
    assert body(inl_r.timestepping_statistics()) == body(inl_m.timestepping_statistics())
    box_r, box_m = ops[anuga][2], ops[ab][2]
    for name in ("get_enquiry_stages", "get_enquiry_depths", "get_enquiry_xmoms", "get_enquiry_ymoms",
                 "get_enquiry_elevations", "get_enquiry_water_depths", "get_enquiry_invert_elevations",
                 "get_enquiry_speeds", "get_enquiry_velocity_heads", "get_enquiry_total_energys",
                 "get_enquiry_specific_energys", "get_enquiry_xvelocitys", "get_enquiry_yvelocitys"):
        assert np.allclose(getattr(box_r, name)(), getattr(box_m, name)(), rtol=1e-14, atol=0.0), name
    for name in ("get_culvert_length", "get_culvert_width", "get_culvert_height", "get_culvert_apron",
                 "get_culvert_slope", "get_culvert_blockage", "get_culvert_barrels", "get_master_proc"):
        assert getattr(box_r, name)() == getattr(box_m, name)(), name
    for k in (0, 1):
        for name in ("get_average_speed", "get_average_velocity_head", "get_average_total_energy",
                     "get_average_specific_energy"):
            assert np.isclose(getattr(box_r.inlets[k], name)(), getattr(box_m.inlets[k], name)(), rtol=1e-13), name
    assert box_r.timestepping_statistics() == box_m.timestepping_statistics()
    assert box_m.discharge_abs_timemean == 0.0
    box_r.print_timestepping_statistics()
    out_r = capsys.readouterr().out
    box_m.print_timestepping_statistics()
    out_m = capsys.readouterr().out
    assert out_r[out_r.index("Type:"):] == out_m[out_m.index("Type:"):]
    upto = lambda s: body(s)[:body(s).index("region")]           # (then the Region objects print themselves)
    assert upto(box_r.statistics()) == upto(box_m.statistics())
    mine.print_operator_timestepping_statistics()
    mine.print_operator_statistics()
    assert "Inlet report for feed" in capsys.readouterr().out
