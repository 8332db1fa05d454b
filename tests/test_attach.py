"""The drop-in for a reference Domain (anuga_core_b200/attach.py).
CPU part: needs the scratch build of the Python reference (skipped elsewhere).
GPU part: the adapter driven through an object with the reference's attribute names."""
import numpy as np
import pytest

import anuga_core_b200 as ab
from golden_util import cases, load, rel_err
from oracle import pyref


@pytest.mark.skipif(not pyref.available(), reason="python reference not built (oracle/build_pyref.py)")
def test_adapter_snapshots_a_real_reference_domain():
    anuga = pyref.import_anuga()
    ref = cases.tsunami_set_stage(anuga, n=8)
    anuga.Rate_operator(ref, rate=2.0e-4, factor=0.5)
    iface = ab.B200_interface(ref)
    d = iface.dev_domain
    # same buffers, not copies
    assert d.quantities["stage"].centroid_values is ref.quantities["stage"].centroid_values
    assert d.quantities["xmomentum"].edge_values is ref.quantities["xmomentum"].edge_values
    assert d.mesh is ref.mesh
    # scalars of the DE1 preset are read from the reference object
    assert (d.CFL, d.timestepping_method, d.minimum_allowed_height, d.beta_w, d.H0) == (1.0, "rk2", 1e-5, 1.0, 1e-5)
    kinds = {t: (None if B is None else B.device_kind) for t, B in d.boundary_map.items()}
    assert kinds == {"left": 4, "right": 3, "top": 1, "bottom": 1}
    assert d._needs_host_stepping()
    assert len(d.fractional_step_operators) == 1 and d.fractional_step_operators[0].rate == 2.0e-4
    assert d.fractional_step_operators[0].factor == 0.5
    # mesh arrays handed to the C ABI are the reference's own
    m = d._mesh_dict()
    assert m["neighbours"] is ref.mesh.neighbours and m["normals"] is ref.mesh.normals
    # a later set_flow_algorithm is picked up at the next evolve (mode 4 of the reference goes stale)
    ref.set_flow_algorithm("DE0")
    iface.refresh()
    assert d.timestepping_method == "euler" and d.CFL == 0.9
    if ab.device_count() == 0:
        with pytest.raises(ab.SwkError):
            ab.set_multiprocessor_mode_b200(ref)


@pytest.mark.skipif(not pyref.available(), reason="python reference not built (oracle/build_pyref.py)")
def test_adapter_rejects_what_it_cannot_run():
    anuga = pyref.import_anuga()
    ref = cases.dam_break_de0(anuga, n=4)
    ref.boundary_map["left"] = type("Inflow_boundary", (object,), {})()
    with pytest.raises(NotImplementedError):
        ab.B200_interface(ref)


@pytest.mark.skipif(not pyref.available(), reason="python reference not built (oracle/build_pyref.py)")
@pytest.mark.parametrize("name", ["inlet_de1", "culvert_de1", "culvert_pipe_de1"])
def test_adapter_adopts_reference_structures(name):
    """Inlet_operator / Boyd operators of a real reference domain are taken over with their resolved
    geometry, parameters and smoothing memory"""
    anuga = pyref.import_anuga()
    ref = cases.CASES[name][0](anuga)
    iface = ab.B200_interface(ref)
    mine = [op for op in iface.dev_domain.fractional_step_operators if getattr(op, "host_side", False)]
    theirs = [op for op in ref.fractional_step_operators if type(op).__name__ != "boundary_flux_integral_operator"]
    assert len(mine) == len(theirs) > 0
    for a, b in zip(mine, theirs):
        assert type(a).__name__ == type(b).__name__
        if hasattr(b, "inlets"):
            for ia, ib in zip(a.inlets, b.inlets):
                assert np.array_equal(ia.triangle_indices, ib.triangle_indices)
                assert ia.enquiry_index == ib.enquiry_index
            assert a.smooth_Q == b.smooth_Q and a.culvert_length == b.culvert_length
        else:
            assert np.array_equal(a.inlet.triangle_indices, b.inlet.triangle_indices)


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["beach_de1", "inlet_de1", "culvert_de1", "culvert_pipe_de1"])
def test_adapter_time_loop_matches_golden(name):
    """evolve through the adapter (host arrays updated in place) == the golden reference run"""
    g = load(name)
    builder, ev = cases.CASES[name]
    ref_like = builder(ab)                       # carries the reference's attribute names
    stage_alias = ref_like.quantities["stage"].centroid_values
    iface = ab.B200_interface(ref_like)
    times = [t for t in iface.evolve_base(**ev)]
    assert np.array_equal(np.array(times), g["yields"])
    assert rel_err(stage_alias, g["final_stage"]) <= 1e-9          # the alias saw the result
    assert rel_err(ref_like.quantities["xmomentum"].centroid_values, g["final_xmom"]) <= 1e-9
    assert ref_like.timestep == g["dts"][-1]


@pytest.mark.gpu
def test_adapter_friction_methods_of_mode_4():
    """gpu_interface.compute_forcing_terms_manning_friction_flat / _sloped (friction.py:133-147)"""
    d = cases.dam_break_de1(ab)
    q = d.quantities
    q["xmomentum"].set_values(lambda x, y: 0.3 + 0.01 * x, location="centroids")
    q["ymomentum"].set_values(lambda x, y: -0.2 + 0.02 * y, location="centroids")
    iface = ab.B200_interface(d)
    q["xmomentum"].semi_implicit_update[:] = 0.0
    q["ymomentum"].semi_implicit_update[:] = 0.0
    iface.compute_forcing_terms_manning_friction_flat()
    w, z = q["stage"].centroid_values, q["elevation"].centroid_values
    uh, vh, eta = q["xmomentum"].centroid_values, q["ymomentum"].centroid_values, q["friction"].centroid_values
    h = w - z
    S = -d.g * eta ** 2 * np.sqrt(uh ** 2 + vh ** 2) / h ** (7.0 / 3.0)
    assert np.all(h > d.minimum_allowed_height)
    assert rel_err(q["xmomentum"].semi_implicit_update, S * uh) <= 1e-14
    assert rel_err(q["ymomentum"].semi_implicit_update, S * vh) <= 1e-14
    flat = q["xmomentum"].semi_implicit_update.copy()
    q["xmomentum"].semi_implicit_update[:] = 0.0
    q["ymomentum"].semi_implicit_update[:] = 0.0
    iface.compute_forcing_terms_manning_friction_sloped()
    assert np.all(np.abs(q["xmomentum"].semi_implicit_update) >= np.abs(flat) * (1 - 1e-14))   # x sqrt(1 + |grad z|^2)


@pytest.mark.gpu
def test_second_evolve_call_does_not_apply_rate_operators_twice():
    """refresh() rebuilds the operator list at every evolve(): the device copies of the previous call must
    go, or rain is applied once more per call.  Two evolve() calls through the adapter == one plain run."""
    def build():
        d = cases.dam_break_de1(ab)
        ab.Rate_operator(d, rate=3.0e-3, factor=2.0)
        ab.Rate_operator(d, rate=lambda t: 1.0e-3 * (1.0 + t))
        return d
    plain = build()
    for _ in plain.evolve(yieldstep=0.1, finaltime=0.4):
        pass
    ref_like = build()
    iface = ab.B200_interface(ref_like)
    for _ in iface.evolve_base(yieldstep=0.1, finaltime=0.2):
        pass
    for _ in iface.evolve_base(yieldstep=0.1, finaltime=0.4):
        pass
    for name in ("stage", "xmomentum", "ymomentum"):
        assert np.array_equal(ref_like.quantities[name].centroid_values, plain.quantities[name].centroid_values), name
    assert iface.dev_domain.fractional_step_volume_integral == plain.fractional_step_volume_integral


@pytest.mark.gpu
def test_set_factor_reaches_an_operator_that_is_already_on_the_device():
    """rate_operators.py:160 reads self.factor at every call"""
    def run(change):
        d = cases.dam_break_de1(ab)
        op = ab.Rate_operator(d, rate=2.0e-3, factor=1.0)
        for t in d.evolve(yieldstep=0.1, finaltime=0.3):
            if change and abs(t - 0.1) < 1e-12:
                op.set_factor(5.0)
        return d
    a, b = run(False), run(True)
    va, vb = a.fractional_step_volume_integral, b.fractional_step_volume_integral
    assert vb > 2.0 * va


@pytest.mark.gpu
@pytest.mark.skipif(not pyref.available(), reason="python reference not built (oracle/build_pyref.py -> baseline/_ref)")
@pytest.mark.parametrize("name", ["beach_de1", "rain_de1", "tsunami_set_stage", "dam_break_de2"])
def test_real_reference_domain_on_the_device_equals_mode_2(name):
    """INTEGRATION.md route A: a REAL anuga.shallow_water.Domain evolves through
    set_multiprocessor_mode_b200 (the reference's own Domain.evolve wrapper around the device time loop)
    and is compared, in the same process, with a twin that runs the reference's mode 2 (C/OpenMP); two
    evolve() calls each (shallow_water_domain.py:2300, 2859-2899)."""
    anuga = pyref.import_anuga()
    builder, ev = cases.CASES[name]
    cpu = builder(anuga)
    cpu.set_multiprocessor_mode(2)
    gpu = builder(anuga)
    iface = ab.set_multiprocessor_mode_b200(gpu)
    assert gpu.multiprocessor_mode == ab.MODE_B200 and gpu.gpu_interface is iface
    half = dict(ev, finaltime=ev["yieldstep"] * max(1, int(round(ev["finaltime"] / ev["yieldstep"])) // 2))
    for stop in (half, ev):
        tc = [t for t in cpu.evolve(**stop)]
        tg = [t for t in gpu.evolve(**stop)]
        assert tc == tg
        assert gpu.timestep == cpu.timestep
        for q in ("stage", "xmomentum", "ymomentum"):
            e = rel_err(gpu.quantities[q].centroid_values, cpu.quantities[q].centroid_values)
            assert e <= 1e-9, (q, e)
    assert iface.dev_domain._dev.kernel_launch_count() > 0


def _weir_rating(hw, tw):
    head = max(hw, tw) - 0.3
    if head <= 0.0:
        return 0.0
    sub = max(min(hw, tw) - 0.3, 0.0) / head
    q = 1.2 * head ** 1.5 * (1.0 - sub ** 1.5) ** 0.385
    return q if hw >= tw else -q


@pytest.mark.gpu
@pytest.mark.skipif(not pyref.available(), reason="python reference not built (oracle/build_pyref.py -> baseline/_ref)")
@pytest.mark.parametrize("implicit", [False, True])
def test_internal_boundary_structures_on_the_device_equal_the_reference(implicit):
    """Internal_boundary_operator with a weir rating and with a pumping station
    (structures/internal_boundary_operator.py, internal_boundary_functions.py:392-477): this package's Domain
    on the device next to the reference's Domain in mode 2, same process, two ponds and a dry embankment."""
    anuga = pyref.import_anuga()
    from anuga.structures.internal_boundary_functions import pumping_station_function as ref_pump

    def build(A, pump):
        d = cases._embankment(A)
        A.Internal_boundary_operator(d, _weir_rating, width=1.5, height=1.0, end_points=[[6.1, 8.3], [9.9, 8.3]],
                                     apron=0.55, enquiry_gap=0.4, smoothing_timescale=0.5,
                                     compute_discharge_implicitly=implicit, verbose=False)
        P = pump(d, pump_capacity=0.6, hw_to_start_pumping=0.5, hw_to_stop_pumping=0.3, initial_pump_rate=0.0,
                 pump_rate_of_increase=0.4, pump_rate_of_decrease=0.4, verbose=False)
        A.Internal_boundary_operator(d, P, width=1.0, height=1.0, end_points=[[6.1, 3.3], [9.9, 3.3]],
                                     apron=0.55, enquiry_gap=0.4, compute_discharge_implicitly=False, verbose=False)
        return d
    cpu = build(anuga, ref_pump)
    cpu.set_multiprocessor_mode(2)
    gpu = build(ab, ab.pumping_station_function)
    tc = [t for t in cpu.evolve(yieldstep=1.0, finaltime=4.0)]
    tg = [t for t in gpu.evolve(yieldstep=1.0, finaltime=4.0)]
    assert tc == tg
    for q in ("stage", "xmomentum", "ymomentum"):
        e = rel_err(gpu.quantities[q].centroid_values, cpu.quantities[q].centroid_values)
        assert e <= 1e-9, (q, e)
    structures = lambda d: [op for op in d.fractional_step_operators if hasattr(op, "accumulated_flow")]
    assert len(structures(cpu)) == len(structures(gpu)) == 2
    for a, b in zip(structures(cpu), structures(gpu)):
        assert abs(a.accumulated_flow - b.accumulated_flow) <= 1e-9 * max(1.0, abs(a.accumulated_flow))
        assert a.accumulated_flow > 0.05
    assert gpu._dev.kernel_launch_count() > 0
