"""CPU: the oracle is pinned.

1. the C port (oracle/sw_oracle.c) and the reference's own C sources (oracle/_ref, compiled
   from /root/reference) give bit-identical results on every golden case;
2. both reproduce the fixtures generated from the unmodified Python reference
   (multiprocessor_mode 2) bit for bit: state after 1 step, at finaltime, the timestep
   sequence, the boundary-flux and fractional-step integrals;
3. the 8-digit expected values embedded in the reference's own unit test hold.
"""
import os

import numpy as np
import pytest

import anuga_core_b200 as ab
from anuga_core_b200.workloads import domain_to_scenario
from golden_util import cases, load, rel_err
from oracle.driver import LIBS, OracleDomain

BACKENDS = ["port"] + (["ref"] if os.path.exists(LIBS["ref"]) else [])


def run_oracle(name, backend):
    builder, ev = cases.CASES[name]
    d = builder(ab)
    o = OracleDomain(domain_to_scenario(d), backend=backend)
    first = {}
    orig = o.apply_fractional_steps

    def hook():
        orig()
        if not first:
            first.update(stage=o.stage_c.copy(), xmom=o.xmom_c.copy(), ymom=o.ymom_c.copy())
    o.apply_fractional_steps = hook
    yields = [t for t in o.evolve(**ev)]
    return o, first, yields


@pytest.mark.parametrize("backend", BACKENDS)
@pytest.mark.parametrize("name", sorted(cases.CASES))
def test_oracle_reproduces_python_reference_bit_for_bit(oracle_libs, name, backend):
    g = load(name)
    o, first, yields = run_oracle(name, backend)
    assert np.array_equal(np.array(o.timestep_history), g["dts"])
    assert np.array_equal(np.array(yields), g["yields"])
    for k, a in (("stage", first["stage"]), ("xmom", first["xmom"]), ("ymom", first["ymom"])):
        assert np.array_equal(a, g["step1_" + k]), k
    assert np.array_equal(o.stage_c, g["final_stage"])
    assert np.array_equal(o.xmom_c, g["final_xmom"])
    assert np.array_equal(o.ymom_c, g["final_ymom"])
    assert np.array_equal(o.stage_e, g["final_stage_edge"])
    assert np.array_equal(o.xmom_v, g["final_xmom_vertex"])
    # sums over edges / cells: the reference reduces them with `omp reduction(+)`, so only the
    # order of additions (not the terms) may differ between runs of the reference itself
    assert abs(o.boundary_flux_integral - g["bfi"][0]) <= 1e-9 * abs(g["bfi"][0]) + 1e-15
    assert abs(o.fractional_step_volume_integral - g["fsvi"][0]) <= 1e-12 * abs(g["fsvi"][0]) + 1e-15


def test_reference_unit_test_expected_values(oracle_libs):
    """test_shallow_water_domain.py:5913 W_EX (first entries), rtol of num.allclose"""
    o, _, _ = run_oracle("kat_bedslope_more_steps", "port")
    assert np.allclose(o.stage_c[:6], cases.KAT_BEDSLOPE_W_EX_HEAD)


@pytest.mark.parametrize("backend", ["port", "ref"])
@pytest.mark.parametrize("key", ["one_step", "two_steps", "more_steps"])
def test_known_answers_embedded_in_the_reference_tests(oracle_libs, key, backend):
    """the expected arrays of test_bedslope_problem_second_order_{one_step,two_steps,more_steps}
    (test_shallow_water_domain.py:5626, 5717, 5913; tests/golden/make_golden_kat.py), with the
    tolerance of the num.allclose those tests use"""
    if backend == "ref" and not os.path.exists(LIBS["ref"]):
        pytest.skip("oracle/_ref not built")
    k = load("kat_reference_tests")
    d = cases.kat_bedslope_more_steps(ab)
    o = OracleDomain(domain_to_scenario(d), backend=backend)
    ys, ft = k[key + "_evolve"]
    for _ in o.evolve(yieldstep=float(ys), finaltime=float(ft)):
        pass
    assert np.allclose(o.stage_c, k[key + "_W_EX"])
    if key + "_UH_EX" in k.files:
        assert np.allclose(o.xmom_c, k[key + "_UH_EX"]) and np.allclose(o.ymom_c, k[key + "_VH_EX"])
    if key + "_recorded_min_timestep" in k.files:
        assert np.allclose(o.recorded_min_timestep, k[key + "_recorded_min_timestep"][0])
        assert np.allclose(o.recorded_max_timestep, k[key + "_recorded_max_timestep"][0])


@pytest.mark.skipif(not os.path.exists(LIBS["ref_fma"]), reason="oracle/_ref not built")
def test_fma_build_of_the_reference_is_only_close(oracle_libs):
    """The reference's timing build (FMA contraction) differs from its own parity build at
    1e-10..1e-8 after O(100) steps: the parity gates need reproducible arithmetic."""
    g = load("dam_break_de1")
    o, _, _ = run_oracle("dam_break_de1", "ref_fma")
    e = rel_err(o.stage_c, g["final_stage"])
    assert 0.0 < e < 1e-6


def test_protect_and_fix_negative_edge_cases(oracle_libs):
    """stage below bed: protect lifts it and reports the added mass; update into negative depth
    is clipped by fix_negative_cells for full cells only."""
    d = ab.rectangular_cross_domain(2, 2)
    d.set_flow_algorithm("DE1")
    d.set_quantity("elevation", 0.0)
    d.set_quantity("stage", -0.5, location="centroids")
    d.set_quantity("xmomentum", 1.0, location="centroids")
    d.set_quantity("ymomentum", 1.0, location="centroids")
    B = ab.Reflective_boundary(d)
    d.set_boundary({t: B for t in d.get_boundary_tags()})
    o = OracleDomain(domain_to_scenario(d), backend="port")
    me = o.protect()
    assert np.isclose(me, 0.5 * d.areas.sum())
    assert np.all(o.stage_c == 0.0) and np.all(o.xmom_c == 0.0)
    assert np.all(o.ymom_c == 1.0)            # the reference never zeroes ymom here (quirk 2)
    o.extrapolate()
    assert np.all(o.ymom_c == 0.0)
    o.stage_c[:] = -1.0
    o.tri_full_flag[0] = 0
    n = o.fn["fix_negative_cells"](__import__("ctypes").byref(o.D))
    assert n == len(o.stage_c) - 1 and o.stage_c[0] == -1.0 and np.all(o.stage_c[1:] == 0.0)


@pytest.mark.skipif(not os.path.exists(LIBS["ref"]), reason="oracle/_ref not built")
def test_riverwall_port_equals_reference_c_code(oracle_libs):
    """the weir branch (sw_domain_openmp.c:324-426, 582-653) of the port against the reference's
    own C code on a synthetic wall (real riverwall meshes need meshpy, absent here)"""
    import importlib.util
    spec = importlib.util.spec_from_file_location(
        "tgp", os.path.join(os.path.dirname(os.path.abspath(__file__)), "test_gpu_parity.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    d = m.riverwall_domain("DE1")
    res = {}
    for be in ("port", "ref"):
        o = OracleDomain(domain_to_scenario(d), backend=be)
        for _ in o.evolve(yieldstep=0.5, finaltime=2.0):
            pass
        res[be] = o
    assert len(res["port"].timestep_history) == len(res["ref"].timestep_history) > 20
    for k in ("stage_c", "xmom_c", "ymom_c", "max_speed"):
        assert np.array_equal(getattr(res["port"], k), getattr(res["ref"], k)), k
    wet_right = (res["ref"].stage_c - res["ref"].bed_c)[d.centroid_coordinates[:, 0] > 6.5].max()
    assert wet_right > 1e-3          # the wall was overtopped: the weir law was exercised


def _random_domain(A, seed):
    """a small randomly configured scenario, written against the shared API (both packages)"""
    rng = np.random.default_rng(seed)
    n = int(rng.integers(6, 11))
    L = float(n)
    d = A.rectangular_cross_domain(n, n, len1=L, len2=L)
    d.set_flow_algorithm(str(rng.choice(["DE0", "DE1", "DE2", "DE0_7", "DE1_7"])))
    d.set_store(False)
    a, b, c0 = rng.uniform(-0.2, 0.2), rng.uniform(-0.1, 0.1), rng.uniform(0.0, 0.4)
    kx, ky = rng.uniform(0.3, 1.2, size=2)
    d.set_quantity("elevation", lambda x, y: a * x + b * y + c0 * (x / L) * (1 - y / L) * 4 + 0.05 * (kx * x) * (ky * y) / L)
    level = float(rng.uniform(-0.3, 0.6))
    bump = float(rng.uniform(0.1, 0.6))
    cx, cy = rng.uniform(0.2 * L, 0.8 * L, size=2)
    d.set_quantity("stage", lambda x, y: level + bump / (1.0 + ((x - cx) ** 2 + (y - cy) ** 2)), location="centroids")
    d.set_quantity("xmomentum", lambda x, y: 0.05 * (y - cy) / L, location="centroids")
    d.set_quantity("friction", float(rng.choice([0.0, 0.02, 0.05])))
    d.set_low_froude(int(rng.integers(0, 3)))
    if rng.random() < 0.4:
        d.set_sloped_mannings_function(True)
    Br = A.Reflective_boundary(d)
    pool = [Br, A.Dirichlet_boundary([level + 0.1, 0.0, 0.0]), A.Transmissive_boundary(d),
            A.Transmissive_stage_zero_momentum_boundary(d)]
    d.set_boundary({t: pool[int(rng.integers(0, len(pool)))] for t in sorted(d.get_boundary_tags())})
    if rng.random() < 0.5:
        A.Rate_operator(d, rate=float(rng.uniform(-0.02, 0.03)))
    return d


@pytest.mark.parametrize("seed", list(range(1, 17)))
def test_oracle_equals_live_python_reference_on_random_scenarios(oracle_libs, seed):
    """fuzzing the pin: randomly configured scenarios (algorithm, bed, wet/dry level, boundaries, friction
    form, low-Froude mode, rain or drain) run by the unmodified Python reference in this process and by the
    oracle - bit for bit"""
    from oracle import pyref
    if not pyref.available():
        pytest.skip("python reference not built (oracle/build_pyref.py)")
    anuga = pyref.import_anuga()
    ref = _random_domain(anuga, seed)
    ref.set_multiprocessor_mode(2)
    ev = dict(yieldstep=0.4, finaltime=1.2)
    dts = []
    orig = ref.apply_fractional_steps

    def hook():
        orig()
        dts.append(ref.timestep)
    ref.apply_fractional_steps = hook
    for _ in ref.evolve(**ev):
        pass
    o = OracleDomain(domain_to_scenario(_random_domain(ab, seed)), backend="port")
    for _ in o.evolve(**ev):
        pass
    assert len(dts) >= 3 and np.array_equal(np.array(o.timestep_history), np.array(dts))
    q = ref.quantities
    assert np.array_equal(o.stage_c, q["stage"].centroid_values)
    assert np.array_equal(o.xmom_c, q["xmomentum"].centroid_values)
    assert np.array_equal(o.ymom_c, q["ymomentum"].centroid_values)


EDGE_CASES = 6


def _edge_case_domain(A, k):
    """the smallest and the degenerate inputs: a 1 x 1 mesh (4 triangles, every one on the boundary), strips one
    cell wide, a completely dry bed with and without rain, stage below the bed, more triangles than a block holds
    by one row"""
    m, n, alg = [(1, 1, "DE1"), (1, 1, "DE0"), (2, 1, "DE1"), (3, 2, "DE1"), (5, 3, "DE2"), (1, 3, "DE1_7")][k]
    d = A.rectangular_cross_domain(m, n, len1=float(m), len2=float(n))
    d.set_flow_algorithm(alg)
    d.set_store(False)
    d.set_quantity("elevation", lambda x, y: 0.1 * x - 0.05 * y)
    d.set_quantity("friction", 0.03)
    Br = A.Reflective_boundary(d)
    bmap = {t: Br for t in d.get_boundary_tags()}
    if k == 0:
        d.set_quantity("stage", lambda x, y: 0.5 + 0.2 * x, location="centroids")
    elif k == 1:
        d.set_quantity("stage", 0.4, location="centroids")
        bmap["left"] = A.Dirichlet_boundary([0.7, 0.1, 0.0])
    elif k == 2:         # dry bed, rain wets it
        d.set_quantity("stage", expression="elevation")
        A.Rate_operator(d, rate=0.05)
    elif k == 3:         # dry bed, nothing happens: the timestep is the maximal one, clipped by the yieldstep
        d.set_quantity("stage", expression="elevation")
    elif k == 4:         # stage below the bed in places (protect raises it), open boundaries
        d.set_quantity("stage", lambda x, y: 0.15 + 0.0 * x - 0.1 * (y > 1.5), location="centroids")
        bmap["right"] = A.Transmissive_boundary(d)
        bmap["top"] = A.Transmissive_stage_zero_momentum_boundary(d)
    else:
        d.set_quantity("stage", lambda x, y: 0.3 + 0.1 * y, location="centroids")
        d.set_quantity("ymomentum", 0.05, location="centroids")
        bmap["top"] = A.Transmissive_stage_zero_momentum_boundary(d)
    d.set_boundary(bmap)
    return d


@pytest.mark.parametrize("k", list(range(EDGE_CASES)))
def test_oracle_equals_live_python_reference_on_edge_cases(oracle_libs, k):
    """the pin on the smallest / degenerate inputs: the unmodified Python reference in this process and the oracle,
    bit for bit (timestep sequence and state)"""
    from oracle import pyref
    if not pyref.available():
        pytest.skip("python reference not built (oracle/build_pyref.py)")
    anuga = pyref.import_anuga()
    ref = _edge_case_domain(anuga, k)
    ref.set_multiprocessor_mode(2)
    ev = dict(yieldstep=0.25, finaltime=0.75)
    dts = []
    orig = ref.apply_fractional_steps

    def hook():
        orig()
        dts.append(ref.timestep)
    ref.apply_fractional_steps = hook
    for _ in ref.evolve(**ev):
        pass
    o = OracleDomain(domain_to_scenario(_edge_case_domain(ab, k)), backend="port")
    for _ in o.evolve(**ev):
        pass
    assert np.array_equal(np.array(o.timestep_history), np.array(dts)), (o.timestep_history, dts)
    q = ref.quantities
    assert np.array_equal(o.stage_c, q["stage"].centroid_values)
    assert np.array_equal(o.xmom_c, q["xmomentum"].centroid_values)
    assert np.array_equal(o.ymom_c, q["ymomentum"].centroid_values)


import merimbula_case  # noqa: E402


def test_merimbula_lake_oracle_reproduces_golden_fixture(oracle_libs):
    """a real unstructured mesh (the reference's Merimbula lake example, 10 785 triangles, node valences 1-9),
    rebuilt from the arrays in the fixture: mesh.py, the neighbour builder and the oracle's time loop give the
    Python reference's timestep sequence and state bit for bit"""
    g = load("merimbula_de1")
    d = merimbula_case.from_fixture(ab, g)
    o = OracleDomain(domain_to_scenario(d), backend="port")
    first = {}
    orig = o.apply_fractional_steps

    def hook():
        orig()
        if not first:
            first.update(stage=o.stage_c.copy(), xmom=o.xmom_c.copy(), ymom=o.ymom_c.copy())
    o.apply_fractional_steps = hook
    yields = [t for t in o.evolve(**merimbula_case.EVOLVE)]
    assert np.array_equal(np.array(yields), g["yields"])
    assert np.array_equal(np.array(o.timestep_history), g["dts"])
    for k in ("stage", "xmom", "ymom"):
        assert np.array_equal(first[k], g["step1_" + k]), k
    assert np.array_equal(o.stage_c, g["final_stage"])
    assert np.array_equal(o.xmom_c, g["final_xmom"]) and np.array_equal(o.ymom_c, g["final_ymom"])


def test_merimbula_fixture_mesh_equals_the_tsh_file():
    """the mesh stored in the fixture is what create_domain_from_file reads (when the reference tree is here)"""
    if not os.path.exists(merimbula_case.TSH):
        pytest.skip("needs the reference tree")
    g = load("merimbula_de1")
    a = merimbula_case.from_fixture(ab, g)
    b = ab.create_domain_from_file(merimbula_case.TSH)
    assert np.array_equal(a.nodes, b.nodes) and np.array_equal(a.triangles, b.triangles)
    assert a.mesh.boundary == b.mesh.boundary
    assert np.array_equal(a.quantities["elevation"].centroid_values, b.quantities["elevation"].centroid_values)
