"""GPU: the PER-CALL layer (host arrays in / host arrays out), one entry point per function the
reference's Cython FFI binds (sw_domain_openmp_ext.pyx:371-459, quantity_ext.pyx:38-117),
checked against the reference's own C code on the same arrays."""
import ctypes as C
import os

import numpy as np
import pytest

import anuga_core_b200 as ab
from anuga_core_b200 import backend as B
from anuga_core_b200.workloads import domain_to_scenario
from golden_util import cases, rel_err
from oracle.driver import LIBS, OracleDomain, _pd

pytestmark = pytest.mark.gpu
REF = "ref" if os.path.exists(LIBS["ref"]) else "port"


def host_view(o, d):
    """SwkHostView over the oracle domain's numpy arrays (they play the reference Domain's arrays)."""
    v = B.SwkHostView()
    keep = B._MeshArrays(d._mesh_dict())
    v.mesh = keep.struct
    v.params = B.make_params(d._param_dict())
    m = {"stage_centroid_values": o.stage_c, "xmom_centroid_values": o.xmom_c, "ymom_centroid_values": o.ymom_c,
         "bed_centroid_values": o.bed_c, "height_centroid_values": o.height_c, "friction_centroid_values": o.friction_c,
         "stage_edge_values": o.stage_e, "xmom_edge_values": o.xmom_e, "ymom_edge_values": o.ymom_e,
         "bed_edge_values": o.bed_e, "height_edge_values": o.height_e,
         "stage_vertex_values": o.stage_v, "xmom_vertex_values": o.xmom_v, "ymom_vertex_values": o.ymom_v,
         "bed_vertex_values": o.bed_v, "height_vertex_values": o.height_v,
         "stage_boundary_values": o.stage_b, "xmom_boundary_values": o.xmom_b, "ymom_boundary_values": o.ymom_b,
         "stage_explicit_update": o.stage_eu, "xmom_explicit_update": o.xmom_eu, "ymom_explicit_update": o.ymom_eu,
         "stage_semi_implicit_update": o.stage_siu, "xmom_semi_implicit_update": o.xmom_siu,
         "ymom_semi_implicit_update": o.ymom_siu, "max_speed": o.max_speed, "boundary_flux_sum": o.boundary_flux_sum}
    for k, a in m.items():
        setattr(v, k, _pd(a))
    return v, keep


def twin(n=12, alg="DE1"):
    d = cases.beach_de1(ab, n=n)
    d.set_flow_algorithm(alg)
    sc = domain_to_scenario(d)
    ref = OracleDomain(sc, backend=REF)
    mine = OracleDomain(sc, backend=REF)      # only its ARRAYS are used; kernels are the CUDA ones
    for o in (ref, mine):
        for _ in o.evolve(yieldstep=0.3, finaltime=0.3):
            pass
        o.stage_c[::7] -= 0.9                 # some cells below the bed
    return d, ref, mine


def test_per_call_sequence_matches_reference_functions():
    lib = B.load_library()
    d, ref, mine = twin()
    v, keep = host_view(mine, d)
    h = B._H()
    B._check(lib.swk_call_open(C.byref(v), 0, C.byref(h)))
    try:
        # protect_new
        me = C.c_double()
        B._check(lib.swk_call_protect_new(h, C.byref(v), C.byref(me)))
        me_ref = ref.protect()
        assert np.isclose(me.value, me_ref, rtol=1e-12) and me_ref > 0
        assert np.array_equal(mine.stage_c, ref.stage_c) and np.array_equal(mine.xmom_c, ref.xmom_c)
        assert np.array_equal(mine.stage_v, ref.stage_v)
        # extrapolate_second_order_edge_sw
        B._check(lib.swk_call_extrapolate_second_order_edge_sw(h, C.byref(v)))
        ref.extrapolate()
        for name in ("stage_e", "xmom_e", "ymom_e", "height_e", "bed_e", "stage_v", "xmom_v", "ymom_v",
                     "height_v", "bed_v", "xmom_c", "ymom_c", "height_c"):
            assert rel_err(getattr(mine, name), getattr(ref, name)) <= 1e-12, name
        # boundary values computed by the host (numpy) as in the reference, then compute_fluxes
        ref.update_boundary()
        mine.update_boundary()
        for substep in (0, 1):
            ft = C.c_double()
            B._check(lib.swk_call_compute_fluxes_ext_central(h, C.byref(v), 1000.0, substep, C.byref(ft)))
            ref.compute_fluxes(substep)
            assert ft.value == ref.flux_timestep
            for name in ("stage_eu", "xmom_eu", "ymom_eu", "max_speed"):
                assert rel_err(getattr(mine, name), getattr(ref, name)) <= 1e-12, (name, substep)
            assert abs(mine.boundary_flux_sum[substep] - ref.boundary_flux_sum[substep]) <= 1e-9
        # fix_negative_cells
        for o in (ref, mine):
            o.stage_c[::5] = o.bed_c[::5] - 0.1
        n = C.c_int64()
        B._check(lib.swk_call_fix_negative_cells(h, C.byref(v), C.byref(n)))
        n_ref = ref.fn["fix_negative_cells"](C.byref(ref.D))
        assert n.value == n_ref > 0
        assert np.array_equal(mine.stage_c, ref.stage_c) and np.array_equal(mine.ymom_c, ref.ymom_c)
    finally:
        lib.swk_call_close(h)


def test_per_call_flux_takes_any_edge_arrays_like_the_reference_entry_point():
    """compute_fluxes_ext_central (sw_domain_openmp_ext.pyx:371) accepts whatever the arrays hold: edge beds
    that are not stage_e - height_e and centroid heights that are not max(stage - bed, 0) are used as given"""
    lib = B.load_library()
    d, ref, mine = twin(n=8)
    for o in (ref, mine):
        o.distribute_to_vertices_and_edges()
        o.update_boundary()
        o.bed_e[::3, 0] += 0.05               # user-modified edge beds
        o.bed_e[1::4, 2] -= 0.02
        o.height_c[::5] *= 1.1                # ... and centroid heights
    v, keep = host_view(mine, d)
    h = B._H()
    B._check(lib.swk_call_open(C.byref(v), 0, C.byref(h)))
    try:
        ft = C.c_double()
        B._check(lib.swk_call_compute_fluxes_ext_central(h, C.byref(v), 1000.0, 0, C.byref(ft)))
        ref.compute_fluxes(0)
        assert ft.value == ref.flux_timestep
        for name in ("stage_eu", "xmom_eu", "ymom_eu", "max_speed"):
            assert rel_err(getattr(mine, name), getattr(ref, name)) <= 1e-12, name
        assert np.any(ref.xmom_eu != 0.0)
    finally:
        lib.swk_call_close(h)


@pytest.mark.parametrize("sloped", [False, True])
def test_manning_friction_entry_points(sloped):
    lib = B.load_library()
    rng = np.random.default_rng(1234)
    N = 5000
    w = rng.uniform(0.0, 2.0, N)
    z = rng.uniform(-0.5, 1.0, N)
    uh, vh = rng.normal(size=N), rng.normal(size=N)
    eta = np.where(rng.uniform(size=N) < 0.1, 0.0, 0.03)
    zv = np.repeat(z[:, None], 3, axis=1) + rng.normal(scale=0.05, size=(N, 3))
    x = rng.uniform(size=(N, 6))
    x[:, 2] += 1.0
    x[:, 5] += 1.0
    from oracle.driver import load
    fn = load(REF)
    xu_r, yu_r = rng.normal(size=N), rng.normal(size=N)
    xu, yu = xu_r.copy(), yu_r.copy()
    if sloped:
        fn["manning_friction_sloped"](9.8, 1e-5, N, _pd(x), _pd(w), _pd(zv), _pd(uh), _pd(vh), _pd(eta), _pd(xu_r), _pd(yu_r))
        B._check(lib.swk_call_manning_friction_sloped(0, 9.8, 1e-5, N, _pd(x), _pd(w), _pd(zv), _pd(uh), _pd(vh), _pd(eta), _pd(xu), _pd(yu)))
    else:
        fn["manning_friction_flat"](9.8, 1e-5, N, _pd(w), _pd(z), _pd(uh), _pd(vh), _pd(eta), _pd(xu_r), _pd(yu_r))
        B._check(lib.swk_call_manning_friction_flat(0, 9.8, 1e-5, N, _pd(w), _pd(z), _pd(uh), _pd(vh), _pd(eta), _pd(xu), _pd(yu)))
    # pow(h, 7/3): correctly rounded on the device, glibc's is 1 ulp off on ~0.1 % of arguments
    assert rel_err(xu, xu_r) <= 1e-14 and rel_err(yu, yu_r) <= 1e-14
    assert np.mean(xu != xu_r) < 0.01


def test_update_backup_saxpy_entry_points():
    lib = B.load_library()
    from oracle.driver import load
    fn = load(REF)
    rng = np.random.default_rng(7)
    N = 4097
    c = rng.normal(size=N)
    c[::11] = 0.0
    eu = rng.normal(size=N)
    siu = -np.abs(rng.normal(size=N)) * np.abs(c)
    c_r, siu_r = c.copy(), siu.copy()
    assert fn["update"](N, 0.01, _pd(c_r), _pd(eu), _pd(siu_r)) == 0
    B._check(lib.swk_call_update(0, N, 0.01, _pd(c), _pd(eu), _pd(siu)))
    assert np.array_equal(c, c_r) and np.all(siu == 0.0)
    # denominator <= 0 -> the reference returns -1, we return SWK_ERR_DENOMINATOR
    c2 = np.ones(8)
    s2 = np.full(8, 1000.0)
    assert lib.swk_call_update(0, 8, 0.01, _pd(c2), _pd(np.zeros(8)), _pd(s2)) == -3
    bk = np.zeros(N)
    B._check(lib.swk_call_backup_centroid_values(0, N, _pd(c), _pd(bk)))
    assert np.array_equal(bk, c)
    a = rng.normal(size=N)
    a_r = a.copy()
    fn["saxpy_centroid_values"](N, 0.25, 0.75, _pd(a_r), _pd(bk))
    B._check(lib.swk_call_saxpy_centroid_values(0, N, 0.25, 0.75, _pd(a), _pd(bk)))
    assert np.array_equal(a, a_r)
    # empty input
    B._check(lib.swk_call_update(0, 0, 0.01, _pd(c), _pd(eu), _pd(siu)))
