"""Golden cases, written ONCE against the reference's public API.

Every builder takes ``A`` - either the reference package (``anuga``, scratch build, used by
make_golden.py in the build container) or this repository's ``anuga_core_b200`` - and uses
only names both export (Domain, rectangular, rectangular_cross_domain, *_boundary,
Rate_operator, set_flow_algorithm, set_quantity, set_boundary, evolve ...).  That the same
script drives both is the drop-in claim of INTEGRATION.md.

CASES maps name -> (builder, evolve kwargs).
"""
import numpy as np


def _reflective_all(A, d):
    Br = A.Reflective_boundary(d)
    d.set_boundary({t: Br for t in d.get_boundary_tags()})


def kat_bedslope_more_steps(A):
    """anuga/shallow_water/tests/test_shallow_water_domain.py:5913-6050
    (test_bedslope_problem_second_order_more_steps): rectangular(6,6), default DE0,
    betas 0.9, elevation -x/3, stage = elevation + 0.05."""
    points, vertices, boundary = A.rectangular(6, 6)
    d = A.Domain(points, vertices, boundary)
    d.set_store(False)
    d.smooth = False
    d.default_order = 2
    d.beta_w = 0.9
    d.beta_w_dry = 0.9
    d.beta_uh = 0.9
    d.beta_uh_dry = 0.9
    d.beta_vh = 0.9
    d.beta_vh_dry = 0.9
    d.set_low_froude(0)
    d.H0 = 0
    d.set_quantity("elevation", lambda x, y: -x / 3)
    Br = A.Reflective_boundary(d)
    d.set_boundary({"left": Br, "right": Br, "top": Br, "bottom": Br})
    d.set_quantity("stage", d.quantities["elevation"].vertex_values + 0.05)
    return d


def dam_break_de0(A, n=30):
    d = A.rectangular_cross_domain(n, n, len1=float(n), len2=float(n))
    d.set_flow_algorithm("DE0")
    d.set_store(False)
    d.set_quantity("elevation", lambda x, y: -x / (n / 2.0))
    d.set_quantity("stage", lambda x, y: np.where(x < n / 2.0, 1.0, 0.2), location="centroids")
    d.set_quantity("friction", 0.03)
    _reflective_all(A, d)
    return d


def dam_break_de0_fixed_dt(A):
    d = dam_break_de0(A, n=24)
    d.set_fixed_flux_timestep(0.01)
    return d


def dam_break_de1(A):
    d = dam_break_de0(A, n=30)
    d.set_flow_algorithm("DE1")
    return d


def dam_break_de2(A):
    d = dam_break_de0(A, n=20)
    d.set_flow_algorithm("DE2")
    return d


def beach_de1(A, n=24):
    """dam break running up a dry beach: protect, dry-cell zeroing, fix-negative"""
    d = A.rectangular_cross_domain(n, n, len1=float(n), len2=float(n))
    d.set_flow_algorithm("DE1")
    d.set_store(False)
    L = float(n)
    d.set_quantity("elevation", lambda x, y: 0.5 * (x - 0.4 * L) / (0.6 * L) * (x > 0.4 * L) + 0.05 * np.sin(y))
    d.set_quantity("stage", lambda x, y: np.where(x < 0.25 * L, 0.8, -1.0), location="centroids")
    d.set_quantity("friction", 0.03)
    _reflective_all(A, d)
    return d


def beach_de0_7(A):
    d = beach_de1(A, n=20)
    d.set_flow_algorithm("DE0_7")
    return d


def tsunami_set_stage(A, n=24):
    """configs[1] style: beach + island, set-stage inflow (time dependent), transmissive outflow"""
    d = A.rectangular_cross_domain(n, n, len1=float(n), len2=float(n))
    d.set_flow_algorithm("DE1")
    d.set_store(False)
    L = float(n)
    d.set_quantity("elevation", lambda x, y: -(10 - 9.9 * x / L) + 0.5 * np.exp(-((x - 0.7 * L) ** 2 + (y - 0.5 * L) ** 2) / (0.05 * L) ** 2))
    d.set_quantity("stage", 0.0)
    d.set_quantity("friction", 0.025)
    Br = A.Reflective_boundary(d)
    Bl = A.Transmissive_n_momentum_zero_t_momentum_set_stage_boundary(d, lambda t: 0.5 * np.sin(2 * np.pi * t / 60.0))
    d.set_boundary({"left": Bl, "right": A.Transmissive_boundary(d), "top": Br, "bottom": Br})
    return d


def tsunami_dirichlet(A, n=24):
    d = tsunami_set_stage(A, n)
    Br = A.Reflective_boundary(d)
    d.set_boundary({"left": A.Dirichlet_boundary([0.3, 0.0, 0.0]), "right": A.Transmissive_boundary(d),
                    "top": Br, "bottom": Br})
    return d


def time_boundary_de1(A, n=16):
    d = tsunami_set_stage(A, n)
    Br = A.Reflective_boundary(d)
    d.set_boundary({"left": A.Time_boundary(d, lambda t: [0.2 * np.sin(t / 3.0), 0.0, 0.0]),
                    "right": A.Transmissive_boundary(d), "top": Br, "bottom": Br})
    return d


def rain_de1(A, n=30):
    """configs[2] fields at small size with the rain Rate_operator"""
    d = A.rectangular_cross_domain(n, n, len1=float(n), len2=float(n))
    d.set_flow_algorithm("DE1")
    d.set_store(False)
    L = float(n)
    elev = lambda x, y: 0.01 * np.sin(2 * np.pi * x / 200.0) * np.cos(2 * np.pi * y / 200.0)
    d.set_quantity("elevation", elev)
    d.set_quantity("stage", lambda x, y: elev(x, y) + 0.5 + 0.1 * np.exp(-((x - L / 2) ** 2 + (y - L / 2) ** 2) / (0.1 * L) ** 2),
                   location="centroids")
    d.set_quantity("friction", 0.03)
    _reflective_all(A, d)
    A.Rate_operator(d, rate=1.0e-4)
    return d


def rain_time_de1(A):
    """rate as a function of time: evaluated at the time the operators see after an rk2 step"""
    d = beach_de1(A, n=16)
    A.Rate_operator(d, rate=lambda t: 2.0e-3 * (1.0 + np.sin(3.0 * t)), factor=0.75)
    return d


def rain_regions_de1(A):
    """rain over a polygon, extraction over a circle (operators resolve their Region)"""
    d = beach_de1(A, n=16)
    A.Rate_operator(d, rate=0.02, polygon=[[1.3, 2.2], [9.7, 1.1], [11.2, 8.4], [3.1, 12.6]])
    A.Rate_operator(d, rate=lambda t: -0.01 * (1.0 + t), center=[4.2, 5.1], radius=2.3)
    return d


def gate_de1(A):
    """Set_stage_operator with a level that varies in time over a circle, after a one-off Set_stage"""
    d = beach_de1(A, n=16)
    A.Set_stage(d, stage=0.9, center=[11.0, 4.0], radius=1.6)()
    A.Set_stage_operator(d, stage=lambda t: 0.7 + 0.2 * np.sin(2.0 * t), center=[3.2, 8.1], radius=1.7)
    return d


def rain_xyt_de1(A):
    """a rain cell that moves across the domain: rate(x, y, t), positive and (over a polygon) negative.
    Only + - * / in the callbacks: numpy's array exp / sin pick a SIMD code path by host CPU and differ
    in the last bit between machines, which a bit-level fixture cannot absorb."""
    d = beach_de1(A, n=16)
    A.Rate_operator(d, rate=lambda x, y, t: 0.02 / (1.0 + ((x - 2.0 - 3.0 * t) ** 2 + (y - 8.0) ** 2) / 6.0))
    A.Rate_operator(d, rate=lambda x, y, t: -0.01 * (1.0 + (x + t) / (1.0 + (x + t) ** 2)), factor=0.5,
                    polygon=[[1.3, 2.2], [6.7, 1.1], [7.2, 6.4], [2.1, 7.6]])
    return d


def breach_de1(A):
    """Set_elevation_operator: an embankment that is lowered over a circle as time goes on"""
    d = _embankment(A)
    A.Set_elevation_operator(d, elevation=lambda t: 1.0 - 0.3 * t, center=[8.0, 8.0], radius=1.6)
    return d


def drain_de1(A):
    """negative rate: the clamped branch of Rate_operator (rate_operators.py:213-245)"""
    d = beach_de1(A, n=16)
    A.Rate_operator(d, rate=-0.05)
    return d


def sloped_manning_de1(A):
    d = dam_break_de0(A, n=16)
    d.set_flow_algorithm("DE1")
    d.set_sloped_mannings_function(True)
    return d


def low_froude_de1(A):
    d = dam_break_de0(A, n=16)
    d.set_flow_algorithm("DE1")
    d.set_low_froude(1)
    return d


def _inlet_indices(d, x0, x1, y0, y1):
    c = d.centroid_coordinates
    return np.flatnonzero((c[:, 0] > x0) & (c[:, 0] < x1) & (c[:, 1] > y0) & (c[:, 1] < y1))


def inlet_de1(A, n=16):
    """config[4] building block: Inlet_operator with a discharge hydrograph on a sloping, partly dry
    bed (structures/inlet_operator.py), plus a second inlet that extracts water"""
    d = beach_de1(A, n=n)
    L = float(n)
    A.Inlet_operator(d, A.Region(d, indices=_inlet_indices(d, 0.55 * L, 0.7 * L, 0.3 * L, 0.6 * L)),
                     Q=lambda t: 2.0 + 1.5 * np.sin(t))
    A.Inlet_operator(d, A.Region(d, indices=_inlet_indices(d, 0.0, 0.15 * L, 0.0, 0.3 * L)), Q=-1.0)
    return d


def flather_de1(A, n=16):
    """open-ocean style boundary: Flather_external_stage_zero_velocity_boundary on two sides; the left
    one sits on a beach whose bed rises above the external stage, so both branches are taken"""
    d = A.rectangular_cross_domain(n, n, len1=float(n), len2=float(n))
    d.set_flow_algorithm("DE1")
    d.set_store(False)
    L = float(n)
    d.set_quantity("elevation", lambda x, y: -0.6 + 1.0 * (y / L) ** 2 + 0.05 * np.sin(x))
    d.set_quantity("stage", lambda x, y: 0.1 + 0.3 * np.exp(-((x - 0.5 * L) ** 2 + (y - 0.3 * L) ** 2) / 4.0),
                   location="centroids")
    d.set_quantity("friction", 0.02)
    Br = A.Reflective_boundary(d)
    Bf = A.Flather_external_stage_zero_velocity_boundary(d, lambda t: 0.1 + 0.05 * np.sin(t))
    d.set_boundary({"left": Bf, "right": Bf, "top": Br, "bottom": Bf})
    return d


def expression_de0(A, n=12):
    """quantities set from expressions over other quantities and over a polygon, as the reference's
    example scripts do (Quantity arithmetic, set_values(expression=...), set_values(polygon=...))"""
    d = A.rectangular_cross_domain(n, n, len1=float(n), len2=float(n))
    d.set_flow_algorithm("DE0")
    d.set_store(False)
    d.set_quantity("elevation", lambda x, y: -0.1 * x + 0.02 * x * y / n)
    d.set_quantity("stage", expression="elevation + 0.35")
    d.set_quantity("stage", 0.8, polygon=[[1.2, 1.1], [5.3, 1.4], [4.8, 6.2], [0.9, 5.1]])
    d.set_quantity("xmomentum", expression="0.1*(stage - elevation)")
    d.set_quantity("ymomentum", expression="(stage - elevation)**2 / (2.0 + elevation*elevation)")
    d.set_quantity("friction", 0.03)
    _reflective_all(A, d)
    return d


def _embankment(A, n=16):
    """two ponds separated by a dry embankment; only a culvert connects them"""
    d = A.rectangular_cross_domain(n, n, len1=float(n), len2=float(n))
    d.set_flow_algorithm("DE1")
    d.set_store(False)
    L = float(n)
    d.set_quantity("elevation", lambda x, y: 1.0 * np.exp(-((x - 0.5 * L) / 1.0) ** 2) + 0.02 * np.cos(y))
    d.set_quantity("stage", lambda x, y: np.where(x < 0.5 * L, 0.8, 0.25), location="centroids")
    d.set_quantity("friction", 0.03)
    _reflective_all(A, d)
    return d


def culvert_de1(A):
    """config[5] building block: Boyd_box_operator given end points (structures/boyd_box_operator.py,
    structure_operator.py), default momentum jet + velocity head"""
    d = _embankment(A)
    A.Boyd_box_operator(d, losses=1.5, width=1.5, height=0.6, end_points=[[6.1, 8.3], [9.9, 8.3]],
                        apron=0.55, enquiry_gap=0.4, manning=0.013, use_momentum_jet=True,
                        use_velocity_head=True, verbose=False)
    return d


def culvert_skew_de1(A):
    """skew culvert from exchange lines, explicit enquiry points and inverts, smoothed discharge,
    two partly blocked barrels, no jet, stage-driven; the head difference reverses the flow"""
    d = _embankment(A)
    d.set_quantity("stage", lambda x, y: np.where(x < 8.0, 0.3, 0.9), location="centroids")
    A.Boyd_box_operator(d, losses={"inlet": 0.5, "outlet": 1.0, "bend": 0.0}, width=1.2, height=0.5,
                        barrels=2.0, blockage=0.2,
                        exchange_lines=[[[6.15, 5.1], [6.4, 3.7]], [[9.8, 6.3], [10.05, 4.9]]],
                        enquiry_points=[[5.1, 4.2], [11.1, 5.8]], invert_elevations=[0.1, 0.05],
                        apron=0.45, manning=0.02, smoothing_timescale=1.5, use_momentum_jet=False,
                        use_velocity_head=False, verbose=False)
    return d


def culvert_pipe_de1(A):
    """Boyd_pipe_operator: two partly blocked circular barrels, smoothed discharge, skew exchange lines"""
    d = _embankment(A)
    A.Boyd_pipe_operator(d, losses=[0.5, 1.0], diameter=0.7, barrels=2.0, blockage=0.15,
                         end_points=[[6.1, 10.3], [9.9, 10.3]], apron=0.55, enquiry_gap=0.45,
                         manning=0.013, smoothing_timescale=0.5, use_momentum_jet=True,
                         use_velocity_head=True, verbose=False)
    return d


def culvert_weir_de1(A):
    """Weir_orifice_trapezoid_operator: trapezoidal opening with side slopes, no jet, smoothed"""
    d = _embankment(A)
    A.Weir_orifice_trapezoid_operator(d, losses=1.2, width=1.1, height=0.55, z1=0.6, z2=1.1, barrels=1.0,
                                      end_points=[[6.1, 5.3], [9.9, 5.3]], apron=0.55, enquiry_gap=0.4,
                                      manning=0.015, smoothing_timescale=0.3, use_momentum_jet=False,
                                      use_velocity_head=True, verbose=False)
    return d


def characteristic_de1(A, n=16):
    """Characteristic_stage_boundary: a solitary-wave style stage signal enters from the left over a
    sloping beach whose upper part is dry (the boundary's dry branch is exercised on the right)"""
    d = A.rectangular_cross_domain(n, n, len1=float(n), len2=float(n))
    d.set_flow_algorithm("DE1")
    d.set_store(False)
    L = float(n)
    d.set_quantity("elevation", lambda x, y: -1.0 + 1.4 * x / L + 0.02 * np.sin(y))
    d.set_quantity("stage", 0.0)
    d.set_quantity("friction", 0.02)
    Br = A.Reflective_boundary(d)
    Bc = A.Characteristic_stage_boundary(d, lambda t: 0.3 / np.cosh(0.8 * (t - 2.5)) ** 2, default_stage=0.0)
    d.set_boundary({"left": Bc, "right": Bc, "top": Br, "bottom": Br})
    return d


def wind_de1(A, n=14):
    """Wind_stress forcing term (shallow_water/forcing.py:80): a gust front that moves across a shallow
    lake, speed a function of (t, x, y), constant direction; plus a steady scalar breeze"""
    d = A.rectangular_cross_domain(n, n, len1=float(n), len2=float(n))
    d.set_flow_algorithm("DE1")
    d.set_store(False)
    L = float(n)
    d.set_quantity("elevation", lambda x, y: -0.6 + 0.3 * x / L)
    d.set_quantity("stage", 0.0)
    d.set_quantity("friction", 0.02)
    _reflective_all(A, d)

    def speed(t, x, y):
        return 30.0 * (x < 2.0 + 4.0 * t) + 0.0 * y
    d.forcing_terms.append(A.Wind_stress(speed, 20.0))
    d.forcing_terms.append(A.Wind_stress(s=8.0, phi=250.0))
    return d


def riverwall_de1(A, n=14):
    """riverwalls from breaklines (structures/riverwall.py:151): a levee along the cell boundaries x = n/2
    with a crest that dips below the upstream water level in the middle (weir flow over it), and a spur with
    its own hydraulic parameters; water behind the levee, nearly dry bed in front"""
    d = A.rectangular_cross_domain(n, n, len1=float(n), len2=float(n))
    d.set_flow_algorithm("DE1")
    d.set_store(False)
    L = float(n)
    d.set_quantity("elevation", lambda x, y: 0.1 * (x > L / 2) * (x - L / 2) / L)
    d.set_quantity("stage", lambda x, y: np.where(x < L / 2, 0.9, 0.05), location="centroids")
    d.set_quantity("friction", 0.03)
    _reflective_all(A, d)
    walls = {"levee": [[L / 2, 0.0, 1.1], [L / 2, 0.4 * L, 0.6], [L / 2, 0.6 * L, 0.7], [L / 2, L, 1.2]],
             "spur": [[L / 2, 0.5 * L, 0.5], [L / 2 + 3.0, 0.5 * L, 0.3]]}
    d.riverwallData.create_riverwalls(walls, riverwallPar={"spur": {"Qfactor": 0.8, "s1": 0.5}}, verbose=False)
    return d


def file_boundary_de1(A, n=8):
    """a small basin nested inside the field stored in file_boundary_source.sww (written by the reference,
    make_file_boundary_source.py): left File_boundary, right Field_boundary with a raised mean stage, top a
    Time_space_boundary function of (t, x, y), bottom Reflective"""
    import os
    src = os.path.join(os.path.dirname(os.path.abspath(__file__)), "file_boundary_source.sww")
    d = A.rectangular_cross_domain(n, n, len1=7.0, len2=6.0, origin=(3.3, 4.7))
    d.set_flow_algorithm("DE1")
    d.set_store(False)
    d.set_quantity("elevation", lambda x, y: -1.0 + 0.02 * x)
    d.set_quantity("stage", 0.0)
    d.set_quantity("friction", 0.01)
    Br = A.Reflective_boundary(d)
    Bf = A.File_boundary(src, d)
    Bm = A.Field_boundary(src, d, mean_stage=0.02)
    Bt = A.Time_space_boundary(d, function=lambda t, x, y: [0.01 * t + 0.001 * (x - 3.0), 0.0, 0.002 * y])
    d.set_boundary({"left": Bf, "right": Bm, "top": Bt, "bottom": Br})
    return d


def forcing_de1(A, n=14):
    """the older forcing-term API (shallow_water/forcing.py): Rainfall as a function of time over a polygon and
    as a constant over the whole domain, Inflow through a circle, all added to the stage update"""
    d = A.rectangular_cross_domain(n, n, len1=float(n), len2=float(n))
    d.set_flow_algorithm("DE1")
    d.set_store(False)
    L = float(n)
    d.set_quantity("elevation", lambda x, y: 0.3 * x / L + 0.05 * np.cos(y))
    d.set_quantity("stage", lambda x, y: np.maximum(0.3 * x / L + 0.05 * np.cos(y), 0.12), location="centroids")
    d.set_quantity("friction", 0.03)
    _reflective_all(A, d)
    d.forcing_terms.append(A.Rainfall(d, rate=lambda t: 40.0 + 10.0 * t,
                                      polygon=[[2.1, 2.2], [9.3, 2.4], [8.7, 9.1], [2.9, 8.3]]))
    d.forcing_terms.append(A.Rainfall(d, rate=3.0))
    d.forcing_terms.append(A.Inflow(d, rate=0.9, center=(10.2, 4.1), radius=1.7))
    return d


def discharge_de1(A, n=14):
    """Dirichlet_discharge_boundary: a fixed stage and a discharge per unit width entering through the left
    boundary of a gently sloping channel, leaving through a Dirichlet stage on the right"""
    d = A.rectangular_cross_domain(n, n, len1=float(n), len2=float(n))
    d.set_flow_algorithm("DE1")
    d.set_store(False)
    L = float(n)
    d.set_quantity("elevation", lambda x, y: -0.2 * x / L)
    d.set_quantity("stage", 0.3)
    d.set_quantity("friction", 0.03)
    Br = A.Reflective_boundary(d)
    # (the reference exports this class from anuga.shallow_water.boundaries only)
    DDB = getattr(A, "Dirichlet_discharge_boundary", None)
    if DDB is None:
        from anuga.shallow_water.boundaries import Dirichlet_discharge_boundary as DDB
    d.set_boundary({"left": DDB(d, 0.45, 0.35), "right": A.Dirichlet_boundary([0.1, 0.0, 0.0]),
                    "top": Br, "bottom": Br})
    return d


CASES = {
    "kat_bedslope_more_steps": (kat_bedslope_more_steps, dict(yieldstep=0.05, finaltime=0.5)),
    "dam_break_de0": (dam_break_de0, dict(yieldstep=1.0, finaltime=6.0)),
    "dam_break_de0_fixed_dt": (dam_break_de0_fixed_dt, dict(yieldstep=5.0, finaltime=10.0)),
    "dam_break_de1": (dam_break_de1, dict(yieldstep=1.0, finaltime=4.0)),
    "dam_break_de2": (dam_break_de2, dict(yieldstep=1.0, finaltime=3.0)),
    "beach_de1": (beach_de1, dict(yieldstep=1.0, finaltime=5.0)),
    "beach_de0_7": (beach_de0_7, dict(yieldstep=1.0, finaltime=3.0)),
    "tsunami_set_stage": (tsunami_set_stage, dict(yieldstep=1.0, finaltime=4.0)),
    "tsunami_dirichlet": (tsunami_dirichlet, dict(yieldstep=1.0, finaltime=4.0)),
    "time_boundary_de1": (time_boundary_de1, dict(yieldstep=1.0, finaltime=3.0)),
    "rain_de1": (rain_de1, dict(yieldstep=1.0, finaltime=4.0)),
    "rain_regions_de1": (rain_regions_de1, dict(yieldstep=1.0, finaltime=3.0)),
    "gate_de1": (gate_de1, dict(yieldstep=1.0, finaltime=3.0)),
    "rain_xyt_de1": (rain_xyt_de1, dict(yieldstep=1.0, finaltime=3.0)),
    "breach_de1": (breach_de1, dict(yieldstep=1.0, finaltime=3.0)),
    "drain_de1": (drain_de1, dict(yieldstep=1.0, finaltime=3.0)),
    "sloped_manning_de1": (sloped_manning_de1, dict(yieldstep=0.5, finaltime=2.0)),
    "low_froude_de1": (low_froude_de1, dict(yieldstep=0.5, finaltime=2.0)),
    "inlet_de1": (inlet_de1, dict(yieldstep=1.0, finaltime=3.0)),
    "rain_time_de1": (rain_time_de1, dict(yieldstep=1.0, finaltime=3.0)),
    "expression_de0": (expression_de0, dict(yieldstep=0.5, finaltime=1.5)),
    "flather_de1": (flather_de1, dict(yieldstep=1.0, finaltime=3.0)),
    "characteristic_de1": (characteristic_de1, dict(yieldstep=1.0, finaltime=4.0)),
    "wind_de1": (wind_de1, dict(yieldstep=1.0, finaltime=3.0)),
    "riverwall_de1": (riverwall_de1, dict(yieldstep=1.0, finaltime=4.0)),
    "forcing_de1": (forcing_de1, dict(yieldstep=1.0, finaltime=3.0)),
    "discharge_de1": (discharge_de1, dict(yieldstep=1.0, finaltime=3.0)),
    "file_boundary_de1": (file_boundary_de1, dict(yieldstep=1.0, finaltime=4.0)),
    "culvert_de1": (culvert_de1, dict(yieldstep=1.0, finaltime=4.0)),
    "culvert_pipe_de1": (culvert_pipe_de1, dict(yieldstep=1.0, finaltime=4.0)),
    "culvert_weir_de1": (culvert_weir_de1, dict(yieldstep=1.0, finaltime=4.0)),
    "culvert_skew_de1": (culvert_skew_de1, dict(yieldstep=1.0, finaltime=4.0)),
}

# 8-digit expected values embedded in the reference's own test (the KAT proper)
KAT_BEDSLOPE_W_EX_HEAD = [-0.0301883, -0.01127593, -0.02834861, -0.0108968, -0.02806583, -0.01074475]
