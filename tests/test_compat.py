"""Scripts written for the reference run against this package after install_as_anuga() (in a
subprocess: the alias must not shadow the real reference used by other tests)."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SCRIPT = r'''
import sys
sys.path.insert(0, %(root)r)
import anuga_core_b200
anuga_core_b200.install_as_anuga()

# ---- from here on: lines as they appear in the reference's example scripts ----
import anuga
from anuga import Domain, Reflective_boundary, Dirichlet_boundary, Time_boundary, rectangular_cross, g
from anuga import distribute, myid, numprocs, finalize, barrier
from anuga.abstract_2d_finite_volumes.mesh_factory import rectangular_cross as rc2
from anuga.abstract_2d_finite_volumes.quantity import Quantity
from anuga.structures.boyd_box_operator import Boyd_box_operator
from anuga.structures.inlet_operator import Inlet_operator
from anuga.operators.rate_operators import Rate_operator
from anuga.geometry.polygon_function import Polygon_function
from anuga.parallel import myid as myid2

assert g == 9.8 and myid == 0 and numprocs == 1 and rc2 is rectangular_cross
points, vertices, boundary = rectangular_cross(20, 10, len1=20.0, len2=10.0)
domain = Domain(points, vertices, boundary)
domain.set_name('compat')
domain.set_flow_algorithm('DE1')
domain.set_quantity('elevation', Polygon_function([([[5, 2], [9, 2], [9, 8], [5, 8]], 0.5)], default=lambda x, y: -x / 40))
domain.set_quantity('friction', 0.03)
domain.set_quantity('stage', expression='elevation + 0.2')
domain.add_quantity('stage', 0.05)
Br = Reflective_boundary(domain)
Bd = Dirichlet_boundary([0.3, 0.0, 0.0])
domain.set_boundary({'left': Bd, 'right': Br, 'top': Br, 'bottom': Br})
Rate_operator(domain, rate=lambda t: 0.001, polygon=[[1, 1], [4, 1], [4, 4], [1, 4]])
Inlet_operator(domain, anuga.Region(domain, center=[15.0, 5.0], radius=1.5), Q=0.5)
Boyd_box_operator(domain, losses=1.5, width=1.0, height=0.5, end_points=[[4.2, 5.1], [9.8, 5.1]],
                  apron=0.6, enquiry_gap=0.3)
domain = distribute(domain)                      # one process: returns the domain itself
z = domain.quantities['elevation'].centroid_values
w = domain.quantities['stage'].centroid_values
assert abs((w - z) - 0.25).max() < 1e-12 and z.max() == 0.5
assert len(domain.fractional_step_operators) == 3
print(domain.statistics().splitlines()[2])
print('compat-ok')
'''


def test_reference_style_script_runs_up_to_evolve(tmp_path):
    script = tmp_path / "script.py"
    script.write_text(SCRIPT % {"root": ROOT})
    res = subprocess.run([sys.executable, str(script)], capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
    assert "compat-ok" in res.stdout and "Number of triangles = 800" in res.stdout
