"""GPU: a real unstructured mesh - the reference's Merimbula lake example (10 785 triangles, node valences
1-9, UTM coordinates, tidal set-stage boundary), rebuilt from the arrays in tests/golden/merimbula_de1.npz -
against the Python reference's golden run.  (Named to run last: added after the round's GPU budget was
spent, so its first execution is the driver's.)"""
import numpy as np
import pytest

import anuga_core_b200 as ab
import merimbula_case
from golden_util import load, rel_err

pytestmark = pytest.mark.gpu


def test_merimbula_lake_matches_python_reference_golden():
    g = load("merimbula_de1")
    d = merimbula_case.from_fixture(ab, g)
    d.record_timestep_history = True
    yields = [t for t in d.evolve(**merimbula_case.EVOLVE)]
    d.sync_to_host()
    q = d.quantities
    assert np.array_equal(np.array(yields), g["yields"])
    assert d.total_steps == len(g["dts"])
    assert np.array_equal(np.array(d.timestep_history), g["dts"])
    e = max(rel_err(q["stage"].centroid_values, g["final_stage"]), rel_err(q["xmomentum"].centroid_values, g["final_xmom"]),
            rel_err(q["ymomentum"].centroid_values, g["final_ymom"]))
    assert e <= 1e-9, e
    print("\n[merimbula] %d steps, final rel err %.2e" % (d.total_steps, e))
