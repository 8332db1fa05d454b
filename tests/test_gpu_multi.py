"""GPU (>= 2 devices): the multi-GPU mode - reference partition / ghost scheme, NCCL halo
exchange and NCCL min-allreduce of dt inside the device time loop - reproduces the single-GPU
run: same timestep sequence, and every full triangle BIT-identical (ghosts are exact copies
and the min is exact, so nothing may differ)."""
import os
import subprocess
import sys

import numpy as np
import pytest

import anuga_core_b200 as ab
from golden_util import cases

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run_distributed(case, nranks, tmp_path, rule, extra=()):
    port = 29600 + (os.getpid() % 300)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(nranks),
           "--master-addr", "127.0.0.1", "--master-port", str(port),
           os.path.join(ROOT, "tests", "multi_gpu_worker.py"), case, str(tmp_path), rule] + list(extra)
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=240)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    return [np.load(os.path.join(str(tmp_path), "rank%d.npz" % r)) for r in range(nranks)]


@pytest.mark.parametrize("case,rule", [("beach_de1", "blocks"), ("dam_break_de0", "quadrants"),
                                       ("dam_break_de2", "blocks"), ("rain_de1", "quadrants"),
                                       ("time_boundary_de1", "blocks"), ("inlet_de1", "quadrants"),
                                       ("culvert_de1", "blocks"), ("culvert_skew_de1", "quadrants")])
def test_multi_gpu_equals_single_gpu_bitwise(case, rule, tmp_path):
    ndev = ab.device_count()
    if ndev < 2:
        pytest.skip("needs >= 2 GPUs")
    nranks = 4 if (ndev >= 4 and rule == "quadrants") else 2
    parts = run_distributed(case, nranks, tmp_path, rule)
    builder, ev = cases.CASES[case]
    d = builder(ab)
    times = [t for t in d.evolve(**ev)]
    q = d.quantities
    seen = np.zeros(d.number_of_triangles, dtype=bool)
    for p in parts:
        ids = p["ids"]
        assert not seen[ids].any()
        seen[ids] = True
        assert np.array_equal(p["times"], np.array(times))
        assert int(p["steps"][0]) == d.total_steps
        assert float(p["dt"][0]) == d.timestep
        assert np.array_equal(p["stage"], q["stage"].centroid_values[ids])
        assert np.array_equal(p["xmom"], q["xmomentum"].centroid_values[ids])
        assert np.array_equal(p["ymom"], q["ymomentum"].centroid_values[ids])
    assert seen.all()


def test_multi_gpu_sww_files_merge_into_the_single_gpu_file(tmp_path):
    """set_store(True) on a distributed run: per-rank SWW files, sww_merge, and the merged file holds
    the single-GPU file's centroid frames (in the partition's triangle order)"""
    if ab.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    from scipy.io import netcdf_file
    case = "beach_de1"
    parts = run_distributed(case, 2, tmp_path, "quadrants", extra=("store",))
    builder, ev = cases.CASES[case]
    d = builder(ab)
    d.set_store(True)
    d.set_name("single_" + case)
    d.set_datadir(str(tmp_path))
    for t in d.evolve(**ev):
        pass
    assert not os.path.exists(os.path.join(str(tmp_path), "multi_%s_P2_0.sww" % case))
    m = netcdf_file(os.path.join(str(tmp_path), "multi_%s.sww" % case), "r", mmap=False)
    s = netcdf_file(os.path.join(str(tmp_path), "single_%s.sww" % case), "r", mmap=False)
    order = np.empty(d.number_of_triangles, dtype=np.int64)
    for p in parts:
        order[p["gids"]] = p["ids"]
    assert np.array_equal(m.variables["time"][:], s.variables["time"][:])
    assert np.array_equal(m.variables["x"][:], s.variables["x"][:])
    assert np.array_equal(m.variables["volumes"][:], s.variables["volumes"][:][order])
    for name in ("stage_c", "xmomentum_c", "ymomentum_c"):
        assert np.array_equal(m.variables[name][:], s.variables[name][:][:, order]), name
    assert np.array_equal(m.variables["elevation_c"][:], s.variables["elevation_c"][:][order])
    a, b = m.variables["stage"][:], s.variables["stage"][:]
    assert a.shape == b.shape and np.mean(a == b) > 0.9
