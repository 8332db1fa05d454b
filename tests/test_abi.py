"""CPU: the C-ABI library builds, loads and exports every symbol include/swk.h declares;
without a device every compute entry point fails loudly (no CPU fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest

import anuga_core_b200 as ab
from anuga_core_b200 import backend, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    build.build()
    return backend.load_library()


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "swk.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(swk_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported_and_bound(lib):
    names = declared_symbols()
    assert len(names) >= 40
    for n in names:
        assert hasattr(lib, n), "libswk.so does not export %s" % n
        assert n in backend.SYMBOLS, "backend.py does not bind %s" % n
    assert sorted(backend.SYMBOLS) == names
    assert lib.swk_abi_version() == 2


def test_library_contains_sm100a_code_only():
    out = os.popen("cuobjdump -lelf %s 2>/dev/null" % backend.LIB_PATH).read()
    if not out:
        pytest.skip("cuobjdump not available")
    assert "sm_100a" in out
    assert not re.search(r"sm_(?!100a)\d+", out)


def test_struct_layouts_match_header_field_order():
    text = open(os.path.join(ROOT, "include", "swk.h")).read()
    body = text[text.index("typedef struct {", text.index("scalar parameters")):text.index("} swk_params;")]
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    fields = []
    for decl in body.split(";"):
        decl = decl.replace("typedef struct {", "").strip()
        if not decl:
            continue
        names = decl.split(None, 1)[1]
        fields += [x.strip().lstrip("*") for x in names.split(",")]
    assert fields == [f[0] for f in backend.SwkParams._fields_]


@pytest.mark.skipif(ab.device_count() > 0, reason="this test documents the no-GPU behaviour")
def test_no_cpu_fallback_without_device(lib):
    d = ab.rectangular_cross_domain(2, 2)
    d.set_quantity("stage", 1.0)
    B = ab.Reflective_boundary(d)
    d.set_boundary({t: B for t in d.get_boundary_tags()})
    with pytest.raises(ab.SwkError) as e:
        for _ in d.evolve(yieldstep=0.1, finaltime=0.1):
            pass
    assert e.value.code == -1
    with pytest.raises(Exception):
        d.set_multiprocessor_mode(2)          # CPU modes are the reference's, not ours
    w = np.ones(4)
    rc = lib.swk_call_update(0, 4, 0.1, w.ctypes.data_as(backend._PD), w.ctypes.data_as(backend._PD),
                             w.ctypes.data_as(backend._PD))
    assert rc == -1 and b"no CUDA device" in lib.swk_last_error()


def test_product_never_imports_the_oracle():
    """only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may touch oracle/"""
    pkg = os.path.join(ROOT, "anuga_core_b200")
    bad = re.compile(r"^\s*(from|import)\s+oracle\b|liboracle|libanuga_ref|oracle/|oracle\.driver", re.M)
    for dirpath, _dirs, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not bad.search(src), "%s reaches into oracle/" % f


def test_enumerations_of_the_header_match_the_python_binding():
    """boundary kinds and quantity ids are plain integers across the C ABI: every SWK_BC_* / SWK_Q_* of
    include/swk.h must have the same value in anuga_core_b200/backend.py, and every boundary class must name a
    kind the header knows"""
    import re
    from anuga_core_b200 import backend, boundaries, file_boundary
    text = open(os.path.join(ROOT, "include", "swk.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    enums = dict((k, int(v)) for k, v in re.findall(r"\b(SWK_(?:BC|Q)_[A-Z0-9_]+)\s*=\s*(\d+)", text))
    bc = {k[len("SWK_"):]: v for k, v in enums.items() if k.startswith("SWK_BC_")}
    assert len(bc) == 12 and sorted(bc.values()) == list(range(12))
    for name, value in bc.items():
        assert getattr(backend, name) == value, name
    for name, value in enums.items():
        if name.startswith("SWK_Q_"):
            assert backend.Q[name[len("SWK_Q_"):]] == value, name
    kinds = set()
    for mod in (boundaries, file_boundary):
        for obj in vars(mod).values():
            if isinstance(obj, type) and issubclass(obj, boundaries.Boundary):
                kinds.add(obj.device_kind)
    assert kinds <= set(bc.values()) and kinds >= {1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11}
