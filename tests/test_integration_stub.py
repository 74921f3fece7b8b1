"""The Option-B ctypes stub of INTEGRATION.md (what a pywfo maintainer would paste into pywfo/main.py),
extracted from the document and run VERBATIM against goldens the unmodified reference produced."""
import os
import re

import numpy as np
import pytest

from conftest import load_golden

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _stub_source():
    md = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    return re.findall(r"```python\n(import ctypes as C, numpy as np.*?)```", md, flags=re.S)[0]


def test_stub_is_in_the_document():
    src = _stub_source()
    assert "wfo_overlaps_ci_host" in src and "def overlaps_cache" in src
    compile(src, "INTEGRATION.md", "exec")


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["r_n17", "r_water", "r_odd"])
def test_stub_runs_verbatim_against_reference_goldens(name):
    src = _stub_source().replace('C.CDLL("libwfo_b200.so")', f'C.CDLL("{ROOT}/pywfo_b200/libwfo_b200.so")')
    ns = {}
    exec(compile(src, "INTEGRATION.md", "exec"), ns)
    g = load_golden(name)
    for ao, key in (("ket", "ref_overlaps_cache"), ("bra", "ref_overlaps_cache_bra")):
        got = ns["overlaps_cache"](g["bra_mos"], g["ket_mos"], g["bra_ci"], g["ket_ci"], int(g["occ"]),
                                   ci_thresh=float(g["ci_thresh"]), ao_ovlps=ao)
        want = g[key]
        np.testing.assert_allclose(got, want, rtol=1e-10, atol=1e-10 * np.abs(want).max())
