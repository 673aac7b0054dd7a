"""The C-ABI shared library loads on a CPU-only box, exports every symbol include/sfb200.h declares, and its numeric
entry points fail loudly (no CPU fallback) when there is no GPU."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import sparseframe_b200 as sf
from sparseframe_b200 import _capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "sfb200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(sfb200_\w+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = C.CDLL(sf.LIB_PATH)
    names = declared_symbols()
    assert len(names) >= 20
    for name in names:
        assert hasattr(lib, name), name
    assert sorted(sf.SYMBOLS) == names
    assert b"sfb200" in _capi.lib().sfb200_version()


def test_analyze_rejects_bad_arguments():
    L = _capi.lib()
    h = C.c_void_p()
    assert L.sfb200_analyze(7, 1, 3, None, None, None, 0, C.byref(h)) == 1
    assert "bad argument" in _capi.last_error()


def test_wrap_structure_roundtrip():
    A = sf.compress(*sf.matrices.generate("lap2d", 12))
    S = sf.analyze(A, "cholesky")
    W = sf.api.wrap_structure("cholesky", A.n, S.array("Super"), S.array("Lsip"), S.array("Lsxp"), S.array("Lsi"),
                              S.array("Lp"), S.array("Li"))
    for k in ("Super", "SuperMap", "Lsip", "Lsxp", "Lsi", "Lp", "Li"):
        assert np.array_equal(W.array(k), S.array(k)), k
    assert W.flops == S.flops


@pytest.mark.parametrize("kind,gen,g", [("cholesky", "lap2d", 12), ("lu", "cd3d", 6)])
def test_wrap_structure_rejects_malformed_structures(kind, gen, g):
    """A structure from somebody else's analysis is checked before the plan builder indexes with it: every corruption is
    SFB200_ERR_ARG with a message, never an out-of-bounds access or an abort inside the library."""
    A = sf.compress(*sf.matrices.generate(gen, g))
    S = sf.analyze(A, kind)
    good = {k: S.array(k) for k in ("Super", "Lsip", "Lsxp", "Lsi", "Lp", "Li")}
    lu = kind == "lu"
    if lu:
        good["Up"], good["Ui"] = S.array("Up"), S.array("Ui")

    def wrap(d):
        return sf.api.wrap_structure(kind, A.n, d["Super"], d["Lsip"], d["Lsxp"], d["Lsi"], d["Lp"], d["Li"],
                                     d.get("Up"), d.get("Ui"))
    wrap(good).close()
    ns = len(good["Super"]) - 1
    big = int(np.argmax(np.diff(good["Lsip"]) - np.diff(good["Super"])))      # a supernode with off-diagonal rows

    def corrupt(name, fn):
        d = {k: v.copy() for k, v in good.items()}
        fn(d[name])
        return d
    cases = [
        corrupt("Super", lambda a: a.__setitem__(ns, A.n + 1)),
        corrupt("Super", lambda a: a.__setitem__(1, a[0])),
        corrupt("Lsip", lambda a: a.__setitem__(1, a[1] + 1)),
        corrupt("Lsxp", lambda a: a.__setitem__(ns, a[ns] + 2)),
        corrupt("Lsi", lambda a: a.__setitem__(good["Lsip"][big], a[good["Lsip"][big]] + 1)),                    # not its own column
        corrupt("Lsi", lambda a: a.__setitem__(good["Lsip"][big + 1] - 1, A.n)),                                 # row outside 0..n
        corrupt("Lsi", lambda a: a.__setitem__(good["Lsip"][big + 1] - 1, a[good["Lsip"][big]])),                # not ascending
        corrupt("Lp", lambda a: a.__setitem__(3, a[2] - 1)),
        corrupt("Li", lambda a: a.__setitem__(0, -1)),
    ]
    if lu:
        cases.append(corrupt("Ui", lambda a: a.__setitem__(0, A.n)))
    for d in cases:
        with pytest.raises(RuntimeError, match="invalid supernodal structure"):
            wrap(d)


@pytest.mark.skipif(sf.device_count() > 0, reason="this check is for the CPU-only box")
def test_numeric_path_fails_loudly_without_gpu():
    A = sf.compress(*sf.matrices.generate("lap2d", 8))
    S = sf.analyze(A, "cholesky")
    with pytest.raises(RuntimeError, match="no usable CUDA device"):
        sf.Plan(S)


def test_missing_library_is_an_error(monkeypatch):
    monkeypatch.setattr(_capi, "_lib", None)
    monkeypatch.setattr(_capi, "LIB_PATH", "/nonexistent/libsfb200.so")
    with pytest.raises(sf.LibraryMissing):
        _capi.lib()
