"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol
include/bdlm.h declares, and refuses to run without a GPU (no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from bayesian_dlms_b200 import _capi as capi
from bayesian_dlms_b200 import dlm
from bayesian_dlms_b200.batch import Model

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "bdlm.h")).read()
    return sorted(set(re.findall(r"BDLM_API[^;(]*?\b(bdlm_[a-z0-9_]+)\s*\(", hdr)))


def test_library_exports_every_declared_symbol():
    lib = capi.load()
    names = _declared_symbols()
    assert len(names) >= 16
    for name in names:
        assert hasattr(lib, name), f"{name} declared in include/bdlm.h but not exported"
    assert lib.bdlm_version() == 100


def test_binding_covers_every_declared_symbol():
    assert set(_declared_symbols()) == set(capi.SYMBOLS)


def test_struct_layout_matches_header(tmp_path):
    """sizeof / offsetof of every struct as gcc sees include/bdlm.h == the ctypes mirrors."""
    import subprocess
    structs = {"bdlm_problem": capi.Problem, "bdlm_kf_out": capi.KfOut,
               "bdlm_smooth_out": capi.SmoothOut, "bdlm_svd_out": capi.SvdOut,
               "bdlm_gibbs_stats": capi.GibbsStats, "bdlm_ar_problem": capi.ArProblem,
               "bdlm_ar_out": capi.ArOut, "bdlm_gibbs_prior": capi.GibbsPrior,
               "bdlm_gibbs_rng": capi.GibbsRng}
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "bdlm.h"', 'int main(void) {']
    for cname, cls in structs.items():
        lines.append(f'printf("{cname} %zu\\n", sizeof({cname}));')
        for fname, _ in cls._fields_:
            lines.append(f'printf("{cname}.{fname} %zu\\n", offsetof({cname}, {fname}));')
    lines += ['return 0; }']
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    inc = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include")
    subprocess.run(["gcc", "-I", inc, str(src), "-o", str(exe)], check=True)
    got = dict(l.split() for l in subprocess.run([str(exe)], capture_output=True, text=True,
                                                 check=True).stdout.splitlines())
    for cname, cls in structs.items():
        assert int(got[cname]) == C.sizeof(cls), cname
        for fname, _ in cls._fields_:
            assert int(got[f"{cname}.{fname}"]) == getattr(cls, fname).offset, (cname, fname)


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(capi.BdlmError) as ei:
        capi.Context(0)
    assert ei.value.code == capi.E_NODEVICE
    assert "no CPU fallback" in str(ei.value)


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "bayesian_dlms_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                src = open(os.path.join(dirpath, fn)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), fn
                assert "liboracle" not in src, fn


def test_model_materialisation_flags():
    m = Model.build(dlm.polynomial(2), T=10)
    assert (m.n, m.p, m.f_tv, m.g_tv, m.times) == (2, 1, False, False, None)
    assert np.array_equal(m.G, [1, 0, 1, 1])  # column-major [[1,1],[0,1]]
    seas = dlm.polynomial(1) + dlm.seasonal(24, 6)
    m = Model.build(seas, T=5)
    assert (m.n, m.p, m.g_tv) == (13, 1, False)  # regular grid: one rotation matrix
    m = Model.build(seas, times=[1.0, 2.0, 3.5, 7.0])
    assert m.g_tv and m.G.size == 4 * 169 and m.times is not None
    x = [np.array([0.5]), np.array([1.5]), np.array([2.5])]
    m = Model.build(dlm.regression(x), T=3)
    assert m.f_tv and m.F.size == 3 * 2
    with pytest.raises(ValueError):
        Model.build(dlm.polynomial(1), times=[])


def test_outer_sum_and_compose_shapes():
    a = dlm.polynomial(1) * dlm.polynomial(1)
    assert a.f(1.0).shape == (2, 2) and a.g(1.0).shape == (2, 2)
    b = dlm.polynomial(1) + dlm.seasonal(24, 3)
    assert b.f(1.0).shape == (7, 1) and b.g(1.0).shape == (7, 7)
    p = dlm.DlmParameters(3.0, 1.0, 0.0, 1.0) * dlm.DlmParameters(2.0, 1.0, 0.0, 1.0)
    assert p.v.shape == (2, 2) and p.m0.shape == (2,)


@pytest.mark.parametrize("n", [1, 2, 3, 4])
def test_scan_combine_is_associative_on_the_host(n):
    """bdlm_scan_combine is host code (the carry folding of the time-sharded scan): the
    filtering and smoothing operators of Sarkka & Garcia-Fernandez must be associative, with the
    documented identity elements -- that is what makes chunks on different GPUs composable."""
    lib = capi.load()
    rng = np.random.default_rng(n)

    def spd(scale):
        a = rng.standard_normal((n, n))
        return scale * (a @ a.T / n + 0.3 * np.eye(n))

    def felem():   # [A | b | C | eta | J], matrices column-major
        return np.concatenate([rng.standard_normal(n * n) * 0.5, rng.standard_normal(n),
                               spd(1.0).T.ravel(), rng.standard_normal(n), spd(0.2).T.ravel()])

    def selem():   # [E | g | L]
        return np.concatenate([rng.standard_normal(n * n) * 0.5, rng.standard_normal(n), spd(1.0).T.ravel()])

    def comb(backward, x, y):
        out = np.empty_like(x)
        assert lib.bdlm_scan_combine(n, backward, x.ctypes.data, y.ctypes.data, out.ctypes.data) == 0
        return out

    for backward, make in ((0, felem), (1, selem)):
        assert lib.bdlm_scan_elem_doubles(n, backward) == make().size
        a, b, c = make(), make(), make()
        left = comb(backward, comb(backward, a, b), c)
        right = comb(backward, a, comb(backward, b, c))
        assert np.allclose(left, right, rtol=1e-9, atol=1e-11), (backward, np.abs(left - right).max())
    # identities: forward (A = I, rest 0), backward (E = I, rest 0)
    ident = np.zeros(3 * n * n + 2 * n)
    ident[: n * n] = np.eye(n).ravel()
    x = felem()
    assert np.allclose(comb(0, ident, x), x, rtol=1e-12, atol=1e-14)
    assert np.allclose(comb(0, x, ident), x, rtol=1e-12, atol=1e-14)
    ident = np.zeros(2 * n * n + n)
    ident[: n * n] = np.eye(n).ravel()
    x = selem()
    assert np.allclose(comb(1, ident, x), x, rtol=1e-12, atol=1e-14)
    assert np.allclose(comb(1, x, ident), x, rtol=1e-12, atol=1e-14)


def test_communicator_needs_a_gpu_and_says_so():
    """bdlm_comm_create wraps bdlm_create per device: no GPU, no communicator, no fallback."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    lib = capi.load()
    h = C.c_void_p()
    devs = (C.c_int32 * 1)(0)
    rc = lib.bdlm_comm_create(devs, 1, 0, 1, None, C.byref(h))
    assert rc in (capi.E_NODEVICE, capi.E_NCCL) and not h.value
    assert lib.bdlm_comm_last_error(None)
    assert lib.bdlm_comm_create(devs, 1, 0, 2, None, C.byref(h)) == capi.E_ARG   # several processes need an id
    assert lib.bdlm_comm_size(None) == 0 and lib.bdlm_comm_ctx(None, 0) is None


def test_null_and_bad_arguments_are_codes_not_crashes():
    """No C++ exception and no segfault may cross the ABI (SURVEY.md 8b "Errors"): misuse that can
    be detected before any device work returns BDLM_E_ARG, also without a GPU."""
    lib = capi.load()
    h = C.c_void_p()
    devs = (C.c_int32 * 2)(0, 1)
    # communicator construction: device list / rank range / missing rendezvous id
    assert lib.bdlm_comm_create(None, 1, 0, 1, None, C.byref(h)) == capi.E_ARG and not h.value
    assert lib.bdlm_comm_create(devs, 0, 0, 1, None, C.byref(h)) == capi.E_ARG
    assert lib.bdlm_comm_create(devs, 2, 0, 1, None, C.byref(h)) == capi.E_ARG      # world < n_local
    assert lib.bdlm_comm_create(devs, 1, 3, 2, None, C.byref(h)) == capi.E_ARG      # rank outside world
    assert lib.bdlm_comm_create(devs, 1, 0, 2, None, C.byref(h)) == capi.E_ARG      # several processes, no id
    assert b"bdlm_comm_unique_id" in lib.bdlm_comm_last_error(None)
    assert lib.bdlm_comm_create(devs, 1, 0, 1, None, None) == capi.E_ARG
    assert lib.bdlm_comm_unique_id(None) == capi.E_ARG
    # calls on a null communicator / context
    assert lib.bdlm_comm_sync(None) == capi.E_ARG
    assert lib.bdlm_comm_allreduce_sum(None, None, 1) == capi.E_ARG
    assert lib.bdlm_comm_kf_filter(None, None, None, None) == capi.E_ARG
    assert lib.bdlm_comm_size(None) == 0 and lib.bdlm_comm_local_size(None) == 0
    assert not lib.bdlm_comm_ctx(None, 0)
    lib.bdlm_comm_destroy(None)
    assert lib.bdlm_kf_filter(None, None, None, None) == capi.E_ARG
    assert lib.bdlm_sync(None) == capi.E_ARG and lib.bdlm_set_rng(None, 1, 0, 0) == capi.E_ARG
    assert lib.bdlm_launch_count(None) == 0
    lib.bdlm_destroy(None)
    # host-side helpers of the time-sharded scan
    assert lib.bdlm_scan_elem_doubles(5, 0) == capi.E_ARG and lib.bdlm_scan_elem_doubles(0, 1) == capi.E_ARG
    x = np.zeros(64)
    assert lib.bdlm_scan_combine(5, 0, x.ctypes.data, x.ctypes.data, x.ctypes.data) == capi.E_ARG
    assert lib.bdlm_scan_combine(2, 0, None, x.ctypes.data, x.ctypes.data) == capi.E_ARG
