"""The C-ABI library loads and exports every symbol include/b200dae.h declares (no GPU)."""
import ctypes
import os
import re

import pytest

from conftest import REPO


def declared_symbols():
    src = open(os.path.join(REPO, "include", "b200dae.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(b200dae_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_header_symbols():
    from scipy_dae_b200 import _lib
    from scipy_dae_b200.build import build
    build()
    L = ctypes.CDLL(_lib.LIB_PATH)
    syms = declared_symbols()
    assert len(syms) >= 9
    for s in syms:
        assert hasattr(L, s), f"{s} declared in include/b200dae.h but not exported"
    assert set(syms) == set(_lib.EXPORTS)


def test_registry_without_gpu():
    from scipy_dae_b200 import _lib
    from scipy_dae_b200.problems import PROBLEMS
    L = _lib.load()
    assert L.b200dae_version() == 106
    assert L.b200dae_num_problems() == len(PROBLEMS)
    for name, pr in PROBLEMS.items():
        pid, n, npar = _lib.problem_lookup(name)
        assert (n, npar) == (pr.n, pr.n_params)
        assert L.b200dae_problem_name(pid).decode() == name
    with pytest.raises(ValueError):
        _lib.problem_lookup("no_such_problem")


def test_struct_layouts_match_header():
    """ctypes mirrors have the sizes the C structs have (checked against a compiled probe)."""
    import subprocess, tempfile
    from scipy_dae_b200 import _lib
    from scipy_dae_b200.constants import CRadauTables
    probe = r'''
    #include <stdio.h>
    #include "b200dae.h"
    int main(void){printf("%zu %zu %zu %zu %zu %zu %zu\n", sizeof(b200dae_radau_tables), sizeof(b200dae_radau_opts),
                          sizeof(b200dae_bdf_opts), sizeof(b200dae_outputs), sizeof(b200dae_event_spec),
                          sizeof(b200dae_events), sizeof(b200dae_cic_opts)); return 0;}
    '''
    with tempfile.TemporaryDirectory() as td:
        c = os.path.join(td, "probe.c")
        open(c, "w").write(probe)
        exe = os.path.join(td, "probe")
        subprocess.run(["gcc", "-I", os.path.join(REPO, "include"), c, "-o", exe], check=True)
        sizes = list(map(int, subprocess.run([exe], check=True, capture_output=True, text=True).stdout.split()))
    assert sizes == [ctypes.sizeof(CRadauTables), ctypes.sizeof(_lib.RadauOpts),
                     ctypes.sizeof(_lib.BdfOpts), ctypes.sizeof(_lib.Outputs), ctypes.sizeof(_lib.EventSpec),
                     ctypes.sizeof(_lib.Events), ctypes.sizeof(_lib.CicOpts)]


def test_no_gpu_fails_loudly():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from scipy_dae_b200 import solve_dae_batched, _lib
    with pytest.raises(_lib.B200DaeError):
        solve_dae_batched("robertson", (0, 1.0), rtol=1e-6, atol=1e-6)


def test_no_gpu_fails_loudly_everywhere():
    import numpy as np
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from scipy_dae_b200 import (solve_dae_batched_host, consistent_initial_conditions_batched, _lib)
    y0, yp0, p = np.array([[1.0, 0, 0]]), np.array([[-0.04, 0.04, 0]]), np.array([[0.04, 1e4, 3e7]])
    with pytest.raises(_lib.B200DaeError):
        solve_dae_batched_host("robertson", (0, 1.0), y0, yp0, p, rtol=1e-6, atol=1e-6)
    with pytest.raises(_lib.B200DaeError):
        consistent_initial_conditions_batched("robertson", 0.0, y0, yp0, p)


def test_event_spec_validation_and_layout():
    """ThresholdEvent -> b200dae_event_spec (host side only): the reference's `terminal` /
    `direction` attributes (dae.py:226-247; scipy prepare_events) and argument errors."""
    from scipy_dae_b200 import ThresholdEvent
    from scipy_dae_b200.batched import _events_struct
    ev = _events_struct([ThresholdEvent(0, 0.5), dict(component=2, value=-1.0, of="yp", direction=-3.0, terminal=True),
                         ThresholdEvent(1, terminal=4, direction=2)], n=3)
    assert ev.spec.n_events == 3
    assert list(ev.spec.of_yp)[:3] == [0, 1, 0] and list(ev.spec.component)[:3] == [0, 2, 1]
    assert list(ev.spec.direction)[:3] == [0, -1, 1] and list(ev.spec.max_events)[:3] == [0, 1, 4]
    with pytest.raises(ValueError):
        _events_struct([ThresholdEvent(5)], n=3)
    with pytest.raises(ValueError):
        _events_struct([ThresholdEvent(0, of="z")], n=3)
    with pytest.raises(TypeError):
        _events_struct([lambda t, y, yp: y[0]], n=3)
    with pytest.raises(ValueError):
        _events_struct([ThresholdEvent(0)] * 5, n=3)


def test_oracle_is_not_imported_by_the_product():
    """The oracle is test infrastructure: nothing under scipy_dae_b200/ may import or load it."""
    import glob
    for f in glob.glob(os.path.join(REPO, "scipy_dae_b200", "**", "*.py"), recursive=True):
        for ln in open(f).read().splitlines():
            code = ln.split("#")[0]
            if re.search(r"\b(import|from|CDLL|LoadLibrary)\b", code):
                assert "oracle" not in code.lower(), (f, ln)
