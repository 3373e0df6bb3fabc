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
    assert L.b200dae_version() == 102
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
    int main(void){printf("%zu %zu %zu %zu\n", sizeof(b200dae_radau_tables), sizeof(b200dae_radau_opts),
                          sizeof(b200dae_bdf_opts), sizeof(b200dae_outputs)); return 0;}
    '''
    with tempfile.TemporaryDirectory() as td:
        c = os.path.join(td, "probe.c")
        open(c, "w").write(probe)
        exe = os.path.join(td, "probe")
        subprocess.run(["gcc", "-I", os.path.join(REPO, "include"), c, "-o", exe], check=True)
        sizes = list(map(int, subprocess.run([exe], check=True, capture_output=True, text=True).stdout.split()))
    assert sizes == [ctypes.sizeof(CRadauTables), ctypes.sizeof(_lib.RadauOpts),
                     ctypes.sizeof(_lib.BdfOpts), ctypes.sizeof(_lib.Outputs)]


def test_no_gpu_fails_loudly():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from scipy_dae_b200 import solve_dae_batched, _lib
    with pytest.raises(_lib.B200DaeError):
        solve_dae_batched("robertson", (0, 1.0), rtol=1e-6, atol=1e-6)
