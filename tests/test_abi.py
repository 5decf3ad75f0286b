"""CPU suite, part 3: the C-ABI shared library loads and exports every symbol include/mdqt.h declares; host-side
helpers (parameter derivation, Philox host replica, planner) behave; and WITHOUT a GPU every compute entry point
fails loudly (there is no CPU fallback). No kernel is launched here."""
import ctypes
import os
import re

import numpy as np
import pytest

import mdqtplasmasims_b200 as pkg
from oracle import pyoracle as po

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from mdqtplasmasims_b200 import build
    build.build()
    return pkg.load_library()


def test_header_symbols_are_all_exported(lib):
    hdr = open(os.path.join(ROOT, "include", "mdqt.h")).read()
    declared = sorted(set(re.findall(r"\b(mdqt_[a-z0-9_]+)\s*\(", hdr)))
    assert declared, "no declarations parsed"
    for name in declared:
        assert hasattr(lib, name), "libmdqt_b200.so does not export %s" % name
    assert sorted(declared) == sorted(pkg.ABI_SYMBOLS)


def test_no_torch_types_in_the_abi():
    hdr = open(os.path.join(ROOT, "include", "mdqt.h")).read()
    assert "torch" not in hdr.lower() and "at::" not in hdr and "std::" not in hdr


def test_params_struct_layout_and_su_constants(lib):
    p = pkg.su_params()
    assert p.struct_bytes == ctypes.sizeof(pkg.Params)
    assert p.scheme == 12 and p.n_ions == 3500 and p.substeps_per_md == 25
    assert p.L == 24.474784873331927 and abs(1 / p.kappa - 1.8257418583505536) < 1e-15 and p.rcut == p.L / 2
    q, ratio = po.su_params()
    for k in ("dtq", "g2E", "pv2qv", "vKick", "vKickDP", "dR", "kRat", "detuning", "detuningDP", "Om", "OmDP"):
        assert getattr(p, k) == getattr(q, k), k
    p2 = pkg.su_params(density=0.7, fracOfSig=0.3)
    q2, r2 = po.su_params(density=0.7, fracOfSig=0.3)
    assert p2.substeps_per_md == r2 and p2.dtq == q2.dtq and p2.g2E == q2.g2E


def test_params_md_constants(lib):
    p = pkg.md_params(scheme=pkg.SCHEME_SR7, n_ions=4096, kappa=0.5, density=2.0)
    q, ratio = po.mc408_params(n=2.0)
    assert p.L == (4096 * 4. * np.pi / 3.) ** (1. / 3) and p.substeps_per_md == ratio == 62
    assert p.dtq == q.dtq and p.g2E == q.g2E and p.pv2qv == q.pv2qv


def test_philox_host_replica_equals_oracle(lib, oracle):
    for seed, traj, sub in ((0, 0, 0), (12345, 3, 7), (2 ** 40 + 17, 1000, 2 ** 33 + 5)):
        assert np.array_equal(pkg.philox_uniforms(seed, traj, 16, sub), oracle.uniforms5(seed, traj, 16, sub))


def test_version_and_error_strings(lib):
    assert b"sm_100a" in lib.mdqt_version()
    assert isinstance(lib.mdqt_last_error(), bytes)


def test_invalid_params_rejected_before_touching_a_device(lib):
    p = pkg.su_params()
    p.struct_bytes = 8
    h = ctypes.c_void_p()
    assert lib.mdqt_create(ctypes.byref(p), ctypes.byref(h)) == -1 and b"ABI" in lib.mdqt_last_error()
    p = pkg.su_params(n_ions=0)
    assert lib.mdqt_create(ctypes.byref(p), ctypes.byref(h)) == -1
    p = pkg.su_params()
    p.rcut = p.L
    assert lib.mdqt_create(ctypes.byref(p), ctypes.byref(h)) == -1 and b"L/2" in lib.mdqt_last_error()
    assert lib.mdqt_forces(None) == -1 and lib.mdqt_substeps(None, 1) == -1 and lib.mdqt_md_steps(None, 1) == -1


def test_no_cpu_fallback_without_gpu(lib):
    if lib.mdqt_device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(pkg.MDQTError) as ei:
        pkg.Engine(pkg.su_params(n_ions=64, N0=64))
    assert "no CUDA device" in str(ei.value) or "-2" in str(ei.value)


def test_product_never_imports_the_oracle():
    """The product path must not route through oracle/ (test infrastructure): grep the package sources."""
    pkgdir = os.path.join(ROOT, "mdqtplasmasims_b200")
    for dirpath, _, files in os.walk(pkgdir):
        for f in files:
            if f.endswith((".py", ".cu", ".h", ".cuh", ".cpp")):
                src = open(os.path.join(dirpath, f)).read()
                assert "pyoracle" not in src and "liboracle" not in src and "oracle/" not in src.replace("oracle/mdqt_oracle.c (orc_", ""), f


def _build_c_client(tmp_path):
    """examples/abi_client.c: a plain C99 program against include/mdqt.h, linked to the in-tree library."""
    import subprocess
    exe = os.path.join(str(tmp_path), "abi_client")
    pkg_dir = os.path.join(ROOT, "mdqtplasmasims_b200")
    cmd = ["gcc", "-std=c99", "-O2", "-Wall", "-Werror", "-I" + os.path.join(ROOT, "include"), os.path.join(ROOT, "examples", "abi_client.c"),
           "-L" + pkg_dir, "-lmdqt_b200", "-Wl,-rpath," + pkg_dir, "-lm", "-o", exe]
    res = subprocess.run(cmd, capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    return exe


def test_c_client_compiles_links_and_fails_loudly_without_gpu(lib, tmp_path):
    """The ABI is usable from plain C (no C++ or torch types in the header); without a device the client gets
    MDQT_ENODEVICE -- there is no CPU fallback to fall into."""
    import subprocess
    exe = _build_c_client(tmp_path)
    if lib.mdqt_device_count() > 0:
        pytest.skip("a GPU is present: the run itself is covered by the gpu-marked test")
    res = subprocess.run([exe, "64", "1"], capture_output=True, text=True)
    assert res.returncode == 3 and "no CPU fallback" in res.stderr


@pytest.mark.gpu
def test_c_client_matches_the_python_mirror(tmp_path):
    """The same run through the C program and through the ctypes mirror: identical observables."""
    import subprocess
    import numpy as np
    from mdqtplasmasims_b200 import Engine, su_params
    exe = _build_c_client(tmp_path)
    n, nsteps = 500, 3
    res = subprocess.run([exe, str(n), str(nsteps)], capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    got = {kv.split("=")[0]: float(kv.split("=")[1]) for kv in res.stdout.split()}
    # replay the client's xorshift64 initial state in Python
    state = 88172645463325252
    M = (1 << 64) - 1

    def urand():
        nonlocal state
        state ^= (state << 13) & M; state ^= state >> 7; state ^= (state << 17) & M
        return (state >> 11) * (1.0 / 9007199254740992.0)
    p = su_params(n_ions=n, N0=n, seed=2024, traj0=1)
    R = np.zeros((3, n)); psi = np.zeros((n, 12, 2))
    for i in range(n):
        for c in range(3):
            R[c, i] = p.L * urand()
        r1, r2 = urand(), urand()
        psi[i, 0, 0] = np.sqrt(r1); psi[i, 1, 0] = np.sqrt(1 - r1) * np.sqrt(r2); psi[i, 1, 1] = np.sqrt(1 - r1) * np.sqrt(1 - r2)
    eng = Engine(p)
    eng.upload(R=R, V=np.zeros((3, n)), psi=psi, tPart=np.zeros(n), t=0.0, substep=0)
    eng.md_steps(nsteps)
    d = eng.diagnostics()
    s = eng.download(("psi",))
    for k in ("t", "ekin_x", "ekin_y", "ekin_z", "epot", "vx_avg"):
        assert got[k] == d[k], k
    assert abs(got["norm"] - (s["psi"] ** 2).sum() / n) < 1e-14
