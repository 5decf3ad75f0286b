"""ctypes front-end for the host-side (no GPU) driver pieces of include/mdqt_io.h: the reference's directory naming,
init(), restart files, output() files and main-loop schedule (SU:289-348, 725-1032, 1147-1159, 1248-1378)."""
import ctypes

import numpy as np

from .engine import MDQTError, load_library

c_double_p = ctypes.POINTER(ctypes.c_double)
N_VINT = 13

IO_SYMBOLS = ["mdqt_io_dirname", "mdqt_io_init_su", "mdqt_io_write_conditions", "mdqt_io_read_conditions",
              "mdqt_io_append_energies", "mdqt_io_write_vel_dist", "mdqt_io_write_populations", "mdqt_schedule_next"]


def _lib():
    L = load_library()
    if getattr(L, "_io_ready", False):
        return L
    d, i, u, vp, cp = ctypes.c_double, ctypes.c_int, ctypes.c_uint, ctypes.c_void_p, ctypes.c_char_p
    L.mdqt_io_dirname.argtypes = [ctypes.c_char_p, i, cp] + [d] * 9 + [i, u, i]
    L.mdqt_io_init_su.argtypes = [ctypes.c_long, i, d, i, vp, vp, vp, vp, c_double_p, c_double_p]
    L.mdqt_io_write_conditions.argtypes = [cp, i, i, u, vp, vp, vp, i, vp]
    L.mdqt_io_read_conditions.argtypes = [cp, i, i, vp, vp, vp, ctypes.POINTER(u), c_double_p, vp]
    L.mdqt_io_append_energies.argtypes = [cp] + [d] * 7
    L.mdqt_io_write_vel_dist.argtypes = [cp, u, vp, d]
    L.mdqt_io_write_populations.argtypes = [cp, u, i, vp, vp]
    ip = ctypes.POINTER(ctypes.c_int)
    L.mdqt_schedule_next.argtypes = [ip, ip, c_double_p, i, i, d, d, ip, ip]
    L._io_ready = True
    return L


def _p(a):
    return None if a is None else ctypes.c_void_p(a.ctypes.data)


def dirname(saveDirectory="dataLaserCool/", Ge=0.1, density=2.0, sig0=4.0, Te=19.0, fracOfSig=0.0, detuning=-1.0,
            detuningDP=1.0, Om=1.0, OmDP=1.0, N0=3500, job=1, create=False):
    buf = ctypes.create_string_buffer(1024)
    if _lib().mdqt_io_dirname(buf, 1024, saveDirectory.encode(), Ge, density, sig0, Te, fracOfSig, detuning, detuningDP, Om,
                              OmDP, N0, job, 1 if create else 0):
        raise MDQTError("directory name too long")
    return buf.value.decode()


def init_su(seed, N0=3500, Ge=0.1):
    """init() of the reference: returns dict(N, R[3][N], V, psi[N][12][2], tPart, L, lDeb)."""
    ld = N0 + 1000
    R, V = np.zeros((3, ld)), np.zeros((3, ld))
    psi, tp = np.zeros((ld, 12, 2)), np.zeros(ld)
    L, lDeb = ctypes.c_double(), ctypes.c_double()
    n = _lib().mdqt_io_init_su(seed, N0, Ge, ld, _p(R), _p(V), _p(psi), _p(tp), ctypes.byref(L), ctypes.byref(lDeb))
    if n < 0:
        raise MDQTError("init drew more than N0+1000 ions")
    return dict(N=n, R=np.ascontiguousarray(R[:, :n]), V=np.ascontiguousarray(V[:, :n]), psi=psi[:n].copy(), tPart=tp[:n].copy(),
                L=L.value, lDeb=lDeb.value)


def write_conditions(dir_, c0, counter, R, V, psi, vholder=None):
    n = R.shape[1]
    R, V, psi = (np.ascontiguousarray(a, dtype=np.float64) for a in (R, V, psi))
    rc = _lib().mdqt_io_write_conditions(dir_.encode(), c0, n, counter, _p(R), _p(V), _p(psi), n, _p(vholder))
    if rc:
        raise MDQTError("write_conditions failed (%d)" % rc)


def read_conditions(dir_, c0, ld=4500):
    R, V = np.zeros((3, ld)), np.zeros((3, ld))
    psi = np.zeros((ld, 12, 2))
    vh = np.zeros((3, N_VINT, ld))
    counter, t = ctypes.c_uint(), ctypes.c_double()
    n = _lib().mdqt_io_read_conditions(dir_.encode(), c0, ld, _p(R), _p(V), _p(psi), ctypes.byref(counter), ctypes.byref(t), _p(vh))
    if n < 0:
        raise MDQTError("read_conditions failed (%d)" % n)
    return dict(N=n, R=np.ascontiguousarray(R[:, :n]), V=np.ascontiguousarray(V[:, :n]), psi=psi[:n].copy(), counter=counter.value,
                t=t.value, vholder=np.ascontiguousarray(vh[:, :, :n]))


def append_energies(dir_, t, ekx, eky, ekz, epot, epot0, vx_avg):
    _lib().mdqt_io_append_energies(dir_.encode(), t, ekx, eky, ekz, epot, epot0, vx_avg)


def write_vel_dist(dir_, counter, pvel, vx_avg):
    pvel = np.ascontiguousarray(pvel, dtype=np.float64)
    _lib().mdqt_io_write_vel_dist(dir_.encode(), counter, _p(pvel), vx_avg)


def write_populations(dir_, counter, Vx, pops):
    Vx, pops = np.ascontiguousarray(Vx, dtype=np.float64), np.ascontiguousarray(pops, dtype=np.float64)
    _lib().mdqt_io_write_populations(dir_.encode(), counter, Vx.shape[0], _p(Vx), _p(pops))


def schedule_next(c0, tsc, t, ratio=25, sampleFreq=40, dtq=0.002 / 25, tmax=30.0):
    """One call of mdqt_schedule_next: returns (n_substeps, do_output, do_forces, c0, tsc, t)."""
    a, b, tt = ctypes.c_int(c0), ctypes.c_int(tsc), ctypes.c_double(t)
    o, f = ctypes.c_int(), ctypes.c_int()
    n = _lib().mdqt_schedule_next(ctypes.byref(a), ctypes.byref(b), ctypes.byref(tt), ratio, sampleFreq, dtq, tmax,
                                  ctypes.byref(o), ctypes.byref(f))
    return n, bool(o.value), bool(f.value), a.value, b.value, tt.value
