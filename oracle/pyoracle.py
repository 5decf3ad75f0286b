"""oracle/pyoracle.py -- TEST INFRASTRUCTURE ONLY.

ctypes front-ends for (a) ``liboracle.so`` -- the independent CPU restatement of the hot path
(``mdqt_oracle.c``) -- and (b) ``_ref/libref_*.so`` -- the UNMODIFIED reference programs compiled through
the hijack harnesses (``ref_*_harness.cpp``).  Only ``tests/``, ``__graft_entry__.smoke()`` and the
CPU-baseline legs of ``bench.py`` may import this module; the product package never does.

All arrays are float64 numpy, C-contiguous: R, V, F, A are ``[3][n]``; psi is ``[n][S][2]``.
"""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
c_double_p = ctypes.POINTER(ctypes.c_double)
c_int_p = ctypes.POINTER(ctypes.c_int)


def _dp(a):
    if a is None:
        return None
    assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(c_double_p)


def build(quiet=True):
    """Build liboracle.so (always) and, when the reference sources are present, oracle/_ref/*.so."""
    out = subprocess.run(["make", "-C", HERE, "all"], capture_output=True, text=True)
    if out.returncode != 0:
        raise RuntimeError("oracle build failed:\n" + out.stdout + out.stderr)
    if not quiet:
        print(out.stdout)


class QTParams(ctypes.Structure):
    _fields_ = [(k, ctypes.c_double) for k in (
        "detuning", "detuningDP", "Om", "OmDP", "dR", "kRat", "vKick", "vKickDP", "g2E", "pv2qv", "dtq",
        "fracOfSig", "Te", "sig0", "density")] + [("renorm", ctypes.c_int), ("quad", ctypes.c_int)]


def su_params(Ge=0.1, density=2.0, sig0=4.0, Te=19.0, fracOfSig=0.0, detuning=-1.0, detuningDP=1.0, Om=1.0,
              OmDP=1.0, renorm=0):
    """Derived constants exactly as the reference evaluates them (SU:79-85, 146-149, 295-297)."""
    import math
    ratio = int(math.ceil(34.81 / math.sqrt(density)))
    pv2qv = 1.1821 * math.pow(density, 1. / 6)
    vKick = 0.001208 / pv2qv
    kRat = 0.395
    p = QTParams(detuning=detuning, detuningDP=detuningDP, Om=Om, OmDP=OmDP, dR=0.0617, kRat=kRat, vKick=vKick,
                 vKickDP=vKick * kRat, g2E=174.07 / math.sqrt(density), pv2qv=pv2qv, dtq=0.002 / ratio,
                 fracOfSig=fracOfSig, Te=Te, sig0=sig0, density=density, renorm=renorm, quad=0)
    return p, ratio


def mc408_params(n=2.0, detuning=-2.5, Om=0.7, quad=0, timeStep=0.005):
    """Derived constants of the 7-level pump stage (MC408L:85-87, 115-122)."""
    import math
    ratio = int(round(87 / math.sqrt(n)))
    pv2qv = 1.1821 * math.pow(n, 1. / 6)
    p = QTParams(detuning=detuning, detuningDP=0.0, Om=Om, OmDP=0.0, dR=0.0617, kRat=0.0, vKick=0.001208 / pv2qv,
                 vKickDP=0.0, g2E=174.07 / math.sqrt(n), pv2qv=pv2qv, dtq=timeStep / ratio, fracOfSig=0.0, Te=0.0,
                 sig0=1.0, density=n, renorm=0, quad=quad)
    return p, ratio


def mc422_params(n=2.0, detuning=-1.0, Om=1.3, timeStep=0.005):
    """Derived constants of the 5-level 422 nm pump stage (MC422L:84-86, 113-121)."""
    import math
    ratio = int(round(87 * .894 / math.sqrt(n)))
    pv2qv = 1.1821 * math.pow(n, 1. / 6) * .967
    p = QTParams(detuning=detuning, detuningDP=0.0, Om=Om, OmDP=0.0, dR=0.0753, kRat=0.0, vKick=0.001257 / pv2qv,
                 vKickDP=0.0, g2E=174.07 * .894 / math.sqrt(n), pv2qv=pv2qv, dtq=timeStep / ratio, fracOfSig=0.0, Te=0.0,
                 sig0=1.0, density=n, renorm=0, quad=0)
    return p, ratio


def ts_params(detuning=-0.5, Om=0.5, dt=0.01):
    """The 3-level test program (TS:55-58, 91, 390): everything already in quantum units."""
    return QTParams(detuning=detuning, detuningDP=0.0, Om=Om, OmDP=0.0, dR=0.0, kRat=0.0, vKick=0.0012076, vKickDP=0.0,
                    g2E=1.0, pv2qv=1.0, dtq=dt, fracOfSig=0.0, Te=0.0, sig0=1.0, density=1.0, renorm=0, quad=0)


class Oracle:
    """The independent restatement (liboracle.so)."""

    def __init__(self):
        path = os.path.join(HERE, "liboracle.so")
        if not os.path.exists(path):
            build()
        L = self.lib = ctypes.CDLL(path)
        L.orc_epot_su.restype = ctypes.c_double
        L.orc_forces_su.argtypes = [ctypes.c_int, c_double_p, ctypes.c_double, ctypes.c_double, c_double_p]
        L.orc_forces_md.argtypes = [ctypes.c_int, c_double_p, ctypes.c_double, ctypes.c_double, ctypes.c_double, c_double_p]
        L.orc_epot_su.argtypes = [ctypes.c_int, c_double_p, ctypes.c_double, ctypes.c_double]
        L.orc_step_su.argtypes = [ctypes.c_int, c_double_p, c_double_p, c_double_p] + [ctypes.c_double] * 3
        L.orc_vv_positions.argtypes = [ctypes.c_int, c_double_p, c_double_p, c_double_p, ctypes.c_double, ctypes.c_double]
        L.orc_vv_velocities.argtypes = [ctypes.c_int, c_double_p, c_double_p, c_double_p, ctypes.c_double, ctypes.c_double,
                                        c_double_p, c_double_p, ctypes.c_int, ctypes.c_double, ctypes.c_double]
        L.orc_qstep12.argtypes = [ctypes.c_int, c_double_p, c_double_p, c_double_p, c_double_p, ctypes.POINTER(QTParams),
                                  c_double_p, ctypes.c_int, ctypes.POINTER(ctypes.c_long), c_int_p]
        L.orc_qstep7.argtypes = [ctypes.c_int, c_double_p, c_double_p, ctypes.POINTER(QTParams), c_double_p, ctypes.c_int,
                                 ctypes.POINTER(ctypes.c_long), c_int_p]
        L.orc_qstep5.argtypes = L.orc_qstep7.argtypes
        L.orc_qstep3.argtypes = [ctypes.c_int, c_double_p, c_double_p, c_double_p, ctypes.POINTER(QTParams), ctypes.c_int,
                                 c_double_p, ctypes.c_int, ctypes.POINTER(ctypes.c_long), c_int_p]
        L.orc_tag.argtypes = [ctypes.c_int, ctypes.c_int, c_double_p, c_double_p, ctypes.c_int,
                              ctypes.POINTER(ctypes.c_long), c_int_p]
        L.orc_lf_drift.argtypes = [ctypes.c_int, c_double_p, c_double_p, c_double_p, ctypes.c_double, ctypes.c_double, ctypes.c_int]
        L.orc_lf_kick.argtypes = [ctypes.c_int, c_double_p, c_double_p, ctypes.c_double]
        L.orc_vaf.argtypes = [ctypes.c_int, c_double_p, c_double_p]
        L.orc_vaf.restype = ctypes.c_double
        L.orc_pair_correlation.argtypes = [ctypes.c_int, c_double_p, ctypes.c_double, ctypes.c_double, ctypes.c_double, c_double_p, c_double_p]
        L.orc_autocorr.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, c_double_p, ctypes.c_double, c_double_p]
        L.orc_uniforms5.argtypes = [ctypes.c_uint64, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint64, c_double_p]
        L.orc_collision_draws.argtypes = [ctypes.c_uint64, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint64, c_double_p, c_double_p]

    def philox(self, ctr, key):
        c = (ctypes.c_uint32 * 4)(*ctr)
        k = (ctypes.c_uint32 * 2)(*key)
        o = (ctypes.c_uint32 * 4)()
        self.lib.orc_philox4x32_10(c, k, o)
        return list(o)

    def uniforms5(self, seed, traj, n, substep):
        """u[n][5] for ions 0..n-1 of trajectory `traj` at substep index `substep`."""
        u = np.empty((n, 5))
        for i in range(n):
            self.lib.orc_uniforms5(seed, traj, i, substep, _dp(u[i]))
        return u

    def collision_draws(self, seed, traj, n, step):
        u = np.empty(n)
        nrm = np.empty((n, 3))
        one = np.empty(1)
        for i in range(n):
            self.lib.orc_collision_draws(seed, traj, i, step, _dp(one), _dp(nrm[i]))
            u[i] = one[0]
        return u, nrm

    def forces_su(self, R, L, lDeb):
        F = np.empty_like(R)
        self.lib.orc_forces_su(R.shape[1], _dp(R), L, lDeb, _dp(F))
        return F

    def forces_md(self, R, L, kappa, rCut):
        A = np.empty_like(R)
        self.lib.orc_forces_md(R.shape[1], _dp(R), L, kappa, rCut, _dp(A))
        return A

    def epot_su(self, R, L, lDeb):
        return self.lib.orc_epot_su(R.shape[1], _dp(R), L, lDeb)

    def step_su(self, R, V, F, L, dtq, t):
        self.lib.orc_step_su(R.shape[1], _dp(R), _dp(V), _dp(F), L, dtq, t)

    def vv_positions(self, R, V, A, L, dt):
        self.lib.orc_vv_positions(R.shape[1], _dp(R), _dp(V), _dp(A), L, dt)

    def vv_velocities(self, V, oldA, A, dt, collisionFreq=0.0, coll_u=None, coll_n=None, laser=0, beta=26000.0, dens=0.4):
        self.lib.orc_vv_velocities(V.shape[1], _dp(V), _dp(oldA), _dp(A), dt, collisionFreq, _dp(coll_u), _dp(coll_n),
                                   laser, beta, dens)

    def qstep12(self, psi, Vx, tPart, t, p, u, sequential=False):
        """One qstep() sweep in place; returns (new t, draws used per ion)."""
        n = psi.shape[0]
        tt = np.array([t], dtype=np.float64)
        cur = ctypes.c_long(0)
        used = np.zeros(n, dtype=np.int32)
        self.lib.orc_qstep12(n, _dp(psi), _dp(Vx), _dp(tPart), _dp(tt), ctypes.byref(p), _dp(u), 1 if sequential else 0,
                             ctypes.byref(cur), used.ctypes.data_as(c_int_p))
        return float(tt[0]), used

    def qstep7(self, psi, Vx, p, u, sequential=False):
        n = psi.shape[0]
        cur = ctypes.c_long(0)
        used = np.zeros(n, dtype=np.int32)
        self.lib.orc_qstep7(n, _dp(psi), _dp(Vx), ctypes.byref(p), _dp(u), 1 if sequential else 0, ctypes.byref(cur),
                            used.ctypes.data_as(c_int_p))
        return used

    def qstep5(self, psi, Vx, p, u, sequential=False):
        """5-level 422 nm pump sweep (MC422L:552-727); table slots: 0 rand, 1 rand2, 2 randDOrS, 4 rand3."""
        n = psi.shape[0]
        cur = ctypes.c_long(0)
        used = np.zeros(n, dtype=np.int32)
        self.lib.orc_qstep5(n, _dp(psi), _dp(Vx), ctypes.byref(p), _dp(u), 1 if sequential else 0, ctypes.byref(cur),
                            used.ctypes.data_as(c_int_p))
        return used

    def qstep3(self, psi, Vx, tPart, p, u, sequential=False, applyForce=True):
        """3-level sweep (TS:140-293), Vx and tPart updated in place; table slots: 0 rand, 3 randDir."""
        n = psi.shape[0]
        cur = ctypes.c_long(0)
        used = np.zeros(n, dtype=np.int32)
        self.lib.orc_qstep3(n, _dp(psi), _dp(Vx), _dp(tPart), ctypes.byref(p), 1 if applyForce else 0, _dp(u),
                            1 if sequential else 0, ctypes.byref(cur), used.ctypes.data_as(c_int_p))
        return used

    def tag(self, psi, u, sequential=False):
        """tagParticles / measureSpinUps; returns (tagged[n] int32, draws consumed when sequential)."""
        n, S = psi.shape[0], psi.shape[1]
        cur = ctypes.c_long(0)
        tagged = np.zeros(n, dtype=np.int32)
        cnt = self.lib.orc_tag(n, S, _dp(psi), _dp(u), 1 if sequential else 0, ctypes.byref(cur), tagged.ctypes.data_as(c_int_p))
        assert cnt == tagged.sum()
        return tagged, cur.value

    def lf_step(self, R, V, L, lDeb, dt, first):
        """FZ-family step() (FZ408L:377-390) with the restated force routine; returns the last F."""
        n = R.shape[1]
        if first:
            self.lib.orc_lf_drift(n, _dp(R), _dp(V), _dp(self.forces_su(R, L, lDeb)), L, 0.5 * dt, 1)
        else:
            self.lib.orc_lf_drift(n, _dp(R), _dp(V), _dp(V), L, 0.5 * dt, 0)
        F = self.forces_su(R, L, lDeb)
        self.lib.orc_lf_kick(n, _dp(V), _dp(F), dt)
        if first:
            F = self.forces_su(R, L, lDeb)
        self.lib.orc_lf_drift(n, _dp(R), _dp(V), _dp(F), L, 0.5 * dt, 1 if first else 0)
        return F

    def pair_correlation(self, R, L, step, rmax):
        """recordPairPairCorr (MD:584-652): (raw ordered-pair counts, normalised g)."""
        nb = int(rmax / step)
        counts, g = np.zeros(nb), np.zeros(nb)
        self.lib.orc_pair_correlation(R.shape[1], _dp(R), L, step, rmax, _dp(counts), _dp(g))
        return counts, g

    def autocorr(self, which, vstore, Gamma, nnorm=None):
        """MD:654-823 on vstore[3][n][T]; nnorm = N of the normalisation (series beyond n count as zeros)."""
        _, n, T = vstore.shape
        out = np.empty(T)
        self.lib.orc_autocorr(which, n, n if nnorm is None else nnorm, T, _dp(np.ascontiguousarray(vstore)), Gamma, _dp(out))
        return out

    def vaf(self, Vhold, Vx):
        return self.lib.orc_vaf(Vx.shape[0], _dp(np.ascontiguousarray(Vhold)), _dp(np.ascontiguousarray(Vx)))


def ref_available(name="su"):
    return os.path.exists(os.path.join(HERE, "_ref", "libref_%s.so" % name))


class RefSU:
    """The unmodified reference SU program behind the hijack harness (oracle/_ref/libref_su.so)."""

    def __init__(self, **kw):
        self.lib = L = ctypes.CDLL(os.path.join(HERE, "_ref", "libref_su.so"))
        L.ref_su_epot.restype = ctypes.c_double
        L.ref_su_get_t.restype = ctypes.c_double
        L.ref_su_get_Epot0.restype = ctypes.c_double
        L.ref_su_set_t.argtypes = [ctypes.c_double]
        L.ref_su_set_Epot0.argtypes = [ctypes.c_double]
        L.ref_su_set_box.argtypes = [ctypes.c_double, ctypes.c_double]
        L.ref_su_init.argtypes = [ctypes.c_long]
        L.ref_su_set_state.argtypes = [ctypes.c_int, c_double_p, c_double_p, c_double_p, c_double_p]
        L.ref_su_get_state.argtypes = [c_double_p] * 5
        L.ref_su_set_F.argtypes = [c_double_p]
        L.ref_su_set_uniforms.argtypes = [c_double_p, ctypes.c_int]
        L.ref_su_qstep_stream.argtypes = [c_double_p, c_int_p]
        L.ref_su_get_tables.argtypes = [c_double_p] * 4
        L.ref_su_set_savedir.argtypes = [ctypes.c_char_p]
        L.ref_su_set_counters.argtypes = [ctypes.c_int, ctypes.c_uint]
        L.ref_su_run_loop.argtypes = [ctypes.c_int] * 3
        self.setup(**kw)

    def setup(self, Ge=0.1, density=2.0, sig0=4.0, Te=19.0, fracOfSig=0.0, detuning=-1.0, detuningDP=1.0, Om=1.0,
              OmDP=1.0, renorm=0):
        p = (ctypes.c_double * 10)(Ge, density, sig0, Te, fracOfSig, detuning, detuningDP, Om, OmDP, renorm)
        assert self.lib.ref_su_setup(p) == 0
        c = (ctypes.c_double * 12)()
        self.lib.ref_su_get_consts(c)
        keys = ["L", "lDeb", "dtq", "g2E", "pv2qv", "vKick", "vKickDP", "ratio", "dR", "kRat", "TIMESTEP", "sampleFreq"]
        self.consts = dict(zip(keys, list(c)))

    def set_box(self, L, lDeb):
        self.lib.ref_su_set_box(L, lDeb)
        self.consts["L"], self.consts["lDeb"] = L, lDeb

    def init(self, seed):
        self.lib.ref_su_init(seed)
        return self.lib.ref_su_get_N()

    @property
    def N(self):
        return self.lib.ref_su_get_N()

    def set_state(self, R=None, V=None, psi=None, tPart=None, t=None, n=None):
        if n is None:
            n = (R if R is not None else V if V is not None else None)
            n = n.shape[1] if n is not None else (psi.shape[0] if psi is not None else tPart.shape[0])
        self.lib.ref_su_set_state(n, _dp(R), _dp(V), _dp(psi), _dp(tPart))
        if t is not None:
            self.lib.ref_su_set_t(t)

    def get_state(self):
        n = self.N
        R, V, F = np.empty((3, n)), np.empty((3, n)), np.empty((3, n))
        psi, tp = np.empty((n, 12, 2)), np.empty(n)
        self.lib.ref_su_get_state(_dp(R), _dp(V), _dp(F), _dp(psi), _dp(tp))
        return dict(R=R, V=V, F=F, psi=psi, tPart=tp, t=self.lib.ref_su_get_t())

    def set_F(self, F):
        self.lib.ref_su_set_F(_dp(F))

    def forces(self):
        self.lib.ref_su_forces()

    def step(self):
        self.lib.ref_su_step()

    def qstep(self, uniforms=None):
        """qstep() with either the reference's own drand48 stream, or a sequential injected stream."""
        if uniforms is not None:
            self._u = np.ascontiguousarray(uniforms, dtype=np.float64)
            self.lib.ref_su_set_uniforms(_dp(self._u), self._u.size)
        self.lib.ref_su_qstep()
        used = self.lib.ref_su_uniforms_used()
        self.lib.ref_su_set_uniforms(None, 0)
        return used

    def qstep_stream(self, u5):
        """qstep() sweep where ion i consumes u5[i][:]; returns draws used per ion."""
        u5 = np.ascontiguousarray(u5, dtype=np.float64)
        used = np.zeros(self.N, dtype=np.int32)
        self.lib.ref_su_qstep_stream(_dp(u5), used.ctypes.data_as(c_int_p))
        return used

    def epot(self):
        return self.lib.ref_su_epot()

    # ---- driver-level functions of the reference (directory naming, restart and output files, main loop) ----
    def savedir(self):
        self.lib.ref_su_get_savedir.restype = ctypes.c_char_p
        return self.lib.ref_su_get_savedir().decode()

    def set_savedir(self, d):
        self.lib.ref_su_set_savedir(d.encode())

    def set_counters(self, c0, counter):
        self.lib.ref_su_set_counters(c0, counter)

    def counters(self):
        return self.lib.ref_su_get_c0(), self.lib.ref_su_get_counter()

    def write_conditions(self, c0):
        self.lib.ref_su_write_conditions(c0)

    def read_conditions(self, c0):
        self.lib.ref_su_read_conditions(c0)

    def output(self):
        self.lib.ref_su_output()

    def run_until(self, tmax, do_output=1):
        """The reference main loop (SU:1248-1381) from the current state; returns the number of substeps."""
        self.lib.ref_su_run_until.restype = ctypes.c_long
        self.lib.ref_su_run_until.argtypes = [ctypes.c_double, ctypes.c_int, ctypes.c_int]
        return self.lib.ref_su_run_until(tmax, int(self.consts["ratio"]), do_output)

    def tables(self):
        c, d, hd = np.empty((12, 12, 2)), np.empty((12, 12, 2)), np.empty((12, 12, 2))
        gs = np.empty(18)
        self.lib.ref_su_get_tables(_dp(c), _dp(d), _dp(hd), _dp(gs))
        return c[..., 0] + 1j * c[..., 1], d[..., 0] + 1j * d[..., 1], hd[..., 0] + 1j * hd[..., 1], gs


class RefMD:
    """The unmodified reference MD-only program (oracle/_ref/libref_md.so); N=4096 fixed."""

    def __init__(self):
        self.lib = L = ctypes.CDLL(os.path.join(HERE, "_ref", "libref_md.so"))
        L.ref_md_set_controls.argtypes = [ctypes.c_double, ctypes.c_int, ctypes.c_int]
        L.ref_md_set_state.argtypes = [c_double_p] * 3
        L.ref_md_get_state.argtypes = [c_double_p] * 3
        c = (ctypes.c_double * 7)()
        L.ref_md_get_consts(c)
        self.consts = dict(zip(["L", "rCut", "kappa", "Gamma", "n", "timeStep", "beta"], list(c)))
        self.N = L.ref_md_N()

    def seed(self, s):
        self.lib.ref_md_seed(s)

    def init(self):
        self.lib.ref_md_init()

    def set_controls(self, collisionFreq=0.0, laser=0, one_axis=0):
        self.lib.ref_md_set_controls(collisionFreq, laser, one_axis)

    def set_state(self, R=None, V=None, A=None):
        self.lib.ref_md_set_state(_dp(R), _dp(V), _dp(A))

    def get_state(self):
        R, V, A = (np.empty((3, self.N)) for _ in range(3))
        self.lib.ref_md_get_state(_dp(R), _dp(V), _dp(A))
        return dict(R=R, V=V, A=A)

    def accelerations(self):
        self.lib.ref_md_accelerations()

    def pair_correlation(self, scratch="/tmp/mdqt_ref_scratch_md/"):
        """recordPairPairCorr() through its output file (values carry the 6 significant digits of %lg)."""
        self.lib.ref_md_pair_correlation.argtypes = [ctypes.c_char_p, c_double_p, c_double_p, ctypes.c_int]
        self.lib.ref_md_pair_step.restype = ctypes.c_double
        self.lib.ref_md_pair_max.restype = ctypes.c_double
        r, g = np.zeros(4096), np.zeros(4096)
        k = self.lib.ref_md_pair_correlation(scratch.encode(), _dp(r), _dp(g), 4096)
        assert k > 0
        return r[:k].copy(), g[:k].copy(), self.lib.ref_md_pair_step(), self.lib.ref_md_pair_max()

    @property
    def T(self):
        return self.lib.ref_md_autocorr_steps()

    def set_vstore(self, v):
        """vStore[3][N][T]: the first v.shape[1] ions from v, zeros beyond."""
        self.lib.ref_md_set_vstore.argtypes = [ctypes.c_int, c_double_p]
        self.lib.ref_md_set_vstore(v.shape[1], _dp(np.ascontiguousarray(v)))

    def autocorr(self, which, scratch="/tmp/mdqt_ref_scratch_md/"):
        self.lib.ref_md_autocorr.argtypes = [ctypes.c_int, ctypes.c_char_p, c_double_p]
        out = np.empty(self.T)
        self.lib.ref_md_autocorr(which, scratch.encode(), _dp(out))
        return out

    def mdstep(self):
        self.lib.ref_md_mdstep()


class RefMC408L:
    """The unmodified reference MC408L program (7-level pump; oracle/_ref/libref_mc408l.so); N=4096 fixed."""

    def __init__(self, detuning=-2.5, Om=0.7, scratch="/tmp/mdqt_ref_scratch/"):
        self.lib = L = ctypes.CDLL(os.path.join(HERE, "_ref", "libref_mc408l.so"))
        L.ref_mc_setup.argtypes = [c_double_p, ctypes.c_char_p]
        L.ref_mc_set_state.argtypes = [c_double_p] * 4
        L.ref_mc_get_state.argtypes = [c_double_p] * 4
        L.ref_mc_set_uniforms.argtypes = [c_double_p, ctypes.c_long]
        L.ref_mc_uniforms_used.restype = ctypes.c_long
        L.ref_mc_set_controls.argtypes = [ctypes.c_double]
        p = (ctypes.c_double * 2)(detuning, Om)
        assert L.ref_mc_setup(p, scratch.encode()) == 0
        c = (ctypes.c_double * 12)()
        L.ref_mc_get_consts(c)
        keys = ["L", "rCut", "kappa", "Gamma", "n", "timeStep", "g2E", "ratio", "dtq", "pv2qv", "dR", "pumpMDTimeSteps"]
        self.consts = dict(zip(keys, list(c)))
        self.N = L.ref_mc_N()

    def set_state(self, R=None, V=None, A=None, psi=None):
        self.lib.ref_mc_set_state(_dp(R), _dp(V), _dp(A), _dp(psi))

    def get_state(self):
        R, V, A = (np.empty((3, self.N)) for _ in range(3))
        psi = np.empty((self.N, 7, 2))
        self.lib.ref_mc_get_state(_dp(R), _dp(V), _dp(A), _dp(psi))
        return dict(R=R, V=V, A=A, psi=psi)

    def qstep(self, uniforms):
        self._u = np.ascontiguousarray(uniforms, dtype=np.float64)
        self.lib.ref_mc_set_uniforms(_dp(self._u), self._u.size)
        self.lib.ref_mc_qstep()
        used = self.lib.ref_mc_uniforms_used()
        self.lib.ref_mc_set_uniforms(None, 0)
        return used

    def mdstep(self):
        self.lib.ref_mc_mdstep()

    def tag(self, uniforms):
        """tagParticles() (MC408L:1022-1067) on the current wavefunctions with one sequential uniform stream."""
        self._u = np.ascontiguousarray(uniforms, dtype=np.float64)
        self.lib.ref_mc_set_uniforms(_dp(self._u), self._u.size)
        out = np.zeros(self.N, dtype=np.int32)
        self.lib.ref_mc_tag(out.ctypes.data_as(c_int_p))
        used = self.lib.ref_mc_uniforms_used()
        self.lib.ref_mc_set_uniforms(None, 0)
        return out, used


class RefMC422L:
    """The unmodified reference MC422L program (5-level 422 nm pump; oracle/_ref/libref_mc422l.so); N=4096 fixed."""

    def __init__(self, detuning=-1.0, Om=1.3, scratch="/tmp/mdqt_ref_scratch422/"):
        self.lib = L = ctypes.CDLL(os.path.join(HERE, "_ref", "libref_mc422l.so"))
        L.ref_m422_setup.argtypes = [c_double_p, ctypes.c_char_p]
        L.ref_m422_set_state.argtypes = [c_double_p] * 4
        L.ref_m422_get_state.argtypes = [c_double_p] * 4
        L.ref_m422_set_uniforms.argtypes = [c_double_p, ctypes.c_long]
        L.ref_m422_uniforms_used.restype = ctypes.c_long
        p = (ctypes.c_double * 2)(detuning, Om)
        assert L.ref_m422_setup(p, scratch.encode()) == 0
        c = (ctypes.c_double * 12)()
        L.ref_m422_get_consts(c)
        keys = ["L", "rCut", "kappa", "Gamma", "n", "timeStep", "g2E", "ratio", "dtq", "pv2qv", "dR", "pumpMDTimeSteps"]
        self.consts = dict(zip(keys, list(c)))
        self.N = L.ref_m422_N()

    def set_state(self, R=None, V=None, A=None, psi=None):
        self.lib.ref_m422_set_state(_dp(R), _dp(V), _dp(A), _dp(psi))

    def get_state(self):
        R, V, A = (np.empty((3, self.N)) for _ in range(3))
        psi = np.empty((self.N, 5, 2))
        self.lib.ref_m422_get_state(_dp(R), _dp(V), _dp(A), _dp(psi))
        return dict(R=R, V=V, A=A, psi=psi)

    def _with_u(self, uniforms, fn):
        self._u = np.ascontiguousarray(uniforms, dtype=np.float64)
        self.lib.ref_m422_set_uniforms(_dp(self._u), self._u.size)
        r = fn()
        used = self.lib.ref_m422_uniforms_used()
        self.lib.ref_m422_set_uniforms(None, 0)
        return r, used

    def qstep(self, uniforms):
        return self._with_u(uniforms, self.lib.ref_m422_qstep)[1]

    def tag(self, uniforms):
        out = np.zeros(self.N, dtype=np.int32)
        _, used = self._with_u(uniforms, lambda: self.lib.ref_m422_tag(out.ctypes.data_as(c_int_p)))
        return out, used


class RefTS:
    """The unmodified reference 3-level program (oracle/_ref/libref_ts.so); N0=1000 fixed."""

    def __init__(self, detuning=-0.5, Om=0.5, dt=0.01, applyForce=True):
        self.lib = L = ctypes.CDLL(os.path.join(HERE, "_ref", "libref_ts.so"))
        L.ref_ts_setup.argtypes = [ctypes.c_double] * 3 + [ctypes.c_int]
        L.ref_ts_set_state.argtypes = [c_double_p] * 3
        L.ref_ts_get_state.argtypes = [c_double_p] * 3
        L.ref_ts_set_uniforms.argtypes = [c_double_p, ctypes.c_long]
        L.ref_ts_uniforms_used.restype = ctypes.c_long
        L.ref_ts_vkick.restype = ctypes.c_double
        assert L.ref_ts_setup(detuning, Om, dt, 1 if applyForce else 0) == 0
        self.N = L.ref_ts_N()
        self.vKick = L.ref_ts_vkick()

    def set_state(self, Vx=None, psi=None, tPart=None):
        self.lib.ref_ts_set_state(_dp(Vx), _dp(psi), _dp(tPart))

    def get_state(self):
        Vx, tp, psi = np.empty(self.N), np.empty(self.N), np.empty((self.N, 3, 2))
        self.lib.ref_ts_get_state(_dp(Vx), _dp(psi), _dp(tp))
        return dict(Vx=Vx, psi=psi, tPart=tp)

    def qstep(self, uniforms):
        self._u = np.ascontiguousarray(uniforms, dtype=np.float64)
        self.lib.ref_ts_set_uniforms(_dp(self._u), self._u.size)
        self.lib.ref_ts_qstep()
        used = self.lib.ref_ts_uniforms_used()
        self.lib.ref_ts_set_uniforms(None, 0)
        return used


class RefFZ408L:
    """The unmodified reference frozen-start pump-window program FZ408L (oracle/_ref/libref_fz408l.so)."""

    def __init__(self, detuning=-2.5, Om=0.7, scratch="/tmp/mdqt_ref_scratch_fz/"):
        self.lib = L = ctypes.CDLL(os.path.join(HERE, "_ref", "libref_fz408l.so"))
        L.ref_fz_setup.argtypes = [ctypes.c_double, ctypes.c_double, ctypes.c_char_p]
        L.ref_fz_init.argtypes = [ctypes.c_long]
        L.ref_fz_get_t.restype = ctypes.c_double
        L.ref_fz_set_t.argtypes = [ctypes.c_double]
        L.ref_fz_set_state.argtypes = [ctypes.c_int] + [c_double_p] * 3
        L.ref_fz_get_state.argtypes = [c_double_p] * 4
        L.ref_fz_set_uniforms.argtypes = [c_double_p, ctypes.c_long]
        L.ref_fz_uniforms_used.restype = ctypes.c_long
        L.ref_fz_measure.argtypes = [c_int_p]
        L.ref_fz_zfunc.argtypes = [ctypes.c_int]
        L.ref_fz_zfunc.restype = ctypes.c_double
        assert L.ref_fz_setup(detuning, Om, scratch.encode()) == 0
        c = (ctypes.c_double * 10)()
        L.ref_fz_get_consts(c)
        keys = ["L", "lDeb", "dtq", "g2E", "pv2qv", "ratio", "dR", "tpump", "tendV0", "TIMESTEP"]
        self.consts = dict(zip(keys, list(c)))

    @property
    def N(self):
        return self.lib.ref_fz_get_N()

    def init(self, seed):
        return self.lib.ref_fz_init(seed)

    def set_state(self, R=None, V=None, psi=None, t=None, n=None):
        if n is None:
            n = R.shape[1] if R is not None else (V.shape[1] if V is not None else psi.shape[0])
        self.lib.ref_fz_set_state(n, _dp(R), _dp(V), _dp(psi))
        if t is not None:
            self.lib.ref_fz_set_t(t)

    def get_state(self):
        n = self.N
        R, V, F = (np.empty((3, n)) for _ in range(3))
        psi = np.empty((n, 7, 2))
        self.lib.ref_fz_get_state(_dp(R), _dp(V), _dp(F), _dp(psi))
        return dict(R=R, V=V, F=F, psi=psi, t=self.lib.ref_fz_get_t())

    def step(self):
        self.lib.ref_fz_step()

    def _with_u(self, uniforms, fn):
        self._u = np.ascontiguousarray(uniforms, dtype=np.float64)
        self.lib.ref_fz_set_uniforms(_dp(self._u), self._u.size)
        r = fn()
        used = self.lib.ref_fz_uniforms_used()
        self.lib.ref_fz_set_uniforms(None, 0)
        return r, used

    def qstep(self, uniforms):
        return self._with_u(uniforms, self.lib.ref_fz_qstep)[1]

    def measure(self, uniforms):
        out = np.zeros(self.N, dtype=np.int32)
        cnt, used = self._with_u(uniforms, lambda: self.lib.ref_fz_measure(out.ctypes.data_as(c_int_p)))
        return out, cnt, used

    def zfunc(self, c1V):
        return self.lib.ref_fz_zfunc(c1V)

    def run_loop(self, tmax, tstart, tend, sampleFreq, uniforms):
        """The reference's main time loop (FZ408L:1040-1072; output()/printVAF() file writes left out) from the current
        state, new-run flags; `uniforms` = ONE sequential stream for every drand48 the loop consumes."""
        self.lib.ref_fz_run_loop.restype = ctypes.c_long
        self.lib.ref_fz_run_loop.argtypes = [ctypes.c_double] * 3 + [ctypes.c_int, c_double_p, c_int_p, c_int_p]
        vaf = np.zeros(2)
        spin = np.zeros(self.N, dtype=np.int32)
        nspin = ctypes.c_int(0)
        (iters), used = self._with_u(uniforms, lambda: self.lib.ref_fz_run_loop(tmax, tstart, tend, sampleFreq, _dp(vaf),
                                                                               spin.ctypes.data_as(c_int_p), ctypes.byref(nspin)))
        return dict(iters=iters, used=used, vaf=vaf, spin=spin, nspin=nspin.value, c0=self.lib.ref_fz_get_c0())

    def set_c0(self, c0):
        self.lib.ref_fz_set_c0(c0)
