"""Thin host drivers: the time loops of the reference programs other than SU, expressed on the C ABI.

The SU production loop lives in C++ (``csrc/mdqt_driver.cpp`` -> ``mdqt_run``). The loops here are a few lines each in
the reference and stay in Python: they only sequence calls that run on the GPU.
"""


def fz_main_loop(eng, tmax, tstartV0, tendV0, c0=-1, sampleFreq=40, new_run=True, on_measure=None, on_sample=None, zfunc=None):
    """The time loop of randomFrozenStartTag408Linear.cpp (FZ408L:1040-1072) on an :class:`Engine` created with the
    7-level scheme, the SU box and ``substeps_per_md = plasmaToQuantumTimestepRatio``:

        while t <= tmax + 0.0009:
            if not recordedSpinUps and t >= tendV0:  measureSpinUps(); output(); Zfunc(0); printVAF(t)
            if (c0+1) % sampleFreq == 0 and timeStepCounter == 1 and recordedSpinUps:  output(); Zfunc(1); printVAF(t)
            if timeStepCounter == ratio:  step(); c0++; timeStepCounter = 0
            if tstartV0 < t < tendV0:  qstep()   (which also does t += quantumTimestep, FZ408L:597)
            else:  t += quantumTimestep
            timeStepCounter++

    File output is the caller's business: ``on_measure(t, tagged, n_up, vaf)`` and ``on_sample(t, c0, vaf)`` are invoked
    where the reference calls ``output()``/``printVAF``. Consecutive pump-window ``qstep()`` calls between two events are
    fused into one launch; ``t`` advances by the reference's repeated addition on host and device alike.
    ``zfunc`` replaces ``eng.Zfunc`` -- the Quad program correlates v_x^2 instead (``eng.ZfuncLongKin``, FZ408Q:942-967); the
    422 nm program is the same loop on an engine with the 5-level scheme (its main() only omits the output() at the measurement).
    Returns ``dict(c0, iters, tagged, n_up, vaf_measure, vaf_last, t)``.
    """
    zfunc = zfunc or eng.Zfunc
    p = eng.params
    ratio, dtq = int(p.substeps_per_md), float(p.dtq)
    t = eng.time()[0]
    tsc = ratio
    recorded = not new_run
    pend_q = 0   # pump-window qstep() calls not yet launched
    pend_t = 0   # time increments not yet applied to the engine's clock
    out = dict(tagged=None, n_up=None, vaf_measure=None, vaf_last=None)
    iters = 0

    def flush():
        nonlocal pend_q, pend_t
        if pend_q:
            eng.qstep7(pend_q)
            pend_q = 0
        if pend_t:
            eng.advance_time(pend_t)
            pend_t = 0

    while t <= tmax + 0.0009:
        if not recorded and t >= tendV0:
            flush()
            tagged, n_up = eng.measureSpinUps()
            recorded = True
            vaf = zfunc(0)
            out.update(tagged=tagged, n_up=n_up, vaf_measure=vaf)
            if on_measure:
                on_measure(t, tagged, n_up, vaf)
        if (c0 + 1) % sampleFreq == 0 and tsc == 1 and recorded:
            flush()
            vaf = zfunc(1)
            out["vaf_last"] = vaf
            if on_sample:
                on_sample(t, c0, vaf)
        if tsc == ratio:
            flush()                    # step() tests the engine's clock (2nd-order start while t <= 0, FZ408L:321)
            eng.step(dtq * ratio)      # dt = quantumTimestep*plasmaToQuantumTimestepRatio (FZ408L:382)
            c0 += 1
            tsc = 0
        if tstartV0 < t < tendV0:
            if pend_t and not pend_q:
                flush()
            pend_q += 1
        pend_t += 1
        t += dtq
        tsc += 1
        iters += 1
    flush()
    out.update(c0=c0, iters=iters, t=t)
    return out


def mc_pump_and_tag(eng, pumpMDTimeSteps, timeStep=0.005):
    """Step 5 of the MC-family mains (MC408L:1222-1233, MC422L:1189-1200): the pump stage

        for k < pumpMDTimeSteps:  { for l < plasmaToQuantumTimestepRatio: qstep(); }  MDStep(k)
        tagParticles()

    at collisionFreq = 0. Returns (tagged, n_tagged)."""
    ratio = int(eng.params.substeps_per_md)
    eng.MDSteps(int(pumpMDTimeSteps), dt=timeStep, qsteps=ratio)  # one replayed CUDA graph; same bits as the single calls
    return eng.tagParticles()


def md_record_stage(eng, numSteps, Gamma, timeStep=0.005, pairPairStep=0.05, gr_every=100, on_gr=None):
    """Steps 6-7 of the MD-family mains (MD:1107-1165, MC408L:1235-1253): collisionless MD with the recorders

        for k < numVelAutoCorrsSteps:  [every 100: recordPairPairCorr(k)]  MDStep(k)  recordVelsForAutocorrelations(k)
        recordVAF(); recordLongViscAutoCorr(); recordVCubeAutoCorr(); recordVFourthAutoCorr()

    Returns the four autocorrelation arrays [4][numSteps]; ``on_gr(k, r, g)`` receives each g(r)."""
    eng.vstore_begin(int(numSteps))
    for k in range(int(numSteps)):
        if k % gr_every == 0 and on_gr:
            r, g, _ = eng.recordPairPairCorr(pairPairStep)
            on_gr(k, r, g)
        eng.MDStep(dt=timeStep)
        eng.recordVelsForAutocorrelations(k)
    return eng.autocorrelations(Gamma)
