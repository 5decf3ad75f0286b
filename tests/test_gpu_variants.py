"""GPU suite, part 3: the opt-in kernel mappings (environment knobs read once per process) stay parity-green:
MDQT_QT_LANES=4|2 (four / two lanes per ion in the fused substep kernel; the default picks by system size), MDQT_K1_ITEMS=0
(the CTA-tile force kernel where the item-walking kernel is the default), MDQT_ITEMS_IPT=2 (two rows per lane in the item kernel), MDQT_GRAPH=0
(stream launches instead of the replayed CUDA graph), MDQT_PDL=1 (programmatic dependent launch),
MDQT_CLUSTER=1 (j chunks combined through distributed shared memory inside a thread-block cluster) and the force-kernel plan
overrides. Each runs __graft_entry__.smoke() -- one MD step with jumps against the oracle -- plus the
no-jump and jump-table goldens in a fresh interpreter."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


T = {"MDQT_K1_ITEMS": "0"}  # the CTA-tile force kernel (default for N > 8192) instead of the item-walking kernel


@pytest.mark.parametrize("env", [{"MDQT_QT_LANES": "4"}, {"MDQT_QT_LANES": "2"}, {"MDQT_GRAPH": "0"}, {"MDQT_PDL": "1"},
                                 {"MDQT_ITEMS_IPT": "2"}, T, dict(T, MDQT_PDL="1"), dict(T, MDQT_CLUSTER="1"),
                                 dict(T, MDQT_CLUSTER="1", MDQT_FORCE_RG="32", MDQT_FORCE_JSUB="4", MDQT_FORCE_NSPLIT="2"),
                                 dict(T, MDQT_FORCE_IPT="2", MDQT_FORCE_JSUB="4"), dict(T, MDQT_FORCE_IPT="2", MDQT_FORCE_NSPLIT="3"),
                                 dict(T, MDQT_FORCE_JSUB="1", MDQT_FORCE_NSPLIT="7")])
def test_variant_parity(env):
    e = dict(os.environ, **env)
    r = subprocess.run([sys.executable, "-c", "import __graft_entry__ as g; g.smoke()"], cwd=ROOT, env=e, capture_output=True,
                       text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    r = subprocess.run([sys.executable, "-m", "pytest", "-q", "-m", "gpu", "tests/test_gpu_parity.py", "-k",
                        "nojump_golden or jump_table or forces_golden or trajectory_golden"], cwd=ROOT, env=e, capture_output=True,
                       text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
