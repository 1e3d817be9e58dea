"""GPU parity cases written after this round's GPU budget was spent: their first run on a B200 is the driver's round-end run.
The file sorts last on purpose.  Each case combines code paths that were verified on the GPU separately
(profiles/r01m_gc_check.log, r01n_pytest_gpu.log) and together in the CPU oracle."""
import numpy as np
import pytest

from common import make_gpu, make_oracle
from test_gpu_parity import check_state

pytestmark = pytest.mark.gpu


def test_cy50r1_configuration_matches_oracle(built):
    """tests/etopo1_oper_an_fc_O48_cy50r1.yml: LLGCBZ0 + LLNORMAGAM + LCIWA3 + LCISCAL in one run (waves under the ice allowed so
    that the attenuation acts)."""
    okw = dict(llgcbz0=1, llnormagam=1, wspmin=0.3, lmaskice=0)
    g, o, f, fl = make_oracle("o48like", lciwa3=1, lciscal=1, **okw)
    _, s, w = make_gpu("o48like", lciwa=12, **okw)
    cith = np.where(f["CICOVER"] > 0, 0.3 + 1.5 * f["CICOVER"], 0.0)
    o.set_field("CITHICK", cith)
    w.set_field("cithick", cith)
    for _ in range(3):
        assert o.step() == 0 and w.step() == 0
    w.synchronize()
    check_state(w, o)


def test_depth_limited_points_with_the_gravity_capillary_physics(built):
    """SDEPTHLIM's factor inside the HALPHAP pass of the cy49r1 instance of k_point (the pass is repeated with the factor for the
    lanes that turn out to be depth-limited), same shallow grid as test_gpu_parity.test_depth_limited_points."""
    def shallow(g):
        n = g.depth.size
        g.depth[n // 2:] = np.minimum(g.depth[n // 2:], 2.0 + 3.0 * (np.arange(n - n // 2) % 7 == 0))
    kw = dict(llgcbz0=1, llnormagam=1, wspmin=0.3)
    g, o, f, fl = make_oracle("o640like", grid_hook=shallow, **kw)
    _, s, w = make_gpu("o640like", grid_hook=shallow, **kw)
    emax = 0.0625 * (0.8 * g.depth) ** 2
    hs_o, _ = o.hs_fm()
    frac = np.mean(hs_o ** 2 / 16.0 > emax)
    assert 0.02 < frac < 0.5
    for _ in range(2):
        assert o.step() == 0 and w.step() == 0
    w.synchronize()
    check_state(w, o)
