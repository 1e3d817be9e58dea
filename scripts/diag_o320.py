"""Diagnostic: where does the full-size O320 case differ from the oracle?  (run on the GPU box)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
from ecwam_b200 import synth, model as M
from oracle import oracle as O

N = int(os.environ.get("DIAG_N", "320"))
A, Fr, P = 24, 29, 64
g = synth.make_grid(N, "continents")
cfg = O.default_config(nang=A, nfre_red=Fr, nproma=P, npr=1, iphys=1, idelt=900.0, idelpro=900.0, delpro_lf=900.0, nthreads=os.cpu_count() or 1)
o = O.Oracle(cfg, g, fast=True)
s = M.WamSetup(g, nproc=1, nang=A, nfre_red=Fr, nproma=P, idelt=900.0, idelpro=900.0, delpro_lf=900.0)
w = M.WamIntgr(s, 0)
w.set_static(g.depth)
f = synth.make_forcing(g)
for k, v in f.items():
    o.set_field(k, v); w.set_field(k, v)
fl = synth.jonswap_cold_start(f["WSWAVE"], f["WDWAVE"], A, 36, Fr)
o.set_fl1(fl); w.set_fl1(fl)


def report(tag):
    w.synchronize()
    a, b = w.get_spec("fl1"), o.get_fl1()[:, :, w.own]
    d = np.abs(a - b)
    print(tag, "max abs", d.max(), "rel to max", d.max() / b.max(), "n(>1e-12 max)", int((d > 1e-12 * b.max()).sum()), flush=True)
    bad_pts = np.unique(np.nonzero(d > 1e-12 * b.max())[2])
    print("  bad points:", bad_pts.size, bad_pts[:20])
    if bad_pts.size:
        for ij in bad_pts[:6]:
            m, k = np.unravel_index(np.argmax(d[:, :, ij]), d[:, :, ij].shape)
            print("   ij", ij, "row", int(g.row_of[ij]), "lat", float(np.atleast_1d(g.lat)[int(g.row_of[ij])] if np.size(g.lat) == g.ngy else np.atleast_1d(g.lat)[ij]), "depth", g.depth[ij], "worst m,k", m, k, "gpu", a[m, k, ij], "orc", b[m, k, ij],
                  "wind", f["WSWAVE"][ij], "ice", f["CICOVER"][ij])
        ms = np.unique(np.nonzero(d > 1e-12 * b.max())[0]); print("  bad m:", ms)
    for nm in ("UFRIC", "TAUW", "Z0M", "PHIAW", "TAUOC"):
        x, y = w.get_field(nm), o.get_field(nm)[w.own]
        print("  ", nm, np.abs(x - y).max() / max(np.abs(y).max(), 1e-300))
    mij = (w.get_field("mij") != o.get_field("MIJ")[w.own]).sum()
    print("   mij mismatches", int(mij))


print("cfl", o.propag(), w.propag())
report("after PROPAG")
o.implsch(); w.implsch()
report("after IMPLSCH")
o.step(); w.step()
report("after step 2")
