"""Verbose GPU-vs-oracle comparison of the LLGCBZ0 / LLNORMAGAM instance of k_point (development aid; the pytest version is
tests/test_gpu_parity.py::test_gravity_capillary_physics_matches_oracle).  Prints per-field errors for every case and exits
non-zero if any of them is outside the tolerances of the parity tests."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np

from common import OUT_FIELDS, make_gpu, make_oracle, relerr

CASES = [("o48_cy49r1", {}), ("o48_iphys0_gc", {}), ("o640like", dict(llgcbz0=1, llnormagam=1, wspmin=0.3)),
         ("o48like", dict(llnormagam=1)), ("o48like", dict(llgcbz0=1, wspmin=0.3)), ("o48_iphys0", dict(llnormagam=1)),
         ("o48like", {})]


def main():
    bad = 0
    for case, extra in CASES:
        g, o, f, fl = make_oracle(case, **extra)
        _, s, w = make_gpu(case, **extra)
        print("==", case, extra, flush=True)
        for step in range(3):
            if step == 0:
                o.implsch(); w.implsch()
            else:
                assert o.step() == 0 and w.step() == 0
            w.synchronize()
            a, b = w.get_spec("fl1"), o.get_fl1()[:, :, w.own]
            big = b > 1e-8 * b.max()
            e1, e2 = relerr(a, b), float((np.abs(a - b)[big] / b[big]).max())
            nx = int((w.get_spec("xllws") != o.get_xllws()[:, :, w.own]).sum())
            nm_ = int((w.get_field("mij") != o.get_field("MIJ")[w.own]).sum())
            errs = {nm: relerr(w.get_field(nm), o.get_field(nm)[w.own]) for nm in OUT_FIELDS}
            worst = max(errs, key=errs.get)
            ok = np.isfinite(a).all() and e1 <= 1e-12 and e2 <= 1e-10 and nx == 0 and nm_ == 0 and errs[worst] <= 1e-10
            bad += 0 if ok else 1
            print(" step %d %s FL1 %.2e / %.2e  xllws %d mij %d  worst field %s %.2e" % (step, "ok " if ok else "BAD", e1, e2, nx, nm_,
                                                                                          worst, errs[worst]), flush=True)
            if not ok:
                for nm in OUT_FIELDS:
                    aa, bb = w.get_field(nm), o.get_field(nm)[w.own]
                    i = int(np.argmax(np.abs(aa - bb)))
                    print("    %-8s rel %.3e  at %d: gpu %.17g oracle %.17g  (n>1e-10: %d)" %
                          (nm, errs[nm], i, aa[i], bb[i], int((np.abs(aa - bb) > 1e-10 * np.abs(bb).max()).sum())))
        w.close()
    print("FAILED cases/steps:", bad)
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
