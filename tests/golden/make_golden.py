"""Generates tests/golden/*.npz from the CPU oracle (oracle/liboracle.so, -ffp-contract=off build).

The reference itself (Fortran + fiat + field_api + eccodes) cannot be built or imported in this image and ships no
per-routine vectors for this path, so these fixtures pin the ORACLE's own outputs: they guard against
regressions of the oracle and give the GPU tests a run-anywhere comparison.  Re-run:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))

from common import CASES, make_oracle, OUT_FIELDS, OUT_ITG, OUT_ICE, OUT_SEA  # noqa: E402

GOLDEN = {"g_iphys1": ("o48like", 8), "g_iphys0": ("o48_iphys0", 8), "g_a36": ("o640like", 6), "g_cy49r1": ("o48_cy49r1", 8)}


def main():
    only = sys.argv[1:]               # optional: the fixtures to (re)generate
    for name, (case, N) in GOLDEN.items():
        if only and name not in only:
            continue
        CASES["_tmp"] = dict(CASES[case], N=N)
        g, o, f, fl0 = make_oracle("_tmp")
        out = dict(case=case, N=N)      # the inputs are regenerated from ecwam_b200.synth (deterministic)
        nsteps = 2
        for _ in range(nsteps):
            assert o.step() == 0
        out["fl"] = o.get_fl1()
        out["xllws"] = np.packbits(o.get_xllws().astype(np.uint8).ravel())
        out["mij"] = o.get_field("MIJ").astype(np.int32)
        for nm in OUT_FIELDS:
            out[nm] = o.get_field(nm)
        hs, fm = o.hs_fm()
        out["hs"], out["fm"] = hs, fm
        out["bout"] = o.outbs(OUT_ITG, OUT_ICE, OUT_SEA)             # OUTBS columns in the order of common.OUT_ITG
        out["bout_itg"] = np.array(OUT_ITG, dtype=np.int32)
        out["wnorm"] = o.outwnorm(True)
        out["nsteps"] = nsteps
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
        print(name, case, "niblo", g.niblo, "Hs mean", hs.mean())


if __name__ == "__main__":
    main()
