"""Verbose GPU-vs-oracle comparison used while developing (the pytest version lives in tests/test_gpu_parity.py)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import oracle as O
from ecwam_b200 import synth, model as M

def relerr(a, b, floor=1e-300):
    return np.abs(a - b).max() / max(np.abs(b).max(), floor)

def run(N, A, Fr, mk, iphys, nsteps, nproma=32, idelt=900.0, fused=False):
    g = synth.make_grid(N, mk)
    cfg = O.default_config(nang=A, nfre_red=Fr, nproma=nproma, npr=1, iphys=iphys, idelt=idelt, idelpro=idelt, delpro_lf=idelt)
    o = O.Oracle(cfg, g)
    s = M.WamSetup(g, nproc=1, nang=A, nfre_red=Fr, iphys=iphys, nproma=nproma, idelt=idelt, idelpro=idelt, delpro_lf=idelt)
    w = M.WamIntgr(s, 0)
    w.set_static(g.depth)
    f = synth.make_forcing(g)
    for k, v in f.items():
        o.set_field(k, v); w.set_field(k, v)
    fl = synth.jonswap_cold_start(f['WSWAVE'], f['WDWAVE'], A, 36, Fr)
    o.set_fl1(fl); w.set_fl1(fl)
    # static fields parity
    for nm in ('WAVNUM', 'CGROUP', 'CINV', 'XK2CG', 'STOKFAC'):
        a = w.get_field3(nm); b = o.get_field3(nm)[:, w.own]
        print('  static', nm, 'maxabs', np.abs(a - b).max())
    for nm in ('DEPTH', 'EMAXDPT', 'COSPHM1'):
        print('  static', nm, np.abs(w.get_field(nm) - o.get_field(nm)[w.own]).max())
    for step in range(nsteps):
        if fused:
            o.step(); w.step(); w.synchronize()
        else:
            cfl_o = o.propag(); cfl_g = w.propag(); w.synchronize()
            a = w.get_spec('fl1'); b = o.get_fl1()[:, :, w.own]
            nbad = (a != b).sum()
            print(f' step {step} propag: cfl {cfl_o}/{cfl_g} bitwise-different bins {nbad} of {a.size} rel {relerr(a,b):.3e}')
            o.implsch(); w.implsch(); w.synchronize()
        a = w.get_spec('fl1'); b = o.get_fl1()[:, :, w.own]
        big = b > 1e-8 * b.max()
        print(f' step {step} implsch: FL1 rel(max) {relerr(a,b):.3e} relbins {np.abs(a-b)[big].max() and (np.abs(a-b)[big]/b[big]).max():.3e} nan {np.isnan(a).sum()}')
        x = w.get_spec('xllws'); y = o.get_xllws()[:, :, w.own]
        print('   xllws mismatches', (x != y).sum())
        for nm in ('UFRIC', 'TAUW', 'TAUWDIR', 'Z0M', 'Z0B', 'CHRNCK', 'USTOKES', 'VSTOKES', 'TAUXD', 'TAUYD', 'TAUOCXD', 'TAUOCYD', 'TAUOC', 'PHIOCD', 'PHIEPS', 'PHIAW'):
            aa = w.get_field(nm); bb = o.get_field(nm)[w.own]
            print(f'   {nm:8s} rel {relerr(aa, bb):.3e}')
        mg = w.get_field('mij'); mo = o.get_field('MIJ')[w.own]
        print('   MIJ mismatches', (mg != mo).sum())
    hs_o, fm_o = o.hs_fm()
    hs_g, fm_g = M.hs_fm(s, w.get_spec('fl1'))
    print(f' Hs mean {hs_o.mean():.6f} rel {relerr(hs_g, hs_o[w.own]):.3e}  FM rel {relerr(fm_g, fm_o[w.own]):.3e}')

if __name__ == '__main__':
    import torch
    print(torch.cuda.get_device_name(0))
    run(48, 12, 25, 'continents', 1, 3)
    run(24, 36, 29, 'continents', 1, 2, nproma=24, idelt=450.0, fused=True)
    run(32, 24, 29, 'continents', 0, 2, nproma=64)
