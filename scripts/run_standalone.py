#!/usr/bin/env python
"""A stand-alone forecast with the B200 hot path, laid out like share/ecwam/scripts/ecwam_run_model.sh + WAMODEL's
ADVECTION loop (wamodel.F90:228-400): cold start (synthetic JONSWAP, GRIB forcing is not available offline), then per
propagation step  PROPAG_WAM -> NEWWIND when a new wind is due -> IMPLSCH  (model.WamIntgr.advection_step), and at every
output step OUTBS + OUTWNORM, printing the WAMNORM lines the reference writes to statistics.log (outwnorm.F90:140-152).

  python scripts/run_standalone.py --grid O48 --hours 6                 # one GPU
  torchrun --nproc-per-node 2 scripts/run_standalone.py --grid O320     # MPDECOMP over 2 GPUs, NCCL halo

With --check the CPU oracle runs the same sequence and the norms are compared (small grids only).
"""
from __future__ import annotations

import argparse
import ctypes as C
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np
import torch

from ecwam_b200 import lib as L, model as M, synth

# the fields tests/etopo1_oper_an_fc_O48.yml asks for: swh mwd mwp pp1d dwi cdww wind -> parameter numbers of mpcrtbl.F90
FIELDS = [("swh", 1), ("mwd", 2), ("mwp", 3), ("pp1d", 6), ("dwi", 5), ("cdww", 7), ("wind", 10)]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--grid", default="O48", choices=sorted(synth.CONFIGS))
    ap.add_argument("--hours", type=float, default=6.0)
    ap.add_argument("--output-every", type=float, default=1.0, help="hours between OUTBS/OUTWNORM (yml output.fields.at.timestep)")
    ap.add_argument("--wind-every", type=float, default=1.0, help="hours between forcing updates (IDELWO)")
    ap.add_argument("--iphys", type=int, default=1)
    ap.add_argument("--mask", default="continents", choices=["continents", "aqua"])
    ap.add_argument("--cycle", default="default", choices=["default", "cy49r1"],
                    help="cy49r1: LLGCBZ0 + LLNORMAGAM, WSPMIN = 0.3 (tests/etopo1_oper_an_fc_O48_cy49r1.yml)")
    ap.add_argument("--grid-tables", default="", help="a wam_grid_tables file (reference binary format) to take the grid and the "
                    "bathymetry from; --grid then only selects the spectral resolution and the time steps")
    ap.add_argument("--restart-in", default="", help="directory with BLS/LAW restart files (reference format) to start from")
    ap.add_argument("--restart-out", default="", help="directory to write the BLS/LAW restart files of the final state to")
    ap.add_argument("--start", default="20220101000000", help="CDATEF, YYYYMMDDHHmmss")
    ap.add_argument("--check", action="store_true", help="run the CPU oracle alongside and compare the norms (test infrastructure)")
    args = ap.parse_args()

    world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("no CUDA device: the WAMINTGR hot path has no CPU fallback")
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    comm = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
        lib = L.load()
        buf = C.create_string_buffer(128)
        if rank == 0:
            L.check(lib.ecwam_b200_nccl_unique_id(buf), "nccl_unique_id")
        t = torch.frombuffer(bytearray(buf.raw), dtype=torch.uint8).clone().to(dev)
        dist.broadcast(t, 0)
        cm = C.c_void_p()
        L.check(lib.ecwam_b200_nccl_comm_init(bytes(t.cpu().numpy().tobytes()), world, rank, C.byref(cm)), "nccl_comm_init")
        comm = cm.value

    cfg = synth.CONFIGS[args.grid]
    nproma = {"O48": 32, "O320": 64}.get(args.grid, 32)
    g = synth.grid_from_tables(args.grid_tables) if args.grid_tables else synth.make_grid(cfg["N"], args.mask)
    kw = dict(nang=cfg["nang"], nfre_red=cfg["nfre_red"], iphys=args.iphys, nproma=nproma, idelt=cfg["idelt"], idelpro=cfg["idelpro"],
              delpro_lf=cfg["delpro_lf"], ifrelfmax=cfg["ifrelfmax"])
    if args.cycle == "cy49r1":
        kw.update(llgcbz0=1, llnormagam=1, wspmin=0.3)
    s = M.WamSetup(g, nproc=world, **kw)
    w = M.WamIntgr(s, rank, device=dev, nccl_comm=comm)
    w.set_static(g.depth)
    f0 = synth.make_forcing(g, t_hours=0.0)
    for k, v in f0.items():
        w.set_field(k, v)
    fl = synth.jonswap_cold_start(f0["WSWAVE"], f0["WDWAVE"], cfg["nang"], 36, cfg["nfre_red"])
    w.set_fl1(fl)

    def restart_names(directory, t_sec):       # GRSTNAME: <ID><CDATEF>_<forecast range> (grstname.F90:88-142)
        import datetime
        cdt = (datetime.datetime.strptime(args.start, "%Y%m%d%H%M%S") + datetime.timedelta(seconds=t_sec)).strftime("%Y%m%d%H%M%S")
        out = []
        for fid in (b"BLS", b"LAW"):
            buf = C.create_string_buffer(400)
            L.check(w.lib.ecwam_b200_grstname(cdt.encode(), args.start.encode(), 0, fid, directory.encode(), buf, 400), "grstname")
            out.append(buf.value.decode())
        return cdt, out[0], out[1]

    def barrier():
        if world > 1:
            torch.distributed.barrier()

    t_start = 0
    if args.restart_in:
        found = sorted(n for n in os.listdir(args.restart_in) if n.startswith("BLS") and "." not in n)
        if not found:
            raise SystemExit("no BLS restart file in " + args.restart_in)
        bls = os.path.join(args.restart_in, found[-1])
        law = os.path.join(args.restart_in, "LAW" + found[-1][3:])
        w.getspec(bls)
        cdt, cdatewo = w.getstress(law)[:2]
        rng = found[-1].split("_")[1]
        t_start = int(rng[:6]) * 86400 + int(rng[6:8]) * 3600 + int(rng[8:10]) * 60 + int(rng[10:12])
        if rank == 0:
            print("restart from %s (CDTPRO %s, +%d s)" % (bls, cdt, t_start))

    def ff_next(t_sec):          # what GETWND would deliver for the wind step starting at t_sec
        f = synth.make_forcing(g, t_hours=(t_start + t_sec) / 3600.0)
        f.update(USTRA=np.zeros(g.niblo), VSTRA=np.zeros(g.niblo))
        return f

    o = None
    if args.check:
        from oracle import oracle as O
        o = O.Oracle(O.default_config(npr=1, **kw), g)
        for k, v in f0.items():
            o.set_field(k, v)
        o.set_fl1(fl)

    itg = [i for _, i in FIELDS]
    ice = [M.OUTBLOCK_PARAMS[i][0] for i in itg]
    sea = [M.OUTBLOCK_PARAMS[i][1] for i in itg]
    clk = M.WamClock(idelpro=cfg["idelpro"], idelt=cfg["idelt"], idelwo=int(args.wind_every * 3600))
    if args.restart_in:          # CDATEWO of the LAW file: when the next forcing fields are due (savstress.F90:157-159)
        import datetime
        fmt = "%Y%m%d%H%M%S"
        clk.cdatewh = clk.cdatewo = int((datetime.datetime.strptime(cdatewo, fmt) - datetime.datetime.strptime(cdt, fmt)).total_seconds())
    nadv = int(round(args.hours * 3600 / cfg["idelpro"]))
    out_every = int(round(args.output_every * 3600))
    if rank == 0:
        print("%s: %d sea points, %dx36(%d) spectrum, IPHYS=%d, %d rank(s), %d propagation steps of %g s" %
              (args.grid, g.niblo, cfg["nang"], cfg["nfre_red"], args.iphys, world, nadv, cfg["idelpro"]))
    t0 = time.perf_counter()
    worst = 0.0
    for kadv in range(nadv):
        imp0 = clk.cdtimp
        cfl = w.advection_step(clk, ff_next)
        if cfl:
            raise SystemExit("CFL violated at %d points" % cfl)
        if o is not None:           # the same call sequence on the CPU oracle
            assert o.propag() == 0
            t_imp = imp0
            while t_imp < clk.cdtpro:
                if t_imp >= clk.idelwo and t_imp % clk.idelwo == 0:
                    nf = ff_next(t_imp)
                    o.newwind(nf)
                o.implsch()
                t_imp += clk.idelt
        if clk.cdtpro % out_every == 0:
            w.outbs(itg, ice, sea)
            wn = w.outwnorm(True)
            if rank == 0:
                print("  WAMNORM ON +%05.1f h" % ((t_start + clk.cdtpro) / 3600.0))
                for (name, _), row in zip(FIELDS, wn):
                    print("    %-5s avg %.14e  min %.14e  max %.14e  n %d" % (name, row[0], row[1], row[2], int(row[3])))
            if o is not None:
                o.outbs(itg, ice, sea)
                wo = o.outwnorm(True)
                err = np.abs(wn[[0, 2, 3, 5, 6], :3] - wo[[0, 2, 3, 5, 6], :3]) / np.maximum(np.abs(wo[[0, 2, 3, 5, 6], :3]), 1e-12)
                worst = max(worst, float(err.max()))
    w.synchronize()
    if rank == 0:
        dt = time.perf_counter() - t0
        print("done: %.2f s wall (%s), %.3g grid-point spectra/s incl. output steps" %
              (dt, "with the CPU oracle alongside" if o is not None else "GPU only", g.niblo * nadv / dt))
        if o is not None:
            print("max relative difference of the swh/mwp/pp1d/cdww/wind norms vs the CPU oracle: %.2e" % worst)
            assert worst < 1e-10
    if args.restart_out:
        os.makedirs(args.restart_out, exist_ok=True)
        cdt, bls, law = restart_names(args.restart_out, t_start + clk.cdtpro)
        cdatewo = restart_names(args.restart_out, t_start + clk.cdatewo)[0]
        if rank == 0:                        # rank 0 lays the files out, then every rank writes its own points in place
            w.savspec(bls, create=True); w.savstress(law, cdt, cdatewo, create=True)
        barrier()
        if rank != 0:
            w.savspec(bls, create=False); w.savstress(law, cdt, cdatewo, create=False)
        barrier()
        if rank == 0:
            print("restart files: %s (%d bytes), %s" % (bls, os.path.getsize(bls), law))
    w.close()
    if world > 1:
        L.check(L.load().ecwam_b200_nccl_comm_destroy(C.c_void_p(comm)), "comm_destroy")
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
