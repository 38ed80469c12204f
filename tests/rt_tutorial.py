"""The canonical driver loop of the reference (tutorials/RT_simple_slabs.py:44-260, 408-433)
expressed against any implementation of the reference API: TF + IF + SF spaces, two eps_r=4
slabs, Gaussian plane source, three Sx collectors.  Used by the golden / collector tests."""
import numpy as np
from scipy.constants import c

um = 1e-6


def run(ns, engine, method, Nx, Ny, Nz, tsteps, dtype=np.complex64, cal_at=(), savedir='/tmp/ies_rt/',
        peak_pos=2000, probe=False):
    Lx, Ly, Lz = 720 * um, 512 * um, 512 * um
    dx, dy, dz = Lx / Nx, Ly / Ny, Lz / Nz
    dt = (1. / 4) * min(dx, dy, dz) / c
    mmtd = np.complex64 if np.dtype(dtype) in (np.dtype('complex64'), np.dtype('float32')) else np.complex128
    mk = lambda cls: cls((Nx, Ny, Nz), (dx, dy, dz), dt, tsteps, dtype, mmtd, method=method, engine=engine)
    TF, IF, SF = mk(ns.space.Basic3D), mk(ns.space.Basic3D), mk(ns.space.Empty3D)
    TF.malloc(); IF.malloc()
    pml = {'x': '+-', 'y': '', 'z': ''}
    for sp in (TF, IF):
        sp.apply_PML(pml, 10)
        sp.apply_BBC({'x': False, 'y': False, 'z': False})
        sp.apply_PBC({'x': False, 'y': True, 'z': True})
    wvc, spread = 100 * um, 0.08
    src = ns.source.Gaussian(dt, wvc, spread, peak_pos, dtype=np.float32)
    w0 = (2 * np.pi * c) / wvc
    w1, w2 = w0 * (1 - spread * 2), w0 * (1 + spread * 2)
    l1, l2 = 2 * np.pi * c / w1 / um, 2 * np.pi * c / w2 / um
    freqs = c / (np.arange(l2, l1, .2) * um)
    xsrt = Lx * 0.2
    sT = ns.source.Setter(TF, (xsrt, 0, 0), (xsrt, Ly, Lz), (0, 0, 0))
    sI = ns.source.Setter(IF, (xsrt, 0, 0), (xsrt, Ly, Lz), (0, 0, 0))
    t1 = 160 * um; t2 = t1 + 100 * um; t3 = t2 + 200 * um; t4 = t3 + 100 * um
    ns.structure.Box('dielectric_slab1', TF, (t1, 0, 0), (t2, Ly, Lz), 4, 1)
    ns.structure.Box('dielectric_slab2', TF, (t3, 0, 0), (t4, Ly, Lz), 4, 1)
    TF.init_update_constants(); IF.init_update_constants()
    col = {
        'TF_R': ns.collector.Sx("TF_R", savedir + "Sx/", TF, Lx * 0.85, (0, 0), (Ly, Lz), freqs, engine),
        'IF_R': ns.collector.Sx("IF_R", savedir + "Sx/", IF, Lx * 0.85, (0, 0), (Ly, Lz), freqs, engine),
        'SF_L': ns.collector.Sx("SF_L", savedir + "Sx/", SF, Lx * 0.15, (0, 0), (Ly, Lz), freqs, engine),
    }
    fap = ns.collector.FieldAtPoint("probe", savedir + "probe/", TF, (Lx * 0.5, Ly * 0.5, Lz * 0.5), engine) if probe else None
    out = {}
    pulse_fn = src.pulse_c if np.dtype(dtype).kind == 'c' else src.pulse_re
    for tstep in range(tsteps + 1):
        pulse = pulse_fn(tstep)
        sT.put_src('Ey', pulse, 'soft')
        sI.put_src('Ey', pulse, 'soft')
        TF.updateH(tstep); TF.updateE(tstep)
        IF.updateH(tstep); IF.updateE(tstep)
        SF.get_SF(TF, IF)
        for cobj in col.values():
            cobj.do_RFT(tstep)
        if fap is not None and tstep < tsteps:
            fap.get_time_signal(tstep)
        if tstep in cal_at:
            for name, cobj in col.items():
                cobj.get_Sx(tstep)
                out[(name, tstep)] = dict(area=np.array(cobj.Sx_area), Ey=np.array(cobj.DFT_Ey),
                                          Hz=np.array(cobj.DFT_Hz), Ez=np.array(cobj.DFT_Ez), Hy=np.array(cobj.DFT_Hy))
    return dict(out=out, freqs=freqs, TF=TF, IF=IF, SF=SF, probe=fap, dt=dt)
