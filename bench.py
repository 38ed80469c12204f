#!/usr/bin/env python
"""bench.py -- headline benchmark of the b200 engine (BASELINE.json):
Mcell-updates/s of fp64 SHPF on 1024x256x256 per GPU, CPML in x, plane-wave
Gaussian source, two eps_r=4 slabs with an air cylinder through each
(examples/reflectance_transmittance/RT_hole_slabs_short_input_src.py geometry).

One step = Setter.put_src + Basic3D.updateH + Basic3D.updateE of ONE space through
the product's public Python API (ctypes -> C-ABI -> sm_100a kernels).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

N > 1 is launched by torchrun (one rank per GPU); each rank owns an x-slab of
1024x256x256 cells (weak scaling, global Nx = 1024*N) and exchanges one yz-plane
per half-step with its neighbours over NCCL.  `--impl reference` times the CPU
restatement of the reference's algorithm (oracle/, the reference itself is pure
Python and cannot travel to the GPU box) on the host cores.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

um = 1e-6
C0 = 299792458.0
NY = NZ = 256
NX_PER_GPU = 1024
BYTES_PER_CELL_UPDATE = 160.0      # fp64 real fields, both half-steps
METRIC = "Mcell-updates/s (fp64 SHPF 1024x256x256 per GPU, CPML-x, plane-wave source)"


def geometry(nx_global):
    dx, dy, dz = 720 * um / NX_PER_GPU, 512 * um / NY, 512 * um / NZ
    dt = 0.25 * min(dx, dy, dz) / C0
    return (dx, dy, dz), dt


def build_space(ns, nx_global, tsteps, comm=None, device=None, method='SHPF'):
    """Space + source + structures of the headline workload through the public API."""
    gap, dt = geometry(nx_global)
    kw = dict(method=method, engine='b200')
    if comm is not None: kw['comm'] = comm
    if device is not None: kw['device'] = device
    sp = ns.space.Basic3D((nx_global, NY, NZ), gap, dt, tsteps, np.float64, np.complex128, **kw)
    sp.malloc()
    sp.apply_PML({'x': '+-', 'y': '', 'z': ''}, 10)
    sp.apply_BBC({'x': False, 'y': False, 'z': False})
    sp.apply_PBC({'x': False, 'y': True, 'z': True})
    Ly, Lz = 512 * um, 512 * um
    xs = 0.2 * 720 * um
    setter = ns.source.Setter(sp, (xs, 0, 0), (xs, Ly, Lz), (0, 0, 0))
    for (a, b) in ((160 * um, 260 * um), (460 * um, 560 * um)):
        ns.structure.Box('slab', sp, (a, 0, 0), (b, Ly, Lz), 4., 1.)
        ns.structure.Cylinder3D('hole', sp, 'x', 128 * um, (a, b), (Ly / 2, Lz / 2), 1., 1.)
    src = ns.source.Gaussian(dt, 100 * um, 0.08, 2000, dtype=np.float64)
    return sp, setter, src


def pinned_like(shape, dtype):
    import torch
    t = torch.empty(int(np.prod(shape)) * np.dtype(dtype).itemsize, dtype=torch.uint8, pin_memory=True)
    return t.numpy().view(dtype).reshape(shape), t


class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu):
        self.gpu, self.p = gpu, None

    def start(self):
        try:
            self.p = subprocess.Popen(['nvidia-smi', '-i', str(self.gpu), f'--query-gpu={self.Q}',
                                       '--format=csv,noheader,nounits', '-lms', '100'],
                                      stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.p.terminate()
        try:
            out, _ = self.p.communicate(timeout=5)
        except Exception:
            self.p.kill(); out = ''
        sm, mx, reasons = [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(',')]
            if len(f) < 9: continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[5:9]):
                if v.lower().startswith('active'): reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def measured_peak():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        try:
            j = json.load(open(p))
            if 'hbm_gbs' in j:
                return float(j['hbm_gbs']), 'measured (MEASURED_PEAKS.json hbm_gbs)'
        except Exception:
            pass
    return 6650.0, 'fallback (B200_PROFILING.md 6.65 TB/s; MEASURED_PEAKS.json absent)'


def ncu_traffic(kernel='k_yline_update'):
    """DRAM bytes per launch of the dominant kernel from the committed ncu capture
    (profiles/ncu_summary.json, written by tools/ncu_summary.py), or None."""
    p = os.path.join(ROOT, 'profiles', 'ncu_summary.json')
    try:
        return json.load(open(p)).get(kernel, {}).get('dram_bytes_per_launch')
    except Exception:
        return None


# ------------------------------------------------------------------ CPU baseline
def oracle_sample_rate(nx_s, steps, warmup, seed=0):
    """Oracle (NumPy restatement of the reference) on an nx_s x 256 x 256 x-slab sample of
    the headline workload; returns (Mcell-updates/s, seconds per step)."""
    from oracle import ies_oracle as O
    gap, dt = geometry(nx_s)
    sp = O.OracleSpace((nx_s, NY, NZ), gap, dt, steps + warmup + 1, np.float64, np.complex128, method='SHPF')
    sp.apply_PML({'x': '+-', 'y': '', 'z': ''}, 10)
    sp.apply_PBC({'x': False, 'y': True, 'z': True})
    setter = O.OracleSetter(sp, (0.2 * nx_s * gap[0], 0, 0), (0.2 * nx_s * gap[0], 512 * um, 512 * um), (0, 0, 0))
    sp.eps[nx_s // 4: nx_s // 2] *= 4.
    sp.init_update_constants()
    rng = np.random.default_rng(seed)
    for n in ('Ex', 'Ey', 'Ez', 'Hx', 'Hy', 'Hz'):
        getattr(sp, n)[...] = rng.uniform(-1, 1, sp.loc_grid)
    t = 0
    for _ in range(warmup):
        setter.put_src('Ey', O.gaussian_pulse_re(t, dt, 100 * um, 0.08, 2000), 'soft')
        sp.update_h(t); sp.update_e(t); t += 1
    t0 = time.perf_counter()
    for _ in range(steps):
        setter.put_src('Ey', O.gaussian_pulse_re(t, dt, 100 * um, 0.08, 2000), 'soft')
        sp.update_h(t); sp.update_e(t); t += 1
    el = time.perf_counter() - t0
    return nx_s * NY * NZ * steps / el / 1e6, el / steps


def _ref_worker(args):
    nx_s, steps, warmup, seed = args
    os.environ['OMP_NUM_THREADS'] = '1'
    return oracle_sample_rate(nx_s, steps, warmup, seed)


def run_reference_arm(args, rank, world):
    if rank != 0:
        return
    import multiprocessing as mp
    try:
        cores = len(os.sched_getaffinity(0))
    except AttributeError:
        cores = os.cpu_count() or 1
    R = max(1, min(cores, 64, int(os.environ.get('IES_BENCH_REF_MAX_RANKS', 64))))
    nx_s = 16
    # keep the whole run within a few minutes: ~0.35 s per rank-step at 16x256x256
    steps = max(1, min(args.steps, 60))
    warm = max(1, min(args.warmup, 3))
    ctx = mp.get_context('fork')
    t0 = time.perf_counter()
    with ctx.Pool(R) as pool:
        res = pool.map(_ref_worker, [(nx_s, steps, warm, s) for s in range(R)])
    wall = time.perf_counter() - t0
    agg = float(sum(r[0] for r in res))
    spp = float(max(r[1] for r in res))
    sample = (f"{R} independent x-slab ranks of {nx_s}x{NY}x{NZ} cells (one per usable host core), "
              f"{steps} timed + {warm} warm-up steps each, NumPy/pocketfft single-threaded per rank; "
              f"halo exchange omitted (2 planes per half-step, <0.1% of rank time); wall {wall:.1f} s")
    line = {
        "impl": "reference", "metric": METRIC, "value": agg, "unit": "Mcell-updates/s",
        "n_gpus": args.gpus, "steps": steps, "warmup": warm, "ms_per_step": spp * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": "SHPF fp64 1024x256x256-shaped x-slab samples on host cores (oracle port of the "
                               "reference's updateH/updateE; the reference is pure Python and is not on this box)",
                   "steps_requested": args.steps},
        "cpu_baseline": {"value": agg, "unit": "Mcell-updates/s", "cores": R, "kind": "port", "sample": sample},
        "e2e": {"value": agg, "unit": "Mcell-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------ GPU arm
def run_b200(args, rank, world, local_rank):
    import types
    import torch
    import ies_b200
    from ies_b200 import _lib, comm as icomm
    ns = types.SimpleNamespace(space=ies_b200.space, source=ies_b200.source,
                               structure=ies_b200.structure, collector=ies_b200.collector)
    lib = _lib.load()
    dist = None
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
        comm = icomm.TorchComm()
    else:
        comm = icomm.SingleComm()
    K, W = args.steps, args.warmup
    nx_global = NX_PER_GPU * world
    import contextlib
    with contextlib.redirect_stdout(sys.stderr):      # the API mirrors the reference's set-up prints
        sp, setter, src = build_space(ns, nx_global, K + W + 8, comm=comm, device=local_rank)
    ncell_local = sp.myNx * NY * NZ

    # host inputs (pinned): random fields and the two coefficient arrays
    rng = np.random.default_rng(1234 + rank)
    host = {}
    keep = []
    for n in ('Ex', 'Ey', 'Ez', 'Hx', 'Hy', 'Hz'):
        a, t = pinned_like(sp.loc_grid, np.float64)
        a[...] = rng.uniform(-1, 1, sp.loc_grid)
        host[n] = a; keep.append(t)
    out_host, t = pinned_like(sp.loc_grid, np.float64); keep.append(t)

    sp.init_update_constants()                         # eps/mu -> CH2/CE2 on the host (set-up, untimed)
    coef = {}
    for half, arr in ((_lib.HALF_H, sp.CHx2), (_lib.HALF_E, sp.CEx2)):
        a, t = pinned_like(sp.loc_grid, np.float64)
        a[...] = arr
        coef[half] = a; keep.append(t)

    def upload_all():
        """H2D of the two coefficient arrays and the six fields, straight through the C-ABI."""
        for half, a in coef.items():
            _lib.check(lib.ies_set_coeff(sp._ctx, half, _lib.dptr(a), a.size))
        for n in host:
            _lib.check(lib.ies_set_field(sp._ctx, _lib.COMP[n], _lib.I3(0, 0, 0), _lib.I3(*sp.loc_grid),
                                         host[n].ctypes.data_as(C.c_void_p)))

    def step(t):
        setter.put_src('Ey', src.pulse_re(t), 'soft')
        sp.updateH(t)
        sp.updateE(t)

    def barrier():
        if dist is not None: dist.barrier()
        sp.sync()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if dist is None: return x
        t = torch.tensor([x], dtype=torch.float64, device='cuda')
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    if world > 1:
        # engine and NCCL share comm's dedicated stream from the first step on
        sp._use_stream(comm._stream_for(local_rank).cuda_stream)

    # ---------------- device-resident throughput (value) ----------------
    upload_all()
    t = 0
    for _ in range(W):
        step(t); t += 1
    barrier()
    clocks = ClockSampler(local_rank)
    if rank == 0: clocks.start()
    l0 = lib.ies_launch_count()
    _lib.check(lib.ies_profile(sp._ctx, 1))
    _lib.check(lib.ies_timer_start(sp._ctx))
    for _ in range(K):
        step(t); t += 1
    ms = C.c_double()
    _lib.check(lib.ies_timer_stop(sp._ctx, C.byref(ms)))
    barrier()
    launches = lib.ies_launch_count() - l0
    clk = clocks.stop() if rank == 0 else None
    prof = {}
    for slot, nm in ((0, 'k_zline'), (1, 'k_yline_update')):
        tot, cnt = C.c_double(), C.c_int64()
        _lib.check(lib.ies_profile_read(sp._ctx, slot, C.byref(tot), C.byref(cnt)))
        prof[nm] = (tot.value, cnt.value)
    _lib.check(lib.ies_profile(sp._ctx, 0))
    ms_total = max_over_ranks(ms.value)
    ms_step = ms_total / K
    value = ncell_local * world * K / (ms_total * 1e-3) / 1e6

    # sanity: the fields are still finite after the timed steps
    probe = np.asarray(sp.Ey[sp.myNx // 2, :4, :4])
    assert np.all(np.isfinite(probe)), "non-finite field after the timed region"

    # ---------------- end to end from host buffers (e2e) ----------------
    # whole job through the public API: H2D of the six fields + two coefficient arrays from
    # pinned host memory, K steps, D2H of the six result fields -- all inside the timed region.
    Ke = K
    barrier()
    t0 = time.perf_counter()
    _lib.check(lib.ies_timer_start(sp._ctx))
    upload_all()
    te = 0
    for _ in range(Ke):
        step(te); te += 1
    d2h = 0
    for n in ('Ex', 'Ey', 'Ez', 'Hx', 'Hy', 'Hz'):
        _lib.check(lib.ies_get_field(sp._ctx, _lib.COMP[n], _lib.I3(0, 0, 0), _lib.I3(*sp.loc_grid),
                                     out_host.ctypes.data_as(C.c_void_p)))
        d2h += out_host.nbytes
    ms2 = C.c_double()
    _lib.check(lib.ies_timer_stop(sp._ctx, C.byref(ms2)))
    barrier()
    wall_e2e = (time.perf_counter() - t0) * 1e3
    e2e_ms = max_over_ranks(max(ms2.value, wall_e2e))
    h2d = 8 * ncell_local * 8
    e2e_val = ncell_local * world * Ke / (e2e_ms * 1e-3) / 1e6

    if rank != 0:
        if dist is not None: dist.destroy_process_group()
        return

    peak, peak_src = measured_peak()
    ky_ms, ky_n = prof['k_yline_update']
    kz_ms, kz_n = prof['k_zline']
    ky_avg = ky_ms / max(ky_n, 1)
    kz_avg = kz_ms / max(kz_n, 1)
    # Algorithmic bytes of one half-step of the slab: 80 B per cell = read 6 fields +
    # 1 coefficient array, write 3 fields.  The half-step is two launches (k_zline derivative
    # pass + k_yline_update); the dominant kernel k_yline_update carries all of these bytes
    # (k_zline only produces scratch), so achieved = 80 B x cells / its launch duration.
    alg_bytes_launch = 0.5 * BYTES_PER_CELL_UPDATE * ncell_local
    achieved = alg_bytes_launch / (ky_avg * 1e-3) / 1e9 if ky_avg > 0 else 0.0
    step_gbs = BYTES_PER_CELL_UPDATE * ncell_local / (ms_step * 1e-3) / 1e9
    traffic = ncu_traffic('k_yline_update')
    roofline = {
        "bound": "hbm", "kernel": "k_yline_update<double,false,256> (y-line FFT derivative + fused update/CPML)",
        "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
        "traffic": traffic, "peak_source": peak_src,
        "algorithmic_bytes_per_launch": alg_bytes_launch,
        "avg_launch_ms": ky_avg, "launches_timed": ky_n,
        "kernel_share_of_step": (ky_ms / K) / ms_step if ms_step > 0 else None,
        "dram_gbs_from_ncu_traffic": (traffic / (ky_avg * 1e-3) / 1e9) if (traffic and ky_avg > 0) else None,
        "zline": {"kernel": "k_zline<double,false,256,16> (z-line FFT derivative pass -> scratch)",
                  "avg_launch_ms": kz_avg, "launches_timed": kz_n,
                  "share_of_step": (kz_ms / K) / ms_step if ms_step > 0 else None,
                  "traffic": ncu_traffic('k_zline')},
        "step": {"achieved": step_gbs, "frac": step_gbs / peak,
                 "note": "whole leap-frog step (both kernels, both half-steps, source injection) at 160 B per "
                         "cell-update: the figure comparable with the 60 % target"},
    }

    cpu = None
    if world == 1:
        nx_s = 64
        v, spp = oracle_sample_rate(nx_s, 2, 1)
        cpu = {"value": v, "unit": "Mcell-updates/s", "cores": 1, "kind": "port",
               "sample": f"oracle (NumPy restatement of the reference's updateH/updateE) on a {nx_s}x{NY}x{NZ} "
                         f"x-slab of the same workload, 1 warm-up + 2 timed steps, one core ({spp:.2f} s/step)"}

    line = {
        "metric": METRIC, "value": value, "unit": "Mcell-updates/s", "n_gpus": world, "steps": K,
        "warmup": W, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"SHPF fp64 {NX_PER_GPU}x{NY}x{NZ} cells per GPU (global {nx_global}x{NY}x{NZ}), "
                               "CPML x+- (npml 10), periodic y/z, 2 eps_r=4 slabs with air cylinder, "
                               "Gaussian plane source Ey (soft); fields pre-filled uniform(-1,1)",
                   "parallelism": f"x-slab x{world}" if world > 1 else "single slab",
                   "l2": "working set 4.8 GB per GPU >> 126 MB L2 (no flush needed)",
                   "timing": "CUDA events on the engine stream, max over ranks"},
        "clocks": clk,
        "e2e": {"value": e2e_val, "unit": "Mcell-updates/s", "h2d_bytes_per_step": h2d / Ke,
                "d2h_bytes_per_step": d2h / Ke, "steps": Ke,
                "note": "H2D of 6 fields + 2 coefficient arrays (pinned), K steps, D2H of 6 fields, all timed"},
        "gpu_launches": int(launches),
        "roofline": roofline,
    }
    if cpu is not None:
        line["cpu_baseline"] = cpu
    print(json.dumps(line), flush=True)
    if dist is not None: dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=100)      # >= 20 warm-up + >= 100 timed steps
    ap.add_argument('--warmup', type=int, default=20)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    rank = int(os.environ.get('RANK', 0))
    world = int(os.environ.get('WORLD_SIZE', 1))
    local_rank = int(os.environ.get('LOCAL_RANK', 0))
    if args.impl == 'reference':
        run_reference_arm(args, rank, world)
        return
    if world != args.gpus and world == 1 and args.gpus > 1:
        # convenience: re-launch under torchrun
        cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', f'--nproc-per-node={args.gpus}',
               '--master-addr', '127.0.0.1', '--master-port', '29531', os.path.abspath(__file__),
               '--gpus', str(args.gpus), '--steps', str(args.steps), '--warmup', str(args.warmup)]
        sys.exit(subprocess.call(cmd))
    run_b200(args, rank, world, local_rank)


if __name__ == '__main__':
    main()
