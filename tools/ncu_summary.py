#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, no GPU needed) into the JSON kept under profiles/.

    python tools/ncu_summary.py gpurun_out/prof_fused.ncu-rep gpurun_out/prof_mie_y.ncu-rep:_mie ... > profiles/ncu_summary.json
"""
import csv
import io
import json
import subprocess
import sys

KEYS = {
    'gpu__time_duration.sum': 'duration',
    'dram__bytes_read.sum': 'dram_read',
    'dram__bytes_write.sum': 'dram_write',
    'dram__bytes_read.sum.per_second': 'dram_read_per_s',
    'dram__bytes_write.sum.per_second': 'dram_write_per_s',
    'launch__registers_per_thread': 'registers_per_thread',
    'launch__grid_size': 'grid',
    'launch__block_size': 'block',
    'launch__shared_mem_per_block_dynamic': 'dyn_smem_per_block',
    'launch__occupancy_limit_registers': 'occ_limit_regs_blocks',
    'launch__occupancy_limit_shared_mem': 'occ_limit_smem_blocks',
    'sm__warps_active.avg.pct_of_peak_sustained_active': 'warps_active_pct',
    'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active': 'fp64_pipe_pct',
    'sm__throughput.avg.pct_of_peak_sustained_elapsed': 'sm_throughput_pct',
    'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed': 'dram_throughput_pct',
    'lts__t_sector_hit_rate.pct': 'l2_hit_pct',
    'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum': 'smem_bank_conflicts',
    'sass__inst_executed_local_loads': 'local_load_insts',
    'sass__inst_executed_local_stores': 'local_store_insts',
    'smsp__inst_executed.sum': 'warp_insts',
}
STALLS = 'smsp__pcsamp_warps_issue_stalled_'


def to_bytes(v, unit):
    mult = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'Tbyte': 1e12}
    for k, m in mult.items():
        if unit.startswith(k):
            return float(v) * m
    return float(v)


def summarise(path):
    out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    res = []
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        u = dict(zip(hdr, units))
        e = {'kernel': d.get('Kernel Name', '')[:120]}
        for k, name in KEYS.items():
            if k in d and d[k] not in ('', 'n/a'):
                try:
                    val = float(d[k].replace(',', ''))
                except ValueError:
                    continue
                if name in ('dram_read', 'dram_write'):
                    val = to_bytes(val, u[k])
                    name += '_bytes'
                elif name == 'duration':
                    val = val * {'ns': 1e-3, 'us': 1.0, 'ms': 1e3, 's': 1e6}.get(u[k].replace('second', 's').replace('usecond', 'us'), 1.0)
                    if u[k] in ('usecond', 'us'): pass
                    name = 'duration_us'
                    if u[k] in ('msecond', 'ms'): val = float(d[k]) * 1e3
                    if u[k] in ('nsecond', 'ns'): val = float(d[k]) * 1e-3
                    if u[k] in ('usecond', 'us'): val = float(d[k])
                    if u[k] in ('second', 's'): val = float(d[k]) * 1e6
                elif name.endswith('_per_s'):
                    val = to_bytes(val, u[k].split('/')[0])
                e[name] = val
        st = {h[len(STALLS):]: float(d[h]) for h in hdr if h.startswith(STALLS) and not h.endswith('_not_issued') and d[h] not in ('', 'n/a')}
        tot = sum(st.values()) or 1.0
        e['stall_pct'] = {k: round(100 * v / tot, 1) for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:8]}
        if 'dram_read_bytes' in e and 'dram_write_bytes' in e:
            e['dram_bytes_per_launch'] = e['dram_read_bytes'] + e['dram_write_bytes']
        res.append(e)
    return res


if __name__ == '__main__':
    # arguments: report.ncu-rep[:suffix]  -- the suffix names the workload ("_mie", "_c128" ...)
    allk = {}
    for arg in sys.argv[1:]:
        p, _, suffix = arg.partition(':')
        for e in summarise(p):
            key = 'k_yline_update' if 'k_yline_update' in e['kernel'] else 'k_zline' if 'k_zline' in e['kernel'] \
                else 'k_shpf_fused' if 'fused' in e['kernel'] else 'k_fdtd' if 'k_fdtd' in e['kernel'] \
                else 'k_xline' if 'k_xline' in e['kernel'] else e['kernel'][:40]
            key += suffix
            allk.setdefault(key, e)
            allk[key]['source_report'] = p
    json.dump(allk, sys.stdout, indent=1, sort_keys=True)
    print()
