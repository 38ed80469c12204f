"""Recipe for oracle/_ref/: the UNMODIFIED reference as it lies under /root/reference, placed
where it can travel to the GPU box -- TEST INFRASTRUCTURE (see oracle/ies_oracle.py header).

The reference is pure Python (no C/C++/CUDA sources), so "building" it is a copy
of the modules on its hot path plus the acceptance scripts:

    python -m oracle.make_ref          # /root/reference -> oracle/_ref/   (build container only)

oracle/_ref/ is git-ignored (no reference source enters the history) but not gpurun-ignored, so
`bench.py --impl reference`, bench.py's cpu_baseline leg and the script-acceptance test can run
the real reference on the GPU box's host cores (kind "reference") through oracle/ref_shims.py.
__graft_entry__.build() calls this when /root/reference is present.
"""
import os
import shutil
import sys

SRC = os.environ.get('IES_REFERENCE_SRC', '/root/reference')
DST = os.path.join(os.path.dirname(os.path.abspath(__file__)), '_ref')
FILES = ['space.py', 'source.py', 'collector.py', 'structure.py', 'plotter.py', 'recorder.py', 'analyzer.py',
         'tutorials/RT_simple_slabs.py', 'examples/mie/mie_scattering.py',
         'examples/reflectance_transmittance/RT_hole_slabs_short_input_src.py']


def main():
    if not os.path.isfile(os.path.join(SRC, 'space.py')):
        print(f'oracle/_ref: {SRC} is not present, nothing to do')
        return 0
    for rel in FILES:
        s, d = os.path.join(SRC, rel), os.path.join(DST, rel)
        if not os.path.isfile(s):
            continue
        os.makedirs(os.path.dirname(d), exist_ok=True)
        shutil.copyfile(s, d)
    print(f'oracle/_ref: {len(FILES)} reference files copied from {SRC}')
    return 0


if __name__ == '__main__':
    sys.exit(main())
