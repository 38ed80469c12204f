"""python -m ies_b200.compat.run <script.py> [args ...]  -- run a reference script unchanged on
the b200 engine (see ies_b200/compat/__init__.py)."""
import runpy
import sys

from . import install


def main():
    if len(sys.argv) < 2:
        print(__doc__)
        return 2
    install()
    script = sys.argv[1]
    sys.argv = sys.argv[1:]
    runpy.run_path(script, run_name='__main__')
    return 0


if __name__ == '__main__':
    sys.exit(main())
