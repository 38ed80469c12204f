"""ies_b200 -- B200-native time-stepping engine behind the IES Python API.

Drop-in: put this directory on sys.path and `import space, source, structure,
collector, plotter, recorder` exactly as the reference's scripts do, or
`from ies_b200 import space, ...`.  See DESIGN.md / INTEGRATION.md.
"""
from . import _lib, comm, space, source, structure, collector, plotter, recorder  # noqa: F401

__all__ = ['space', 'source', 'structure', 'collector', 'plotter', 'recorder', 'comm']
