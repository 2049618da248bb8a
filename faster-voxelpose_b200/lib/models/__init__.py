"""`models` package as the reference's run scripts resolve it: ``eval('models.' + config.MODEL + '.get')``
(run/validate.py:66).  Only the hot-path model lives here; put this directory's parent (``lib/``) on
``sys.path`` *before* the reference's ``lib/`` and everything else (``dataset``, ``utils``, ``core``) keeps
resolving to the reference."""
from . import faster_voxelpose  # noqa: F401
