"""`models` package as the reference's run scripts resolve it: ``eval('models.' + config.MODEL + '.get')`` and
``eval('models.' + config.BACKBONE + '.get')`` (run/validate.py:66,70).  The hot-path model and the PoseResNet backbone live
here; put this directory's parent (``lib/``) on ``sys.path`` *before* the reference's ``lib/`` and everything else
(``dataset``, ``utils``, ``core``) keeps resolving to the reference."""
from . import faster_voxelpose  # noqa: F401
from . import resnet  # noqa: F401
