"""Drop-in registration: make the reference's scripts (`demo.py`, `train.py`, `test_*.py`) resolve
`from src.model import ViTEss` and `from lietorch import SE3` to this package, unchanged.

    import rel_pose_b200.dropin as d; d.install()          # then run / import the reference script
    python -m rel_pose_b200.run /path/to/rel_pose/demo.py --img1 a.png --img2 b.png --ckpt matterport.pth

The reference imports (`demo.py:5,18,22`, `train.py:10-11`, `test_matterport.py:13-17`):
    import lietorch / from lietorch import SE3      -> rel_pose_b200.lietorch
    from src.model import ViTEss                    -> rel_pose_b200.model.ViTEss
Everything else under the reference's `src/` package (data readers, logger, geom.losses) keeps resolving to
the reference's own files: only the `src.model` entry of sys.modules is replaced.
"""
import importlib
import sys
import types

from .lietorch import install_as_lietorch


def install(force=True):
    install_as_lietorch(force=force)
    from . import model as _model
    shim = types.ModuleType("src.model")
    shim.__doc__ = "rel_pose_b200 stand-in for the reference's src/model.py"
    shim.ViTEss = _model.ViTEss
    try:                                   # the reference tree is importable: keep its package, swap one module
        pkg = importlib.import_module("src")
    except Exception:
        pkg = types.ModuleType("src")
        pkg.__path__ = []
        sys.modules["src"] = pkg
    sys.modules["src.model"] = shim
    setattr(pkg, "model", shim)
    return shim
