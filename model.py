"""Reference-compatible module name: ``from model import ModelB_2`` and
``torch.load('modelB.pt', weights_only=False)`` (a pickle that names ``model.ModelB_2``,
``model.DoubleConvolution``, ``model.DownBlock_pool``, ``model.UpBlock``,
``model.ResidualConnection`` -- reference utils.py:802-826) resolve to the B200-native
implementation when the repository root is on ``sys.path``."""
import sifnn_b200 as _pkg
from sifnn_b200.model import *  # noqa: F401,F403
from sifnn_b200.model import __all__ as _names

for _n in _names:
    _obj = getattr(_pkg.model, _n)
    if isinstance(_obj, type):
        _obj.__module__ = "model"
