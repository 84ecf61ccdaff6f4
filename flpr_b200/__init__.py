"""Import shim: ``import flpr_b200`` resolves to the ``federated-lifelong-person-reid_b200/`` source tree.

The product directory carries the name mandated for this project (it contains hyphens, so it cannot be imported
directly); this package simply points its ``__path__`` there so that ``flpr_b200.ops``, ``flpr_b200.models``,
``flpr_b200.parallel`` ... are ordinary sub-packages.
"""
import os as _os

_ROOT = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                      "federated-lifelong-person-reid_b200")
__path__ = [_ROOT]
with open(_os.path.join(_ROOT, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_ROOT, "__init__.py"), "exec"))
