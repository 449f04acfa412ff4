"""Import the UNMODIFIED reference (read-only at /root/reference) in this container.

Test/fixture tooling only: used by tools/make_golden.py to generate the committed golden vectors under
tests/golden/.  Nothing in the product path, the -m gpu tests, smoke() or bench.py imports this module, because
/root/reference does not exist on the GPU box.

The reference imports four viz-only packages at module scope (rerun, matplotlib, plotly, IPython) that are not
installed here; they are replaced by empty stub modules.
"""
import importlib.machinery
import sys
import types

REF_ROOT = "/root/reference"


def _stub(name, attrs=()):
    m = types.ModuleType(name)
    m.__spec__ = importlib.machinery.ModuleSpec(name, None)
    m.__path__ = []
    for a in attrs:
        setattr(m, a, type(a, (), {}))
    sys.modules[name] = m
    return m


def load_reference():
    """Returns a namespace with the reference's MSCKF, MSCKFParameters, IMUMeasurement, CameraMeasurement, ..."""
    if "rerun" not in sys.modules:
        _stub("rerun")
    for name in ("matplotlib", "matplotlib.pyplot", "plotly", "plotly.graph_objs", "IPython"):
        if name not in sys.modules:
            _stub(name)
    if "matplotlib.patches" not in sys.modules:
        _stub("matplotlib.patches", ["Ellipse"])
    if "IPython.display" not in sys.modules:
        d = _stub("IPython.display")
        d.display = d.clear_output = lambda *a, **k: None
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    from src.msckf.MSCKF import MSCKF, MSCKFParameters
    from src.msckf.IMU import IMUMeasurement
    from src.msckf.FeatureExtractor import CameraMeasurement, Feature
    from src.utils.geometry import Isometry3D
    ns = types.SimpleNamespace(MSCKF=MSCKF, MSCKFParameters=MSCKFParameters, IMUMeasurement=IMUMeasurement,
                               CameraMeasurement=CameraMeasurement, Feature=Feature, Isometry3D=Isometry3D)
    return ns
