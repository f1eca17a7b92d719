"""Import the UNMODIFIED reference modules from ``/root/reference`` (build container only).

TEST INFRASTRUCTURE -- see ``oracle/__init__.py``.  The reference is pure
Python but its third-party imports (matplotlib, asteroid, pesq, pystoi, librosa,
tensorboardX, S3PRL's ``utility`` / ``transformer`` / ``downstream``) are absent
here and there is no network; they are replaced by ``MagicMock`` entries in
``sys.modules`` so that the in-repo arithmetic (objective.py, evaluation.py,
utils.py, dataset.py, model.py, runner.py, sampler.py) imports and runs as
written.  Recipe: SURVEY.md Appendix A.  ``/root/reference`` does not exist on
the GPU box, so this is only used by ``oracle/make_golden.py`` and by CPU tests
that skip when the directory is missing.
"""
import importlib
import os
import sys
from unittest.mock import MagicMock

REFERENCE_ROOT = "/root/reference"

_STUBS = [
    "matplotlib", "matplotlib.pyplot", "asteroid", "asteroid.losses", "asteroid.losses.sdr",
    "asteroid.losses.stoi", "asteroid.losses.pmsqe", "pesq", "pystoi", "librosa", "librosa.util",
    "utility", "utility.preprocessor", "transformer", "transformer.nn_transformer",
    "transformer.model", "downstream", "downstream.solver", "downstream.model", "tensorboardX", "ipdb",
]


def available():
    return os.path.isdir(REFERENCE_ROOT)


def load(names=("utils", "objective", "evaluation", "dataset", "model", "sampler", "runner")):
    """Return {name: module} of reference modules, imported from REFERENCE_ROOT."""
    if not available():
        raise RuntimeError(f"{REFERENCE_ROOT} is not present (it exists only in the build container)")
    import numpy
    import scipy
    if not hasattr(scipy, "hanning"):                 # utils.py:13 calls the removed scipy.hanning
        scipy.hanning = numpy.hanning
    for name in _STUBS:
        if name not in sys.modules:
            try:
                importlib.import_module(name)
            except Exception:
                sys.modules[name] = MagicMock()
    # the reference's module names (utils, model, dataset, ...) are generic: import them in
    # isolation and remove them from sys.modules afterwards so they cannot shadow anything.
    saved = {n: sys.modules.pop(n) for n in list(names) if n in sys.modules}
    sys.path.insert(0, REFERENCE_ROOT)
    try:
        mods = {n: importlib.import_module(n) for n in names}
    finally:
        sys.path.remove(REFERENCE_ROOT)
        for n in names:
            sys.modules.pop(n, None)
        sys.modules.update(saved)
    return mods
