"""TEST INFRASTRUCTURE ONLY -- never imported by the product path.

Loads the *unmodified* reference (usra-riacs/Nonlocal-Monte-Carlo, mounted read-only at
/root/reference) so that it can act as the live oracle in THIS container:

* every reference module imports matplotlib at top level (NMC/nmc.py:4, NPT/npt.py:6,
  NPT/apt_preprocessor.py:5, NPT/apt_ICM.py:6) and matplotlib is not installed here, so a
  stub is inserted into ``sys.modules`` first;
* NPT.run / APT_preprocessor.run use a ProcessPoolExecutor (NPT/npt.py:616,
  NPT/apt_preprocessor.py:160).  With ``num_cores=1`` there is one forked worker that inherits
  the parent's global ``np.random`` state at the first ``submit`` -- that is the only
  reproducible configuration of the reference and therefore the oracle configuration.

``/root/reference`` does not exist on the GPU box; anything that needs this module must be
skipped there (see ``available()``).  Golden vectors produced with it are committed under
``tests/golden`` by ``oracle/make_golden.py``.
"""
from __future__ import annotations

import contextlib
import importlib.util
import io
import os
import random
import sys
import tempfile
import types

import numpy as np

REFERENCE_ROOT = os.environ.get("NLMC_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "NMC", "nmc.py"))


class _Anything:
    """Object that tolerates every use the reference makes of matplotlib objects."""

    def __call__(self, *a, **k):
        return _Anything()

    def __getattr__(self, name):
        return _Anything()

    def __iter__(self):  # fig, ax = plt.subplots()
        return iter((_Anything(), _Anything()))

    def __getitem__(self, key):  # axes[0]
        return _Anything()

    def __len__(self):
        return 0

    def __add__(self, other):  # ax.get_xticklabels() + ax.get_yticklabels()
        return []

    def __radd__(self, other):
        return []


def _install_matplotlib_stub() -> None:
    if "matplotlib" in sys.modules and not isinstance(sys.modules["matplotlib"], types.ModuleType):
        return
    try:  # a real matplotlib is fine too
        import matplotlib  # noqa: F401
        import matplotlib.pyplot  # noqa: F401
        return
    except Exception:
        pass
    mpl = types.ModuleType("matplotlib")
    plt = types.ModuleType("matplotlib.pyplot")
    anything = _Anything()

    def _getattr(name):
        return anything

    plt.__getattr__ = _getattr  # type: ignore[attr-defined]
    mpl.__getattr__ = _getattr  # type: ignore[attr-defined]
    mpl.pyplot = plt  # type: ignore[attr-defined]
    sys.modules["matplotlib"] = mpl
    sys.modules["matplotlib.pyplot"] = plt


_CACHE: dict[str, types.ModuleType] = {}


def _load(name: str, rel: str) -> types.ModuleType:
    if name in _CACHE:
        return _CACHE[name]
    if not available():
        raise RuntimeError(f"reference not mounted at {REFERENCE_ROOT}")
    _install_matplotlib_stub()
    saved = np.random.get_state()  # NMC/nmc.py:10 calls np.random.seed(0) at import
    spec = importlib.util.spec_from_file_location(f"_nlmc_reference_{name}",
                                                  os.path.join(REFERENCE_ROOT, rel))
    mod = importlib.util.module_from_spec(spec)
    sys.modules[spec.name] = mod  # needed for pickling bound methods to pool workers
    spec.loader.exec_module(mod)
    np.random.set_state(saved)
    _CACHE[name] = mod
    return mod


def nmc():
    return _load("nmc", "NMC/nmc.py")


def npt():
    return _load("npt", "NPT/npt.py")


def apt_preprocessor():
    return _load("apt_preprocessor", "NPT/apt_preprocessor.py")


def apt_icm():
    return _load("apt_ICM", "NPT/apt_ICM.py")


def seed_all(seed: int) -> None:
    """Seed both generators the reference draws from (np.random legacy MT19937, `random`)."""
    np.random.seed(seed)
    random.seed(seed)


@contextlib.contextmanager
def quiet_tmp_cwd(silence: bool = True):
    """Run reference code in a scratch cwd (it writes PNG/NPY files) with stdout silenced."""
    old = os.getcwd()
    with tempfile.TemporaryDirectory() as tmp:
        os.chdir(tmp)
        try:
            if silence:
                with contextlib.redirect_stdout(io.StringIO()):
                    yield tmp
            else:
                yield tmp
        finally:
            os.chdir(old)
