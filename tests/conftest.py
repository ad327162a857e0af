import importlib.util
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def load_package():
    """The package directory is named after the reference repo (hyphenated), so
    it is imported through importlib under the module name `klu_b200`."""
    if "klu_b200" in sys.modules:
        return sys.modules["klu_b200"]
    pkg = os.path.join(ROOT, "kaldi-lattice-utils_b200")
    spec = importlib.util.spec_from_file_location("klu_b200", os.path.join(pkg, "__init__.py"),
                                                  submodule_search_locations=[pkg])
    mod = importlib.util.module_from_spec(spec)
    sys.modules["klu_b200"] = mod
    spec.loader.exec_module(mod)
    return mod


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_sessionstart(session):
    """Built artefacts are not in the history: a fresh checkout builds them once (nvcc
    cross-compiles without a GPU), as `__graft_entry__.build()` does."""
    pkg = os.path.join(ROOT, "kaldi-lattice-utils_b200")
    need = [os.path.join(pkg, "libklu_b200.so"), os.path.join(pkg, "libklu_host.so"),
            os.path.join(pkg, "bin", "klu-copy-lattices"), os.path.join(pkg, "bin", "lattice-to-word-frame-post")]
    if not all(os.path.exists(p) for p in need):
        import __graft_entry__
        __graft_entry__.build()


@pytest.fixture(scope="session")
def klu():
    return load_package()


@pytest.fixture(scope="session")
def ora():
    from oracle import ora as o
    o.build()
    return o


@pytest.fixture(scope="session")
def engine(klu):
    eng = klu.Engine(0)
    yield eng
    eng.close()
