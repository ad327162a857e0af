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
