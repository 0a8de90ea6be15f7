import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def orc():
    """The CPU oracle (test infrastructure) -- built on demand."""
    from oracle import oracle as o
    o.build()
    o.lib()
    return o


@pytest.fixture(scope="session")
def jtm_fix():
    return dict(np.load(os.path.join(GOLDEN, "jtm_fixture.npz")))


@pytest.fixture(scope="session")
def otm_fix():
    return dict(np.load(os.path.join(GOLDEN, "otm_fixture.npz")))


@pytest.fixture(scope="session")
def dr_fix():
    return dict(np.load(os.path.join(GOLDEN, "dr_fixture.npz")))


@pytest.fixture(scope="session")
def queries():
    return dict(np.load(os.path.join(GOLDEN, "queries.npz")))


@pytest.fixture(scope="session")
def golden_out():
    return dict(np.load(os.path.join(GOLDEN, "oracle_outputs.npz")))


@pytest.fixture(scope="session")
def jtm_oracle(orc, jtm_fix):
    f = jtm_fix
    tree = orc.Tree(int(f["max_level"]), f["codes"], f["node_ids"], f["is_leaf"], f["leaf_ids"], f["leaf_codes"])
    model = orc.TdmModel(f["params"], 8191, int(f["E"]), int(f["T"]))
    return tree, model


@pytest.fixture(scope="session")
def otm_oracle(orc, otm_fix):
    f = otm_fix
    model = orc.OtmModel(f["params"], 8191, int(f["E"]), int(f["T"]))
    n = len(f["items"])
    leaf_level = int(np.ceil(np.log(n) / np.log(2)))
    leaf_item = np.full(1 << leaf_level, -1, np.int32)
    leaf_item[f["leaf_ids"] - ((1 << leaf_level) - 1)] = f["items"]
    item_leaf = {int(a): int(b) for a, b in zip(f["items"], f["leaf_ids"])}
    return model, leaf_level, leaf_item, item_leaf


@pytest.fixture(scope="session")
def engine():
    """A live GPU engine; only -m gpu tests may request it."""
    from dismember_b200 import Engine
    e = Engine(0)
    yield e
    e.close()


def new_engine():
    from dismember_b200 import Engine
    return Engine(0)
