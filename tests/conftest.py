import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    """The CPU oracle (test infrastructure).  Built on demand from oracle/ludvm_oracle.c."""
    from oracle import ludvm_oracle
    ludvm_oracle.build()
    return ludvm_oracle


def load_golden(name):
    import json
    import numpy as np
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    d = {k: z[k] for k in z.files}
    if "kw" in d:
        d["kw"] = json.loads(str(d["kw"]))
    if "ff_kw" in d:
        d["ff_kw"] = json.loads(str(d["ff_kw"]))
    return d


def golden_tables(g):
    """Rebuild the step-table dict of a simulation fixture (bit-identical host inputs)."""
    import numpy as np
    tb = {}
    for k, v in g.items():
        if k.startswith("tb_"):
            v = np.asarray(v)
            tb[k[3:]] = v.item() if v.ndim == 0 else np.ascontiguousarray(v)
    return tb


def biteq(a, b):
    import numpy as np
    a = np.ascontiguousarray(np.asarray(a, dtype=np.float64))
    b = np.ascontiguousarray(np.asarray(b, dtype=np.float64))
    return a.shape == b.shape and bool(np.array_equal(a.view(np.uint64), b.view(np.uint64)))
