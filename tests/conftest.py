import json
import sys
from pathlib import Path

import numpy as np
import pytest

REPO = Path(__file__).resolve().parents[1]
GOLDEN = REPO / "tests" / "golden"
if str(REPO) not in sys.path:
    sys.path.insert(0, str(REPO))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run by the driver with -m gpu)")


@pytest.fixture(scope="session")
def meta():
    return json.loads((GOLDEN / "meta.json").read_text())


@pytest.fixture(scope="session")
def fixtures():
    """name -> dict(audio f32 [-1,1), sr, probs) from the reference's own clips (see oracle/gen_golden.py)."""
    out = {}
    for name in ("test16k", "aepyx16k", "aepyx8k"):
        z = np.load(GOLDEN / f"{name}.npz")
        out[name] = {"audio": z["pcm"].astype(np.float32) / 32768.0, "pcm": z["pcm"], "sr": int(z["sr"]), "probs": z["probs"]}
    return out


@pytest.fixture(scope="session")
def synthetic():
    return dict(np.load(GOLDEN / "synthetic.npz"))


@pytest.fixture(scope="session")
def sm_cases():
    return json.loads((GOLDEN / "state_machine_cases.json").read_text())


@pytest.fixture(scope="session")
def oracle():
    from oracle.oracle import Oracle
    return Oracle()
