import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


@pytest.fixture(scope="session")
def s288c():
    """The reference's tests/S288c fixtures, bundled by tests/golden/make_golden.py."""
    z = np.load(ROOT / "tests" / "golden" / "s288c_fixture.npz")
    return {k: z[k] for k in z.files}


@pytest.fixture(scope="session")
def orc():
    from oracle import oracle
    oracle.lib()
    return oracle


def cut_ctgs(orc, s288c, piece):
    """`gams gen --piece N` on the fixture genome → list of (ctg_id, chr_id, chr_start, chr_end, seq)."""
    out = []
    for name in s288c["chr_names"].tolist():
        seq = s288c["seq_" + name]
        for i, (s, e) in enumerate(orc.gen_ctgs(seq, piece=piece), 1):
            out.append((f"ctg:{name}:{i}", name, s, e, seq[s - 1:e]))
    return out
