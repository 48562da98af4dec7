import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def golden():
    def load(name):
        return dict(np.load(os.path.join(GOLDEN, name)))
    return load


@pytest.fixture(scope="session")
def reference_plonk_key():
    return _reference_plonk_key


def _reference_plonk_key(g, tag):
    """Rebuilds a PLONK zkey the reference ships (tests/golden/plonk_setup_cases.npz: its r1cs + the ptau slices `plonk setup`
    reads) with oracle.plonk.plonk_setup, and checks it is the reference's file byte for byte (sha256)."""
    import hashlib
    from oracle import oracle as orc
    from oracle import plonk
    n = int(g[f"{tag}_n"][0])
    ptau = orc.write_binfile("ptau", 1, [(1, bytes(g["ptau_header"])), (2, bytes(g[f"{tag}_ptau2"])), (3, bytes(g["ptau3"])),
                                         (12, bytes((n - 1) * 64) + bytes(g[f"{tag}_ptau12"]))])
    zkey = plonk.plonk_setup(bytes(g[f"{tag}_r1cs"]), ptau)
    assert hashlib.sha256(zkey).digest() == bytes(g[f"{tag}_zkey_sha256"]), "plonk_setup does not reproduce the reference zkey"
    return zkey, bytes(g[f"{tag}_wtns"])
