import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def read_fasta_records(path):
    """Golden segments files are FASTA-style '>name<idx>' records even with a .fastq suffix (chiron_eval.py:213)."""
    with open(path) as f:
        return [l.strip() for i, l in enumerate(f) if i % 2 == 1]


@pytest.fixture(scope="session")
def dna_model():
    from chiron_b200.model import load_model
    cfg, tensors, blob = load_model("DNA_default")
    return cfg, tensors, blob


@pytest.fixture(scope="session")
def rna_model():
    from chiron_b200.model import load_model
    cfg, tensors, blob = load_model("RNA_default")
    return cfg, tensors, blob
