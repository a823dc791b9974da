import os
import sys

import pytest

# Parity with the oracle / the reference fixtures is defined at dropout p = 0 (oracle/make_golden.py runs the reference in
# eval mode): models built by the suites get BERT's train-mode dropout switched off unless a test sets the probabilities
# itself (tests/test_dropout_gpu.py does, and checks the p = 0.1 default of the constructor).
os.environ.setdefault("SIMSEG_BERT_DROPOUT", "0")

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu on the GPU box")


@pytest.fixture(scope="session")
def cuda():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")
