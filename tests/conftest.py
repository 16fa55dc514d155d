import os
import sys

import numpy as np
import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


@pytest.fixture(scope="session")
def state_dict():
    from tokenize_audio_b200 import synth
    return synth.synth_state_dict(0)


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    return {k: z[k] for k in z.files}


def golden_input(g):
    """int16 PCM [B, Nmax] -> input_values [B,1,N] fp32 exactly as the fixture generator fed the reference."""
    return (g["pcm"].astype(np.float32) / np.float32(32768.0))[:, None, :]


@pytest.fixture(scope="session")
def b200_model(state_dict):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from tokenize_audio_b200.encoder import MimiB200Model
    return MimiB200Model(state_dict, device="cuda:0")
