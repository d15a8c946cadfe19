"""The oracle must keep reproducing the committed golden vectors (tests/golden/*.npz)."""
import os
import sys

import numpy as np
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import make_golden  # noqa: E402


@pytest.mark.parametrize("name", sorted(make_golden.CASES))
def test_oracle_reproduces_golden(name):
    torch.set_num_threads(4)
    with torch.no_grad():
        got = make_golden.run_case(name, make_golden.CASES[name])
    want = np.load(os.path.join(HERE, "golden", name + ".npz"))["latents_per_step"]
    assert got.shape == want.shape
    err = np.linalg.norm(got - want) / np.linalg.norm(want)
    assert err < 1e-5, err
