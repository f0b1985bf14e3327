"""Replays the committed golden fixtures (tests/golden/*.json, made by tools/make_golden.py) on the oracle."""
import glob
import json
import os

import pytest

from strolle_b200 import scenes
from tools.make_golden import CASES, run_case

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "*.json")))


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-5] for p in GOLDEN])
def test_oracle_matches_golden(oracle, blue_noise, path):
    doc = json.load(open(path))
    name = os.path.basename(path)[:-5]
    got = run_case(oracle.OracleEngine(blue_noise=blue_noise), CASES[name])
    for buf, want in doc["buffers"].items():
        assert got[buf]["sha256"] == want["sha256"], f"{name}:{buf}"


@pytest.mark.gpu
@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-5] for p in GOLDEN])
def test_cuda_matches_golden(blue_noise, path):
    """The CUDA path (exact mode, through the C ABI) reproduces the committed digests without running the oracle."""
    import strolle_b200
    doc = json.load(open(path))
    name = os.path.basename(path)[:-5]
    got = run_case(strolle_b200.Engine(blue_noise=blue_noise, exact=True), CASES[name])
    for buf, want in doc["buffers"].items():
        assert got[buf]["sha256"] == want["sha256"], f"{name}:{buf}"
