"""GPU: the CUDA path (through the C-ABI) against the committed golden fixtures of the reference.
No file outside the repository is read: this is the parity gate that runs on the B200 box."""
import os

import pytest

from common import GOLDEN, check_cuda_against_records, load_golden

pytestmark = pytest.mark.gpu
# 'langevin' carries thermal noise from the reference's sequential std::mt19937 stream: bit-wise
# unpinnable on a per-particle GPU RNG; see test_gpu_langevin_statistics.py
NAMES = sorted(f[:-4] for f in os.listdir(GOLDEN) if f.endswith(".npz") and f[:-4] != "langevin")


@pytest.mark.parametrize("name", NAMES)
def test_cuda_matches_reference_golden(name):
    rec, cfg, steps, init_build = load_golden(name)
    seed = int(cfg["Seed"]) if cfg.get("DumpInit") == "Langevin" else None
    worst = check_cuda_against_records(rec, steps, init_build, seed)
    print(name, worst)
