"""CPU: the plain-C oracle against the committed golden fixtures (generated from the unmodified
reference by oracle/make_golden.py) -- bit for bit, including the pair generation ORDER."""
import os

import pytest

from common import GOLDEN, check_oracle_against_records, load_golden

NAMES = sorted(f[:-4] for f in os.listdir(GOLDEN) if f.endswith(".npz"))


@pytest.mark.parametrize("name", NAMES)
def test_oracle_reproduces_golden_bitwise(name):
    rec, cfg, steps, init_build = load_golden(name)
    seed = int(cfg["Seed"]) if cfg.get("DumpInit") == "Langevin" else None
    check_oracle_against_records(rec, steps, init_build, seed)
