"""GPU: the interior / boundary split of the force launch (the two launches the multi-process path puts
around the ghost exchange, api.cu:calculate) must give the same parity as the single launch.  The hook
MDB200_FORCE_SPLIT runs the two launches back to back on one GPU."""
import os

import pytest

from common import check_cuda_against_records, load_golden

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ["bench1", "bench8", "slab2"])
def test_split_force_launch_matches_reference_golden(name, monkeypatch):
    monkeypatch.setenv("MDB200_FORCE_SPLIT", "1")
    rec, cfg, steps, init_build = load_golden(name)
    worst = check_cuda_against_records(rec, steps, init_build, None)
    print(name, worst)
