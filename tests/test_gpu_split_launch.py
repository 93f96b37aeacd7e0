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


# The fused force kernel also performs the NEXT step's first half drift speculatively (force.cu MODE 3) and
# api.cu undoes it whenever end-of-step positions are needed.  Observing every step (test_gpu_golden.py) undoes
# it every step; observing rarely lets the pre-drifted state run through rebuilds and expiry tests.
@pytest.mark.parametrize("name", ["bench1", "bench8", "slab2", "collision2"])
def test_predrifted_steps_match_reference_golden(name):
    rec, cfg, steps, init_build = load_golden(name)
    worst = check_cuda_against_records(rec, steps, init_build, None, observe_every=25)
    print(name, worst)


def test_plain_path_without_predrift(monkeypatch):
    monkeypatch.setenv("MDB200_NO_PREDRIFT", "1")
    rec, cfg, steps, init_build = load_golden("bench1")
    check_cuda_against_records(rec, steps, init_build, None, observe_every=25)


# The fused NVE launch normally walks the INNER list (entries within cutoff + delta, re-derived every ~16 steps,
# force.cu PRUNE); MDB200_NO_PRUNE keeps it on the full list.  Both must meet the same parity.
@pytest.mark.parametrize("name", ["bench1", "bench8"])
def test_outer_list_only_matches_reference_golden(name, monkeypatch):
    monkeypatch.setenv("MDB200_NO_PRUNE", "1")
    rec, cfg, steps, init_build = load_golden(name)
    check_cuda_against_records(rec, steps, init_build, None, observe_every=25)


# Where the whole unit goes through one persistent force launch, that launch also derives the inner list (force.cu
# DERIVE); MDB200_FUSED_DERIVE=0 keeps the separate compaction pass.  Both must meet the same parity, also with rare
# observations (the pre-drifted state runs through the deriving launches).
@pytest.mark.parametrize("split", [False, True])
@pytest.mark.parametrize("fused", ["1", "0"])
@pytest.mark.parametrize("name", ["bench1", "bench8", "collision2", "langevinT0"])
def test_fused_inner_list_derivation(name, fused, split, monkeypatch):
    monkeypatch.setenv("MDB200_FORCE_PERSIST", "1")
    monkeypatch.setenv("MDB200_FUSED_DERIVE", fused)
    if split:  # interior and boundary tiles in two launches (the overlapped multi-process step): both derive
        monkeypatch.setenv("MDB200_FORCE_SPLIT", "1")
    monkeypatch.setenv("MDB200_TILE", "2,2,2")
    rec, cfg, steps, init_build = load_golden(name)
    seed = int(cfg["Seed"]) if cfg.get("DumpInit") == "Langevin" else None
    check_cuda_against_records(rec, steps, init_build, seed, observe_every=25)


# The first build of a store has no previous longest row to size a one-pass search with: on large units a count pass over
# every 8th tile provides one (tile_search.cu).  Forced onto the small fixtures here, where the sample is a handful of
# tiles: some cases build in one pass from it, others overflow and fall back to the two passes -- same lists either way.
@pytest.mark.parametrize("name", ["bench1", "bench8", "collision2", "slab2", "bench1_c25"])
def test_first_build_sized_from_a_tile_sample(name, monkeypatch):
    monkeypatch.setenv("MDB200_SAMPLE_TILES", "1")
    rec, cfg, steps, init_build = load_golden(name)
    check_cuda_against_records(rec, steps, init_build, None, observe_every=25)
