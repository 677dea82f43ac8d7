"""Full-size parity (BASELINE.json sizes) that the CPU oracle cannot deliver in seconds: the oracle restatement runs as
torch-eager bf16 ON the GPU over the native model's own weight storage (tests/eager_oracle_compare.py) and the final
velocity is compared. Both sides are bf16 end to end through 48-85 blocks, so the bar is the north-star's final-latent
one (cosine >= 0.999) plus routing agreement, not the per-block fp32-oracle bar of the small-size tests."""
import os
import sys
from pathlib import Path

import pytest

pytestmark = pytest.mark.gpu
sys.path.insert(0, str(Path(__file__).resolve().parent))


def test_sd35_medium_full_size_matches_eager_oracle():
    """cfg5 architecture: SD3.5-medium UniGenSD3, 1024^2 + 1 condition (4096 + 4096 + 333 tokens), forward batch 2."""
    import eager_oracle_compare as E
    rec = E.main(["--workload", "cfg5", "--batch", "2", "--steps", "1"])
    par = rec["full_size_parity"]
    assert par["cosine"] >= 0.999 and par["routing_agreement"] >= 0.98, rec


@pytest.mark.skipif(os.environ.get("UG_FULL_SIZE_TESTS", "0") != "1",
                    reason="37 GB of weights + two torch-eager steps (~2 min): set UG_FULL_SIZE_TESTS=1; result committed as "
                           "profiles/r01_eager_oracle_cfg3.json")
def test_flux_cfg3_full_size_matches_eager_oracle():
    import eager_oracle_compare as E
    rec = E.main(["--workload", "cfg3", "--steps", "1"])
    par = rec["full_size_parity"]
    assert par["cosine"] >= 0.999 and par["routing_agreement"] >= 0.98, rec
