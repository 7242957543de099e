"""
The oracle against the golden vectors: outputs of the unmodified reference driver
(DataProcessing.main_loop with the compiled SlidingWindow.pyx) captured by
tests/golden/make_golden.py.  Covers the whole path: index matching on all three platforms,
phasing, platform transforms, filters, sliding window, HR donor, table order and counts.
"""
import pytest

import golden_util as gu
from oracle import scar_oracle as orc


@pytest.mark.parametrize("name", gu.golden_names())
def test_oracle_reproduces_reference_driver(name):
    blob = gu.load(name)
    case, golden = blob["case"], blob["golden"]
    prob = orc.Problem(**gu.problem_inputs(case))
    f0, f1 = gu.index_fields(case)
    records = orc.classify_batch(prob, case["reads"], f0, f1)
    assert len(records) == len(case["reads"])
    gu.compare_with_golden(case, golden, records)


def test_golden_fixtures_present():
    names = gu.golden_names()
    for need in ("c1_single_target", "c2_96_samples", "c3_12_targets_mh", "c4_long_300bp",
                 "ragged_hr_donor", "truseq", "ramsden"):
        assert need in names
