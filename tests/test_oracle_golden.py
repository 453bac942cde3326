"""Pins the oracle against the reference's golden output files
(test/integration/single_barcode.jl:2-45 -> test/results/{demo1_R1,demo1_R2,demo2})."""
import pytest

import hostref


@pytest.mark.parametrize("idx", [0, 1, 2])
def test_oracle_reproduces_reference_golden_files(refdata, tmp_path, idx):
    case = hostref.demo_cases(refdata)[idx]
    n = hostref.run_case(case, str(tmp_path / "out"), hostref.oracle_classifier)
    assert n == {0: 24, 1: 24, 2: 76}[idx]
