"""CPU: the algorithmic-FLOP formulas behind every TFLOP/s figure (`bench.py::step_flops`, `tools/bench_sd3.py`) reproduce
the per-step totals SURVEY.md §8(d) / BASELINE.md §3 derive for the named configurations."""
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tools"))


def test_flux_step_flops_match_the_survey_totals():
    import bench
    D, H, T = 3072, 24, 512
    for N, n_cond, E, want_total, want_gemm, want_attn in ((1024, 1, 6, 45.7, 42.3, 3.4),      # cfg2
                                                            (4096, 1, 6, 159.3, 127.8, 31.5),   # cfg3
                                                            (4096, 3, 12, 170.9, 135.9, 35.0)):  # cfg4 (S-variant)
        g, a = bench.step_flops(D, H, N, T, 19, 38, 19, 38, E, executed=False, n_cond=n_cond)
        assert (g + a) / 1e12 == pytest.approx(want_total, rel=5e-3)
        assert g / 1e12 == pytest.approx(want_gemm, rel=1e-2) and a / 1e12 == pytest.approx(want_attn, rel=1.5e-2)
    g_ex, a_ex = bench.step_flops(D, H, 4096, T, 19, 38, 19, 38, 6, executed=True)
    g, a = bench.step_flops(D, H, 4096, T, 19, 38, 19, 38, 6, executed=False)
    # the native path skips only work whose result the reference discards: ~1 % of the step, attention untouched
    assert a_ex == a and 0.985 < (g_ex + a_ex) / (g + a) < 1.0


def test_sd3_step_flops_match_the_survey_total():
    import bench_sd3
    g, a = bench_sd3.sd3_step_flops(1536, 4096, 333, 24, 13, 13, 6, 683, executed=False)
    assert (g + a) / 1e12 == pytest.approx(25.9, rel=1e-2)
    assert g / 1e12 == pytest.approx(16.1, rel=3e-2) and a / 1e12 == pytest.approx(9.8, rel=3e-2)
