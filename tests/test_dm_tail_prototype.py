"""The float64 prototype of the dispersion-tail quadrature (tools/dm_tail_prototype.py, the specification of
WarpTiles::dm_tail in -DNGRTD_DM_TAIL builds) stays pinned to the reference's golden vectors.  CPU only."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))


def test_prototype_matches_reference_golden_vectors():
    import dm_tail_prototype as P
    worst, med, mx, skipped = P.run(verbose=False)
    assert skipped == 0
    assert worst < 1e-12, worst                    # tolerance of the path is 1e-10 (north_star); measured 8.4e-15
    assert mx <= 512 and med <= 256, (med, mx)     # ~216 nodes instead of 25,128 tail lags
