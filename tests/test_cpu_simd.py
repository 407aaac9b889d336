"""The tuned CPU arm of the benchmark (oracle/jv_cpu_simd.c) against the bit-exact checker (oracle/jv_oracle.c).

Same algorithm, free summation order: ids may differ where scores differ in the last ulp, so the gate is recall equality within
sampling error, scores within 1e-5 relative where the same doc is returned, and identical counters up to a few visits."""
import numpy as np
import pytest

from oracle import oracle as O
from tests.helpers import clustered, embedded, make_fixture, recall


@pytest.mark.parametrize("sim,dim,m", [(O.SIM_DOT, 64, 16), (O.SIM_EUCLIDEAN, 96, 48), (O.SIM_COSINE, 128, 16), (O.SIM_DOT, 64, 0),
                                       (O.SIM_MIP, 64, 0)])
def test_simd_arm_matches_the_checker(sim, dim, m):
    base, q = clustered(4000, dim, 64, seed=50 + dim + m, normalize=sim in (O.SIM_DOT, O.SIM_MIP))
    fx = make_fixture(sim, base, q, max_degree=16, pq_m=m)
    ora = fx.oracle_index()
    fast = O.SimdIndex(ora)
    assert fast.isa in ("avx512", "avx2")
    wd, ws, wc, wst = ora.search(q, 10, 50)
    fd, fs, fc, fst = fast.search(q, 10, 50, threads=2)
    gt = ora.exact_topk(q, 10)[0]
    assert abs(recall(fd, gt) - recall(wd, gt)) <= 0.01
    np.testing.assert_array_equal(fc, wc)
    assert np.mean([np.array_equal(a, b) for a, b in zip(fd, wd)]) >= 0.9
    for i in range(len(q)):
        ref = {int(d): s for d, s in zip(wd[i], ws[i]) if d >= 0}
        for d, s in zip(fd[i], fs[i]):
            if int(d) in ref:
                assert abs(s - ref[int(d)]) <= 1e-5 * abs(ref[int(d)])
    assert abs(fst[:, 0].mean() - wst[:, 0].mean()) <= 0.02 * wst[:, 0].mean() + 2
    assert (fst[:, 3] == wst[:, 3]).all()


def test_simd_arm_headline_shape():
    base, q = embedded(3000, 768, 32, seed=9)
    fx = make_fixture(O.SIM_DOT, base, q, max_degree=32, pq_m=192)
    ora = fx.oracle_index()
    fast = O.SimdIndex(ora)
    wd = ora.search(q, 10, 50)[0]
    fd = fast.search(q, 10, 50)[0]
    gt = ora.exact_topk(q, 10)[0]
    assert abs(recall(fd, gt) - recall(wd, gt)) <= 0.01
