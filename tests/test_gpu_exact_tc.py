"""K5 on the tensor cores (csrc/jv_exact_tc.cu): bf16 tcgen05 candidate generation + canonical fp32 re-scoring must return exactly
what the fp32 brute-force kernel and the oracle return — ids bit for bit (ties -> lower docId), scores bit for bit."""
import os

import numpy as np
import pytest

from oracle import oracle as O
from tests.helpers import Fixture, clustered

pytestmark = pytest.mark.gpu


@pytest.fixture
def force_tc(jv):
    os.environ["JVGPU_EXACT_TC"] = "1"
    yield
    os.environ.pop("JVGPU_EXACT_TC", None)


def ring_fixture(sim, base, queries, ord_to_doc=None):
    """Brute force never reads the graph: a ring keeps the fixture cheap."""
    n = base.shape[0]
    adj = np.full((n, 4), -1, np.int32)
    adj[:, 0] = (np.arange(n) + 1) % n
    return Fixture(sim, np.ascontiguousarray(base, np.float32), np.ascontiguousarray(queries, np.float32), adj, 0, ord_to_doc,
                   None if ord_to_doc is None else n)


def _check(jv, fx, q, k, fallbacks=0):
    ora = fx.oracle_index()
    wd, ws, wc = ora.exact_topk(q, k)
    with fx.gpu_index(jv) as gi:
        gi.refresh_knobs()
        gd, gs, gc = gi.exact_topk(q, k)
        done, redone = gi.exact_tc_counters()
        assert done == 1, "the tensor-core path did not answer the batch"
        assert fallbacks <= redone <= 4 * fallbacks, (redone, fallbacks)  # queries re-answered by the fp32 kernel
    np.testing.assert_array_equal(gc, wc)
    np.testing.assert_array_equal(gd, wd)
    np.testing.assert_array_equal(gs.view(np.uint32), ws.view(np.uint32))


@pytest.mark.parametrize("sim", [O.SIM_EUCLIDEAN, O.SIM_DOT, O.SIM_COSINE, O.SIM_MIP])
@pytest.mark.parametrize("n,dim,k", [(9000, 96, 10), (20000, 128, 100), (5000, 768, 1)])
def test_tensor_core_brute_force_matches_oracle(jv, force_tc, sim, n, dim, k):
    base, q = clustered(n, dim, 150, seed=31 + n + dim, normalize=sim in (O.SIM_DOT,))
    if sim == O.SIM_EUCLIDEAN:
        base, q = base * 3.0, q * 3.0  # non-unit norms: the per-vector bias and the norm-scaled error bound are exercised
    fx = ring_fixture(sim, base, q[:150])
    _check(jv, fx, q[:150], k)


def test_tensor_core_brute_force_ties_and_deleted_docs(jv, force_tc):
    base, q = clustered(6000, 64, 130, seed=5, normalize=True)
    base[100:140] = base[99]            # 41 identical vectors: ties broken by the lower docId
    q[0] = base[99]
    rng = np.random.default_rng(3)
    ord_to_doc = np.arange(6000, dtype=np.int32)
    ord_to_doc[rng.random(6000) < 0.2] = -1   # deleted
    fx = ring_fixture(O.SIM_DOT, base, q[:130], ord_to_doc)
    _check(jv, fx, q[:130], 10)


def test_tensor_core_brute_force_odd_shapes(jv, force_tc):
    # dim not a multiple of 64 (zero-padded K blocks), n not a multiple of 256 (zero-filled rows), nq not a multiple of 128
    base, q = clustered(4099, 100, 37, seed=77, normalize=True)
    fx = ring_fixture(O.SIM_DOT, base, q[:37])
    _check(jv, fx, q[:37], 10)


def test_overflowing_queries_are_answered_by_the_fp32_kernel(jv, force_tc):
    """3 000 vectors within 1e-4 of each other: for the queries next to them more candidates lie inside the 2-eps margin than the
    re-scoring stage holds, so exactly those queries go through the fp32 kernel — same answers, the rest of the batch stays on
    the tensor cores."""
    base, q = clustered(9000, 64, 140, seed=11, normalize=True)
    rng = np.random.default_rng(8)
    base[1000:4000] = base[999] + 1e-4 * rng.standard_normal((3000, 64)).astype(np.float32)
    q[3] = base[999]
    q[77] = base[2000]
    fx = ring_fixture(O.SIM_DOT, base, q[:140])
    _check(jv, fx, q[:140], 10, fallbacks=2)
