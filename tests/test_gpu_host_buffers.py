"""Page-locked host buffers of the C-ABI (jv_host_alloc / jv_host_free / jv_host_register / jv_host_unregister, include/jvgpu.h;
INTEGRATION.md section 3b): the batch entry point returns the same answers from library-allocated, registered and pageable
buffers."""
import ctypes as C

import numpy as np
import pytest

from oracle import oracle as O
from tests.helpers import clustered, make_fixture

pytestmark = pytest.mark.gpu


def test_search_batch_from_allocated_registered_and_pageable_buffers(jv):
    N = jv.native
    lib = N.load()
    base, q = clustered(4000, 64, 300, seed=9, normalize=True)
    fx = make_fixture(O.SIM_DOT, base, q, max_degree=16, pq_m=16)
    nq, k, rk = q.shape[0], 10, 50
    with fx.gpu_index(jv, flags=N.FLAG_LUT_U8) as gi:
        want = gi.search(q, k, rk)                                   # pageable numpy buffers
        p = gi._params(k, rk, 0.0, 0.0, None, 0, 0)
        # (a) library-allocated
        hq, hq_p = N.host_alloc((nq, 64), np.float32)
        hq[:] = q
        hd, hd_p = N.host_alloc((nq, k), np.int32)
        hs, hs_p = N.host_alloc((nq, k), np.float32)
        hc, hc_p = N.host_alloc((nq,), np.int32)
        N.check(lib.jv_search_batch(gi.handle, hq_p, nq, C.addressof(p), hd_p, hs_p, hc_p, None, None))
        np.testing.assert_array_equal(hd, want.docs)
        np.testing.assert_array_equal(hs.view(np.uint32), want.scores.view(np.uint32))
        np.testing.assert_array_equal(hc, want.counts)
        # (b) registered memory the caller owns
        rq = np.ascontiguousarray(q)
        rd = np.empty((nq, k), np.int32)
        N.check(lib.jv_host_register(rq.ctypes.data, rq.nbytes))
        N.check(lib.jv_host_register(rd.ctypes.data, rd.nbytes))
        try:
            rs, rc = np.empty((nq, k), np.float32), np.empty(nq, np.int32)
            N.check(lib.jv_search_batch(gi.handle, rq.ctypes.data, nq, C.addressof(p), rd.ctypes.data, rs.ctypes.data, rc.ctypes.data, None, None))
            np.testing.assert_array_equal(rd, want.docs)
        finally:
            N.check(lib.jv_host_unregister(rq.ctypes.data))
            N.check(lib.jv_host_unregister(rd.ctypes.data))
        for ptr in (hq_p, hd_p, hs_p, hc_p):
            N.host_free(ptr)
    assert lib.jv_host_free(None) == N.JV_OK                          # free(NULL) is a no-op
    bad = C.c_void_p()
    assert lib.jv_host_alloc(0, C.byref(bad)) != N.JV_OK              # zero bytes: invalid argument
    assert lib.jv_host_unregister(None) != N.JV_OK


def test_index_create_rejects_out_of_range_arrays(jv):
    """jv_index_create range-checks the caller's decoded arrays on the device (a bad id would be an out-of-bounds read later)."""
    base, q = clustered(500, 16, 4, seed=2)
    fx = make_fixture(O.SIM_EUCLIDEAN, base, q, max_degree=8, pq_m=4, pq_k=64)
    ok = fx.gpu_index(jv)
    ok.close()
    bad_adj = fx.adjacency.copy()
    bad_adj[17, 3] = 500                                             # == n
    with pytest.raises(ValueError, match="adjacency"):
        jv.GpuIndex(fx.sim, fx.base, bad_adj, fx.entry, pq_m=fx.pq_m, pq_k=fx.pq_k, pq_codebooks=fx.codebooks, pq_global_centroid=fx.gcent,
                    pq_codes=fx.codes)
    bad_map = np.arange(500, dtype=np.int32)
    bad_map[3] = 900
    with pytest.raises(ValueError, match="ord_to_doc"):
        jv.GpuIndex(fx.sim, fx.base, fx.adjacency, fx.entry, ord_to_doc=bad_map, max_doc=600)
    bad_codes = fx.codes.copy()
    bad_codes[5, 1] = 200                                            # >= K = 64
    with pytest.raises(ValueError, match="pq_codes"):
        jv.GpuIndex(fx.sim, fx.base, fx.adjacency, fx.entry, pq_m=fx.pq_m, pq_k=fx.pq_k, pq_codebooks=fx.codebooks, pq_global_centroid=fx.gcent,
                    pq_codes=bad_codes)
    with pytest.raises(ValueError, match="max_doc"):
        jv.GpuIndex(fx.sim, fx.base, fx.adjacency, fx.entry, max_doc=10)
