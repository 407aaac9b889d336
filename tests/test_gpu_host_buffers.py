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
