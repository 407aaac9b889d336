"""Writes tests/golden/segment/: one tiny persisted segment (meta file, vestigial data file, one field data file) in the layout
of SURVEY Appendix B, from seeded inputs.  tests/test_segment_files.py::test_committed_golden_segment parses the committed bytes
with libjvgpu's loader and checks that today's writer mirror still produces exactly these bytes, so neither side can drift alone.

    python tests/golden/make_golden_segment.py      # rewrites the three files; commit the result
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
import jvpkg  # noqa: E402

OUT = Path(__file__).resolve().parent / "segment"
SEGMENT_ID = bytes(range(0xA0, 0xB0))
N, DIM, R, M, K = 40, 6, 4, 3, 16


def arrays():
    rng = np.random.default_rng(20261017)
    vec = (rng.integers(-8, 9, (N, DIM)) / 8.0).astype(np.float32)          # exactly representable: no platform rounding
    adj = np.full((N, R), -1, np.int32)
    for i in range(N):
        d = int(rng.integers(1, R + 1))
        adj[i, :d] = rng.choice(np.delete(np.arange(N), i), d, replace=False)
    docs = np.arange(N, dtype=np.int32) * 3 + 1                              # ordinal != docId; maxDoc 130 -> two-byte vints
    docs[7] = -1                                                             # a deleted document
    cb = (rng.integers(-16, 17, K * DIM) / 16.0).astype(np.float32)
    gcent = (rng.integers(-4, 5, DIM) / 4.0).astype(np.float32)
    codes = rng.integers(0, K, (N, M)).astype(np.uint8)
    return vec, adj, docs, cb, gcent, codes


def segment(jv):
    vec, adj, docs, cb, gcent, codes = arrays()
    fd = jv.FieldData(jv.VectorSimilarityFunction.EUCLIDEAN, vec, adj, 11, jv.GraphNodeIdToDocMap(docs, 130), M, K, cb, gcent, codes)
    return jv.Segment(130, {"vec": fd})


def write(directory):
    jv = jvpkg.load()
    return jv.JVectorWriter.write(segment(jv), directory, "_g", "JVector_0", field_numbers={"vec": 2}, segment_id=SEGMENT_ID)


if __name__ == "__main__":
    for p in write(OUT).values():
        print(p, p.stat().st_size)
