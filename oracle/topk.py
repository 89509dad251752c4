"""CPU oracle of the evaluation search: exact inner-product k-nearest-neighbours with hnswlib's "ip" conventions.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py). The reference builds an hnswlib index (dalm/eval/utils.py:18-42:
`hnswlib.Index(space="ip", dim)`, `init_index(max_elements, ef_construction=200, M=100)`, `add_items(data, arange(n))`) and
queries it with `set_ef(100); knn_query(q, k)` (:53-56), which returns (labels, distances) with distance = 1 - <q,p>, nearest
first. hnswlib is a third-party dependency that is absent offline and unpinned in the reference's pyproject ("hnswlib" without
a version): "parity unpinned" for its APPROXIMATE graph search; what is restated here is the exact search it approximates
(its published definition of the inner-product space), in float64, ties broken towards the lower id."""
from __future__ import annotations

import numpy as np


def knn_query(data: np.ndarray, queries: np.ndarray, k: int):
    """-> (labels [nq,k] int64, distances [nq,k] float64), distance = 1 - inner product, ascending"""
    d = np.asarray(data, dtype=np.float64)
    q = np.atleast_2d(np.asarray(queries, dtype=np.float64))
    scores = q @ d.T
    order = np.lexsort((np.broadcast_to(np.arange(d.shape[0]), scores.shape), -scores), axis=1)[:, :k]
    return order.astype(np.int64), 1.0 - np.take_along_axis(scores, order, axis=1)
