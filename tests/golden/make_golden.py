#!/usr/bin/env python3
"""
Generates tests/golden/*.json|*.npz — the known-answer vectors the reference's own tests hold for the
pmm.topk / pmm.matmul path.  Run from the repo root:  python tests/golden/make_golden.py

Two kinds of vectors:
  (1) hand-computed known answers transcribed from the reference's unit/integration tests
      (inputs and expected outputs are literal values in those tests; file:line cited per case);
  (2) the reference's seeded NumPy-equivalence tests: inputs regenerated with the same
      np.random.seed(42) + randn call sequence, expected values computed with the same NumPy
      expression the reference test asserts against (rtol stated per case).
The reference extension itself cannot be imported here (Rust + polars absent), so no output of the
reference binary is recorded; these are the reference's *test oracles*, not its outputs.
"""
import json
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))

known = {
    "_about": "Known-answer vectors from NivekNey/polars-matmul v0.1.4 tests; see make_golden.py",
    "select": [
        {"src": "src/topk.rs:83-95", "matrix": [[0.1, 0.9, 0.5], [0.8, 0.2, 0.6]], "k": 2, "higher": True,
         "dtype": "f64", "index": [[1, 2], [0, 2]]},
        {"src": "src/topk.rs:98-110", "matrix": [[0.1, 0.9, 0.5], [0.8, 0.2, 0.6]], "k": 2, "higher": True,
         "dtype": "f32", "index": [[1, 2], [0, 2]]},
        {"src": "src/topk.rs:113-125", "matrix": [[0.1, 0.9, 0.5], [0.8, 0.2, 0.6]], "k": 2, "higher": False,
         "dtype": "f64", "index": [[0, 2], [1, 2]]},
    ],
    "scores": [
        {"src": "src/metrics.rs:401-410", "dtype": "f64", "metric": "dot",
         "query": [[1.0, 0.0], [0.0, 1.0]], "corpus": [[1.0, 0.0], [0.0, 1.0], [1.0, 1.0]],
         "expect": {"0,0": 1.0, "0,1": 0.0, "1,1": 1.0}, "atol": 1e-10},
        {"src": "src/metrics.rs:413-422", "dtype": "f32", "metric": "dot",
         "query": [[1.0, 0.0], [0.0, 1.0]], "corpus": [[1.0, 0.0], [0.0, 1.0], [1.0, 1.0]],
         "expect": {"0,0": 1.0, "0,1": 0.0, "1,1": 1.0}, "atol": 1e-5},
        {"src": "src/metrics.rs:425-434", "dtype": "f64", "metric": "cosine",
         "query": [[1.0, 0.0], [0.0, 1.0]], "corpus": [[2.0, 0.0], [0.0, 3.0]],
         "expect": {"0,0": 1.0, "1,1": 1.0, "1,0": 0.0}, "atol": 1e-10},
    ],
    "list_to_dense": [
        {"src": "src/matmul.rs:526-538", "dtype": "f64", "values": [1, 2, 3, 4, 5, 6], "offsets": [0, 3, 6],
         "shape": [2, 3], "expect": {"0,0": 1.0, "1,2": 6.0}, "atol": 1e-10},
        {"src": "src/matmul.rs:541-553", "dtype": "f32", "values": [1, 2, 3, 4, 5, 6], "offsets": [0, 3, 6],
         "shape": [2, 3], "expect": {"0,0": 1.0, "1,2": 6.0}, "atol": 1e-5},
    ],
    "topk": [
        {"src": "tests/test_polars_matmul.py:13-53", "dtype": "f64", "metric": "cosine", "k": 2,
         "query": [[1.0, 0.0, 0.0], [0.0, 1.0, 0.0]],
         "corpus": [[1.0, 0.0, 0.0], [0.0, 1.0, 0.0], [0.0, 0.0, 1.0]],
         "top1_index": [0, 1], "top1_score": [1.0, 1.0], "atol": 1e-6, "n_results": 2},
        {"src": "tests/test_polars_matmul.py:77-95", "dtype": "f64", "metric": "dot", "k": 2,
         "query": [[2.0, 0.0]], "corpus": [[1.0, 0.0], [3.0, 0.0]],
         "top1_index": [1], "top1_score": [6.0], "atol": 1e-6, "n_results": 2},
        {"src": "tests/test_polars_matmul.py:97-115", "dtype": "f64", "metric": "euclidean", "k": 2,
         "query": [[0.0, 0.0]], "corpus": [[3.0, 4.0], [1.0, 0.0]],
         "top1_index": [1], "top1_score": [1.0], "atol": 1e-6, "n_results": 2},
        {"src": "tests/test_polars_matmul.py:117-133 (k > corpus clamps)", "dtype": "f64", "metric": "cosine",
         "k": 10, "query": [[1.0, 0.0]], "corpus": [[1.0, 0.0], [0.0, 1.0]],
         "top1_index": [0], "top1_score": [1.0], "atol": 1e-6, "n_results": 2},
        {"src": "tests/test_polars_matmul.py:56-75 (explode/unnest: 2 queries x k=2 -> 4 rows)", "dtype": "f64",
         "metric": "cosine", "k": 2, "query": [[1.0, 0.0], [0.0, 1.0]],
         "corpus": [[1.0, 0.0], [0.0, 1.0], [0.5, 0.5]],
         "top1_index": [0, 1], "top1_score": [1.0, 1.0], "atol": 1e-6, "n_results": 2},
    ],
    "matmul": [
        {"src": "tests/test_polars_matmul.py:169-184", "dtype": "f64",
         "left": [[1.0, 2.0], [3.0, 4.0]], "right": [[1.0, 0.0], [0.0, 1.0]],
         "expect": [[1.0, 2.0], [3.0, 4.0]], "rtol": 1e-6},
        {"src": "tests/test_polars_matmul.py:204-222 (flatten, row-major)", "dtype": "f64",
         "left": [[1.0, 0.0], [0.0, 1.0], [1.0, 1.0]], "right": [[1.0, 0.0], [0.0, 1.0]],
         "expect": [[1.0, 0.0], [0.0, 1.0], [1.0, 1.0]], "rtol": 1e-5, "flat": [1.0, 0.0, 0.0, 1.0, 1.0, 1.0]},
        {"src": "tests/test_polars_matmul.py:241-258 (Array input)", "dtype": "f64",
         "left": [[1.0, 2.0, 3.0, 4.0], [5.0, 6.0, 7.0, 8.0]],
         "right": [[1.0, 0.0, 0.0, 0.0], [0.0, 1.0, 0.0, 0.0]],
         "expect": [[1.0, 2.0], [5.0, 6.0]], "rtol": 1e-5},
        {"src": "tests/test_polars_matmul.py:369-388 (f32 in -> f32 out)", "dtype": "f32",
         "left": [[1.0, 2.0], [3.0, 4.0]], "right": [[1.0, 0.0], [0.0, 1.0]],
         "expect": [[1.0, 2.0], [3.0, 4.0]], "rtol": 1e-5},
    ],
    "dtype_dispatch": [
        {"src": "tests/test_polars_matmul.py:369-399,434-447,449-464; src/matmul.rs:308",
         "cases": [["f32", "f32", "f32"], ["f64", "f64", "f64"], ["f32", "f64", "f64"], ["f64", "f32", "f64"]]}
    ],
    "errors": [
        {"src": "tests/test_polars_matmul.py:302-309", "call": "topk", "metric": "invalid_metric",
         "query": [[1.0, 0.0]], "corpus": [[1.0, 0.0]], "k": 1, "match": "Unknown metric"},
        {"src": "tests/test_polars_matmul.py:334-343", "call": "topk", "metric": "cosine",
         "query": [[1.0, 0.0]], "corpus": [], "k": 1, "match": "Empty"},
        {"src": "tests/test_polars_matmul.py:345-353", "call": "matmul",
         "query": [[1.0, 2.0]], "corpus": [[1.0, 2.0, 3.0]], "match": "Dimension mismatch"},
        {"src": "tests/test_polars_matmul.py:355-363", "call": "topk", "metric": "cosine",
         "query": [[1.0, 2.0]], "corpus": [[1.0, 2.0, 3.0]], "k": 1, "match": "Dimension mismatch"},
    ],
    "metric_strings": {"src": "src/metrics.rs:19-27",
                       "ok": {"cosine": 0, "COSINE": 0, "Dot": 1, "dot": 1, "euclidean": 2, "L2": 2, "l2": 2},
                       "bad": ["invalid_metric", "manhattan", ""]},
}

with open(os.path.join(HERE, "reference_known_answers.json"), "w") as f:
    json.dump(known, f, indent=1)

# ---- (2) seeded NumPy-equivalence fixtures ------------------------------------------------------
# tests/test_polars_matmul.py:186-202 and tests/test_performance.py:78-97: matmul vs np.dot, rtol 1e-5
np.random.seed(42)
left = np.random.randn(10, 32)
right = np.random.randn(20, 32)
np.savez(os.path.join(HERE, "matmul_seed42_10x20x32_f64.npz"), left=left, right=right,
         expect=np.dot(left, right.T), rtol=1e-5)

# tests/test_polars_matmul.py:264-296: cosine k=20 of seed-42 randn(5,16) x randn(20,16); the
# reference test compares the per-query SORTED score lists with rtol 1e-5 (indices are not asserted).
np.random.seed(42)
q = np.random.randn(5, 16)
c = np.random.randn(20, 16)
qn = q / np.linalg.norm(q, axis=1, keepdims=True)
cn = c / np.linalg.norm(c, axis=1, keepdims=True)
expected = np.dot(qn, cn.T)
np.savez(os.path.join(HERE, "cosine_seed42_5x20x16_f64.npz"), query=q, corpus=c,
         expect_sorted_desc=-np.sort(-expected, axis=1), rtol=1e-5, k=20)

# examples/benchmark_topk.py:187-203 + :122-138: self-check randn(100,64) x randn(500,64) f64, k=10,
# sorted scores rtol 1e-4, comparator numpy_topk_cosine (:14-33).
np.random.seed(42)
q = np.random.randn(100, 64)
c = np.random.randn(500, 64)
qn = q / np.sqrt(np.sum(q ** 2, axis=1, keepdims=True))
cn = c / np.sqrt(np.sum(c ** 2, axis=1, keepdims=True))
sim = np.dot(qn, cn.T)
top = -np.sort(-sim, axis=1)[:, :10]
np.savez(os.path.join(HERE, "bench_selfcheck_seed42_100x500x64_f64.npz"), query=q, corpus=c,
         expect_topk_scores=top, rtol=1e-4, k=10)
print("golden vectors written to", HERE)
