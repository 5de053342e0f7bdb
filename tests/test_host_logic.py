"""CPU-only tests: C-ABI loading/exports, argument validation that needs no GPU, Arrow marshalling,
the packed candidate format, shard bounds."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_abi_library_loads_and_exports_every_declared_symbol():
    from polars_matmul_b200 import _native
    L = _native.lib()
    header = open(os.path.join(ROOT, "include", "pmm.h")).read()
    declared = sorted(set(re.findall(r"PMM_API[^;(]*?\b(pmm_[a-z_0-9]+)\s*\(", header)))
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(L, name), f"{name} is declared in include/pmm.h but not exported by libpmm_b200.so"
    assert sorted(_native.EXPORTED_SYMBOLS) == declared
    assert L.pmm_version().decode().startswith("0.1.4")


def test_matrix_struct_layout_matches_header():
    from polars_matmul_b200._native import PmmMatrix
    assert ctypes.sizeof(PmmMatrix) == 4 * 8 + 2 * 8 + 2 * 4
    assert PmmMatrix.n_rows.offset == 32 and PmmMatrix.dim.offset == 40 and PmmMatrix.dtype.offset == 48


def test_metric_parsing_without_gpu(golden_dir):
    import json
    from polars_matmul_b200 import _native
    ms = json.load(open(os.path.join(golden_dir, "reference_known_answers.json")))["metric_strings"]
    for s, want in ms["ok"].items():
        assert _native.metric_from_str(s) == want
    for s in ms["bad"]:
        with pytest.raises(RuntimeError, match="Unknown metric"):
            _native.metric_from_str(s)
    L = _native.lib()
    assert L.pmm_higher_is_better(0) == 1 and L.pmm_higher_is_better(1) == 1 and L.pmm_higher_is_better(2) == 0
    f16, f32, f64 = _native.DTYPE_F16, _native.DTYPE_F32, _native.DTYPE_F64
    assert L.pmm_working_dtype(f32, f32) == f32 and L.pmm_working_dtype(f32, f64) == f64
    assert L.pmm_working_dtype(f64, f32) == f64 and L.pmm_working_dtype(f16, f32) == f32


def test_validation_order_needs_no_gpu():
    """Checks that precede any device work follow the reference's order (src/matmul.rs:473-490, :433-441)."""
    import polars_matmul_b200 as pmm
    # empty queries: OK even with an invalid metric
    assert len(pmm._topk(np.empty((0, 2)), np.ones((1, 2)), 1, "invalid")) == 0
    with pytest.raises(RuntimeError, match="Unknown metric"):
        pmm.topk_arrays(np.ones((1, 2)), np.ones((1, 2)), 1, "invalid")
    with pytest.raises(RuntimeError, match="Empty"):
        pmm.topk_arrays(np.ones((1, 2)), np.empty((0, 2)), 1, "cosine")
    with pytest.raises(RuntimeError, match="Dimension mismatch: left has 2 dimensional vectors, right has 3"):
        pmm.topk_arrays(np.ones((1, 2)), np.ones((1, 3)), 1, "cosine")
    with pytest.raises(RuntimeError, match="Dimension mismatch"):
        pmm.matmul_array(np.ones((1, 2)), np.ones((1, 3)))
    with pytest.raises(OverflowError):
        pmm.topk_arrays(np.ones((1, 2)), np.ones((1, 2)), -1, "cosine")
    assert pmm._matmul(np.empty((0, 2), np.float32), np.ones((1, 2), np.float32)).type.value_type == __import__("pyarrow").float32()


def test_no_cpu_fallback_without_device():
    """On a box without a GPU a compute call must fail loudly, not fall back."""
    import polars_matmul_b200 as pmm
    from polars_matmul_b200 import _native
    if _native.device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(_native.PmmError, match="no CUDA device"):
        pmm.topk_arrays(np.ones((2, 4), np.float32), np.ones((3, 4), np.float32), 1, "cosine")
    with pytest.raises(_native.PmmError, match="no CUDA device"):
        pmm.matmul_array(np.ones((2, 4), np.float32), np.ones((3, 4), np.float32))


def test_product_never_imports_the_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "polars_matmul_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("the oracle", "").replace("oracle's", "").replace("CPU oracle", "") \
                    or f in ("pmm_generic.cu", "pmm_prep.cu"), f"{f} mentions the oracle"
                assert "import oracle" not in src and "from oracle" not in src and "libpmm_oracle" not in src, f


def test_arrow_marshalling():
    import pyarrow as pa
    from polars_matmul_b200.arrow import to_host_matrix
    h = to_host_matrix(pa.array([[1.0, 2.0], [3.0, 4.0]], type=pa.list_(pa.float32())))
    assert h.values.dtype == np.float32 and h.offsets.tolist() == [0, 2, 4] and (h.n_rows, h.dim) == (2, 2)
    fsl = pa.array([[1.0, 2.0], [3.0, None], [5.0, 6.0]], type=pa.list_(pa.float64(), 2))
    h = to_host_matrix(fsl.slice(1))
    assert h.offsets is None and (h.n_rows, h.dim) == (2, 2) and h.values[0] == 3.0 and h.values[2:].tolist() == [5.0, 6.0]
    assert h.validity is not None and (h.validity[0] & 0b1111) == 0b1101
    ll = pa.array([[1, 2], [3, 4], None, [5]], type=pa.large_list(pa.int32()))
    h = to_host_matrix(ll.slice(1))
    assert h.values.dtype == np.float64 and h.offsets.tolist() == [2, 4, 4, 5] and h.dim == 2
    assert h.row_validity is not None and (h.row_validity[0] & 0b111) == 0b101
    with pytest.raises(RuntimeError, match="First element is null"):
        to_host_matrix(pa.array([None, [1.0]], type=pa.list_(pa.float32())))
    ch = pa.chunked_array([pa.array([[1.0, 2.0]], type=pa.list_(pa.float32(), 2)),
                           pa.array([[3.0, 4.0]], type=pa.list_(pa.float32(), 2))])
    h = to_host_matrix(ch)      # multi-chunk Array column: views of the chunks' own buffers, no host-side concatenation
    assert h.n_rows == 2 and h.dim == 2 and [c.tolist() for c in h.chunks] == [[1.0, 2.0], [3.0, 4.0]]
    m = h.c_struct()
    assert m.reserved == 2 and m.n_rows == 2 and m.offsets is None
    chn = pa.chunked_array([pa.array([[1.0, None]], type=pa.list_(pa.float32(), 2)),      # nulls: concatenated on the host
                            pa.array([[3.0, 4.0]], type=pa.list_(pa.float32(), 2))])
    h = to_host_matrix(chn)
    assert h.chunks is None and h.n_rows == 2 and h.validity is not None
    chl = pa.chunked_array([pa.array([[1.0, 2.0]], type=pa.large_list(pa.float32())),      # List columns likewise
                            pa.array([[3.0, 4.0]], type=pa.large_list(pa.float32()))])
    h = to_host_matrix(chl)
    assert h.chunks is None and h.n_rows == 2 and h.values.tolist() == [1.0, 2.0, 3.0, 4.0]
    h = to_host_matrix(np.arange(6, dtype=np.int32).reshape(2, 3))
    assert h.values.dtype == np.float64
    h = to_host_matrix(np.ones((2, 3), np.float16))
    assert h.dtype_code == 0


def test_result_builders_zero_copy_layout():
    import pyarrow as pa
    from polars_matmul_b200 import arrow
    idx = np.array([[1, 2], [0, 2]], np.uint32)
    sc = np.array([[0.9, 0.5], [0.8, 0.6]])
    r = arrow.topk_to_arrow(idx, sc)
    assert r.type == pa.large_list(pa.struct([("index", pa.uint32()), ("score", pa.float64())]))
    assert r.to_pylist()[1] == [{"index": 0, "score": 0.8}, {"index": 2, "score": 0.6}]
    m = arrow.matmul_to_arrow(np.arange(6, dtype=np.float32).reshape(2, 3))
    assert m.type == pa.list_(pa.float32(), 3) and m.to_pylist() == [[0, 1, 2], [3, 4, 5]]


def test_packed_candidate_format_orders_like_the_oracle(oracle):
    from polars_matmul_b200.sharded import pack_candidates, unpack_candidates
    rng = np.random.default_rng(0)
    s = rng.standard_normal(1000).astype(np.float32)
    s[[3, 500]] = np.nan
    s[[10, 11, 12]] = 0.25          # ties -> lower index first
    s[20], s[21] = 0.0, -0.0        # -0.0 == +0.0
    s[30], s[31] = np.inf, -np.inf
    idx = np.arange(1000, dtype=np.uint32)
    for higher in (True, False):
        c = pack_candidates(idx, s, higher)
        order = np.argsort(c)[::-1]                                     # larger u64 = better
        oi, osc = oracle.select(s[None, :], 1000, higher)
        assert np.array_equal(order, oi[0])
        i2, s2 = unpack_candidates(c, higher)
        assert np.array_equal(i2, idx)
        assert np.array_equal(np.isnan(s2), np.isnan(s))
        ok = ~np.isnan(s)
        assert np.array_equal(s2[ok], s[ok].astype(np.float64))


def test_shard_bounds():
    from polars_matmul_b200.sharded import shard_bounds
    assert shard_bounds(10, 4) == [(0, 3), (3, 6), (6, 9), (9, 10)]
    assert shard_bounds(8, 8) == [(i, i + 1) for i in range(8)]
    assert shard_bounds(3, 4) == [(0, 1), (1, 2), (2, 3), (3, 3)]
    b = shard_bounds(10_000_000, 8)
    assert b[0] == (0, 1_250_000) and b[-1] == (8_750_000, 10_000_000)


def test_small_results_are_plain_arrays_and_device_flag_is_declared():
    """result_empty only takes page-locked pool blocks for results of at least 1 MB (no GPU needed below that);
    the PMM_MATRIX_ON_DEVICE flag of the shim equals the header's."""
    from polars_matmul_b200 import _native
    a = _native.result_empty((10, 7), np.float64)
    assert a.shape == (10, 7) and a.dtype == np.float64 and a.base is None
    header = open(os.path.join(ROOT, "include", "pmm.h")).read()
    m = re.search(r"#define\s+PMM_MATRIX_ON_DEVICE\s+(\d+)", header)
    assert m and int(m.group(1)) == _native.PMM_MATRIX_ON_DEVICE
    dm = _native.dev_matrix(0x1000, 5, 8, _native.DTYPE_F32, flags=_native.PMM_MATRIX_ON_DEVICE)
    assert dm.reserved == 1 and dm.n_rows == 5 and dm.dim == 8


def test_bind_near_gpu_is_best_effort_without_a_device():
    from polars_matmul_b200 import sharded
    before = os.sched_getaffinity(0)
    assert sharded.bind_near_gpu(0) is None          # no GPU / no topology here: nothing may change
    assert os.sched_getaffinity(0) == before


def test_every_documented_option_is_accepted_and_unknown_ones_are_not():
    """include/pmm.h lists the runtime options; each must be known to pmm_set_option (setting an option needs no GPU)."""
    from polars_matmul_b200 import _native
    header = open(os.path.join(ROOT, "include", "pmm.h")).read()
    block = header[header.index("Tuning / diagnostics"):header.index("PMM_API int pmm_set_option")]
    names = set(re.findall(r'"([a-z0-9_]+)"', block)) - {"kernel"}
    assert {"tc_levels", "f16r_wide", "host_chunked", "verify", "profile"} <= names
    defaults = {"tc_levels": 3, "tc_cg": 2, "tc_sync_tiles": 32, "host_chunked": 1, "verify": 1, "f16r_wide": 1, "tc_clm": 1,
                "host_chunk_min_rows": 16384, "host_chunk_min_mb": 64, "generic_workspace_mb": 0, "seed_retry": 1, "f64_tc": 1,
                "multi_gpu": 1, "multi_gpu_min_gflop": 4000, "stage": 1, "stage_slot_mb": 32, "stage_slots": 4,
                "workspace_cache_mb": 24576, "pipeline": 0, "pipeline_min_gflop": 2000, "prep_fast": 1, "rescore_stream_loads": 1, "multipass": 1, "rescore_fixed": 0}
    for n in sorted(names):
        if n in ("release_workspace", "generic_workspace_mb") or n.startswith("tc_dbg"):
            continue
        _native.set_option(n, defaults.get(n, 0))      # restores the default as it goes
    with pytest.raises(_native.PmmError):
        _native.set_option("no_such_option", 1)
    # the wrong-result timing modes are compiled out of the default build
    for v in (1, 2, 3):
        with pytest.raises(_native.PmmError, match="PMM_DIAG"):
            _native.set_option("tc_debug_skip", v)
    _native.set_option("tc_debug_skip", 8)
    _native.set_option("tc_debug_skip", 0)
    # per-thread overrides: same key space (minus the immediate, process-wide ones), dropped with key=None
    _native.set_thread_option("tc_levels", 2)
    with pytest.raises(_native.PmmError):
        _native.set_thread_option("stage_threads", 2)
    with pytest.raises(_native.PmmError):
        _native.set_thread_option("no_such_option", 1)
    _native.set_thread_option(None)
