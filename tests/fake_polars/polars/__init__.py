"""
A minimal stand-in for the handful of Polars names the `pmm` expression namespace touches — TEST INFRASTRUCTURE.

The image has no Polars wheel, so `tests/test_polars_api.py` (written against real Polars) is skipped here.  This stub
lets `tests/test_namespace_stub.py` EXECUTE `PmmNamespace.topk / .matmul`, the `_topk` / `_matmul` Series branch and the
declared-versus-returned dtype contract anyway: Series are thin wrappers over pyarrow arrays, an Expr records its
`map_batches` call, and `DataFrame.select` evaluates it and checks the returned Series against the declared
`return_dtype` (what Polars itself would reject as a schema mismatch).  Nothing here is used by the product.
"""
from __future__ import annotations

import numpy as np
import pyarrow as pa

__version__ = "0.0-stub"


class DataType:
    def __eq__(self, other):
        other = other() if isinstance(other, type) else other
        return type(self) is type(other) and self.__dict__ == other.__dict__

    def __hash__(self):
        return hash((type(self).__name__, repr(sorted(self.__dict__.items(), key=str))))

    def __repr__(self):
        return f"{type(self).__name__}({', '.join(f'{k}={v!r}' for k, v in self.__dict__.items())})"


class Float16(DataType):
    pass


class Float32(DataType):
    pass


class Float64(DataType):
    pass


class UInt32(DataType):
    pass


class Int64(DataType):
    pass


def _inst(t):
    return t() if isinstance(t, type) else t


class List(DataType):
    def __init__(self, inner):
        self.inner = _inst(inner)


class Array(DataType):
    def __init__(self, inner, shape):
        self.inner = _inst(inner)
        self.size = int(shape)


class Struct(DataType):
    def __init__(self, fields):
        self.fields = {k: _inst(v) for k, v in dict(fields).items()}


def _dtype_of(t: pa.DataType) -> DataType:
    if pa.types.is_float16(t):
        return Float16()
    if pa.types.is_float32(t):
        return Float32()
    if pa.types.is_float64(t):
        return Float64()
    if pa.types.is_uint32(t):
        return UInt32()
    if pa.types.is_integer(t):
        return Int64()
    if pa.types.is_fixed_size_list(t):
        return Array(_dtype_of(t.value_type), t.list_size)
    if pa.types.is_list(t) or pa.types.is_large_list(t):
        return List(_dtype_of(t.value_type))
    if pa.types.is_struct(t):
        return Struct({t.field(i).name: _dtype_of(t.field(i).type) for i in range(t.num_fields)})
    raise TypeError(f"stub: unsupported arrow type {t}")


class Series:
    def __init__(self, name="", values=None, dtype=None):
        if not isinstance(name, str):
            name, values = "", name
        if isinstance(values, Series):
            values = values._a
        if isinstance(values, np.ndarray) and values.ndim == 2:
            values = pa.FixedSizeListArray.from_arrays(pa.array(values.reshape(-1)), values.shape[1])
        if not isinstance(values, (pa.Array, pa.ChunkedArray)):
            values = pa.array(values)
        self.name, self._a = name, values

    def to_arrow(self):
        return self._a

    @property
    def dtype(self):
        return _dtype_of(self._a.type)

    def __len__(self):
        return len(self._a)

    def to_list(self):
        return self._a.to_pylist()


class Expr:
    def __init__(self, column=None, parent=None, fn=None, is_elementwise=False, return_dtype=None):
        self.column, self.parent, self.fn = column, parent, fn
        self.is_elementwise, self.return_dtype = is_elementwise, return_dtype

    def map_batches(self, function, return_dtype=None, *, is_elementwise=False, **_):
        return Expr(parent=self, fn=function, is_elementwise=is_elementwise, return_dtype=_inst(return_dtype) if return_dtype is not None else None)

    def _evaluate(self, frame):
        if self.parent is None:
            return frame[self.column]
        out = self.fn(self.parent._evaluate(frame))
        if not isinstance(out, Series):
            raise TypeError("stub: a map_batches function must return a Series")
        if self.return_dtype is not None and out.dtype != self.return_dtype and len(out) > 0:
            raise TypeError(f"schema mismatch: map_batches declared {self.return_dtype!r} but the function returned {out.dtype!r}")
        return out


class _Api:
    @staticmethod
    def register_expr_namespace(name):
        def deco(cls):
            setattr(Expr, name, property(lambda self: cls(self)))
            return cls
        return deco


api = _Api()


def col(name):
    return Expr(column=name)


class DataFrame:
    def __init__(self, data):
        self._cols = {k: (v if isinstance(v, Series) else Series(k, v)) for k, v in dict(data).items()}

    def __getitem__(self, name):
        return self._cols[name]

    def get_column(self, name):
        return self._cols[name]

    def select(self, *exprs):
        out = {}
        for i, e in enumerate(exprs):
            s = e._evaluate(self)
            out[s.name or f"col{i}"] = s
        return DataFrame(out)
