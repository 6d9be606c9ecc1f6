"""ctypes wrapper for oracle/c_oracle.c.  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py)."""

from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libc_oracle.so")
_lib = None

QUANT = {"fp16": 0, "int8": 1, "int4": 2, "fp32": 3}
OUT = {"bf16": 0, "fp16": 1}


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "c_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "libc_oracle.so"])
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        L.oracle_index_create.restype = C.c_void_p
        L.oracle_index_create.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32]
        L.oracle_index_destroy.argtypes = [C.c_void_p]
        L.oracle_match.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_int]
        L.oracle_match_all.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_int]
        L.oracle_embed.restype = C.c_int
        L.oracle_embed.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64,
                                   C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int64, C.c_int, C.c_void_p,
                                   C.c_void_p, C.c_void_p, C.c_int]
        L.oracle_embed_ex.restype = C.c_int
        L.oracle_embed_ex.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64,
                                      C.c_void_p, C.c_int64, C.c_void_p, C.c_int, C.c_void_p, C.c_int64, C.c_int64, C.c_int,
                                      C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        L.oracle_f32_to_f16.restype = C.c_uint16
        L.oracle_f32_to_f16.argtypes = [C.c_float]
        L.oracle_f32_to_bf16.restype = C.c_uint16
        L.oracle_f32_to_bf16.argtypes = [C.c_float]
        L.oracle_f16_to_f32.restype = C.c_float
        L.oracle_f16_to_f32.argtypes = [C.c_uint16]
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


class COracleIndex:
    """Flat-array f-gram vocabulary: tokens int32 [N, max_n] (pad -1), lens uint8 [N], id = row."""

    def __init__(self, vocab_tokens: np.ndarray, vocab_lens: np.ndarray):
        self.toks = np.ascontiguousarray(vocab_tokens, dtype=np.int32)
        self.lens = np.ascontiguousarray(vocab_lens, dtype=np.uint8)
        self.n, self.max_n = self.toks.shape
        self.h = lib().oracle_index_create(_p(self.toks), _p(self.lens), self.n, self.max_n)
        if not self.h:
            raise ValueError("oracle_index_create failed (duplicate key, bad length, or out of memory)")

    def __del__(self):
        if getattr(self, "h", None):
            lib().oracle_index_destroy(self.h)
            self.h = None

    def match(self, ids2d: np.ndarray, nthreads: int = 1):
        ids2d = np.ascontiguousarray(ids2d, dtype=np.int64)
        B, L = ids2d.shape
        oid = np.empty((B, L), dtype=np.int32)
        olen = np.empty((B, L), dtype=np.uint8)
        lib().oracle_match(self.h, _p(ids2d), B, L, _p(oid), _p(olen), nthreads)
        return oid, olen

    def match_all(self, ids2d: np.ndarray, nthreads: int = 1):
        ids2d = np.ascontiguousarray(ids2d, dtype=np.int64)
        B, L = ids2d.shape
        out = np.empty((B, L, self.max_n), dtype=np.int32)
        lib().oracle_match_all(self.h, _p(ids2d), B, L, _p(out), nthreads)
        return out

    def embed(self, quant: str, D: int, group: int, payload: np.ndarray, row_stride: int, scales, scale_stride: int,
              base_bits: np.ndarray, ids2d: np.ndarray, out_dtype: str, nthreads: int = 1, out=None,
              pos_bits: np.ndarray = None, additive: bool = False):
        """payload / scales are raw byte views; strides in bytes.  Returns (out_bits, id, len, err).
        ``pos_bits`` ([>= L, D] uint16 in out_dtype) / ``additive``: as py_oracle.embed_forward."""
        ids2d = np.ascontiguousarray(ids2d, dtype=np.int64)
        B, L = ids2d.shape
        base_bits = np.ascontiguousarray(base_bits, dtype=np.uint16)
        if out is None:
            out = np.empty((B, L, D), dtype=np.uint16)
        oid = np.empty((B, L), dtype=np.int32)
        olen = np.empty((B, L), dtype=np.uint8)
        if pos_bits is None and not additive:
            err = lib().oracle_embed(self.h, QUANT[quant], D, group, _p(payload), row_stride, _p(scales), scale_stride,
                                     _p(base_bits), base_bits.shape[0], _p(ids2d), B, L, OUT[out_dtype], _p(out),
                                     _p(oid), _p(olen), nthreads)
            return out, oid, olen, err
        if pos_bits is not None:
            pos_bits = np.ascontiguousarray(pos_bits, dtype=np.uint16)
            if pos_bits.shape[0] < L or pos_bits.shape[1] != D:
                raise ValueError("pos_bits must be [>= L, D]")
        err = lib().oracle_embed_ex(self.h, QUANT[quant], D, group, _p(payload), row_stride, _p(scales), scale_stride,
                                    _p(base_bits), base_bits.shape[0], _p(pos_bits) if pos_bits is not None else None,
                                    1 if additive else 0, _p(ids2d), B, L, OUT[out_dtype], _p(out), _p(oid), _p(olen), nthreads)
        return out, oid, olen, err
