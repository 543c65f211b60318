"""ctypes binding of oracle/ref_eval.c (the C restatement of the reference evaluator).
TEST INFRASTRUCTURE / CPU BASELINE ONLY: see the header of ref_eval.c."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(_HERE, "_build", "liboracle.so")


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE])
    return LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        src = os.path.join(_HERE, "ref_eval.c")
        if not os.path.exists(LIB) or os.path.getmtime(src) > os.path.getmtime(LIB):
            build()
        L = ctypes.CDLL(LIB)
        L.oracle_load.restype = ctypes.c_void_p
        L.oracle_load.argtypes = [ctypes.c_char_p, ctypes.c_size_t]
        L.oracle_free.argtypes = [ctypes.c_void_p]
        L.oracle_info.argtypes = [ctypes.c_void_p, ctypes.POINTER(ctypes.c_uint64)]
        L.oracle_evaluate_batch.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_int]
        L.oracle_op_duo.argtypes = [ctypes.c_uint, ctypes.c_char_p, ctypes.c_char_p, ctypes.c_char_p]
        L.oracle_op_uno.argtypes = [ctypes.c_uint, ctypes.c_char_p, ctypes.c_char_p]
        _lib = L
    return _lib


class CGraph:
    def __init__(self, data: bytes):
        self.L = lib()
        self.h = self.L.oracle_load(data, len(data))
        if not self.h:
            raise ValueError("oracle: cannot parse graph")
        info = (ctypes.c_uint64 * 4)()
        self.L.oracle_info(self.h, info)
        self.n_nodes, self.n_inputs, self.n_witness, self.n_ops = [int(x) for x in info]

    def evaluate_batch(self, inputs: np.ndarray, n_threads=1) -> np.ndarray:
        """inputs uint8 [B, I, 32] -> witness uint8 [B, W, 32]"""
        inputs = np.ascontiguousarray(inputs, dtype=np.uint8)
        B = inputs.shape[0]
        assert inputs.shape[1:] == (self.n_inputs, 32), inputs.shape
        out = np.empty((B, self.n_witness, 32), dtype=np.uint8)
        self.L.oracle_evaluate_batch(self.h, inputs.ctypes.data, B, out.ctypes.data, n_threads)
        return out

    def __del__(self):
        if getattr(self, "h", None):
            self.L.oracle_free(self.h)
            self.h = None


def op_duo(op: int, a: int, b: int) -> int:
    r = ctypes.create_string_buffer(32)
    lib().oracle_op_duo(op, a.to_bytes(32, "little"), b.to_bytes(32, "little"), r)
    return int.from_bytes(r.raw, "little")


def op_uno(op: int, a: int) -> int:
    r = ctypes.create_string_buffer(32)
    lib().oracle_op_uno(op, a.to_bytes(32, "little"), r)
    return int.from_bytes(r.raw, "little")
