"""``pyopencl.array`` stand-in: host ndarrays posing as device arrays (see package docstring)."""
import numpy as np


class Array:
    def __init__(self, a):
        self.data = a  # what the reference passes to the kernel launch
        self.shape = a.shape
        self.dtype = a.dtype

    def get(self):
        return self.data


def to_device(queue, ary):
    return Array(np.ascontiguousarray(ary).copy())


def empty(queue, shape, dtype):
    # cl_array.empty leaves memory uninitialised; a fixed fill keeps the
    # golden vectors reproducible without changing any defined output.
    return Array(np.full(shape, -7, dtype=dtype) if np.issubdtype(np.dtype(dtype), np.integer)
                 else np.full(shape, -7.0, dtype=dtype))
