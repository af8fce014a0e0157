"""opensot_b200 -- B200-native batched Stack-of-Tasks QP engine (hot path of OpenSoT's
AutoStack::update() + iHQP::solve()). See DESIGN.md."""
from . import _abi
from ._abi import StackDesc, Layout, make_desc
from .engine import (BatchedStackSolver, EngineError, NativeLibraryMissing, blob_layout,
                     smem_bytes_per_instance, measure_fp64_tflops, measure_fp32_tflops, lib, lib_path, host_empty)

from . import synth  # noqa: E402

__all__ = ["synth", "StackDesc", "Layout", "make_desc", "BatchedStackSolver", "EngineError",
           "NativeLibraryMissing", "blob_layout", "smem_bytes_per_instance",
           "measure_fp64_tflops", "measure_fp32_tflops", "lib", "lib_path", "host_empty", "_abi"]
