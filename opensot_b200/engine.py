"""Python binding of the C ABI (include/opensot_b200.h) via ctypes.

PyTorch is used only for device memory and streams in the device-resident entry points;
the host-buffer entry point (`solve_host`) takes numpy arrays, exactly what the C ABI takes.
There is no CPU fallback: if the native library is missing or no CUDA device is present,
every compute call raises."""
import ctypes as C
import os

import numpy as np

from . import _abi
from ._abi import StackDesc, Layout, Options, EngineInfo, GatherSrc, PostureInputs

# (OSB_LIB: another build of the same library, for A/B measurements of kernel variants)
_LIB_PATH = os.environ.get("OSB_LIB") or os.path.join(os.path.dirname(os.path.abspath(__file__)), "lib", "libopensot_b200.so")
_lib = None

EXPORTS = [
    "osb_blob_layout", "osb_struct_sizes", "osb_smem_bytes_per_instance", "osb_last_error", "osb_version",
    "osb_device_count", "osb_engine_create", "osb_engine_create_f32", "osb_blob_stride_f32",
    "osb_engine_destroy", "osb_set_options", "osb_set_active_levels", "osb_solve_host",
    "osb_solve_device", "osb_solve_host_f32", "osb_solve_device_f32", "osb_stack_gather_device",
    "osb_produce_posture_device",
    "osb_engine_get_info", "osb_engine_device_buffers", "osb_engine_stream", "osb_engine_kernel_name",
    "osb_measure_fp64_tflops", "osb_measure_fp32_tflops",
    "osb_kinematics_create", "osb_produce_kinematics_device",
    "osb_states_configure", "osb_solve_states_host", "osb_states_download_blobs",
    "osb_set_eps_level", "osb_get_eps_level", "osb_get_options", "osb_host_alloc", "osb_host_free",
    "osb_gather_create", "osb_gather_connect", "osb_solve_device_gather", "osb_gather_barrier",
    "osb_gather_buffers", "osb_gather_barrier_failures",
    "osb_solve_host_submit", "osb_solve_host_submit_f32", "osb_solve_states_submit", "osb_solve_wait", "osb_solve_in_flight",
    "osb_wire_configure", "osb_wire_stride", "osb_wire_pack", "osb_solve_host_wire", "osb_solve_host_wire_submit",
]


class NativeLibraryMissing(RuntimeError):
    pass


def lib_path():
    return _LIB_PATH


def lib():
    """Loads libopensot_b200.so (built in-tree by opensot_b200/build.py)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_LIB_PATH):
        raise NativeLibraryMissing(
            "%s not found: run `python -c 'import __graft_entry__ as g; g.build()'` "
            "(there is no CPU fallback)" % _LIB_PATH)
    L = C.CDLL(_LIB_PATH)
    vp, i32, i64, dbl = C.c_void_p, C.c_int32, C.c_int64, C.c_double
    L.osb_blob_layout.argtypes = [C.POINTER(StackDesc), C.POINTER(Layout)]
    L.osb_blob_layout.restype = C.c_int
    L.osb_struct_sizes.argtypes = [C.POINTER(i32 * 6)]
    L.osb_struct_sizes.restype = None
    sizes = (i32 * 6)()
    L.osb_struct_sizes(C.byref(sizes))
    mine = [C.sizeof(t) for t in (StackDesc, Layout, Options, EngineInfo, GatherSrc, PostureInputs)]
    if list(sizes) != mine:
        raise RuntimeError("ctypes mirrors out of date: C %s vs python %s" % (list(sizes), mine))
    L.osb_smem_bytes_per_instance.argtypes = [C.POINTER(StackDesc)]
    L.osb_smem_bytes_per_instance.restype = i64
    L.osb_last_error.restype = C.c_char_p
    L.osb_version.restype = C.c_char_p
    L.osb_device_count.restype = C.c_int
    L.osb_engine_create.argtypes = [C.POINTER(vp), C.POINTER(StackDesc), i32, i32]
    L.osb_engine_create_f32.argtypes = [C.POINTER(vp), C.POINTER(StackDesc), i32, i32]
    L.osb_blob_stride_f32.argtypes = [C.POINTER(StackDesc)]
    L.osb_blob_stride_f32.restype = i32
    L.osb_solve_host_f32.argtypes = [vp, vp, i32, vp, vp, vp]
    L.osb_solve_device_f32.argtypes = [vp, vp, i32, vp, vp, vp, vp, vp]
    L.osb_engine_destroy.argtypes = [vp]
    L.osb_engine_destroy.restype = None
    L.osb_set_options.argtypes = [vp, C.POINTER(Options)]
    L.osb_set_active_levels.argtypes = [vp, C.POINTER(i32)]
    L.osb_solve_host.argtypes = [vp, vp, i32, vp, vp, vp]
    L.osb_solve_device.argtypes = [vp, vp, i32, vp, vp, vp, vp, vp]
    L.osb_stack_gather_device.argtypes = [vp, C.POINTER(GatherSrc), i32, vp, i32, vp]
    L.osb_produce_posture_device.argtypes = [vp, C.POINTER(PostureInputs), vp, i32, vp]
    L.osb_engine_get_info.argtypes = [vp, C.POINTER(EngineInfo)]
    L.osb_engine_device_buffers.argtypes = [vp, C.POINTER(vp), C.POINTER(vp), C.POINTER(vp)]
    L.osb_engine_stream.argtypes = [vp]
    L.osb_engine_stream.restype = vp
    L.osb_engine_kernel_name.argtypes = [vp]
    L.osb_engine_kernel_name.restype = C.c_char_p
    L.osb_measure_fp64_tflops.argtypes = [i32, i32]
    L.osb_measure_fp64_tflops.restype = dbl
    L.osb_kinematics_create.argtypes = [vp, vp]
    L.osb_produce_kinematics_device.argtypes = [vp, vp, i64, vp, i32, vp, i32, vp]
    L.osb_states_configure.argtypes = [vp, vp]
    L.osb_solve_states_host.argtypes = [vp, vp, i32, vp, vp]
    L.osb_solve_host_submit.argtypes = [vp, vp, i32, vp, vp, vp]
    L.osb_solve_host_submit_f32.argtypes = [vp, vp, i32, vp, vp, vp]
    L.osb_solve_states_submit.argtypes = [vp, vp, i32, vp, vp]
    L.osb_solve_wait.argtypes = [vp]
    L.osb_wire_configure.argtypes = [vp, vp, i32, vp]
    L.osb_wire_stride.argtypes = [vp]
    L.osb_wire_stride.restype = i32
    L.osb_wire_pack.argtypes = [vp, vp, i32, vp]
    L.osb_solve_host_wire.argtypes = [vp, vp, i32, vp, vp, vp]
    L.osb_solve_host_wire_submit.argtypes = [vp, vp, i32, vp, vp, vp]
    L.osb_solve_in_flight.argtypes = [vp]
    L.osb_solve_in_flight.restype = i32
    L.osb_states_download_blobs.argtypes = [vp, vp, i32]
    L.osb_set_eps_level.argtypes = [vp, i32, dbl]
    L.osb_get_eps_level.argtypes = [vp, i32]
    L.osb_get_eps_level.restype = dbl
    L.osb_get_options.argtypes = [vp, C.POINTER(Options)]
    L.osb_host_alloc.argtypes = [i64]
    L.osb_host_alloc.restype = vp
    L.osb_host_free.argtypes = [vp]
    L.osb_host_free.restype = None
    L.osb_measure_fp32_tflops.argtypes = [i32, i32]
    L.osb_measure_fp32_tflops.restype = dbl
    L.osb_gather_create.argtypes = [vp, i64, i32, i32, i64, C.c_char_p]
    L.osb_gather_connect.argtypes = [vp, i32, C.c_char_p]
    L.osb_solve_device_gather.argtypes = [vp, vp, i32, vp]
    L.osb_gather_barrier.argtypes = [vp, vp, i32]
    L.osb_gather_buffers.argtypes = [vp, C.POINTER(vp), C.POINTER(vp)]
    L.osb_gather_barrier_failures.argtypes = [vp]
    _lib = L
    return L


class EngineError(RuntimeError):
    pass


def _check(rc):
    if rc != 0:
        raise EngineError("opensot_b200 error %d: %s" % (rc, lib().osb_last_error().decode()))


class _PinnedBlock(object):
    def __init__(self, nbytes):
        self.ptr = lib().osb_host_alloc(nbytes)
        if not self.ptr:
            raise MemoryError("osb_host_alloc(%d) failed" % nbytes)

    def __del__(self):
        try:
            lib().osb_host_free(self.ptr)
        except Exception:       # interpreter shutdown
            pass


def host_empty(shape, dtype=np.float64):
    """numpy array over page-locked host memory from osb_host_alloc (cudaHostAlloc): the buffers osb_solve_host and its
    submitted / wire-format twins copy from and to at the full PCIe rate; freed with the array."""
    dtype = np.dtype(dtype)
    count = int(np.prod(shape))
    blk = _PinnedBlock(max(16, count * dtype.itemsize))
    buf = (C.c_char * (count * dtype.itemsize)).from_address(blk.ptr)
    buf._blk = blk                      # the array keeps the ctypes view alive, the view keeps the block alive
    return np.frombuffer(buf, dtype=dtype).reshape(shape)


def blob_layout(desc):
    L = Layout()
    _check(lib().osb_blob_layout(C.byref(desc), C.byref(L)))
    return L


def smem_bytes_per_instance(desc):
    return int(lib().osb_smem_bytes_per_instance(C.byref(desc)))


def measure_fp64_tflops(device=0, iters=4096):
    return float(lib().osb_measure_fp64_tflops(device, iters))


def measure_fp32_tflops(device=0, iters=4096):
    return float(lib().osb_measure_fp32_tflops(device, iters))


class BatchedStackSolver:
    """A batch of structurally identical stacks solved on one GPU.

    The batched counterpart of one `OpenSoT::solvers::iHQP` object
    (reference include/OpenSoT/solvers/iHQP.h:59-193): `solve_host` is iHQP::solve for B
    instances at once, `set_active_levels` is setActiveStack/activateAllStacks."""

    def __init__(self, desc, max_batch, device=0, store_levels=False, max_iter=0,
                 tol_feas=0.0, tol_rank=0.0, precision="f64"):
        assert precision in ("f64", "f32")
        self.desc = desc
        self.layout = blob_layout(desc)
        self.max_batch = int(max_batch)
        self.device = int(device)
        self.precision = precision
        self.np_dtype = np.float64 if precision == "f64" else np.float32
        self._stride = self.layout.stride if precision == "f64" else int(lib().osb_blob_stride_f32(C.byref(desc)))
        self._h = C.c_void_p()
        create = lib().osb_engine_create if precision == "f64" else lib().osb_engine_create_f32
        _check(create(C.byref(self._h), C.byref(desc), self.max_batch, self.device))
        self.store_levels = bool(store_levels)
        opt = Options(max_iter, 1 if store_levels else 0, tol_feas, tol_rank)
        _check(lib().osb_set_options(self._h, C.byref(opt)))

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            lib().osb_engine_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def n(self):
        return self.desc.n

    @property
    def stride(self):
        """elements between consecutive instance blobs (fp32 engines pad to a multiple of 4)"""
        return self._stride

    def pack(self, blobs64):
        """fp64 blobs [B, layout.stride] -> blobs of this engine's precision and stride."""
        blobs64 = np.asarray(blobs64)
        if self.precision == "f64":
            return np.ascontiguousarray(blobs64, dtype=np.float64)
        out = np.zeros((blobs64.shape[0], self._stride), dtype=np.float32)
        out[:, :blobs64.shape[1]] = blobs64
        return out

    def info(self):
        out = EngineInfo()
        _check(lib().osb_engine_get_info(self._h, C.byref(out)))
        return {f: getattr(out, f) for f, _ in EngineInfo._fields_}

    def kernel_name(self):
        """the solver kernel this engine launches (register-row instance or the generic kernel)"""
        return lib().osb_engine_kernel_name(self._h).decode()

    def host_stream(self):
        """torch view of the stream `solve_host` works on (for event timing)."""
        import torch
        return torch.cuda.ExternalStream(lib().osb_engine_stream(self._h), device=self.device)

    def set_active_levels(self, active):
        arr = (C.c_int32 * self.desc.n_levels)(*[1 if a else 0 for a in active])
        _check(lib().osb_set_active_levels(self._h, arr))

    def set_eps_level(self, level, eps):
        """absolute diagonal shift of one level (level < 0: all): iHQP::setEpsRegularisation(eps[, i])"""
        _check(lib().osb_set_eps_level(self._h, level, eps))

    def get_eps_level(self, level):
        return float(lib().osb_get_eps_level(self._h, level))

    def activate_all_levels(self):
        self.set_active_levels([True] * self.desc.n_levels)

    def solve_host(self, blobs, x=None, status=None, x_levels=None):
        """blobs: float64 [B, stride] host array (pinned or pageable). Returns (x, status[, x_levels])."""
        dt = self.np_dtype
        blobs = np.ascontiguousarray(blobs, dtype=dt)
        B = blobs.shape[0]
        assert blobs.shape[1] == self.stride, "blob stride mismatch"
        if x is None:
            x = np.empty((B, self.n), dtype=dt)
        if status is None:
            status = np.empty(B, dtype=np.int32)
        if self.store_levels and x_levels is None:
            x_levels = np.empty((B, self.desc.n_levels, self.n), dtype=dt)
        assert x.dtype == dt
        fn = lib().osb_solve_host if self.precision == "f64" else lib().osb_solve_host_f32
        _check(fn(self._h, blobs.ctypes.data, B, x.ctypes.data, status.ctypes.data,
                  x_levels.ctypes.data if x_levels is not None else None))
        return (x, status, x_levels) if self.store_levels else (x, status)

    def submit_host(self, blobs, x, status, x_levels=None):
        """Throughput mode (osb_solve_host_submit): enqueue one host-buffer call and return; at most two in flight.
        `blobs`, `x`, `status` are the caller's (page-locked) numpy arrays of the engine's dtype; they must stay alive and
        untouched until the matching wait()."""
        dt = self.np_dtype
        assert blobs.dtype == dt and blobs.flags.c_contiguous and blobs.shape[1] == self.stride
        assert x.dtype == dt and x.flags.c_contiguous and status.dtype == np.int32
        fn = lib().osb_solve_host_submit if self.precision == "f64" else lib().osb_solve_host_submit_f32
        _check(fn(self._h, blobs.ctypes.data, blobs.shape[0], x.ctypes.data, status.ctypes.data,
                  x_levels.ctypes.data if x_levels is not None else None))

    def wire_configure(self, static_segments, template_blob):
        """Declares the fields of a blob that never change between cycles (osb_wire_configure): `static_segments` is a list
        of (offset, count) in elements (layout offsets), `template_blob` a float64 array of layout.stride values carrying
        them. Returns the wire stride (elements of the engine's precision per instance)."""
        segs = np.ascontiguousarray(np.asarray(sorted(static_segments), dtype=np.int32).reshape(-1, 2))
        tmpl = np.ascontiguousarray(template_blob, dtype=np.float64)
        assert tmpl.shape[0] >= self.layout.stride
        _check(lib().osb_wire_configure(self._h, segs.ctypes.data if len(segs) else None, len(segs), tmpl.ctypes.data))
        return int(lib().osb_wire_stride(self._h))

    def wire_stride(self):
        return int(lib().osb_wire_stride(self._h))

    def wire_pack(self, blobs, out=None):
        """Full blobs of the engine's precision ([B, stride]) -> wire rows ([B, wire_stride])."""
        dt = self.np_dtype
        assert blobs.dtype == dt and blobs.flags.c_contiguous and blobs.shape[1] == self.stride
        if out is None:
            out = np.empty((blobs.shape[0], self.wire_stride()), dtype=dt)
        _check(lib().osb_wire_pack(self._h, blobs.ctypes.data, blobs.shape[0], out.ctypes.data))
        return out

    def solve_host_wire(self, wire, x=None, status=None):
        dt = self.np_dtype
        assert wire.dtype == dt and wire.flags.c_contiguous and wire.shape[1] == self.wire_stride()
        B = wire.shape[0]
        if x is None:
            x = np.empty((B, self.n), dtype=dt)
        if status is None:
            status = np.empty(B, dtype=np.int32)
        _check(lib().osb_solve_host_wire(self._h, wire.ctypes.data, B, x.ctypes.data, status.ctypes.data, None))
        return x, status

    def submit_host_wire(self, wire, x, status):
        """submit_host for wire rows (same buffer rules; wait() completes it)."""
        dt = self.np_dtype
        assert wire.dtype == dt and wire.flags.c_contiguous and wire.shape[1] == self.wire_stride()
        assert x.dtype == dt and x.flags.c_contiguous and status.dtype == np.int32
        _check(lib().osb_solve_host_wire_submit(self._h, wire.ctypes.data, wire.shape[0], x.ctypes.data, status.ctypes.data, None))

    def submit_states(self, q, x, status):
        """Submitted twin of solve_states_host (same buffer rules as submit_host)."""
        assert q.dtype == np.float64 and q.flags.c_contiguous and x.dtype == np.float64 and status.dtype == np.int32
        _check(lib().osb_solve_states_submit(self._h, q.ctypes.data, q.shape[0], x.ctypes.data, status.ctypes.data))

    def wait(self):
        """Blocks until the oldest submitted call has delivered its results (osb_solve_wait)."""
        _check(lib().osb_solve_wait(self._h))

    def in_flight(self):
        return int(lib().osb_solve_in_flight(self._h))

    def solve_device(self, d_blobs, d_x, d_status, d_x_levels=None, d_iters=None, stream=None):
        """All arguments are torch CUDA tensors on this engine's device; the launch is enqueued
        on `stream` (a torch.cuda.Stream, default: the current stream) and not synchronised."""
        import torch
        B = d_blobs.shape[0]
        tdt = torch.float64 if self.precision == "f64" else torch.float32
        assert d_blobs.dtype == tdt and d_blobs.is_contiguous() and d_blobs.shape[1] == self.stride
        assert d_x.dtype == tdt and d_x.is_contiguous() and d_status.dtype == torch.int32
        if stream is None:
            stream = torch.cuda.current_stream(self.device)
        fn = lib().osb_solve_device if self.precision == "f64" else lib().osb_solve_device_f32
        _check(fn(
            self._h, d_blobs.data_ptr(), B, d_x.data_ptr(), d_status.data_ptr(),
            d_x_levels.data_ptr() if d_x_levels is not None else None,
            d_iters.data_ptr() if d_iters is not None else None, stream.cuda_stream))

    def gather_device(self, sources, d_blobs, stream=None):
        """sources: list of dicts(src=tensor [B, rows_src, ld], rows, cols, dst_off, scale,
        row_index=int32 tensor or None). Device-side Aggregated/AutoStack::update()."""
        import torch
        if stream is None:
            stream = torch.cuda.current_stream(self.device)
        arr = (GatherSrc * len(sources))()
        for i, s in enumerate(sources):
            t = s["src"]
            assert t.dtype == torch.float64 and t.is_cuda
            arr[i].src = t.data_ptr()
            arr[i].src_stride = t.stride(0)
            arr[i].rows, arr[i].cols = s["rows"], s["cols"]
            arr[i].src_ld = s.get("src_ld", s["cols"])
            ri = s.get("row_index")
            arr[i].row_index = ri.data_ptr() if ri is not None else None
            arr[i].dst_off = s["dst_off"]
            arr[i].scale = s.get("scale", 1.0)
        _check(lib().osb_stack_gather_device(self._h, arr, len(sources), d_blobs.data_ptr(),
                                             d_blobs.shape[0], stream.cuda_stream))

    def log(self, path, blobs, x_levels=None, active=None, task_ids=None):
        """iHQP::log for a batch (src/solvers/iHQP.cpp:403-415, BackEnd.cpp:111-130, Task.h:437-453):
        the per-level QPs and task matrices under the reference's names, rebuilt on the host from
        fp64 blobs [B, layout.stride] and the per-level solutions of a store_levels engine; .mat or
        .npz by extension. Debugging aid, off the solve path (opensot_b200/qplog.py)."""
        from . import qplog
        arrays = qplog.level_qps(self.desc, self.layout, blobs, x_levels, active)
        arrays.update(qplog.task_logs(self.desc, self.layout, blobs, task_ids))
        qplog.save(path, arrays)
        return arrays

    # --- solve fused with the all-gather of the solutions over peer memory (osb_gather_*) ----
    def gather_create(self, global_batch, rank, world, first_instance):
        """allocates this rank's gather buffer; returns its CUDA IPC handle (bytes) for the peers"""
        h = C.create_string_buffer(_abi.OSB_IPC_HANDLE_BYTES)
        _check(lib().osb_gather_create(self._h, global_batch, rank, world, first_instance, h))
        self._gather_shape = (int(global_batch), self.desc.n)
        return h.raw

    def gather_connect(self, peer, handle):
        _check(lib().osb_gather_connect(self._h, peer, C.create_string_buffer(handle, _abi.OSB_IPC_HANDLE_BYTES)))

    def solve_device_gather(self, d_blobs, batch=None, stream=None):
        import torch
        if stream is None:
            stream = torch.cuda.current_stream(self.device)
        B = d_blobs.shape[0] if batch is None else batch
        assert d_blobs.is_cuda and d_blobs.dtype == (torch.float64 if self.precision == "f64" else torch.float32)
        _check(lib().osb_solve_device_gather(self._h, d_blobs.data_ptr(), B, stream.cuda_stream))

    def gather_barrier(self, stream=None, timeout_ms=2000):
        import torch
        if stream is None:
            stream = torch.cuda.current_stream(self.device)
        _check(lib().osb_gather_barrier(self._h, stream.cuda_stream, timeout_ms))

    def gather_barrier_failures(self):
        return int(lib().osb_gather_barrier_failures(self._h))

    def gather_tensors(self):
        """(x_all [global_batch, n], status_all [global_batch]) as torch views of the gather buffer"""
        import torch
        px, ps = C.c_void_p(), C.c_void_p()
        _check(lib().osb_gather_buffers(self._h, C.byref(px), C.byref(ps)))
        Bg, n = self._gather_shape
        f64 = self.precision == "f64"

        class _Raw:
            def __init__(self, ptr, shape, typestr):
                self.__cuda_array_interface__ = {"shape": shape, "typestr": typestr, "data": (ptr, False), "version": 2}
        dev = torch.device("cuda", self.device)
        x = torch.as_tensor(_Raw(px.value, (Bg, n), "<f8" if f64 else "<f4"), device=dev)
        st = torch.as_tensor(_Raw(ps.value, (Bg,), "<i4"), device=dev)
        return x, st

    def kinematics_create(self, tree):
        """uploads a kinematic tree (opensot_b200.kinematics.KinTree, e.g. from load_urdf) for produce_kinematics_device"""
        t = tree.c_struct()
        _check(lib().osb_kinematics_create(self._h, C.byref(t)))
        self._kin_tree = tree

    def produce_kinematics_device(self, d_blobs, q, tasks, stream=None):
        """Cartesian / CoM tasks from joint positions, on device (Cartesian.cpp:68-105, CoM.cpp:59-74).
        q: torch CUDA fp64 [B, n]; tasks: dicts(kind='cartesian'|'com', distal=link name or joint index, base=None,
        level, row0=0, lam=1.0, orientation_gain=1.0, desired=tensor [B,12]|[12]|[B,3]|[3] or None, desired_vel=...)."""
        import torch
        from . import kinematics as K
        if stream is None:
            stream = torch.cuda.current_stream(self.device)
        arr = (K.KinTaskC * len(tasks))()
        keep = []
        for i, t in enumerate(tasks):
            arr[i].kind = K.COM if t["kind"] == "com" else K.CARTESIAN
            link = lambda v: -1 if v is None else (self._kin_tree.link(v) if isinstance(v, str) else int(v))  # noqa: E731
            arr[i].distal, arr[i].base = link(t.get("distal")), link(t.get("base"))
            arr[i].level, arr[i].row0 = t["level"], t.get("row0", 0)
            arr[i].lambda_, arr[i].orientation_gain = t.get("lam", 1.0), t.get("orientation_gain", 1.0)
            for key in ("desired", "desired_vel"):
                v = t.get(key)
                if v is not None:
                    assert v.dtype == torch.float64 and v.is_cuda and v.is_contiguous()
                    keep.append(v)
                    setattr(arr[i], key, v.data_ptr())
                    setattr(arr[i], key + "_stride", v.stride(0) if v.dim() == 2 else 0)
        assert q.dtype == torch.float64 and q.is_cuda and q.is_contiguous()
        _check(lib().osb_produce_kinematics_device(self._h, q.data_ptr(), q.stride(0), arr, len(tasks), d_blobs.data_ptr(),
                                                   d_blobs.shape[0], stream.cuda_stream))

    def states_configure(self, blob_template, tasks=(), q_ref=None, q_min=None, q_max=None, q_neutral=None, qdot_max=None,
                         lam=1.0, bound_scaling=1.0, dT=1.0, postural_level=-1, joint_limit_set=-1, velocity_limit_set=-1):
        """State-input path (osb_states_configure): everything that does not depend on q, from HOST arrays.
        tasks: like produce_kinematics_device, with numpy `desired` ([12] pose | [3] CoM) / `desired_vel` shared by the batch."""
        from . import kinematics as K
        keep = [np.ascontiguousarray(blob_template, dtype=np.float64)]
        assert keep[0].size == self.layout.stride
        cfg = K.StatesConfigC()
        cfg.blob_template = keep[0].ctypes.data
        for name, v in (("q_ref", q_ref), ("q_min", q_min), ("q_max", q_max), ("q_neutral", q_neutral), ("qdot_max", qdot_max)):
            if v is not None:
                a = np.ascontiguousarray(v, dtype=np.float64); keep.append(a)
                assert a.size == self.n
                setattr(cfg, name, a.ctypes.data)
        cfg.lambda_, cfg.bound_scaling, cfg.dT = lam, bound_scaling, dT
        cfg.postural_level, cfg.joint_limit_set, cfg.velocity_limit_set = postural_level, joint_limit_set, velocity_limit_set
        arr = (K.KinTaskC * max(1, len(tasks)))()
        for i, t in enumerate(tasks):
            arr[i].kind = K.COM if t["kind"] == "com" else K.CARTESIAN
            link = lambda v: -1 if v is None else (self._kin_tree.link(v) if isinstance(v, str) else int(v))  # noqa: E731
            arr[i].distal, arr[i].base = link(t.get("distal")), link(t.get("base"))
            arr[i].level, arr[i].row0 = t["level"], t.get("row0", 0)
            arr[i].lambda_, arr[i].orientation_gain = t.get("lam", 1.0), t.get("orientation_gain", 1.0)
            for key in ("desired", "desired_vel"):
                if t.get(key) is not None:
                    a = np.ascontiguousarray(t[key], dtype=np.float64); keep.append(a)
                    setattr(arr[i], key, a.ctypes.data)
        cfg.n_tasks, cfg.tasks = len(tasks), C.cast(arr, C.c_void_p)
        _check(lib().osb_states_configure(self._h, C.byref(cfg)))

    def solve_states_host(self, q, x=None, status=None):
        """q: float64 host array [B, n] (joint positions). H2D q -> producers -> solve -> D2H. Returns (x, status)."""
        q = np.ascontiguousarray(q, dtype=np.float64)
        B = q.shape[0]
        if x is None:
            x = np.empty((B, self.n), dtype=np.float64)
        if status is None:
            status = np.empty(B, dtype=np.int32)
        _check(lib().osb_solve_states_host(self._h, q.ctypes.data, B, x.ctypes.data, status.ctypes.data))
        return x, status

    def states_download_blobs(self, B):
        blobs = np.empty((B, self.layout.stride), dtype=np.float64)
        _check(lib().osb_states_download_blobs(self._h, blobs.ctypes.data, B))
        return blobs

    def produce_posture_device(self, d_blobs, q, q_ref=None, v_ref=None, q_min=None, q_max=None, q_neutral=None,
                               qdot_max=None, lam=1.0, bound_scaling=1.0, dT=1.0, postural_level=-1,
                               joint_limit_set=-1, velocity_limit_set=-1, stream=None):
        """Postural error and JointLimits / VelocityLimits boxes from joint states, on device
        (Postural.cpp:97-100, JointLimits.cpp:37-59, VelocityLimits.cpp:72-89). q, q_ref, v_ref:
        torch CUDA fp64 tensors [B, n] or [n]; limit vectors [n]."""
        import torch
        if stream is None:
            stream = torch.cuda.current_stream(self.device)
        P = PostureInputs()

        def ptr(t):
            if t is None:
                return None, 0
            assert t.dtype == torch.float64 and t.is_cuda and t.is_contiguous()
            return t.data_ptr(), (t.stride(0) if t.dim() == 2 else 0)
        P.q, P.q_stride = ptr(q)
        P.q_ref, P.q_ref_stride = ptr(q_ref)
        P.v_ref, P.v_ref_stride = ptr(v_ref)
        P.q_min, P.q_max, P.q_neutral, P.qdot_max = ptr(q_min)[0], ptr(q_max)[0], ptr(q_neutral)[0], ptr(qdot_max)[0]
        P.lambda_, P.bound_scaling, P.dT = lam, bound_scaling, dT
        P.postural_level, P.joint_limit_set, P.velocity_limit_set = postural_level, joint_limit_set, velocity_limit_set
        _check(lib().osb_produce_posture_device(self._h, C.byref(P), d_blobs.data_ptr(), d_blobs.shape[0],
                                                stream.cuda_stream))
