"""Batched Math-of-Tasks front-end over CUDA tensors.

The caller side the reference offers through `pyopensot` (bindings/python/autostack.hpp:9-107,
bindings/python/solver.hpp:64-89, bindings/python/pyopensot.cpp:14-67), with a batch axis: a task
holds `A[B, m, n]`, `b[B, m]` as torch CUDA tensors, stacks are built with the same operators
(`+` aggregate, `/` hard priority, `%` sub-task rows, `*` weight, `<<` subject-to) and the same
free functions (`sum`, `hard`, `sub`, `mul`, `subj`), and `iHQP(stack).solve()` solves all B
instances in one launch. Per solve the tensors are piled into instance blobs ON THE DEVICE
(osb_stack_gather_device: the Aggregated/MatrixPiler copies of src/tasks/Aggregated.cpp:113-132
and src/constraints/Aggregated.cpp:73-257 become one gather kernel) and handed to
osb_solve_device; nothing crosses PCIe.

Semantics kept from the reference: rows of `t1 + t2` are piled in order and their
weights form a block-diagonal W; weights enter as A'WA (iHQP.cpp:152); `w * task` scales the
weight, not A or b; `task % rows` selects rows of A, b and of a diagonal weight
(SubTask.cpp:39-58, lambda of the sub-task = 1); bounds are intersected (l = max, u = min,
constraints/Aggregated.cpp:141-148); `stack << task` turns the task into the equality A x = b
(TaskToConstraint.cpp:61-66); solve() returns False if any instance failed (iHQP.cpp:279,335-347).
Not covered: dense (non-diagonal) weights, task-local constraints, regularisation tasks -- use
the C ABI directly for those.
"""
import numpy as np

from . import _abi
from .engine import BatchedStackSolver, blob_layout

DEFAULT_EPS_REGULARISATION = 2e2   # include/OpenSoT/solvers/iHQP.h:32


def _t(x):
    import torch
    assert isinstance(x, torch.Tensor) and x.is_cuda and x.dtype == torch.float64, "fp64 CUDA tensors only"
    return x.contiguous()


class Task:
    """A [B, m, n], b [B, m]; weight: None (identity) or a [m] / [B-less] diagonal as a python list/array."""

    def __init__(self, task_id, A, b, weight=None, rows=None, parts=None):
        self.task_id = task_id
        self._parts = parts          # aggregated: list of Task
        if parts is None:
            self.A, self.b = _t(A), _t(b)
            assert self.A.dim() == 3 and self.b.dim() == 2 and self.A.shape[:2] == self.b.shape
            self.rows = None if rows is None else [int(r) for r in rows]
            m = self.A.shape[1] if self.rows is None else len(self.rows)
            self.weight = None if weight is None else np.broadcast_to(np.asarray(weight, dtype=np.float64), (m,)).copy()
        self.identity = False

    # --- reference Task surface (include/OpenSoT/Task.h) -------------------------------------
    def getTaskID(self):
        return self.task_id

    def getTaskSize(self):
        if self._parts:
            return int(np.sum([p.getTaskSize() for p in self._parts]))
        return self.A.shape[1] if self.rows is None else len(self.rows)

    def getXSize(self):
        return self.leaves()[0].A.shape[2]

    def setA(self, A):
        self.A = _t(A)

    def setb(self, b):
        self.b = _t(b)

    def setAb(self, A, b):
        self.setA(A); self.setb(b)

    def setWeight(self, w):
        self.weight = np.broadcast_to(np.asarray(w, dtype=np.float64), (self.getTaskSize(),)).copy()

    def getWeight(self):
        if self._parts:
            return np.concatenate([p.getWeight() for p in self._parts])
        return np.ones(self.getTaskSize()) if self.weight is None else self.weight

    def leaves(self):
        return [l for p in self._parts for l in p.leaves()] if self._parts else [self]

    def has_weight(self):
        return any(l.weight is not None for l in self.leaves())

    # --- Math of Tasks (include/OpenSoT/utils/AutoStack.h:115-336) ----------------------------
    def __add__(self, other):
        return sum_(self, other)

    def __truediv__(self, other):
        return hard(self, other)

    def __mod__(self, rows):
        return sub(self, rows)

    def __rmul__(self, w):
        return mul(w, self)

    def __lshift__(self, constraint):
        return subj(AutoStack(self), constraint)


class IdentityTask(Task):
    """A = I (Postural-like, src/tasks/velocity/Postural.cpp): only b [B, n] is stored."""

    def __init__(self, task_id, b):
        self.task_id = task_id
        self._parts = None
        self.b = _t(b)
        self.A = None
        self.rows = None
        self.weight = None
        self.identity = True

    def getTaskSize(self):
        return self.b.shape[1]

    def getXSize(self):
        return self.b.shape[1]


class Constraint:
    """Bounds: lower/upper [B, n] (or [n]); general rows: A [B, c, n] with bLower/bUpper [B, c]."""

    def __init__(self, constraint_id, upper, lower, A=None):
        self.constraint_id = constraint_id
        self.upper, self.lower = _t(upper), _t(lower)
        self.A = None if A is None else _t(A)

    def isBound(self):
        return self.A is None


def GenericTask(task_id, A, b):          # bindings/python/tasks: GenericTask(id, A, b)
    return Task(task_id, A, b)


def GenericConstraint(constraint_id, upper, lower, A=None):
    return Constraint(constraint_id, upper, lower, A)


class AutoStack:
    def __init__(self, first):
        self.levels = [first] if isinstance(first, Task) else list(first)
        self.bounds, self.constraints, self.equalities = [], [], []

    def getStack(self):
        return self.levels

    def getBoundsList(self):
        return self.bounds + self.constraints

    def update(self):
        """tensors are referenced, not copied: nothing to do (AutoStack.cpp:385-393 calls the tasks' update)"""

    def __truediv__(self, other):
        return hard(self, other)

    def __lshift__(self, c):
        return subj(self, c)


def sum_(t1, t2):
    return Task(t1.task_id + "+" + t2.task_id, None, None, parts=[t1, t2])


def hard(a, b):
    sa = a if isinstance(a, AutoStack) else AutoStack(a)
    sb = b if isinstance(b, AutoStack) else AutoStack(b)
    out = AutoStack(sa.levels + sb.levels)
    out.bounds, out.constraints, out.equalities = sa.bounds + sb.bounds, sa.constraints + sb.constraints, sa.equalities + sb.equalities
    return out


def sub(task, rows):
    assert task._parts is None and not task.identity, "sub-tasks of plain tasks only"
    w = None if task.weight is None else task.weight[list(rows)]
    return Task(task.task_id + "_sub", task.A, task.b, weight=w, rows=rows)


def mul(w, task):
    out = Task(task.task_id, None, None, parts=[_scaled(l, w) for l in task.leaves()]) if task._parts else _scaled(task, w)
    return out


def _scaled(leaf, w):
    assert not leaf.identity, "weighted identity levels need the C ABI (OSB_W_DIAG on an identity task)"
    new = Task(leaf.task_id, leaf.A, leaf.b, weight=leaf.getWeight() * np.asarray(w, dtype=np.float64), rows=leaf.rows)
    return new


def subj(stack, c):
    stack = stack if isinstance(stack, AutoStack) else AutoStack(stack)
    if isinstance(c, Task):
        stack.equalities.append(c)          # TaskToConstraint: A x = b
    elif c.isBound():
        stack.bounds.append(c)
    else:
        stack.constraints.append(c)
    return stack


globals()["sum"] = sum_       # pyopensot exposes `sum` (autostack.hpp:57-71)


class iHQP:
    """Batched iHQP over an AutoStack of tensor tasks (bindings/python/solver.hpp:64-89)."""

    def __init__(self, stack, eps_regularisation=DEFAULT_EPS_REGULARISATION, be_solver=None, device=0):
        import torch
        self.stack = stack if isinstance(stack, AutoStack) else AutoStack(stack)
        lv = self.stack.levels
        if not lv:
            raise RuntimeError("Can Not initizalize SoT: empty stack")       # iHQP.cpp:24
        first = lv[0].leaves()[0]
        self.B = first.b.shape[0]
        self.n = lv[0].getXSize()
        levels = []
        for t in lv:
            if t.identity:
                levels.append(dict(m=self.n, kind=_abi.TASK_IDENTITY))
            else:
                levels.append(dict(m=t.getTaskSize(), kind=_abi.TASK_DENSE,
                                   wkind=_abi.W_DIAG if t.has_weight() else _abi.W_IDENTITY))
        c_eq = int(np.sum([t.getTaskSize() for t in self.stack.equalities])) if self.stack.equalities else 0
        c_in = int(np.sum([c.A.shape[1] for c in self.stack.constraints])) if self.stack.constraints else 0
        if len(self.stack.bounds) > _abi.OSB_MAX_BOUND_SETS:
            raise RuntimeError("Can Not initizalize SoT: more than %d bound sets" % _abi.OSB_MAX_BOUND_SETS)
        self.desc = _abi.make_desc(self.n, levels, c_eq=c_eq, c_in=c_in, n_bound_sets=len(self.stack.bounds),
                                   eps_factor=eps_regularisation)
        self.layout = blob_layout(self.desc)
        self.engine = BatchedStackSolver(self.desc, self.B, device=device)
        dev = torch.device("cuda", device)
        self.d_blobs = torch.zeros((self.B, self.layout.stride), dtype=torch.float64, device=dev)
        self.d_x = torch.empty((self.B, self.n), dtype=torch.float64, device=dev)
        self.d_status = torch.empty(self.B, dtype=torch.int32, device=dev)
        self._keep = []
        self._active = [True] * len(lv)

    # --- piling: tensors -> gather sources --------------------------------------------------
    def _rows_of(self, leaf):
        import torch
        if leaf.rows is None:
            return None, leaf.A.shape[1]
        idx = torch.tensor(leaf.rows, dtype=torch.int32, device=leaf.A.device)
        self._keep.append(idx)
        return idx, len(leaf.rows)

    def _pile(self, leaf, off_A, off_b, n):
        srcs = []
        idx, m = self._rows_of(leaf)
        srcs.append(dict(src=leaf.A, rows=m, cols=n, dst_off=off_A, row_index=idx))
        srcs.append(dict(src=leaf.b.unsqueeze(2), rows=m, cols=1, dst_off=off_b, row_index=idx))
        return srcs, m

    def _sources(self):
        import torch
        L, n, B = self.layout, self.n, self.B
        srcs, self._keep = [], []
        for k, t in enumerate(self.stack.levels):
            if t.identity:
                srcs.append(dict(src=t.b.unsqueeze(2), rows=n, cols=1, dst_off=L.b[k]))
                continue
            row = 0
            for leaf in t.leaves():
                s, m = self._pile(leaf, L.A[k] + row * n, L.b[k] + row, n)
                srcs += s
                row += m
            if t.has_weight():
                w = torch.from_numpy(np.ascontiguousarray(t.getWeight())).to(self.d_blobs.device)
                self.d_blobs[:, L.w[k]:L.w[k] + w.numel()] = w      # constant per solve: a plain device copy
        row = 0
        for t in self.stack.equalities:
            for leaf in t.leaves():
                s, m = self._pile(leaf, L.Ceq + row * n, L.deq + row, n)
                srcs += s
                row += m
        row = 0
        for c in self.stack.constraints:
            m = c.A.shape[1]
            srcs.append(dict(src=c.A, rows=m, cols=n, dst_off=L.Cin + row * n))
            srcs.append(dict(src=self._bn(c.lower, m).unsqueeze(2), rows=m, cols=1, dst_off=L.lA + row))
            srcs.append(dict(src=self._bn(c.upper, m).unsqueeze(2), rows=m, cols=1, dst_off=L.uA + row))
            row += m
        for s_, c in enumerate(self.stack.bounds):
            srcs.append(dict(src=self._bn(c.lower, n).unsqueeze(2), rows=n, cols=1, dst_off=L.l[s_]))
            srcs.append(dict(src=self._bn(c.upper, n).unsqueeze(2), rows=n, cols=1, dst_off=L.u[s_]))
        return srcs

    def _bn(self, v, m):
        if v.dim() == 1:
            v = v.unsqueeze(0).expand(self.B, m).contiguous()
            self._keep.append(v)
        return v

    # --- reference iHQP surface -------------------------------------------------------------
    def solve(self):
        """returns (ok, x [B, n] CUDA tensor); per-instance osb_status words in getStatus()"""
        import torch
        self.engine.gather_device(self._sources(), self.d_blobs)
        self.engine.solve_device(self.d_blobs, self.d_x, self.d_status)
        ok = bool((self.d_status == 0).all().item())
        return ok, self.d_x

    def getStatus(self):
        return self.d_status

    def getNumberOfTasks(self):
        return len(self.stack.levels)

    def setActiveStack(self, i, flag):
        if i < 0 or i >= len(self._active):
            return False
        self._active[i] = bool(flag)
        self.engine.set_active_levels(self._active)
        return True

    def activateAllStacks(self):
        self._active = [True] * len(self._active)
        self.engine.activate_all_levels()

    def getBackEndName(self, i=0):
        return "B200:" + self.engine.kernel_name()
