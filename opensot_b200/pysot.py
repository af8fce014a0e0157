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
(TaskToConstraint.cpp:61-66); `task << constraint` attaches the constraint to the TASK: it holds
at that task's level only (AutoStack.cpp:291-303, 320-325; an aggregate inherits the constraints
of its children, tasks/Aggregated.cpp:134-200), a bound given that way becomes identity rows of
the level's local block; `W * task` with a matrix gives a dense weight (AutoStack.h:127);
`stack.setRegularisationTask(t)` adds t's cost to every level and never to the optimality
conditions (AutoStack.h:78-90, iHQP.cpp:184-188, 274-278); solve() returns False if any instance
failed (iHQP.cpp:279,335-347). Weights given as numbers / numpy arrays are per stack (one W for
the whole batch, like one reference task has one W); given as TORCH tensors with a leading batch
axis they are per instance: [B] a scalar per instance (`w * task`: a policy that outputs its task
weights), [B, m] a diagonal, [B, m, m] a dense matrix per instance; they are copied into the blobs
on the device at every solve (nothing crosses PCIe).
"""
import numpy as np

from . import _abi
from .engine import BatchedStackSolver, blob_layout

DEFAULT_EPS_REGULARISATION = 2e2   # include/OpenSoT/solvers/iHQP.h:32


def _t(x):
    """fp64 tensors; they must live on the solver's GPU by the time iHQP is built (checked there)"""
    import torch
    assert isinstance(x, torch.Tensor) and x.dtype == torch.float64, "fp64 torch tensors only"
    return x.contiguous()


class Task:
    """A [B, m, n], b [B, m]; weight: None (identity), a diagonal [m] or a dense [m, m] array (per
    stack), or a torch tensor [B, m] / [B, m, m] (per instance)."""
    __array_ufunc__ = None       # `ndarray * task` goes to Task.__rmul__, not to numpy broadcasting

    def __init__(self, task_id, A, b, weight=None, rows=None, parts=None):
        self.task_id = task_id
        self._parts = parts          # aggregated: list of Task
        if parts is None:
            self.A, self.b = _t(A), _t(b)
            assert self.A.dim() == 3 and self.b.dim() == 2 and self.A.shape[:2] == self.b.shape
            self.rows = None if rows is None else [int(r) for r in rows]
            m = self.A.shape[1] if self.rows is None else len(self.rows)
            self.weight = _as_weight(weight, m)
        self.identity = False
        self.local = []              # task-local constraints (Task.h: getConstraints())
        self.active = True

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

    def setActive(self, flag):
        """an inactive task keeps its rows in the stack with A = 0 and b unchanged (Task.h:382-386)"""
        for l in self.leaves():
            l.active = bool(flag)

    def getActive(self):
        return all(l.active for l in self.leaves())

    def getA(self):
        assert self._parts is None
        return self.A

    def getb(self):
        assert self._parts is None
        return self.b

    def setWeight(self, w):
        assert self._parts is None, "set the weight of the children (or use w * aggregate)"
        self.weight = _as_weight(w, self.getTaskSize())

    def getWeight(self):
        """diagonal as a vector while every part is diagonal, else the block-diagonal matrix
        (tasks/Aggregated.cpp:265-279)"""
        if self._parts:
            ws = [p.getWeight() for p in self._parts]
            if any(_batched(w) for w in ws):
                return _block_diag_batched(ws)
            if all(w.ndim == 1 for w in ws):
                return np.concatenate(ws)
            m = int(np.sum([w.shape[0] for w in ws]))
            W, at = np.zeros((m, m)), 0
            for w in ws:
                k = w.shape[0]
                W[at:at + k, at:at + k] = np.diag(w) if w.ndim == 1 else w
                at += k
            return W
        return np.ones(self.getTaskSize()) if self.weight is None else self.weight

    def leaves(self):
        return [l for p in self._parts for l in p.leaves()] if self._parts else [self]

    def has_weight(self):
        return any(l.weight is not None for l in self.leaves())

    def weight_kind(self):
        if not self.has_weight():
            return _abi.W_IDENTITY
        return _abi.W_DENSE if any(l.weight is not None and _is_dense(l.weight) for l in self.leaves()) else _abi.W_DIAG

    def getConstraints(self):
        """own constraints, then the children's (tasks/Aggregated.cpp:134-200)"""
        out = list(self.local)
        for p in (self._parts or []):
            out += p.getConstraints()
        return out

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
        self.local.append(constraint)        # AutoStack.cpp:291-303, 320-325: the task is returned
        return self


def _batched(w):
    """a per-instance weight: a torch tensor with a leading batch axis"""
    import torch
    return isinstance(w, torch.Tensor)


def _is_dense(w):
    return (w.dim() == 3) if _batched(w) else (w.ndim == 2)


def _block_diag_batched(ws):
    """weights of the parts of an aggregate, some of them per instance -> [B, M] while every part is diagonal, else the
    block-diagonal [B, M, M] (tasks/Aggregated.cpp:265-279, with a batch axis)"""
    import torch
    ref = next(w for w in ws if _batched(w))
    B, dev = ref.shape[0], ref.device
    tw = [w if _batched(w) else torch.from_numpy(np.ascontiguousarray(w)).to(dev).unsqueeze(0).expand(B, *w.shape) for w in ws]
    if all(w.dim() == 2 for w in tw):
        return torch.cat(tw, dim=1).contiguous()
    M = int(np.sum([w.shape[1] for w in tw]))
    W, at = torch.zeros((B, M, M), dtype=torch.float64, device=dev), 0
    for w in tw:
        k = w.shape[1]
        W[:, at:at + k, at:at + k] = torch.diag_embed(w) if w.dim() == 2 else w
        at += k
    return W


def _as_weight(w, m):
    if w is None:
        return None
    if _batched(w):
        import torch
        assert w.dtype == torch.float64 and w.dim() in (2, 3) and tuple(w.shape[1:]) in ((m,), (m, m)), \
            "per-instance weight: fp64 torch tensor [B, %d] or [B, %d, %d]" % (m, m, m)
        return w.contiguous()
    w = np.asarray(w, dtype=np.float64)
    if w.ndim == 2:
        assert w.shape == (m, m), "weight must be %d x %d" % (m, m)
        if np.count_nonzero(w - np.diag(np.diagonal(w))) == 0:
            return np.diagonal(w).copy()
        return w.copy()
    return np.broadcast_to(w, (m,)).copy()


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
        self.local = []
        self.active = True

    def getTaskSize(self):
        return self.b.shape[1]

    def getXSize(self):
        return self.b.shape[1]


class Constraint:
    """Bounds: lower/upper [B, n] (or [n]); general rows: A [B, c, n] with bLower/bUpper [B, c].

    max_rows: row SLOTS of a constraint whose number of rows changes at run time or differs between instances
    (constraints/velocity/ConvexHull.cpp:47-52: one row per hull edge; the reference rebuilds its QP object,
    QPOasesBackEnd.cpp:229-244). The engine has fixed shapes: the tensors are padded to max_rows with the null row
    -inf <= 0 x <= +inf (the convention of CollisionAvoidance.cpp:104-107), setRows() refills them in place on the
    device, a per-instance count masks the slots an instance does not use (same scheme as BatchediHQP's row slots)."""

    def __init__(self, constraint_id, upper, lower, A=None, max_rows=None):
        self.constraint_id = constraint_id
        self.upper, self.lower = _t(upper), _t(lower)
        self.A = None if A is None else _t(A)
        if max_rows is not None:
            assert self.A is not None and self.A.dim() == 3 and self.upper.dim() == 2, "row slots are for general rows [B, c, n]"
            A0, u0, l0 = self.A, self.upper, self.lower
            B, c, n = A0.shape
            assert c <= max_rows
            import torch
            self.A = torch.zeros((B, max_rows, n), dtype=torch.float64, device=A0.device)
            self.upper = torch.full((B, max_rows), float("inf"), dtype=torch.float64, device=A0.device)
            self.lower = torch.full((B, max_rows), float("-inf"), dtype=torch.float64, device=A0.device)
            self.setRows(A0, u0, l0)

    def isBound(self):
        return self.A is None

    def setRows(self, A, upper, lower, count=None):
        """new rows for the next solves, written into the slots in place: A [B, c, n], upper / lower [B, c], c <= max_rows;
        count: how many of the c rows are real -- an int, or a [B] integer tensor (per instance, stays on the device);
        every other slot holds the null row"""
        import torch
        A, upper, lower = _t(A), _t(upper), _t(lower)
        c, K = A.shape[1], self.A.shape[1]
        assert self.A.dim() == 3 and c <= K and A.shape[0] == self.A.shape[0] and A.shape[2] == self.A.shape[2], \
            "setRows: [B, c <= %d, n] rows" % K
        self.A[:, :c] = A; self.upper[:, :c] = upper; self.lower[:, :c] = lower
        self.A[:, c:] = 0.0; self.upper[:, c:] = float("inf"); self.lower[:, c:] = float("-inf")
        if count is not None:
            cnt = count if isinstance(count, torch.Tensor) else torch.full((A.shape[0],), int(count), device=self.A.device)
            unused = torch.arange(K, device=self.A.device).unsqueeze(0) >= cnt.to(self.A.device).reshape(-1, 1)
            self.A.masked_fill_(unused.unsqueeze(2), 0.0)
            self.upper.masked_fill_(unused, float("inf")); self.lower.masked_fill_(unused, float("-inf"))


def GenericTask(task_id, A, b):          # bindings/python/tasks: GenericTask(id, A, b)
    return Task(task_id, A, b)


def GenericConstraint(constraint_id, upper, lower, A=None, max_rows=None):
    return Constraint(constraint_id, upper, lower, A, max_rows)


class AutoStack:
    def __init__(self, first):
        self.levels = [first] if isinstance(first, Task) else list(first)
        self.bounds, self.constraints, self.equalities = [], [], []
        self.regularisation = None

    def setRegularisationTask(self, task):      # AutoStack.h:78-85
        self.regularisation = task

    def getRegularisationTask(self):
        return self.regularisation

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
    out.regularisation = sa.regularisation or sb.regularisation
    return out


def sub(task, rows):
    assert task._parts is None and not task.identity, "sub-tasks of plain tasks only"
    rows = list(rows)
    w = task.weight
    if w is not None and _batched(w):
        w = w[:, rows] if w.dim() == 2 else w[:, rows][:, :, rows]
    elif w is not None:
        w = w[rows] if w.ndim == 1 else w[np.ix_(rows, rows)]      # SubTask.cpp:39-58
    out = Task(task.task_id + "_sub", task.A, task.b, weight=w, rows=rows)
    out.active = task.active
    return out


def mul(w, task):
    if not task._parts:
        return _scaled(task, w)
    assert (w.dim() == 1) if _batched(w) else (np.ndim(w) == 0), "a matrix weight applies to a plain task"
    out = Task(task.task_id, None, None, parts=[_scaled(l, w) for l in task.leaves()])
    out.local = task.getConstraints()
    for l in out._parts:
        l.local = []
    return out


def _scaled_batched(old, w, m):
    """w * W with a batch axis on either side: w a torch [B] (scalar per instance), [B, m] or [B, m, m], or a per-stack
    number / array against a per-instance W"""
    import torch
    ref = w if _batched(w) else old
    B, dev = ref.shape[0], ref.device

    def lift(x):
        if _batched(x):
            return x
        x = torch.from_numpy(np.ascontiguousarray(np.asarray(x, dtype=np.float64))).to(dev)
        return x.unsqueeze(0).expand(B, *x.shape) if x.dim() else x
    o, ww = lift(old), lift(w)
    if _batched(w) and w.dim() == 1:
        ww = w.reshape(B, *([1] * (o.dim() - 1)))          # a scalar per instance
        return (o * ww).contiguous()
    if ww.dim() == 3:                                       # W * task: W @ old, per instance (AutoStack.h:127)
        return torch.matmul(ww, torch.diag_embed(o) if o.dim() == 2 else o).contiguous()
    if ww.dim() == 0:
        return (o * ww).contiguous()
    return (o * (ww if o.dim() == 2 else ww.unsqueeze(2))).contiguous()     # diagonal w: scales the rows of W, like diag(w) @ W


def _scaled(leaf, w):
    old = leaf.getWeight()
    if _batched(w) or _batched(old):
        neww = _scaled_batched(old, w, leaf.getTaskSize())
        if leaf.identity:
            new = IdentityTask(leaf.task_id, leaf.b)
            new.weight = _as_weight(neww, leaf.getTaskSize())
        else:
            new = Task(leaf.task_id, leaf.A, leaf.b, weight=neww, rows=leaf.rows)
        new.local = list(leaf.local)
        new.active = leaf.active
        return new
    w = np.asarray(w, dtype=np.float64)
    if w.ndim == 2:                          # W * task: setWeight(W * task->getWeight()), AutoStack.h:127
        neww = w @ (np.diag(old) if old.ndim == 1 else old)
    else:
        neww = old * w
    if leaf.identity:
        new = IdentityTask(leaf.task_id, leaf.b)
        new.weight = _as_weight(neww, leaf.getTaskSize())
    else:
        new = Task(leaf.task_id, leaf.A, leaf.b, weight=neww, rows=leaf.rows)
    new.local = list(leaf.local)
    new.active = leaf.active
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

    def __init__(self, stack, eps_regularisation=DEFAULT_EPS_REGULARISATION, be_solver=None, device=0, plan_only=False):
        """plan_only: derive the stack descriptor, the blob layout and the piling plan without a
        GPU (no engine is created, solve() is unavailable) -- what the CPU test-suite checks."""
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
            c_loc = int(np.sum([self._local_rows(c) for c in t.getConstraints()])) if t.getConstraints() else 0
            if t.identity:
                levels.append(dict(m=self.n, kind=_abi.TASK_IDENTITY, wkind=t.weight_kind(), c_loc=c_loc))
            else:
                levels.append(dict(m=t.getTaskSize(), kind=_abi.TASK_DENSE, wkind=t.weight_kind(), c_loc=c_loc))
        reg = self.stack.regularisation
        m_reg, reg_wkind = 0, _abi.W_IDENTITY
        if reg is not None:
            m_reg, reg_wkind = reg.getTaskSize(), reg.weight_kind()
            if reg_wkind == _abi.W_DENSE:
                raise RuntimeError("regularisation task: identity or diagonal weight only")
        c_eq = int(np.sum([t.getTaskSize() for t in self.stack.equalities])) if self.stack.equalities else 0
        c_in = int(np.sum([c.A.shape[1] for c in self.stack.constraints])) if self.stack.constraints else 0
        if len(self.stack.bounds) > _abi.OSB_MAX_BOUND_SETS:
            raise RuntimeError("Can Not initizalize SoT: more than %d bound sets" % _abi.OSB_MAX_BOUND_SETS)
        self.desc = _abi.make_desc(self.n, levels, c_eq=c_eq, c_in=c_in, n_bound_sets=len(self.stack.bounds),
                                   eps_factor=eps_regularisation, m_reg=m_reg, reg_weight_kind=reg_wkind)
        self.layout = blob_layout(self.desc)
        if plan_only:
            self.engine, dev = None, torch.device("cpu")
        else:
            dev = torch.device("cuda", device)
            assert first.b.device == dev, "task tensors must live on the solver's device (cuda:%d)" % device
            self.engine = BatchedStackSolver(self.desc, self.B, device=device)
        self.d_blobs = torch.zeros((self.B, self.layout.stride), dtype=torch.float64, device=dev)
        self.d_x = torch.empty((self.B, self.n), dtype=torch.float64, device=dev)
        self.d_status = torch.empty(self.B, dtype=torch.int32, device=dev)
        self._keep = []
        self._eye = None
        self._active = [True] * len(lv)

    # --- piling: tensors -> gather sources --------------------------------------------------
    def _rows_of(self, leaf):
        import torch
        if leaf.rows is None:
            return None, leaf.A.shape[1]
        idx = torch.tensor(leaf.rows, dtype=torch.int32, device=leaf.A.device)
        self._keep.append(idx)
        return idx, len(leaf.rows)

    def _identity(self):
        import torch
        if self._eye is None:
            self._eye = torch.eye(self.n, dtype=torch.float64, device=self.d_blobs.device).unsqueeze(0).expand(self.B, self.n, self.n)   # stride 0 over the batch
        return self._eye

    def _pile(self, leaf, off_A, off_b, n, off_b2=None):
        """rows of one plain task -> A block and right-hand side(s)"""
        srcs = []
        if leaf.identity:
            idx, m, A = None, n, self._identity()
        else:
            idx, m = self._rows_of(leaf)
            A = leaf.A
        srcs.append(dict(src=A, rows=m, cols=n, dst_off=off_A, row_index=idx, scale=1.0 if leaf.active else 0.0))
        for off in (off_b, off_b2):
            if off is not None:
                srcs.append(dict(src=leaf.b.unsqueeze(2), rows=m, cols=1, dst_off=off, row_index=idx))
        return srcs, m

    def _local_rows(self, c):
        if isinstance(c, Task):
            return c.getTaskSize()
        return self.n if c.isBound() else c.A.shape[1]

    def _pile_rows(self, c, off_C, off_l, off_u, row):
        """one constraint as two-sided rows lo <= C x <= hi at row `row` of a row block: general
        rows as they are, a bound as identity rows, a task as the equality A x = b"""
        n, srcs = self.n, []
        if isinstance(c, Task):
            for leaf in c.leaves():
                s, m = self._pile(leaf, off_C + row * n, off_l + row, n, off_b2=off_u + row)
                srcs += s
                row += m
            return srcs, row
        m = self._local_rows(c)
        srcs.append(dict(src=self._identity() if c.isBound() else c.A, rows=m, cols=n, dst_off=off_C + row * n))
        srcs.append(dict(src=self._bn(c.lower, m).unsqueeze(2), rows=m, cols=1, dst_off=off_l + row))
        srcs.append(dict(src=self._bn(c.upper, m).unsqueeze(2), rows=m, cols=1, dst_off=off_u + row))
        return srcs, row + m

    def _put_weight(self, off, W):
        import torch
        if _batched(W):                               # per instance: [B, m] or [B, m, m], already on the device
            assert W.shape[0] == self.B and W.device == self.d_blobs.device, "per-instance weights: [B, ...] on the solver's device"
            w = W.reshape(self.B, -1)
            self.d_blobs[:, off:off + w.shape[1]] = w
            return
        w = torch.from_numpy(np.ascontiguousarray(W).reshape(-1)).to(self.d_blobs.device)
        self.d_blobs[:, off:off + w.numel()] = w      # constant per solve: a plain device copy

    def _sources(self):
        import torch
        L, n, B = self.layout, self.n, self.B
        srcs, self._keep = [], []
        for k, t in enumerate(self.stack.levels):
            if t.identity:
                if not t.active:     # A = I is implicit in an identity level: nothing to zero
                    raise NotImplementedError("setActive(False) on an identity level: use GenericTask with A = I")
                srcs.append(dict(src=t.b.unsqueeze(2), rows=n, cols=1, dst_off=L.b[k]))
            else:
                row = 0
                for leaf in t.leaves():
                    s, m = self._pile(leaf, L.A[k] + row * n, L.b[k] + row, n)
                    srcs += s
                    row += m
            if t.has_weight():
                self._put_weight(L.w[k], t.getWeight())
            row = 0
            for c in t.getConstraints():
                s, row = self._pile_rows(c, L.Cl[k], L.lCl[k], L.uCl[k], row)
                srcs += s
        reg = self.stack.regularisation
        if reg is not None:
            row = 0
            for leaf in reg.leaves():
                s, m = self._pile(leaf, L.Ar + row * n, L.br + row, n)
                srcs += s
                row += m
            if reg.has_weight():
                self._put_weight(L.wr, reg.getWeight())
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
        if self.engine is None:
            raise RuntimeError("plan-only solver: no engine (there is no CPU solve path)")
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

    def setEpsRegularisation(self, eps, i=None):
        """the diagonal shift of every level's QP. As in the reference the value given HERE is
        absolute (QPOasesBackEnd.cpp:358-366 stores it as is), while the constructor's
        eps_regularisation is multiplied by 2.22e-13 (QPOasesBackEnd.cpp:57,67). The engine has one
        shift per stack: the per-level overload (iHQP.cpp:438-447) is not offered."""
        if i is not None:
            raise NotImplementedError("one eps per stack")
        if eps < 0.0:
            return False                      # "Negative eps is not allowed!"
        self.desc.eps = float(eps)
        if self.engine is not None:           # the shift is part of the engine's stack descriptor
            device = self.engine.device
            self.engine.close()
            self.engine = BatchedStackSolver(self.desc, self.B, device=device)
            self.engine.set_active_levels(self._active)
        return True

    def getBackEndName(self, i=0):
        return "B200:" + self.engine.kernel_name()
