"""Truncation strategies (src/interface/truncation.jl:37-275, src/implementations/truncation.jl:45-174).
Single matrix: the values vector is short (k reals) and the index search runs on a host copy, like the
reference's GPU path (MatrixAlgebraKitCUDAExt.jl:64-66).  Batches (``trunc_select_batched_``): rank and
truncation error of every block from one kernel launch and one device->host read."""
import ctypes as C
from dataclasses import dataclass

import numpy as np
import torch


@dataclass(frozen=True)
class NoTruncation:
    pass


@dataclass(frozen=True)
class TruncationByOrder:
    howmany: int
    rev: bool = True


@dataclass(frozen=True)
class TruncationByValue:
    atol: float = 0.0
    rtol: float = 0.0
    p: float = 2
    keep_below: bool = False


@dataclass(frozen=True)
class TruncationByError:
    atol: float = 0.0
    rtol: float = 0.0
    p: float = 2


@dataclass(frozen=True)
class TruncationIntersection:
    components: tuple

    def __and__(self, other):
        return TruncationIntersection(self.components + (other,))


@dataclass(frozen=True)
class TruncationUnion:
    components: tuple


def notrunc():
    return NoTruncation()


def truncrank(howmany, rev=True):
    return TruncationByOrder(int(howmany), rev)


def trunctol(atol=0.0, rtol=0.0, p=2, keep_below=False):
    return TruncationByValue(float(atol), float(rtol), p, keep_below)


def truncerror(atol=0.0, rtol=0.0, p=2):
    return TruncationByError(float(atol), float(rtol), p)


def trunc_and(*c):
    return TruncationIntersection(tuple(c))


def trunc_or(*c):
    if any(isinstance(x, NoTruncation) for x in c):
        return NoTruncation()  # notrunc is absorbing for | (test/common/truncate.jl:88-89)
    return TruncationUnion(tuple(c))


def select_truncation(trunc):
    """``select_truncation`` / ``TruncationStrategy(; atol, rtol, maxrank, minrank, maxerror)``
    (src/interface/truncation.jl:37-66)."""
    if trunc is None:
        return NoTruncation()
    if isinstance(trunc, dict):
        allowed = {"atol", "rtol", "maxrank", "minrank", "maxerror"}
        bad = set(trunc) - allowed
        if bad:
            raise ValueError(f"unknown truncation keyword(s) {sorted(bad)}")
        comps = []
        if trunc.get("atol") is not None or trunc.get("rtol") is not None:
            comps.append(trunctol(trunc.get("atol") or 0.0, trunc.get("rtol") or 0.0))
        if trunc.get("maxrank") is not None:
            comps.append(truncrank(trunc["maxrank"]))
        if trunc.get("maxerror") is not None:
            comps.append(truncerror(atol=trunc["maxerror"]))
        s = NoTruncation() if not comps else (comps[0] if len(comps) == 1 else trunc_and(*comps))
        if trunc.get("minrank") is not None:
            s = truncrank(trunc["minrank"]) if not comps else trunc_or(s, truncrank(trunc["minrank"]))
        return s
    return trunc


def _pnorm(v, p):
    v = np.abs(v)
    if p == 2:
        return float(np.sqrt(np.sum(v * v)))
    if np.isinf(p):
        return float(v.max()) if v.size else 0.0
    return float(np.sum(v ** p) ** (1.0 / p))


def _truncerr_rank(vals_desc, s):
    vp = np.abs(vals_desc) ** s.p
    Np = float(vp.sum())
    ep = max(s.atol ** s.p, s.rtol ** s.p * Np)
    if ep >= Np:
        return 0
    cs = np.cumsum(vp[::-1])
    return len(vals_desc) - int(np.argmax(cs >= ep))


def _find(values, s, svd):
    n = len(values)
    if isinstance(s, NoTruncation):
        return np.arange(n)
    if isinstance(s, TruncationByOrder):
        hm = min(s.howmany, n)
        if svd:
            return np.arange(hm) if s.rev else np.arange(n - hm, n)
        order = np.argsort(-np.abs(values) if s.rev else np.abs(values), kind="stable")
        return order[:hm]
    if isinstance(s, TruncationByValue):
        thr = max(s.atol, s.rtol * _pnorm(values, s.p))
        if svd:
            if s.keep_below:
                i = int(np.searchsorted(-np.abs(values), -thr, side="left"))
                return np.arange(i, n)
            i = int(np.searchsorted(-np.abs(values), -thr, side="right"))
            return np.arange(i)
        m = (np.abs(values) <= thr) if s.keep_below else (np.abs(values) >= thr)
        return np.nonzero(m)[0]
    if isinstance(s, TruncationByError):
        if svd:
            return np.arange(_truncerr_rank(values, s))
        order = np.argsort(-np.abs(values), kind="stable")
        return order[: _truncerr_rank(values[order], s)]
    if isinstance(s, (TruncationIntersection, TruncationUnion)):
        # the reference's _ind_intersect / _ind_union (implementations/truncation.jl:104-164) keep the order of the FIRST
        # component: intersection = filter(in(rest), first); union = first, then the new entries of the rest in their order
        # (Julia's intersect/union).  eigh_trunc with truncrank(r) & trunctol(...) therefore returns descending-|lambda| order.
        inds = [np.asarray(_find(values, c, svd), dtype=np.int64) for c in s.components]
        if not inds:
            return np.arange(n) if isinstance(s, TruncationIntersection) else np.arange(0)
        out = inds[-1]
        for first in reversed(inds[:-1]):        # right fold, as the reference recurses on Base.tail
            rest = set(int(i) for i in out)
            if isinstance(s, TruncationIntersection):
                first_is_range = len(first) == 0 or np.array_equal(first, np.arange(first[0], first[0] + len(first)))
                out_is_range = len(out) == 0 or np.array_equal(out, np.arange(out[0], out[0] + len(out)))
                if first_is_range and not out_is_range:
                    # _ind_intersect(::UnitRange, ::Vector) = filter(in(range), vector): the vector's order survives
                    fs = set(int(i) for i in first)
                    out = np.array([i for i in out if int(i) in fs], dtype=np.int64)
                else:
                    out = np.array([i for i in first if int(i) in rest], dtype=np.int64)
            else:
                seen = set(int(i) for i in first)
                out = np.concatenate([first, np.array([i for i in out if int(i) not in seen], dtype=np.int64)])
        return out
    raise ValueError(f"unknown truncation strategy {s}")


def _host(values):
    return values.detach().cpu().numpy() if isinstance(values, torch.Tensor) else np.asarray(values)


def findtruncated(values, strategy):
    """generic ``findtruncated`` (implementations/truncation.jl:48-83); 0-based indices."""
    ind = _find(_host(values).astype(np.float64), strategy, svd=False)
    return torch.as_tensor(ind, dtype=torch.long, device=values.device if isinstance(values, torch.Tensor) else "cpu")


def findtruncated_svd(values, strategy):
    """``findtruncated_svd`` (implementations/truncation.jl:54-58,69-79,86-102): assumes the
    values are sorted descending."""
    ind = _find(_host(values).astype(np.float64), strategy, svd=True)
    return torch.as_tensor(ind, dtype=torch.long, device=values.device if isinstance(values, torch.Tensor) else "cpu")


def truncation_error_(values, ind):
    """``truncation_error!`` (implementations/truncation.jl:168-174): zero the kept entries
    in place, return the 2-norm of the rest (a device scalar read)."""
    values[ind] = 0.0
    return float(torch.linalg.vector_norm(values).item())


# ---------------------------------------------------------------------------------------------------
# batched, device side (makb200_trunc_select_batched)
# ---------------------------------------------------------------------------------------------------
class TruncSpec(C.Structure):
    """``makb200_trunc_spec`` (include/makb200.h)."""
    _fields_ = [("maxrank", C.c_int), ("minrank", C.c_int), ("by_value", C.c_int), ("by_error", C.c_int),
                ("vatol", C.c_double), ("vrtol", C.c_double), ("vp", C.c_double),
                ("eatol", C.c_double), ("ertol", C.c_double), ("ep", C.c_double)]


def _spec_and(spec, s):
    """Fold one prefix-keeping component into ``spec``; False if it cannot be encoded."""
    if isinstance(s, NoTruncation):
        return True
    if isinstance(s, TruncationByOrder) and s.rev:
        spec.maxrank = s.howmany if spec.maxrank < 0 else min(spec.maxrank, s.howmany)
        return True
    if isinstance(s, TruncationByValue) and not s.keep_below and not spec.by_value and np.isfinite(s.p) and s.p > 0:
        spec.by_value, spec.vatol, spec.vrtol, spec.vp = 1, s.atol, s.rtol, float(s.p)
        return True
    if isinstance(s, TruncationByError) and not spec.by_error and np.isfinite(s.p) and s.p > 0:
        spec.by_error, spec.eatol, spec.ertol, spec.ep = 1, s.atol, s.rtol, float(s.p)
        return True
    if isinstance(s, TruncationIntersection):
        return all(_spec_and(spec, c) for c in s.components)
    return False


def device_spec(strategy):
    """``makb200_trunc_spec`` of a strategy of the shape ``TruncationStrategy(; atol, rtol, maxrank,
    maxerror, minrank)`` builds — (rank & tol & error) | minrank — or None when the strategy keeps
    something other than a prefix of the sorted values (``rev=false``, ``keep_below``) or nests deeper."""
    spec = TruncSpec(-1, -1, 0, 0, 0.0, 0.0, 2.0, 0.0, 0.0, 2.0)
    s = strategy
    if isinstance(s, TruncationUnion):
        mins = [c for c in s.components if isinstance(c, TruncationByOrder) and c.rev]
        rest = [c for c in s.components if not (isinstance(c, TruncationByOrder) and c.rev)]
        if len(rest) != 1 or not mins:
            return None
        spec.minrank = max(c.howmany for c in mins)
        s = rest[0]
        if isinstance(s, NoTruncation):
            return None
        before = (spec.maxrank, spec.by_value, spec.by_error)
        if not _spec_and(spec, s) or (spec.maxrank, spec.by_value, spec.by_error) == before:
            return None
        return spec
    return spec if _spec_and(spec, s) else None


def trunc_select_batched_(Ss, spec, maxranks=None):
    """(ranks, eps) as host lists for the sorted spectra ``Ss`` (device vectors): one launch, one read.
    ``maxranks``: optional per-block rank caps (intersected with the strategy)."""
    from . import _core
    b = len(Ss)
    if b == 0:
        return [], []
    dev = Ss[0].device
    h = _core.Handle.get(dev)
    for S in Ss:
        if S.dtype != torch.float64 or S.dim() != 1 or (S.numel() > 1 and S.stride(0) != 1):
            raise ValueError("trunc_select_batched_: contiguous real vectors expected")
    k = (C.c_int * b)(*[S.numel() for S in Ss])
    Sp = (C.c_void_p * b)(*[S.data_ptr() for S in Ss])
    out = torch.empty(2 * b, dtype=torch.float64, device=dev)
    rank = out[:b].view(torch.int32)[:b]          # int32[b] carved from the same allocation
    eps = out[b:]
    work = h.workspace(h.lib.makb200_trunc_select_batched_worksize(h.h, b))
    caps = None
    if maxranks is not None:
        if len(maxranks) != b:
            raise ValueError("maxranks: one entry per block expected")
        caps = (C.c_int * b)(*[int(r) for r in maxranks])
    rc = h.lib.makb200_trunc_select_batched(h.h, b, k, Sp, C.byref(spec), caps, _core.ptr(rank), _core.ptr(eps),
                                            _core.ptr(work), work.numel())
    h.check(rc, "makb200_trunc_select_batched")
    host = out.cpu()                               # the one device->host read
    return host[:b].view(torch.int32)[:b].tolist(), host[b:].tolist()
