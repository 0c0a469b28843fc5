"""TSQR: qr_compact! of a tall-skinny matrix row-sharded over ranks (BASELINE configs[3]).

New capability: the result must equal the single-GPU ``qr_compact!`` of the
row-concatenated matrix.  The product path is ONE C-ABI call, ``makb200_tsqr(h, ncclComm_t, ...)``
(csrc/polar.cu: tsqr_t): local CholeskyQR2 on the DMMA GEMM, binary-tree reduction of the n x n R
factors with ncclSend/ncclRecv issued from the producing stream, Householder QR of each stacked pair,
the tree's path product folded into the last local solve (``Q_p = Q1_p (L2^-H T_p)``) and an
ncclBroadcast of R.  This module only builds the communicator (``makb200_comm_create``; the
ncclUniqueId travels through torch.distributed) and marshals pointers.

``ops`` injects the numerical steps into a Python model of the same tree so that the topology
(send/recv pairing, uneven world sizes, down-sweep order) is tested with gloo on CPU
(tests/test_tsqr_gloo.py supplies a numpy stand-in).  ``robust=True`` (shifted CholeskyQR local step for
ill-conditioned shards) also takes that tree with the CUDA library as ``ops``.  Nothing here falls back to
the CPU: the default path fails loudly without a GPU."""

import ctypes as C

import torch
import torch.distributed as dist

from . import _core


class _CudaOps:
    """Numerical steps of TSQR on libmakb200 (no CPU fallback).  ``nshift`` > 0 selects the shifted
    CholeskyQR local step (``makb200_tsqr_local_ex``) for ill-conditioned shards."""

    def __init__(self, nshift=0):
        self.nshift = int(nshift)

    def local_qr(self, A):
        m, n = A.shape
        h = _core.Handle.get(A.device)
        dt = _core.dtype_code(A)
        Q = _core.colmajor_empty(m, n, A.dtype, A.device)
        R = _core.colmajor_empty(n, n, A.dtype, A.device)
        info = torch.zeros(1, dtype=torch.int32, device=A.device)
        lw = h.lib.makb200_tsqr_local_worksize(h.h, dt, m, n)
        work = h.workspace(lw)
        rc = h.lib.makb200_tsqr_local_ex(h.h, dt, m, n, _core.ptr(A), _core.ld(A), _core.ptr(Q), _core.ld(Q),
                                         _core.ptr(R), _core.ld(R), self.nshift, _core.ptr(work), work.numel(),
                                         _core.ptr(info))
        h.check(rc, "makb200_tsqr_local_ex")
        self._info = info
        return Q, R

    def check(self):
        if int(self._info.item()) != 0:
            raise _core.MakError("tsqr: Cholesky breakdown in the local factorization (matrix too ill-conditioned "
                                 "for CholeskyQR2, kappa >~ 1e7): call tsqr_(A, robust=True) on a fresh copy "
                                 "(the input was destroyed)")

    def small_qr(self, S):
        from .qr import qr_compact_
        return qr_compact_(S)

    def matmul(self, A, B):
        from .gemm import gemm_
        out = _core.colmajor_empty(A.shape[0], B.shape[1], A.dtype, A.device)
        return gemm_(out, A, B)

    def stack(self, Ra, Rb):
        n = Ra.shape[0]
        S = _core.colmajor_empty(2 * n, n, Ra.dtype, Ra.device)
        S[:n].copy_(Ra)
        S[n:].copy_(Rb)
        return S

    def empty(self, n, like):
        return _core.colmajor_empty(n, n, like.dtype, like.device)

    def eye(self, n, like):
        T = _core.colmajor_zeros(n, n, like.dtype, like.device)
        T.diagonal().fill_(1)
        return T


def _send(t, dst, group):
    # column-major n x n block: ship the underlying contiguous (n, n) buffer
    dist.send(t.t().contiguous(), dst, group=group)


def _recv(buf, src, group):
    tmp = torch.empty((buf.shape[1], buf.shape[0]), dtype=buf.dtype, device=buf.device)
    dist.recv(tmp, src, group=group)
    buf.copy_(tmp.t())
    return buf


_COMMS = {}


def nccl_comm(group=None, device=None):
    """ncclComm_t (as a ctypes void pointer) spanning ``group`` in rank order, created once per (group, device)
    by ``makb200_comm_create``; the 128-byte ncclUniqueId is broadcast through torch.distributed."""
    from . import _lib
    lib = _lib.load()
    dev = torch.device(device if device is not None else torch.device("cuda", torch.cuda.current_device()))
    key = (id(group) if group is not None else 0, dev.index)
    if key in _COMMS:
        return _COMMS[key]
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    buf = (C.c_char * 128)()
    if rank == 0:
        rc = lib.makb200_nccl_unique_id(buf)
        if rc != 0:
            raise _core.MakError(f"makb200_nccl_unique_id failed ({rc}): libnccl.so.2 not loadable")
    box = [bytes(buf) if rank == 0 else None]
    dist.broadcast_object_list(box, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
    idbytes = (C.c_char * 128).from_buffer_copy(box[0])
    comm = C.c_void_p()
    with torch.cuda.device(dev):
        rc = lib.makb200_comm_create(C.byref(comm), world, rank, idbytes)
    if rc != 0:
        raise _core.MakError(f"makb200_comm_create failed ({rc})")
    _COMMS[key] = comm
    return comm


def _tsqr_cabi(A_local, group, check):
    m, n = A_local.shape
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    h = _core.Handle.get(A_local.device)
    dt = _core.dtype_code(A_local)
    if not _core.is_colmajor(A_local):
        raise ValueError("tsqr_: column-major shard expected")
    comm = nccl_comm(group, A_local.device) if world > 1 else C.c_void_p(0)
    Q = _core.colmajor_empty(m, n, A_local.dtype, A_local.device)
    R = _core.colmajor_empty(n, n, A_local.dtype, A_local.device)
    info = torch.zeros(1, dtype=torch.int32, device=A_local.device)
    lw = h.lib.makb200_tsqr_worksize(h.h, dt, m, n, world)
    work = h.workspace(lw)
    rc = h.lib.makb200_tsqr(h.h, comm, dt, m, n, _core.ptr(A_local), _core.ld(A_local), _core.ptr(Q), _core.ld(Q),
                            _core.ptr(R), _core.ld(R), _core.ptr(work), work.numel(), _core.ptr(info))
    h.check(rc, "makb200_tsqr")
    if check and int(info.item()) != 0:
        raise _core.MakError("tsqr: Cholesky breakdown in the local factorization (matrix too ill-conditioned "
                             "for CholeskyQR2, kappa >~ 1e7): call tsqr_(A, robust=True) on a fresh copy "
                             "(the input was destroyed)")
    return Q, R


def tsqr_(A_local, group=None, ops=None, check=True, robust=False):
    """Row-sharded ``qr_compact!``: every rank passes its (m_loc x n) shard (destroyed) and gets
    back (Q_local, R) with R identical on all ranks, diag(R) >= 0.

    ``robust=True`` (or an int 1..3 = number of preconditioning passes) runs the local step as
    shifted CholeskyQR (2 extra passes by default): any numerically full-rank shard, at twice the
    cost of the default CholeskyQR2 (kappa <~ 1e7)."""
    if ops is None and not robust:
        return _tsqr_cabi(A_local, group, check)
    if ops is None:
        ops = _CudaOps(nshift=(2 if robust is True else int(robust)))
    n = A_local.shape[1]
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    Q, R = ops.local_qr(A_local)
    if world == 1:
        if check and hasattr(ops, "check"):
            ops.check()
        return Q, R
    # ---- up-sweep: binary tree over ranks ------------------------------------------------
    factors = []  # (round stride, Qs) kept by receivers
    stride = 1
    active = True
    sent_to = None
    while stride < world:
        if active:
            if rank % (2 * stride) == 0:
                partner = rank + stride
                if partner < world:
                    Rb = _recv(ops.empty(n, R), partner, group)
                    Qs, R = ops.small_qr(ops.stack(R, Rb))
                    factors.append((stride, partner, Qs))
            else:
                partner = rank - stride
                _send(R, partner, group)
                sent_to = partner
                active = False
        stride *= 2
    # ---- down-sweep: path products T_p, then Q_p <- Q_p T_p ----------------------------------
    if rank == 0:
        T = ops.eye(n, R)
    else:
        T = _recv(ops.empty(n, R), sent_to, group)
    for stride, partner, Qs in reversed(factors):
        Tb = ops.matmul(Qs[n:], T)
        _send(Tb, partner, group)
        T = ops.matmul(Qs[:n], T)
    Q = ops.matmul(Q, T)
    # ---- final R to everybody ---------------------------------------------------------------
    Rbuf = R.t().contiguous() if rank == 0 else torch.empty((n, n), dtype=R.dtype, device=R.device)
    dist.broadcast(Rbuf, 0, group=group)
    R = ops.empty(n, Q)
    R.copy_(Rbuf.t())
    if check and hasattr(ops, "check"):
        ops.check()
    return Q, R
