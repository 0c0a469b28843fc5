"""TSQR: qr_compact! of a tall-skinny matrix row-sharded over ranks (BASELINE configs[3]).

New capability: the result must equal the single-GPU ``qr_compact!`` of the
row-concatenated matrix.  Per rank: local factorization (``makb200_tsqr_local``, CholeskyQR2 on
the DMMA GEMM).  Across ranks: binary-tree reduction of the n x n R factors — each round one
``send/recv`` of a 256x256 block (512 KiB) over NCCL/NVLink, a Householder QR of the stacked
2n x n pair on the receiver, the tree factors pushed back down and one local GEMM
``Q_p <- Q_p T_p``.  The exchange is issued from the stream that produced R (no host sync in
the data path); it is latency-bound (~10-20 us per round) next to >= 15 ms of local work.

The numerical kernels are injected through ``ops`` so the tree logic can be tested with gloo on
CPU (tests/test_tsqr_gloo.py supplies a numpy stand-in); the default ``ops`` is the CUDA library
and fails loudly without a GPU."""

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


def tsqr_(A_local, group=None, ops=None, check=True, robust=False):
    """Row-sharded ``qr_compact!``: every rank passes its (m_loc x n) shard (destroyed) and gets
    back (Q_local, R) with R identical on all ranks, diag(R) >= 0.

    ``robust=True`` (or an int 1..3 = number of preconditioning passes) runs the local step as
    shifted CholeskyQR (2 extra passes by default): any numerically full-rank shard, at twice the
    cost of the default CholeskyQR2 (kappa <~ 1e7)."""
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
