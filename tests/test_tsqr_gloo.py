"""TSQR tree logic over world_size = 2 and 3 with gloo on CPU (the numerical steps are injected:
a numpy stand-in built on the oracle replaces the CUDA library here; the GPU test runs the real
one)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import mak_oracle as O


class NumpyOps:
    def local_qr(self, A):
        Q, R = O.qr_compact(A.numpy())
        return torch.from_numpy(np.ascontiguousarray(Q)), torch.from_numpy(np.ascontiguousarray(R))

    def small_qr(self, S):
        Q, R = O.qr_compact(S.numpy())
        return torch.from_numpy(np.ascontiguousarray(Q)), torch.from_numpy(np.ascontiguousarray(R))

    def matmul(self, A, B):
        return A @ B

    def stack(self, Ra, Rb):
        return torch.cat([Ra, Rb], 0)

    def empty(self, n, like):
        return torch.empty((n, n), dtype=like.dtype)

    def eye(self, n, like):
        return torch.eye(n, dtype=like.dtype)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, m_loc, n, dtype, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import makb200
    A = torch.from_numpy(np.ascontiguousarray(O.randn_matrix(m_loc, n, dtype, seed=5 + rank)))
    Q, R = makb200.tsqr_(A.clone(), ops=NumpyOps())
    out[rank] = (Q.numpy(), R.numpy())
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
@pytest.mark.parametrize("dtype", ["f64", "c128"])
def test_tsqr_tree_gloo(world, dtype):
    m_loc, n = 40, 8
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), m_loc, n, dtype, out), nprocs=world, join=True)
    A = np.vstack([O.randn_matrix(m_loc, n, dtype, seed=5 + r) for r in range(world)])
    Qo, Ro = O.qr_compact(A)
    Q = np.vstack([out[r][0] for r in range(world)])
    for r in range(world):
        assert np.linalg.norm(out[r][1] - Ro) < 1e-12      # same R everywhere = single-matrix qr_compact
    assert np.linalg.norm(Q - Qo) < 1e-12
    assert O.orth_err(Q) < O.tol_for(m_loc * world, n)
