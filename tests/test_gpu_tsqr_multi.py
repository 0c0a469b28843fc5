"""Multi-rank TSQR through the C ABI (`makb200_tsqr(h, ncclComm_t, ...)`): binary-tree R-reduction over NCCL.

Runs only under a multi-process launch, one rank per GPU:
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 -m pytest tests/test_gpu_tsqr_multi.py -q
(skipped in the single-process `pytest -m gpu` run).  Every rank checks its slice of Q and the common R against
the oracle's qr_compact of the row-concatenated matrix: the north_star requirement "result must equal the
single-GPU qr_compact! of the concatenated matrix", tolerance 10 n eps (100x slack on the direct factor comparison)."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist

from oracle import mak_oracle as O

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(int(os.environ.get("WORLD_SIZE", "1")) < 2, reason="needs torchrun with >= 2 ranks")]


@pytest.fixture(scope="module")
def ranks():
    world, rank, local = (int(os.environ[k]) for k in ("WORLD_SIZE", "RANK", "LOCAL_RANK"))
    torch.cuda.set_device(local)
    if not dist.is_initialized():
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    yield world, rank, torch.device("cuda", local)
    dist.barrier()
    torch.cuda.synchronize()


def _rows(world, base, uneven):
    return [base + (37 * r if uneven else 0) for r in range(world)]


@pytest.mark.parametrize("dtype", ["f64", "c128"])
@pytest.mark.parametrize("base,n,uneven", [(3000, 64, False), (5000, 256, True), (700, 300, True)])
def test_tsqr_tree_vs_oracle(ranks, base, n, uneven, dtype):
    import makb200
    world, rank, dev = ranks
    rows = _rows(world, base, uneven)
    shards = [O.randn_matrix(rows[r], n, dtype, seed=5 + r) for r in range(world)]
    Q, R = makb200.tsqr_(makb200.to_device(shards[rank], dev))
    torch.cuda.synchronize()
    Qn, Rn = makb200.to_numpy(Q), makb200.to_numpy(R)
    Afull = np.vstack(shards)
    Qo, Ro = O.qr_compact(Afull)
    r0 = sum(rows[:rank])
    tol = O.tol_for(Afull.shape[0], n)
    assert np.array_equal(Rn, np.triu(Rn)) and np.all(np.diagonal(Rn).real > 0) and np.all(np.diagonal(Rn).imag == 0)
    assert np.linalg.norm(Rn - Ro) <= 100 * tol * np.linalg.norm(Ro)
    assert np.linalg.norm(Qn - Qo[r0:r0 + rows[rank]]) <= 100 * tol
    # R is bit-identical on every rank (broadcast), and the global Q is orthonormal: sum_p Q_p^H Q_p = I
    Rall = [torch.empty_like(R.t().contiguous()) for _ in range(world)]
    dist.all_gather(Rall, R.t().contiguous())
    for Rp in Rall:
        assert torch.equal(Rp, Rall[0])
    G = (Q.conj().t() @ Q).contiguous()
    dist.all_reduce(G)
    assert float(torch.linalg.matrix_norm(G - torch.eye(n, dtype=G.dtype, device=dev))) <= tol
    # residual of this rank's rows
    assert np.linalg.norm(shards[rank] - Qn @ Rn) / np.linalg.norm(shards[rank]) <= tol


def test_tsqr_tree_matches_python_model(ranks):
    """The C-ABI tree and the injectable Python model of it (tsqr.py, gloo-tested on CPU) agree."""
    import makb200
    from makb200.tsqr import _CudaOps
    world, rank, dev = ranks
    A0 = O.randn_matrix(4000, 128, "f64", seed=50 + rank)
    Q1, R1 = makb200.tsqr_(makb200.to_device(A0, dev))
    Q2, R2 = makb200.tsqr_(makb200.to_device(A0, dev), ops=_CudaOps())
    torch.cuda.synchronize()
    tol = O.tol_for(4000 * world, 128)
    assert float(torch.linalg.matrix_norm(R1 - R2) / torch.linalg.matrix_norm(R2)) <= 100 * tol
    assert float(torch.linalg.matrix_norm(Q1 - Q2)) <= 100 * tol
