"""makb200 — B200-native dense factorizations behind MatrixAlgebraKit.jl's algorithm API.

Host layer in Python (the reference's Julia toolchain is absent from this image; the Julia
extension that binds the same C ABI is in ``ext/MatrixAlgebraKitB200Ext``).  All numerical work
happens in ``libmakb200.so`` (hand-written sm_100a CUDA); there is no CPU fallback."""
from . import _lib
from ._core import (Handle, MakError, as_colmajor, colmajor_empty, colmajor_zeros, is_colmajor, to_device,
                    to_numpy)
from .algorithms import *  # noqa: F401,F403
from .algorithms import (Algorithm, TruncatedAlgorithm, default_algorithm, select_algorithm)
from .gemm import gemm_
from .qr import (geqrf_, qr_compact, qr_compact_, qr_compact_batched_, qr_full, qr_full_, qr_householder_, qr_null_householder_,
                 ungqr_, unmqr_)
from .eigh import (DomainError, check_hermitian, eigh_full, eigh_full_, eigh_trunc, eigh_trunc_, eigh_vals,
                   eigh_vals_)
from .truncation import (findtruncated, findtruncated_svd, notrunc, select_truncation, trunc_and, trunc_or,
                         truncerror, truncrank, trunctol)
from .polar import left_polar, left_polar_
from .svd import (svd_compact, svd_compact_, svd_full, svd_full_, svd_trunc, svd_trunc_, svd_trunc_no_error, svd_trunc_no_error_,
                  svd_vals, svd_vals_)
from . import eigh, polar, qr, svd, truncation  # noqa: E402,F401
from .tsqr import tsqr_
from .svd import svd_compact_batched_, svd_trunc_batched_
from .qr import BatchedQRPlan
from .svd import BatchedSVDPlan
from .eigh import BatchedEighPlan, eigh_full_batched_
from .orthnull import (adjoint_, left_null, left_null_, left_orth, left_orth_, lq_compact, lq_compact_, lq_full, lq_full_,
                       lq_null, lq_null_, qr_null, qr_null_, right_null, right_null_, right_orth, right_orth_)
from . import partition  # noqa: E402,F401
from .partition import gather_block_info, lpt_partition, my_blocks  # noqa: E402,F401
from .sbr import sbr_apply_q2_, sbr_chase_, sy2sb_  # noqa: E402,F401  (experimental)
from .projections import (defaulttol, is_left_isometric, is_right_isometric, isantihermitian, ishermitian,  # noqa: E402,F401
                          isisometric, isunitary, project_antihermitian, project_antihermitian_, project_hermitian,
                          project_hermitian_, project_isometric, project_isometric_)
from .eigh import eigh_vals_batched_  # noqa: E402,F401
from .svd import svd_vals_batched_  # noqa: E402,F401
from .projections import lowertriangular_, one_, uppertriangular_  # noqa: E402,F401
