"""Host-side mirror of the reference's algorithm-struct API (src/algorithms.jl:10-45,106-145,
171-238; src/interface/decompositions.jl).  An ``Algorithm`` is a name plus a kwargs mapping,
exactly like ``Algorithm{name,KW}``; ``select_algorithm`` accepts the same five ``alg`` forms
(None, str/symbol, class, dict/NamedTuple, instance) and raises ``ValueError`` (Julia:
``ArgumentError``) for kwargs combined with an instance (algorithms.jl:117-120).

The new names this backend adds (declared next to CUSOLVER_* in the core package,
interface/decompositions.jl:365-430, implemented in ext/MatrixAlgebraKitB200Ext):
``B200`` driver, ``B200_HouseholderQR``, ``B200_DivideAndConquer`` (eigh: tridiagonal D&C),
``B200_SVDViaPolar`` (QDWH + eigh, the reference's ``SVDViaPolar`` tag), ``B200_QDWH`` (polar),
``B200_Jacobi`` (batched small blocks)."""
from dataclasses import dataclass, field


class Driver:
    """``abstract type Driver`` (src/algorithms.jl:171-199)."""

    def __repr__(self):
        return f"{type(self).__name__}()"

    def __eq__(self, other):
        return type(self) is type(other)

    def __hash__(self):
        return hash(type(self).__name__)


class DefaultDriver(Driver):
    pass


class B200(Driver):
    """New driver: hand-written sm_100a kernels behind libmakb200 (no cuSOLVER, no CPU)."""


class LAPACK(Driver):
    """Named so that requesting it raises the same way an unavailable driver does."""


class CUSOLVER(Driver):
    pass


@dataclass(frozen=True)
class Algorithm:
    """``Algorithm{name,KW}`` (src/algorithms.jl:20-45)."""
    name: str
    kwargs: dict = field(default_factory=dict)

    def get(self, key, default=None):
        return self.kwargs.get(key, default)

    def __hash__(self):
        return hash((self.name, tuple(sorted(self.kwargs.items(), key=lambda kv: kv[0]))))


def _algdef(name, allowed, defaults=None):
    defaults = dict(defaults or {})

    def ctor(**kw):
        for k in kw:
            if k not in allowed:
                raise ValueError(f"{name}: unknown keyword argument `{k}`")  # ArgumentError
        d = dict(defaults)
        d.update(kw)
        return Algorithm(name, d)

    ctor.__name__ = name
    ctor.algname = name
    return ctor


# reference names (interface/decompositions.jl:80-86,97,116,137,170,183,294-305)
Householder = _algdef("Householder", {"blocksize", "driver", "pivoted", "positive"})
DivideAndConquer = _algdef("DivideAndConquer", {"driver", "fixgauge", "hermitian_tol"})
SafeDivideAndConquer = _algdef("SafeDivideAndConquer", {"driver", "fixgauge"})
QRIteration = _algdef("QRIteration", {"driver", "fixgauge", "hermitian_tol"})
RobustRepresentations = _algdef("RobustRepresentations", {"driver", "fixgauge", "hermitian_tol"})
Jacobi = _algdef("Jacobi", {"driver", "fixgauge", "tol", "max_sweeps", "hermitian_tol"})
SVDViaPolar = _algdef("SVDViaPolar", {"driver", "fixgauge", "tol"})
PolarNewton = _algdef("PolarNewton", {"maxiter", "tol"})


def PolarViaSVD(svd_alg=None):
    return Algorithm("PolarViaSVD", {"svd_alg": svd_alg})


# B200-native names
B200_HouseholderQR = _algdef("Householder", {"blocksize", "pivoted", "positive", "driver"}, {"driver": B200()})
B200_DivideAndConquer = _algdef("DivideAndConquer", {"fixgauge", "hermitian_tol", "driver"}, {"driver": B200()})
B200_SVDViaPolar = _algdef("SVDViaPolar", {"fixgauge", "tol", "driver"}, {"driver": B200()})
B200_Jacobi = _algdef("Jacobi", {"fixgauge", "tol", "max_sweeps", "driver", "hermitian_tol"}, {"driver": B200()})
B200_QDWH = _algdef("QDWH", {"tol", "maxiter", "driver"}, {"driver": B200()})


@dataclass(frozen=True)
class TruncatedAlgorithm:
    """``TruncatedAlgorithm(alg, trunc)`` (src/algorithms.jl:325-330)."""
    alg: Algorithm
    trunc: object


_DEFAULT_FIXGAUGE = [True]


def default_fixgauge(new=None):
    """``default_fixgauge()`` global (src/common/defaults.jl:47-61)."""
    if new is not None:
        _DEFAULT_FIXGAUGE[0] = bool(new)
    return _DEFAULT_FIXGAUGE[0]


def default_driver(A):
    """``default_driver`` for device matrices = B200() (pattern: MatrixAlgebraKitCUDAExt.jl:19)."""
    return B200()


def resolve_driver(driver, A):
    if driver is None or isinstance(driver, DefaultDriver):
        return default_driver(A)
    if not isinstance(driver, B200):
        raise ValueError(f"driver {driver!r} is not available for B200 device matrices")
    return driver


# default algorithm per op for device matrices (pattern: MatrixAlgebraKitCUDAExt.jl:21-29)
def default_qr_algorithm(A, **kw):
    return Householder(**kw)


def default_svd_algorithm(A, **kw):
    return SVDViaPolar(**kw)


def default_eigh_algorithm(A, **kw):
    return DivideAndConquer(**kw)


def default_polar_algorithm(A, **kw):
    return B200_QDWH(**kw)


_DEFAULTS = {
    "qr_compact": default_qr_algorithm, "qr_full": default_qr_algorithm, "qr_null": default_qr_algorithm,
    "svd_compact": default_svd_algorithm, "svd_full": default_svd_algorithm,
    "svd_vals": default_svd_algorithm,
    "eigh_full": default_eigh_algorithm, "eigh_vals": default_eigh_algorithm,
    "left_polar": default_polar_algorithm,
}

_BY_NAME = {
    "Householder": Householder, "DivideAndConquer": DivideAndConquer,
    "SafeDivideAndConquer": SafeDivideAndConquer, "QRIteration": QRIteration,
    "RobustRepresentations": RobustRepresentations, "Jacobi": Jacobi, "SVDViaPolar": SVDViaPolar,
    "PolarNewton": PolarNewton, "QDWH": B200_QDWH,
}


def default_algorithm(f, A, **kw):
    return _DEFAULTS[f](A, **kw)


def select_algorithm(f, A, alg=None, **kw):
    """``select_algorithm`` (src/algorithms.jl:106-124)."""
    if alg is None:
        return default_algorithm(f, A, **kw)
    if isinstance(alg, str):
        if alg not in _BY_NAME:
            raise ValueError(f"Unknown alg {alg}")
        return _BY_NAME[alg](**kw)
    if callable(alg) and hasattr(alg, "algname"):
        return alg(**kw)
    if isinstance(alg, dict):
        if kw:
            raise ValueError("Additional keyword arguments are not allowed when algorithm parameters are specified.")
        return default_algorithm(f, A, **alg)
    if isinstance(alg, (Algorithm, TruncatedAlgorithm)):
        if kw:
            raise ValueError("Additional keyword arguments are not allowed when algorithm parameters are specified.")
        return alg
    raise ValueError(f"Unknown alg {alg}")
