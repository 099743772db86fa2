"""TEST INFRASTRUCTURE — ctypes binding of the compiled, unmodified reference.

``oracle/_ref/libptoy_ref.so`` is built by ``oracle/Makefile`` from
``/root/reference/src/particles.cpp`` + ``geometry.cpp`` (scalar build,
``-DUSE_AVX=0 -ffp-contract=off``: the parity oracle of SURVEY.md §8c) plus the
forwarding wrapper ``oracle/ref_wrap.cpp``.  Variants ``omp`` / ``avx_omp`` are the
timing baselines.  This module never touches ``/root/reference`` at run time.
"""
import os

from ._binding import OracleParticles, _Lib

_HERE = os.path.dirname(os.path.abspath(__file__))
_VARIANTS = {
    "scalar": "libptoy_ref.so",
    "omp": "libptoy_ref_omp.so",
    "avx_omp": "libptoy_ref_avx_omp.so",
}
_libs = {}


def lib_path(variant="scalar"):
    return os.path.join(_HERE, "_ref", _VARIANTS[variant])


def available(variant="scalar"):
    return os.path.exists(lib_path(variant))


def load(variant="scalar"):
    if variant not in _libs:
        _libs[variant] = _Lib(lib_path(variant), "ref_")
    return _libs[variant]


class RefParticles(OracleParticles):
    kind = "reference"

    def __init__(self, variant="scalar"):
        super().__init__(load(variant))
