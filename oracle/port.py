"""TEST INFRASTRUCTURE — ctypes binding of the plain-C restatement
(``oracle/ptoy_oracle.c`` → ``oracle/libptoy_oracle.so``, built by ``oracle/Makefile``
or ``__graft_entry__.build()``)."""
import os

from ._binding import OracleParticles, _Lib

_HERE = os.path.dirname(os.path.abspath(__file__))
_lib = None


def lib_path():
    return os.path.join(_HERE, "libptoy_oracle.so")


def available():
    return os.path.exists(lib_path())


def load():
    global _lib
    if _lib is None:
        _lib = _Lib(lib_path(), "orc_")
    return _lib


class PortParticles(OracleParticles):
    kind = "port"

    def __init__(self):
        super().__init__(load())
