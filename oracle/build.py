"""Build and load the plain-C oracle (TEST INFRASTRUCTURE ONLY).

The reference is pure Python (no C sources to compile into oracle/_ref -- see
DESIGN.md), so the compiled CPU checker is this repo's own C restatement.
"""
import ctypes
import os
import subprocess

_DIR = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(_DIR, "liboracle_cube.so")
_SOURCES = ("cube_oracle.c", "cube_oracle_tables.h", "Makefile")


def build(force=False):
    stale = force or not os.path.exists(LIB) or any(
        os.path.getmtime(os.path.join(_DIR, s)) > os.path.getmtime(LIB) for s in _SOURCES)
    if stale:
        subprocess.check_call(["make", "-s", "-C", _DIR, "liboracle_cube.so"] + (["-B"] if force else []))
    return LIB


_lib = None


def load():
    """ctypes handle to liboracle_cube.so (built on demand when gcc is present)."""
    global _lib
    if _lib is not None:
        return _lib
    try:
        build()
    except (OSError, subprocess.CalledProcessError):
        if not os.path.exists(LIB):
            raise
    lib = ctypes.CDLL(LIB)
    u8p, f32p, i64 = ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64
    lib.orc_num_threads.restype = ctypes.c_int
    lib.orc_set_threads.argtypes = [ctypes.c_int]
    lib.orc_step.argtypes = [ctypes.c_int, u8p, u8p, i64, u8p, f32p]
    lib.orc_scramble.argtypes = [ctypes.c_int, u8p, u8p, i64, ctypes.c_int, u8p, u8p, f32p, u8p,
                                 ctypes.c_void_p]
    lib.orc_solved.argtypes = [ctypes.c_int, u8p, i64, u8p, f32p]
    lib.orc_columns.argtypes = [ctypes.c_int, u8p, i64, u8p]
    lib.orc_encode_u8.argtypes = [ctypes.c_int, u8p, i64, u8p]
    lib.orc_expand.argtypes = [ctypes.c_int, u8p, i64, u8p, u8p, u8p]
    for f in (lib.orc_step, lib.orc_scramble, lib.orc_solved, lib.orc_columns,
              lib.orc_encode_u8, lib.orc_expand):
        f.restype = ctypes.c_int
    _lib = lib
    return lib


if __name__ == "__main__":
    print(build(force=True))
