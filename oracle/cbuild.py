"""Compiles the oracle's C restatements (oracle/tf_nms.c) with gcc into oracle/_build/ (git-ignored; travels to the GPU box with
the snapshot).  Test / baseline infrastructure only: nothing in ssd_keras_b200/ loads it."""
import ctypes as C
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, '_build')
_lib = None


def build_c(force=False):
    """Returns the path of libtfnms.so, or None where no C compiler is available (callers fall back to the NumPy restatement)."""
    src = os.path.join(HERE, 'tf_nms.c')
    so = os.path.join(OUT, 'libtfnms.so')
    if not force and os.path.exists(so) and os.path.getmtime(so) >= os.path.getmtime(src):
        return so
    gcc = shutil.which('gcc') or shutil.which('cc')
    if gcc is None:
        return so if os.path.exists(so) else None
    os.makedirs(OUT, exist_ok=True)
    tmp = so + '.%d.tmp' % os.getpid()
    r = subprocess.run([gcc, '-O3', '-ffp-contract=off', '-fno-fast-math', '-shared', '-fPIC', '-o', tmp, src], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError('gcc failed on oracle/tf_nms.c:\n' + r.stderr)
    os.replace(tmp, so)
    return so


def tf_nms_lib():
    global _lib
    if _lib is None:
        so = build_c()
        if so is None:
            return None
        lib = C.CDLL(so)
        lib.tf_nms_f32.argtypes = [C.POINTER(C.c_float), C.POINTER(C.c_float), C.c_int, C.c_int, C.c_float, C.POINTER(C.c_int)]
        lib.tf_nms_f32.restype = C.c_int
        _lib = lib
    return _lib
