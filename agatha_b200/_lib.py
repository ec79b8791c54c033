"""ctypes binding of libagatha_b200.so. Fails loudly when the library is missing: there is no fallback."""
import ctypes
import os

PKG = os.path.dirname(os.path.abspath(__file__))
# AGATHA_B200_LIB: load another build of the same library (kernel A/B measurements, tools/kperf.py)
_LIB_PATH = os.environ.get("AGATHA_B200_LIB") or os.path.join(PKG, "lib", "libagatha_b200.so")
_lib = None

STOP_END, STOP_ZDROP, STOP_BANDEXIT = 0, 1, 2


class AgathaError(RuntimeError):
    pass


class Params(ctypes.Structure):
    """agatha_params_t == gasal_subst_scores (AGAThA/src/gasal.h:165-173)."""
    _fields_ = [("match", ctypes.c_int32), ("mismatch", ctypes.c_int32), ("gap_open", ctypes.c_int32),
                ("gap_extend", ctypes.c_int32), ("slice_width", ctypes.c_int32),
                ("z_threshold", ctypes.c_int32), ("band_width", ctypes.c_int32)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


# AGAThA.sh:44
DEFAULT_PARAMS = dict(match=1, mismatch=4, gap_open=6, gap_extend=2, slice_width=3, z_threshold=400, band_width=751)


def make_params(**kw):
    d = dict(DEFAULT_PARAMS)
    d.update(kw)
    return Params(**d)


def lib_path():
    return _LIB_PATH


def lib():
    """The loaded shared library. Raises AgathaError if it has not been built (python -m agatha_b200.build)."""
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            raise AgathaError("libagatha_b200.so not found at %s: run `python -m agatha_b200.build` "
                              "(there is no CPU fallback)" % _LIB_PATH)
        L = ctypes.CDLL(_LIB_PATH)
        L.agatha_last_error.restype = ctypes.c_char_p
        L.agatha_launch_count.restype = ctypes.c_uint64
        _lib = L
    return _lib


def check(rc):
    if rc != 0:
        raise AgathaError("agatha_b200 error %d: %s" % (rc, lib().agatha_last_error().decode()))
