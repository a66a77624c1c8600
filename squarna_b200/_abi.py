"""ctypes mirror of include/sqrn.h (structs only; the loader lives in _lib.py)."""
import ctypes as C

import numpy as np

MAX_BPKEYS = 32
MAX_LEN = 16000

OK, E_BADARG, E_CAPACITY, E_CUDA, E_UNSUPPORTED, E_NOMEM = 0, -1, -2, -3, -4, -5
RC_UNPAIRED, RC_NOLEFT, RC_NORIGHT = 1, 2, 4


class ParamSet(C.Structure):
    """sqrn_paramset"""
    _fields_ = [("n_bp", C.c_int32),
                ("bp_keys", C.c_uint8 * (2 * MAX_BPKEYS)),
                ("bp_vals", C.c_double * MAX_BPKEYS),
                ("suboptmax", C.c_double), ("suboptmin", C.c_double), ("suboptsteps", C.c_double),
                ("minlen", C.c_double), ("minbpscore", C.c_double), ("minfinscorefactor", C.c_double),
                ("bracketweight", C.c_double), ("distcoef", C.c_double), ("orderpenalty", C.c_double),
                ("loopbonus", C.c_double), ("maxstemnum", C.c_double)]

    @classmethod
    def from_dict(cls, ps):
        """ps: a parameter-set dict as returned by ParseConfig."""
        out = cls()
        items = list(ps["bpweights"].items())
        if len(items) > MAX_BPKEYS:
            raise ValueError("too many bpweights entries")
        out.n_bp = len(items)
        for k, (key, val) in enumerate(items):
            if len(key) != 2:
                raise ValueError("bpweights keys must be two symbols: %r" % (key,))
            kb = key.encode("latin-1", "replace")
            out.bp_keys[2 * k] = kb[0]
            out.bp_keys[2 * k + 1] = kb[1]
            out.bp_vals[k] = float(val)
        for f in ("suboptmax", "suboptmin", "suboptsteps", "minlen", "minbpscore", "minfinscorefactor",
                  "bracketweight", "distcoef", "orderpenalty", "loopbonus", "maxstemnum"):
            setattr(out, f, float(ps[f]))
        return out


def paramset_array(paramsets):
    arr = (ParamSet * max(len(paramsets), 1))()
    for k, ps in enumerate(paramsets):
        arr[k] = ParamSet.from_dict(ps)
    return arr


class Batch(C.Structure):
    """sqrn_batch"""
    _fields_ = [("n_seqs", C.c_int64), ("offsets", C.c_void_p), ("symbols", C.c_void_p),
                ("react_code", C.c_void_p), ("react_values", C.c_void_p), ("n_react_values", C.c_int32),
                ("react_sum_compensated", C.c_int32),
                ("restr_class", C.c_void_p), ("rbp_offsets", C.c_void_p), ("rbps", C.c_void_p),
                ("smat", C.c_void_p), ("smat_L", C.c_int32), ("cols", C.c_void_p),
                ("interchainonly", C.c_int32), ("hardrest", C.c_int32), ("rankbydiff", C.c_int32),
                ("poollim", C.c_int32), ("conslim", C.c_int32), ("max_structs", C.c_int32),
                ("rankby", C.c_int32 * 3), ("priority_mask", C.c_uint64),
                ("bpp_term", C.c_void_p), ("bpp_offsets", C.c_void_p), ("bpp_mode", C.c_int32)]


class Result(C.Structure):
    """sqrn_result"""
    _fields_ = [("cap_structs", C.c_int64), ("cap_stems", C.c_int64), ("cap_dbn", C.c_int64),
                ("struct_offsets", C.c_void_p), ("scores", C.c_void_p), ("struct_is_int0", C.c_void_p),
                ("psmask", C.c_void_p), ("n_total", C.c_void_p), ("stem_offsets", C.c_void_p),
                ("stems", C.c_void_p), ("dbn_offsets", C.c_void_p), ("dbn", C.c_void_p), ("cons", C.c_void_p),
                ("need_structs", C.c_int64), ("need_stems", C.c_int64), ("need_dbn", C.c_int64)]


class Stems(C.Structure):
    """sqrn_stems"""
    _fields_ = [("cap_stems", C.c_int64), ("stem_offsets", C.c_void_p), ("stems", C.c_void_p),
                ("scores", C.c_void_p), ("need_stems", C.c_int64)]


def ptr(a):
    """void* of a numpy array (None -> NULL)."""
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def pack_sequences(seqs):
    """list of str/bytes -> (uint8 symbols, int64 offsets)"""
    enc = [s.encode("latin-1", "replace") if isinstance(s, str) else bytes(s) for s in seqs]
    lens = np.fromiter((len(e) for e in enc), dtype=np.int64, count=len(enc))
    offsets = np.zeros(len(enc) + 1, dtype=np.int64)
    np.cumsum(lens, out=offsets[1:])
    symbols = np.frombuffer(b"".join(enc), dtype=np.uint8).copy() if enc else np.zeros(0, np.uint8)
    if symbols.size == 0:
        symbols = np.zeros(1, np.uint8)
    return symbols, offsets
