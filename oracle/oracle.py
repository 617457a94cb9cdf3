"""ctypes wrapper around the CPU oracle (TEST INFRASTRUCTURE ONLY -- see bfhe_oracle.h).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
"""
import ctypes as C
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "libbfhe_oracle.so")

TOY, STD128_OPT = 0, 5
AP, GINX = 0, 1
OR, AND, NOR, NAND, XOR_FAST, XNOR_FAST, XOR, XNOR, BOOTSTRAP = range(9)
NEG0, NEG1 = 0x100, 0x200


class Params(C.Structure):
    _fields_ = [("paramset", C.c_uint32), ("method", C.c_uint32), ("n", C.c_uint32), ("N", C.c_uint32),
                ("q", C.c_uint32), ("Q", C.c_uint64), ("qKS", C.c_uint64), ("baseKS", C.c_uint32),
                ("dKS", C.c_uint32), ("baseG", C.c_uint32), ("dG", C.c_uint32), ("baseR", C.c_uint32),
                ("dR", C.c_uint32), ("ct_words", C.c_uint32), ("ct_stride", C.c_uint32)]


GATE_DTYPE = np.dtype([("op", "<u4"), ("in0", "<u4"), ("in1", "<u4"), ("out", "<u4")])


def build(force=False):
    src = [os.path.join(_HERE, f) for f in ("bfhe_oracle.c", "bfhe_oracle.h")]
    if force or not os.path.exists(_LIB) or any(os.path.getmtime(s) > os.path.getmtime(_LIB) for s in src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB):
            build()
        L = C.CDLL(_LIB)
        u32p, vp = C.POINTER(C.c_uint32), C.c_void_p
        L.orc_create.restype = vp
        L.orc_create.argtypes = [C.c_int, C.c_int]
        L.orc_destroy.argtypes = [vp]
        L.orc_get_params.argtypes = [vp, C.POINTER(Params)]
        L.orc_keygen.argtypes = [vp, C.c_uint64]
        L.orc_keyblob_size.restype = C.c_size_t
        L.orc_keyblob_size.argtypes = [vp]
        L.orc_export_keys.argtypes = [vp, vp, C.c_size_t]
        L.orc_import_keys.argtypes = [vp, vp, C.c_size_t]
        L.orc_encrypt_fresh.argtypes = [vp, C.c_int, C.c_uint64, u32p]
        L.orc_decrypt.argtypes = [vp, u32p]
        L.orc_eval_not.argtypes = [vp, u32p, u32p]
        L.orc_eval_bingate.argtypes = [vp, C.c_int, u32p, u32p, u32p]
        L.orc_bootstrap.argtypes = [vp, u32p, u32p]
        L.orc_eval_gates.argtypes = [vp, vp, C.c_int, u32p, C.c_int]
        L.orc_ntt_fwd.argtypes = [vp, u32p]
        L.orc_ntt_inv.argtypes = [vp, u32p]
        L.orc_signed_digit_decompose.argtypes = [vp, u32p, u32p]
        L.orc_prep.argtypes = [vp, C.c_uint32, u32p, u32p, u32p]
        L.orc_blind_rotate.argtypes = [vp, C.c_int, u32p, u32p]
        L.orc_extract_modswitch.argtypes = [vp, u32p, u32p]
        L.orc_keyswitch_modswitch.argtypes = [vp, u32p, u32p]
        L.orc_modulus_Q.restype = C.c_uint64
        L.orc_modulus_Q.argtypes = [C.c_int]
        _lib = L
    return _lib


def _p(a):
    assert a.dtype == np.uint32 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(C.POINTER(C.c_uint32))


class Oracle:
    def __init__(self, paramset=STD128_OPT, method=GINX):
        self.L = lib()
        self.h = self.L.orc_create(paramset, method)
        if not self.h:
            raise ValueError("unsupported paramset/method")
        self.p = Params()
        self.L.orc_get_params(self.h, C.byref(self.p))
        self.stride = self.p.ct_stride

    def __del__(self):
        if getattr(self, "h", None):
            self.L.orc_destroy(self.h)
            self.h = None

    def keygen(self, seed=1):
        assert self.L.orc_keygen(self.h, seed) == 0

    def export_keys(self):
        n = self.L.orc_keyblob_size(self.h)
        buf = np.empty(n, dtype=np.uint8)
        assert self.L.orc_export_keys(self.h, buf.ctypes.data, n) == 0
        return buf

    def import_keys(self, blob):
        blob = np.ascontiguousarray(blob, dtype=np.uint8)
        rc = self.L.orc_import_keys(self.h, blob.ctypes.data, blob.size)
        if rc:
            raise ValueError("orc_import_keys failed: %d" % rc)

    def new_slab(self, rows):
        return np.zeros((rows, self.stride), dtype=np.uint32)

    def encrypt(self, bits, seed=0):
        bits = np.asarray(bits).ravel()
        out = self.new_slab(len(bits))
        for i, b in enumerate(bits):
            self.L.orc_encrypt_fresh(self.h, int(b), seed * 1000003 + i, _p(out[i]))
        return out

    def decrypt(self, cts):
        cts = np.ascontiguousarray(cts, dtype=np.uint32)
        if cts.ndim == 1:
            return self.L.orc_decrypt(self.h, _p(cts))
        return np.array([self.L.orc_decrypt(self.h, _p(cts[i])) for i in range(cts.shape[0])])

    def eval_not(self, ct):
        out = np.zeros(self.stride, dtype=np.uint32)
        self.L.orc_eval_not(self.h, _p(ct), _p(out))
        return out

    def eval_bingate(self, op, ct1, ct2):
        out = np.zeros(self.stride, dtype=np.uint32)
        rc = self.L.orc_eval_bingate(self.h, op, _p(ct1), _p(ct2), _p(out))
        if rc:
            raise RuntimeError("EvalBinGate: inputs must be independent ciphertexts")
        return out

    def bootstrap(self, ct):
        out = np.zeros(self.stride, dtype=np.uint32)
        self.L.orc_bootstrap(self.h, _p(ct), _p(out))
        return out

    def eval_gates(self, gates, slab, nthreads=0):
        gates = np.ascontiguousarray(gates, dtype=GATE_DTYPE)
        assert slab.dtype == np.uint32 and slab.flags["C_CONTIGUOUS"] and slab.shape[1] == self.stride
        rc = self.L.orc_eval_gates(self.h, gates.ctypes.data, len(gates), _p(slab), nthreads)
        if rc:
            raise RuntimeError("orc_eval_gates failed: %d" % rc)

    def ntt_fwd(self, poly):
        a = np.array(poly, dtype=np.uint32)
        self.L.orc_ntt_fwd(self.h, _p(a))
        return a

    def ntt_inv(self, poly):
        a = np.array(poly, dtype=np.uint32)
        self.L.orc_ntt_inv(self.h, _p(a))
        return a

    def decompose(self, two_polys):
        a = np.ascontiguousarray(two_polys, dtype=np.uint32).reshape(-1)
        out = np.zeros(2 * self.p.dG * self.p.N, dtype=np.uint32)
        self.L.orc_signed_digit_decompose(self.h, _p(a), _p(out))
        return out.reshape(2 * self.p.dG, self.p.N)

    def prep(self, op, ct1, ct2):
        out = np.zeros(self.stride, dtype=np.uint32)
        self.L.orc_prep(self.h, op, _p(ct1), _p(ct2), _p(out))
        return out

    def blind_rotate(self, gate, prep):
        acc = np.zeros(2 * self.p.N, dtype=np.uint32)
        self.L.orc_blind_rotate(self.h, gate, _p(np.ascontiguousarray(prep)), _p(acc))
        return acc.reshape(2, self.p.N)

    def extract_modswitch(self, acc):
        ext = np.zeros(self.p.N + 1, dtype=np.uint32)
        self.L.orc_extract_modswitch(self.h, _p(np.ascontiguousarray(acc).reshape(-1)), _p(ext))
        return ext

    def keyswitch_modswitch(self, ext):
        out = np.zeros(self.stride, dtype=np.uint32)
        self.L.orc_keyswitch_modswitch(self.h, _p(np.ascontiguousarray(ext)), _p(out))
        return out
