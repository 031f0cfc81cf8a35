"""ctypes binding of the synthetic stream generator (tools/streamgen, test + bench input)."""
import ctypes as C
import os

import numpy as np

from . import build as _build
from .abi import MbInfo, PicParams, mb_dtype


class GenParams(C.Structure):
    _fields_ = [("width", C.c_int32), ("height", C.c_int32), ("chroma_format", C.c_int32),
                ("n_gops", C.c_int32), ("gop_n", C.c_int32), ("gop_m", C.c_int32), ("intra_only", C.c_int32),
                ("seed", C.c_uint64), ("mode", C.c_int32), ("mv_range", C.c_int32), ("qscale_code_max", C.c_int32),
                ("alternate_scan", C.c_int32), ("q_scale_type", C.c_int32), ("intra_dc_precision", C.c_int32),
                ("pct_skipped", C.c_int32), ("pct_intra_in_pb", C.c_int32), ("pct_coded", C.c_int32),
                ("pct_mb_quant", C.c_int32), ("pct_big_levels", C.c_int32), ("all_blocks_coded", C.c_int32),
                ("natural_mean_coefs", C.c_int32), ("unclamped_mv", C.c_int32),
                ("user_data_bytes", C.c_int32), ("texture_noise", C.c_int32), ("matrices_once", C.c_int32),
                ("pct_field_dct", C.c_int32), ("intra_vlc_table0", C.c_int32)]


class GenPicture(C.Structure):
    _fields_ = [("params", PicParams), ("mb", C.POINTER(MbInfo)), ("coef", C.POINTER(C.c_uint32)),
                ("mb_count", C.c_uint32), ("n_coef", C.c_uint32), ("display_index", C.c_int32), ("gop", C.c_int32),
                ("q_scale_type", C.c_int32), ("intra_dc_precision", C.c_int32),
                ("tx", (C.c_uint8 * 64) * 4), ("tx_loaded", C.c_int32 * 4)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        path = _build.STREAMGEN_LIB
        if not os.path.exists(path):
            _build.build_streamgen()
        L = C.CDLL(path)
        L.mp2v_gen_default_params.argtypes = [C.POINTER(GenParams), C.c_int, C.c_int, C.c_int]
        L.mp2v_gen_create.restype = C.c_void_p
        L.mp2v_gen_create.argtypes = [C.POINTER(GenParams)]
        L.mp2v_gen_destroy.argtypes = [C.c_void_p]
        L.mp2v_gen_error.restype = C.c_char_p
        L.mp2v_gen_error.argtypes = [C.c_void_p]
        L.mp2v_gen_stream.restype = C.c_size_t
        L.mp2v_gen_stream.argtypes = [C.c_void_p, C.POINTER(C.POINTER(C.c_uint8))]
        L.mp2v_gen_gop_offset.restype = C.c_size_t
        L.mp2v_gen_gop_offset.argtypes = [C.c_void_p, C.c_int]
        L.mp2v_gen_num_pictures.argtypes = [C.c_void_p]
        L.mp2v_gen_picture.argtypes = [C.c_void_p, C.c_int, C.POINTER(GenPicture)]
        _lib = L
    return _lib


class Picture:
    """Ground truth of one coded picture: numpy copies of the generator's records."""

    def __init__(self, gp):
        self.params = PicParams.from_buffer_copy(gp.params)
        n_mb, n_coef = gp.mb_count, gp.n_coef
        self.mb = np.ctypeslib.as_array(C.cast(gp.mb, C.POINTER(C.c_uint8)), shape=(n_mb * 16,)).copy().view(mb_dtype)
        if n_coef:
            self.coef = np.ctypeslib.as_array(gp.coef, shape=(n_coef,)).copy()
        else:
            self.coef = np.zeros(0, np.uint32)
        self.display_index = gp.display_index
        self.gop = gp.gop
        self.q_scale_type = gp.q_scale_type
        self.intra_dc_precision = gp.intra_dc_precision
        self.tx = np.ctypeslib.as_array(gp.tx).copy()
        self.tx_loaded = list(gp.tx_loaded)

    @property
    def type(self):
        return self.params.picture_coding_type


class Stream:
    """A generated elementary stream plus its ground truth."""

    def __init__(self, width, height, chroma_format, **kw):
        L = lib()
        p = GenParams()
        L.mp2v_gen_default_params(C.byref(p), width, height, chroma_format)
        for k, v in kw.items():
            if not hasattr(p, k):
                raise TypeError("unknown generator parameter %r" % k)
            setattr(p, k, v)
        self.gen_params = p
        g = L.mp2v_gen_create(C.byref(p))
        try:
            e = L.mp2v_gen_error(g)
            if e:
                raise RuntimeError("stream generator: " + e.decode())
            ptr = C.POINTER(C.c_uint8)()
            n = L.mp2v_gen_stream(g, C.byref(ptr))
            # keep the 256 zero bytes of padding the decoders' over-reads need
            self.padded = np.ctypeslib.as_array(ptr, shape=(n + 256,)).copy()
            self.size = n
            self.gop_offsets = [L.mp2v_gen_gop_offset(g, i) for i in range(p.n_gops + 1)]
            self.pictures = []
            gp = GenPicture()
            for i in range(L.mp2v_gen_num_pictures(g)):
                L.mp2v_gen_picture(g, i, C.byref(gp))
                self.pictures.append(Picture(gp))
        finally:
            L.mp2v_gen_destroy(g)
        self.width, self.height, self.chroma_format = width, height, chroma_format

    @property
    def data(self):
        return self.padded[:self.size]

    def display_order(self):
        """coded indices sorted by display position"""
        return sorted(range(len(self.pictures)), key=lambda i: self.pictures[i].display_index)
