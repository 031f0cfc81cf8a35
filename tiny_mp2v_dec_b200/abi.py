"""ctypes / numpy mirrors of include/mp2v_recon.h (one definition for every Python binding)."""
import ctypes as C

import numpy as np


class MbInfo(C.Structure):
    _fields_ = [("coef_off", C.c_uint32), ("bits", C.c_uint32), ("mv", (C.c_int16 * 2) * 2)]


mb_dtype = np.dtype([("coef_off", "<u4"), ("bits", "<u4"), ("mv", "<i2", (2, 2))])
assert mb_dtype.itemsize == 16 and C.sizeof(MbInfo) == 16

MB_INTRA, MB_FWD, MB_BWD = 1 << 29, 1 << 30, 1 << 31
MB_FIELD_DCT = 1 << 31      # in coef_off


def mb_coef_off(coef_off):
    return coef_off & 0x7fffffff

COEF_RAW, COEF_FIRST = 1 << 26, 1 << 27


def mb_ncoef(bits):
    return bits & 0x3ff


def mb_qscale(bits):
    return (bits >> 10) & 0x7f


def mb_cbp(bits):
    return (bits >> 17) & 0xfff


class PicParams(C.Structure):
    _fields_ = [("W", (C.c_uint8 * 64) * 4), ("picture_coding_type", C.c_int32), ("alternate_scan", C.c_int32),
                ("dst_frame", C.c_int32), ("l0_frame", C.c_int32), ("l1_frame", C.c_int32), ("n_coef", C.c_uint32),
                ("reserved", C.c_uint32 * 2)]


class Picture(C.Structure):
    _fields_ = [("params", C.POINTER(PicParams)), ("mb", C.POINTER(MbInfo)), ("coef", C.POINTER(C.c_uint32)),
                ("mb_count", C.c_uint32), ("coef_capacity", C.c_uint32), ("slot", C.c_int32), ("reserved", C.c_int32)]


class ReconConfig(C.Structure):
    _fields_ = [("device", C.c_int32), ("width", C.c_int32), ("height", C.c_int32), ("chroma_format", C.c_int32),
                ("n_frames", C.c_int32), ("n_pictures", C.c_int32), ("max_batch", C.c_int32), ("flags", C.c_int32),
                ("coef_capacity", C.c_uint32), ("bitstream_capacity", C.c_uint32)]


class FrameLayout(C.Structure):
    _fields_ = [("width", C.c_int32 * 3), ("height", C.c_int32 * 3), ("stride", C.c_int32 * 3),
                ("plane_offset", C.c_size_t * 3), ("bytes", C.c_size_t)]


class ReconStats(C.Structure):
    _fields_ = [("pictures", C.c_uint64), ("launches", C.c_uint64), ("h2d_bytes", C.c_uint64), ("d2h_bytes", C.c_uint64),
                ("algorithmic_bytes", C.c_uint64), ("kernel_ms", C.c_double),
                ("vlc_launches", C.c_uint64), ("vlc_slices", C.c_uint64), ("vlc_coefs", C.c_uint64),
                ("idct_batches", C.c_uint64), ("idct_exact_pass2", C.c_uint64), ("idct_exact_pass1", C.c_uint64)]


class PicSyntax(C.Structure):
    _fields_ = [("f_code", (C.c_int32 * 2) * 2), ("intra_dc_precision", C.c_int32), ("q_scale_type", C.c_int32),
                ("intra_vlc_format", C.c_int32), ("field_dct_syntax", C.c_int32)]


class SliceRef(C.Structure):
    _fields_ = [("payload", C.c_void_p), ("bytes", C.c_uint32), ("code", C.c_int32)]


RECON_VALIDATE = 1
RECON_DEVICE_VLC = 2
RECON_AUTO_DOWNLOAD = 4
RECON_THROUGHPUT = 8
OK, ERR_ARG, ERR_CUDA, ERR_NOMEM, ERR_STATE, ERR_RANGE = 0, -1, -2, -3, -4, -5
