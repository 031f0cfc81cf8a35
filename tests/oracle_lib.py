"""Test-side bindings of the CHECKERS: oracle/libmp2v_oracle.so (C restatement) and, when it was
built, oracle/_ref/libmp2v_ref.so (the unmodified reference).  Test infrastructure only."""
import ctypes as C
import os
import subprocess

import numpy as np

from tiny_mp2v_dec_b200.abi import FrameLayout, MbInfo, PicParams

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_LIB = os.path.join(ROOT, "oracle", "libmp2v_oracle.so")
REF_LIB = os.path.join(ROOT, "oracle", "_ref", "libmp2v_ref.so")

_oracle = None
_ref = None

U8P = C.POINTER(C.c_uint8)


def oracle():
    global _oracle
    if _oracle is None:
        if not os.path.exists(ORACLE_LIB):
            subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "oracle"], stdout=subprocess.DEVNULL)
        L = C.CDLL(ORACLE_LIB)
        L.orc_recon_picture.argtypes = [C.POINTER(PicParams), C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                        U8P * 3, U8P * 3, U8P * 3]
        L.orc_idct_sse2.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_dequant_block.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        L.orc_mc_fetch.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
        L.orc_build_W.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.orc_scan_tables.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.mp2v_frame_layout.argtypes = [C.c_int, C.c_int, C.c_int, C.POINTER(FrameLayout)]
        _oracle = L
    return _oracle


def have_ref():
    return os.path.exists(REF_LIB)


def ref():
    global _ref
    if _ref is None:
        L = C.CDLL(REF_LIB)
        L.ref_decode_serial.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_size_t,
                                        C.POINTER(C.c_size_t), C.POINTER(C.c_uint64)]
        L.ref_decode_mt.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                    C.c_size_t, C.POINTER(C.c_size_t), C.POINTER(C.c_uint64), C.POINTER(C.c_double)]
        _ref = L
    return _ref


def frame_layout(width, height, cf):
    lay = FrameLayout()
    assert oracle().mp2v_frame_layout(width, height, cf, C.byref(lay)) == 0
    return lay


def yuv_frame_bytes(width, height, cf):
    return width * height * {1: 3, 2: 4, 3: 6}[cf] // 2


class Frame:
    """A host frame with the reference's frame_c geometry (padded strides)."""

    def __init__(self, width, height, cf, fill=None):
        self.lay = frame_layout(width, height, cf)
        self.planes = []
        for p in range(3):
            a = np.zeros((self.lay.height[p], self.lay.stride[p]), np.uint8)
            if fill is not None:
                a[:] = fill
            self.planes.append(a)

    def ptrs(self):
        return (U8P * 3)(*[p.ctypes.data_as(U8P) for p in self.planes])

    def cropped(self):
        return b"".join(self.planes[p][:, :self.lay.width[p]].tobytes() for p in range(3))


def fnv1a64(data):
    """same hash as oracle/ref_driver.cpp's sink (vectorised is not possible; small inputs only)"""
    h = 1469598103934665603
    for b in data:
        h = ((h ^ b) * 1099511628211) & 0xffffffffffffffff
    return h


def oracle_decode_stream(stream):
    """Reconstruct every picture of a generated stream from its GROUND-TRUTH records with the C
    oracle; returns the cropped planar YUV of all frames in display order."""
    L = oracle()
    frames = {}
    for idx, pic in enumerate(stream.pictures):
        dst = Frame(stream.width, stream.height, stream.chroma_format)
        l0 = frames.get(pic.params.l0_frame)
        l1 = frames.get(pic.params.l1_frame)
        null = (U8P * 3)()
        rc = L.orc_recon_picture(C.byref(pic.params), pic.mb.ctypes.data, pic.coef.ctypes.data,
                                 stream.width, stream.height, stream.chroma_format,
                                 dst.ptrs(), l0.ptrs() if l0 else null, l1.ptrs() if l1 else null)
        assert rc == 0, "oracle recon failed rc=%d picture %d" % (rc, idx)
        frames[idx] = dst
    return b"".join(frames[i].cropped() for i in stream.display_order())


def ref_decode_serial(stream):
    """The unmodified reference, serial driver; returns cropped planar YUV in display order."""
    L = ref()
    n_frames = len(stream.pictures)
    cap = n_frames * yuv_frame_bytes(stream.width, stream.height, stream.chroma_format)
    out = np.zeros(cap, np.uint8)
    nbytes = C.c_size_t()
    h = C.c_uint64()
    buf = stream.padded.copy()
    got = L.ref_decode_serial(buf.ctypes.data, stream.size, stream.width, stream.height, stream.chroma_format,
                              out.ctypes.data, cap, C.byref(nbytes), C.byref(h))
    assert got == n_frames and nbytes.value == cap, (got, n_frames, nbytes.value, cap)
    return out.tobytes()


def ref_decode_mt(stream, threads=8, pool=10, want_output=True):
    L = ref()
    n_frames = len(stream.pictures)
    cap = n_frames * yuv_frame_bytes(stream.width, stream.height, stream.chroma_format) if want_output else 0
    out = np.zeros(max(cap, 1), np.uint8)
    nbytes = C.c_size_t()
    h = C.c_uint64()
    secs = C.c_double()
    buf = stream.padded.copy()
    got = L.ref_decode_mt(buf.ctypes.data, stream.size, stream.width, stream.height, stream.chroma_format, pool, threads,
                          1 if want_output else 0, out.ctypes.data, cap, C.byref(nbytes), C.byref(h), C.byref(secs))
    return got, (out[:cap].tobytes() if want_output else None), secs.value
