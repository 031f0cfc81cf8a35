"""Shared test helpers."""
import ctypes as C
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))

import oracle_lib as O  # noqa: E402
from cases import CASES as GOLDEN_CASES  # noqa: E402

GOLDEN = json.load(open(os.path.join(HERE, "golden", "golden.json")))


def sha(b):
    return hashlib.sha256(b).hexdigest()


def display_order_of(types):
    """coded order -> display order with the reference's reorder (decoder.cpp:350-378):
    B pictures at once, I/P pictures when the next I/P (or the end) arrives."""
    order, held = [], None
    for i, t in enumerate(types):
        if t == 3:
            order.append(i)
        else:
            if held is not None:
                order.append(held)
            held = i
    if held is not None:
        order.append(held)
    return order


def oracle_decode_parsed(pics, width, height, cf):
    """C-oracle reconstruction of host-parser output (ParsedPicture list) -> cropped YUV, display order"""
    L = O.oracle()
    frames = {}
    null = (O.U8P * 3)()
    for i, p in enumerate(pics):
        dst = O.Frame(width, height, cf)
        l0, l1 = frames.get(p.params.l0_frame), frames.get(p.params.l1_frame)
        rc = L.orc_recon_picture(C.byref(p.params), p.mb.ctypes.data, p.coef.ctypes.data, width, height, cf,
                                 dst.ptrs(), l0.ptrs() if l0 else null, l1.ptrs() if l1 else null)
        assert rc == 0, (rc, i)
        frames[i] = dst
    order = display_order_of([p.params.picture_coding_type for p in pics])
    return b"".join(frames[i].cropped() for i in order)
