"""The C-ABI library loads on a CPU-only box and exports every symbol include/*.h declares
(no compute calls); record layouts match the Python mirrors; creation fails loudly without a GPU."""
import ctypes as C
import os
import re

import pytest

from tiny_mp2v_dec_b200 import abi, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols(header):
    src = open(os.path.join(ROOT, "include", header)).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"MP2V_API\s+[\w\s\*]+?\b(mp2v_\w+)\s*\(", src)))


@pytest.mark.parametrize("header", ["mp2v_recon.h", "mp2v_decode_c.h"])
def test_library_exports_every_declared_symbol(header):
    L = C.CDLL(build.PRODUCT_LIB)
    names = declared_symbols(header)
    assert len(names) >= 7
    for n in names:
        assert hasattr(L, n), "libmp2v_b200.so does not export %s declared in include/%s" % (n, header)


def test_python_bindings_cover_the_headers():
    from tiny_mp2v_dec_b200 import decoder, recon
    assert sorted(recon.EXPORTS) == declared_symbols("mp2v_recon.h")
    assert sorted(decoder.DECODE_EXPORTS) == declared_symbols("mp2v_decode_c.h")


def test_record_layouts():
    assert C.sizeof(abi.MbInfo) == 16 and abi.mb_dtype.itemsize == 16
    assert C.sizeof(abi.PicParams) == 256 + 4 * 8
    assert C.sizeof(abi.ReconConfig) == 40
    bits = (5 & 0x3ff) | (112 << 10) | (0xabc << 17) | abi.MB_FWD
    assert abi.mb_ncoef(bits) == 5 and abi.mb_qscale(bits) == 112 and abi.mb_cbp(bits) == 0xabc


def test_frame_layout_is_the_reference_rule():
    """frame_c (decoder.cpp:44-66): stride = align64(width), chroma per format"""
    from tiny_mp2v_dec_b200.recon import frame_layout
    import oracle_lib as O
    for w, h, cf, want in [(1920, 1088, 1, ([1920, 960, 960], [1088, 544, 544], [1920, 960, 960])),
                           (1920, 1088, 2, ([1920, 960, 960], [1088, 1088, 1088], [1920, 960, 960])),
                           (3840, 2160, 3, ([3840] * 3, [2160] * 3, [3840] * 3)),
                           (48, 32, 1, ([48, 24, 24], [32, 16, 16], [64, 64, 64])),
                           (176, 144, 2, ([176, 88, 88], [144, 144, 144], [192, 128, 128]))]:
        lay = frame_layout(w, h, cf)
        assert (list(lay.width), list(lay.height), list(lay.stride)) == want
        olay = O.frame_layout(w, h, cf)
        assert list(olay.stride) == list(lay.stride) and olay.bytes == lay.bytes
    with pytest.raises(Exception):
        frame_layout(100, 64, 1)


def test_no_gpu_no_fallback():
    """without a usable CUDA device the product refuses to run: there is no CPU reconstruction path"""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from tiny_mp2v_dec_b200.recon import Recon, ReconError
    with pytest.raises(ReconError) as e:
        Recon(64, 48, 1)
    assert "CUDA" in str(e.value) or "device" in str(e.value)
    from tiny_mp2v_dec_b200.decoder import Decoder
    from tiny_mp2v_dec_b200.streamgen import Stream
    s = Stream(64, 48, 1, seed=1, gop_n=2, gop_m=1)
    with pytest.raises(ReconError):
        Decoder(64, 48, 1, num_threads=1).decode(s.padded, s.size)


def test_numa_cpu_list_parser():
    """sysfs cpulist format (host/numa.cpp): ranges and singles, malformed input gives nothing"""
    L = C.CDLL(build.PRODUCT_LIB)
    L.mp2v_numa_parse_cpu_list.argtypes = [C.c_char_p, C.POINTER(C.c_int32), C.c_int]
    out = (C.c_int32 * 64)()

    def parse(text):
        n = L.mp2v_numa_parse_cpu_list(text.encode(), out, 64)
        return list(out[:min(n, 64)]) if n else []
    assert parse("0-3,8,10-11\n") == [0, 1, 2, 3, 8, 10, 11]
    assert parse("5") == [5]
    assert parse("0-23") == list(range(24))
    assert parse(" 1 , 3-4 ") == [1, 3, 4]
    assert parse("") == [] and parse("3-1") == [] and parse("a-b") == [] and parse("1,,2") == [] and parse("1-") == []
    assert L.mp2v_numa_parse_cpu_list(b"0-127", out, 64) == 128      # counts past the caller's capacity, stores 64
    assert L.mp2v_recon_numa_node(None) == -1
