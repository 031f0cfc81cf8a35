"""Annex B tables: internal consistency everywhere, and row-by-row equality with the reference's
forward tables (mp2v_luts.hpp) where the reference tree is mounted."""
import os
import re
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HDR = os.path.join(ROOT, "tiny_mp2v_dec_b200", "csrc", "host", "vlc_tables.h")


def own_tables():
    src = open(HDR).read()
    out = {}
    for name in re.findall(r"static const vlc_row_t (\w+)\[\]", src):
        body = re.search(r"%s\[\] = \{(.*?)\n\};" % name, src, re.S).group(1)
        out[name] = [(b, int(a, 0), int(c, 0)) for b, a, c in re.findall(r'\{"([01]+)",\s*(-?\w+),\s*(-?\w+)\}', body)]
    return out


def prefix_free(codes):
    codes = sorted(codes)
    return all(not b.startswith(a) for a, b in zip(codes, codes[1:]))


def test_tables_are_prefix_free_and_complete():
    t = own_tables()
    assert len(t["kTabMbAddrInc"]) == 34 and len(t["kTabCbp"]) == 64 and len(t["kTabMotionCode"]) == 17
    assert len(t["kTabCoefB14"]) == 111 and len(t["kTabCoefB15"]) == 111
    assert prefix_free([r[0] for r in t["kTabMbAddrInc"]])
    assert prefix_free([r[0] for r in t["kTabCbp"]])
    assert prefix_free([r[0] for r in t["kTabMotionCode"]])
    for pct in (1, 2, 3):
        assert prefix_free([r[0] for r in t["kTabMbType"] if r[2] == pct])
    for lc in (0, 1):
        assert prefix_free([r[0] for r in t["kTabDcSize"] if r[2] == lc])
    assert prefix_free([r[0] for r in t["kTabCoefB14"]] + ["10", "000001"])
    assert prefix_free([r[0] for r in t["kTabCoefB15"]] + ["0110", "000001"])
    # both run/level tables cover the same (run, level) pairs
    assert sorted((a, b) for _, a, b in t["kTabCoefB14"]) == sorted((a, b) for _, a, b in t["kTabCoefB15"])


@pytest.mark.skipif(not os.path.exists("/root/reference/src/core/mp2v_luts.hpp"), reason="reference tree not mounted")
def test_tables_equal_the_reference_forward_tables():
    sys.path.insert(0, os.path.join(ROOT, "tools", "dev"))
    import ref_vlc_tables
    r = ref_vlc_tables.load()
    t = own_tables()
    assert {a: b for b, a, _ in t["kTabMbAddrInc"] if a} == {i: r["mba"][i] for i in range(1, 34)}
    assert {a: b for b, a, _ in t["kTabCbp"]} == dict(enumerate(r["cbp"]))
    for b, mag, _ in t["kTabMotionCode"]:
        if mag == 0:
            assert r["motion"][16] == b
        else:
            assert r["motion"][16 + mag] == b + "0" and r["motion"][16 - mag] == b + "1"
    assert {(lc, a): b for b, a, lc in t["kTabDcSize"]} == {**{(0, i): c for i, c in enumerate(r["dc_luma"])},
                                                             **{(1, i): c for i, c in enumerate(r["dc_chroma"])}}
    assert {(a, b): c for c, a, b in t["kTabCoefB14"]} == r["b14"]
    assert {(a, b): c for c, a, b in t["kTabCoefB15"]} == r["b15"]
    for pct, name in ((1, "mbtype_i"), (2, "mbtype_p"), (3, "mbtype_b")):
        assert {a: b for b, a, p in t["kTabMbType"] if p == pct} == r[name]
