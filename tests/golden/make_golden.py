"""Regenerate tests/golden/golden.json.  Run in the authoring container, where the UNMODIFIED
reference has been compiled into oracle/_ref (oracle/Makefile): every entry records the SHA-256 of the
generated stream and of the YUV the real reference decoder (serial driver) produces from it, plus the
reference's own sample binary for the hard-wired 1920x1088 4:2:2 geometry.  The files travel; the
reference does not."""
import hashlib
import json
import os
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, HERE)

import oracle_lib as O  # noqa: E402
from cases import CASES  # noqa: E402
from tiny_mp2v_dec_b200.streamgen import Stream  # noqa: E402


def main():
    assert O.have_ref(), "oracle/_ref is missing: run `make -C oracle ref` where /root/reference is mounted"
    out = {}
    for name, (w, h, cf, kw) in CASES.items():
        s = Stream(w, h, cf, **kw)
        yuv = O.ref_decode_serial(s)
        entry = dict(width=w, height=h, chroma_format=cf, gen=kw, frames=len(s.pictures),
                     stream_sha256=hashlib.sha256(s.data.tobytes()).hexdigest(),
                     yuv_sha256=hashlib.sha256(yuv).hexdigest(), source="oracle/_ref serial driver over the unmodified reference")
        if (w, h, cf) == (1920, 1088, 2):
            # config 1 of BASELINE.json: the reference's own sample binary, 8 threads, YUV written
            with tempfile.TemporaryDirectory() as d:
                m2v, yo = os.path.join(d, "s.m2v"), os.path.join(d, "s.yuv")
                with open(m2v, "wb") as f:
                    f.write(s.data.tobytes())
                subprocess.check_call([os.path.join(ROOT, "oracle", "_ref", "tiny_mp2v_dec_sample"), "-v", m2v, "-o", yo],
                                      stdout=subprocess.DEVNULL)
                sample = open(yo, "rb").read()
            entry["sample_yuv_sha256"] = hashlib.sha256(sample).hexdigest()
            assert entry["sample_yuv_sha256"] == entry["yuv_sha256"], "reference sample binary and serial driver disagree"
        out[name] = entry
        print(name, entry["frames"], entry["yuv_sha256"][:16])
    # two tiny raw fixtures, small enough to commit, so that one parity check needs no generator at all
    s = Stream(48, 32, 1, **CASES["tiny420_m1"][3])
    with open(os.path.join(HERE, "tiny420_m1.m2v"), "wb") as f:
        f.write(s.data.tobytes())
    with open(os.path.join(HERE, "tiny420_m1.yuv"), "wb") as f:
        f.write(O.ref_decode_serial(s))
    json.dump(out, open(os.path.join(HERE, "golden.json"), "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
