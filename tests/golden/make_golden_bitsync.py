"""Golden vectors for the preamble search / bit summation, produced by the
REFERENCE's own ``NavigationResult.findPreambles`` (postNavigation.py:524-631) and the bit summation of
``postNavigate`` (postNavigation.py:125-134) through the Python-3 shim.  Build container only:

    python tests/golden/make_golden_bitsync.py

Inputs are regenerated from seeds (tests/cases.py: integer-only construction); their SHA-1 is stored."""
import contextlib
import hashlib
import io
import os
import sys

os.environ.setdefault("OPENBLAS_NUM_THREADS", "1")   # np.correlate = 74 000 small BLAS dots: threads only spin
os.environ.setdefault("OMP_NUM_THREADS", "1")

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import make_ref_shim                                  # noqa: E402
from oracle.gnss_oracle import TRACK_FIELDS                      # noqa: E402
from tests.cases import build_bitsync_case, build_pseudo_case, BITSYNC_MS          # noqa: E402


def main():
    ref = make_ref_shim.import_ref()
    ips = build_bitsync_case()
    s = ref["initialize"].Settings()
    s.msToProcess = float(BITSYNC_MS)
    s.numberOfChannels = len(ips)
    dtype = [('status', 'S1')] + [(f, 'object') for f in TRACK_FIELDS] + [('PRN', 'int64')]
    zero = np.zeros(BITSYNC_MS)
    rec = [(b'T',) + tuple(ips[c] if f == "I_P" else zero for f in TRACK_FIELDS) + (c + 1,) for c in range(len(ips))]
    res = np.rec.fromrecords(rec, dtype=dtype)

    class T(object):
        results = res
        channels = None
        settings = s
    nav = ref["postNavigation"].NavigationResult(T())
    with contextlib.redirect_stdout(io.StringIO()):
        first, active = nav.findPreambles()
    bits = np.zeros((len(ips), 1501), dtype=np.uint8)
    bits_valid = np.zeros(len(ips), dtype=np.uint8)
    for ch in active:                                             # postNavigation.py:125-134, verbatim
        if first[ch] + 1500 * 20 > BITSYNC_MS:                    # (the reference's reshape raises here)
            continue
        bits_valid[ch] = 1
        x = res[ch].I_P[first[ch] - 20:first[ch] + 1500 * 20].copy()
        x = x.reshape(20, -1, order='F')
        bits[ch] = (x.sum(0) > 0) * 1
    np.savez_compressed(os.path.join(HERE, "bitsync.npz"),
                        first=np.asarray(first, dtype=np.int64), active=np.asarray(active, dtype=np.int64),
                        nav_bits=np.packbits(bits, axis=1), bits_valid=bits_valid,
                        input_sha1=hashlib.sha1(np.ascontiguousarray(np.stack(ips)).tobytes()).hexdigest())
    print("firstSubFrame", first, "active", active)

    # ---- calculatePseudoranges (postNavigation.py:27-72) on synthetic absoluteSample series ----------------
    abs_sample, ms_index, act = build_pseudo_case()
    n_ch, n_ms = abs_sample.shape
    s2 = ref["initialize"].Settings()
    s2.numberOfChannels = n_ch
    zero = np.zeros(n_ms)
    rec = [(b'T',) + tuple(abs_sample[c] if f == "absoluteSample" else zero for f in TRACK_FIELDS) + (c + 1,)
           for c in range(n_ch)]
    res2 = np.rec.fromrecords(rec, dtype=dtype)

    class T2(object):
        results = res2
        channels = None
        settings = s2
    nav2 = ref["postNavigation"].NavigationResult(T2())
    pr = np.zeros(ms_index.shape)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for e in range(ms_index.shape[0]):
            pr[e] = nav2.calculatePseudoranges(ms_index[e].astype(float), act[e].nonzero()[0])
    np.savez_compressed(os.path.join(HERE, "pseudo.npz"), pseudoranges=pr,
                        input_sha1=hashlib.sha1(np.ascontiguousarray(abs_sample).tobytes()).hexdigest())
    print("pseudoranges epoch 0:", pr[0][:3], "epoch 3 (empty list):", pr[3][:2])


if __name__ == "__main__":
    main()
