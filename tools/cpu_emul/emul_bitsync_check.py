"""Developer check: run the bit-sync kernel under the CPU fiber emulator against the reference's golden output."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from softgnss_python_b200 import _native                       # noqa: E402
from tests.cases import build_bitsync_case, build_bitsync_channel, BITSYNC_EARLY, BITSYNC_MS   # noqa: E402

L = _native.Lib(os.environ.get("SGX_EMUL_LIB", os.path.join(ROOT, "tools", "cpu_emul", "libsoftgnss_emul.so")))
g = np.load(os.path.join(ROOT, "tests", "golden", "bitsync.npz"))
ips = np.stack(build_bitsync_case() + [build_bitsync_channel(BITSYNC_EARLY, 600)])
t = time.time()
first, bits, valid = L.find_preambles(ips, ips.shape[1], ips.shape[0], BITSYNC_MS)
print("first", first.tolist(), "%.1fs" % (time.time() - t))
print("golden", g["first"].tolist(), "match:", np.array_equal(first[:-1], g["first"]), "early ->", first[-1], "(oracle 6020)")
gb = np.unpackbits(g["nav_bits"], axis=1)[:, :1501]
print("bits_valid", valid.tolist(), "golden", g["bits_valid"].tolist())
ok = all(np.array_equal(bits[c], gb[c]) for c in range(len(gb)) if g["bits_valid"][c])
print("nav bits identical:", ok)
