"""Developer check: preamble search -> ephemeris decoding -> measurement loop (postnav.post_navigate_batch) under the
CPU fiber emulator against the oracle chain (same comparison as tests/test_gpu_nav.py::test_post_navigate_chain)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from softgnss_python_b200 import _native, postnav              # noqa: E402
from softgnss_python_b200.settings import Settings             # noqa: E402
from tests.cases import NAV_MS, build_chain_case, oracle_chain  # noqa: E402
from tests import nav_util                                      # noqa: E402

_native._LIB = _native.Lib(os.environ.get("SGX_EMUL_LIB", os.path.join(ROOT, "tools", "cpu_emul", "libsoftgnss_emul.so")))
tr0, prn0, truth0 = build_chain_case(seed=2)
tr1, prn1, truth1 = build_chain_case(seed=5, drop=(1, 4, 6, 7, 2))      # three satellites left: no solution
s = Settings(numberOfChannels=8, msToProcess=float(NAV_MS))
out = postnav.post_navigate_batch(np.stack([tr0, tr1]), np.stack([prn0, prn1]), s)
print("subFrameStart", out["subFrameStart"].tolist(), "epochs", out["n_epochs"].tolist(), "tow", out["tow"].tolist())
f0, o0 = oracle_chain(tr0, prn0, dict(elevation_mask=s.elevationMask, use_trop_corr=s.useTropCorr))
f1, o1 = oracle_chain(tr1, prn1, dict(elevation_mask=s.elevationMask, use_trop_corr=s.useTropCorr))
assert np.array_equal(out["subFrameStart"][0], f0) and np.array_equal(out["subFrameStart"][1], f1)
assert o1 is None and out["n_epochs"][1] == 0
print(nav_util.compare_nav(out, 0, o0, "chain"))
sol = out["sol"][0, :out["n_epochs"][0]]
err = np.sqrt(((sol[:, :3] - truth0["rx"]) ** 2).sum(1))
print("3-D error to the true antenna: median %.1f m, max %.1f m" % (np.median(err), err.max()))
