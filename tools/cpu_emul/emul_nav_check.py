"""Developer check: run the navigation-solution kernel under the CPU fiber emulator against the reference's golden
output and the oracle (same comparison as tests/test_gpu_nav.py)."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from softgnss_python_b200 import _native, postnav              # noqa: E402
from oracle import gnss_oracle as orc                           # noqa: E402
from tests.cases import NAV_MS, load_nav_cases, nav_abs_sample  # noqa: E402
from tests import nav_util                                      # noqa: E402

L = _native.Lib(os.environ.get("SGX_EMUL_LIB", os.path.join(ROOT, "tools", "cpu_emul", "libsoftgnss_emul.so")))
_native._LIB = L
for idx, case in enumerate(load_nav_cases()):
    n_ch = len(case["prn"])
    s = nav_util.settings_for(case, n_ch, NAV_MS)
    sfs, ready, eph = nav_util.case_inputs(case)
    abs_sample = nav_abs_sample(case["coef"])
    t = time.time()
    out = postnav.nav_solve_batch(abs_sample[None], sfs[None], ready[None], eph[None], [case["tow"]], s)
    dt = time.time() - t
    worst = nav_util.compare_nav(out, 0, case["ref"], "case %d vs reference" % idx)
    o = orc.nav_solve(abs_sample, case["prn"], case["sub_frame_start"], case["ready"], case["eph"], case["tow"],
                      float(NAV_MS), 38192, elevation_mask=case["elevation_mask"], use_trop_corr=case["use_trop_corr"])
    w2 = nav_util.compare_nav(out, 0, o, "case %d vs oracle" % idx)
    print("case %d ok (%.2fs): worst errors" % (idx, dt), {k: float("%.2g" % v) for k, v in w2.items()})
