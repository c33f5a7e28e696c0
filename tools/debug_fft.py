import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from softgnss_python_b200 import _native
if len(sys.argv) > 1 and sys.argv[1] == "emul":
    L = _native.Lib(os.path.join(ROOT, "tools", "cpu_emul", "libsoftgnss_emul.so"))
else:
    L = _native.lib()
rng = np.random.default_rng(0)
for n in (16, 64, 256, 4096, 65536, 38192, 16368, 4000, 1 << 18, 1 << 22):
    if len(sys.argv) > 1 and sys.argv[1] == "emul" and n > 70000: continue
    x = (rng.standard_normal((2, n)) + 1j * rng.standard_normal((2, n))).astype(np.complex64)
    for inv in (False, True):
        y = L.fft(x, inverse=inv)
        ref = np.fft.ifft(x.astype(np.complex128), axis=1) * n if inv else np.fft.fft(x.astype(np.complex128), axis=1)
        err = np.abs(y - ref).max() / np.abs(ref).max()
        print("n=%8d inv=%d rel err %.2e" % (n, inv, err))
