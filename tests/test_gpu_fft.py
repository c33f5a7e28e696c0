"""The acquisition path's FFT engine alone (sgx_fft_c2c test hook) against numpy.fft -- the np.fft.fft / ifft call
sites of acquisition.py:95-126,182.  Covers the mixed-radix plans of the sampling-rate sweep (BASELINE config 5), the
38192-point search length, and the power-of-two plans (balanced passes; 2^19 is the fine search's sub-transform)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

#   length: why
SIZES = {
    2: "single radix-2 pass", 32: "16 x 2", 512: "two passes (16 x 2, 16)", 1024: "two passes of 32",
    4000: "fs = 4 MHz (2^5 x 5^3)", 16368: "fs = 16.3676 MHz rounded (2^4 x 3 x 11 x 31)", 38192: "search length (217 x 176)",
    64000: "fs = 64 MHz", 2 ** 19: "fine-search sub-transform (128 x 64 x 64)",
    381920: "10 ms coherent at 38.192 MHz (config 3): three passes, reordered 80 x 62 x 77 for the persistent kernels",
    320000: "5 ms at 64 MHz (80 x 50 x 80)",
}
# float32 transform against float64 numpy: relative l2 error per row (observed 1e-7 .. 4e-7)
REL_L2 = 2e-6


@pytest.mark.parametrize("n", sorted(SIZES))
@pytest.mark.parametrize("inverse", [False, True])
@pytest.mark.parametrize("persistent", [False, True])
def test_fft_engine_matches_numpy(n, inverse, persistent):
    from softgnss_python_b200 import _native
    L = _native.lib()
    rng = np.random.default_rng(n + int(inverse))
    batch = 3 if n < 100000 else 2
    x = (rng.standard_normal((batch, n)) + 1j * rng.standard_normal((batch, n))).astype(np.complex64)
    x[0, :] = 0
    x[0, n // 3] = 1.0                                            # a unit impulse: every twiddle appears once
    y = L.fft(x, inverse=inverse, persistent=persistent)   # persistent: the search's pass kernels (compile-time radices for 38192 / 2^19)
    x64 = x.astype(np.complex128)
    ref = np.fft.ifft(x64, axis=1) * n if inverse else np.fft.fft(x64, axis=1)     # unnormalised in both directions
    err = np.linalg.norm(y - ref, axis=1) / np.linalg.norm(ref, axis=1)
    assert err.max() < REL_L2, (n, inverse, err)
    assert np.abs(np.abs(y[0]) - 1.0).max() < 1e-5               # impulse: flat magnitude spectrum


def test_fft_rejects_unsupported_length():
    from softgnss_python_b200 import _native
    with pytest.raises(_native.NativeError):
        _native.lib().fft(np.zeros((1, 13 * 17), dtype=np.complex64))          # prime factors outside {2,3,5,7,11,31}
