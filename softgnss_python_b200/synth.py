"""Synthetic GPS L1 C/A int8 IF recordings with known truth.

The reference's data file (``initialize.py:99``) is not distributed, so every parity
and throughput run uses recordings produced by this integer-only signal model.  It is
evaluated by two bit-identical implementations: ``generate_cpu`` (numpy, here) and the
``sgx_synth_generate`` CUDA kernel (``csrc/sgx_synth.cu``) -- identical because every
step is integer arithmetic with explicit wrap-around:

  sample[n] = clip( ( sum_s A_s * d_s(n) * c_s(n) * COS[ph_s(n) >> 52]
                      + K * (bytesum(splitmix64(seed + n*GOLDEN)) - 1020) + 2^21 ) >> 22 )

  ph_s(n) = phi0_s + n*dphi_s          (uint64, 2^64 = one carrier cycle)
  cp_s(n) = cp0_s  + n*dcp_s           (uint64, 32.32 fixed-point chips)
  c_s     = C/A chip   (cp >> 32) mod 1023,  d_s = nav bit of code period (cp>>32)/1023
  COS     = round(2^14 cos(2 pi k / 4096)),  A_s = round(256 * amplitude in LSB)
  noise   : Irwin-Hall sum of the 8 bytes of a counter hash, sigma = K * 209.0215 / 2^22 LSB

File format is what the reference reads: raw int8, real IF, one byte per sample
(``initialize.py:102``, ``:481``; ``tracking.py:107``).
"""
from fractions import Fraction

import numpy as np

from .settings import ca_code_bits

LUT_BITS = 12
LUT_SCALE = 1 << 14
AMP_SCALE = 1 << 8
OUT_SHIFT = 22                      # LUT_SCALE * AMP_SCALE
GOLDEN = 0x9E3779B97F4A7C15
HASH_SIGMA = float(np.sqrt(8 * (256 ** 2 - 1) / 12.0))   # std of the sum of 8 uniform bytes
L1_HZ = 1575.42e6
MAX_SATS = 12
M64 = (1 << 64) - 1


def cos_lut():
    k = np.arange(1 << LUT_BITS)
    return np.round(LUT_SCALE * np.cos(2 * np.pi * k / (1 << LUT_BITS))).astype(np.int16)


class SatSpec(object):
    """One satellite of a recording.  ``code_phase`` = sample index (0..N-1) at which a code
    period starts; ``doppler`` in Hz; ``cn0`` in dB-Hz; ``nav_bits`` +-1 per 20 ms bit;
    ``bit_offset_ms`` = code periods already elapsed in the current bit at that start."""

    def __init__(self, prn, doppler, code_phase, cn0=45.0, nav_bits=None, bit_offset_ms=0,
                 carrier_phase=0.0):
        self.prn = int(prn)
        self.doppler = float(doppler)
        self.code_phase = int(code_phase)
        self.cn0 = float(cn0)
        self.nav_bits = None if nav_bits is None else np.asarray(nav_bits, dtype=np.int8)
        self.bit_offset_ms = int(bit_offset_ms)
        self.carrier_phase = float(carrier_phase)


class RecordingSpec(object):
    """Integer parameter block shared by the CPU and CUDA generators."""

    def __init__(self, sats, fs=38.192e6, f_if=9.548e6, sigma=12.0, seed=1, n_bits=2048):
        assert 1 <= len(sats) <= MAX_SATS
        self.fs, self.f_if, self.sigma, self.seed = float(fs), float(f_if), float(sigma), int(seed)
        self.sats = list(sats)
        ns = len(sats)
        self.n_bits = int(n_bits)
        self.prn = np.zeros(ns, dtype=np.int32)
        self.amp = np.zeros(ns, dtype=np.int32)
        self.phi0 = np.zeros(ns, dtype=np.uint64)
        self.dphi = np.zeros(ns, dtype=np.uint64)
        self.cp0 = np.zeros(ns, dtype=np.uint64)
        self.dcp = np.zeros(ns, dtype=np.uint64)
        self.per0 = np.zeros(ns, dtype=np.int32)      # code periods elapsed in bit 0 at cp=0
        self.bits = np.ones((ns, self.n_bits), dtype=np.int8)
        rng = np.random.default_rng(self.seed + 7919)
        code_mod = 1023 << 32
        for i, s in enumerate(sats):
            self.prn[i] = s.prn
            a = 2.0 * self.sigma * np.sqrt(10.0 ** (s.cn0 / 10.0) / self.fs)
            self.amp[i] = int(round(a * AMP_SCALE))
            f = Fraction(self.f_if) + Fraction(s.doppler)
            self.dphi[i] = int(round(f / Fraction(self.fs) * (1 << 64))) & M64
            self.phi0[i] = int(round((s.carrier_phase % 1.0) * (1 << 64))) & M64
            fcode = Fraction(1023000) * (1 + Fraction(s.doppler) / Fraction(L1_HZ))
            dcp = int(round(fcode / Fraction(self.fs) * (1 << 32)))
            self.dcp[i] = dcp
            sync = getattr(s, "frame_sync", None)
            if sync is not None:
                # (sample index n_s at which nav bit number b0 of the stream starts, b0): the code period
                # that starts at n_s is the first of that bit
                n_s, b0 = int(sync[0]), int(sync[1])
                self.cp0[i] = (code_mod - (n_s * dcp) % code_mod) % code_mod
                periods_at_ns = (int(self.cp0[i]) + n_s * dcp) // code_mod
                self.per0[i] = 20 * b0 - periods_at_ns
                assert self.per0[i] >= 0, "nav stream must start before sample 0"
                cp0 = int(self.cp0[i])
                s.code_phase = 0 if cp0 == 0 else -(-(code_mod - cp0) // dcp)   # first period start >= 0
            else:
                # a code period starts exactly at sample code_phase: cp0 + code_phase*dcp == 0 (mod 1023
                # chips); one full period is added so that cp never has to go negative.
                self.cp0[i] = (code_mod - (s.code_phase * dcp) % code_mod) % code_mod
                # the period running at n=0 is period 0; the one starting at code_phase is period 1
                # (or 0 when code_phase == 0).  bit index = (period + per0) // 20.
                first = 0 if s.code_phase == 0 or self.cp0[i] == 0 else 1
                self.per0[i] = (s.bit_offset_ms - first) % 20
            if s.nav_bits is not None:
                nb = min(len(s.nav_bits), self.n_bits)
                self.bits[i, :nb] = s.nav_bits[:nb]
            else:
                self.bits[i] = rng.integers(0, 2, self.n_bits).astype(np.int8) * 2 - 1
        self.noise_k = int(round(self.sigma / HASH_SIGMA * (1 << OUT_SHIFT)))
        worst = int(self.amp.astype(np.int64).sum()) * LUT_SCALE + self.noise_k * 1020 + (1 << 21)
        assert worst < (1 << 31), "amplitudes overflow the int32 accumulator of the CUDA generator"

    # ---- truth in the units the receiver reports -------------------------------------------
    def true_carr_freq(self, i):
        return float(Fraction(int(self.dphi[i]), 1 << 64) * Fraction(self.fs))

    def true_code_freq(self, i):
        return float(Fraction(int(self.dcp[i]), 1 << 32) * Fraction(self.fs))

    def nav_bit_at_period(self, i, period):
        return self.bits[i, ((period + int(self.per0[i])) // 20) % self.n_bits]


def _splitmix64(z):
    z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return z ^ (z >> np.uint64(31))


def generate_cpu(spec, n_samples, start=0, chunk=1 << 20):
    """int8[n_samples] for absolute sample indices start .. start+n_samples-1."""
    lut = cos_lut().astype(np.int64)
    codes = np.stack([ca_code_bits(p - 1).astype(np.int64) * 2 - 1 for p in spec.prn])
    out = np.empty(n_samples, dtype=np.int8)
    seed = np.uint64(spec.seed & M64)
    with np.errstate(over="ignore"):
        for lo in range(0, n_samples, chunk):
            hi = min(lo + chunk, n_samples)
            n = np.arange(start + lo, start + hi, dtype=np.uint64)
            acc = np.zeros(hi - lo, dtype=np.int64)
            for i in range(len(spec.prn)):
                ph = spec.phi0[i] + n * spec.dphi[i]
                cp = spec.cp0[i] + n * spec.dcp[i]
                chips = (cp >> np.uint64(32)).astype(np.int64)
                chip = chips % 1023
                period = chips // 1023
                bit = ((period + int(spec.per0[i])) // 20) % spec.n_bits
                sign = codes[i][chip] * spec.bits[i][bit].astype(np.int64)
                acc += int(spec.amp[i]) * sign * lut[(ph >> np.uint64(64 - LUT_BITS)).astype(np.int64)]
            h = _splitmix64(seed + n * np.uint64(GOLDEN))
            bsum = h.view(np.uint8).reshape(-1, 8).sum(axis=1).astype(np.int64)
            acc += spec.noise_k * (bsum - 1020)
            q = (acc + (1 << (OUT_SHIFT - 1))) >> OUT_SHIFT
            out[lo:hi] = np.clip(q, -128, 127).astype(np.int8)
    return out


def default_constellation(seed, n_sats=8, fs=38.192e6, cn0=45.0, n_code=38192, avoid=(37,)):
    """A reproducible set of satellites: PRNs, Doppler in +-6.5 kHz, code phase uniform in
    0..N-1 except the values the reference cannot report."""
    rng = np.random.default_rng(seed)
    prns = np.sort(rng.choice(np.arange(1, 33), size=n_sats, replace=False))
    sats = []
    for p in prns:
        while True:
            cph = int(rng.integers(0, n_code))
            if (cph + 1) % n_code not in avoid and cph not in avoid:
                break
        sats.append(SatSpec(p, float(rng.uniform(-6500, 6500)), cph, cn0=cn0,
                            bit_offset_ms=int(rng.integers(0, 20)),
                            carrier_phase=float(rng.uniform(0, 1))))
    return sats
