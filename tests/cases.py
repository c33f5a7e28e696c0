"""Seeded synthetic test cases shared by the golden generator, the CPU oracle tests and the
GPU parity tests.  Inputs are always regenerated from these specs (never stored)."""
import numpy as np

from softgnss_python_b200 import synth

N = 38192  # samplesPerCode at the reference's default fs


def _sats(lst, cn0=47.0):
    return [synth.SatSpec(p, d, c, cn0=k.get("cn0", cn0), bit_offset_ms=k.get("bo", 3 * p % 20),
                          carrier_phase=(0.137 * p) % 1.0)
            for (p, d, c, k) in [(x[0], x[1], x[2], x[3] if len(x) > 3 else {}) for x in lst]]


CASES = {
    # BASELINE.json config 1: default acquisition on an 11 ms recording, 8 satellites
    "acq_c1": dict(seed=1, ms=0, n_ms=11, sats=_sats([
        (1, 1500.0, 1000), (3, -3250.0, 12345), (7, -2500.0, 40), (11, 6200.0, 38000),
        (14, 250.0, 20000), (19, -6400.0, 5), (22, 4000.0, 38181), (31, -7000.0, 27001)], cn0=45.0)),
    # code phases around the second-peak window branches (acquisition.py:147-159); the
    # reference reports start+1 (appendix A.1-9), 37 is its IndexError and is avoided here
    "acq_edges": dict(seed=2, ms=0, n_ms=11, sats=_sats([
        (2, 0.0, 0), (5, 3000.0, 35), (9, -3000.0, 37), (12, 7000.0, N - 40),
        (17, -500.0, N - 39), (25, 500.0, N - 2)], cn0=50.0)),
    # acquisition + preRun + 300 ms of tracking on 4 channels out of 5 satellites
    "trk_small": dict(seed=3, ms=300, n_ms=303, settings=dict(numberOfChannels=4), sats=_sats([
        (4, 2100.0, 777), (8, -4300.0, 30011), (15, 5600.0, 15000, dict(cn0=50.0)),
        (21, -900.0, 38100, dict(cn0=44.0)), (30, 3333.0, 9000)], cn0=47.0)),
    # non-zero skipNumberOfBytes and a weak satellite near the detection threshold
    "trk_skip": dict(seed=4, ms=120, n_ms=125, settings=dict(numberOfChannels=3, skipNumberOfBytes=5000),
                     sats=_sats([(6, -1234.0, 100), (13, 4321.0, 22222),
                                 (27, 600.0, 31000, dict(cn0=39.0))], cn0=48.0)),
}


def build_recording(case):
    spec = synth.RecordingSpec(case["sats"], seed=case["seed"])
    data = synth.generate_cpu(spec, case["n_ms"] * N)
    return spec, data


class SettingsLike(object):
    """Minimal settings bag for the oracle (same attribute names as the reference)."""

    def __init__(self, **kw):
        from softgnss_python_b200.settings import Settings
        base = Settings(**kw)
        self.__dict__.update(base.__dict__)
        self.acqSatelliteList = range(1, 33)


def case_settings(case):
    from softgnss_python_b200.settings import Settings
    s = Settings(**case.get("settings", {}))
    s.msToProcess = float(case["ms"])
    return s
