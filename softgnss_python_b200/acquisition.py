"""Acquisition stage: same call surface as the reference, computed by the sm_100a pipeline.

Reference surface (``acquisition.py:6-336``):
    ``a = AcquisitionResult(settings); a.acquire(longSignal); a.carrFreq/.codePhase/.peakMetric;
    a.preRun(); a.showChannelStatus(); a.channels``
and the Matlab-style functions it keeps as comments (``acquisition.py:34``, ``:266``):
    ``acqResults = acquisition(longSignal, settings)``, ``channel = preRun(acqResults, settings)``.

``acqResults`` is the reference's recarray of 32 rows, fields ``carrFreq, codePhase, peakMetric``
(float64; ``carrFreq == 0`` <=> not detected, ``acquisition.py:201-203``).
"""
import numpy as np

from . import _native
from .settings import fine_code_index, make_ca_table, to_pod
from .tracking import Result


_TABLES = {}
ACQ_SPECTRA_BYTES = 24 << 30     # device memory one call may use for the spectra (the search work buffers come on top)


def _tables(settings, fine_ms):
    """Host-built tables (A3 code table, A10 chip index, C/A chips), cached per signal definition."""
    key = (float(settings.samplingFreq), float(settings.codeFreqBasis), int(settings.codeLength), int(fine_ms))
    if key not in _TABLES:
        _TABLES.clear()
        _TABLES[key] = (np.ascontiguousarray(make_ca_table(settings).astype(np.int8)),
                        np.ascontiguousarray(fine_code_index(settings, fine_ms)),
                        _native.ca_chips_int8())
    return _TABLES[key]


def acquire_batch(signals, settings, prn_first=0, prn_count=None, stream=0, diagnostics=False):
    """``signals``: int8 [R, n_samples] (numpy or CUDA tensor).  Searches PRN indices
    [prn_first, prn_first+prn_count) for every recording; returns dict of float64 [R, prn_count]."""
    L = _native.lib()
    L.require_device()
    pod = to_pod(settings)
    nsat = min(32, len(settings.acqSatelliteList))          # acquisition.py:92 -- only the length matters
    if prn_count is None:
        prn_count = nsat - prn_first
    r, ns = int(signals.shape[0]), int(signals.shape[1])
    stride = signals.stride(0) if hasattr(signals, "data_ptr") else signals.strides[0]
    table, fidx, chips = _tables(settings, pod.fineMs)   # named: the buffers must outlive the ctypes call
    carr = np.zeros((r, prn_count))
    cph = np.zeros((r, prn_count))
    met = np.zeros((r, prn_count))
    fbin = np.zeros((r, prn_count), dtype=np.int32)
    fine = np.zeros((r, prn_count), dtype=np.int32)
    import ctypes
    # one sgx_acquire call holds the wiped-off spectra of all its recordings (blocks x bins x n complex64 each):
    # large batches (BASELINE config 3: 141 bins x 10 ms) are walked in slices of recordings
    per_rec = int(pod.acqNonCoherentBlocks) * int(pod.numFrqBins)
    spec_bytes = per_rec * int(pod.samplesPerCode) * int(pod.acqCoherentMs) * 8
    rmax = max(1, min(32768 // per_rec, ACQ_SPECTRA_BYTES // spec_bytes, 32768 // prn_count))
    for r0 in range(0, r, rmax):
        r1 = min(r, r0 + rmax)
        rc = L.dll.sgx_acquire(_native._ptr(signals[r0:r1]), int(stride), ns, r1 - r0, ctypes.byref(pod),
                               _native._ptr(table), _native._ptr(chips), _native._ptr(fidx), int(prn_first),
                               int(prn_count), _native._ptr(carr[r0:r1]), _native._ptr(cph[r0:r1]),
                               _native._ptr(met[r0:r1]), _native._ptr(fbin[r0:r1]), _native._ptr(fine[r0:r1]),
                               ctypes.c_void_p(stream))
        L.check(rc)
    out = dict(carrFreq=carr, codePhase=cph, peakMetric=met)
    if diagnostics:
        out.update(frqBin=fbin, finePeakIndex=fine)
    return out


def acquisition(longSignal, settings):
    """``acqResults = acquisition(longSignal, settings)`` (reference acquisition.py:27-204)."""
    x = longSignal
    if not hasattr(x, "data_ptr"):
        x = np.ascontiguousarray(np.asarray(longSignal, dtype=np.int8))
    res = acquire_batch(x.reshape(1, -1), settings)
    carr = np.zeros(32)
    cph = np.zeros(32)
    met = np.zeros(32)
    k = res["carrFreq"].shape[1]
    carr[:k], cph[:k], met[:k] = res["carrFreq"][0], res["codePhase"][0], res["peakMetric"][0]
    return np.rec.fromarrays([carr, cph, met], names='carrFreq,codePhase,peakMetric')


def preRun(acqResults, settings):
    """``channel = preRun(acqResults, settings)`` (reference acquisition.py:259-306): the strongest
    ``numberOfChannels`` detected PRNs in order of peakMetric; the rest idle (PRN 0, status '-')."""
    nch = int(settings.numberOfChannels)
    prn = np.zeros(nch, dtype='int64')
    freq = np.zeros(nch)
    cph = np.zeros(nch)
    status = ['-'] * nch
    order = sorted(range(len(acqResults.peakMetric)), key=lambda i: acqResults.peakMetric[i], reverse=True)
    for slot in range(min(nch, int(np.sum(acqResults.carrFreq > 0)))):
        i = order[slot]
        prn[slot] = i + 1
        freq[slot] = acqResults.carrFreq[i]
        cph[slot] = acqResults.codePhase[i]
        status[slot] = 'T'
    return np.rec.fromarrays([prn, freq, cph, status], names='PRN,acquiredFreq,codePhase,status')


def showChannelStatus(channel, settings):
    """Channel table (reference acquisition.py:308-336)."""
    bar = '*=========*=====*===============*===========*=============*========*'
    print('\n' + bar)
    print('| Channel | PRN |   Frequency   |  Doppler  | Code Offset | Status |')
    print(bar)
    for n in range(int(settings.numberOfChannels)):
        ch = channel[n]
        if ch.status != '-':
            print('|      %2d | %3d |  %2.5e |   %5.0f   |    %6d   |     %1s  |' % (
                n, ch.PRN, ch.acquiredFreq, ch.acquiredFreq - settings.IF, ch.codePhase, ch.status))
        else:
            print('|      %2d | --- |  ------------ |   -----   |    ------   |   Off  |' % n)
    print(bar + '\n')


class AcquisitionResult(Result):
    """Drop-in for the reference class of the same name (``acquisition.py:6-25``)."""

    def __init__(self, settings):
        Result.__init__(self, settings)

    @property
    def peakMetric(self):
        assert isinstance(self._results, np.recarray)
        return self._results.peakMetric

    @property
    def carrFreq(self):
        assert isinstance(self._results, np.recarray)
        return self._results.carrFreq

    @property
    def codePhase(self):
        assert isinstance(self._results, np.recarray)
        return self._results.codePhase

    def acquire(self, longSignal):
        self._results = acquisition(longSignal, self._settings)
        return

    def preRun(self):
        assert isinstance(self._results, np.recarray)
        self._channels = preRun(self._results, self._settings)
        return

    def showChannelStatus(self):
        assert isinstance(self._channels, np.recarray)
        showChannelStatus(self._channels, self._settings)
