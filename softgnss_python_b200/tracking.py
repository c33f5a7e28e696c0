"""Tracking stage: same call surface as the reference, computed by the sm_100a kernel.

Reference surface (``tracking.py:6-295``):
    ``TrackingResult(acqResult).track(fid)`` then ``.results`` / ``.channels`` / ``.settings``;
the Matlab-style function the reference keeps as a comment (``tracking.py:18``):
    ``trackResults, channel = tracking(fid, channel, settings)``.

Both are provided.  ``fid`` may be an open binary file (the reference's only form), a path, an int8
numpy array holding the file's bytes, or an int8 CUDA tensor already resident in HBM.  The result is
the reference's recarray (``tracking.py:285-294``): one record per channel whose PRN != 0, fields
``status`` ('S1'), thirteen float64[msToProcess] series as object fields, ``PRN`` int64.
On a short recording the reference prints a message and returns None (``tracking.py:159-163``);
so does this.
"""
import numpy as np

from . import _native
from .settings import to_pod

FIELDS = _native.TRACK_FIELDS
RESULT_DTYPE = [('status', 'S1')] + [(f, 'object') for f in FIELDS] + [('PRN', 'int64')]


class Result(object):
    """Base of the stage objects (reference ``initialize.py:20-46``)."""

    def __init__(self, settings):
        self._settings = settings
        self._results = None
        self._channels = None

    @property
    def settings(self):
        return self._settings

    @property
    def channels(self):
        assert isinstance(self._channels, np.recarray)
        return self._channels

    @property
    def results(self):
        assert isinstance(self._results, np.recarray)
        return self._results

    @results.setter
    def results(self, records):
        assert isinstance(records, np.recarray)
        self._results = records

    def plot(self):
        pass


def _load_recording(fid):
    """The bytes the reference would reach through fid.seek/np.fromfile, as one int8 array
    (or a CUDA tensor passed through)."""
    if hasattr(fid, "data_ptr"):            # torch tensor
        return fid
    if isinstance(fid, np.ndarray):
        assert fid.dtype == np.int8 and fid.ndim == 1
        return np.ascontiguousarray(fid)
    if isinstance(fid, str):
        return np.fromfile(fid, dtype=np.int8)
    name = getattr(fid, "name", None)
    if isinstance(name, str):
        return np.fromfile(name, dtype=np.int8)
    fid.seek(0)
    return np.frombuffer(fid.read(), dtype=np.int8)


def track_batch(recordings, rec_len, channel_sets, settings, out=None, stream=0):
    """Batched entry: ``recordings`` int8 [R, stride] (numpy or CUDA tensor), ``channel_sets`` a list of
    R channel recarrays (``preRun`` output).  Returns (rc, out[R, C, 13, ms], ms_done[R, C])."""
    L = _native.lib()
    pod = to_pod(settings)
    r = len(channel_sets)
    c = int(settings.numberOfChannels)
    prn, freq, cph = [], [], []
    for ch in channel_sets:
        for i in range(c):
            prn.append(int(ch.PRN[i])); freq.append(float(ch.acquiredFreq[i])); cph.append(float(ch.codePhase[i]))
    chans = _native.make_channels(prn, freq, cph)
    ms = pod.msToProcess
    if out is None:
        out = np.empty((r, c, len(FIELDS), ms), dtype=np.float64)
    stride = recordings.stride(0) if hasattr(recordings, "data_ptr") else recordings.strides[0]
    rc, done = L.track(recordings, stride, rec_len, chans, pod, _native.ca_chips_int8(), out, stream)
    return rc, out, done


def tracking(fid, channel, settings):
    """``[trackResults, channel] = tracking(fid, channel, settings)`` (reference tracking.py:13-295)."""
    data = _load_recording(fid)
    n = int(data.numel() if hasattr(data, "data_ptr") else data.size)
    rec = data.reshape(1, n) if not hasattr(data, "data_ptr") else data.view(1, n)
    rc, out, done = track_batch(rec, [n], [channel], settings)
    if rc == _native.SGX_ERR_SHORT:
        print('Not able to read the specified number of samples for tracking, exiting!')
        if hasattr(fid, "close"):
            fid.close()
        return None, channel
    _native.lib().check(rc)
    recs = []
    for i in range(int(settings.numberOfChannels)):
        if channel.PRN[i] == 0:           # tracking.py:99, :280-283 -- idle channels produce no record
            continue
        series = tuple(np.array(out[0, i, f]) for f in range(len(FIELDS)))
        recs.append((channel.status[i],) + series + (int(channel.PRN[i]),))
    if recs:
        res = np.rec.fromrecords(recs, dtype=RESULT_DTYPE)
    else:
        res = np.recarray((0,), dtype=RESULT_DTYPE)
    return res, channel


class TrackingResult(Result):
    """Drop-in for the reference class of the same name (``tracking.py:6-13``)."""

    def __init__(self, acqResult):
        self._results = None
        self._channels = acqResult.channels
        self._settings = acqResult.settings

    def track(self, fid):
        res, _ = tracking(fid, self._channels, self._settings)
        if res is None:
            return None
        self._results = res
        return
