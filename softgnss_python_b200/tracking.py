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
FILE_CHUNK_SAMPLES = 0          # samples per staging buffer of the file ingest (0: the library's default, 1024 code periods)
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


SAMPLE_BYTES = {"int8": 1, "b": 1, "schar": 1, "int16": 2, "short": 2}


def _sample_bytes(settings):
    """``settings.dataType`` (initialize.py:102) -> bytes per sample; anything but int8 / int16 is rejected."""
    dt = str(getattr(settings, "dataType", "int8"))
    if dt not in SAMPLE_BYTES:
        raise _native.NativeError(-2, "dataType %r is not supported (int8, int16)" % dt)
    return SAMPLE_BYTES[dt]


def _file_path(fid):
    """Path of a recording given as a path or as an open file object (the reference's form), else None."""
    import os
    if isinstance(fid, (str, bytes, os.PathLike)):
        return os.fspath(fid)
    name = getattr(fid, "name", None)
    if isinstance(name, (str, bytes)) and os.path.exists(name):
        return name
    return None


def _load_recording(fid, settings=None):
    """In-memory forms of a recording: an int8 numpy array, a CUDA tensor (passed through), or a file-like object
    without a path (read once).  Files with a path never come here: they are streamed (``sgx_track_file``)."""
    if hasattr(fid, "data_ptr"):            # torch tensor
        return fid
    if isinstance(fid, np.ndarray):
        assert fid.dtype == np.int8 and fid.ndim == 1
        return np.ascontiguousarray(fid)
    fid.seek(0)
    raw = fid.read()
    if settings is not None and _sample_bytes(settings) == 2:
        x = np.frombuffer(raw, dtype="<i2")
        if x.size and (x.min() < -128 or x.max() > 127):
            raise _native.NativeError(-2, "int16 sample outside the int8 range of the correlators")
        return x.astype(np.int8)
    return np.frombuffer(raw, dtype=np.int8)


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
    path = _file_path(fid)
    if path is not None:
        # on-disk recording: pread -> two pinned staging buffers -> HBM, overlapped with tracking; only the window the
        # channels can touch is read (tracking.py:107, :154; initialize.py:102, :466-481)
        L = _native.lib()
        pod = to_pod(settings)
        c = int(settings.numberOfChannels)
        chans = _native.make_channels([int(channel.PRN[i]) for i in range(c)],
                                      [float(channel.acquiredFreq[i]) for i in range(c)],
                                      [float(channel.codePhase[i]) for i in range(c)])
        out = np.zeros((1, c, len(FIELDS), pod.msToProcess), dtype=np.float64)
        rc, done, _ = L.track_file(path, _sample_bytes(settings), chans, pod, _native.ca_chips_int8(), out,
                                   chunk_samples=FILE_CHUNK_SAMPLES)
    else:
        data = _load_recording(fid, settings)
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
