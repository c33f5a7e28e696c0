"""Developer check under the CPU emulator: other sampling rates and the extension settings vs the oracle."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from softgnss_python_b200 import _native, synth
_native.LIB_PATH = os.environ.get("SGX_EMUL_LIB", os.path.join(ROOT, "tools", "cpu_emul", "libsoftgnss_emul.so"))
_native._LIB = _native.Lib(_native.LIB_PATH)
from softgnss_python_b200.acquisition import acquire_batch
from softgnss_python_b200.settings import Settings
from softgnss_python_b200.tracking import track_batch
from oracle import gnss_oracle as orc

def run(fs, f_if, nprn=6, coh=1, blocks=2, step=500.0, ms=20, cn0=52.0):
    s = Settings(samplingFreq=fs, IF=f_if, acqCoherentMs=coh, acqNonCoherentBlocks=blocks, acqDopplerStep=step,
                 numberOfChannels=2, msToProcess=float(ms))
    s.acqSatelliteList = range(1, nprn + 1)
    n = s.samplesPerCode
    sats = [synth.SatSpec(2, 1250.0, n // 3, cn0=cn0, bit_offset_ms=3), synth.SatSpec(5, -2750.0, n - 7, cn0=cn0, bit_offset_ms=11)]
    spec = synth.RecordingSpec(sats, fs=fs, f_if=f_if, seed=11)
    nlong = max(11, coh * blocks) * n
    data = synth.generate_cpu(spec, max(nlong, (ms + 3) * n))
    t = time.time()
    ref = orc.acquire(data[:nlong], s, coherent_ms=coh, noncoh_blocks=blocks, doppler_step=step, clamp_window=True)
    t1 = time.time()
    got = acquire_batch(data[:nlong].reshape(1, -1), s)
    print("fs=%.4g N=%d coh=%d blocks=%d step=%g: oracle %.1fs emul %.1fs" % (fs, n, coh, blocks, step, t1 - t, time.time() - t1))
    print("   detected ref", np.nonzero(ref["carrFreq"][:nprn])[0] + 1, "got", np.nonzero(got["carrFreq"][0])[0] + 1)
    print("   codePhase ref", ref["codePhase"][:nprn], "got", got["codePhase"][0])
    print("   carrFreq diff", np.abs(ref["carrFreq"][:nprn] - got["carrFreq"][0]).max(), " metric rel",
          np.abs(got["peakMetric"][0] / ref["peakMetric"][:nprn] - 1).max())
    if ms > 0 and coh == 1:
        ch = orc.pre_run(ref, s)
        chr_ = np.rec.fromarrays([ch["PRN"], ch["acquiredFreq"], ch["codePhase"], ch["status"]], names="PRN,acquiredFreq,codePhase,status")
        recs = orc.track(data, ch, s)
        rc, out, done = track_batch(data.reshape(1, -1), [data.size], [chr_], s)
        print("   track rc", rc, "done", done.tolist())
        act = [i for i in range(2) if ch["PRN"][i] != 0]
        for fi, f in enumerate(_native.TRACK_FIELDS):
            refv = np.stack([r[2][f] for r in recs]); gotv = out[0, act, fi, :]
            if f in ("absoluteSample", "I_P", "Q_P", "carrFreq", "codeFreq"):
                print("     %-14s max|d|=%.3e (scale %.3g)" % (f, np.abs(gotv - refv).max(), np.abs(refv).max()))

if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    if which in ("all", "fs16"): run(16.3676e6, 4.1304e6)
    if which in ("all", "fs4"): run(4.092e6, 1.023e6, ms=10)
    if which in ("all", "fs64"): run(64e6, 16e6, nprn=5, ms=6)
    if which in ("all", "ext"): run(38.192e6, 9.548e6, nprn=5, coh=2, blocks=3, step=250.0, ms=0, cn0=47.0)
