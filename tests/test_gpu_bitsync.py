"""sgx_find_preambles on the B200 against the reference's golden output, the oracle, and -- on a real tracking
run of the LNAV scenario -- the oracle applied to the same tracking output."""
import os

import numpy as np
import pytest

from oracle import gnss_oracle as orc
from tests.cases import BITSYNC_EARLY, BITSYNC_MS, build_bitsync_case, build_bitsync_channel

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden", "bitsync.npz")


def test_bitsync_matches_reference_golden():
    from softgnss_python_b200 import postnav
    g = np.load(GOLD, allow_pickle=False)
    ips = np.stack(build_bitsync_case() + [build_bitsync_channel(BITSYNC_EARLY, 600)])
    first, bits, valid = postnav.find_preambles_batch(ips)
    assert np.array_equal(first[:-1], g["first"])
    assert first[-1] == 6020                                   # candidate at 20 ms passed over (see header)
    gb = np.unpackbits(g["nav_bits"], axis=1)[:, :1501]
    assert np.array_equal(valid[:-1], g["bits_valid"])
    for ch in range(len(gb)):
        if g["bits_valid"][ch]:
            assert np.array_equal(bits[ch], gb[ch]), ch


def test_bitsync_device_tensor_and_strided_rows():
    import torch
    from softgnss_python_b200 import postnav
    ips = np.stack(build_bitsync_case())
    pad = np.zeros((ips.shape[0], BITSYNC_MS + 13))
    pad[:, :BITSYNC_MS] = ips
    t = torch.from_numpy(pad).cuda()
    first, bits, valid = postnav.find_preambles_batch(t[:, :BITSYNC_MS])
    f0, b0, v0 = postnav.find_preambles_batch(ips)
    assert np.array_equal(first, f0) and np.array_equal(bits, b0) and np.array_equal(valid, v0)


@pytest.mark.parametrize("ms", [160, 6199, 12345, 36000])
def test_bitsync_ragged_lengths_match_oracle(ms):
    from softgnss_python_b200 import postnav
    ips = [x[:ms] for x in build_bitsync_case()]
    first, bits, valid = postnav.find_preambles_batch(np.stack(ips))
    of, _ = orc.find_preambles(ips)
    assert np.array_equal(first, of)
    for ch in range(len(ips)):
        ok = of[ch] != 0 and of[ch] + 30000 <= ms
        assert bool(valid[ch]) == bool(ok)
        if ok:
            assert np.array_equal(bits[ch], orc.nav_bits(ips[ch], int(of[ch])))


def test_bitsync_mirror_of_findPreambles(capsys):
    from softgnss_python_b200 import postnav
    from softgnss_python_b200.settings import Settings
    ips = build_bitsync_case()
    zero = np.zeros(BITSYNC_MS)
    dtype = [('status', 'U1')] + [(f, 'object') for f in orc.TRACK_FIELDS] + [('PRN', 'int64')]
    rec = [('T',) + tuple(ips[c] if f == "I_P" else zero for f in orc.TRACK_FIELDS) + (c + 1,) for c in range(len(ips))]
    res = np.rec.fromrecords(rec, dtype=dtype)
    g = np.load(GOLD, allow_pickle=False)
    first, active = postnav.findPreambles(res, Settings(numberOfChannels=len(ips)))
    assert np.array_equal(first, g["first"]) and np.array_equal(active, g["active"])
    assert capsys.readouterr().out.count("Could not find valid preambles") == 2


def test_bitsync_on_tracking_output_of_the_lnav_scenario():
    """Track 8 channels x 37 000 ms of the LNAV scenario on the device, search the preambles in the device-resident
    result, and compare with the oracle applied to the same I_P series."""
    import torch
    from softgnss_python_b200 import _native, navsynth, postnav, synth
    from softgnss_python_b200.settings import Settings, to_pod
    ms, n_code = 37000, 38192
    spec, truth = navsynth.build_scenario(seed=2)
    L = _native.lib()
    n = (ms + 2) * n_code
    stride = (n + 15) // 16 * 16
    dev = torch.empty((1, stride), dtype=torch.int8, device="cuda")
    sp, bits = _native.make_synth_specs([spec])
    stream = torch.cuda.current_stream().cuda_stream
    L.synth(dev, stride, n, 0, sp, bits, synth.cos_lut(), _native.ca_chips_int8(), stream)
    s = Settings(msToProcess=float(ms))
    prn = [x.prn for x in spec.sats]
    freq = [spec.true_carr_freq(i) - 20.0 for i in range(len(prn))]
    cph = [(x.code_phase + 1) % n_code for x in spec.sats]
    out = torch.zeros((1, 8, 13, ms), dtype=torch.float64, device="cuda")
    rc, done = L.track(dev, stride, [n], _native.make_channels(prn, freq, cph), to_pod(s), _native.ca_chips_int8(),
                       out, stream)
    assert rc == 0 and int(done.min()) == ms
    ip_dev = out[0, :, 3, :]                                    # [8, ms] view, row stride 13 * ms
    first, nb, valid = postnav.find_preambles_batch(ip_dev)
    ip = ip_dev.cpu().numpy()
    of, oa = orc.find_preambles(list(ip))
    assert np.array_equal(first, of)
    assert len(oa) == 8 and (first > 0).all()
    # the subframe boundaries arrive staggered by the geometric ranges of the scenario
    b = np.array(truth["boundary_sample"]) / n_code
    assert np.all(np.abs(((first - b) + 3000) % 6000 - 3000) <= 1.5)
    for ch in range(8):
        assert valid[ch] == 1
        assert np.array_equal(nb[ch], orc.nav_bits(ip[ch], int(of[ch])))


def test_pseudoranges_bit_identical_to_reference():
    import torch
    from softgnss_python_b200 import postnav
    from softgnss_python_b200.settings import Settings
    from tests.cases import build_pseudo_case
    abs_sample, ms_index, act = build_pseudo_case()
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "pseudo.npz"), allow_pickle=False)
    n_ch, ms = abs_sample.shape
    s = Settings(numberOfChannels=n_ch)
    # batch of 3 "recordings": the case, the case shifted by whole code periods (same pseudoranges), a copy
    trk = np.zeros((3, n_ch, 13, ms))
    trk[0, :, 0] = abs_sample
    trk[1, :, 0] = abs_sample + 5 * 38192
    trk[2, :, 0] = abs_sample
    idx = np.broadcast_to(ms_index, (3,) + ms_index.shape).copy()
    a = np.broadcast_to(act, (3,) + act.shape).copy()
    pr = postnav.pseudoranges_batch(trk, idx, a, s)
    assert np.array_equal(pr[0], g["pseudoranges"], equal_nan=True)
    assert np.array_equal(pr[2], pr[0], equal_nan=True)
    assert np.allclose(pr[1], pr[0], rtol=0, atol=1e-6, equal_nan=True)
    prd = postnav.pseudoranges_batch(torch.from_numpy(trk).cuda(), idx, a, s)          # device-resident input
    assert np.array_equal(prd, pr, equal_nan=True)
    # function-style mirror of the reference method, one epoch
    dtype = [('status', 'U1')] + [(f, 'object') for f in orc.TRACK_FIELDS] + [('PRN', 'int64')]
    zero = np.zeros(ms)
    rec = [('T',) + tuple(abs_sample[c] if f == "absoluteSample" else zero for f in orc.TRACK_FIELDS) + (c + 1,)
           for c in range(n_ch)]
    res = np.rec.fromrecords(rec, dtype=dtype)
    one = postnav.calculatePseudoranges(res, ms_index[1].astype(float), act[1].nonzero()[0], s)
    assert np.array_equal(one, g["pseudoranges"][1])
