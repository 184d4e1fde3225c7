"""GPU tests of the big-block overlap-save route (upols.cu: olsb_filter): ONE partition, 2^18..2^22-point blocks through
the two-pass engine with the fused middle pass.  Checked against the CPU oracle (scipy fftconvolve / the reference's
exact-N masks) and against the 4096-frame partitioned route it replaces; `ars_olsb_count` asserts which route ran."""
import numpy as np
import pytest

import ars_oracle as orc

pytestmark = pytest.mark.gpu
TOL = 1e-5
RATE = 48000


def rel_err(got, ref):
    got = np.asarray(got, np.float64)
    ref = np.asarray(ref, np.float64)
    assert got.shape == ref.shape, (got.shape, ref.shape)
    return float(np.max(np.abs(got - ref)) / max(1.0, float(np.max(np.abs(ref)))))


def snr_db(got, ref):
    noise = np.sum((np.asarray(got, np.float64) - ref) ** 2)
    return np.inf if noise == 0 else 10 * np.log10(np.sum(np.asarray(ref, np.float64) ** 2) / noise)


@pytest.fixture()
def opts():
    from ars_b200 import _capi
    yield _capi
    for k, v in (("olsb", 1), ("olsb_logf", 0), ("olsb_stripe", 0), ("air_fold", 1)):
        _capi.set_option(k, v)


def _count(capi):
    return int(capi.load_library().ars_olsb_count())


def _dense_ir(seconds, seed):
    L = int(seconds * RATE)
    g = np.random.default_rng(seed)
    ir = g.standard_normal((L, 2)) * np.exp(-6.9 * np.arange(L) / (0.6 * L))[:, None]
    return (ir / np.max(np.abs(ir)) / 8.0).astype(np.float32)


@pytest.mark.parametrize("cin", [1, 2, 3, 6])
def test_split_mask_free_vs_oracle_and_partitioned_route(rs, opts, cin):
    """convolve_audio_split_3d with air off and EQ flat (rs.py:338-408), 12 s clips of 1 / 2 / 3 / 6 channels."""
    g = np.random.default_rng(40 + cin)
    x = (0.3 * g.standard_normal((12 * RATE, cin) if cin > 1 else 12 * RATE)).astype(np.float32)
    np.random.seed(9)
    early, late = rs.generate_impulse_response_split_3d(RATE, 1.5, 30, 0.06, "Holz", 0.5, 0.08, 0.5)
    want = orc.convolve_split(x, early, late, 0.7, 0.9, 0.6, 1.0, 1.0, RATE, 0.5, 0.0)
    c0 = _count(opts)
    got = rs.convolve_audio_split_3d(x, early, late, 0.7, 0.9, 0.6, 1.0, 1.0, RATE, 0.5, 0.0)
    assert _count(opts) == c0 + 1
    opts.set_option("olsb", 0)
    old = rs.convolve_audio_split_3d(x, early, late, 0.7, 0.9, 0.6, 1.0, 1.0, RATE, 0.5, 0.0)
    assert _count(opts) == c0 + 1
    assert rel_err(got, want) <= TOL and snr_db(got, want) >= 100.0, (rel_err(got, want), snr_db(got, want))
    assert rel_err(got, old) <= 2e-6, rel_err(got, old)


@pytest.mark.parametrize("logf", [0, 18, 19, 20, 21, 22])
def test_external_stereo_ir_every_block_length(rs, opts, logf):
    """convolve_audio_external_ir (rs.py:410-462): the mirror form of the middle pass (per-channel spectra from the
    bin pair k, F - k) at every block length / first-pass shape (2^6 .. 2^10 x 2^12)."""
    g = np.random.default_rng(7)
    x = (0.2 * g.standard_normal((50 * RATE, 2))).astype(np.float32)
    ir = _dense_ir(1.5, 8)
    want = orc.convolve_external(x, ir, 0.5, 1.0, 1.0, RATE, 0.5)
    opts.set_option("olsb_logf", logf)
    c0 = _count(opts)
    got = rs.convolve_audio_external_ir(x, ir, 0.5, 1.0, 1.0, RATE, 0.5)
    assert _count(opts) == c0 + 1
    assert rel_err(got, want) <= TOL and snr_db(got, want) >= 100.0, (logf, rel_err(got, want), snr_db(got, want))


@pytest.mark.parametrize("stripe", [1, 3, 64])
def test_stripes_do_not_change_the_result(rs, opts, stripe):
    g = np.random.default_rng(5)
    x = (0.2 * g.standard_normal((40 * RATE, 6))).astype(np.float32)
    ir = _dense_ir(0.4, 3)
    ref = rs.convolve_audio_external_ir(x, ir, 0.4, 1.0, 1.0, RATE, 0.5)
    opts.set_option("olsb_stripe", stripe)
    got = rs.convolve_audio_external_ir(x, ir, 0.4, 1.0, 1.0, RATE, 0.5)
    assert np.array_equal(got, ref)


def test_folded_air_route_through_big_blocks(rs, opts):
    """Air ramp folded into the IR (circular form: taps before time zero, N-periodic signal) on 30 s clips."""
    g = np.random.default_rng(21)
    np.random.seed(5)
    early, late = rs.generate_impulse_response_split_3d(RATE, 1.5, 30, 0.06, "Holz", 0.5, 0.08, 0.5)
    for cin, air in ((2, 0.1), (6, 0.25), (1, 0.05)):
        x = (0.3 * g.standard_normal((30 * RATE, cin) if cin > 1 else 30 * RATE)).astype(np.float32)
        want = orc.convolve_split(x, early, late, 0.7, 0.9, 0.6, 1.0, 1.0, RATE, 0.5, air)
        c0 = _count(opts)
        got = rs.convolve_audio_split_3d(x, early, late, 0.7, 0.9, 0.6, 1.0, 1.0, RATE, 0.5, air)
        assert _count(opts) == c0 + 1, "the folded-air stage must take the big-block route at this size"
        opts.set_option("olsb", 0)
        old = rs.convolve_audio_split_3d(x, early, late, 0.7, 0.9, 0.6, 1.0, 1.0, RATE, 0.5, air)
        opts.set_option("olsb", 1)
        assert rel_err(got, want) <= TOL and snr_db(got, want) >= 100.0, (cin, air, rel_err(got, want))
        assert rel_err(got, old) <= 2e-6, rel_err(got, old)


def test_air_kernel_reaching_before_time_zero_wraps(rs, opts):
    """A late part that starts right at the beginning of the IR: the folded taps start before time zero (adv > 0) and the
    pre-ring wraps around the period N as in the reference's N-point filter."""
    n = 20 * RATE
    t = np.arange(n)
    late = np.zeros(600, np.float32)
    late[20:600] = (np.random.default_rng(2).standard_normal(580) * np.exp(-np.arange(580) / 80.0)).astype(np.float32)
    late *= 0.7 / np.max(np.abs(late))
    early = np.zeros(600, np.float32)
    early[[3, 40, 77]] = [0.9, -0.4, 0.2]
    sigs = {"noise": (np.random.default_rng(3).standard_normal((n, 2)) * 0.3).astype(np.float32),
            "tones": np.stack([0.5 * np.sin(2 * np.pi * 2000.0 * t / RATE), 0.5 * np.cos(np.pi * t)], 1).astype(np.float32)}
    for name, x in sigs.items():
        want = orc.convolve_split(x, early, late, 0.8, 1.0, 0.7, 1.0, 1.0, RATE, 0.5, 0.12)
        c0 = _count(opts)
        got = rs.convolve_audio_split_3d(x, early, late, 0.8, 1.0, 0.7, 1.0, 1.0, RATE, 0.5, 0.12)
        assert _count(opts) == c0 + 1
        assert rel_err(got, want) <= TOL, (name, rel_err(got, want))
