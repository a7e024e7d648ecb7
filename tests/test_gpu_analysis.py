"""Analysis tasks that produce numbers (reference: dedalus/analysis/analysis_set.py): TrackMode, PowerSpectrum, VolumeAverage under
an AnalysisSet, through the drop-in package.  PowerSpectrum against a numpy restatement of _compute_spectrum (:778-816)."""
import numpy as np
import pytest

from devutil import dev_physics, oracle_physics, set_state

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", autouse=True)
def _native_lib_loaded():
    from conftest import native_lib_expected
    native_lib_expected()
    yield


def spectrum_numpy(g, comps, norm=1.0):
    """analysis_set.py:778-816 with numpy on the oracle's grid."""
    kmag = np.sqrt(g.k2())
    power = np.zeros_like(kmag)
    for c in comps:
        power += np.abs(c) ** 2
    power1d = norm * power * (2. * np.pi * kmag if g.ndim == 2 else 4. * np.pi * kmag ** 2)
    kmax = np.sqrt(np.sum(g.kny ** 2))
    n = int(np.min([np.min(g.shape) / 2., 100]))
    kbottom = np.linspace(0, kmax, n, endpoint=False)
    ktop = kbottom + kbottom[1]
    spec, cnt = np.zeros(n), np.zeros(n)
    for i in range(n):
        mask = (kmag >= kbottom[i]) & (kmag < ktop[i]) & (power1d != 0)
        spec[i] = power1d[mask].sum()
        cnt[i] = mask.sum()
    cnt[cnt == 0] = 1.
    return kbottom + kbottom[1] / 2., spec / cnt


@pytest.mark.parametrize("physics,shape,cfg", [("IncompressibleHydro", (64, 48), 1), ("IncompressibleMHD", (16, 32, 24), 5)])
def test_tasks_under_an_analysis_set(tmp_path, monkeypatch, physics, shape, cfg):
    import dedalus_oracle as orc
    from dedalus.mods import AnalysisSet, TrackMode, PowerSpectrum, VolumeAverage, VolumeAverageSet, Snapshot, RK2mid
    monkeypatch.chdir(tmp_path)
    Po = oracle_physics(physics, shape, None, dict(nu=1e-3))
    do = orc.synthetic_ic(Po, cfg)
    P = dev_physics(physics, shape, None, dict(nu=1e-3))
    data = P.create_fields(0.)
    set_state(data, do.kvector())
    ti = RK2mid(P)
    vs = VolumeAverageSet(data)
    vs.add("ekin", "%20.12e")
    an = AnalysisSet(data, ti)
    index = (1, 2) if len(shape) == 2 else (2, 1, 3)
    an.add(TrackMode(1, fieldlist=["u"], indexlist=[index]))
    an.add(PowerSpectrum(2, fieldlist=["u"], plot=False))
    an.add(VolumeAverage(1, vs))
    an.add(Snapshot(1))
    an.run()
    for _ in range(2):
        ti.advance(data, 1e-3)
        an.run()
    an.cleanup()
    # tracked amplitudes: the first line is the initial state's entry
    lines = [l for l in open("ux_mode_amplitudes.dat").read().splitlines() if not l.startswith("#")]
    assert len(lines) == 3
    t0, amp0 = lines[0].split("\t")
    assert float(t0) == 0.0 and abs(complex(amp0) - do.kvector()[0][index]) < 1e-15
    # spectra: iteration 0 and 2; the first against the numpy restatement
    rows = open("u_power_spectra.dat").read().splitlines()
    assert rows[0].startswith("# Dedalus Power Spectrum") and rows[1].startswith("time\t") and len(rows) == 4
    k_ref, s_ref = spectrum_numpy(Po.g, [c["kspace"] for _, c in do["u"]])
    assert np.allclose(np.array(rows[1].split("\t")[1:], dtype=float), k_ref, rtol=1e-14)
    got = np.array(rows[2].split("\t")[1:], dtype=float)
    assert np.allclose(got, s_ref, rtol=1e-12, atol=1e-18)
    ts = [l for l in open("time_series.dat").read().splitlines() if not l.startswith("#")]
    assert len(ts) == 3 and abs(float(ts[0].split("\t")[1]) - orc.invariants(do)["ekin"]) < 1e-11
