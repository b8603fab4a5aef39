"""CPU tier: the KRAKEN .mod reader / writer (bogp_b200/kraken_mod.py).  The binary layout is restated from memory of the
Acoustics Toolbox's reader (parity UNPINNED: no KRAKEN, no toolbox, no real file here), so what is tested is the round
trip through this module's own writer, the record arithmetic of broadband files, and that the ModeSet assembled from
files equals the one built from the same tables directly."""
import struct

import numpy as np
import pytest

from bogp_b200 import kraken_mod
from bogp_b200.modes import INV_TONALS_3, REC_Z_21, modeset_from_tables, pekeris_modes, replica_field


def _tables(freqs, z, complex_modes=False):
    """(k[f], phi[f] on z) of the synthetic Pekeris waveguide, through ModeSet.phi_src (the data contract's interpolation)."""
    ms = pekeris_modes(freqs, REC_Z_21, complex_modes=complex_modes)
    ks = [np.asarray(k) for k in ms.k]
    phis = [np.stack([ms.phi_src(f, float(zz))[0] for zz in z]) for f in range(len(freqs))]
    return ms, ks, phis


def test_round_trip_single_and_broadband(tmp_path):
    z = np.linspace(0.0, 217.0, 435)
    ms, ks, phis = _tables(INV_TONALS_3, z)
    # one broadband file: every tonal found by frequency, mode counts differ per profile
    path = tmp_path / "pekeris.mod"
    kraken_mod.write_mod(path, INV_TONALS_3, z, ks, phis, title="Pekeris 217 m")
    for f, freq in enumerate(INV_TONALS_3):
        m = kraken_mod.read_mod(path, freq)
        assert m["freq"] == freq and m["M"] == len(ks[f]) and m["title"] == "Pekeris 217 m"
        np.testing.assert_allclose(m["z"], z, rtol=1e-6)
        np.testing.assert_allclose(m["k"], ks[f], rtol=2e-7)                      # single precision on disk
        np.testing.assert_allclose(m["phi"], phis[f].real, rtol=0, atol=2e-7 * np.abs(phis[f]).max())
        assert not np.iscomplexobj(m["phi"]) and m["bottom"]["bc"] == "A" and m["top"]["bc"] == "V"
    first = kraken_mod.read_mod(path)                                             # freq = 0: the first profile
    assert first["freq"] == INV_TONALS_3[0]
    sub = kraken_mod.read_mod(path, INV_TONALS_3[2], modes=[1, 3, 999])          # mode subset, out-of-range index dropped
    full = kraken_mod.read_mod(path, INV_TONALS_3[2])
    np.testing.assert_array_equal(sub["k"], full["k"][[0, 2]])
    np.testing.assert_array_equal(sub["phi"], full["phi"][:, [0, 2]])
    # the record arithmetic: first word = record length in words, profile p advances by 3 + M + floor(4 (2M - 1) / lrecl)
    raw = path.read_bytes()
    lrecl = 4 * struct.unpack_from("<i", raw, 0)[0]
    p, counts = 5, []
    for _ in INV_TONALS_3:
        M = struct.unpack_from("<i", raw, p * lrecl)[0]
        counts.append(M)
        p += 3 + M + (4 * (2 * M - 1)) // lrecl
    assert counts == [len(k) for k in ks] and p * lrecl == len(raw)


def test_modeset_from_files_equals_modeset_from_tables(tmp_path):
    z = np.linspace(0.0, 217.0, 869)
    ms, ks, phis = _tables(INV_TONALS_3, z, complex_modes=True)                  # KRAKENC: complex mode shapes survive
    paths = []
    for f, freq in enumerate(INV_TONALS_3):                                      # one file per tonal, named as TritonOA names a run
        p = tmp_path / f"{freq:.1f}Hz.mod"
        kraken_mod.write_mod(p, [freq], z, [ks[f]], [phis[f]])
        paths.append(p)
    got = kraken_mod.modeset_from_mod(paths, INV_TONALS_3, REC_Z_21, z_pivot=ms.z_pivot)
    ref = modeset_from_tables(INV_TONALS_3, ks, [z] * 3, phis, REC_Z_21, z_pivot=ms.z_pivot)
    assert got.n_modes == ref.n_modes and got.n_rec == ref.n_rec and got.nz_tab == ref.nz_tab
    for f in range(3):
        np.testing.assert_allclose(got.k[f], ref.k[f], rtol=2e-7)
        np.testing.assert_allclose(got.phi_rec[f], ref.phi_rec[f], atol=3e-7 * np.abs(ref.phi_rec[f]).max())
        np.testing.assert_allclose(got.phi_src_tab[f], ref.phi_src_tab[f], atol=3e-7 * np.abs(ref.phi_src_tab[f]).max())
        assert np.iscomplexobj(got.phi_rec[f])
    # and the field a run_kraken call would return from it matches the analytic mode set to the table's resolution
    p_file = replica_field(got, 1, 71.11, 1.07)
    p_true = replica_field(ms, 1, 71.11, 1.07)
    assert np.abs(p_file - p_true).max() <= 2e-3 * np.abs(p_true).max()
    one = tmp_path / "all.mod"
    kraken_mod.write_mod(one, INV_TONALS_3, z, ks, phis)
    same = kraken_mod.modeset_from_mod(one, INV_TONALS_3, REC_Z_21, z_pivot=ms.z_pivot)
    for f in range(3):
        np.testing.assert_array_equal(same.phi_rec[f], got.phi_rec[f])


def test_reader_rejects_what_it_cannot_read(tmp_path):
    bad = tmp_path / "short.mod"
    bad.write_bytes(b"\x01\x02")
    with pytest.raises(ValueError):
        kraken_mod.read_mod(bad)
    z = np.linspace(0.0, 100.0, 51)
    ms, ks, phis = _tables([148.0], z)
    path = tmp_path / "a.mod"
    kraken_mod.write_mod(path, [148.0], z, ks, phis)
    raw = bytearray(path.read_bytes())
    with pytest.raises(ValueError):
        (tmp_path / "trunc.mod").write_bytes(bytes(raw[: len(raw) // 2]))
        kraken_mod.read_mod(tmp_path / "trunc.mod")
    struct.pack_into("<i", raw, 84 + 12, 2 * len(z))                            # NMat != Ntot: an elastic medium
    (tmp_path / "elastic.mod").write_bytes(bytes(raw))
    with pytest.raises(NotImplementedError):
        kraken_mod.read_mod(tmp_path / "elastic.mod")
    with pytest.raises(ValueError):
        kraken_mod.modeset_from_mod(path, [235.0], REC_Z_21)                     # tonal not in the file
    with pytest.raises(ValueError):
        kraken_mod.modeset_from_mod([path, path], [148.0], REC_Z_21)
