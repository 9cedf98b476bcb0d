"""BAL reader / writer: native parser vs the Python reader vs the reference fixture inputs."""
import os
import time

import numpy as np
import pytest

from conftest import golden_problem, load_golden

FIELDS = ["cam_id", "lmk_id", "z", "cam_means", "lmk_means", "K4"]


def same(a, b):
    return all(np.array_equal(getattr(a, k), getattr(b, k)) for k in FIELDS)


def test_roundtrip_text_and_npz(tmp_path, built_library):
    from gbp_b200 import balio
    from gbp_b200.synthetic import make_synthetic
    p = make_synthetic(7, 300, 4, seed=2)
    for name in ("p.txt", "p.npz"):
        path = str(tmp_path / name)
        balio.write_bal(path, p, ["synthetic", "two comment lines"])
        assert same(balio.read_bal(path), p)
    assert same(balio.read_bal_python(str(tmp_path / "p.txt")), p)
    t = balio.read_bal(str(tmp_path / "p.txt")).as_tuple()
    assert t[0] == 7 and t[1] == 300 and t[2] == p.n_edges and t[8].shape == (3, 3) and isinstance(t[6], list)


def test_fixture_inputs_survive_the_text_format(tmp_path, built_library):
    """fr1desk_vsmall: fixture arrays -> BAL text -> native parser gives back the same bits."""
    from gbp_b200 import balio
    G = load_golden("fr1desk_vsmall")
    p = golden_problem(G)
    path = str(tmp_path / "vsmall.txt")
    balio.write_bal(path, p)
    q = balio.read_bal(path)
    assert same(q, p) and (q.n_keyframes, q.n_points, q.n_edges) == (10, 640, 1801)


def test_reference_data_files_if_present(built_library):
    """In the build container: the native parser reproduces what the reference's reader produced (fixture in_* arrays)."""
    from gbp_b200 import balio
    path = "/root/reference/data/fr1desk.txt"
    if not os.path.exists(path):
        pytest.skip("reference checkout not present")
    G = load_golden("fr1desk")
    assert same(balio.read_bal(path), golden_problem(G))
    assert same(balio.read_bal_python(path), golden_problem(G))


def test_reference_acceptance_rules(tmp_path, built_library):
    """Comment / blank header lines, extra columns on measurement lines, trailing tokens on parameter lines,
    CRLF line ends, '+' signs and exponents."""
    from gbp_b200 import balio
    text = ("# Dataset: x\r\n#\n\n   \n# Camera noise: 0.07 m\n"
            "2 1 3\n500.5 +501 320 2.4e2\n"
            "0 0   1.5e+02 2.0 extra tokens here\n1 0 3 4\n0 0\t-5.25 +6\n"
            + "".join(f"{v} trailing\n" for v in range(12)) + "0.5\n-1e-3\n7\n")
    path = tmp_path / "odd.txt"
    path.write_text(text)
    for rd in (balio.read_bal, balio.read_bal_python):
        p = rd(str(path))
        assert p.cam_id.tolist() == [0, 1, 0] and p.lmk_id.tolist() == [0, 0, 0]
        np.testing.assert_array_equal(p.z, [[150.0, 2.0], [3.0, 4.0], [-5.25, 6.0]])
        np.testing.assert_array_equal(p.cam_means, np.arange(12.0).reshape(2, 6))
        np.testing.assert_array_equal(p.lmk_means, [[0.5, -1e-3, 7.0]])
        np.testing.assert_array_equal(p.K4, [500.5, 501.0, 320.0, 240.0])


def test_parser_errors(tmp_path, built_library):
    from gbp_b200 import balio, _lib
    bad = tmp_path / "bad.txt"
    bad.write_text("1 1 2\n1 1 1 1\n0 0 1 2\n")          # truncated
    with pytest.raises(_lib.GbpError, match="file ends|bad"):
        balio.read_bal(str(bad))
    bad.write_text("1 1 1\n1 1 1 1\n0 zero 1 2\n" + "0\n" * 9)
    with pytest.raises(_lib.GbpError, match="bad measurement line"):
        balio.read_bal(str(bad))
    with pytest.raises(_lib.GbpError, match="cannot open"):
        balio.read_bal(str(tmp_path / "missing.txt"))


def test_native_parser_is_fast(tmp_path, built_library):
    from gbp_b200 import balio
    from gbp_b200.synthetic import make_synthetic
    p = make_synthetic(50, 40000, 10, seed=0)          # 400k measurements
    path = str(tmp_path / "big.txt")
    balio.write_bal(path, p)
    t0 = time.perf_counter(); q = balio.read_bal(path); t_native = time.perf_counter() - t0
    t0 = time.perf_counter(); r = balio.read_bal_python(path); t_py = time.perf_counter() - t0
    assert same(q, p) and same(r, p)
    assert t_native < t_py, (t_native, t_py)
    print(f"native {t_native:.3f}s python {t_py:.3f}s")
