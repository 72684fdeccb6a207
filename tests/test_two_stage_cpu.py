"""CPU checks of the numpy statement (tools/proto_two_stage.py) of the two-stage tridiagonalisation that csrc/sbr.cu implements
block for block: compact-WY band reduction (k_sbr_qr / av / vtz / st / w / r2k) and bulge chasing (k_sbr_chase).  The GPU tests
(test_tps_gpu.py) compare the kernels with the one-stage reduction; these tests pin the algorithm itself - in particular the
dependency rule the persistent chase kernel spins on - against LAPACK."""
import importlib.util
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec = importlib.util.spec_from_file_location("proto_two_stage", os.path.join(ROOT, "tools", "proto_two_stage.py"))
proto = importlib.util.module_from_spec(spec)
spec.loader.exec_module(proto)


@pytest.mark.parametrize("m,b", [(3, 8), (10, 8), (34, 32), (35, 32), (37, 8), (65, 8), (131, 32), (200, 32)])
def test_two_stage_matches_lapack(m, b):
    """eigenvalues of the band matrix and of the final tridiagonal, |Q'z| = |z|, RSS(lambda) through Q'z"""
    r = proto.check(m, b, L=2, seed=m)
    assert r["below_band"] == 0.0
    assert r["ev_band"] < 5e-14 and r["ev_tri"] < 5e-14
    assert r["znorm"] < 1e-13 and r["rss_rel"] < 1e-11


@pytest.mark.parametrize("m,b,seed", [(90, 8, 3), (150, 32, 4), (70, 8, 11)])
def test_chase_dependency_rule_is_sufficient(m, b, seed):
    """Any schedule that only honours 'sweep s runs step k once sweep s-1 has finished step k+1' (what k_sbr_chase waits for)
    gives the same tridiagonal matrix as running the sweeps one after the other - bit for bit."""
    rng = np.random.default_rng(seed)
    A = rng.standard_normal((m, m))
    A = A @ A.T
    z = rng.standard_normal((m, 2))
    A1, z1 = A.copy(), z.copy()
    proto.stage1(A1, z1, b)
    B_seq, z_seq = proto.to_band(A1, b), z1.copy()
    B_rnd, z_rnd = B_seq.copy(), z1.copy()
    d0, e0 = proto.stage2(B_seq, z_seq, m, b)
    d1, e1 = proto.stage2(B_rnd, z_rnd, m, b, rng=np.random.default_rng(seed + 100))
    np.testing.assert_array_equal(d0, d1)
    np.testing.assert_array_equal(e0, e1)
    np.testing.assert_array_equal(z_seq, z_rnd)


def test_panel_qr_is_a_qr():
    """the fused one-reduction-per-column panel QR: Q = I - V T V' is orthogonal and Q'P = [R; 0]"""
    rng = np.random.default_rng(5)
    for r, b in [(40, 8), (9, 8), (5, 8), (100, 32)]:
        P = rng.standard_normal((r, b))
        X = P.copy()
        V, T, nref = proto.panel_qr(X)
        Q = np.eye(r) - V @ T @ V.T
        assert np.abs(Q.T @ Q - np.eye(r)).max() < 1e-13
        QtP = Q.T @ P
        assert np.abs(np.tril(QtP, -1)).max() < 1e-12 * np.abs(P).max()
        assert np.abs(np.triu(QtP)[:min(r, b)] - np.triu(X)[:min(r, b)]).max() < 1e-12


@pytest.mark.parametrize("m,b", [(40, 8), (131, 32), (200, 32), (70, 32)])
def test_coefficients_from_the_band_form(m, b):
    """(M + lambda I)^-1 z = Q1 (B + lambda I)^-1 Q1'z with a block band Cholesky: the planned replacement of the dense
    Cholesky at the selected lambda agrees with the dense solve."""
    rng = np.random.default_rng(m)
    Qr, _ = np.linalg.qr(rng.standard_normal((m, m)))
    M = (Qr * (6.0 * np.exp(-np.linspace(0, 20, m)))) @ Qr.T
    M = 0.5 * (M + M.T)
    z = rng.standard_normal((m, 2))
    lam = 3e-3
    ref = np.linalg.solve(M + lam * np.eye(m), z)
    got = proto.coefficients_from_band(M, z, lam, b)
    assert np.abs(got - ref).max() < 1e-10 * np.abs(ref).max()


@pytest.mark.parametrize("m,b,g,seed", [(90, 8, 4, 1), (150, 32, 4, 2), (70, 8, 3, 3), (64, 8, 8, 5), (40, 32, 4, 6)])
def test_grouped_chase_with_a_row_window(m, b, g, seed):
    """g sweeps per CTA on a sliding window of band ROWS (the planned shared-memory variant of k_sbr_chase): loading rows only
    after the previous group has retired them and retiring rows no sweep of the group touches again reproduces the sequential
    chase bit for bit under a random interleaving of the groups, with at most ~(2 g - 1) row blocks resident."""
    rng = np.random.default_rng(seed)
    A = rng.standard_normal((m, m))
    A = A @ A.T
    z = rng.standard_normal((m, 2))
    A1, z1 = A.copy(), z.copy()
    proto.stage1(A1, z1, b)
    B_seq, z_seq = proto.to_band(A1, b), z1.copy()
    B_grp, z_grp = B_seq.copy(), z1.copy()
    d0, e0 = proto.stage2(B_seq, z_seq, m, b)
    d1, e1 = proto.stage2_grouped(B_grp, z_grp, m, b, g, np.random.default_rng(seed + 50))
    np.testing.assert_array_equal(d0, d1)
    np.testing.assert_array_equal(e0, e1)
    np.testing.assert_array_equal(z_seq, z_grp)
    assert proto.stage2_grouped.max_resident <= (2 * g) * b + g
