"""The oracle against (a) the reference's own known-answer vectors (restated from the
reference's *_test.py files, cited per test) and (b) fixtures produced by running the
unmodified reference (tests/golden/make_golden.py)."""
import math
import os

import numpy as np
import pytest

import v3s_oracle as orc


def hybrid_ok(x, y, p):
    return bool(np.all(orc.almost_equal_hybrid(x, y, p)))


# ---------------- reference known-answer vectors ------------------------------------


def test_knn_from_round_distance_golden():
    # distance_test.py:11-41
    B = np.array([[0, -0.8, -0.6], [-0.8, 0, 0.6], [0.8, -0.6, 0]])
    S2, N = orc.knn_from_round_distance(B, 2)
    assert np.array_equal(N, [[1, 2], [0, 1], [1, 2]])
    assert np.array_equal(S2, [[-0.8, -0.6], [-0.8, 0], [-0.6, 0]])


def test_convert_to_round_distance_properties():
    # distance_test.py:44-57
    A = np.array([[3, 4], [-3, 4], [6, -8]])
    B = orc.convert_to_round_distance(A)
    assert B.shape == (3, 3)
    assert hybrid_ok(B, B.T, 1e-15)
    assert np.all(np.diag(B) <= 1e-6) and np.all(B <= np.pi)


def test_set_norm_to_one_and_remove_offset_golden():
    # distance_test.py:60-79
    assert np.allclose(orc.set_norm_to_one(np.array([[3, 4], [-3, 4], [6, -8]])),
                       [[0.6, 0.8], [-0.6, 0.8], [0.6, -0.8]], atol=1e-12, rtol=0)
    assert np.allclose(orc.remove_offset(np.array([[1, 2], [-1, -1], [3, 5]])),
                       [[0, 0], [-2, -3], [2, 3]], atol=1e-12, rtol=0)


S2_4 = np.array([[0.0, 16.0, 9.0, 25.0], [16.0, 0.0, 25.0, 9.0], [9.0, 25.0, 0.0, 16.0], [25.0, 9.0, 16.0, 0.0]])
N_4 = np.array([[1, 3, 2, 4], [2, 4, 1, 3], [3, 1, 4, 2], [4, 2, 3, 1]]) - 1
W1 = np.array([[1.0, 0.48675, 0.27804, 0.13534], [0.13534, 0.27804, 0.48675, 1.0],
               [0.13534, 0.27804, 0.48675, 1.0], [1.0, 0.48675, 0.27804, 0.13534]])
W2 = np.array([[0.27697, 0.08615, 0.057246, 0.15723], [0.08615, 0.077009, 0.10591, 0.20589],
               [0.057246, 0.10591, 0.13482, 0.17699], [0.15723, 0.20589, 0.17699, 0.037484]])


def test_s2_to_w_matrix_golden():
    # diffusion_test.py:27-55
    W = orc.s2_to_w_matrix(S2_4.copy(), N_4, 1.0, s2_median_index=3, index_offset=0)
    assert hybrid_ok(W, W1, 1e-4)


def test_anisotropic_norm_golden():
    # diffusion_test.py:58-80
    assert hybrid_ok(orc.anisotropic_norm(W1.copy()), W2, 1e-4)


def test_dm_cov_golden():
    # diffusion_test.py:83-113
    P_exp = np.array([[0.47952, 0.16448, 0.1093, 0.27221], [0.16448, 0.16213, 0.22299, 0.3931],
                      [0.1093, 0.22299, 0.28385, 0.33791], [0.27221, 0.3931, 0.33791, 0.064897]])
    D_exp = np.diag([1.3158, 1.4510, 1.4510, 1.3158])
    P, D = orc.dm_cov(W2.copy())
    assert hybrid_ok(P, P_exp, 1e-4) and hybrid_ok(D, D_exp, 1e-4)


def test_diffusion_coordinate_golden_up_to_sign():
    # diffusion_test.py:116-149; the reference test itself is sign-flaky,
    # so columns are compared up to sign.
    Ps_exp = np.array([[-0.689225, 0.060209], [-0.689225, 0.235587], [-0.689225, 0.115495], [-0.689225, -0.348909]])
    Ps, Lam = orc.diffusion_coordinate(S2_4.copy(), N_4, 1.0, 2, False)
    assert hybrid_ok(Lam, [1.0, -0.3242], 1e-4)
    for j in range(2):
        assert hybrid_ok(Ps[:, j], Ps_exp[:, j], 2e-4) or hybrid_ok(-Ps[:, j], Ps_exp[:, j], 2e-4)


def test_lsq_linear_golden():
    # compute_test.py:10-29
    M = np.array([[1, -1], [2, 3]])
    A = np.array([[0, -4], [-1, 1], [5, 0], [1, 2]])
    assert hybrid_ok(orc.lsq_linear(A, A @ M), M, 1e-13)


def test_rotation_rz90_golden():
    # rotation_test.py:91-153
    Rz = np.array([[0.0, -1.0, 0.0], [1.0, 0.0, 0.0], [0.0, 0.0, 1.0]])
    assert hybrid_ok(orc.axis_to_rotmat((np.pi / 2) * np.array([0.0, 0.0, 1.0])), Rz, 1e-15)
    assert hybrid_ok(orc.rot_mat_to_axis(Rz), (np.pi / 2) * np.array([[0.0], [0.0], [1.0]]), 1e-15)
    assert hybrid_ok(orc.rot_mat_to_quat(Rz), math.sqrt(0.5) * np.array([[1.0], [0.0], [0.0], [1.0]]), 1e-15)
    assert hybrid_ok(orc.rot_mat_to_quat_batch(Rz[None]), math.sqrt(0.5) * np.array([[1.0], [0.0], [0.0], [1.0]]), 1e-15)


def test_mean_geodesic_statistical():
    # rotation_test.py:72-88 (seeded here; N=12^3 keeps it fast)
    rng = np.random.default_rng(0)
    q1 = orc.uniform_so3_hopf(12 ** 3)
    for deg in (40, 10, 5, 1, 0.1):
        q2 = q1 + 0.99 * (np.pi / 180) * (0.5 * deg) * rng.standard_normal(q1.shape)
        assert orc.mean_geodesic_distance_degrees(q1, q2) < deg


def test_shapes_like_reference_smoke():
    # rotation_test.py:12-69, projection_test.py:76-88, solution_test.py:12-56 (shapes)
    assert orc.uniform_so3_hopf(1).shape == (4, 1)
    R, Axis, Q = orc.all_rot_variables(8)
    assert R.shape == (9, 8) and Axis.shape == (3, 8) and Q.shape == (4, 8)
    with pytest.raises(Exception):
        orc.uniform_so3_hopf(10)
    assert orc.experiment_parameters(31)["Experiment"]["N_p"] == 63
    assert orc.validate_parameters(orc.default_parameters())
    assert not orc.validate_parameters({})


def test_guards():
    with pytest.raises(Exception, match="Non-square"):
        orc.knn_from_round_distance(np.zeros((2, 3)), 1)
    with pytest.raises(Exception, match="too high"):
        orc.knn_from_round_distance(np.zeros((3, 3)), 4)
    with pytest.raises(Exception, match="S2_Max = 0"):
        orc.s2_to_w_matrix(np.zeros((4, 4)), N_4, 1.0)


# ---------------- fixtures generated from the unmodified reference -------------------


@pytest.mark.parametrize("n", [7, 15, 31])
def test_single_images_vs_reference(golden_dir, n):
    g = np.load(os.path.join(golden_dir, f"images_n{n}.npz"))
    assert np.array_equal(orc.synthesize_3d(n)["ED"], g["ED"])
    imgs = orc.generate_from_rotations(g["R9"], n)
    assert np.max(np.abs(imgs - g["Images"])) <= 1e-12 * np.max(np.abs(g["Images"]))
    er = orc.rotate_structure_index(g["ED"], g["R9"][:, 0].reshape(3, 3).T)
    assert np.max(np.abs(er - g["ED_rot0"])) < 1e-13


def test_hopf_rotations_vs_reference(golden_dir):
    g = np.load(os.path.join(golden_dir, "hopf_8.npz"))
    R, Axis, Quat = orc.all_rot_variables(512)
    for a, b in ((R, g["R"]), (Axis, g["Axis"]), (Quat, g["Quat"])):
        assert np.allclose(a, b, rtol=0, atol=1e-15)


def test_default_config_vs_reference(golden_dir):
    g = np.load(os.path.join(golden_dir, "default_cfg.npz"))
    Images = orc.generate(8, 7)[0]
    assert np.max(np.abs(Images - g["Images"])) <= 1e-12 * np.max(np.abs(g["Images"]))
    B = orc.convert_to_round_distance(np.asfortranarray(g["Images"]))  # the reference passes Images.T
    assert np.allclose(B[0], g["B_row0"], rtol=0, atol=1e-12) and np.allclose(B[100], g["B_row100"], rtol=0, atol=1e-12)
    S2, N = orc.knn_from_round_distance(B, 24)
    assert np.allclose(S2, g["S2"], rtol=0, atol=1e-12)
    # indices: compared where the reference's distances are not tied (duplicate images exist)
    untied = np.ones_like(S2, dtype=bool)
    untied[:, 1:] &= np.abs(np.diff(g["S2"], axis=1)) > 1e-9
    untied[:, :-1] &= np.abs(np.diff(g["S2"], axis=1)) > 1e-9
    assert np.array_equal(N[untied], g["N"][untied])
    S2m = g["S2"].copy()
    W = orc.s2_to_w_matrix(S2m, g["N"], 0.7)
    assert np.array_equal(S2m, g["S2_mutated"])
    assert np.allclose(W.sum(1), g["W_rowsum"], rtol=1e-14) and np.allclose(W.sum(0), g["W_colsum"], rtol=1e-14)
    W2 = orc.anisotropic_norm(W)
    assert np.allclose(W2.sum(1), g["W2_rowsum"], rtol=1e-13)
    P, di = orc.dm_cov(W2, dense_D=False)
    assert np.allclose(di, g["D_diag"], rtol=1e-14)
    assert np.allclose(P @ g["z"], g["P_ep_z"], rtol=1e-12, atol=1e-13)
    assert np.allclose(P[0], g["P_ep_row0"], rtol=1e-13) and np.allclose(P[511], g["P_ep_row511"], rtol=1e-13)
    lam = np.sort(np.linalg.eigvalsh(P))[::-1][:200]
    assert np.allclose(lam, g["lam_top"], atol=1e-10)


def test_connected_case_vs_reference(golden_dir):
    g = np.load(os.path.join(golden_dir, "connected_n15.npz"))
    Imgs = g["Images_f32"].astype(np.float64)
    S2, N = orc.knn(Imgs, int(g["k"]))
    assert np.allclose(S2, g["S2"], rtol=0, atol=1e-12) and np.array_equal(N, g["N"])
    Ps, Lam, _, _, P, di = orc.diffusion_coordinate(S2.copy(), N, float(g["eps"]), return_all=True)
    assert np.allclose(Lam, g["Lambda"], rtol=1e-12)
    assert np.allclose(P @ g["z"], g["P_ep_z"], rtol=1e-11, atol=1e-13)
    assert np.max(orc.principal_angles(Ps, g["Ps"])) < 1e-10
    c, _, sig, recon_pre, _, _, _ = orc.dm_fit_all(Ps, g["R9"], True, return_all=True)
    assert np.max(np.abs(recon_pre - g["Recon_pre"])) < 1e-10
    assert abs(sig - float(g["Sigma_T"])) < 1e-6


def test_helpers_vs_reference(golden_dir):
    g = np.load(os.path.join(golden_dir, "helpers.npz"))
    assert np.allclose(orc.lsq_linear(g["A"], g["b"]), g["M"], rtol=1e-12)
    pol = np.array([orc.polar_decompose(m) for m in g["Ms"]])
    assert np.allclose(pol, g["polar_compat"], atol=1e-13)
    assert np.allclose(orc.rot_mat_to_quat_batch(g["Rb"]), g["quats"], atol=1e-14)
    assert abs(orc.mean_geodesic_distance_degrees(g["quats"], g["q2"]) - float(g["sigma"])) < 1e-10
