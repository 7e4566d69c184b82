"""GMM emission kernels (csrc/gmm_exact.cu) against the reference golden vectors and the numpy oracle."""
import numpy as np
import pytest

from helpers import assert_close

pytestmark = pytest.mark.gpu


def test_gmm_module_functions_match_reference(golden):
    from lnasr import gmm
    g = golden("gmm")
    x, mu, Sigma = g["x"], g["mu"], g["Sigma"]
    np.testing.assert_allclose(gmm.gaussian_multi_distribution_log(x, mu, Sigma), g["gmlog"], rtol=1e-10, atol=1e-10)
    np.testing.assert_allclose(gmm.gaussian_multi_distribution(x, mu, Sigma), g["gm"], rtol=1e-9, atol=1e-300)
    np.testing.assert_allclose(gmm.gaussian_mixture_distribution_log(np.log(g["c"]), x, g["mus"], g["Sigmas"]), g["gmixlog"],
                               rtol=1e-10, atol=1e-10)
    out = gmm.gaussian_mixture_distribution(g["c"], x, g["mus"], g["Sigmas"])
    assert out.shape == g["gmix"].shape
    np.testing.assert_allclose(out, g["gmix"], rtol=1e-9)
    np.testing.assert_allclose(gmm.gaussian_distribution_log(x[:, 0], 0.3, 1.7), g["g1log"], rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(gmm.gaussian_distribution(x[:, 0], 0.3, 1.7), g["g1"], rtol=1e-9)


def test_map_b_full_covariance_matches_reference(golden):
    """GMMHMM._map_b with the reference's full si (gmmhmm.py:46-67)."""
    from lnasr.gmmhmm import GMMHMM
    g = golden("gmmhmm_5_4_39")
    m = GMMHMM(5, 4, 39, A=g["A"], pi=g["pi"], w=g["w"], mu=g["mu"], si=g["si"])
    m._map_b(g["obs"])
    assert m.b.shape == (5, 135) and m.bm.shape == (5, 4, 135)
    assert_close(m.b, g["b"], 1e-9, "b")
    assert_close(m.bm, g["bm"], 1e-9, "bm")


def test_map_b_diagonal_matches_reference(golden):
    from lnasr.gmmhmm import GMMHMM
    g = golden("gmmhmm_5_4_39")
    var = g["dg_var"]
    si = var[..., None] * np.eye(39)
    m = GMMHMM(5, 4, 39, A=g["A"], pi=g["pi"], w=g["w"], mu=g["mu"], si=si)
    m._map_b(g["obs"])
    assert_close(m.b, g["dg_b"], 1e-10, "b")
    assert_close(m.bm, g["dg_bm"], 1e-10, "bm")


@pytest.fixture
def gmm_kernel_switch():
    """lnasr_debug_set_gmm_simt: 0 = default (persistent DMMA when the image fits shared memory, else streaming DMMA, else
    SIMT), 1 = SIMT, 2 = streaming DMMA.  Restored to 0 afterwards."""
    from lnasr import _capi
    yield _capi.lib().lnasr_debug_set_gmm_simt
    _capi.lib().lnasr_debug_set_gmm_simt(0)


@pytest.mark.parametrize("kernel", [0, 1, 2])
@pytest.mark.parametrize("n,m,d,T", [(1, 1, 1, 7), (3, 8, 39, 300), (50, 4, 39, 257), (7, 3, 2, 40), (200, 16, 39, 130),
                                     (9, 2, 13, 1000), (33, 5, 39, 77), (400, 4, 39, 33), (6, 1, 39, 129)])
def test_exact_kernel_against_oracle(n, m, d, T, kernel, gmm_kernel_switch):
    import torch
    from lnasr.gmm import PackedGmm
    from oracle import lnasr_oracle as O
    gmm_kernel_switch(kernel)
    rng = np.random.default_rng(n * 1000 + m)
    mdl = O.random_gmmhmm(n, m, d, rng)
    x = rng.standard_normal((T, d)) * 0.5
    b_ref, bm_ref = O.map_b(x, mdl["w"], mdl["mu"], mdl["var"])
    g = PackedGmm(mdl["w"], mdl["mu"], mdl["var"])
    b, bm = g.loglik(torch.from_numpy(x).cuda(), want_bm=True)
    assert_close(b.cpu().numpy().T, b_ref, 1e-10, "b")
    assert_close(bm.cpu().numpy().reshape(T, n, m).transpose(1, 2, 0), bm_ref, 1e-10, "bm")
    # float32 features in, float32 emissions out: the layout the batched pipeline uses
    b32, _ = g.loglik(torch.from_numpy(x.astype(np.float32)).cuda(), out_dtype=torch.float32)
    b_ref32, _ = O.map_b(x.astype(np.float32).astype(np.float64), mdl["w"], mdl["mu"], mdl["var"])
    assert_close(b32.cpu().numpy().T, b_ref32, 1e-6, "b32")


def test_zero_weight_mixture_and_empty_input():
    import torch
    from lnasr.gmm import PackedGmm
    from oracle import lnasr_oracle as O
    rng = np.random.default_rng(3)
    mdl = O.random_gmmhmm(2, 3, 4, rng)
    mdl["w"][0, 1] = -np.inf
    mdl["w"][1, :] = -np.inf                      # a state with no mass: b = -inf
    x = rng.standard_normal((9, 4))
    with np.errstate(all="ignore"):
        b_ref, _ = O.map_b(x, mdl["w"], mdl["mu"], mdl["var"])
    g = PackedGmm(mdl["w"], mdl["mu"], mdl["var"])
    b, _ = g.loglik(torch.from_numpy(x).cuda())
    assert_close(b.cpu().numpy().T, b_ref, 1e-10, "b")
    b0, _ = g.loglik(torch.zeros((0, 4), dtype=torch.float64, device="cuda"))
    assert b0.shape == (0, 2)
