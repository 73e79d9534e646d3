"""Forward-process, loss and evaluation-statistics kernels against the reference's torch expressions."""
import numpy as np
import pytest
import torch

from diffusionmodelscustom_b200 import DiffusionUtils, DiffusionUtilsV2, SDFWeightedMSELoss, evaluation

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("scaled", [False, True])
def test_noise_image_bit_exact_with_injected_noise(scaled):
    """diffusion_DANRA_conditional.py:85-103 / src/diffusion_modules.py:71-99: identical x, t, eps give bit-identical x_t."""
    g = torch.Generator().manual_seed(3)
    x, eps = torch.randn(5, 1, 64, 64, generator=g), torch.randn(5, 1, 64, 64, generator=g)
    t = torch.tensor([1, 999, 500, 42, 7])
    du = DiffusionUtilsV2(1000, 1e-4, 0.02, "cpu", "linear", 64, data_scaled=scaled)
    ref_noise = eps * 0.005 if scaled else eps
    ref = torch.sqrt(du.alpha_hat[t])[:, None, None, None] * x + torch.sqrt(1 - du.alpha_hat[t])[:, None, None, None] * ref_noise
    x_t, n = du.noiseImage(x.cuda(), t.cuda(), noise=eps.cuda())
    assert torch.equal(x_t.cpu(), ref) and torch.equal(n.cpu(), ref_noise)


def test_noise_image_draws_standard_normals_keyed_by_seed():
    du = DiffusionUtils(1000, 1e-4, 0.02, "cpu")
    x = torch.zeros(8, 1, 128, 128, device="cuda")
    t = torch.full((8,), 999, device="cuda")
    _, n1 = du.noiseImage(x, t, seed=5)
    _, n2 = du.noiseImage(x, t, seed=5)
    _, n3 = du.noiseImage(x, t, seed=6)
    assert torch.equal(n1, n2) and not torch.equal(n1, n3)
    assert abs(float(n1.mean())) < 1e-2 and abs(float(n1.std()) - 1.0) < 1e-2
    assert abs(float(torch.corrcoef(torch.stack([n1[0].flatten(), n1[1].flatten()]))[0, 1])) < 2e-2   # samples independent


def test_sdf_weighted_mse_matches_the_reference_expression():
    g = torch.Generator().manual_seed(1)
    a, b = torch.randn(6, 1, 64, 64, generator=g), torch.randn(6, 1, 64, 64, generator=g)
    sdf = torch.randn(6, 1, 64, 64, generator=g) * 3
    ref = ((torch.sigmoid(sdf) * (1.0 - 0.5) + 0.5) * (a - b) ** 2).mean()
    loss = SDFWeightedMSELoss(1.0, 0.5)
    with torch.no_grad():
        got = loss(a.cuda(), b.cuda(), sdf.cuda())
    assert abs(float(got) - float(ref)) <= 2e-6 * abs(float(ref))
    assert abs(float(evaluation.weighted_mse(a.cuda(), b.cuda())) - float(((a - b) ** 2).mean())) <= 2e-6 * float(((a - b) ** 2).mean())
    ac = a.cuda().requires_grad_(True)            # with autograd: the reference's differentiable expression
    loss(ac, b.cuda(), sdf.cuda()).backward()
    assert ac.grad is not None and torch.isfinite(ac.grad).all()


def test_daily_and_pixel_errors_are_nan_aware():
    """evaluation_DANRA_conditional.py:121-122 (nanmean over the spatial dimensions) and the pixel-wise counterparts."""
    g = torch.Generator().manual_seed(2)
    gen, ev = torch.randn(9, 64, 64, generator=g) * 4 + 8, torch.randn(9, 64, 64, generator=g) * 4 + 8
    ev[:, :5, :7] = float("nan")                 # masked sea points
    ev[3] = float("nan")                         # a day without truth
    mae, rmse = evaluation.daily_errors(gen.cuda(), ev.cuda())
    ref_mae = torch.abs(gen - ev).nanmean(dim=(1, 2))
    ref_rmse = torch.sqrt(torch.square(gen - ev).nanmean(dim=(1, 2)))
    np.testing.assert_allclose(mae.cpu().numpy(), ref_mae.numpy(), rtol=1e-5, equal_nan=True)
    np.testing.assert_allclose(rmse.cpu().numpy(), ref_rmse.numpy(), rtol=1e-5, equal_nan=True)
    pm, pr, pb = evaluation.pixel_errors(gen.cuda(), ev.cuda())
    np.testing.assert_allclose(pm.cpu().numpy(), torch.abs(gen - ev).nanmean(dim=0).numpy(), rtol=1e-5, equal_nan=True)
    np.testing.assert_allclose(pr.cpu().numpy(), torch.sqrt(torch.square(gen - ev).nanmean(dim=0)).numpy(), rtol=1e-5, equal_nan=True)
    np.testing.assert_allclose(pb.cpu().numpy(), (gen - ev).nanmean(dim=0).numpy(), rtol=1e-4, atol=1e-6, equal_nan=True)
    assert abs(float(evaluation.bias(gen.cuda(), ev.cuda())) - float(np.nanmean(gen.numpy()) - np.nanmean(ev.numpy()))) < 1e-4


@pytest.mark.parametrize("bins,rng", [(50, (0.0, 25.0)), (1200, (-3.0, 3.0)), (3000, (-40.0, 60.0))])
def test_histogram_matches_numpy(bins, rng):
    g = torch.Generator().manual_seed(bins)
    x = torch.randn(200000, generator=g) * 6 + 4
    x[::97] = float("nan")
    x[5] = rng[1]                               # right edge belongs to the last bin
    ref, _ = np.histogram(x.numpy()[~np.isnan(x.numpy())], bins=bins, range=rng)
    got = evaluation.histogram(x.cuda(), bins, rng).cpu().numpy()
    # bin edges are computed in fp32 on the device and fp64 in numpy: values within 1 ulp of an edge may land next door
    assert got.sum() == ref.sum() and np.abs(got - ref).sum() <= 4
