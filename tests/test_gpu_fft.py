"""GPU parity of the FFT-derived fields (c2g_fft_derivative, grid3%fft) and of the NCIPLOT FOURIER loop
(c2g_nci_rdg_fourier) against the oracle, through the C ABI.

Tolerance: these are floating-point transforms; cuFFT and the oracle's DFT use different factorisations, so
each output agrees to rounding of the transform: |gpu - oracle| <= 1e-12 * max|oracle| (written below).  The
RDG built from them keeps north_star's 1e-12 (relative to max RDG away from rho -> 0 amplification: the same
rule as the tricubic test)."""
import numpy as np
import pytest

import cases
from critic2_b200 import capi
from oracle import oracle as orc

pytestmark = pytest.mark.gpu

FFT_TOL = 1e-12


@pytest.fixture(scope="module")
def ctx():
    c = capi.Context(0)
    yield c
    c.close()


@pytest.mark.parametrize("name", ["cubic48", "triclinic", "odd_dims"])
def test_fft_derivative_all_codes(ctx, name):
    c = cases.make_case(name)
    h = ctx.upload(c["f"])
    for w in capi.FT_CODES:
        ref = orc.fft_derivative(c["f"], c["x2c"], w)
        ho = ctx.fft_derivative(h, c["x2c"], w)
        out = ctx.download(ho, c["f"].shape)
        ctx.free(ho)
        err = np.abs(out - ref).max() / np.abs(ref).max()
        assert err <= FFT_TOL, f"{name} {w}: {err:.3e}"
    ctx.free(h)


def test_fft_random_field_even_and_odd_nyquist(ctx):
    """White noise has full weight on the Nyquist planes, where the real-part rule matters."""
    rng = np.random.default_rng(11)
    x2c = np.array([[5.0, 0.3, 0.1], [0.0, 4.5, 0.2], [0.0, 0.0, 6.0]])
    for n in [(12, 10, 9), (8, 8, 8), (7, 5, 3), (16, 6, 15)]:
        f = np.asfortranarray(rng.standard_normal(n))
        h = ctx.upload(f)
        for w in capi.FT_CODES:
            ref = orc.fft_derivative(f, x2c, w)
            ho = ctx.fft_derivative(h, x2c, w)
            out = ctx.download(ho, f.shape)
            ctx.free(ho)
            err = np.abs(out - ref).max() / np.abs(ref).max()
            assert err <= FFT_TOL, f"{n} {w}: {err:.3e}"
        ctx.free(h)


def test_fft_laplacian_integrates_to_zero_per_cell_and_feeds_integrable(ctx):
    """INTEGRABLE lap on Bader basins (the reference's default integrable set, systemmod@proc.f90:187-190): the
    FFT Laplacian has no G = 0 component, so the basin integrals sum to ~0; compare with the oracle per basin."""
    import systems as S
    import helpers
    c = cases.make_case("cubic48")
    f, x2c = c["f"], c["x2c"]
    idg, nattr, _, _ = orc.bader_integrate(f, x2c, atoms=c["atoms"])
    lap_o = orc.fft_derivative(f, x2c, "lap")
    vref, pref = orc.integrate_bader(idg, [f, lap_o], nattr, S.omega(x2c))
    h = ctx.upload(f)
    hl = ctx.fft_derivative(h, x2c, "lap")
    _, car2lat, lid = orc.bader_metrics(x2c, c["n"])
    b = ctx.bader_assign(h, car2lat, lid)
    mp, na, _ = helpers.assign_attractors(b.maxima(), c["n"], x2c, c["atoms"])
    assert na == nattr
    b.set_map(na, mp)
    vol, ps = ctx.integrate(b, [h, hl], S.omega(x2c))
    assert np.abs(ps[:, 0] - pref[:, 0]).max() <= 1e-10 * np.abs(pref[:, 0]).max()
    scale = np.abs(lap_o).sum() * S.omega(x2c) / f.size
    assert np.abs(ps[:, 1] - pref[:, 1]).max() <= 1e-10 * scale
    assert abs(ps[:, 1].sum()) <= 1e-10 * scale
    b.free()
    ctx.free(h)
    ctx.free(hl)


def _fourier_compare(crho, cgrad, crho_o, cgrad_o, der_o):
    rel = np.abs(cgrad - cgrad_o) / np.abs(cgrad_o).max()
    assert rel.max() <= 1e-12 or (np.abs(cgrad - cgrad_o) <= 1e-11 * np.abs(cgrad_o)).all(), rel.max()
    assert np.abs(np.abs(crho) - np.abs(crho_o)).max() <= 1e-12 * np.abs(crho_o).max()
    flips = np.sign(crho) != np.sign(crho_o)
    # the sign comes from count(Hii > 0): it may only differ where some Hii is at rounding level
    assert flips.mean() <= 1e-4, flips.mean()


@pytest.mark.parametrize("name", ["cubic48", "triclinic"])
def test_nci_fourier_node_aligned(ctx, name):
    c = cases.make_case(name)
    f, x2c = c["f"], c["x2c"]
    der_o = tuple(orc.fft_derivative(f, x2c, w) for w in ("grad", "xx", "yy", "zz"))
    crho_o, cgrad_o = orc.nci_rdg_fourier(f, x2c, derived=der_o)
    h = ctx.upload(f)
    hd = [ctx.fft_derivative(h, x2c, w) for w in ("grad", "xx", "yy", "zz")]
    crho, cgrad = ctx.nci_rdg_fourier([h] + hd, x2c, c["n"])
    _fourier_compare(crho, cgrad, crho_o, cgrad_o, der_o)
    for hh in [h] + hd:
        ctx.free(hh)


def test_nci_fourier_general_lattice_same_derived_grids(ctx):
    """Off-node lattice: tricubic rho + trilinear derived grids.  The derived grids are uploaded from the oracle so
    that only the interpolation path is compared (bit-level agreement of the inputs)."""
    c = cases.make_case("triclinic")
    f, x2c = c["f"], c["x2c"]
    der_o = tuple(orc.fft_derivative(f, x2c, w) for w in ("grad", "xx", "yy", "zz"))
    nstep = (21, 17, 19)
    x0 = x2c @ np.array([0.013, -0.021, 1.034])
    xmat = x2c / np.array(nstep, dtype=float)[None, :] * 0.93
    crho_o, cgrad_o = orc.nci_rdg_fourier(f, x2c, nstep=nstep, x0=x0, xmat=xmat, derived=der_o)
    hs = [ctx.upload(f)] + [ctx.upload(d) for d in der_o]
    crho, cgrad = ctx.nci_rdg_fourier(hs, x2c, c["n"], nstep=nstep, x0=x0, xmat=xmat)
    assert np.abs(cgrad - cgrad_o).max() <= 1e-12 * np.abs(cgrad_o).max()
    assert np.abs(crho - crho_o).max() <= 1e-12 * np.abs(crho_o).max()
    for hh in hs:
        ctx.free(hh)


def test_fft_and_fourier_error_paths(ctx):
    """Bad calls fail with a status and a message (ferror on the Fortran side), they do not crash."""
    c = cases.make_case("tiny")
    h = ctx.upload(c["f"])
    with pytest.raises(capi.C2GError):
        ctx.fft_derivative(h, c["x2c"], 99)              # not an ifformat_as_ft_* code
    with pytest.raises(capi.C2GError):
        ctx.fft_derivative(12345, c["x2c"], "lap")       # invalid handle
    with pytest.raises(capi.C2GError):
        ctx.fft_derivative(h, np.zeros((3, 3)), "lap")   # singular cell
    other = ctx.upload(np.asfortranarray(np.ones((4, 5, 6))))
    with pytest.raises(capi.C2GError):
        ctx.nci_rdg_fourier([h, h, h, h, other], c["x2c"], c["n"])   # derived grid of another shape
    ctx.free(h); ctx.free(other)
