"""CPU tests of the oracle's Bader restatement (bader@proc.f90) -- no GPU."""
import hashlib
import json
import os

import numpy as np
import pytest

import cases
import helpers as H
import systems as S
from oracle import oracle as orc

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "oracle_golden.json")


def _atom_labels_from_terminals(term, atoms, n):
    ai = np.round(atoms * np.array(n)).astype(int) % np.array(n)
    alin = ai[:, 0] + n[0] * (ai[:, 1] + n[1] * ai[:, 2])
    lab = np.zeros(term.size, dtype=np.int32)
    t = term.ravel(order="F")
    for m in np.unique(t):
        w = np.where(alin == m)[0]
        assert len(w) == 1, "terminal maximum is not an atom position"
        lab[t == m] = w[0] + 1
    return lab.reshape(n, order="F")


@pytest.mark.parametrize("name", ["cubic48", "triclinic", "odd_dims", "tiny"])
def test_reference_scan_equals_own_trajectory_labels(name):
    """The sequential, order-dependent reference algorithm ends in the labelling where every point
    carries the terminal maximum of its own trajectory -- the property the
    parallel implementation computes."""
    c = cases.make_case(name)
    idg, nattr, xattr, stats = orc.bader_integrate(c["f"], c["x2c"], atoms=c["atoms"])
    term, st2 = orc.bader_canonical(c["f"], c["x2c"])
    lab = _atom_labels_from_terminals(term, c["atoms"], c["n"])
    assert nattr == len(c["atoms"])
    assert np.count_nonzero(lab != idg) == 0
    assert idg.min() >= 1 and idg.max() <= nattr


def test_single_atom_is_one_basin():
    c = cases.make_case("tiny")
    idg, nattr, _, _ = orc.bader_integrate(c["f"], c["x2c"], atoms=c["atoms"])
    assert nattr == 1 and (idg == 1).all()


def test_mirror_symmetric_dimer():
    """Two identical atoms at +-d on a cubic grid: the basins are mirror images and have equal volume."""
    n = (40, 24, 24)
    x2c = S.cell_x2c(10, 6, 6)
    at = np.array([[0.25, 0.5, 0.5], [0.75, 0.5, 0.5]])
    f = orc.promolecular(n, x2c, at, [3.0, 3.0], [1.7, 1.7], nimg=1)
    idg, nattr, _, _ = orc.bader_integrate(f, x2c, atoms=at)
    assert nattr == 2
    vol, _ = orc.integrate_bader(idg, [], nattr, S.omega(x2c))
    # the plane x = 0 and x = 0.5 hold ties; everything else must be mirror symmetric
    a = idg[1:20, :, :]
    b = idg[21:40, :, :][::-1, :, :]
    assert np.count_nonzero((a == 1) != (b == 2)) == 0
    assert abs(vol.sum() - S.omega(x2c)) < 1e-9


def test_noatoms_numbering_and_nnm_merge():
    c = cases.make_case("cubic48")
    idg_a, na, _, _ = orc.bader_integrate(c["f"], c["x2c"], atoms=c["atoms"])
    idg_n, nn_, xattr, _ = orc.bader_integrate(c["f"], c["x2c"], atoms=None, atexist=False)
    assert nn_ == na
    # same partition, different numbering
    pairs = set(zip(idg_a.ravel().tolist(), idg_n.ravel().tolist()))
    assert len(pairs) == na
    # attractor positions are grid nodes
    assert np.allclose(xattr * np.array(c["n"])[:, None], np.round(xattr * np.array(c["n"])[:, None]))


def test_integrate_bader_known_answer():
    rng = np.random.default_rng(0)
    n = (6, 5, 4)
    idg = rng.integers(1, 4, size=n).astype(np.int32)
    f = rng.random(n)
    vol, ps = orc.integrate_bader(idg, [f], 3, 2.5)
    for i in range(3):
        assert np.isclose(vol[i], np.count_nonzero(idg == i + 1) * 2.5 / idg.size, rtol=0, atol=1e-15)
        assert np.isclose(ps[i, 0], f[idg == i + 1].sum() * 2.5 / idg.size, rtol=1e-14)


def test_golden_fixture():
    """Labels / integrals committed under tests/golden (made by tests/golden/make_golden.py with this
    oracle): guards the oracle itself against silent changes."""
    g = json.load(open(GOLDEN))
    for name, ref in g["bader"].items():
        c = cases.make_case(name)
        idg, nattr, _, stats = orc.bader_integrate(c["f"], c["x2c"], atoms=c["atoms"])
        assert nattr == ref["nattr"]
        assert hashlib.sha256(np.ascontiguousarray(idg.ravel(order="F")).tobytes()).hexdigest() == ref["labels_sha256"]
        vol, ps = orc.integrate_bader(idg, [c["f"]], nattr, S.omega(c["x2c"]))
        assert np.allclose(ps[:, 0], ref["pop"], rtol=1e-12, atol=0)
        assert np.allclose(vol, ref["vol"], rtol=1e-13, atol=0)


def test_assign_attractors_helper_matches_oracle_logic():
    c = cases.make_case("cubic48")
    idg, nattr, xattr, _ = orc.bader_integrate(c["f"], c["x2c"], atoms=c["atoms"])
    # feed the atom nodes as "maxima" in a shuffled order
    n = np.array(c["n"])
    pm = (np.round(c["atoms"] * n).astype(int) % n) + 1
    perm = np.array([3, 1, 5, 0, 2, 4])
    mp, na, xa = H.assign_attractors(pm[perm], c["n"], c["x2c"], c["atoms"])
    assert na == nattr and (mp == perm + 1).all()
