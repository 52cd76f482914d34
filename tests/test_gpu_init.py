"""GPU: the initial-configuration generator (csrc/kmc_init.cu, replaces main.cpp:273-456).

(a) It IS the reference's sequential insertion: molecule i takes the first of its keyed candidate positions that clashes with no
    molecule placed before it (main.cpp:284-296, 354-383). A plain sequential numpy restatement with the same keyed draws must give
    the same configuration bit for bit (dense system: many clashes and retries).
(b) The exclusion radii at t = 0 (main.cpp:293, 368, 380) on a large membrane, by KD-tree.
(c) The distributions the reference draws from: positions uniform in the box, receptor psi ~ U(-pai, pai) (main.cpp:330), ligand
    Euler angles ~ U(-pai, pai) each (main.cpp:422-424), rigid bodies with the template edge lengths."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import kmc_b200

M0, M1, W0, W1 = 0xD2511F53, 0xCD9E8D57, 0x9E3779B9, 0xBB67AE85


def philox(c, k0, k1):
    """Philox4x32-10 on python ints (c = 4 words)"""
    c = list(c)
    for _ in range(10):
        p0, p1 = M0 * c[0], M1 * c[2]
        c = [((p1 >> 32) ^ c[1] ^ k0) & 0xffffffff, p1 & 0xffffffff, ((p0 >> 32) ^ c[3] ^ k1) & 0xffffffff, p0 & 0xffffffff]
        k0 = (k0 + W0) & 0xffffffff; k1 = (k1 + W1) & 0xffffffff
    return c


def uniform(seed, mol, partner, slot):
    """the library's keyed_uniform at step 0 (csrc/kmc_philox.cuh): two draws per block"""
    c = philox([mol, partner, 0, slot & ~1], seed & 0xffffffff, seed >> 32)
    hi, lo = (c[3], c[2]) if slot & 1 else (c[1], c[0])
    return float(((hi << 32) | lo) >> 11) * (1.0 / 9007199254740992.0)


def test_generator_equals_sequential_insertion():
    na, nb, box, seed = 1500, 500, (7500.0, 7500.0, 400.0), 12345
    p = kmc_b200.default_params(box=box, n_receptor=na, n_ligand=nb)
    k = kmc_b200.Kmc(p)
    k.init_random(seed=seed, sort_cells=False)
    rec, lig, _, _, _ = k.get_packed()
    rA, rB = p.rA, p.rB
    rs = rB * 2 / np.sqrt(3.0)
    exRR, exRL, exLL = 2 * rA, rA + rs + rB, 2 * rs + 2 * rB
    A = np.zeros((na, 2)); B = np.zeros((nb, 3))
    retries = 0
    for a in range(na):
        t = 0
        while True:
            x = uniform(seed, a + 1, t, 32) * box[0] - box[0] / 2; y = uniform(seed, a + 1, t, 33) * box[1] - box[1] / 2
            if a == 0 or not (np.sqrt((A[:a, 0] - x) ** 2 + (A[:a, 1] - y) ** 2) <= exRR).any():
                break
            t += 1; retries += 1
        A[a] = (x, y)
    beads_z = np.array([0.0, 2 * rA, 4 * rA, 6 * rA])
    for b in range(nb):
        t = 0
        while True:
            x = uniform(seed, na + b + 1, t, 32) * box[0] - box[0] / 2; y = uniform(seed, na + b + 1, t, 33) * box[1] - box[1] / 2
            z = uniform(seed, na + b + 1, t, 34) * box[2]
            d2 = (A[:, 0] - x) ** 2 + (A[:, 1] - y) ** 2
            clash = (np.sqrt(d2[:, None] + (z - beads_z[None, :]) ** 2) <= exRL).any()
            clash = clash or (b > 0 and (np.sqrt((B[:b, 0] - x) ** 2 + (B[:b, 1] - y) ** 2 + (B[:b, 2] - z) ** 2) <= exLL).any())
            if not clash:
                break
            t += 1; retries += 1
        B[b] = (x, y, z)
    assert retries > 50                                  # the dense box really exercises the retry chains
    assert np.array_equal(rec[:, 0:2], A), "receptor centres differ from the sequential insertion"
    assert np.array_equal(lig[:, 0:3], B), "ligand centres differ from the sequential insertion"


def test_exclusion_radii_and_distributions_large_membrane():
    from scipy import stats
    from scipy.spatial import cKDTree
    M = 400000
    na, nb = 3 * M // 4, M // 4
    box = kmc_b200.scaled_box(M)
    p = kmc_b200.default_params(box=box, n_receptor=na, n_ligand=nb)
    k = kmc_b200.Kmc(p)
    k.init_random(seed=7, sort_cells=True)
    rec, lig, rl, rs_, rc = k.get_packed()
    assert (rl < 0).all() and (rc < 0).all()
    rA, rB = p.rA, p.rB
    rs = rB * 2 / np.sqrt(3.0)
    c = rec[:, 0:2]
    L = lig.reshape(nb, 8, 3)
    # (b) exclusions: main.cpp:293 (receptor centres > 2 rA), 368 (ligand centre > rA + rs + rB from every receptor bead), 380 (ligand centres > 2 rs + 2 rB)
    assert len(cKDTree(c).query_pairs(2 * rA)) == 0
    rec_beads = np.concatenate([np.column_stack([c, np.full(na, z)]) for z in (0.0, 2 * rA, 4 * rA, 6 * rA)])
    assert cKDTree(rec_beads).query_ball_point(L[:, 0], rA + rs + rB, return_length=True).sum() == 0
    assert len(cKDTree(L[:, 0]).query_pairs(2 * rs + 2 * rB)) == 0
    # (c) positions uniform in the box (chi-square on a 16x16 histogram; z of the ligands uniform in [0, Lz] up to the receptor exclusion)
    for pts in (c, L[:, 0, :2]):
        hist, _, _ = np.histogram2d(pts[:, 0], pts[:, 1], bins=16, range=[[-box[0] / 2, box[0] / 2], [-box[1] / 2, box[1] / 2]])
        assert stats.chisquare(hist.ravel()).pvalue > 1e-4
    assert L[:, 0, 2].min() >= 0 and L[:, 0, 2].max() < box[2]
    assert stats.kstest(L[L[:, 0, 2] > 250.0, 0, 2], "uniform", args=(250.0, box[2] - 250.0)).pvalue > 1e-4
    # receptor: sites at rA from the centre, opposite each other; psi uniform in (-pai, pai)
    d2, d3 = rec[:, 2:4] - c, rec[:, 4:6] - c
    assert np.allclose(np.linalg.norm(d2, axis=1), rA, rtol=0, atol=1e-9) and np.allclose(d2, -d3, rtol=0, atol=1e-9)
    psi = np.arctan2(d2[:, 1], d2[:, 0])
    assert stats.kstest(psi, "uniform", args=(-np.pi, 2 * np.pi)).pvalue > 1e-4
    # ligand: template edge lengths; theta ~ U(-pai, pai) makes the polar angle of the normal (centre -> marker) uniform in [0, pi]
    for q in (1, 2, 3):
        assert np.allclose(np.linalg.norm(L[:, q] - L[:, 0], axis=1), rs, rtol=0, atol=1e-9)
        assert np.allclose(np.linalg.norm(L[:, 4 + q] - L[:, 0], axis=1), rs + rB, rtol=0, atol=1e-9)
    nrm = (L[:, 4] - L[:, 0]) / rB
    assert np.allclose(np.linalg.norm(nrm, axis=1), 1.0, rtol=0, atol=1e-12)
    assert stats.kstest(np.arccos(np.clip(nrm[:, 2], -1, 1)), "uniform", args=(0.0, np.pi)).pvalue > 1e-4
    azim = np.arctan2(nrm[:, 1], nrm[:, 0])
    assert stats.kstest(azim, "uniform", args=(-np.pi, 2 * np.pi)).pvalue > 1e-4
    # cell-major numbering: consecutive receptors are neighbours in space (memory locality of the gathers)
    assert np.median(np.linalg.norm(np.diff(c, axis=0), axis=1)) < 2000.0
    # deterministic: the same seed gives the same membrane, another seed a different one
    k2 = kmc_b200.Kmc(p); k2.init_random(seed=7, sort_cells=True)
    assert np.array_equal(k2.get_packed()[0], rec)
    k2.init_random(seed=8, sort_cells=True)
    assert not np.array_equal(k2.get_packed()[0], rec)


def test_one_million_molecules_in_milliseconds():
    """the whole 1e7-molecule membrane in well under a second (the host generator of round 1 took 2 s); here 2e6 to keep the test light"""
    import time
    M = 2000000
    k = kmc_b200.Kmc(kmc_b200.default_params(box=kmc_b200.scaled_box(M), n_receptor=3 * M // 4, n_ligand=M // 4))
    k.init_random(seed=1, sort_cells=True)          # warm
    t0 = time.perf_counter(); k.init_random(seed=2, sort_cells=True); dt = time.perf_counter() - t0
    print("GPU initial configuration: %d molecules in %.1f ms" % (M, 1e3 * dt))
    assert dt < 0.25
