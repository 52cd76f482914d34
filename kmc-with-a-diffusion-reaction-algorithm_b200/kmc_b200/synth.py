"""Synthetic OLIGOMERISED start states for measurements (bench.py extras, tools/): the membrane at its default density with most
ligands already bound in aligned complexes, which a run from a bond-free start only reaches after ~10^6 steps.

Complexes are assembled with the reference's own alignment geometry (main.cpp:1157-1228 ligand template and receptor snap,
786-798 cis snap, 1439-1501 re-seat of a bridged ligand), so the sweep finds them aligned (AreSame, 1e-8) and lying flat:
  type A  one ligand, receptors on 1-3 of its sites, some of them with a ligand-free cis partner          (2-7 molecules)
  type B  two ligands bridged by a cis pair of receptors (ligand - receptor ~ receptor - ligand) plus further receptors
The rest of the receptors and ligands are free. Host-side numpy; used for benchmarks only, never in a parity test."""
import numpy as np


def _rot(v, ang):
    c, s = np.cos(ang), np.sin(ang)
    return np.stack([v[..., 0] * c - v[..., 1] * s, v[..., 0] * s + v[..., 1] * c], axis=-1)


def oligomerised_state(p, seed=1, bound_fraction=0.6, two_ligand_share=0.4):
    """-> rec[na,6], lig[nb,24], rec_lig, rec_site, rec_cis (kmc_set_packed conventions) for the parameters p (n_replicas = 1)"""
    from scipy.spatial import cKDTree
    rng = np.random.default_rng(seed)
    na, nb = p.n_receptor, p.n_ligand
    Lx, Ly, Lz = p.box[0], p.box[1], p.box[2]
    rA, rB = p.rA, p.rB
    rs = rB * 2 / np.sqrt(3.0)
    half_b, half_c = p.bond_dist_cut / 2, p.cis_dist_cut / 2
    fRL1, fRL3, fRL2 = (half_b + rA) / rB, (half_b + 2 * rA) / rB, half_b / rB
    fC1, fC3, fC2 = (half_c + rA) / rA, half_c / rA, (half_c + 2 * rA) / rA
    fSeat = (half_b + rs + rB) / rA
    ghost = np.array([[0, 0], [0, rs], [-rB, -rB / np.sqrt(3)], [rB, -rB / np.sqrt(3)], [0, 0],
                      [0, rB * (2 / np.sqrt(3) + 1)], [-rB * (np.sqrt(3) / 2 + 1), -rB / np.sqrt(3) - rB / 2], [rB * (np.sqrt(3) / 2 + 1), -rB / np.sqrt(3) - rB / 2]])
    zA = 4 * rA                                  # receptor bead 3 (main.cpp:301)

    def seat(angle, centre):                     # flat ligand(s): [n,8,3]
        n = len(angle)
        out = np.zeros((n, 8, 3))
        out[:, :, :2] = _rot(ghost[None, :, :], angle[:, None]) + centre[:, None, :]
        out[:, :, 2] = zA
        out[:, 4, 2] = zA + rB
        return out

    def snap_rec(L, s):                          # receptor on site s of ligand(s) L: [n,6]
        idx = np.arange(len(L))
        bead, site = L[idx, 1 + s, :2], L[idx, 5 + s, :2]
        u = site - bead
        return np.concatenate([fRL1 * u + site, fRL2 * u + site, fRL3 * u + site], axis=1)

    def snap_cis(src):                           # cis partner rebuilt from the centre -> site-3 axis of src
        c, s3 = src[:, 0:2], src[:, 4:6]
        u = s3 - c
        return np.concatenate([fC1 * u + s3, fC2 * u + s3, fC3 * u + s3], axis=1)

    def reseat(r, s):                            # ligand seated on site s by receptor r (main.cpp:1439-1501)
        ax1 = ghost[1 + s]
        ax2 = r[:, 0:2] - r[:, 2:4]
        dot = ax1[:, 0] * ax2[:, 0] + ax1[:, 1] * ax2[:, 1]
        det = ax1[:, 0] * ax2[:, 1] - ax1[:, 1] * ax2[:, 0]
        angle = np.arctan2(-det, -dot) + p.pai
        centre = fSeat * (r[:, 2:4] - r[:, 0:2]) + r[:, 2:4]
        return seat(angle, centre)

    n_bound = int(bound_fraction * nb)
    nB2 = int(n_bound * two_ligand_share / 2)          # type B anchors (2 ligands each)
    nA1 = n_bound - 2 * nB2                            # type A anchors
    n_anchor = nA1 + nB2
    # anchors on a jittered square grid: complexes (<= ~400 A long) never touch each other
    g = int(np.ceil(np.sqrt(n_anchor)))
    pitch = min(Lx, Ly) / g
    if pitch < 900.0:
        raise ValueError("too dense for pre-assembled complexes: lower bound_fraction")
    cells = rng.permutation(g * g)[:n_anchor]
    jit = (pitch - 800.0) / 2
    ax = -Lx / 2 + (cells % g + 0.5) * pitch + rng.uniform(-jit, jit, n_anchor)
    ay = -Ly / 2 + (cells // g + 0.5) * pitch + rng.uniform(-jit, jit, n_anchor)
    anchor = np.stack([ax, ay], axis=1)
    typeB = np.zeros(n_anchor, bool); typeB[:nB2] = True
    rng.shuffle(typeB)

    ligs, recs = [], []            # lists of arrays; bonds as (receptor index, ligand index, site) / (receptor, receptor)
    rl, cis = [], []
    nlig = nrec = 0

    def add_ligs(L):
        nonlocal nlig
        ligs.append(L); idx = np.arange(nlig, nlig + len(L)); nlig += len(L); return idx

    def add_recs(R):
        nonlocal nrec
        recs.append(R); idx = np.arange(nrec, nrec + len(R)); nrec += len(R); return idx

    # first ligand of every anchor
    L1 = seat(rng.uniform(-np.pi, np.pi, n_anchor), anchor)
    iL1 = add_ligs(L1)
    occ = rng.random((n_anchor, 3)) < 0.6
    occ[:, 0] = True                                   # site 0 always carries a receptor (the bridge of type B uses it)
    first = {}
    for s in range(3):
        m = np.nonzero(occ[:, s])[0]
        R = snap_rec(L1[m], np.full(len(m), s))
        iR = add_recs(R)
        rl.append(np.stack([iR, iL1[m], np.full(len(m), s)], axis=1))
        if s == 0:
            first = dict(idx=iR, pose=R, anchors=m)
        else:                                          # ligand-free cis partner on some of them
            c = rng.random(len(m)) < 0.3
            P = snap_cis(R[c]); iP = add_recs(P)
            cis.append(np.stack([iR[c], iP], axis=1))
    # site-0 receptors: type B bridges to a second ligand, type A gets a ligand-free partner sometimes
    b = typeB[first["anchors"]]
    P = snap_cis(first["pose"][b]); iP = add_recs(P)
    cis.append(np.stack([first["idx"][b], iP], axis=1))
    L2 = reseat(P, np.zeros(len(P), int)); iL2 = add_ligs(L2)
    rl.append(np.stack([iP, iL2, np.zeros(len(P), int)], axis=1))
    for s in (1, 2):
        m = np.nonzero(rng.random(len(L2)) < 0.5)[0]
        R = snap_rec(L2[m], np.full(len(m), s)); iR = add_recs(R)
        rl.append(np.stack([iR, iL2[m], np.full(len(m), s)], axis=1))
    a = ~b & (rng.random(len(b)) < 0.3)
    P = snap_cis(first["pose"][a]); iP = add_recs(P)
    cis.append(np.stack([first["idx"][a], iP], axis=1))

    rec = np.concatenate(recs); lig = np.concatenate(ligs).reshape(-1, 24)
    if len(rec) > na or len(lig) > nb:
        raise ValueError("not enough receptors for this bound fraction (%d needed, %d available)" % (len(rec), na))
    # free receptors: uniform, >= 2 rA + 5 from every receptor, outside the footprint of every bound ligand
    tree_lig = cKDTree(lig[:, 0:2])
    placed = rec[:, 0:2]
    free = np.zeros((0, 2))
    need = na - len(rec)
    while need > 0:
        cand = np.stack([rng.uniform(-Lx / 2, Lx / 2, int(need * 1.3) + 16), rng.uniform(-Ly / 2, Ly / 2, int(need * 1.3) + 16)], axis=1)
        ok = tree_lig.query(cand, distance_upper_bound=rs + rB + rA + 10)[0] == np.inf
        cand = cand[ok]
        cand = cand[cKDTree(np.concatenate([placed, free])).query(cand, distance_upper_bound=2 * rA + 5)[0] == np.inf]
        keep = np.ones(len(cand), bool)                 # among the candidates themselves: drop the later one of each close pair
        for i, j in cKDTree(cand).query_pairs(2 * rA + 5):
            keep[max(i, j)] = False
        free = np.concatenate([free, cand[keep][:need]])
        need = na - len(rec) - len(free)
    psi = rng.uniform(-np.pi, np.pi, len(free))
    d = np.stack([np.cos(psi), np.sin(psi)], axis=1) * rA
    rec = np.concatenate([rec, np.concatenate([free, free + d, free - d], axis=1)])
    # free ligands: well above the receptors' bead stacks, >= 2 rs + 2 rB apart
    need = nb - len(lig)
    freeL = np.zeros((0, 3))
    while need > 0:
        cand = np.stack([rng.uniform(-Lx / 2, Lx / 2, int(need * 1.2) + 16), rng.uniform(-Ly / 2, Ly / 2, int(need * 1.2) + 16),
                         rng.uniform(8 * rA + rs + rB, Lz - rs - rB, int(need * 1.2) + 16)], axis=1)
        if len(freeL):
            cand = cand[cKDTree(freeL).query(cand, distance_upper_bound=2 * rs + 2 * rB + 5)[0] == np.inf]
        keep = np.ones(len(cand), bool)
        for i, j in cKDTree(cand).query_pairs(2 * rs + 2 * rB + 5):
            keep[max(i, j)] = False
        freeL = np.concatenate([freeL, cand[keep][:need]])
        need = nb - len(lig) - len(freeL)
    tpl = np.zeros((8, 3)); tpl[:, :2] = ghost; tpl[4, 2] = rB
    th, ph, ps = (rng.uniform(-np.pi, np.pi, len(freeL)) for _ in range(3))
    ct, st, cp, sp, cs, ss = np.cos(th), np.sin(th), np.cos(ph), np.sin(ph), np.cos(ps), np.sin(ps)
    T = np.stack([np.stack([cs * cp - ct * sp * ss, -ss * cp - ct * sp * cs, st * sp], axis=1),
                  np.stack([cs * sp + ct * cp * ss, -ss * sp + ct * cp * cs, -st * cp], axis=1),
                  np.stack([ss * st, cs * st, ct], axis=1)], axis=1)                    # [n,3,3]
    FL = np.einsum("nij,qj->nqi", T, tpl) + freeL[:, None, :]
    lig = np.concatenate([lig, FL.reshape(-1, 24)])
    # bond table, then cell-major renumbering (neighbours in space are neighbours in memory, like kmc_init_random(sort_cells=1))
    rec_lig = np.full(na, -1, np.int32); rec_site = np.zeros(na, np.int32); rec_cis = np.full(na, -1, np.int32)
    RL = np.concatenate(rl); CI = np.concatenate(cis)
    rec_lig[RL[:, 0]] = RL[:, 1]; rec_site[RL[:, 0]] = RL[:, 2] + 2
    rec_cis[CI[:, 0]] = CI[:, 1]; rec_cis[CI[:, 1]] = CI[:, 0]
    edge = 256.0
    ka = np.floor((rec[:, 1] + Ly / 2) / edge).astype(np.int64) * (1 << 20) + np.floor((rec[:, 0] + Lx / 2) / edge).astype(np.int64)
    kb = np.floor((lig[:, 1] + Ly / 2) / edge).astype(np.int64) * (1 << 20) + np.floor((lig[:, 0] + Lx / 2) / edge).astype(np.int64)
    oa, ob = np.argsort(ka, kind="stable"), np.argsort(kb, kind="stable")
    inv_a = np.empty(na, np.int32); inv_a[oa] = np.arange(na, dtype=np.int32)
    inv_b = np.empty(nb, np.int32); inv_b[ob] = np.arange(nb, dtype=np.int32)
    rec, lig = rec[oa], lig[ob]
    rl2, rs2, rc2 = rec_lig[oa], rec_site[oa], rec_cis[oa]
    rl2 = np.where(rl2 >= 0, inv_b[np.maximum(rl2, 0)], -1).astype(np.int32)
    rc2 = np.where(rc2 >= 0, inv_a[np.maximum(rc2, 0)], -1).astype(np.int32)
    return np.ascontiguousarray(rec), np.ascontiguousarray(lig), rl2, rs2.astype(np.int32), rc2
