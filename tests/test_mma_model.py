"""CPU model of the tensor-core ED projection (msc-project_b200/csrc/ed_mma.cu) against the CPU oracle.

The kernel maps the two passes of EDMD::calculate_ED_Forces (edmd.cpp:106-127) onto mma.sync.m8n8k4.f64 with eight
tetrads of one parameter set as the N dimension.  This test replays the kernel's index arithmetic lane by lane in
numpy — which atom / eigenvector / tetrad every lane feeds into the A and B fragments, where its D fragment goes, the
padded shared-memory row stride, the partial group — with the fragment layouts of the PTX ISA, and checks the result
against the oracle's ED pass on the same tetrads.  It pins the decomposition (and documents it) without a GPU; the
GPU tests check the kernel itself (tests/test_gpu_edge.py).
"""
import numpy as np

S = 256                      # atom stride of the state planes
SP = 3 * S + 4               # row stride of the staged eigenvectors (ed_mma.cu: 2 SP = 8 mod 32 words)
ROWS = 12


def dmma(acc, a, b):
    """D = A.B + C for one warp.  acc: [32][2]; a, b: [32].  Lane l = 4 g + r holds A[g][r], B[r][g], C/D[g][2r], C/D[g][2r+1]."""
    lanes = np.arange(32)
    g, r = lanes >> 2, lanes & 3
    A = np.zeros((8, 4)); B = np.zeros((4, 8)); C = np.zeros((8, 8))
    A[g, r] = a; B[r, g] = b
    C[g, 2 * r] = acc[:, 0]; C[g, 2 * r + 1] = acc[:, 1]
    D = A @ B + C
    acc[:, 0] = D[g, 2 * r]; acc[:, 1] = D[g, 2 * r + 1]


def bank_conflict_free(addr_of_lane):
    """64-bit shared-memory loads are served per half-warp: 16 lanes must touch 16 different bank pairs."""
    for half in (0, 1):
        banks = {(2 * addr_of_lane(l)) % 32 for l in range(16 * half, 16 * half + 16)}
        if len(banks) != 16:
            return False
    return True


def test_fragment_loads_are_conflict_free():
    for s in range(8):
        for c in range(3):
            assert bank_conflict_free(lambda l: (l >> 2) * SP + c * S + 4 * s + (l & 3))               # pass 1, rows 0-7
            # rows 8-11: only lanes 0-15 (g < 4) load; the other half-warp feeds zero rows
            assert len({(2 * ((8 + (l >> 2)) * SP + c * S + 4 * s + (l & 3))) % 32 for l in range(16)}) == 16
    for q in range(3):
        for c in range(3):
            for a0 in (0, 8, 248):
                assert bank_conflict_free(lambda l: (4 * q + (l & 3)) * SP + c * S + a0 + (l >> 2))    # pass 2
    # the unpadded stride 3 S would put all rows of a fragment on the same banks
    assert not bank_conflict_free(lambda l: (l >> 2) * 3 * S + 4 + (l & 3))


def model_group(ps, crd, sup, scaled, ng):
    """One group of `ng` <= 8 tetrads sharing parameter set `ps`.  crd: [8][3 S] planes, sup: [8][24] (R, shift,
    centroid of avg, centroid of x).  Returns new coordinate planes, force planes, ED energies."""
    na, nev = ps.num_atoms, ps.num_evecs
    sEV = np.zeros(ROWS * SP)
    sAV = np.zeros(3 * S)
    for c in range(3):
        sAV[c * S:c * S + na] = ps.avg[c::3]
        for j in range(nev):
            sEV[j * SP + c * S:j * SP + c * S + na] = ps.evecs[j, c::3]
    lanes = np.arange(32)
    g, r = lanes >> 2, lanes & 3
    colv = g < ng
    # ---- pass 1: K = atoms, four per MMA; rows 0-7 and 8-11 (+ 4 zero rows) of E
    acc = [np.zeros((32, 2)), np.zeros((32, 2))]
    for s in range(S // 4):
        a = 4 * s + r
        ok = colv & (a < na)
        d = np.zeros((32, 3))
        for l in np.nonzero(ok)[0]:
            R = sup[g[l], :9].reshape(3, 3); ca = sup[g[l], 12:15]; cx = sup[g[l], 15:18]
            xc = crd[g[l], [a[l], S + a[l], 2 * S + a[l]]] - cx
            ac = sAV[[a[l], S + a[l], 2 * S + a[l]]] - ca
            d[l] = R @ xc - ac
        for c in range(3):
            a_lo = sEV[g * SP + c * S + a]
            a_hi = np.where(g < 4, sEV[np.minimum(8 + g, ROWS - 1) * SP + c * S + a], 0.0)
            dmma(acc[0], a_lo, d[:, c]); dmma(acc[1], a_hi, d[:, c])
    # ---- D fragments -> scratch [12][8] -> B fragments of pass 2
    xp = np.full(ROWS * 8, np.nan)
    for l in range(32):
        xp[g[l] * 8 + 2 * r[l]:g[l] * 8 + 2 * r[l] + 2] = acc[0][l]
        if g[l] < 4:
            xp[(8 + g[l]) * 8 + 2 * r[l]:(8 + g[l]) * 8 + 2 * r[l] + 2] = acc[1][l]
    assert not np.isnan(xp).any()
    P = np.zeros((32, 3)); W = np.zeros((32, 3)); term = np.zeros((32, 3))
    for q in range(3):
        j = r + 4 * q
        el = np.where(j < nev, ps.evals[np.minimum(j, nev - 1)], 1.0)
        P[:, q] = xp[j * 8 + g]
        W[:, q] = -(P[:, q] * scaled / el)
        term[:, q] = P[:, q] * P[:, q] / el
    en = np.zeros(8)
    for l in range(32):
        if r[l] == 0 and colv[l]:
            e = 0.0
            for j in range(ROWS):
                if j < nev:
                    e += term[(l & ~3) | (j & 3), j >> 2]
            en[g[l]] = e * (0.5 * scaled)
    # ---- pass 2: 8-atom M tiles, K = eigenvectors
    X = np.zeros((8, 3 * S)); F = np.zeros((8, 3 * S))
    for a0 in range(0, S, 8):
        a = a0 + g
        sv = [np.zeros((32, 2)) for _ in range(3)]; fv = [np.zeros((32, 2)) for _ in range(3)]
        for c in range(3):
            sv[c][:, 0] = sv[c][:, 1] = sAV[c * S + a]
            for q in range(3):
                ev = sEV[(4 * q + r) * SP + c * S + a]
                dmma(sv[c], ev, P[:, q]); dmma(fv[c], ev, W[:, q])
        for l in range(32):
            if a[l] >= na:
                continue
            for u in range(2):
                t = 2 * r[l] + u
                if t >= ng:
                    continue
                R = sup[t, :9].reshape(3, 3); sh = sup[t, 9:12]
                s3 = np.array([sv[0][l, u], sv[1][l, u], sv[2][l, u]]) - sh
                f3 = np.array([fv[0][l, u], fv[1][l, u], fv[2][l, u]])
                X[t, [a[l], S + a[l], 2 * S + a[l]]] = R.T @ s3
                F[t, [a[l], S + a[l], 2 * S + a[l]]] = R.T @ f3
    return X, F, en


def test_mma_decomposition_matches_the_oracle(edmd, po):
    s = edmd.make_ring(120, kind="coil", laps=3, lap_gap=9.0)          # 10 parameter sets x 12 tetrads
    o = po.Oracle(s, "port")
    scaled = float(s.params[4])
    off = s.off3
    for pid, ng in ((0, 8), (3, 5)):                                  # a full group and a partial one
        ps = s.psets[pid]
        na = ps.num_atoms
        tets = np.nonzero(s.param_id == pid)[0][:ng]
        crd = np.zeros((8, 3 * S)); sup = np.zeros((8, 24))
        for k, t in enumerate(tets):
            x = s.crd[off[t]:off[t] + 3 * na].reshape(na, 3)
            avg = ps.avg.reshape(na, 3)
            R, _ = o.qcp(avg, x)                                       # rotates x onto avg (edmd.cpp:88)
            for c in range(3):
                crd[k, c * S:c * S + na] = x[:, c]
            sup[k, :9] = R.reshape(-1)
            sup[k, 9:12] = (avg - x @ R.T).mean(axis=0)                # offset vector, edmd.cpp:97-102
            sup[k, 12:15] = avg.mean(axis=0); sup[k, 15:18] = x.mean(axis=0)
        X, F, en = model_group(ps, crd, sup, scaled, ng)
        for k, t in enumerate(tets):
            o.ed_forces(int(t))
        fo = o.get_forces(); co, _ = o.get_state()
        for k, t in enumerate(tets):
            f_ref = fo["ed"][off[t]:off[t] + 3 * na].reshape(na, 3)
            x_ref = co[off[t]:off[t] + 3 * na].reshape(na, 3)
            f_mod = np.stack([F[k, c * S:c * S + na] for c in range(3)], axis=1)
            x_mod = np.stack([X[k, c * S:c * S + na] for c in range(3)], axis=1)
            assert np.abs(f_mod - f_ref).max() <= 1e-11 * np.abs(f_ref).max()
            assert np.abs(x_mod - x_ref).max() <= 1e-12 * np.abs(x_ref).max()
            assert abs(en[k] - fo["ed_energy"][t]) <= 1e-11 * abs(fo["ed_energy"][t])
        assert not X[ng:].any() and not F[ng:].any()                  # columns beyond the group are never stored
