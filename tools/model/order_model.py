"""Block-liveness model of the generation-4/5 tile kernel for different in-cell orders and prefilters.

Design-time numpy model (CPU only); not used by the product, the tests or the bench.

For sampled tiles of the bench geometries it counts, per (layer of 32 i, quad of 4 j) block:
  pass   the box prefilter lets it through (per layer)
  tested exact tests run under the "every layer of a live quad" policy of MODE 1
  live   at least one of its 128 pairs is in range
and the lane efficiency of the live blocks, for in-cell orders Morton-64 / Hilbert-64 / Hilbert-512 and for
1 / 2 / 4 boxes per layer.
"""
import sys
import numpy as np

rng = np.random.default_rng(1)


def hilbert_index(x, y, z, bits):
    """3-D Hilbert index (Skilling's transform), vectorised over numpy int arrays."""
    X = [x.astype(np.int64).copy(), y.astype(np.int64).copy(), z.astype(np.int64).copy()]
    n = 3
    M = 1 << (bits - 1)
    Q = M
    while Q > 1:
        P = Q - 1
        for i in range(n):
            m = (X[i] & Q) != 0
            # invert
            X[0] = np.where(m, X[0] ^ P, X[0])
            # exchange
            t = np.where(m, 0, (X[0] ^ X[i]) & P)
            X[0] ^= t
            X[i] ^= t
        Q >>= 1
    for i in range(1, n):
        X[i] ^= X[i - 1]
    t = np.zeros_like(X[0])
    Q = M
    while Q > 1:
        t = np.where((X[n - 1] & Q) != 0, t ^ (Q - 1), t)
        Q >>= 1
    for i in range(n):
        X[i] ^= t
    h = np.zeros_like(X[0])
    for b in range(bits - 1, -1, -1):
        for i in range(n):
            h = (h << 1) | ((X[i] >> b) & 1)
    return h


def morton_index(x, y, z, bits):
    h = np.zeros_like(x, dtype=np.int64)
    for b in range(bits - 1, -1, -1):
        h = (h << 3) | (((x >> b) & 1) << 2) | (((y >> b) & 1) << 1) | ((z >> b) & 1)
    return h


def run(N, W, R, radio=None, ratio=0.0, T=6, ntiles=24, name='', order='morton', sub_bits=2, nbox=1, homog=True,
        seed=1):
    rng = np.random.default_rng(seed)
    a = np.ones(T) if radio is None else 1 + np.array(radio) * ratio
    h = 0.5 * R * a
    uniform = radio is None
    Rmax = 2 * h.max()
    nc = int(W // (Rmax * (1 + 1e-5)))
    pos = rng.random((N, 3)) * W
    typ = rng.integers(0, T, N)
    S = 1 << sub_bits
    c = np.minimum((pos * nc / W).astype(int), nc - 1)
    sub = np.minimum((pos * S * nc / W).astype(int), S * nc - 1) - S * c
    if order == 'morton':
        code = morton_index(sub[:, 0], sub[:, 1], sub[:, 2], sub_bits)
    else:
        code = hilbert_index(sub[:, 0], sub[:, 1], sub[:, 2], sub_bits)
    cell = (c[:, 0] * nc + c[:, 1]) * nc + c[:, 2]
    key = cell * (S ** 3) + code
    o = np.argsort(key, kind='stable')
    pos, typ, cell, code = pos[o], typ[o], cell[o], code[o]
    start = np.searchsorted(cell, np.arange(nc ** 3 + 1))
    # j copy: (row, type, zcell, code) when radii differ per type
    if not uniform and homog:
        row = cell // nc
        cz = cell % nc
        jkey = ((row * T + typ) * nc + cz)
        oj = np.argsort(jkey, kind='stable')
        jpos, jtyp, jk = pos[oj], typ[oj], jkey[oj]
        jstart = np.searchsorted(jk, np.arange(nc * nc * T * nc + 1))
    tot = dict(blocks=0, passed=0, quads=0, livequads=0, tested=0, live=0, acc=0, chunks=0)
    cells = rng.integers(0, nc ** 3, ntiles)
    for ce in cells:
        cz = ce % nc
        cy = (ce // nc) % nc
        cx = ce // (nc * nc)
        i0 = start[ce]
        n = start[ce + 1] - i0
        if n == 0:
            continue
        k = rng.integers(0, (n + 127) // 128)
        ib = i0 + k * 128
        ni = min(n - k * 128, 128)
        nl = (ni + 31) // 32
        layers = [(pos[ib + 32 * l: ib + min(32 * l + 32, ni)], typ[ib + 32 * l: ib + min(32 * l + 32, ni)]) for l in range(nl)]
        for dx in (-1, 0, 1):
            for dy in (-1, 0, 1):
                x = (cx + dx) % nc
                y = (cy + dy) % nc
                sx = (-W if cx + dx < 0 else (W if cx + dx >= nc else 0))
                sy = (-W if cy + dy < 0 else (W if cy + dy >= nc else 0))
                segs = [(max(cz - 1, 0), min(cz + 1, nc - 1), 0.0)]
                if cz == 0:
                    segs.append((nc - 1, nc - 1, -W))
                if cz == nc - 1:
                    segs.append((0, 0, W))
                for z0, z1, sz in segs:
                    rw = x * nc + y
                    subruns = []
                    if not uniform and homog:
                        for t in range(T):
                            b = (rw * T + t) * nc
                            subruns.append((jpos, jtyp, jstart[b + z0], jstart[b + z1 + 1]))
                    else:
                        subruns.append((pos, typ, start[rw * nc + z0], start[rw * nc + z1 + 1]))
                    for P, Ty, j0, j1 in subruns:
                        if j1 <= j0:
                            continue
                        pj = P[j0:j1] + np.array([sx, sy, sz])
                        tj = Ty[j0:j1]
                        nj = j1 - j0
                        nq = (nj + 3) // 4
                        tot['chunks'] += (nj + 127) // 128
                        # quad boxes
                        padn = nq * 4 - nj
                        pjp = np.concatenate([pj, np.repeat(pj[-1:], padn, 0)]) if padn else pj
                        tjp = np.concatenate([tj, np.repeat(tj[-1:], padn)]) if padn else tj
                        qlo = pjp.reshape(nq, 4, 3).min(1)
                        qhi = pjp.reshape(nq, 4, 3).max(1)
                        qH = h[tjp].reshape(nq, 4).max(1)
                        anyq = np.zeros(nq, bool)
                        passes = []
                        lives = []
                        for (pi, ti) in layers:
                            m = len(pi)
                            # prefilter with nbox boxes per layer
                            pas = np.zeros(nq, bool)
                            for bx in range(nbox):
                                s0 = bx * 32 // nbox
                                s1 = min((bx + 1) * 32 // nbox, m)
                                if s1 <= s0:
                                    continue
                                lo = pi[s0:s1].min(0)
                                hi = pi[s0:s1].max(0)
                                Hi = h[ti[s0:s1]].max()
                                gap = np.maximum(0, np.maximum(qlo - hi, lo - qhi))
                                pas |= (gap * gap).sum(1) < (Hi + qH) ** 2
                            d = pjp[None, :, :] - pi[:, None, :]
                            d2 = (d * d).sum(-1)
                            cut = (h[ti][:, None] + h[tjp][None, :]) ** 2
                            ok = d2 < cut
                            if padn:
                                ok[:, nj:] = False
                            lv = ok.reshape(m, nq, 4).any(2).any(0)
                            passes.append(pas)
                            lives.append(lv)
                            tot['acc'] += ok.sum()
                            tot['blocks'] += nq
                            tot['passed'] += pas.sum()
                            anyq |= pas
                        tot['quads'] += nq
                        tot['livequads'] += anyq.sum()
                        for pas, lv in zip(passes, lives):
                            tot['tested'] += anyq.sum()
                            tot['live'] += (lv & anyq).sum()
                            assert not (lv & ~pas).any()
    b = tot['blocks']
    print(f"{name:10s} order={order}{S**3:<4d} nbox={nbox}: pass {tot['passed']/b:.3f} tested(all layers of live quad) {tot['tested']/b:.3f} "
          f"live {tot['live']/b:.3f}  lane-eff {tot['acc']/(128.0*tot['live']):.3f}  livequads {tot['livequads']/tot['quads']:.3f}")
    return tot


if __name__ == '__main__':
    which = sys.argv[1] if len(sys.argv) > 1 else 'eater'
    cfgs = {
        'pulser': dict(N=1000000, W=8000., R=386.),
        'eater': dict(N=1000000, W=8000., R=397., radio=[1, .5, 0, 0, -.5, 1], ratio=0.5),
        'settings2M': dict(N=2000000, W=8000., R=285., T=8),
    }
    for order, bits in (('morton', 2), ('hilbert', 2), ('hilbert', 3), ('morton', 3)):
        for nbox in (1, 2, 4):
            run(name=which, order=order, sub_bits=bits, nbox=nbox, **cfgs[which])
