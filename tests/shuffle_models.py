"""Numpy models of the tiled forms of partition_shuffle (crates/bvh/src/blas.rs:168-182) used or planned by the grid tier
of the CUDA BLAS build.  Test infrastructure only (imported by tests/test_oracle.py); nothing here is on a product path.

`two_phase`  the form k_t1_coop runs today: PA builds the part of the rank->position table that can be looked up, a grid
             barrier, PB gathers from it and scatters.
`one_phase`  groundwork for DESIGN.md section 9 item 2 (one barrier per shuffle): no global table.  The only elements that
             change place are the front R's (before the boundary element f) and the back L's (after f):
                 m-th front R (m >= 1)  ->  position of the (m-1)-th L from the back, minus one   (the 0-th goes to n-1)
                 k-th L from the back   ->  position of the k-th front R
             so every move needs one rank-select in a tile on the other side of f.  A tile can answer those for itself
             from the per-tile L counts (known before the shuffle starts) plus the flags of the partner tile, which it
             loads.  To keep the number of partner tiles per tile bounded whatever the plane's left fraction is, a move
             that involves front tile A and back tile B is executed by A if A holds no more R's than B holds L's, else
             by B: both sides evaluate the same rule from the per-tile counts alone, and a tile that drives a partner
             spends at most its own element count of that partner's ranks, so it drives ~2 partners per direction.
             The model executes the shuffle tile by tile under exactly that information budget and reports how many
             partner tiles every tile had to load.
`axis_with_frozen_prefix`  groundwork for DESIGN.md section 9 item 0: the buffer plan that lets the ping-pong tiers shuffle
             only the suffix behind the previous pivot on the 2nd..7th plane of an axis."""
import numpy as np


def axis_with_frozen_prefix(ids, k_axis, shuffle):
    """Buffer plan for DESIGN.md section 9 item 0 in the ping-pong tiers (grid, block, warp): the 7 planes of one axis with
    only the suffix behind the previous pivot taking part in the ping-pong.  `ids`: the order before the axis' first plane;
    `k_axis[id]`: the primitive's plane count on this axis; `shuffle(ids, flags_by_id) -> (pivot, ids)`: the sequential
    partition_shuffle.  Shuffle b reads the suffix [pivot_{b-1}, n) from the current buffer and writes it to the other one,
    so segment [pivot_{b-1}, pivot_b) stays frozen in the buffer shuffle b wrote; one merge pass at the end copies the
    segments that are stale in the final buffer.  Returns (pivots, merged order, slots copied by the merge)."""
    n = len(ids)
    buf = [np.array(ids, dtype=np.uint32), np.full(n, 0xFFFFFFFF, dtype=np.uint32)]  # the other buffer holds garbage
    cur, off = 0, 0
    pivots, owner = [], np.zeros(n, dtype=np.int64)  # owner[j]: buffer that holds the valid value of slot j
    for b in range(1, 8):
        fl = (np.asarray(k_axis) < b).astype(np.uint8)
        piv, out = shuffle(buf[cur][off:], fl) if n > off else (0, buf[cur][off:])
        buf[cur ^ 1][off:] = out
        owner[off:] = cur ^ 1
        cur ^= 1
        off += piv
        pivots.append(off)
    stale = owner != cur
    buf[cur][stale] = buf[cur ^ 1][stale]  # the merge pass (before an axis change, the bins, or the final shuffle)
    return pivots, buf[cur], int(stale.sum())


def boundary(L, n, nL):
    """f and pivot from nL and the flags at nL-1, nL, nL+1 (the closed form the kernels use)."""
    at = lambda j: int(L[j]) if 0 <= j < n else 0
    l0, l1, l2 = (at(nL - 1) if nL else 0), at(nL), at(nL + 1)
    if nL >= 1 and not (nL + 1 <= n and l0 + l1 <= 1):
        f, lf = nL - 1, l0
    elif not (nL + 2 <= n and l1 + l2 == 0):
        f, lf = nL, l1
    else:
        f, lf = nL + 1, l2
    return f, nL - lf


def two_phase(flags, tile):
    """dest[] and pivot as PA + PB compute them, tile by tile (the table is global, as in the kernel)."""
    n = len(flags)
    L = np.asarray(flags, dtype=bool)
    nt = (n + tile - 1) // tile
    cnt = np.array([int(L[t * tile:(t + 1) * tile].sum()) for t in range(nt)])
    pre = np.concatenate([[0], np.cumsum(cnt)])
    nL = int(pre[-1])
    f, pivot = boundary(L, n, nL)
    table = np.full(n, -10**9, dtype=np.int64)
    for t in range(nt):  # PA
        lf = pre[t]
        for j in range(t * tile, min(n, (t + 1) * tile)):
            if L[j]:
                if j >= nL:
                    table[n - 1 - (nL - lf - 1)] = j
                lf += 1
            elif j <= nL:
                table[j - lf] = j
    dest = np.zeros(n, dtype=np.int64)
    for t in range(nt):  # PB
        lf = pre[t]
        for j in range(t * tile, min(n, (t + 1) * tile)):
            rf = j - lf
            if j < f:
                dest[j] = j if L[j] else (n - 1 if rf == 0 else table[n - rf] - 1)
            elif j == f:
                dest[j] = pivot
            else:
                dest[j] = table[nL - lf - 1] if L[j] else j - 1
            lf += int(L[j])
    return dest, pivot


def one_phase(flags, tile):
    """dest[] (-1 where nobody moved the element: a bug), pivot, and the number of partner tiles each tile loaded."""
    n = len(flags)
    L = np.asarray(flags, dtype=bool)
    nt = (n + tile - 1) // tile
    size = np.array([min(n, (t + 1) * tile) - t * tile for t in range(nt)])
    cnt = np.array([int(L[t * tile:(t + 1) * tile].sum()) for t in range(nt)])  # published before the shuffle starts
    pre = np.concatenate([[0], np.cumsum(cnt)])                                # L's before tile t
    nL = int(pre[-1])
    f, pivot = boundary(L, n, nL)  # three flag reads, every tile does them for itself
    dest = np.full(n, -1, dtype=np.int64)
    moved = np.zeros(n, dtype=np.int64)
    loads = np.zeros(nt, dtype=np.int64)

    def put(j, d):
        dest[j] = d
        moved[j] += 1

    # what a tile may know about another tile without loading it: its L count, hence the rank ranges it holds
    def r_front_range(t):  # ranks (RF) of the R's of tile t, as [lo, hi): exact for tiles that end before f
        lo = t * tile - pre[t]
        return lo, lo + (size[t] - cnt[t])

    def l_back_range(t):  # ranks from the back (LB) of the L's of tile t, as [lo, hi)
        return nL - pre[t + 1], nL - pre[t]

    def a_drives(a, b):  # the rule both sides evaluate: front tile a, back tile b
        return (size[a] - cnt[a]) <= cnt[b]

    def load_positions(t):  # "load the flags of tile t": positions of its R's by RF and of its L's by LB
        lf = pre[t]
        rpos, lpos = {}, {}
        for j in range(t * tile, min(n, (t + 1) * tile)):
            if L[j]:
                lpos[nL - lf - 1] = j
                lf += 1
            else:
                rpos[j - lf] = j
        return rpos, lpos

    for x in range(nt):
        own_r, own_l = load_positions(x)
        cache = {}

        def partner(t):
            if t == x:
                return own_r, own_l
            if t not in cache:
                cache[t] = load_positions(t)
                loads[x] += 1
            return cache[t]

        def tile_of_l_back(k):  # tile holding the k-th L from the back: per-tile counts only
            for t in range(nt - 1, -1, -1):
                lo, hi = l_back_range(t)
                if lo <= k < hi:
                    return t
            raise AssertionError("rank out of range")

        def tile_of_r_front(k):
            for t in range(nt):
                lo, hi = r_front_range(t)
                if lo <= k < hi:
                    return t
            raise AssertionError("rank out of range")

        lf = pre[x]
        for j in range(x * tile, min(n, (x + 1) * tile)):
            rf = j - lf
            if j == f:
                put(j, pivot)
            elif j < f and L[j]:
                put(j, j)                      # front L stays
            elif j > f and not L[j]:
                put(j, j - 1)                  # back R shifts by one
            elif j < f:                        # front R number rf
                if rf == 0:
                    put(j, n - 1)
                else:
                    b = tile_of_l_back(rf - 1)
                    if b == x or a_drives(x, b):
                        put(j, partner(b)[1][rf - 1] - 1)
            else:                              # back L number lb: goes to the rf == lb front R
                lb = nL - lf - 1
                a = tile_of_r_front(lb)
                if a == x or not a_drives(a, x):
                    put(j, partner(a)[0][lb])
            lf += int(L[j])
        # moves this tile executes on behalf of partners that do not drive them
        # (a) as a front tile for back L's: L_back(k) -> position of my R with RF == k, for the partners b I drive
        for k, q in own_r.items():
            if q < f:
                # my front R number k is the landing place of L_back(k), if that L exists behind f
                if k < nL:
                    try:
                        b = tile_of_l_back(k)
                    except AssertionError:
                        b = None
                    if b is not None and b != x and a_drives(x, b):
                        p = partner(b)[1].get(k)
                        if p is not None and p > f:
                            put(p, q)
        # (b) as a back tile for front R's: R_front(m) -> (position of my L with LB == m-1) - 1
        for k, q in own_l.items():
            if q > f:
                try:
                    a = tile_of_r_front(k + 1)
                except AssertionError:
                    a = None
                if a is not None and a != x and not a_drives(a, x):
                    p = partner(a)[0].get(k + 1)
                    if p is not None and p < f:
                        put(p, q - 1)
    return dest, pivot, loads, moved


def pull_form(flags, tile, warp_slots=None):
    """The form the grid tier's resident-tile path runs (`t1_level_pull`): ONE barrier per shuffle.
    Before the barrier every tile publishes its L count and its ballot words (+ the exclusive L prefix of each of its warps);
    after it, every tile builds its own OUTPUT slots: src[p] = input position of the element that lands at p.  A slot is
    either kept (front L), shifted (the back R one to the right), the boundary element f (at `pivot`), or a hole that is
    filled by a rank-select in the ballots of a tile on the other side of f:
        p <  f, R(p)                         <-  the L that has RF(p) L's after it            (a back L)
        q = p+1 > f, L(q)  (or p == n-1)     <-  the R that has (L's after q) + 1 R's before it  (a front R; RF 0 for p = n-1)
    Only per-tile counts, the boundary flags and the partner's ballots are used.  Returns (src, pivot, partner tiles
    touched per tile)."""
    n = len(flags)
    L = np.asarray(flags, dtype=bool)
    nt = (n + tile - 1) // tile
    cnt = np.array([int(L[t * tile:(t + 1) * tile].sum()) for t in range(nt)], dtype=np.int64)
    pre = np.concatenate([[0], np.cumsum(cnt)])
    nL = int(pre[-1])
    f, pivot = boundary(L, n, nL)
    ws = warp_slots or max(1, tile // 8)  # slots per "warp" inside a tile: meta = per-warp exclusive L prefix + the bits

    def select_in_tile(t, r, want_L):
        """position of the r-th (0-based, from the tile's front) L (or R) of tile t, from its meta only"""
        lo, hi = t * tile, min(n, (t + 1) * tile)
        bits = L[lo:hi]
        nw = (tile + ws - 1) // ws
        wpre = [int(bits[:w * ws].sum()) for w in range(nw)]  # published: exclusive L prefix per warp
        if not want_L:
            wpre = [w * ws - wpre[w] for w in range(nw)]        # R prefix follows from it (tail slots beyond n count as R,
        w = max(i for i in range(nw) if wpre[i] <= r)           #  but ranks never reach them)
        r -= wpre[w]
        for j in range(lo + w * ws, min(lo + (w + 1) * ws, hi)):
            if bool(L[j]) == want_L:
                if r == 0:
                    return j
                r -= 1
        raise AssertionError("select past the end of the warp")

    def pos_L_with_after(k):  # the L that has k L's after it
        # tile t holds the L's whose L's-after count lies in [nL - pre[t+1], nL - pre[t])
        t = int(np.searchsorted(pre, nL - k, side="left")) - 1   # largest t with pre[t] < nL - k  <=>  nL - pre[t] > k
        assert nL - pre[t + 1] <= k < nL - pre[t]
        kk = k - (nL - pre[t + 1])                 # from the tile's back
        return select_in_tile(t, int(cnt[t]) - 1 - kk, True), t

    rpre = np.array([t * tile for t in range(nt)] + [n], dtype=np.int64) - pre  # R's before tile t

    def pos_R_with_before(k):  # the R that has k R's before it
        t = int(np.searchsorted(rpre, k, side="right")) - 1      # largest t with rpre[t] <= k
        assert rpre[t] <= k < rpre[t + 1]
        return select_in_tile(t, k - int(rpre[t]), False), t

    src = np.full(n, -1, dtype=np.int64)
    partners = np.zeros(nt, dtype=np.int64)
    for x in range(nt):
        seen = set()
        lf = int(pre[x])  # L's before p
        for p in range(x * tile, min(n, (x + 1) * tile)):
            lp = bool(L[p])
            q = p + 1
            if p == pivot:
                src[p] = f
            elif p < f and lp:
                src[p] = p
            elif p < f:
                src[p], t = pos_L_with_after(p - lf)
                seen.add(t)
            elif q < n and q > f and not L[q]:
                src[p] = q
            else:
                assert q == n or (q > f and L[q]), (p, f, pivot, n)
                k = 0 if q == n else (nL - (lf + int(lp)) - 1) + 1
                src[p], t = pos_R_with_before(k)
                seen.add(t)
            lf += int(lp)
        seen.discard(x)
        partners[x] = len(seen)
    return src, pivot, partners
