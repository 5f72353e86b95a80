"""Numpy model of Tlas::build (crates/bvh/src/tlas.rs:56-105) with a BEST-MATCH CACHE: every live slot remembers the answer
of find_best_match for itself, the cache is repaired in the same pass that evaluates the node a merge creates, and the chain
walk (a, b = best(a), c = best(b), ...) reads the cache instead of scanning.  Test infrastructure only (imported by
tests/test_oracle.py): it pins the cache rules — which entries a merge invalidates, what swap-remove does to the first-index
tie-break, the stale slot `a` — against the oracle's byte output before the rules go into csrc/tlas.cu, and it counts the full
scans that remain.

Rules (slots a, b merge into N; `last` = count - 1 is swap-removed into slot b):
  * N lives at slot a, or at slot b when a == last (then node_indices[b] = node_indices[last] copies N and a goes stale);
  * an entry whose partner is a or b pointed at a consumed node: INVALID (rescanned when the walk asks for it);
  * an entry whose partner is `last` follows the move: same area, partner b (a lower index can only confirm a first-index win);
  * every other valid entry is offered two candidates: (area with N, slot of N) and (area with the moved node, b) — the moved
    node is not new, but its index dropped, so it can now win a tie it used to lose;
  * the moved node brings its own entry along (same rules);  N's entry is the pass's reduction;
  * find_best_match for the stale slot a scans every live slot, the copy of N in slot b included.
"""
import numpy as np

NONE = np.uint64(0xFFFFFFFFFFFFFFFF)   # no candidate below 1e30: find_best_match returns the target itself
F32 = np.float32


def _area(lo, hi):
    """Aabb::area (intersection.rs:16-19): (dx*dy + dx*dz + dy*dz) * 2, left to right, in f32."""
    d = (hi - lo).astype(F32)
    s = ((d[..., 0] * d[..., 1]).astype(F32) + (d[..., 0] * d[..., 2]).astype(F32)).astype(F32)
    s = (s + (d[..., 1] * d[..., 2]).astype(F32)).astype(F32)
    return (s * F32(2.0)).astype(F32)


def _keys(tlo, thi, lo, hi, idx):
    """key = area bits << 32 | slot for every candidate box, NONE where the union's area is not < 1e30."""
    a = _area(np.minimum(tlo, lo), np.maximum(thi, hi))
    ok = a < F32(1e30)
    k = (a.view(np.uint32).astype(np.uint64) << np.uint64(32)) | idx.astype(np.uint64)
    return np.where(ok, k, NONE)


def _min_rs(acc, v):
    return np.where(v < acc, v, acc)


def _max_rs(acc, v):
    return np.where(v > acc, v, acc)


def build(leaf_lo, leaf_hi, eager=True):
    """leaf boxes [I,3] (slots 1..I of the TLAS) -> (lo[2I+1,3], hi[2I+1,3], left_right[2I+1], kids[2I+1,2], stats)."""
    n = leaf_lo.shape[0]
    lo = np.zeros((2 * n + 1, 3), F32); hi = np.zeros((2 * n + 1, 3), F32)
    lr = np.zeros(2 * n + 1, np.uint32); kids = np.zeros((2 * n + 1, 2), np.uint32)
    lo[1:n + 1], hi[1:n + 1] = leaf_lo, leaf_hi
    blo, bhi = leaf_lo.astype(F32).copy(), leaf_hi.astype(F32).copy()   # live slot boxes
    ni = np.arange(1, n + 1, dtype=np.int64)
    ck = np.full(n, NONE, np.uint64); valid = np.zeros(n, bool)
    count, used = n, n + 1
    stats = {"scans": 0, "lookups": 0, "merges": 0, "init_scans": 0}
    ar = np.arange(n, dtype=np.int64)

    def scan(tlo, thi, exclude):
        k = _keys(tlo, thi, blo[:count], bhi[:count], ar[:count])
        if 0 <= exclude < count:
            k[exclude] = NONE
        return k.min() if count else NONE

    def best(slot):
        """find_best_match(target = slot) for a live slot, through the cache."""
        if valid[slot]:
            stats["lookups"] += 1
        else:
            stats["scans"] += 1
            ck[slot] = scan(blo[slot], bhi[slot], slot)
            valid[slot] = True
        return slot if ck[slot] == NONE else int(ck[slot] & np.uint64(0xFFFFFFFF))

    if eager:
        for s in range(n):
            ck[s] = scan(blo[s], bhi[s], s); valid[s] = True
        stats["init_scans"] = n
    a = 0
    b = best(a)
    while count > 0:
        c = best(b)
        if a != c:
            a, b = b, c
            continue
        # ---- merge (tlas.rs:62-79) ----
        stats["merges"] += 1
        ia, ib = int(ni[a]), int(ni[b])
        ulo = _min_rs(blo[a], blo[b]); uhi = _max_rs(bhi[a], bhi[b])
        lo[used], hi[used] = ulo, uhi
        lr[used] = np.uint32((ia + (ib << 16)) & 0xFFFFFFFF); kids[used] = (ia, ib)
        last = count - 1
        if a == b:
            # find_best_match found nothing below 1e30 and returned its target: the slot merges with itself (always the
            # end of the build, tlas.rs:61; with more than one live slot only for boxes of area >= 1e30).  The reference's
            # three assignments are replayed literally and the cache starts over.
            ni[a] = used; blo[a], bhi[a] = ulo, uhi
            ni[b] = ni[last]; blo[b], bhi[b] = blo[last].copy(), bhi[last].copy()
            count -= 1; used += 1
            valid[:] = False
            if count == 0:
                break
            stats["scans"] += 1
            r = scan(blo[a], bhi[a], a if a < count else -1)
            b = a if r == NONE else int(r & np.uint64(0xFFFFFFFF))
            continue
        moved = last != a and last != b  # the node in slot `last` is swap-removed into slot b
        mlo, mhi = blo[last].copy(), bhi[last].copy()
        mk, mv = ck[last], valid[last]
        ni[a] = used; blo[a], bhi[a] = ulo, uhi
        ni[b] = ni[last]; blo[b], bhi[b] = blo[last].copy(), bhi[last].copy()   # copies N when last == a
        if moved:
            ck[b], valid[b] = mk, mv
        count -= 1; used += 1
        s_new = a if a < count else b    # a == last: the live copy of N is the one in slot b
        if count == 0:
            break
        # ---- cache repair + N's own entry: one pass over the live slots ----
        stats["scans"] += 1
        live = ar[:count]
        k = ck[:count].copy(); v = valid[:count].copy()
        part = (k & np.uint64(0xFFFFFFFF)).astype(np.int64)
        has = k != NONE
        dead = v & has & ((part == a) | (part == b))
        v &= ~dead
        if moved:
            ren = v & has & (part == last)
            k = np.where(ren, (k & ~np.uint64(0xFFFFFFFF)) | np.uint64(b), k)
        cand_n = _keys(ulo, uhi, blo[:count], bhi[:count], np.full(count, s_new, np.int64))
        k = np.where(v, np.minimum(k, cand_n), k)
        if moved:
            cand_m = _keys(mlo, mhi, blo[:count], bhi[:count], np.full(count, b, np.int64))
            cand_m[b] = NONE
            k = np.where(v, np.minimum(k, cand_m), k)
        q = _keys(ulo, uhi, blo[:count], bhi[:count], live)
        q[s_new] = NONE
        k[s_new] = q.min(); v[s_new] = True
        ck[:count], valid[:count] = k, v
        # ---- b = find_best_match(a) (tlas.rs:79); a may be the stale slot ----
        if a < count:
            b = a if ck[a] == NONE else int(ck[a] & np.uint64(0xFFFFFFFF))
        else:
            self_key = _keys(ulo, uhi, ulo[None, :], uhi[None, :], np.array([s_new]))[0]
            r = min(ck[s_new], self_key)
            b = a if r == NONE else int(r & np.uint64(0xFFFFFFFF))
    root = int(ni[a])
    lo[0], hi[0], lr[0], kids[0] = lo[root], hi[root], lr[root], kids[root]
    return lo, hi, lr, kids, stats
