"""CPU: the byte arithmetic of the peer-to-peer slab exchange (dedalus/data_objects/slab.py: arena_layout) -
every block a rank pushes lands inside the destination arena, blocks of different sources never overlap, the
forward exchange is the mirror of the inverse one, and the peer-store tables address exactly the same bytes as
the copy lists.  Covers ranks that own no retained ky row (block slabs under 2/3 dealiasing)."""
import importlib.util
import itertools
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec = importlib.util.spec_from_file_location("ddl_slab_layout", os.path.join(ROOT, "dedalus-1.0_b200", "dedalus", "data_objects", "slab.py"))
slab = importlib.util.module_from_spec(spec)
spec.loader.exec_module(slab)

CASES = [
    (8, [64, 64, 43, 0, 0, 42, 64, 64], 64, 176, 512, 9),      # 512^3, block ky slabs
    (8, [43, 43, 43, 43, 43, 42, 42, 42], 64, 176, 512, 9),     # 512^3, cyclic
    (2, [11, 10], 16, 16, 32, 6),
    (4, [3, 0, 0, 2], 4, 8, 16, 4),
]


@pytest.mark.parametrize("P,rows,nzl,cx,nz,nmax", CASES)
def test_exchange_blocks_tile_the_destination_arenas(P, rows, nzl, cx, nz, nmax):
    lays = [slab.arena_layout(P, me, rows, nzl, cx, nz, nmax) for me in range(P)]
    blk = nzl * cx * 16
    cy = sum(rows)
    for direction in ("inv", "fwd"):
        for f in range(nmax):
            landed = {r: [] for r in range(P)}
            for me in range(P):
                ranks, src, dst, nb = lays[me][direction][f]
                assert sorted(ranks) == list(range(P)) and ranks[0] == me
                for s, so, do, n in zip(ranks, src, dst, nb):
                    assert 0 <= so and so + n <= lays[me]["bytes"]                  # source inside my arena
                    assert 0 <= do and do + n <= lays[s]["bytes"]                   # destination inside the peer's
                    landed[s].append((do, do + n, me))
                    # source region: my k-side field (inverse) / my x-side field (forward)
                    region = lays[me]["ks"][f] if direction == "inv" else lays[me]["xs"][f]
                    size = rows[me] * nz * cx * 16 if direction == "inv" else cy * blk
                    assert region <= so and so + n <= region + size
            for r in range(P):
                spans = sorted((a, b) for a, b, _ in landed[r] if b > a)
                for (a0, b0), (a1, b1) in zip(spans[:-1], spans[1:]):
                    assert b0 <= a1                                                 # no overlap
                total = sum(b - a for a, b in spans)
                want = cy * blk if direction == "inv" else rows[r] * nz * cx * 16   # exactly one field is filled
                assert total == want
                if spans:
                    base = lays[r]["xs"][f] if direction == "inv" else lays[r]["ks"][f]
                    assert spans[0][0] == base and spans[-1][1] == base + want


@pytest.mark.parametrize("P,rows,nzl,cx,nz,nmax", CASES)
def test_forward_is_the_mirror_of_inverse(P, rows, nzl, cx, nz, nmax):
    lays = [slab.arena_layout(P, me, rows, nzl, cx, nz, nmax) for me in range(P)]
    for f, me in itertools.product(range(nmax), range(P)):
        inv = {s: (so, do, n) for s, so, do, n in zip(*lays[me]["inv"][f])}
        for s in range(P):
            fwd_s = {t: (so, do, n) for t, so, do, n in zip(*lays[s]["fwd"][f])}
            so, do, n = inv[s]
            so2, do2, n2 = fwd_s[me]
            assert (so2, do2, n2) == (do, so, n)       # what I pushed to s comes back from s to where it came from


@pytest.mark.parametrize("P,rows,nzl,cx,nz,nmax", CASES)
def test_peer_store_tables_address_the_copy_destinations(P, rows, nzl, cx, nz, nmax):
    cy0 = [sum(rows[:r]) for r in range(P)]
    blk = nzl * cx * 16
    for me in range(P):
        lay = slab.arena_layout(P, me, rows, nzl, cx, nz, nmax)
        for f in range(nmax):
            inv = {s: do for s, so, do, n in zip(*lay["inv"][f])}
            fwd = {s: do for s, so, do, n in zip(*lay["fwd"][f])}
            for s in range(P):
                assert lay["zinv_peer"][f][s] == inv[s]                              # first row of my block in s's x-side field
                # y pass: x-side position p = cy0[s] + j (rank s's j-th row) -> base + p*blk = block `me`, row j of s's k-side field
                if rows[s]:
                    assert lay["yfwd_peer"][f][s] + cy0[s] * blk == fwd[s]


def test_retained_boxes_cover_exactly_the_kept_modes():
    """Plan.retained_boxes() (the index boxes comp.upload_retained / download_retained hand to ddl_copy_boxes): their union is
    exactly the set of modes the plan's mask keeps, for every rank of block and cyclic ky ownership; no two boxes overlap.
    One child process (host-emulation harness) for all cases."""
    code = (
        "import os, sys, numpy as np\n"
        "os.environ['DDL_TEST_HOST_EMUL'] = '1'\n"
        "sys.path.insert(0, os.path.join(%r, 'tests'))\n"
        "import conftest\n"
        "from dedalus.data_objects.plan import Plan\n"
        "done = 0\n"
        "for shape, dealiasing in (((12, 16, 20), '2/3 cython'), ((16, 16, 32), '2/3 cython'), ((16, 24), '2/3 cython'), ((8, 16, 16), 'None')):\n"
        "  for nranks, layout in ((1, 'block'), (2, 'block'), (4, 'block'), (2, 'cyclic'), (4, 'cyclic')):\n"
        "    if len(shape) == 2 and nranks > 1:\n"
        "        continue\n"
        "    for rank in range(nranks):\n"
        "        pl = Plan(shape, (2 * np.pi,) * len(shape), dealiasing, nranks, rank, layout)\n"
        "        shape3, boxes = pl.retained_boxes()\n"
        "        local = tuple(int(n) for n in pl.kshape_local)\n"
        "        assert tuple(shape3) == (1,) * (3 - len(local)) + local\n"
        "        hits = np.zeros(tuple(shape3), dtype=int)\n"
        "        for b in boxes:\n"
        "            hits[b[0]:b[1], b[2]:b[3], b[4]:b[5]] += 1\n"
        "        keep = np.ones(local, dtype=bool)\n"
        "        for name, axis in pl.ktrans.items():\n"
        "            if not isinstance(name, str):\n"
        "                continue\n"
        "            kp = np.asarray(pl.keep_np[name]).astype(bool).ravel()\n"
        "            if axis == 0 and nranks > 1:\n"
        "                kp = kp[pl.krows]\n"
        "            sh = [1] * len(local); sh[axis] = len(kp)\n"
        "            keep = keep & kp.reshape(sh)\n"
        "        assert hits.max() == 1 and np.array_equal(hits.reshape(local) == 1, keep), (shape, nranks, layout, rank)\n"
        "        assert len(boxes) <= 4 and 0 < keep.sum() < keep.size\n"
        "        done += 1\n"
        "print('ok', done)\n") % (ROOT,)
    r = subprocess.run([sys.executable, "-c", code], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert r.returncode == 0 and "ok 40" in r.stdout, r.stdout[-2000:]


@pytest.mark.parametrize("P,rows,nzl,cx,nz,nmax", CASES)
def test_push_tables_move_exactly_the_bytes_of_the_copy_lists(P, rows, nzl, cx, nz, nmax):
    """The "push" exchange (SlabPipeline._push_table -> ddl_p2p_push) moves rows x pitch rectangles: per field group in the inverse
    direction, per plane chunk in the forward one.  Over all groups / chunks they must cover, byte for byte and without overlap,
    what the whole-field copy lists of arena_layout cover (which the tests above tie to the arenas)."""
    for me in range(P):
        lay = slab.arena_layout(P, me, rows, nzl, cx, nz, nmax)
        pipe = type("P", (), {})()                           # the four attributes _push_table reads
        pipe._p2p_layout, pipe.nzl, pipe.cx, pipe.rows = lay, nzl, cx, rows

        def spans(key):
            n, ranks, src, dst, rb, nr, pitch = slab.SlabPipeline._push_table(pipe, key)
            out = []
            for i in range(n):
                for r in range(nr[i]):
                    out.append((ranks[i], src[i] + r * pitch[i], dst[i] + r * pitch[i], rb[i]))
            return out

        def cover(sp):
            """{(rank, dst byte)} and {src byte} sets, asserting no byte is written or read twice"""
            d, s = set(), set()
            for rank, so, do, nb in sp:
                if nb == 0:
                    continue
                for b in range(0, nb, 16):
                    assert (rank, do + b) not in d and (so + b) not in s
                    d.add((rank, do + b))
                    s.add(so + b)
            return d, s

        small = nzl * cx * sum(rows) * nmax <= 40000          # exhaustive byte sets only where they stay small
        # inverse: groups [0, 2), [2, nmax) against the whole-field lists of the same fields
        for f0, f1 in ((0, 2), (2, nmax)):
            got = spans(("inv", f0, f1))
            want = [(s, so, do, nb) for f in range(f0, f1) for s, so, do, nb in zip(*lay["inv"][f])]
            assert sorted(got) == sorted(want)
        # forward: chunks of the local planes
        for nch in (1, 2, nzl):
            zc = nzl // nch
            got = [x for c in range(nch) for x in spans(("fwd", nmax, c * zc, zc))]
            assert sum(nb for _, _, _, nb in got) == sum(nb for f in range(nmax) for nb in lay["fwd"][f][3])
            if small:
                gd, gs = cover(got)
                wd, ws = cover([(s, so, do, nb) for f in range(nmax) for s, so, do, nb in zip(*lay["fwd"][f])])
                assert gd == wd and gs == ws
