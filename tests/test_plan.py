"""The native host graph compiler (gbp_plan_*: the storage order gbp_ba_create lays out), checked on CPU against an
independent NumPy statement of its rules.  The reference builds its factor list with an O(C F) scan
(gbp/gbp_ba.py:128-143); the order that scan produces -- camera-major, file order inside a camera -- is the factor
order here, everything else (tiles, landmark blocks, CSR tables) is engine layout with no reference counterpart."""
import numpy as np
import pytest

from conftest import golden_problem, load_golden


def _plan(cam, lmk, C, L, **kw):
    from gbp_b200.engine import compile_plan
    return compile_plan(cam, lmk, C, L, **kw)


def _blocks(L, lblock, K):
    """Landmark -> (block, chunk): every chunk [L k / K, L (k + 1) / K) is cut into blocks of lblock landmarks, numbered chunk by chunk."""
    blk, chunk, nb = np.zeros(L, np.int64), np.zeros(L, np.int64), 0
    for k in range(K):
        l0, l1 = L * k // K, L * (k + 1) // K
        for b0 in range(l0, l1, lblock):
            blk[b0:min(l1, b0 + lblock)] = nb
            nb += 1
        chunk[l0:l1] = k
    return blk, chunk


def _check(plan, cam, lmk, C, L, lblock, K=1):
    T, tiles, F = plan["T"], plan["tiles"], len(cam)
    blk_of, chunk_of = _blocks(L, lblock, K)
    order = np.argsort(cam, kind="stable")
    # factor order = the reference's: stable sort of the measurement list by camera
    assert np.array_equal(plan["file_of_factor"], order)
    assert np.array_equal(plan["adj"][:, 0], cam[order]) and np.array_equal(plan["adj"][:, 1], lmk[order])
    fcam, flmk = cam[order], lmk[order]
    # slots: injective, inside the tile range, padding slots unmapped
    sl = plan["slot_of_factor"]
    assert len(np.unique(sl)) == F and (sl >= 0).all() and (sl < plan["n_slots"]).all() and plan["n_slots"] == len(tiles) * T
    t_of, pos = sl // T, sl % T
    assert (pos < tiles[t_of, 1]).all()                              # inside the valid part of its tile
    assert np.array_equal(np.bincount(t_of, minlength=len(tiles)), tiles[:, 1]) and (tiles[:, 1] >= 1).all() and (tiles[:, 1] <= T).all()
    assert np.array_equal(tiles[t_of, 0], fcam)                      # every tile holds edges of ONE keyframe
    # tiles are sorted by (landmark block, keyframe); inside a run the factor order is kept and only the last tile is ragged
    key = blk_of[flmk] * max(C, 1) + fcam
    tile_key = np.full(len(tiles), -1, np.int64)
    tile_key[t_of] = key
    assert (np.diff(tile_key) >= 0).all()
    assert np.array_equal(np.argsort(key, kind="stable"), np.argsort(sl))      # storage order = stable sort by run key
    for k in np.unique(tile_key):
        cnt = tiles[tile_key == k, 1]
        assert (cnt[:-1] == T).all()
    # per-slot landmark index, zero in padding
    idx = np.zeros(plan["n_slots"], np.int32)
    idx[sl] = flmk
    assert np.array_equal(plan["lmk_idx"], idx)
    # CSR by landmark over slots, factor order inside a landmark (= adj_factors order of the reference)
    assert np.array_equal(np.diff(plan["lmk_ptr"]), np.bincount(flmk, minlength=L))
    lorder = np.argsort(flmk, kind="stable")
    assert np.array_equal(plan["lmk_slots"], sl[lorder])
    # CSR by keyframe over tiles, tile order inside a keyframe
    assert np.array_equal(np.diff(plan["cam_tile_ptr"]), np.bincount(tiles[:, 0], minlength=C))
    assert np.array_equal(plan["cam_tiles"], np.argsort(tiles[:, 0], kind="stable"))
    # landmark chunks: a tile lies in ONE chunk; a keyframe's tile list is chunk-major and cam_chunk_ptr marks the chunk starts
    assert plan["n_chunks"] == K
    tchunk = np.full(len(tiles), -1, np.int64)
    tchunk[t_of] = chunk_of[flmk]
    for t in range(len(tiles)):
        assert (chunk_of[flmk[t_of == t]] == tchunk[t]).all()
    assert np.array_equal(plan["tile_chunk"], tchunk)
    ccp = plan["cam_chunk_ptr"]
    assert ccp.shape == (C, K + 1)
    for c in range(C):
        assert ccp[c, 0] == plan["cam_tile_ptr"][c] and ccp[c, K] == plan["cam_tile_ptr"][c + 1] and (np.diff(ccp[c]) >= 0).all()
        for k in range(K):
            assert (tchunk[plan["cam_tiles"][ccp[c, k]:ccp[c, k + 1]]] == k).all()


@pytest.mark.parametrize("name", ["fr1desk_vsmall", "fr1desk"])
@pytest.mark.parametrize("tile,block", [(0, 0), (32, 100), (64, 0), (128, 64)])
def test_plan_of_the_reference_problems(built_library, name, tile, block):
    P = golden_problem(load_golden(name))
    plan = _plan(P.cam_id, P.lmk_id, P.n_keyframes, P.n_points, tile_edges=tile, lmk_block=block)
    assert plan["T"] == (tile or 32)                                   # small graphs: 32-edge tiles
    _check(plan, np.asarray(P.cam_id), np.asarray(P.lmk_id), P.n_keyframes, P.n_points, block or max(P.n_points, 1))


@pytest.mark.parametrize("K,block", [(2, 0), (8, 0), (4, 50), (3, 7)])
def test_plan_with_landmark_chunks(built_library, K, block):
    """gbp_config.lmk_chunks: blocks never straddle a chunk, so that a rank holding whole chunks lays out exactly the tiles the
    single-GPU plan has for those chunks (the basis of bit-identical results on 1, 2, 4 and 8 GPUs)."""
    P = golden_problem(load_golden("fr1desk"))
    cam, lmk = np.asarray(P.cam_id), np.asarray(P.lmk_id)
    plan = _plan(cam, lmk, P.n_keyframes, P.n_points, tile_edges=32, lmk_block=block, chunks=K)
    _check(plan, cam, lmk, P.n_keyframes, P.n_points, block or max(P.n_points, 1), K)
    # the sub-problem of chunk range [k0, k1) planned on its own = the corresponding tiles of the whole plan
    from gbp_b200.dist import local_problem
    for world in (2,) if K % 2 == 0 else ():
        for r in range(world):
            sub, sel, (l0, l1) = local_problem(P, r, world)
            lp = _plan(sub.cam_id, sub.lmk_id, P.n_keyframes, sub.n_points, tile_edges=32, lmk_block=block or max(P.n_points, 1),
                       chunks=(K // world, r * K // world, K, l0, P.n_points))
            mine = np.nonzero((plan["tile_chunk"] >= r * K // world) & (plan["tile_chunk"] < (r + 1) * K // world))[0]
            assert np.array_equal(lp["tiles"], plan["tiles"][mine])
            whole = plan["lmk_idx"].reshape(-1, 32)[mine]
            local = lp["lmk_idx"].reshape(-1, 32)
            valid = np.arange(32)[None, :] < lp["tiles"][:, 1:2]
            assert np.array_equal((local + l0)[valid], whole[valid])            # the same edges in the same slots
            assert np.array_equal(lp["tile_chunk"] + r * K // world, plan["tile_chunk"][mine])


def test_plan_shuffled_file_order_and_isolated_variables(built_library):
    """Measurements in random file order, keyframes and landmarks without any measurement, a landmark seen 70 times."""
    rng = np.random.default_rng(3)
    C, L, F = 7, 50, 400
    cam = rng.integers(0, C - 2, F).astype(np.int32)          # keyframes 5, 6 never observed
    lmk = rng.integers(0, L - 5, F).astype(np.int32)          # landmarks 45.. never observed
    lmk[:70] = 3
    plan = _plan(cam, lmk, C, L, tile_edges=32, lmk_block=16)
    _check(plan, cam, lmk, C, L, 16)
    assert plan["cam_tile_ptr"][5] == plan["cam_tile_ptr"][7] and plan["lmk_ptr"][45] == plan["lmk_ptr"][50]


def test_plan_large_graph_auto_tiling(built_library):
    """Automatic choices: 64-edge tiles from 56832 factors on; landmark chunks of at least 125000 landmarks, a power of two <= 8
    of them (the keyframe-side sums are associated chunk by chunk, identically on 1, 2, 4 and 8 GPUs); landmark blocks of at
    most 262144 inside a chunk."""
    from gbp_b200.synthetic import make_synthetic
    prob = make_synthetic(50, 300_000, 4, seed=2)
    plan = _plan(prob.cam_id, prob.lmk_id, prob.n_keyframes, prob.n_points)
    assert plan["T"] == 64
    _check(plan, np.asarray(prob.cam_id), np.asarray(prob.lmk_id), prob.n_keyframes, prob.n_points, 262144, K=2)
    assert plan["n_chunks"] == 2 and len(np.unique(plan["tile_chunk"])) == 2
    for n_lmk, want in ((100, 1), (124_999, 1), (249_999, 1), (250_000, 2), (999_999, 4), (1_000_000, 8), (5_000_000, 8)):
        tiny = _plan(np.zeros(1, np.int32), np.zeros(1, np.int32), 1, n_lmk)
        assert tiny["n_chunks"] == want, (n_lmk, tiny["n_chunks"], want)
    waste = plan["n_slots"] / len(prob.cam_id) - 1.0
    assert waste < 0.05                                                 # padding slots: < 5 % on this graph


def test_plan_errors_and_empty(built_library):
    from gbp_b200 import _lib as L
    with pytest.raises(L.GbpError, match="landmark id"):
        _plan(np.array([0, 1], np.int32), np.array([0, 9], np.int32), 2, 3)
    with pytest.raises(L.GbpError, match="camera id"):
        _plan(np.array([0, -1], np.int32), np.array([0, 1], np.int32), 2, 3)
    with pytest.raises(L.GbpError, match="tile_edges"):
        _plan(np.array([0], np.int32), np.array([0], np.int32), 1, 1, tile_edges=48)
    empty = _plan(np.zeros(0, np.int32), np.zeros(0, np.int32), 2, 3)
    assert empty["n_tiles"] == 0 and empty["n_slots"] == 0 and np.array_equal(empty["lmk_ptr"], np.zeros(4, np.int32))
