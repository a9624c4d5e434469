"""Cell-identified state (api/api_state.h): the oracle's restatement through the reference's own test story
(test/api_test.cpp:42-98, test_state_with_id_functionality)."""
import numpy as np


def _four_cells():
    # geo_cell_data(geo_point(x, y, 1), area 10, cid): (1,1) (1,2) in catchment 1, (2,1) (2,2) in catchment 2; kirchner.q 1.1 1.2 2.1 2.2
    geo = np.zeros((4, 12))
    geo[:, 0] = [1, 1, 2, 2]
    geo[:, 1] = [1, 2, 1, 2]
    geo[:, 2] = 1
    geo[:, 3] = 10
    geo[:, 4] = [1, 1, 2, 2]
    st = np.tile(np.array([0.4, 0.1, 30000.0, 1.26, 0.0, 0.0, 0.0, 0.0, 0.0]), (4, 1))
    st[:, 8] = [1.1, 1.2, 2.1, 2.2]
    return geo, st


def test_reference_story(oracle):
    geo, st = _four_cells()
    a, b, c = (1, 2, 3, 10), (1, 2, 4, 20), (2, 2, 4, 20)
    assert a != b and c not in {a: 10, b: 20} and a in {a: 10, b: 20}
    assert a < b < c     # cell_state_id::operator< is the lexicographic order of (cid, x, y, area)
    s0 = oracle.extract_state(geo, st)
    assert len(s0) == 4 and [s[0] for s in s0] == [oracle.cell_state_id_of(g) for g in geo]
    s1 = oracle.extract_state(geo, st, [2])
    assert [s[0] for s in s1] == [oracle.cell_state_id_of(g) for g in geo[2:]]
    assert oracle.extract_state(geo, st, [3]) == []
    _, m0 = oracle.apply_state(geo, st, s0)
    assert m0 == []
    _, m0_x = oracle.apply_state(geo, st, s0, [4])
    assert m0_x == []          # "because we passed in states not containing 4"
    s0[0] = ((4,) + s0[0][0][1:], s0[0][1])
    _, m0_y = oracle.apply_state(geo, st, s0, [4])
    assert m0_y == [0]


def test_ids_truncate_toward_zero_and_states_land_in_their_cells(oracle):
    geo, st = _four_cells()
    geo[:, 0] = [1.9, -1.9, 2.2, 2.7]
    geo[:, 3] = 10.99
    assert [oracle.cell_state_id_of(g) for g in geo] == [(1, 1, 1, 10), (1, -1, 2, 10), (2, 2, 1, 10), (2, 2, 2, 10)]
    ids = oracle.extract_state(geo, st)
    shuffled = [ids[3], ids[0], ids[2], ids[1]]
    new = [(sid, row * 2.0) for sid, row in shuffled]
    st2, missing = oracle.apply_state(geo, st, new, [2])
    assert missing == []
    assert np.array_equal(st2[:2], st[:2]) and np.array_equal(st2[2:], st[2:] * 2.0)
