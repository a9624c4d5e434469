"""Cell-identified state through the C ABI (sb2_extract_state / sb2_apply_state): the reference's test story
(test/api_test.cpp:42-98; shyft/tests/api/test_region_model_stacks.py:71-79) and exact agreement with the oracle."""
import numpy as np
import pytest

from fixtures import geo_matrix

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def sb():
    import shyft_b200
    return shyft_b200


def _ids(v):
    return [tuple(int(t) for t in r) for r in v.ids.tolist()]


def test_reference_story(sb):
    geo = sb.geo_cell_data_vector([1.0, 1.0, 2.0, 2.0], [1.0, 2.0, 1.0, 2.0], [1.0] * 4, area=10.0, catchment_id=np.array([1, 1, 2, 2]))
    m = sb.PTGSKOptModel(geo)
    st = np.tile(np.array([0.4, 0.1, 30000.0, 1.26, 0.0, 0.0, 0.0, 0.0, 0.0]), (4, 1))
    st[:, 8] = [1.1, 1.2, 2.1, 2.2]
    m.set_states(st)
    s0 = m.state.extract_state([])
    assert len(s0) == m.size()
    assert _ids(s0) == [(1, 1, 1, 10), (1, 1, 2, 10), (2, 2, 1, 10), (2, 2, 2, 10)]
    assert np.array_equal(s0.states, st) and np.array_equal(s0.state_vector, st)
    blob = s0.serialize_to_bytes()
    assert len(blob) > 10
    s0_x = sb.StateWithIdVector.deserialize_from_bytes(blob)
    assert np.array_equal(s0_x.ids, s0.ids) and np.array_equal(s0_x.states, s0.states)
    s1 = m.state.extract_state([2])
    assert len(s1) == 2 and _ids(s1) == _ids(s0)[2:] and np.array_equal(s1.states, st[2:])
    assert len(m.state.extract_state([3])) == 0
    assert m.state.apply_state(s0, []) == []
    assert m.state.apply_state(s0, [4]) == []
    s0.ids["cid"][0] = 4
    assert m.state.apply_state(s0, [4]) == [0]
    assert np.array_equal(m.get_states(), st)


@pytest.mark.parametrize("cls_name,stack", [("PTGSKModel", 0), ("PTHSKModel", 1), ("HbvStackModel", 2)])
def test_against_the_oracle_on_a_region(sb, oracle, cls_name, stack):
    from shyft_b200 import synthetic
    n = 3000
    geo, ta, env = synthetic.make_region(n, 24, 9, config_index=1, cells_per_catchment=250)
    G = geo_matrix(geo)
    m = getattr(sb, cls_name)(geo)
    rng = np.random.default_rng(5)
    st = rng.uniform(0.1, 5.0, size=(n, m.state_size))
    m.set_states(st)
    cids = [2, 5, 11]
    got = m.state.extract_state(cids)
    want = oracle.extract_state(G, st, cids)
    assert _ids(got) == [w[0] for w in want] and np.array_equal(got.states, np.stack([w[1] for w in want]))
    assert np.array_equal(sb.cell_state_id_of(geo), m.state.extract_state([]).ids)
    # apply: shuffled, scaled states for catchments 5 and 11 plus two strangers; scope = [5, 9, 11]
    perm = rng.permutation(len(got))
    ids = got.ids[perm].copy()
    rows = got.states[perm] * 3.0
    ids["x"][:2] += 7           # no such cell
    v = sb.StateWithIdVector(ids, rows)
    missing = m.state.apply_state(v, [5, 9, 11])
    st_want, missing_want = oracle.apply_state(G, st, [(tuple(int(t) for t in ids[k].tolist()), rows[k]) for k in range(len(ids))], [5, 9, 11])
    assert missing == missing_want
    assert np.array_equal(m.get_states(), st_want)


def test_reference_python_state_with_id_handler_story(sb):
    """shyft/tests/api/test_region_model_stacks.py:556-600 (test_state_with_id_handler), two catchments"""
    n = 20
    cid = np.array([1 + (i * 7) % 2 for i in range(n)])
    geo = sb.geo_cell_data_vector(500 + 1000.0 * np.arange(n), np.full(n, 500.0), 500.0 * np.arange(n) / n, catchment_id=cid, glacier=0.01,
                                  lake=0.05, reservoir=0.19, forest=0.3)
    m = sb.PTGSKModel(geo)
    s12, s1, s2 = m.state.extract_state([]), m.state.extract_state([1]), m.state.extract_state([2])
    assert len(s1) + len(s2) == len(s12) and len(s1) > 0 and len(s2) > 0
    ms2 = sb.StateWithIdVector.deserialize_from_bytes(s2.serialize_to_bytes())
    assert len(ms2) == len(s2) and np.array_equal(ms2.ids, s2.ids) and np.array_equal(ms2.states[:, 8], s2.states[:, 8])
    assert np.all(s1.ids["cid"] == 1) and np.all(s2.ids["cid"] == 2)
    s12.states[:, 8] = 100 + np.arange(len(s12))
    m.state.apply_state(s12, [])
    ms_12 = m.state.extract_state([])
    assert np.array_equal(ms_12.states[:, 8], 100.0 + np.arange(n))
    s2.states[:, 8] = 200 + np.arange(len(s2))
    assert m.state.apply_state(s2, [2]) == []
    ms_12 = m.state.extract_state([])
    for i in range(n):
        if ms_12.ids["cid"][i] == 1:
            assert ms_12.states[i, 8] == 100 + i
    assert np.array_equal(m.state.extract_state([2]).states[:, 8], 200.0 + np.arange(len(s2)))
