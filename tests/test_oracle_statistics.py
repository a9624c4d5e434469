"""The oracle's restatement of cell_statistics (core/cell_model.h:194-406) on hand-checkable numbers."""
import numpy as np
import pytest


def test_sum_and_area_average_by_catchment_and_cell(oracle):
    series = np.array([[1.0, 2.0, 3.0, 4.0], [10.0, 20.0, 30.0, 40.0]])   # [T=2][cells=4]
    area = np.array([1.0, 1.0, 2.0, 4.0])
    cids = [7, 7, 9, 9]
    assert np.array_equal(oracle.sum_catchment_feature(series, cids), [10.0, 100.0])
    assert np.array_equal(oracle.sum_catchment_feature(series, cids, [9]), [7.0, 70.0])
    assert np.array_equal(oracle.sum_catchment_feature(series, cids, [0, 3], oracle.CELL_IX), [5.0, 50.0])
    avg = oracle.average_catchment_feature(series, area, cids, [9])
    assert avg == pytest.approx([(3 * 2 + 4 * 4) / 6.0, (30 * 2 + 40 * 4) / 6.0], rel=1e-15)
    assert oracle.average_catchment_feature_value(series, area, cids, [], 1) == pytest.approx((10 + 20 + 60 + 160) / 8.0, rel=1e-15)
    assert oracle.sum_catchment_feature_value(series, cids, [7], 0) == 3.0
    assert np.array_equal(oracle.catchment_feature(series, cids, [9, 7], 1), [10.0, 20.0, 30.0, 40.0])   # cell order, not index order
    assert np.array_equal(oracle.catchment_feature(series, cids, [2], 0, oracle.CELL_IX), [3.0])


def test_a_cell_matching_two_indexes_counts_once(oracle):
    series = np.ones((1, 3))
    assert oracle.sum_catchment_feature(series, [1, 1, 2], [1, 1, 2])[0] == 3.0     # `break` after the first match (:253,:325)


def test_unknown_indexes_raise_like_verify_cids_exist(oracle):
    series = np.ones((1, 3))
    with pytest.raises(RuntimeError, match="one or more supplied catchment_indexes does not exist:5"):
        oracle.sum_catchment_feature(series, [1, 1, 2], [5])
    with pytest.raises(RuntimeError, match="Supplied cell index reference 4 is ouside valid range 0 ..3"):
        oracle.sum_catchment_feature(series, [1, 1, 2], [4], oracle.CELL_IX)
    assert oracle.sum_catchment_feature(series, [1, 1, 2], [3], oracle.CELL_IX)[0] == 0.0   # == size passes the reference's check
