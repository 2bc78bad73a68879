import numpy as np


def assert_leafsets_equal(a, b, what=""):
    """Bit-exact comparison of two canonical (origin-sorted) leaf sets."""
    assert len(a) == len(b), f"{what}: leaf count {len(a)} != {len(b)}"
    if len(a) == 0:
        return
    assert np.array_equal(a.origins, b.origins), f"{what}: leaf origins differ"
    assert np.array_equal(a.active, b.active), f"{what}: active masks differ"
    if a.valmask is not None or b.valmask is not None:
        assert np.array_equal(a.valmask, b.valmask), f"{what}: value masks differ"
    if a.values is not None or b.values is not None:
        # bit patterns, not float equality (so -0.0 / NaN payloads would be caught too)
        assert np.array_equal(a.values.view(np.uint32), b.values.view(np.uint32)), f"{what}: leaf values differ"


def popcount64(a):
    a = np.ascontiguousarray(a, dtype=np.uint64)
    return int(np.unpackbits(a.view(np.uint8)).sum())


CFG_GTEST = (0.9, 0.1, 0.49, 0.51)     # tests/mapping.cpp config: the tile-probe quirk fires
CFG_ROS = (0.7, 0.4, 0.12, 0.97)       # typical vdb_mapping_ros config
