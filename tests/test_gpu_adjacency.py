"""GPU parity: voxel adjacency builder against the oracle restating pyfunc.py:48-76 and the tools variant."""
import numpy as np
import pytest

import oracle

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("conn", [26, 6])
@pytest.mark.parametrize("shape,frac", [((9, 11, 7), 0.5), ((6, 5, 4), 1.0), ((12, 8, 10), 0.15)])
def test_voxel_adjacency_matches_oracle(conn, shape, frac):
    from tfce_mediation_b200.pyfunc import create_adjac_voxel
    rs = np.random.RandomState(1)
    mask = rs.rand(*shape) < frac
    got = create_adjac_voxel(mask, mask.astype(np.float32), int(mask.sum()), dirtype=conn)
    want = oracle.voxel_adjacency(mask, conn)
    assert len(got) == len(want)
    assert all(list(g) == list(w) for g, w in zip(got, want))


@pytest.mark.parametrize("conn", [26, 6])
def test_voxel_adjacency_tools_variant(conn):
    from tfce_mediation_b200.pyfunc import create_adjac_voxel_tools
    rs = np.random.RandomState(2)
    mask = rs.rand(10, 9, 8) < 0.4
    got = create_adjac_voxel_tools(mask, dirtype=conn)
    want = oracle.voxel_adjacency_tools(mask, conn)
    assert len(got) == len(want)
    assert all(set(g) == set(w) for g, w in zip(got, want))
