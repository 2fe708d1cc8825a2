"""The sparse-pool entry of the oracle (orc_raycast_slots) that the full-size parity gate uses (tests/parity_gate.py,
bench.py): same pass as orc_raycast on the dense atlas, bit for bit; reads of slots that were not supplied are counted."""
import numpy as np
import pytest

from oracle import orc
from scene import Scene
from tuvok_b200 import synth


def _slots_from_atlas(st, brick):
    pool, atlas = st["pool"], st["atlas"]
    cap = pool.capacity
    slots = {}
    for v in st["meta"]:
        if v < orc.BI_FLAG_COUNT:
            continue
        s = int(v) - orc.BI_FLAG_COUNT
        sx, sy, sz = s % cap[0], (s // cap[0]) % cap[1], s // (cap[0] * cap[1])
        slots[s] = atlas[sz * brick:(sz + 1) * brick, sy * brick:(sy + 1) * brick, sx * brick:(sx + 1) * brick].copy().reshape(-1)
    return slots


@pytest.mark.parametrize("mode,lighting,dtype", [(orc.RM_1DTRANS, False, orc.U8), (orc.RM_2DTRANS, True, orc.U16),
                                                   (orc.RM_ISOSURFACE, False, orc.F32)])
def test_sparse_pool_pass_equals_dense_pass(mode, lighting, dtype):
    s = Scene(kind=synth.V_SPH, size=(56, 48, 40), dtype=dtype, brick=20, overlap=2, mode=mode, lighting=lighting,
              width=48, height=40, tf_center=0.3, tf_inv_gradient=0.3)
    st = s.oracle_render(threads=4)
    p = st["params"]
    zeros = np.zeros_like(st["entry"])
    dense, ds = orc.raycast(p, st["atlas"], st["meta"], st["tf"], st["entry"], zeros, st["exit"], st["covered"], None, 4)
    slots = _slots_from_atlas(st, 20)
    sparse, ss, absent = orc.raycast_slots(p, slots, st["meta"], st["tf"], st["entry"], zeros, st["exit"], st["covered"], 4)
    assert absent == 0 and ss.samples == ds.samples
    for a, b in zip(dense, sparse):
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32))
    # a strided coverage mask traces exactly those pixels
    mask = np.zeros((40, 48), np.uint8)
    mask[2::4, 2::4] = 1
    mask = mask.reshape(-1) & st["covered"]
    part, _, _ = orc.raycast_slots(p, slots, st["meta"], st["tf"], st["entry"], zeros, st["exit"], mask, 4)
    sel = np.flatnonzero(mask)
    assert sel.size and np.array_equal(part[0][sel].view(np.uint32), dense[0][sel].view(np.uint32))
    assert not part[0][np.flatnonzero(mask == 0)].any()
    # a missing slot is detected, not silently read as zeros
    victim = next(iter(k for k in slots if k != max(slots)))
    del slots[victim]
    _, _, absent = orc.raycast_slots(p, slots, st["meta"], st["tf"], st["entry"], zeros, st["exit"], st["covered"], 4)
    assert absent > 0
