"""CPU-side check of the claim behind the temporal-coherence mode: with persistent
fat boxes the candidate set is a superset of the tight one at every step, and the contact set is the
same, because the narrowphase works on the exact shapes.  The CUDA path is compared with this stateful
restatement in tests/test_widen_gpu.py::test_temporal_coherence_fat_boxes_and_pair_cache."""
import numpy as np

import axcd
import oracle_lib as O


def _expand(tight, margin):
    m = np.float32(margin)
    return np.concatenate([tight[:, :3] - m, tight[:, 3:] + m], axis=1).astype(np.float32)


def test_fat_boxes_keep_the_contact_set_while_bodies_drift():
    s = axcd.config_scene("C1", scale=0.05)
    rng = np.random.default_rng(4)
    xf = s.xf.copy()
    margin = 0.04
    fat = None
    rebuilt_total = 0
    for step in range(6):
        xf[:, :3] += rng.normal(size=(s.n, 3)).astype(np.float32) * 0.02      # slow drift
        rc, tight = O.refit(xf, s.shapes, s.hull)
        if fat is None:
            fat = _expand(tight, margin)
        else:
            inside = (fat[:, :3] <= tight[:, :3]).all(1) & (tight[:, 3:] <= fat[:, 3:]).all(1)
            fat[~inside] = _expand(tight[~inside], margin)
            rebuilt_total += int((~inside).sum())
            assert (~inside).sum() < s.n // 2          # most boxes survive a step
        # every tight box lies inside its fat box
        assert (fat[:, :3] <= tight[:, :3]).all() and (tight[:, 3:] <= fat[:, 3:]).all()
        p_tight = O.broadphase(tight)
        p_fat = O.broadphase(fat)
        st, sf = set(map(tuple, p_tight)), set(map(tuple, p_fat))
        assert st <= sf and len(sf) > len(st)
        c_tight, _, _ = O.narrowphase(xf, s.shapes, p_tight, s.hull)
        c_fat, _, _ = O.narrowphase(xf, s.shapes, p_fat, s.hull)
        assert c_tight.tobytes() == c_fat.tobytes()      # same contacts, bit for bit
    assert rebuilt_total > 0
