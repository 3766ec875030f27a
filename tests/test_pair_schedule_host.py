"""Host logic of the CTA-pair convolution kernel (no GPU): the work lists conv_pair.cu hands to the pairs
(sr_object_detection_b200/csrc/cuda/conv_pair.cu, pair_schedule).  Every (position tile, 64-filter unit) must be
owned by exactly one piece, pieces are 64..256 filters wide inside one position tile, and the balanced cut never
costs more than whole tiles round-robin under the kernel's own cost table."""
import ctypes as C

import numpy as np
import pytest

from sr_object_detection_b200 import _lib

COST = [0.0, 2.97, 3.28, 3.6, 4.0]  # pair_piece_costs() defaults (measured)


def _schedule(rows, units, pairs, balanced):
    lib = _lib.load()
    lib.y2_pair_schedule.restype = C.c_int
    cap = rows * units + 8 * pairs + 64
    out = np.zeros((cap, 4), np.int32)
    stride = C.c_int()
    n = lib.y2_pair_schedule(rows, units, pairs, balanced, out.ctypes.data_as(C.POINTER(C.c_int)), cap, C.byref(stride))
    assert n == pairs * stride.value, n
    return out[:n].reshape(pairs, stride.value, 4)


def _check(rows, units, work):
    cover = np.zeros((rows, units), np.int32)
    costs = []
    for lst in work:
        cost, ended = 0.0, False
        for m, n0, nc, _ in lst:
            if nc == 0:
                ended = True
                continue
            assert not ended, "entries after the terminator"
            assert n0 % 64 == 0 and nc in (64, 128, 192, 256) and 0 <= m < rows and n0 + nc <= units * 64
            cover[m, n0 // 64:(n0 + nc) // 64] += 1
            cost += COST[nc // 64]
        assert ended, "every list ends with a terminator"
        costs.append(cost)
    assert (cover == 1).all(), "every unit belongs to exactly one piece"
    return max(costs)


# (position tiles, 64-filter units, pairs): yolo-voc 13x13 b64 with 1024 filters, 52x52 with 256, 26x26 with 512,
# yolo 608 19x19 b32, a 1280-filter layer, fewer tiles than pairs, few pairs
SHAPES = [(49, 16, 74), (703, 4, 74), (183, 8, 74), (50, 16, 74), (49, 20, 74), (5, 4, 5), (100, 4, 3), (1, 4, 1),
          (37, 12, 74), (3000, 4, 74)]


@pytest.mark.parametrize("rows,units,pairs", SHAPES)
def test_work_lists_cover_every_unit_once(rows, units, pairs):
    rr = _check(rows, units, _schedule(rows, units, pairs, 0))
    bal = _check(rows, units, _schedule(rows, units, pairs, 1))
    assert rr == 4.0 * -(-(rows * (units // 4)) // pairs)   # whole tiles: ceil(tiles / pairs) each
    assert bal <= rr + 1e-6
    if (rows, units, pairs) == (49, 16, 74):
        assert bal < 0.98 * rr   # the yolo-voc 13x13 layers: 196 tiles on 74 pairs no longer run as 3 full waves


def test_rejects_bad_arguments():
    lib = _lib.load()
    lib.y2_pair_schedule.restype = C.c_int
    out = (C.c_int * 16)()
    stride = C.c_int()
    assert lib.y2_pair_schedule(4, 6, 2, 1, out, 4, C.byref(stride)) < 0     # units not a multiple of 4
    assert lib.y2_pair_schedule(100, 4, 2, 1, out, 4, C.byref(stride)) < 0   # capacity too small
