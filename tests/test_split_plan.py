"""The launch planner of the section-split biquad kernel (K1s, zg_split_plan_query): what its persistent CTAs rely on,
checked over thousands of shapes on the host.  A violated invariant would not be a wrong sample but a launch that never
ends (a group of warps waiting for a row that no other group hands over), so it is checked where no GPU is needed."""
import itertools

import pytest


def _check(zg, sections, C, T, exact, sm, smem, segs=1, warm=0):
    g = zg.split_plan(sections, C, T, exact=exact, sm_count=sm, max_smem=smem, segments=segs, warmup_samples=warm)
    if g is None:
        return None
    n_cg = (C + 31) // 32
    rows = n_cg * g.segments
    slots = g.grid * g.groups_per_cta
    assert T % 32 == 0 and T >= 128
    assert 1 <= g.grid <= sm
    assert g.threads_per_cta == g.groups_per_cta * g.warps_per_group * 32 <= 512
    assert g.warps_per_group * g.sections_per_warp == sections
    # every group's range of the tile sequence is at least one row long: a row is handed over at most once
    assert slots <= rows, (sections, C, T, g.grid, g.groups_per_cta, rows)
    assert 2 <= g.stages <= 6 and g.boxes_per_tile >= 1
    assert g.boxes_per_handover in (1, 2, 4)
    assert g.boxes_per_tile % g.boxes_per_handover == 0 and (g.boxes_per_handover == 1 or g.boxes_per_tile >= 2 * g.boxes_per_handover)
    assert g.boxes_per_handover == 1 or (g.groups_per_cta == 1 and sections == 4)
    # the ring (+ barriers, alignment slack) fits
    assert g.groups_per_cta * g.stages * g.boxes_per_tile * 4096 < g.smem_bytes <= smem
    if g.segments > 1:
        assert not exact and sections == 4
        assert g.segment_boxes % g.boxes_per_tile == 0 and g.warmup_boxes % g.boxes_per_tile == 0
        assert g.warmup_boxes * 32 >= warm and g.segment_boxes >= 2 * g.warmup_boxes > 0
        # the last segment has samples of its own
        assert (g.segments - 1) * g.segment_boxes + g.warmup_boxes < T // 32
        assert g.segments * g.segment_boxes + g.warmup_boxes >= T // 32
    else:
        assert g.segment_boxes == 0 and g.warmup_boxes == 0
    return g


def test_planner_invariants_over_many_shapes(zg):
    chans = [1, 31, 32, 33, 96, 1000, 4096, 4737, 6400, 9471, 9472, 12800, 16384, 28415, 28416, 65536, 131072, 1 << 20]
    samples = [128, 256, 4096, 4128, 8192, 65536, 1 << 20]
    n = 0
    for sections, C, T, exact in itertools.product(range(2, 9), chans, samples, (True, False)):
        for sm, smem in ((148, 232448), (132, 232448), (8, 101376)):
            if _check(zg, sections, C, T, exact, sm, smem) is not None:
                n += 1
    assert n > 1000


def test_planner_invariants_cut_in_time(zg):
    n = 0
    for C, T, segs, warm in itertools.product([32, 96, 4096, 8192, 20000], [4096, 8192, 65536, 65536 + 96, 1 << 20],
                                              [2, 4, 8, 12, 33], [128, 640, 2048, 8192]):
        if _check(zg, 4, C, T, False, 148, 232448, segs, warm) is not None:
            n += 1
    assert n > 100


def test_planner_choices_on_the_baseline_shapes(zg):
    ns = zg.split_plan(4, 65536, 8192)                       # north star: three groups per CTA on every SM, 1 KB runs
    assert (ns.grid, ns.groups_per_cta, ns.warps_per_group, ns.stages, ns.boxes_per_tile, ns.boxes_per_handover) == (148, 3, 4, 2, 8, 1)
    c2 = zg.split_plan(4, 4096, 65536)                       # configs[1] EXACT: one group per SM, nothing is cut
    assert (c2.grid, c2.groups_per_cta, c2.stages, c2.boxes_per_tile, c2.boxes_per_handover) == (128, 1, 3, 16, 4)
    mid = zg.split_plan(4, 8192, 32768)                      # two groups on 128 SMs: whole rows
    assert (mid.grid, mid.groups_per_cta) == (128, 2)
    cut = zg.split_plan(4, 4096, 65536, exact=False, segments=8, warmup_samples=640)    # configs[1] FAST, cut in time
    assert cut.segments >= 4 and cut.warmup_boxes * 32 >= 640 and cut.groups_per_cta == 3 and cut.grid == 148
    assert zg.split_plan(4, 65536, 8200) is None             # ragged blocks stay on the lane-per-channel kernel
    assert zg.split_plan(2, 65536, 8192) is None             # two sections: K1
    assert zg.split_plan(8, 4096, 65536) is None             # few channels, other section counts: K1 / K1b
    assert zg.split_plan(8, 16384, 16384).groups_per_cta >= 2
