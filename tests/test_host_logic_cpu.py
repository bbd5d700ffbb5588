

def test_owned_rows_cover_every_window_once():
    """Row bands of the sharded host tile path (tile.py: owned_rows): the ranks' window blocks are disjoint, in reference order, and complete."""
    from sifnn_b200.tile import owned_rows, window_list
    for (ht, wt, world) in [(1200, 1200, 1), (1200, 1200, 8), (1200, 1200, 5), (336, 256, 3), (64, 64, 4), (128, 640, 7)]:
        wy_all, wx_all = window_list(ht, wt)
        seen = []
        for r in range(world):
            y0, y1, wy, wx = owned_rows(ht, wt, r, world)
            assert wy.numel() == wx.numel()
            if wy.numel() == 0:
                assert y0 == y1
                continue
            assert int(wy.min()) == 0 and int(wy.max()) == y1 - y0 - 1
            seen += [(int(a) + y0, int(b)) for a, b in zip(wy, wx)]
        assert seen == [(int(a), int(b)) for a, b in zip(wy_all, wx_all)]
