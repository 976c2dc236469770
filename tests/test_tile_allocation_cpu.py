"""CPU: host logic of the tile allocation mirror (tile_allocation.py; preprocess/build_tiles.py:94-245 of the
reference): tile lattice, tile / view selection rules, file formats."""
import os
import tempfile

import torch

from conftest import load_pkg


def test_tile_grid_lattice():
    load_pkg()
    import tile_allocation as ta
    corners = ta.tile_grid([0.0, 0.0, 0.0, 50.0, 9.0, 30.0], [20.0, 10.0, 20.0], 0.2, offset=(1.0, 0.0, -1.0), max_num_tile=(100, 1, 100))
    # ceil((50 - 1) / 20) = 3, 1 (capped), ceil(31 / 20) = 2 tiles; stride = 0.8 * size
    assert corners.shape == (6, 3)
    assert torch.allclose(corners[0], torch.tensor([1.0, 0.0, -1.0]))
    assert torch.allclose(corners[-1], torch.tensor([1.0 + 2 * 16.0, 0.0, -1.0 + 16.0]))


def test_selection_rules_and_files():
    load_pkg()
    import tile_allocation as ta
    tile_size = [10.0, 10.0, 10.0]
    corners = torch.tensor([[0.0, 0, 0], [10.0, 0, 0], [20.0, 0, 0], [30.0, 0, 0]])
    cams = torch.tensor([[5.0, 5, 5], [6.0, 5, 5], [15.0, 5, 5], [16.0, 5, 5], [17.0, 5, 5], [100.0, 5, 5]])
    related = torch.tensor([[0.9, 0.8, 0.05, 0.0, 0.0, 0.0],
                            [0.3, 0.02, 0.7, 0.6, 0.5, 0.0],
                            [0.0, 0.0, 0.4, 0.3, 0.2, 0.0],
                            [0.0, 0.0, 0.0, 0.0, 0.0, 0.0]])
    # tiles 0 and 1 hold cameras; expect 3 -> the closest empty tile (2) is added; tile 3 never
    kept, views = ta.select_tiles_and_views(related, cams, corners, tile_size, expect_num=3, min_num_image=1, scene_type="outdoor", thresh=0.1)
    assert kept == [0, 1, 2]
    assert views[0] == [0, 1]                            # 0.1 * inside + related: 1.0, 0.9 | 0.05 and below fall under the threshold
    assert views[1] == [2, 3, 4, 0]                      # 0.8, 0.7, 0.6, 0.3
    assert views[2] == [2, 3, 4]
    # a tile with too few views is dropped and the positions close up
    kept2, views2 = ta.select_tiles_and_views(related, cams, corners, tile_size, expect_num=3, min_num_image=2, thresh=0.1)
    assert kept2 == [1, 2] and views2[0] == [2, 3, 4, 0] and views2[1] == [2, 3, 4]
    # ignored cameras and the indoor rule (no bonus for cameras inside the tile)
    kept3, views3 = ta.select_tiles_and_views(related, cams, corners, tile_size, expect_num=2, min_num_image=0, scene_type="indoor", ignore=[2])
    assert kept3 == [0, 1] and views3[1] == [3, 4, 0]
    out = tempfile.mkdtemp()
    ta.write_tile_files(out, corners, tile_size, kept, views)
    lines = open(os.path.join(out, "training_views.txt")).read().split("\n")
    assert lines[:6] == ["0", "0 1", "1", "2 3 4 0", "2", "2 3 4"]
    info = open(os.path.join(out, "tile_info.txt")).read().split("\n")
    assert info[0].startswith("# TILEID(1)") and info[2] == "1 10.00 0.00 0.00 10.00 10.00 10.00 32 8192 0"
