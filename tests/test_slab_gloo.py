"""CPU, world_size 2 and 4 over gloo: the slab-decomposed pipeline (product class
dedalus/data_objects/slab.py + the C-ABI phases compiled for host emulation) reproduces the
reference goldens rank by rank."""
import json
import os
import socket

import pytest
import torch.multiprocessing as mp

import slab_worker


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("world,case,layout", [(2, "mhd3d_16_rk2mid", 0), (4, "mhd3d_16_rk2mid", 0), (2, "bouss3d_16_rk2mid", 0),
                                               (2, "hydro3d_16_rk2mid", 1), (4, "mhd3d_8x16x32_rk2trap", 0),
                                               (4, "mhd3d_16_rk2mid", 1), (2, "mhd3d_8x16x32_rk2trap", 1),
                                               (8, "mhd3d_16_rk2mid", 0),      # block slabs: two ranks own NO retained ky row
                                               (8, "mhd3d_16_rk2mid", 1),
                                               # grids that are not powers of two (runtime-length kernels): 12 x 20 x 24
                                               (2, "hydro3d_12x20x24_rk2mid", 0), (4, "hydro3d_12x20x24_rk2mid", 1)])
def test_slab_pipeline_matches_reference(tmp_path, world, case, layout):
    """layout 0: the reference's block ky slabs; 1: cyclic ky ownership (balanced under dealiasing)."""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "host"))
    import emul
    emul.load()          # build the emulation library once, before the ranks race for it
    out = str(tmp_path / "res.json")
    mp.spawn(slab_worker.worker, args=(world, _free_port(), case, out, layout), nprocs=world, join=True)
    res = json.load(open(out))
    assert len(res) == world
    for r in res:
        assert r["bwd"] < 1e-13 and r["fwd"] < 1e-13 and r["bwd_dealias"] == 0.0, r
        assert r["rhs"] < 1e-12 and r["state_after"] < 1e-13, r
        assert sum(r["rows"]) > 0
    if world == 8 and layout == 0:
        assert 0 in res[0]["rows"]            # the empty-rank edge case really occurred
