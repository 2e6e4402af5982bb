"""World-size-2 `gloo` tests of the multi-GPU host logic (CPU only).

The CUDA kernels cannot run here, so each rank stands in for its GPU with the CPU oracle restricted to the tiles
it owns (tile % world == rank, the rule the kernels use) and the real `triangle_splatting_b200.distributed`
code assembles the frame and reduces the gradients.  Checks: ownership is a partition of the tiles; the
assembled forward outputs equal the single-rank render; the all-reduced partial gradients equal the full ones."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import harness  # noqa: F401
from triangle_splatting_b200 import distributed as tsd
from triangle_splatting_b200.scenes import make_scene

WORLD = 2


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _scene():
    return make_scene("d", 600, 80, 64, sh_degree=1, rich_info=True, geometry_grads=True, seed=9, rho_px=5.0)


def _oracle_shard(sc, rank, world):
    from oracle.oracle import Oracle

    o = Oracle("f64")
    kw = sc.settings_kwargs()
    kw.pop("debug")
    kw = {k: (v.numpy() if isinstance(v, torch.Tensor) else v) for k, v in kw.items()}
    st = o.forward(**kw, vertex=sc.vertex.numpy(), shs=sc.shs.numpy(), feature=None, opacity=sc.opacity.numpy(), tile_step=world,
                   tile_offset=rank)
    g = o.backward(st, sc.grads["dL_dout_feature"].numpy(), sc.grads["dL_dout_depth"].numpy(), sc.grads["dL_dout_normal"].numpy())
    return st, g


def _worker(rank, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(WORLD))
    dist.init_process_group("gloo", rank=rank, world_size=WORLD)
    try:
        assert tsd.current_shard() == (0, 1)
        r, w = tsd.enable_tile_sharding()
        assert (r, w) == (rank, WORLD) and tsd.current_shard() == (rank, WORLD)
        sc = _scene()
        st, g = _oracle_shard(sc, rank, WORLD)
        # un-owned tiles must be zero before assembly (the CUDA path allocates zero-filled planes when sharded)
        gx = (sc.cam["image_width"] + 15) // 16
        own = np.zeros((sc.cam["image_height"], sc.cam["image_width"]), bool)
        for t in tsd.owned_tiles(gx * ((sc.cam["image_height"] + 15) // 16), rank, WORLD):
            ty, tx = divmod(t, gx)
            own[ty * 16:(ty + 1) * 16, tx * 16:(tx + 1) * 16] = True
        assert np.all(st["out_feature"][:, ~own] == 0)
        img, dep, nrm = (torch.from_numpy(st[k].copy()) for k in ("out_feature", "out_depth", "out_normal"))
        cs, cm = torch.from_numpy(st["contrib_sum"].copy()), torch.from_numpy(st["contrib_max"].copy())
        tsd.assemble_forward(img, dep, nrm, cs, cm)
        # the sharded CUDA forward carves image | depth | normal | contrib_sum back to back out of one buffer so that the frame
        # is assembled by ONE in-place all-reduce: same answer as the packed path above
        parts = [torch.from_numpy(st[k].copy()) for k in ("out_feature", "out_depth", "out_normal", "contrib_sum")]
        flat = torch.cat([t.reshape(-1) for t in parts])
        views, off = [], 0
        for t in parts:
            views.append(flat[off:off + t.numel()].view(t.shape))
            off += t.numel()
        tsd.assemble_forward(*views, torch.from_numpy(st["contrib_max"].copy()))
        for v, full in zip(views, (img, dep, nrm, cs)):
            assert torch.equal(v, full)
        grads = [torch.from_numpy(g[k].copy()) for k in ("dL_dvertex", "dL_dcenter2D", "dL_dshs", "dL_dfeature", "dL_dopacity")]
        tsd.reduce_gradients(*grads)
        if rank == 0:
            ret.update(img=img.numpy(), dep=dep.numpy(), nrm=nrm.numpy(), cs=cs.numpy(), cm=cm.numpy(), grads=[t.numpy() for t in grads])
        tsd.disable_tile_sharding()
        assert tsd.current_shard() == (0, 1)
    finally:
        dist.destroy_process_group()


def test_tile_ownership_is_a_partition():
    for n_tiles in (1, 7, 8160):
        for world in (1, 2, 3, 8):
            seen = []
            for r in range(world):
                seen += list(tsd.owned_tiles(n_tiles, r, world))
            assert sorted(seen) == list(range(n_tiles))


def test_sharded_render_assembles_to_the_full_frame():
    mgr = mp.Manager()
    ret = mgr.dict()
    port = _free_port()
    mp.spawn(_worker, args=(port, ret), nprocs=WORLD, join=True)
    sc = _scene()
    st, g = _oracle_shard(sc, 0, 1)  # the full, unsharded render
    assert np.allclose(ret["img"], st["out_feature"], rtol=0, atol=1e-12)
    assert np.allclose(ret["dep"], st["out_depth"], rtol=0, atol=1e-9)
    assert np.allclose(ret["nrm"], st["out_normal"], rtol=0, atol=1e-12)
    assert np.allclose(ret["cs"], st["contrib_sum"], rtol=1e-12, atol=1e-12)
    assert np.array_equal(ret["cm"], st["contrib_max"])
    for got, k in zip(ret["grads"], ("dL_dvertex", "dL_dcenter2D", "dL_dshs", "dL_dfeature", "dL_dopacity")):
        # K9 is linear in the per-triangle screen-space sums, so per-rank K9 + all-reduce == K9 of the total
        assert np.allclose(got, g[k], rtol=1e-9, atol=1e-14), k


def test_enable_requires_initialised_process_group():
    with pytest.raises(RuntimeError, match="not initialised"):
        tsd.enable_tile_sharding()
