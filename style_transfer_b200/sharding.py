"""Tile sharding across ranks: the replacement of the reference's fork()/queue/shared-memory tile
dispatcher (style_transfer.py:267-337, 614-645).

Tile ``i`` of the row-major grid goes to rank ``i % world`` -- the round-robin of
``TileWorkerPool.request`` (:284-288), restarted at worker 0 for every evaluation (:290-298).  Each
rank writes the gradients of its tiles into consecutive *slots* of its chunk of the exchange buffer,
``[tiles_per_rank][3][tile_h_max][tile_w_max]`` floats padded to a multiple of four plus a four-float
tail that carries the rank's loss as a double; ONE all-gather of the chunks stitches the image
gradient and delivers the losses (the reference's ``resp_q.get()`` loop, :639-643).  On the GPU the
collective is ``st_allgather_grad`` (ncclAllGather issued from the C ABI); the functions here are
pure host logic (they run on CPU tensors with the gloo backend too); the CUDA side computes the
same geometry in ``st_tile_grid`` / ``st_packed_floats`` / ``st_eval_sc_grad_tiles`` /
``st_unpack_grad``.
"""

import numpy as np


def tile_grid(H, W, tile_size):
    """(ntiles_y, ntiles_x, tile_h, tile_w, tile_h_max, tile_w_max) -- style_transfer.py:619-631:
    ``ntiles = (size-1)//tile_size + 1`` tiles of ``size//ntiles``, the last row/column absorbing
    the remainder."""
    nty, ntx = (H - 1) // tile_size + 1, (W - 1) // tile_size + 1
    th, tw = H // nty, W // ntx
    return nty, ntx, th, tw, H - (nty - 1) * th, W - (ntx - 1) * tw


def tile_boxes(H, W, tile_size):
    """[(start_y, start_x, end_y, end_x)] in row-major (request) order."""
    nty, ntx, th, tw, _, _ = tile_grid(H, W, tile_size)
    boxes = []
    for y in range(nty):
        for x in range(ntx):
            boxes.append((y * th, x * tw, H if y == nty - 1 else (y + 1) * th,
                          W if x == ntx - 1 else (x + 1) * tw))
    return boxes


def local_tiles(H, W, tile_size, rank, world):
    """[(slot, box)] of the tiles rank ``rank`` evaluates."""
    boxes = tile_boxes(H, W, tile_size)
    return [(slot, boxes[t]) for slot, t in enumerate(range(rank, len(boxes), world))]


def packed_shape(H, W, tile_size, world):
    nty, ntx, _, _, thmax, twmax = tile_grid(H, W, tile_size)
    per_rank = (nty * ntx + world - 1) // world
    return (per_rank, 3, thmax, twmax)


def packed_floats(H, W, tile_size, world):
    """Floats of one rank's chunk (``st_packed_floats``): tiles, padded to 4, + the 4-float tail."""
    tiles = int(np.prod(packed_shape(H, W, tile_size, world)))
    return (tiles + 3) // 4 * 4 + 4


def tiles_view(chunk, H, W, tile_size, world):
    """[tiles_per_rank, 3, thmax, twmax] view of the tile part of one chunk (numpy or torch)."""
    shape = packed_shape(H, W, tile_size, world)
    return chunk[:int(np.prod(shape))].reshape(shape)


def loss_view(chunk):
    """The loss (one float64) in the tail of a chunk: numpy array or torch tensor view."""
    return chunk[-4:-2].view(np.float64) if isinstance(chunk, np.ndarray) else \
        chunk[-4:-2].view(__import__('torch').float64)


def exchange(chunk, world, group=None):
    """All-gather of the ranks' chunks through ``torch.distributed`` (any backend: the CPU tests run
    it over gloo).  Returns [world, floats_per_rank]; ``world == 1`` is a no-op."""
    if world == 1:
        return chunk.reshape((1,) + tuple(chunk.shape))
    import torch
    import torch.distributed as dist
    flat = torch.empty(world * chunk.numel(), dtype=chunk.dtype, device=chunk.device)
    dist.all_gather_into_tensor(flat, chunk.contiguous().view(-1), group=group)   # rank-major
    return flat.view((world,) + tuple(chunk.shape))


def unpack_numpy(packed_all, H, W, tile_size, roll_y=0, roll_x=0):
    """Host restatement of ``st_unpack_grad``: pastes the slots back (:642), rolls the result back
    into the un-rolled frame (:805) and sums the ranks' losses.  ``packed_all``: [world, floats]."""
    packed_all = np.asarray(packed_all)
    world = packed_all.shape[0]
    grad = np.zeros((3, H, W), dtype=packed_all.dtype)
    tiles = [tiles_view(packed_all[r], H, W, tile_size, world) for r in range(world)]
    for t, (sy, sx, ey, ex) in enumerate(tile_boxes(H, W, tile_size)):
        grad[:, sy:ey, sx:ex] = tiles[t % world][t // world, :, :ey - sy, :ex - sx]
    loss = sum(float(loss_view(np.ascontiguousarray(packed_all[r]))[0]) for r in range(world))
    return np.roll(grad, (-roll_y, -roll_x), axis=(1, 2)), loss
