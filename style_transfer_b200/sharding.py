"""Tile sharding across ranks: the replacement of the reference's fork()/queue/shared-memory tile
dispatcher (style_transfer.py:267-337, 614-645).

Tile ``i`` of the row-major grid goes to rank ``i % world`` -- the round-robin of
``TileWorkerPool.request`` (:284-288), restarted at worker 0 for every evaluation (:290-298).  Each
rank writes the gradients of its tiles into consecutive *slots* of a packed buffer
``[tiles_per_rank][3][tile_h_max][tile_w_max]``; one all-gather stitches the image gradient and one
all-reduce sums the loss (the reference's ``resp_q.get()`` loop, :639-643).  The functions here are
pure host logic (they run on CPU tensors with the gloo backend too); the CUDA side computes the
same geometry in ``st_tile_grid`` / ``st_eval_sc_grad_tiles`` / ``st_unpack_grad``.
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


def exchange(packed, loss, world, group=None):
    """All-gather of the packed gradient tiles + all-reduce(sum) of the loss.  Returns
    (packed_all [world, tiles_per_rank, 3, thmax, twmax], loss).  ``world == 1`` is a no-op."""
    if world == 1:
        return packed.reshape((1,) + tuple(packed.shape)), loss
    import torch
    import torch.distributed as dist
    shape = tuple(packed.shape)
    flat = torch.empty((world * shape[0],) + shape[1:], dtype=packed.dtype, device=packed.device)
    dist.all_gather_into_tensor(flat, packed.contiguous(), group=group)   # rank-major concatenation
    dist.all_reduce(loss, group=group)
    return flat.view((world,) + shape), loss


def unpack_numpy(packed_all, H, W, tile_size, roll_y=0, roll_x=0):
    """Host restatement of ``st_unpack_grad``: pastes the slots back (:642) and rolls the result
    back into the un-rolled frame (:805).  Used by the CPU tests of the sharding logic."""
    packed_all = np.asarray(packed_all)
    world = packed_all.shape[0]
    grad = np.zeros((3, H, W), dtype=packed_all.dtype)
    for t, (sy, sx, ey, ex) in enumerate(tile_boxes(H, W, tile_size)):
        grad[:, sy:ey, sx:ex] = packed_all[t % world, t // world, :, :ey - sy, :ex - sx]
    return np.roll(grad, (-roll_y, -roll_x), axis=(1, 2))
