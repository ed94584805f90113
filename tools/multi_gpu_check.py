"""Run under torchrun with N >= 2 GPUs: the N-rank evaluation (tiles dealt round-robin, ONE NCCL
all-gather of the ranks' chunks -- gradient tiles + loss -- issued from the C ABI by
st_allgather_grad) must equal the single-rank evaluation of the same image bit for bit (tiles are
independent and the gather is a permutation).  Also checks the host-image path (every rank uploads its
slab, NVLink all-gather) against the resident image.  tests/test_gpu_multi.py runs this under
torch.distributed.run when the machine has two GPUs."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

from style_transfer_b200 import netdesc, weights
from style_transfer_b200.engine import TileEngine
from style_transfer_b200.transfer import StyleTransfer, default_args


def main():
    rank, world, local = (int(os.environ[k]) for k in ('RANK', 'WORLD_SIZE', 'LOCAL_RANK'))
    torch.cuda.set_device(local)
    dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    size, tile = int(os.environ.get('CHECK_SIZE', '1024')), int(os.environ.get('CHECK_TILE', '256'))
    args = default_args(size=size, min_size=size, tile_size=tile)
    net = netdesc.from_model(args.model)
    params = weights.he_normal(net)
    rs = np.random.RandomState(3)
    content, style = (rs.randint(0, 256, (size, size, 3)).astype(np.uint8) for _ in range(2))
    results = []
    for r, w in ((rank, world), (0, 1)):
        eng = TileEngine(net, params, mean=args.mean, device=local, precision='fp16', rank=r, world=w)
        eng.init_comm()
        assert w == 1 or eng._comm_ready
        st = StyleTransfer(eng, args)
        np.random.seed(0)
        st.init_first_scale(size, size)
        st.prepare([eng.pil_to_image(content)], [eng.pil_to_image(style)])
        for i in range(3):
            if i == 2:                    # last step: the image comes from pinned host memory
                host = torch.empty(eng.img.shape, dtype=torch.float32).pin_memory()
                host.copy_(eng.img)
                torch.cuda.synchronize()
                eng.img.zero_()           # the resident copy must not be what gets evaluated
                eng.stage_host_image(host)
            avg, loss = st.step()
        torch.cuda.synchronize()
        results.append((avg.clone(), float(loss)))
    same = torch.equal(results[0][0], results[1][0])
    dl = abs(results[0][1] - results[1][1]) / abs(results[1][1])
    print('rank %d/%d: image after 3 steps identical to 1-rank run: %s, loss rel diff %.2e' %
          (rank, world, same, dl), flush=True)
    ok = torch.tensor([int(same and dl < 1e-9)], device='cuda')
    dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    dist.destroy_process_group()
    sys.exit(0 if int(ok.item()) == 1 else 1)


if __name__ == '__main__':
    main()
