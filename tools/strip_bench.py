"""Strong scaling of ONE frame over row strips (BASELINE config 4 style: 7680x4320 4:2:0), run under
torchrun.  Prints one line per run from rank 0.  Measurement aid for profiles/ — bench.py stays the
headline (independent frames)."""
import argparse
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from jpeg2png_b200 import abi, strips, synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--width', type=int, default=7680)
ap.add_argument('--height', type=int, default=4320)
ap.add_argument('--subsampling', default='4:2:0')
ap.add_argument('--quality', type=int, default=10)
ap.add_argument('--iterations', type=int, default=100)
args = ap.parse_args()

rank, world, local = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1)), int(os.environ.get('LOCAL_RANK', 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group('nccl', device_id=torch.device('cuda', local))
lib = abi.load_product()
# a quarter-size cartoon tiled 4x4 at block level: same statistics, a fraction of the host time
base = synth.synth_coefs(-(-args.width // 64) * 16, -(-args.height // 64) * 16, args.quality, args.subsampling, seed=1238)
img = synth.tile_coefs(base, 4, 4, args.width, args.height)
mcu = 8 * max(p.h_samp for p in img.planes)
row0, rows = strips.plan_strips(img.frame_h, mcu, world)[rank]
be = strips.ProductStrip(lib, img, 0.3, [0.001] * 3, args.iterations, row0, rows, local)


comm = strips.native_comm(be, dist, rank, world)


def run(mode, n):
    if mode == 'native':
        strips.solve_strips_native(be, comm, n)       # NCCL queued by the library on the session stream
    else:
        strips.solve_strips(be, dist, rank, world, n)  # torch.distributed drives both exchanges


for mode in ('native', 'torchdist'):
    lib.j2p_session_reset(be.s)
    run(mode, 5)                         # warm-up (also NCCL connection setup)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier(device_ids=[local])
    lib.j2p_session_reset(be.s)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record(be.stream)
    run(mode, args.iterations)
    e1.record(be.stream)
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    t = torch.tensor([e0.elapsed_time(e1) * 1e-3, wall], dtype=torch.float64, device='cuda')
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        pix = args.width * args.height * args.iterations
        dev, wl = t[0].item(), t[1].item()
        print(f'strips[{mode}] N={world} {args.width}x{args.height} {args.subsampling} {args.iterations} it: device {dev*1e3:.1f} ms '
              f'(wall {wl*1e3:.1f} ms)  {pix / dev / 1e6:.0f} Mpix-it/s  ({dev / args.iterations * 1e6:.0f} us/iteration)', flush=True)
        if mode == 'native':
            print(f'strips[native] protocol: {"peer memory" if lib.j2p_comm_protocol(comm) else "NCCL"}, status {lib.j2p_comm_status(comm)}', flush=True)
        print(f'strips[{mode}] N={world} checksum {float(np.float64(be.download(0)).sum()):.4f}', flush=True)
lib.j2p_comm_destroy(comm)
be.close()
if world > 1:
    dist.destroy_process_group()
