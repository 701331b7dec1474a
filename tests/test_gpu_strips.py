"""Row strips of one frame over several B200s (pytest -m gpu on a box with >= 2 GPUs): the product's
strip sessions + NCCL, against the single-GPU product result (which the parity tests tie to the
reference).  Skipped on a single-GPU box.  With J2P_STRIP_P2P=1 in the environment the 'native'
cases run the peer-memory protocol instead of NCCL inside the loop (same assertions)."""
import os
import tempfile

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

CASE = dict(w=640, h=512, q=20, ss='4:2:0', weight=0.3, pw=[0.001] * 3, iters=25)


def _worker(rank, world, init_file, out_dir, native):
    import torch
    import torch.distributed as dist
    from jpeg2png_b200 import abi, strips, synth
    torch.cuda.set_device(rank)
    dist.init_process_group('nccl', init_method=f'file://{init_file}', rank=rank, world_size=world,
                            device_id=torch.device('cuda', rank))
    try:
        lib = abi.load_product()
        img = synth.synth_coefs(CASE['w'], CASE['h'], CASE['q'], CASE['ss'], seed=777)
        mcu = 8 * max(p.h_samp for p in img.planes)
        row0, rows = strips.plan_strips(img.frame_h, mcu, world)[rank]
        be = strips.ProductStrip(lib, img, CASE['weight'], CASE['pw'], CASE['iters'], row0, rows, rank)
        if native:
            # the library's own loop: NCCL all-gather + halo send/recv queued on the session stream,
            # in two calls to cover the continuation path
            comm = strips.native_comm(be, dist, rank, world)
            strips.solve_strips_native(be, comm, 10)
            strips.solve_strips_native(be, comm, CASE['iters'] - 10)
        else:
            strips.solve_strips(be, dist, rank, world, CASE['iters'])
        np.savez(os.path.join(out_dir, f'rank{rank}.npz'), rows=rows, **{f'p{c}': be.download(c) for c in range(3)})
        if native:
            assert lib.j2p_comm_status(comm) == 0, lib.j2p_last_error().decode()
            if os.environ.get('J2P_STRIP_P2P') == '1' and world <= 8:
                assert lib.j2p_comm_protocol(comm) == 1, 'the peer-memory protocol was requested but the ranks fell back to NCCL'
            lib.j2p_comm_destroy(comm)
        be.close()
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('native', [False, True], ids=['torchdist', 'native'])
@pytest.mark.parametrize('world', [2, 4, 8])
def test_strips_match_single_gpu(world, native, tmp_path):
    import torch
    import torch.multiprocessing as mp
    from jpeg2png_b200 import synth
    from tests import helpers as H
    if torch.cuda.device_count() < world:
        pytest.skip(f'needs {world} GPUs')
    init_file = tempfile.mktemp(dir=str(tmp_path))
    mp.spawn(_worker, args=(world, init_file, str(tmp_path), native), nprocs=world, join=True)
    img = synth.synth_coefs(CASE['w'], CASE['h'], CASE['q'], CASE['ss'], seed=777)
    want = H.run_compute('product', img, [0, 1, 2], CASE['weight'], CASE['pw'], CASE['iters'])
    parts = [np.load(os.path.join(str(tmp_path), f'rank{r}.npz')) for r in range(world)]
    got = [np.concatenate([p[f'p{c}'] for p in parts], axis=0) for c in range(3)]
    H.assert_bit_identical(got, want, f'{world} strips vs one GPU')
