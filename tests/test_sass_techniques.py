"""The compiled product must contain the Blackwell instructions DESIGN.md says the kernels are built
on (CPU test: cuobjdump over the in-tree libjpeg2png_b200.so; no GPU needed).  A refactor that
silently falls back to scalar fp32, register-held prefetch or plain launches fails here before it
reaches a GPU box.  `tools/sass_summary.py` prints the same table for profiles/."""
import collections
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, 'jpeg2png_b200', 'csrc', 'libjpeg2png_b200.so')


def sass_by_kernel():
    cuobjdump = shutil.which('cuobjdump') or '/usr/local/cuda/bin/cuobjdump'
    if not os.path.exists(cuobjdump) or not os.path.exists(LIB):
        pytest.skip('CUDA toolkit or the built library is missing')
    txt = subprocess.run([cuobjdump, '-sass', LIB], check=True, capture_output=True, text=True).stdout
    assert 'sm_100a' in txt or 'SM100' in txt.upper() or 'EF_CUDA_SM100' in txt, 'no sm_100a code in the library?'
    out = {}
    for fun in re.split(r'\n\s+Function : ', txt)[1:]:
        name = fun.split('\n')[0].strip()
        ops = collections.Counter()
        for line in fun.split('\n'):
            m = re.match(r'\s*/\*[0-9a-f]{4,5}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_]+)', line)
            if m:
                ops[m.group(1)] += 1
        out[name] = ops
    return out


@pytest.fixture(scope='module')
def sass():
    return sass_by_kernel()


def _kernels(sass, fragment):
    ks = {n: o for n, o in sass.items() if fragment in n}
    assert ks, f'no kernel matching {fragment} in the library'
    return ks


def test_gradient_kernel_runs_on_packed_fp32_with_an_async_row_ring(sass):
    # every instantiation of k_gradient_packed: FADD2 / FMUL2 / FFMA2 carry the arithmetic, cp.async
    # (LDGSTS) feeds the per-warp row ring, and the kernel takes part in the programmatic launch chain
    for name, ops in _kernels(sass, 'k_gradient_packed').items():
        assert ops['FFMA2'] > 30 and ops['FMUL2'] > 20 and ops['FADD2'] > 10, (name, dict(ops))    # the one-channel TV-only build is the smallest
        assert ops['LDGSTS'] >= 6, name
        assert ops['ACQBULK'] >= 1 and ops['PREEXIT'] >= 1, f'{name}: griddepcontrol.wait / launch_dependents missing'


def test_projection_tile_kernels_stage_with_cp_async_and_join_the_launch_chain(sass):
    for frag in ('k_project_tileILb', 'k_project_tile22'):
        for name, ops in _kernels(sass, frag).items():
            assert ops['LDGSTS'] >= 6, name
            assert ops['ACQBULK'] >= 1 and ops['PREEXIT'] >= 1, name
            assert ops['F2F'] >= 100 and ops['DMUL'] >= 100, f'{name}: the fp64-promoted transforms are gone?'


def test_tma_projection_kernel_uses_the_tensor_memory_accelerator(sass):
    for name, ops in _kernels(sass, 'k_project_tma').items():
        assert ops['UTMALDG'] >= 3, f'{name}: cp.async.bulk.tensor loads (UTMALDG) missing'
        assert ops['UTMASTG'] >= 2, f'{name}: cp.async.bulk.tensor stores (UTMASTG) missing'
        assert ops['SYNCS'] >= 3, f'{name}: mbarrier instructions missing'


def test_no_kernel_uses_tensor_cores_or_other_architectures(sass):
    # the path is a stencil + 8-point transforms: no MMA of any generation belongs in it
    for name, ops in sass.items():
        assert not any(op.startswith(('HMMA', 'IMMA', 'UTCMMA', 'UTCHMMA', 'WGMMA')) for op in ops), name
