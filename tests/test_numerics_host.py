"""The exact shared-reciprocal division of numerics.cuh, restated for the host and checked against
IEEE division on the CPU (tests/qdiv_check.c); the exhaustive GPU-side check is tools/divcheck.cu."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def test_shared_reciprocal_division_is_correctly_rounded(tmp_path):
    exe = tmp_path / 'qdiv_check'
    cc = '/usr/bin/gcc' if os.path.exists('/usr/bin/gcc') else 'gcc'
    subprocess.run([cc, '-O2', '-std=c11', '-ffp-contract=off', '-msse2', '-mfpmath=sse', '-o', str(exe),
                    os.path.join(HERE, 'qdiv_check.c'), '-lm'], check=True)
    r = subprocess.run([str(exe), '30000000'], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert ' 0 mismatches' in r.stdout


def test_row_guard_property(tmp_path):
    """One magnitude test per loaded value implies the per-numerator guard (tests/rowguard_check.c)."""
    exe = tmp_path / 'rowguard_check'
    cc = '/usr/bin/gcc' if os.path.exists('/usr/bin/gcc') else 'gcc'
    subprocess.run([cc, '-O2', '-std=c11', '-ffp-contract=off', '-msse2', '-mfpmath=sse', '-o', str(exe),
                    os.path.join(HERE, 'rowguard_check.c'), '-lm'], check=True)
    r = subprocess.run([str(exe), '5000000'], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert ' 0 below 2^-60' in r.stdout


def test_packed_products_are_not_contracted(tmp_path):
    """ptxas 12.9 fuses mul.rn.f32x2 + add/sub.rn.f32x2 into one FFMA2 (a single rounding where
    the reference has two) although the operations carry an explicit .rn and the build passes
    --fmad=false; NVIDIA's own __fadd2_rn(__fmul2_rn(x, y), z) is affected too.  The packed
    gradient kernel therefore routes every sum with a product through addm2() (numerics.cuh).
    This test is the tripwire: for every instantiation of k_gradient_packed the SASS must contain
    exactly as many FFMA2 / FMUL2 / FADD2 as the PTX has fma / mul / add+sub .rn.f32x2 — a
    contraction would turn one FMUL2 and one FADD2 into an extra FFMA2."""
    import collections
    import re
    import shutil
    import subprocess
    nvcc = shutil.which('nvcc') or '/usr/local/cuda/bin/nvcc'
    cuobjdump = shutil.which('cuobjdump') or '/usr/local/cuda/bin/cuobjdump'
    if not (os.path.exists(nvcc) and os.path.exists(cuobjdump)):
        import pytest
        pytest.skip('CUDA toolkit not installed')
    flags = ['-O3', '-std=c++17', '-fmad=false', '-prec-div=true', '-prec-sqrt=true', '-ftz=false']
    # (source, packed FMAs a kernel must at least have, kernels expected): the gradient kernel and the
    # projection kernels whose stepper / clamp sections run on packed fp32
    for base, min_fma, min_kernels in (('kernels_gradient_packed', 30, 12), ('kernels_project_tma', 10, 2)):
        src = os.path.join(ROOT, 'jpeg2png_b200', 'csrc', base + '.cu')
        ptx, cubin = str(tmp_path / (base + '.ptx')), str(tmp_path / (base + '.cubin'))
        subprocess.run([nvcc, '-gencode', 'arch=compute_100a,code=compute_100a', *flags, '-ptx', '-o', ptx, src], check=True, capture_output=True)
        subprocess.run([nvcc, '-gencode', 'arch=compute_100a,code=sm_100a', *flags, '-cubin', '-o', cubin, src], check=True, capture_output=True)
        want = {}
        for entry in re.split(r'\n\.visible \.entry ', open(ptx).read())[1:]:
            c = collections.Counter(re.findall(r'\b(fma|mul|add|sub)\.rn\.f32x2\b', entry))
            want[entry.split('(')[0].strip()] = (c['fma'], c['mul'], c['add'] + c['sub'])
        sass = subprocess.run([cuobjdump, '-sass', cubin], check=True, capture_output=True, text=True).stdout
        seen = 0
        for fun in re.split(r'\n\s+Function : ', sass)[1:]:
            name = fun.split('\n')[0].strip()
            if name not in want or want[name] == (0, 0, 0):
                continue
            c = collections.Counter(re.findall(r'\b(FFMA2|FMUL2|FADD2)\b', fun))
            assert (c['FFMA2'], c['FMUL2'], c['FADD2']) == want[name], f'{name}: SASS {dict(c)} vs PTX fma/mul/add+sub {want[name]}'
            assert want[name][0] > min_fma, f'{name} no longer uses packed fp32?'
            seen += 1
        assert seen >= min_kernels, base


def test_small_int_to_float_trick():
    """project_common.cuh small_int_to_float: (float)d for a quantised coefficient without the
    conversion pipe — integer add on the bit pattern of 1.5 * 2^23, then an exact fp32 subtraction.
    Exhaustive over int16 (the coefficient type) and beyond, against numpy's conversion."""
    import numpy as np
    d = np.arange(-(1 << 22) + 1, 1 << 22, dtype=np.int64)
    bits = (np.int64(0x4B400000) + d).astype(np.uint32)
    got = bits.view(np.float32) - np.float32(12582912.0)
    assert got.dtype == np.float32
    assert (got == d.astype(np.float32)).all()
    assert (np.signbit(got) == (d < 0)).all()          # and +0 for d == 0, like (float)0


def test_product_sum_through_fma_with_one():
    """numerics.cuh addm2: fma(m, 1, b) == RN(m + b) bit for bit (the form that keeps ptxas from
    contracting a packed product into the sum that follows it).  fp64 has fp32's exact products
    and sums to spare, so RN32 of the exact value is one cast away."""
    import numpy as np
    rng = np.random.default_rng(3)
    m = (rng.standard_normal(2_000_000) * np.exp2(rng.integers(-40, 40, 2_000_000))).astype(np.float32)
    b = (rng.standard_normal(2_000_000) * np.exp2(rng.integers(-40, 40, 2_000_000))).astype(np.float32)
    b[::7] = -m[::7]                                     # exact cancellations
    b[::11] = 0.0
    m[::13] = 0.0
    fma = (m.astype(np.float64) * 1.0 + b.astype(np.float64)).astype(np.float32)   # exact in fp64, one rounding
    add = m + b
    assert (fma.view(np.uint32) == add.view(np.uint32)).all()
