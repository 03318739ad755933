"""utils/cuda_vectors.h (reference utils/cuda_vectors.h:7-141): every float4 / double2 operator,
evaluated on the HOST through its __host__ __device__ definition (nvcc compiles tests/aux/cuda_vectors_host.cu here, no
GPU needed) against numpy.  The device side is exercised by every benchmark01-03 kernel: csrc/vec_kernels.cu includes
the header and accumulates through its operators (tests/test_vec_gpu.py, tests/test_reference_kernels_gpu.py)."""
import os
import shutil
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NVCC = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"


@pytest.fixture(scope="module")
def results(tmp_path_factory):
    if not os.path.exists(NVCC):
        pytest.skip("nvcc not found")
    exe = tmp_path_factory.mktemp("cv") / "cuda_vectors_host"
    hostcxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    subprocess.check_call([NVCC, "-ccbin", hostcxx, "-std=c++17", "-O1", "-o", str(exe),
                           os.path.join(ROOT, "tests", "aux", "cuda_vectors_host.cu")])
    out = subprocess.run([str(exe)], check=True, stdout=subprocess.PIPE).stdout.decode().splitlines()
    return {l.split()[0]: [float(v) for v in l.split()[1:]] for l in out}


def test_every_operator_matches_componentwise_arithmetic(results):
    a, b, s = np.array([1.25, -3.5]), np.array([0.1, 7.0]), 2.5
    f = np.array([1.5, -2.25, 0.3, 8.0], dtype=np.float32)
    g = np.array([0.7, 4.0, -1.1, 0.125], dtype=np.float32)
    t = np.float32(-0.75)
    want = {"d+v": a + b, "d-v": a - b, "d*v": a * b, "d+s": a + s, "d-s": a - s, "d*s": a * s,
            "f+v": f + g, "f-v": f - g, "f*v": f * g, "f+s": f + t, "f-s": f - t, "f*s": f * t}
    for k in list(want):
        want[k[0] + k[1] + "=" + k[2]] = want[k]          # the compound forms give the same values
    assert len(results) == 25
    for k, w in want.items():
        got = np.array(results[k], dtype=w.dtype)
        assert np.array_equal(got, w), (k, got, w)        # one IEEE operation per component: exact
    # host compilers may or may not contract a*a + acc: accept either rounding of the chained form
    chained = a * a + b * b
    assert np.allclose(results["dsq"], chained, rtol=4e-16, atol=0)


def test_the_vector_kernels_include_the_header():
    src = open(os.path.join(ROOT, "gpu-benchmarking_b200", "csrc", "vec_kernels.cu")).read()
    assert '#include "../../utils/cuda_vectors.h"' in src and "a += v * v;" in src and "a += u * v;" in src
