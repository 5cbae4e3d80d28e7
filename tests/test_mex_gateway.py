"""The MEX gateway (matlab/redmax_mex.cpp) compiled by g++ against a functional stand-in for mex.h (tests/stub/mex.h) and
driven through mexFunction by tests/stub/mex_driver.cpp: MATLAB is not in this image, so this is how the binding of
INTEGRATION.md sees a compiler and runs.  CPU part: scene create / destroy with index arrays of either class, argument
validation (errors are raised as redmax:arg, never a crash).  GPU part: 'rollout' and 'resume' through the gateway."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIBDIR = os.path.join(ROOT, 'redmax_b200', 'lib')


@pytest.fixture(scope='module')
def driver(tmp_path_factory):
    import __graft_entry__ as ge
    ge.build_cuda()
    exe = str(tmp_path_factory.mktemp('mex') / 'mex_driver')
    cmd = ['g++', '-std=c++14', '-Wall', '-Werror', '-DMATLAB_MEX_FILE', '-I' + os.path.join(ROOT, 'tests', 'stub'),
           '-I' + os.path.join(ROOT, 'include'), '-o', exe, os.path.join(ROOT, 'matlab', 'redmax_mex.cpp'),
           os.path.join(ROOT, 'tests', 'stub', 'mex_driver.cpp'), '-L' + LIBDIR, '-lredmax_b200', '-Wl,-rpath,' + LIBDIR]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    return exe


def test_gateway_compiles_and_creates_scenes(driver):
    r = subprocess.run([driver, 'create'], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and r.stdout.strip().endswith('PASS'), r.stdout + r.stderr
    assert 'create (chart as int32)' in r.stdout and 'create (chart as double)' in r.stdout


@pytest.mark.gpu
def test_gateway_rollout_and_resume(driver):
    r = subprocess.run([driver, 'rollout'], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and r.stdout.strip().endswith('PASS'), r.stdout + r.stderr
    assert 'resume from step 5 (int32 kbegin): max |dq| = 0' in r.stdout
