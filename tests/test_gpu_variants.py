"""The opt-in / fallback kernel variants behind the developer switches of fastpath.cu must stay
bit-for-bit as correct as the default path: each variant re-runs the cfg1-shaped fused parity test (through the C ABI,
against the oracle) in a fresh process, because the switches are read once per process.  The switches exist only in
the developer build libb200dsp_dev.so (`make -C dspsr_b200/csrc dev`, -DB200_TUNING); the product library never reads
the environment."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DEVLIB = os.path.join(ROOT, "dspsr_b200", "libb200dsp_dev.so")

VARIANTS = {
    "z_natural_order": {"B200_Z_TILED": "0"},
    "k2_single_group": {"B200_K2_G2": "0"},
    "first_generation_plans": {"B200_K2_R32": "0", "B200_K3_R32": "0"},
    "generic_kernels": {"B200_FAST": "0"},
}


# longconv.cu: each of the three c2 kernels alone between the generic kernels (plane-per-polarisation
# scratch), and the generic three-kernel path the product falls back to for real input
LONG_VARIANTS = {
    "long_only_k1": {"B200_BC_K2": "0", "B200_BC_K3": "0"},
    "long_only_k2": {"B200_BC_K1": "0", "B200_BC_K3": "0"},
    "long_only_k3": {"B200_BC_K1": "0", "B200_BC_K2": "0"},
    "long_generic": {"B200_BIG_CONV": "0"},
}


@pytest.mark.parametrize("name", sorted(LONG_VARIANTS))
def test_long_convolution_variant_matches_oracle(name):
    if not os.path.exists(DEVLIB):
        pytest.skip("developer build libb200dsp_dev.so not present")
    env = dict(os.environ)
    env.update(LONG_VARIANTS[name])
    env["B200_LIB"] = DEVLIB
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(ROOT, "tests", "test_gpu_parity.py"), "-q", "-x",
                        "-m", "gpu", "-k", "test_long_convolution_kernels and (262144 or 1048576 or 2097152)",
                        "-p", "no:cacheprovider"],
                       cwd=ROOT, env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, (name, r.stdout[-2000:], r.stderr[-2000:])
    assert "3 passed" in r.stdout, r.stdout[-500:]


@pytest.mark.parametrize("name", sorted(VARIANTS))
def test_variant_matches_oracle(name):
    if not os.path.exists(DEVLIB):
        pytest.skip("developer build libb200dsp_dev.so not present")
    env = dict(os.environ)
    env.update(VARIANTS[name])
    env["B200_LIB"] = DEVLIB
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(ROOT, "tests", "test_gpu_parity.py"), "-q", "-x",
                        "-m", "gpu", "-k", "test_pipeline_cfg1 and not bench_scale or test_filterbank_cfg1_shape",
                        "-p", "no:cacheprovider"],
                       cwd=ROOT, env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, (name, r.stdout[-2000:], r.stderr[-2000:])
    assert "2 passed" in r.stdout, r.stdout[-500:]
