import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")
    if os.environ.get("MADGPU_EMULATED_DEVICE") == "1":
        _bind_emulated_device()


def _bind_emulated_device():
    """MADGPU_EMULATED_DEVICE=1 python -m pytest tests/test_gpu_ved.py -m gpu ...: dry-run GPU test files on the CPU.  The
    product's Python binding (and, through LD_LIBRARY_PATH, the C++ drop-in test programs) is pointed at the host build of the CUDA
    source (tests/mad_host/), so the TEST CODE can be checked before it is spent on a GPU box.  Slow, and never a substitute for the
    GPU run: the driver's `-m gpu` run does not set this variable."""
    import shutil
    import subprocess
    import tempfile
    # never on a box with a GPU: there the -m gpu tests must run the product (libmadgpu.so on the device), and a stray variable
    # must not turn them into a CPU dry run that reports green
    has_gpu = os.path.exists("/dev/nvidiactl") or os.path.exists("/dev/nvidia0")
    if not has_gpu and shutil.which("nvidia-smi"):
        has_gpu = subprocess.run(["nvidia-smi", "-L"], capture_output=True, text=True).stdout.strip().startswith("GPU")
    if has_gpu:
        raise pytest.UsageError("MADGPU_EMULATED_DEVICE=1 is refused on a machine with a CUDA device: the GPU tests must run "
                                "libmadgpu.so on the device, not the CPU dry-run build (unset the variable)")
    print("\n" + "=" * 100 + "\n  MADGPU_EMULATED_DEVICE=1: `-m gpu` tests run on the CPU DRY-RUN BUILD of the CUDA source (tests/mad_host), NOT on a GPU.\n"
          "  Results of this session say nothing about the product on hardware.\n" + "=" * 100, file=sys.stderr, flush=True)
    sys.path.insert(0, os.path.join(ROOT, "tests", "mad_host"))
    import hostlib
    hostlib.bind(hostlib.load())
    d = tempfile.mkdtemp(prefix="madgpu_emulated_")
    os.symlink(hostlib.HOST_LIB, os.path.join(d, "libmadgpu.so"))
    os.environ["LD_LIBRARY_PATH"] = d + os.pathsep + os.environ.get("LD_LIBRARY_PATH", "")


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")
