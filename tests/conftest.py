import os
import subprocess
import sys

import pytest

# Multi-rank tests keep several device objects (6 streams each) on ONE GPU; with the default of 8 hardware queues their
# streams would alias and a rank's bounded spin-wait for a peer could sit in front of that peer's work.
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with `-m gpu` on the GPU box)")


def _ensure_built():
    """Build the CUDA extension and the reference oracle if they are missing (cross-compiles without a GPU).
    On the GPU box /root/reference does not exist; the prebuilt files travel with the snapshot."""
    lib = os.path.join(ROOT, "malevich_b200", "csrc", "libmalevich_b200.so")
    if not os.path.exists(lib):
        subprocess.run(["make", "-C", os.path.join(ROOT, "malevich_b200", "csrc")], check=True, capture_output=True)
    ref_dir = os.path.join(ROOT, "oracle", "_ref")
    want = ["320x200", "1200x720", "1280x720", "1920x1080", "3840x2160"]
    missing = [r for r in want if not os.path.exists(os.path.join(ref_dir, f"libmalevich_ref_{r}.so"))]
    if missing and os.path.exists("/root/reference/source/main.c"):
        subprocess.run([os.path.join(ROOT, "oracle", "build_ref.sh")] + missing, check=True, capture_output=True)


_ensure_built()


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")


def have_ref(w, h):
    return os.path.exists(os.path.join(ROOT, "oracle", "_ref", f"libmalevich_ref_{w}x{h}.so"))
