"""-m gpu: the drop-in (libcrnlib_b200.so over the CUDA library) against the unmodified reference through the same test program, at 256 x 256
(see tests/test_dropin_cpu.py).  The binaries are prebuilt by __graft_entry__.build() where the reference's headers exist."""
import pytest

from test_dropin_cpu import compare_demo, run_demo

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("size,exact_vq", [(64, True), (256, False), (512, False)])
def test_gpu_dropin_matches_reference(size, exact_vq):
    compare_demo(run_demo("demo_b200", size, exact_vq), run_demo("demo_ref", size))
