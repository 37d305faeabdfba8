import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def oracle():
    import orc
    return orc.Oracle()


@pytest.fixture(scope="session")
def ref_c():
    import orc
    if not orc.have_ref():
        pytest.skip("oracle/_ref/libhavoc_ref.so not built (needs /root/reference; run make -C oracle ref)")
    return orc.Ref(use_asm=False)


@pytest.fixture(scope="session")
def ref_asm():
    import orc
    if not orc.have_ref():
        pytest.skip("oracle/_ref/libhavoc_ref.so not built")
    return orc.Ref(use_asm=True)
