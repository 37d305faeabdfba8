"""__graft_entry__.smoke()'s tiny frame pass -- motion search, intra sweep and the TU pipeline with RDOQ over a 128x64
picture, checked against the oracle's frame-pass loops -- on the kernels' own source under the host emulator
(tests/emu_context.py) instead of on cuda:0: the driver's GPU smoke test, in the CPU-only suite."""
import __graft_entry__ as entry
from emu_context import EmuContext


def test_smoke_pass_under_emulation():
    units, _ = entry.smoke_pass(EmuContext(0, 1, 8))
    assert units["pu_searches"] > 100 and units["intra_partitions"] > 100 and units["tus"] > 100
