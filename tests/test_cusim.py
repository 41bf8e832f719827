"""The CUDA kernels themselves, compiled for the CPU SIMT emulator (tests/cusim), against the reference:
same .cu sources, every lane a fiber, warp collectives as rendezvous.  Checks kernel logic without a GPU."""
import pytest

from conftest import CUSIMBIN, ref_sam, run_sam

CASES = [
    ("c1_se", "c1", dict(tag="cs", n=160, seed=31), []),
    ("stress_se", "stress", dict(tag="cs", n=120, seed=33, err=(0.016, 0.002, 0.002), chimeric=0.05), []),
    ("stress_pe", "stress", dict(tag="cspe", n=60, seed=34, paired=True, err=(0.016, 0.002, 0.002), chimeric=0.05), []),
    ("two_1k", "two", dict(tag="cs1k", n=6, length=1000, seed=35), []),
]


@pytest.mark.parametrize("name,ref,kw,extra", CASES, ids=[c[0] for c in CASES])
def test_emulated_kernels_sam_vs_reference(data, name, ref, kw, extra):
    fa, fqs = data.reads(ref, **kw)
    args = extra + ["-K", "100000000", "-t", "2", fa] + fqs
    assert run_sam(CUSIMBIN, args) == ref_sam(args)
