"""-m gpu: hardware assumption behind the resident-weight conv kernels (conv3_res.cu, wgrad.cu): the tensor core applies the
128B swizzle on ABSOLUTE shared-memory address bits, so a K-major UMMA operand may start at any 128-byte row of a swizzled
tile and its 8-row groups may be `sbo` bytes apart with sbo not a multiple of 1024 (the 16x8-pixel tile inside an 18x10 halo
tile uses sbo = 1280; wgrad's stacked taps use 2304). Checked on the device with a TEST-ONLY object (tests/probe/)."""
import ctypes
import os

import pytest
import torch

pytestmark = pytest.mark.gpu
LIB = os.path.join(os.path.dirname(os.path.abspath(__file__)), "probe", "libb200probe.so")


def test_umma_descriptor_start_and_group_stride_need_no_atom_alignment():
    if not os.path.exists(LIB):
        pytest.skip("tests/probe/libb200probe.so not built (__graft_entry__.build() builds it)")
    lib = ctypes.CDLL(LIB)
    fn = lib.b200probe_shift
    fn.restype = ctypes.c_int
    fn.argtypes = [ctypes.c_void_p] * 3 + [ctypes.c_int] * 3 + [ctypes.c_void_p]
    g = torch.Generator().manual_seed(0)
    a = torch.randn(256, 64, generator=g).to(torch.bfloat16).cuda()
    b = torch.randn(64, 64, generator=g).to(torch.bfloat16).cuda()
    checked = 0
    for sbo in (1024, 1280, 2304):
        for shift in (0, 1, 2, 3, 7, 8, 10, 11, 12, 20, 21, 22):
            rows = torch.tensor([shift + (m // 8) * (sbo // 128) + (m % 8) for m in range(128)])
            if int(rows.max()) >= 256:
                continue
            out = torch.zeros(128, 64, device="cuda")
            assert fn(a.data_ptr(), b.data_ptr(), out.data_ptr(), shift, 0, sbo, torch.cuda.current_stream().cuda_stream) == 0
            torch.cuda.synchronize()
            want = a[rows.cuda()].float() @ b.float().t()
            assert float((out - want).abs().max()) < 1e-3, (sbo, shift)
            checked += 1
    assert checked >= 20
