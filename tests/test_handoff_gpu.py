"""Device-side pipeline hand-off kernels (csrc/handoff.cu) on ONE GPU: the mailbox is this process's own memory, the
"peer" pointer is the same allocation, sender and receiver run on two streams.  Covers the sequence-number protocol
(the k-th wait pairs with the k-th send), the first-wait-without-a-send rule of stage 0, the
payload copy and the time-out path.  The two-process / two-GPU form runs in tools/pipeline_check.py."""
import ctypes

import pytest
import torch

pytestmark = pytest.mark.gpu


def _mailbox(nbytes):
    from quip_for_all_b200._native import check, lib
    L = lib()
    ptr, handle = ctypes.c_void_p(), ctypes.create_string_buffer(64)
    check(L.quipb200_mailbox_create(nbytes, ctypes.byref(ptr), handle), "mailbox_create")
    assert any(handle.raw)          # a CUDA IPC handle was exported
    return L, ptr.value


def test_handoff_send_wait_protocol():
    from quip_for_all_b200._native import check
    dev = torch.device("cuda:0")
    FLAG = 16384
    L, box = _mailbox(FLAG + 256)
    try:
        send_ctr = torch.zeros(1, dtype=torch.int64, device=dev)
        wait_ctr = torch.zeros(1, dtype=torch.int64, device=dev)
        err = torch.zeros(1, dtype=torch.int32, device=dev)
        s_send, s_recv = torch.cuda.Stream(), torch.cuda.Stream()
        torch.cuda.synchronize()
        srcs = [torch.randn(4096, device=dev).half() for _ in range(3)]
        dsts = [torch.zeros(4096, dtype=torch.float16, device=dev) for _ in range(3)]
        torch.cuda.synchronize()

        def send(i):
            with torch.cuda.stream(s_send):
                check(L.quipb200_handoff_send(srcs[i].data_ptr(), box, 8192, box + FLAG, send_ctr.data_ptr(),
                                              ctypes.c_void_p(s_send.cuda_stream)), "send")

        def wait(i):
            with torch.cuda.stream(s_recv):
                check(L.quipb200_handoff_wait(box + FLAG, wait_ctr.data_ptr(), box, dsts[i].data_ptr(), 8192, err.data_ptr(),
                                              ctypes.c_void_p(s_recv.cuda_stream)), "wait")

        # producer before consumer on every round (two streams of ONE device are not guaranteed to run concurrently -- a
        # box with CUDA_DEVICE_MAX_CONNECTIONS=1 would serialise a spinning consumer in front of its producer; the
        # consumer-first order is exercised across two GPUs by tools/pipeline_check.py)
        for i in range(3):
            send(i)
            s_send.synchronize()
            wait(i)
            s_recv.synchronize()
            assert torch.equal(dsts[i], srcs[i]) and int(err.item()) == 0, i
        assert int(send_ctr.item()) == 3 and int(wait_ctr.item()) == 3
        # argument checks of the C ABI
        assert L.quipb200_handoff_send(srcs[0].data_ptr(), box, 12, box + FLAG, send_ctr.data_ptr(), None) == -1
        assert L.quipb200_handoff_wait(box + FLAG, wait_ctr.data_ptr(), box, dsts[0].data_ptr() + 2, 8192, err.data_ptr(), None) == -2
    finally:
        torch.cuda.synchronize()
        L.quipb200_mailbox_destroy(box)


def test_handoff_first_wait_passes_and_timeout_is_reported():
    from quip_for_all_b200._native import check
    dev = torch.device("cuda:0")
    FLAG = 4096
    L, box = _mailbox(FLAG + 256)
    try:
        wait_ctr = torch.full((1,), -1, dtype=torch.int64, device=dev)     # stage 0: the first run consumes the prefill's token
        err = torch.zeros(1, dtype=torch.int32, device=dev)
        tok = torch.tensor([1234], dtype=torch.int64, device=dev)
        st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        check(L.quipb200_handoff_wait(box + FLAG, wait_ctr.data_ptr(), box, tok.data_ptr(), 8, err.data_ptr(), st), "wait")
        torch.cuda.synchronize()
        assert int(tok.item()) == 1234 and int(err.item()) == 0 and int(wait_ctr.item()) == 0
        # second wait with no producer: gives up after ~2 s of GPU time, leaves the buffer alone and counts the error
        check(L.quipb200_handoff_wait(box + FLAG, wait_ctr.data_ptr(), box, tok.data_ptr(), 8, err.data_ptr(), st), "wait")
        torch.cuda.synchronize()
        assert int(err.item()) == 1 and int(tok.item()) == 1234
    finally:
        L.quipb200_mailbox_destroy(box)
