"""GPU tier: the peer-memory record exchange (bogp_peer_*, csrc/kernels_peer.cu) that replaces the NCCL all-gather of the
final (arg-max index, value) records.  World 1 through the C ABI on any GPU box; world 2 over CUDA IPC when the box has
two GPUs (skipped otherwise; run with `gpurun --gpus 2 -- python -m pytest tests/test_peer.py -m gpu`)."""
import ctypes as C
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_peer_exchange_world1_self_exchange():
    """One rank exchanging with itself: every epoch returns the record just written, the double-buffered slots and the
    epoch flags advance, and a closed handle can be re-created."""
    from bogp_b200 import _lib
    for _ in range(2):
        h = C.c_void_p()
        handle = (C.c_ubyte * 64)()
        _lib.call("bogp_peer_create", C.byref(h), 0, 1, C.cast(handle, C.c_void_p))
        _lib.call("bogp_peer_open", h, C.cast(handle, C.c_void_p))
        rec = torch.zeros(2, dtype=torch.int64, device="cuda")
        out = torch.full((1, 2), -7, dtype=torch.int64, device="cuda")
        for step in range(5):
            rec.copy_(torch.tensor([1000 + step, int(np.float32(0.5 + step).view(np.int32))], dtype=torch.int64))
            _lib.call("bogp_peer_exchange", h, C.c_void_p(rec.data_ptr()), C.c_void_p(out.data_ptr()),
                      C.c_void_p(_lib.current_stream_ptr()))
            torch.cuda.synchronize()
            assert out.cpu().tolist() == [rec.cpu().tolist()]
        with pytest.raises(_lib.BogpError):
            _lib.call("bogp_peer_open", h, C.cast(handle, C.c_void_p))          # already open
        _lib.lib().bogp_peer_destroy(h)
    with pytest.raises(_lib.BogpError):
        _lib.call("bogp_peer_create", C.byref(C.c_void_p()), 3, 2, C.cast((C.c_ubyte * 64)(), C.c_void_p))


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _peer_worker(rank: int, world: int, port: int, q):
    import torch.distributed as dist
    from bogp_b200.sharding import PeerExchange
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world)
    try:
        px = PeerExchange()
        rec = torch.zeros(2, dtype=torch.int64, device="cuda")
        out = torch.zeros(world, 2, dtype=torch.int64, device="cuda")
        check = torch.zeros_like(out)
        ok = True
        for step in range(20):
            val = np.float32(0.1 * ((rank + step) % world) + 0.01 * step)
            rec.copy_(torch.tensor([rank * 1_000_000 + step, int(val.view(np.int32))], dtype=torch.int64))
            px.exchange(rec, out)
            dist.all_gather_into_tensor(check.view(-1), rec)
            torch.cuda.synchronize()
            ok = ok and torch.equal(out, check)
        px.close()
        q.put((rank, ok))
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs on one box")
def test_peer_exchange_two_gpus_matches_nccl_allgather():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_peer_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=180) for _ in procs)
    for p in procs:
        p.join(60)
    assert [r[1] for r in res] == [True, True]
