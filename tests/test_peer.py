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


@pytest.mark.parametrize("engine", ["umma", "simt"])
def test_surface_with_fused_exchange_world1(engine):
    """bogp_bartlett_grid_argext_peer / _host_argmax_peer on one GPU: the record that comes back through the exchange is
    (index_offset + arg-max index, value bits) of the surface -- published from inside the surface kernel on the
    FP16-split engine (no extra launch), by the pack + exchange kernels on the others."""
    from bogp_b200 import _lib, mfp
    from bogp_b200.modes import SWELLEX_TONALS_13, pekeris_modes, synthetic_covariance
    from bogp_b200.sharding import PeerExchange, reduce_records
    modes = pekeris_modes(SWELLEX_TONALS_13)
    r, z = np.linspace(0.1, 10.0, 300), np.linspace(1.0, 200.0, 120)
    dms = mfp.DeviceModeSet(modes, synthetic_covariance(modes, z[37], r[140]))
    px = PeerExchange()
    pair = torch.zeros(2, dtype=torch.int64, device="cuda")
    pairs = torch.zeros(1, 2, dtype=torch.int64, device="cuda")
    scratch = torch.empty(8192, dtype=torch.uint8, device="cuda")
    for step, off in enumerate((0, 7_000_000, 123)):
        n0 = _lib.kernel_launches()
        surf = dms.bartlett_grid(r, z, engine=engine, argext=(pair, scratch, True), peer=(px, off, pairs))
        launches = _lib.kernel_launches() - n0
        torch.cuda.synchronize()
        host = surf.cpu().numpy()
        val, idx = reduce_records(pairs)
        assert idx == off + int(host.argmax()) == off + 37 * 300 + 140 and val == host.max()
        assert int(pair[0]) == int(host.argmax())
        if engine == "umma" and step > 0:
            assert launches == 2          # table kernel + surface kernel: arg-max and exchange are inside the latter
    out, recs = dms.bartlett_grid_host(r, z, engine=engine, argext="max", peer=(px, 55))
    np.testing.assert_array_equal(out, host)
    assert reduce_records(recs) == (float(host.max()), 55 + int(host.argmax()))
    _, recs = dms.bartlett_grid_host(r, z, engine=engine, argext="min", peer=(px, 0))
    assert reduce_records(recs, maximize=False) == (float(host.min()), int(host.argmin()))
    px.close()


def test_pipelined_exchange_world1():
    """bogp_peer_set_pipelined: an exchange hands back the records of the PREVIOUS step (index -1 on the first), collect
    returns those of the last one; a collective step after pipelined ones still returns its own records.  Stand-alone
    kernel and the surface kernel's fused exchange."""
    from bogp_b200 import mfp
    from bogp_b200.modes import SWELLEX_TONALS_13, pekeris_modes, synthetic_covariance
    from bogp_b200.sharding import PeerExchange, reduce_records
    px = PeerExchange()
    rec = torch.zeros(2, dtype=torch.int64, device="cuda")
    out = torch.full((1, 2), -7, dtype=torch.int64, device="cuda")
    px.set_pipelined(True)
    sent = []
    for step in range(5):
        sent.append([500 + step, int(np.float32(0.25 * step).view(np.int32))])
        rec.copy_(torch.tensor(sent[-1], dtype=torch.int64))
        px.exchange(rec, out)
        torch.cuda.synchronize()
        assert out.cpu().tolist() == ([sent[-2]] if step else [[-1, 0]])
    px.collect(out)
    torch.cuda.synchronize()
    assert out.cpu().tolist() == [sent[-1]]
    px.set_pipelined(False)
    rec.copy_(torch.tensor([77, 5], dtype=torch.int64))
    px.exchange(rec, out)
    torch.cuda.synchronize()
    assert out.cpu().tolist() == [[77, 5]]
    # fused: three surfaces with different truths, each step's records arrive with the next call
    modes = pekeris_modes(SWELLEX_TONALS_13)
    r, z = np.linspace(0.1, 10.0, 300), np.linspace(1.0, 200.0, 120)
    pair = torch.zeros(2, dtype=torch.int64, device="cuda")
    pairs = torch.zeros(1, 2, dtype=torch.int64, device="cuda")
    scratch = torch.empty(8192, dtype=torch.uint8, device="cuda")
    px.set_pipelined(True)
    truths = [(37, 140), (80, 20), (5, 299)]
    prev = 77
    for jz, ir in truths:
        dms = mfp.DeviceModeSet(modes, synthetic_covariance(modes, z[jz], r[ir]))
        dms.bartlett_grid(r, z, engine="umma", argext=(pair, scratch, True), peer=(px, 1000, pairs))
        torch.cuda.synchronize()
        assert reduce_records(pairs)[1] == prev
        prev = 1000 + jz * 300 + ir
    px.collect(pairs)
    torch.cuda.synchronize()
    assert reduce_records(pairs)[1] == prev
    px.close()


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
        # the same through the surface kernel: every rank evaluates its slab of depth rows, the exchange runs in the
        # surface kernel's last CTA
        from bogp_b200 import mfp
        from bogp_b200.modes import SWELLEX_TONALS_13, pekeris_modes, synthetic_covariance
        from bogp_b200.sharding import reduce_records, slab
        modes = pekeris_modes(SWELLEX_TONALS_13)
        r, z = np.linspace(0.1, 10.0, 300), np.linspace(1.0, 200.0, 128)
        dms = mfp.DeviceModeSet(modes, synthetic_covariance(modes, z[90], r[140]))
        lo, hi = slab(len(z), rank, world)
        pair = torch.zeros(2, dtype=torch.int64, device="cuda")
        scratch = torch.empty(8192, dtype=torch.uint8, device="cuda")
        for _ in range(3):
            dms.bartlett_grid(r, z[lo:hi], engine="umma", argext=(pair, scratch, True), peer=(px, lo * len(r), out))
            torch.cuda.synchronize()
            ok = ok and reduce_records(out)[1] == 90 * 300 + 140
        _, recs = dms.bartlett_grid_host(r, z[lo:hi], engine="umma", argext="max", peer=(px, lo * len(r)))
        ok = ok and reduce_records(recs)[1] == 90 * 300 + 140
        # pipelined: no rank waits for the others inside a step; step k hands back the global record of step k - 1
        px.set_pipelined(True)
        truth = [(90, 140), (10, 7), (120, 250), (64, 64), (3, 299)]
        sets = [mfp.DeviceModeSet(modes, synthetic_covariance(modes, z[jz], r[ir])) for jz, ir in truth]
        prev = None
        for k, d in enumerate(sets):
            d.bartlett_grid(r, z[lo:hi], engine="umma", argext=(pair, scratch, True), peer=(px, lo * len(r), out))
            torch.cuda.synchronize()
            if prev is not None:
                ok = ok and reduce_records(out)[1] == prev
            prev = truth[k][0] * 300 + truth[k][1]
        px.collect(out)
        torch.cuda.synchronize()
        ok = ok and reduce_records(out)[1] == prev
        px.set_pipelined(False)
        sets[0].bartlett_grid(r, z[lo:hi], engine="umma", argext=(pair, scratch, True), peer=(px, lo * len(r), out))
        torch.cuda.synchronize()
        ok = ok and reduce_records(out)[1] == 90 * 300 + 140
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


def _shard_worker(rank: int, world: int, port: int, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world)
    try:
        from bogp_b200 import mfp
        from bogp_b200.envbatch import EnvBatch
        from bogp_b200.modes import (INV_TONALS_3, REC_Z_21, SWELLEX_TONALS_13, pekeris_modes, perturbed_pekeris_batch,
                                     synthetic_covariance)
        from bogp_b200.sharding import sharded_candidates, sharded_tilt_volume
        # configs[4]: per-candidate environments, slabs of candidates per rank, arg-min of the cbf_ml objective
        truth = pekeris_modes(INV_TONALS_3, REC_Z_21)
        K = synthetic_covariance(truth, 71.11, 1.07)
        rng = np.random.default_rng(4)
        Cn = 61
        depths, cbs = rng.uniform(212.0, 222.0, Cn), rng.uniform(1600.0, 1700.0, Cn)
        depths[40], cbs[40] = 217.0, 1650.0                       # the true environment sits in the second rank's slab
        sets = perturbed_pekeris_batch(INV_TONALS_3, REC_Z_21, depths, cbs)
        r, zs = np.full(Cn, 1.07), np.full(Cn, 71.11)

        def score(lo, hi):
            with EnvBatch(sets[lo:hi], r[lo:hi], zs[lo:hi], K) as eb:
                return eb.bartlett(atype="cbf_ml", multifreq_method="product")

        local, (val, idx), full = sharded_candidates(score, Cn, maximize=False, gather=True)
        whole = score(0, Cn)
        ok = idx == 40 and abs(val) < 2e-5 and bool(torch.equal(full, whole))
        # configs[3]: the (tilt, depth, range) volume split over tilt planes
        modes = pekeris_modes(SWELLEX_TONALS_13)
        rr, zz, tilt = np.linspace(0.5, 9.0, 150), np.linspace(5.0, 195.0, 24), np.linspace(-2.0, 2.0, 5)
        dms = mfp.DeviceModeSet(modes, synthetic_covariance(modes, zz[13], rr[77], tilt_deg=tilt[3]))
        planes, (vmax, pos) = sharded_tilt_volume(lambda lo, hi: dms.bartlett_grid(rr, zz, tilt[lo:hi]), len(tilt), len(zz), len(rr))
        vol = dms.bartlett_grid(rr, zz, tilt)
        flat = int(vol.reshape(-1).argmax())
        ok = ok and pos == tuple(int(x) for x in np.unravel_index(flat, tuple(vol.shape))) and pos == (3, 13, 77)
        ok = ok and vmax == float(vol.reshape(-1)[flat])
        q.put((rank, ok))
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs on one box")
def test_environment_batch_and_tilt_volume_sharded_over_two_gpus():
    """BASELINE configs[4] (own mode set per candidate) and configs[3] (tilt volume) over two ranks: each rank scores its
    slab with its own EnvBatch / tilt-plane call, the gathered values equal the one-GPU call bit for bit and the global
    extremum is the truth (environment 40; tilt 3, depth 13, range 77)."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_shard_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=240) for _ in procs)
    for p in procs:
        p.join(60)
    assert [r[1] for r in res] == [True, True]

