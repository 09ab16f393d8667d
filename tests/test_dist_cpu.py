"""Multi-rank host logic on CPU (gloo, world_size 2): utterance sharding and the single flat-buffer
weight broadcast that is the only collective of the inference path."""
import os
import socket
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from fac_via_ppg_b200 import dist as fdist, synth
    from fac_via_ppg_b200.packing import PackedTacotron, PackedWaveGlow

    class Holder:                                    # the packed()/empty_packed()/use_packed() protocol
        def __init__(self, make_full, make_empty):
            self._full, self._empty, self.adopted = make_full, make_empty, None

        def packed(self):
            return self._full()

        def empty_packed(self):
            return self._empty()

        def use_packed(self, p):
            self.adopted = p

    cfg = synth.WAVEGLOW_CONFIG_SMALL
    wg = Holder(lambda: PackedWaveGlow.from_state(synth.waveglow_state(cfg=cfg), cfg, "cpu"),
                lambda: PackedWaveGlow(cfg, "cpu"))
    got = fdist.broadcast_packed(wg, src=0)
    ref = PackedWaveGlow.from_state(synth.waveglow_state(cfg=cfg), cfg, "cpu")
    ok_wg = torch.equal(got.flat, ref.flat) and wg.adopted is got
    hp = synth.TACOTRON_HPARAMS
    taco = Holder(lambda: PackedTacotron.from_state(synth.tacotron_state(), hp, "cpu"), lambda: PackedTacotron(hp, "cpu"))
    got_t = fdist.broadcast_packed(taco, src=0)
    ok_t = torch.equal(got_t.flat, PackedTacotron.from_state(synth.tacotron_state(), hp, "cpu").flat)
    shard = fdist.shard_indices(7, rank, world)
    shard_len = fdist.shard_indices(5, rank, world, lengths=[10, 50, 30, 20, 40])
    gathered = [None] * world
    dist.all_gather_object(gathered, (shard, shard_len))
    if rank == 0:
        out.put((ok_wg, ok_t, gathered))
    else:
        assert ok_wg and ok_t
    dist.destroy_process_group()


def test_broadcast_and_sharding_world2():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    ok_wg, ok_t, gathered = out.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert ok_wg and ok_t
    (s0, l0), (s1, l1) = gathered
    assert sorted(s0 + s1) == list(range(7)) and not set(s0) & set(s1)
    assert sorted(l0 + l1) == list(range(5))
    # longest-first round robin: rank 0 gets lengths 50, 30, 10; rank 1 gets 40, 20
    assert l0 == [0, 1, 2] and l1 == [3, 4]


def test_shard_indices_cover_everything_once():
    from fac_via_ppg_b200.dist import shard_indices
    for n in (0, 1, 8, 13):
        for world in (1, 2, 4, 8):
            parts = [shard_indices(n, r, world) for r in range(world)]
            assert sorted(sum(parts, [])) == list(range(n))
