"""World-size-2 `gloo` tests (CPU) of the N>1 host logic: contiguous sharding
with the reference's rule (helper_multi_gpu.cu:64-101) and key replication by
one broadcast.  The compute on each shard is done here by the oracle, only to
show that shard results concatenate to the unsharded result."""
import os
import socket

import numpy as np
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import sys

    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import torch.distributed as dist

    from oracle import oracle as O
    from tfhe_rs_b200 import multi_gpu

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    P = O.TOY_K1
    total = 11  # ragged split: 6 + 5
    if rank == 0:
        keys = O.keygen(P, 4242)
        bsk, ksk, sk, gsk = keys.bsk, keys.ksk, keys.lwe_sk, keys.glwe_sk
    else:
        bsk = ksk = sk = gsk = None
    # one broadcast per key (and the secret keys, for the test's decryption)
    bsk = multi_gpu.broadcast_host_array(bsk, P.num_ggsw * P.ggsw_polys * P.N, np.uint64)
    ksk = multi_gpu.broadcast_host_array(ksk, P.big_n * P.ks_level * (P.n + 1), np.uint64)
    sk = multi_gpu.broadcast_host_array(sk, P.n, np.uint64)
    gsk = multi_gpu.broadcast_host_array(gsk, P.big_n, np.uint64)
    keys = O.KeySet(P, sk, gsk, bsk, ksk)
    msgs = np.arange(total) % P.p
    cts = O.lwe_encrypt_batch(O.Rng(5), sk, msgs.astype(np.uint64) * np.uint64(P.delta), P.lwe_noise_log2)
    lo, hi = multi_gpu.shard_range(total, rank, world)
    lut = O.make_lut(P, [(i + 3) % P.p for i in range(P.p)])
    out = O.pbs_batch(keys, lut, cts[lo:hi])
    dec = O.decode(O.lwe_decrypt_batch(gsk, out), P.delta, P.p)
    q.put((rank, lo, hi, dec.tolist()))
    dist.barrier()
    dist.destroy_process_group()


def test_shard_rule_matches_reference():
    from tfhe_rs_b200 import multi_gpu

    for total in (0, 1, 7, 4096, 4097, 65536):
        for world in (1, 2, 3, 8):
            sizes = [multi_gpu.get_num_inputs_on_gpu(total, r, world) for r in range(world)]
            offs = [multi_gpu.get_gpu_offset(total, r, world) for r in range(world)]
            assert sum(sizes) == total
            assert offs == [sum(sizes[:r]) for r in range(world)]
            assert max(sizes) - min(sizes) <= 1 and sizes == sorted(sizes, reverse=True)


def test_two_rank_sharded_pbs_with_broadcast_keys():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert [(r[1], r[2]) for r in results] == [(0, 6), (6, 11)]
    dec = sum((r[3] for r in results), [])
    assert dec == [((m % 16) + 3) % 16 for m in range(11)]
