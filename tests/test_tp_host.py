"""CPU test (gloo, world size 2) of the host-side tensor-parallel wiring: the handle exchange is an ordered all-gather
of opaque 64-byte blobs; no GPU is involved."""
import multiprocessing as mp
import os


def _worker(rank, world, port, q):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import torch.distributed as dist

    from metalchat_b200 import tp

    dist.init_process_group("gloo", rank=rank, world_size=world)
    blob = bytes([rank]) * 64
    got = tp.gather_blobs(blob)
    r, w, l = tp.env_rank_world()
    q.put((rank, [g[0] for g in got], all(len(g) == 64 for g in got), (r, w, l)))
    dist.barrier()
    dist.destroy_process_group()


def test_handle_exchange_gloo_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, 29517, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, firsts, sizes_ok, env in res:
        assert firsts == [0, 1] and sizes_ok and env == (rank, 2, rank)
