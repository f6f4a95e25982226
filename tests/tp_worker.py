"""Multi-GPU parity worker: run as  torchrun --nproc-per-node N tests/tp_worker.py  (see tests/test_gpu_tp.py).

Every rank holds one shard of the 1B bf16 model (tensor parallel over WORLD_SIZE GPUs) plus a full single-GPU copy;
both decode the golden prompt greedily.  Checks: TP tokens == single-GPU tokens == oracle fixture; the gathered logits
shards match the single-GPU logits within the fp32 re-association tolerance."""
import json
import os
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def main():
    import torch
    import torch.distributed as dist

    from metalchat_b200 import capi, tp
    from oracle import orc

    rank, world, local = tp.env_rank_world()
    torch.cuda.set_device(local)
    dist.init_process_group("gloo")
    small = len(sys.argv) > 1 and sys.argv[1] == "small"
    if small:
        cfgd = dict(dim=512, n_layers=3, n_heads=8, n_kv_heads=2, head_dim=64, ffn_dim=1024, vocab=2000, max_seq_len=96)
        prompt, steps, golden = [3, 77, 512, 999, 0, 41, 41, 7, 1500, 2], 24, None
    else:
        g = json.loads((ROOT / "tests/golden/llama1b_L16_q0_p512_s64.json").read_text())
        cfgd = dict(dim=2048, n_layers=16, n_heads=32, n_kv_heads=8, head_dim=64, ffn_dim=8192, vocab=128256, max_seq_len=1024)
        prompt = [int(orc.lib().orc_hash_int(0x5EED, 0xFFFF, i, 0, cfgd["vocab"])) for i in range(g["prompt_len"])]
        steps, golden = g["steps"], g["tokens"]
    dev = capi.Device(local)
    m = tp.create(dev, **cfgd)
    m.init_random(0x5EED)
    m.finalize()
    single = capi.Llama(dev, capi.llama_config(**cfgd))
    single.init_random(0x5EED)
    single.finalize()
    m.prefill(prompt)
    single.prefill(prompt)
    shard = m.logits()
    shards = [None] * world
    dist.all_gather_object(shards, shard)
    full = np.concatenate(shards)
    ref = single.logits()
    f = lambda a: (a.astype(np.uint32) << 16).view(np.float32)
    err = float(np.abs(f(full) - f(ref)).max() / np.abs(f(ref)).max())
    first = int(np.lexsort((np.arange(cfgd["vocab"]), -f(ref)))[0])
    assert int(np.lexsort((np.arange(cfgd["vocab"]), -f(full)))[0]) == first, "TP argmax differs after prefill"
    t_tp, ms_tp = m.decode_loop([first], [len(prompt)], steps - 1)
    t_1, ms_1 = single.decode_loop([first], [len(prompt)], steps - 1)
    got_tp, got_1 = [first] + t_tp[:, 0].tolist(), [first] + t_1[:, 0].tolist()
    all_tp = [None] * world
    dist.all_gather_object(all_tp, got_tp)
    ok = err < 1e-2 and got_tp == got_1 and all(a == got_tp for a in all_tp) and (golden is None or got_tp == golden)
    if rank == 0:
        print(json.dumps({"world": world, "logits_max_rel_vs_single": err, "tokens_equal_single": got_tp == got_1,
                          "tokens_equal_golden": None if golden is None else got_tp == golden, "all_ranks_agree": all(a == got_tp for a in all_tp),
                          "ms_per_token_tp": ms_tp / (steps - 1), "ms_per_token_single": ms_1 / (steps - 1)}), flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
