"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: unique-id exchange, row sharding, sharded sweeps."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, tmp):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from multimodalautoencoder_b200 import dp
    from multimodalautoencoder_b200.autoencoder_wrapper import MMAEWrapper
    from multimodalautoencoder_b200.data_funcs import DataLoader
    from multimodalautoencoder_b200.synthetic import make_frame
    # 1. the 128-byte id reaches every rank unchanged
    uid = dp.exchange_unique_id(lambda: bytes(range(128)), rank, dist)
    assert uid == bytes(range(128))
    # 2. contiguous row shards tile the global batch
    first, rows = dp.shard_rows(65536, rank, world)
    t = torch.tensor([first, rows])
    got = [torch.zeros(2, dtype=torch.long) for _ in range(world)]
    dist.all_gather(got, t)
    assert [int(g[0]) for g in got] == [r * 65536 // world for r in range(world)] and sum(int(g[1]) for g in got) == 65536
    # 3. a sharded grid sweep: each rank writes its own CSV with its share of the settings; rank 0 merges
    df = make_frame(200, seed=1)
    dl = DataLoader(df=df, supervised=False, cross_validation=True, normalize_and_fill=False, suppress_output=True)
    cdl = DataLoader(df=df, supervised=True, cross_validation=True, normalize_and_fill=False, suppress_output=True)

    class Fake(MMAEWrapper):            # no GPU here: replace the fit by a deterministic score
        def get_cross_validation_results(self, param_dict):
            param_dict[self.optimize_for] = float(len(str(sorted(param_dict.items(), key=str))))
            return param_dict

    w = Fake('synthetic.csv', dropbox_path=tmp + '/', data_loader=dl, classification_data_loader=cdl, shard=(rank, world))
    w.sweep_all_parameters()
    assert len(w.val_results_df) == 54
    dist.barrier()
    if rank == 0:
        merged = w.merge_shard_results()
        assert len(merged) == 108 and len(merged.drop_duplicates()) == 108
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_host_logic(tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)


def test_shard_rows_rejects_ragged():
    from multimodalautoencoder_b200 import dp
    with pytest.raises(ValueError):
        dp.shard_rows(10, 0, 3)
    assert dp.shard_rows(12, 2, 3) == (8, 4)
