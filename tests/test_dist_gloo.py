"""World-size-2 `gloo` test of the N>1 host logic: deterministic sharding of work items
and the gather of 64-byte result records on rank 0 (fixed-size and ragged).  The records
come from the oracle here (no GPU in this container); on the GPU box the same gather runs
over NCCL with records written by the kernels (bench.py)."""

import os
import socket
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parents[1]


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _records_for(items, mats):
    """Oracle-made vsp_record array for the given item indices."""
    import spectral_oracle as orc

    from vision_spectra_b200 import _native as nat

    rec = np.zeros(len(items), nat.RECORD_DTYPE)
    for j, i in enumerate(items):
        m = orc.get_spectral_metrics(mats[i])
        io = orc.integer_outputs(mats[i])
        rec[j] = (i, 0, io["m"], io["start"], io["end"], io["k"], min(mats[i].shape), 0, [m[k] for k in orc.METRIC_KEYS])
    return rec


def _worker(rank, world, port, out_path):
    for p in (ROOT, ROOT / "tests", ROOT / "oracle"):
        sys.path.insert(0, str(p))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from _inputs import vit_block_matrices

    from vision_spectra_b200 import _native as nat
    from vision_spectra_b200.sweep import gather_records, gather_records_ragged, matrix_cost, partition_lpt, records_from_bytes

    rng = np.random.default_rng(11)
    mats = [w for _ in range(3) for _, w in vit_block_matrices(32, rng)]  # 18 matrices, same on every rank
    costs = [matrix_cost(*w.shape) for w in mats]
    groups = [(i // 6) * 10 + (0 if i % 6 < 3 else i % 6) for i in range(len(mats))]
    shards = partition_lpt(costs, world, groups)
    mine = shards[rank]
    local = torch.from_numpy(_records_for(mine, mats).view(np.uint8).copy())
    # ragged gather (shards may differ in length)
    full = gather_records_ragged(local, dst=0)
    # fixed-size gather (pad to equal count, as bench.py's equal shards are)
    n_max = max(len(s) for s in shards)
    padded = torch.zeros(n_max * 64, dtype=torch.uint8)
    padded[: local.numel()] = local
    fixed = gather_records(padded, dst=0)
    if rank == 0:
        rec = records_from_bytes(full)
        assert sorted(rec["item"].tolist()) == list(range(len(mats)))
        ref = _records_for(list(range(len(mats))), mats)
        got = rec[np.argsort(rec["item"])]
        assert got.tobytes() == ref.tobytes()  # bit-identical records after the exchange
        assert fixed.numel() == world * n_max * 64
        first = records_from_bytes(fixed[: len(shards[0]) * 64])
        assert first["item"].tolist() == shards[0]
        Path(out_path).write_text("ok")
    else:
        assert full is None and fixed is None
    dist.barrier()
    dist.destroy_process_group()


def test_gather_records_world2(tmp_path):
    out = tmp_path / "rank0.txt"
    mp.spawn(_worker, args=(2, _free_port(), str(out)), nprocs=2, join=True)
    assert out.read_text() == "ok"
