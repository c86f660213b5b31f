"""N > 1 path: world_size-2 gloo runs.  CPU: the partition / collective / merge logic on host tensors.
GPU (-m gpu): two ranks sharing cuda:0 run the row-partitioned model against the single-GPU model."""
import os
import socket
import tempfile

import pytest
import torch
import torch.multiprocessing as mp

import _parallel_worker as W


def free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def run(fn, world=2):
    with tempfile.TemporaryDirectory() as d:
        mp.spawn(fn, args=(world, free_port(), d), nprocs=world, join=True)
        assert sorted(os.listdir(d)) == [f"ok{r}" for r in range(world)]


def test_partition_collectives_and_merge_cpu_gloo():
    run(W.cpu_collectives, 2)


def test_partition_three_ranks_cpu_gloo():
    run(W.cpu_collectives, 3)


@pytest.mark.gpu
def test_partitioned_model_matches_single_gpu():
    run(W.gpu_partitioned_model, 2)
