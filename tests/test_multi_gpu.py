"""Two-rank NCCL data-parallel test: one KD train step on 2 GPUs with the
impressions sharded by rank and the bucketed gradient all-reduce must move the parameters exactly
like one GPU stepping on the whole batch.  Needs >= 2 GPUs (`gpurun --gpus 2 -- pytest -m gpu
tests/test_multi_gpu.py`); skipped on a single-GPU box."""
import os
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _problem(B, seed=7):
    import tinyrec.synth as synth
    H, K, L, M, D = 10, 5, 16, 2, 256
    news = synth.news_table(400, L=L, seed=3)
    hist_idx, hmask, cand_idx, label = synth.train_impressions(B, 400, H, K, seed=seed)
    tables = synth.teacher_tables(400, M, D, seed=5)
    history = torch.from_numpy(news[hist_idx].astype(np.int64))
    candidate = torch.from_numpy(news[cand_idx].astype(np.int64))
    th = [torch.from_numpy(t[hist_idx]) for t in tables]
    tc = [torch.from_numpy(t[cand_idx]) for t in tables]
    return (history, torch.from_numpy(hmask), candidate, torch.from_numpy(label), th, tc), (H, M)


def _model(H, M, device):
    import tinyrec.model_bert as mb
    import tinyrec.synth as synth
    layers = 2
    args = synth.demo_args(num_student_layers=layers, num_teachers=M, user_log_length=H)
    m = mb.Model(args)
    m.load_state_dict(synth.kd_model_state(layers, M, 11, noisy=True), strict=True)
    m.to(device).eval()                      # dropout off: ranks must be comparable with the 1-GPU run
    for p in m.teachers.parameters():
        p.requires_grad = False
    bm = m.student.news_encoder.bert_model
    for p in bm.parameters():
        p.requires_grad = False
    for layer in bm.bert.encoder.layer:
        for p in layer.parameters():
            p.requires_grad = True
    return m


def _slice(batch, lo, hi, device):
    h, hm, c, lab, th, tc = batch
    return (h[lo:hi].to(device), hm[lo:hi].to(device), c[lo:hi].to(device), lab[lo:hi].to(device),
            [t[lo:hi].to(device) for t in th], [t[lo:hi].to(device) for t in tc])


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    import tinyrec.optim as topt
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    B = 8
    batch, (H, M) = _problem(B)
    m = _model(H, M, dev)
    opt = topt.DistributedOptimizer(topt.Adam(m, lr=1e-3))
    topt.broadcast_parameters(m, 0)
    per = B // world
    for _ in range(2):
        opt.zero_grad()
        loss = m(*_slice(batch, rank * per, (rank + 1) * per, dev))[0]
        loss.backward()
        opt.step()
    torch.cuda.synchronize()
    flat = m.train_state().flat
    if rank == 0:
        torch.save(flat.data.cpu(), out)
    gathered = [torch.empty_like(flat.data) for _ in range(world)]
    dist.all_gather(gathered, flat.data)
    assert all(torch.equal(g, gathered[0]) for g in gathered), "replicas diverged"
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_rank_step_equals_single_gpu_step(tmp_path):
    import torch.multiprocessing as mp
    import tinyrec.optim as topt
    out = str(tmp_path / "dp.pt")
    mp.spawn(_worker, args=(2, 29533, out), nprocs=2, join=True)
    dp = torch.load(out)
    dev = torch.device("cuda", 0)
    B = 8
    batch, (H, M) = _problem(B)
    m = _model(H, M, dev)
    opt = topt.Adam(m, lr=1e-3)
    p0 = m.train_state().flat.data.clone()
    for _ in range(2):
        opt.zero_grad()
        m(*_slice(batch, 0, B, dev))[0].backward()
        opt.step()
    single = m.train_state().flat.data.cpu()
    moved = (single - p0.cpu()).norm()
    diff = (dp - single).norm()
    assert float(moved) > 0
    # Adam normalises the update, so compare the parameter movement: bf16 gradient noise between a batch
    # of 8 and two batches of 4 stays far below the update itself
    assert float(diff) < 0.05 * float(moved), (float(diff), float(moved))
