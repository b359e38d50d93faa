"""Two-rank NCCL tests.  (1) One KD train step on 2 GPUs with the
impressions sharded by rank and the bucketed gradient all-reduce must move the parameters exactly
like one GPU stepping on the whole batch.  (2) The sharded drivers: ``run.build_news_table`` (rows sharded +
all-gather) and ``run.evaluate`` (impressions sharded + 5-double all-reduce) return what the unsharded
calls return; ``run.train`` from ``IndexBatches`` with an odd number of batches runs the same number of
steps on both ranks and leaves identical replicas.  Needs >= 2 GPUs (`gpurun --gpus 2 -- pytest -m gpu
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
        print("gradient exchange:", "tnr_allreduce_p2p" if flat.symm is not None and opt.use_p2p else f"nccl ({flat.symm_error})")
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


# ---------------------------------------------------------------------------------------------- sharded drivers
def _driver_problem():
    import tinyrec.synth as synth
    layers, L, n_news, H, D = 2, 12, 300, 10, 256
    news = synth.news_table(n_news, L=L, seed=13)                       # 301 rows: odd, so the shards are ragged
    sd = synth.model_bert_state("", layers, 17, noisy=True)
    eh, em, ptr, cand, lab = synth.eval_impressions(501, n_news, H, seed=19)
    return layers, H, news, sd, (eh, em, ptr, cand, lab)


def _driver_model(layers, H, sd, device, ulm=True):
    import tinyrec.model_bert as mb
    import tinyrec.synth as synth
    m = mb.ModelBert(synth.demo_args(num_student_layers=layers, user_log_length=H, user_log_mask=ulm))
    m.load_state_dict(sd, strict=True)
    return m.to(device).eval()


def _driver_worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    import tinyrec.model_bert as mb
    import tinyrec.run as trun
    import tinyrec.synth as synth
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    layers, H, news, sd, ev = _driver_problem()
    m = _driver_model(layers, H, sd, dev)
    table = trun.build_news_table(m.news_encoder, news, batch_size=64)            # rows sharded, all-gathered
    mean, total = trun.evaluate(m.user_encoder, table, *ev, batch_size=128)       # impressions sharded, all-reduced
    # run.train on 2 ranks: 5 batches -> 2 steps per rank (the fifth is dropped on both), replicas stay identical
    B, Hh, K, L, M, D, tl = 4, 6, 3, 12, 2, 256, 2
    tnews = synth.news_table(300, L=L, seed=21)
    hist_idx, hmask, cand_idx, label = synth.train_impressions(5 * B, 300, Hh, K, seed=22)
    tables = synth.teacher_tables(300, M, D, seed=23)
    args = synth.demo_args(num_student_layers=tl, num_teachers=M, user_log_length=Hh, npratio=K - 1, batch_size=B,
                           bert_trainable_layer=[1], lr=1e-3, epochs=2, max_steps_per_epoch=100, log_steps=100,
                           enable_hvd=True, model_dir=None)
    model = mb.Model(args)
    model.load_state_dict(synth.kd_model_state(tl, M, 24 + rank, noisy=True), strict=True)    # ranks differ: broadcast must fix it
    model.to(dev).eval()
    batches = trun.IndexBatches(hist_idx, hmask, cand_idx, label, B, rank=rank, world=world, device=dev)
    hist = []
    trun.train(args, tnews, tables, batches, model=model, history=hist)
    flat = model.train_state().flat
    gathered = [torch.empty_like(flat.data) for _ in range(world)]
    dist.all_gather(gathered, flat.data)
    assert all(torch.equal(g, gathered[0]) for g in gathered), "replicas diverged"
    if rank == 0:
        torch.save(dict(table=table.cpu(), mean=mean.cpu(), total=total, steps=len(hist), flat=flat.data.cpu()), out)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_sharded_drivers_equal_unsharded(tmp_path):
    import torch.multiprocessing as mp
    import tinyrec.run as trun
    out = str(tmp_path / "drv.pt")
    mp.spawn(_driver_worker, args=(2, 29541, out), nprocs=2, join=True)
    got = torch.load(out)
    dev = torch.device("cuda", 0)
    layers, H, news, sd, ev = _driver_problem()
    m = _driver_model(layers, H, sd, dev)
    table = trun.build_news_table(m.news_encoder, news, batch_size=64)
    assert torch.equal(got["table"], table.cpu())                       # row sharding + all-gather is invisible
    mean, total = trun.evaluate(m.user_encoder, table, *ev, batch_size=128)
    assert got["total"] == total == 501
    assert torch.allclose(got["mean"], mean.cpu(), rtol=0, atol=1e-12)
    assert got["steps"] == 4                                            # 2 epochs x 2 steps per rank (5 batches, world 2)
