"""Mode drivers: ``train`` / ``test`` / ``get_teacher_emb`` on the B200 path.

Same three modes, argument namespace (``parameters.py`` flags) and on-disk formats as the
reference's ``Tiny-NewsRec/run.py`` (:20 train, :219 test, :382 get_teacher_emb):
checkpoints are ``torch.save({'model_state_dict', 'category_dict', 'word_dict',
'subcategory_dict'}, model_dir/epoch-{n}.pt)`` (run.py:205-214), teacher tables are pickled
float32 ndarrays [N+1, D] (run.py:458-459).  What changes underneath: Horovod -> NCCL
(``tinyrec.parallel``), host batch assembly -> device gathers (``tinyrec.dataloader``), the
per-impression Python metric loop -> ``tnr_eval_metrics``, the single-GPU table build ->
row-sharded build + all-gather.

Text ingestion (news.tsv tokenisation, TF ``TextLineDataset`` streaming) is outside the hot
path: callers hand in ``news_combined`` (int32 [N+1, 2L]) and an iterable
of behaviour lines or pre-parsed index arrays.
"""
import logging
import os
import random

import numpy as np
import torch

from . import checkpoint as ckpt
from . import dataloader as dl
from . import ops
from . import parallel as par
from .model_bert import Model
from .optim import Adam, DistributedOptimizer, broadcast_parameters


def apply_freeze_policy(model, trainable_layers, student=True):
    """run.py:98-112: teachers frozen; whole bert_model frozen except the listed encoder layers."""
    if student:
        for p in model.teachers.parameters():
            p.requires_grad = False
        bm = model.student.news_encoder.bert_model
    else:
        bm = model.news_encoder.bert_model
    for p in bm.parameters():
        p.requires_grad = False
    for i, layer in enumerate(bm.bert.encoder.layer):
        if i in trainable_layers:
            logging.info(f"finetune block {i}")
            for p in layer.parameters():
                p.requires_grad = True


def load_teacher_user_encoders(model, teacher_state_dicts):
    """run.py:61-70: keys ``user_encoder.*`` of teacher i -> ``teachers.{i}.*``."""
    sd = model.state_dict()
    loaded = []
    for i, tsd in enumerate(teacher_state_dicts):
        for k, v in tsd.items():
            if not k.startswith("user_encoder"):
                continue
            key = ".".join(["teachers", str(i)] + k.split(".")[1:])
            sd[key].copy_(v)
            loaded.append(key)
    return loaded


def load_student_from_first_stage(model, pretrained_state_dict):
    """run.py:76-85: every key starting with ``student`` is copied."""
    sd = model.state_dict()
    loaded = []
    for k, v in pretrained_state_dict.items():
        if k.startswith("student"):
            sd[k].copy_(v)
            loaded.append(k)
    return loaded


def accuracy(y_true, y_hat):
    """utils.py:79-83 (device scalar; no host sync)."""
    return (y_true == y_hat.argmax(dim=-1)).sum().float() / y_true.shape[0]


class IndexBatches:
    """Iterator of pre-parsed training impressions -> device index tensors.
    ``hist_idx`` int32 [n,H], ``hist_mask`` f32 [n,H], ``cand_idx`` int32 [n,K], ``label`` int64 [n];
    rank r takes batches r, r+world, ... (impression sharding, streaming.py:53-54)."""

    def __init__(self, hist_idx, hist_mask, cand_idx, label, batch_size, rank=0, world=1, device="cuda"):
        self.arr = (hist_idx, hist_mask, cand_idx, label)
        self.bs, self.rank, self.world, self.device = batch_size, rank, world, device

    def __iter__(self):
        n = self.arr[0].shape[0] // self.bs
        for b in range(self.rank, n, self.world):
            sl = slice(b * self.bs, (b + 1) * self.bs)
            yield tuple(torch.from_numpy(np.ascontiguousarray(a[sl])).to(self.device, non_blocking=True) for a in self.arr)


def batches_from_lines(lines, news_index, args, rank=0, world=1, device="cuda", seed=0):
    """Parse ``behaviors_np{K}_*.tsv`` lines (dataloader.py:118-148) into index batches."""
    rng = random.Random(seed)
    H, K = args.user_log_length, args.npratio + 1
    buf = []
    for line in lines:
        buf.append(dl.parse_train_line(line, news_index, H, args.npratio, rng))
        if len(buf) == args.batch_size:
            hist = np.array([b[0] for b in buf], dtype=np.int32)
            mask = np.array([b[1] for b in buf], dtype=np.float32)
            cand = np.array([b[2] for b in buf], dtype=np.int32).reshape(len(buf), K)
            lab = np.array([b[3] for b in buf], dtype=np.int64)
            buf = []
            yield tuple(torch.from_numpy(a).to(device, non_blocking=True) for a in (hist, mask, cand, lab))


class GraphedTrainStep:
    """The body of the reference train loop (run.py:178-197: forward, zero_grad, backward, optimizer step) captured
    ONCE into a CUDA graph and replayed per batch: the ~100 kernel launches and ~40 small torch ops of a step become
    one graph launch (no launch gaps, no Python between kernels).  Inputs are copied into static device buffers;
    the dropout seed and the Adam step counter live on the device and advance inside the graph, so every replay
    is a new step.  Warm-up steps needed before the capture run on a snapshot that is restored afterwards:
    constructing this object does not train the model.

        step = GraphedTrainStep(model, optimizer, first_batch)
        total, distill, emb, target, score = step(history, mask, candidate, label, th_list, tc_list)

    The returned tensors are static outputs overwritten by the next call."""

    def __init__(self, model, optimizer, batch, warmup=3):
        self.model, self.opt = model, optimizer
        inner = getattr(optimizer, "opt", optimizer)
        inner.make_capturable()
        self.static = self._clone(batch)
        st = model.train_state()
        inner.ensure_state()
        ne = model.student.news_encoder if hasattr(model, "student") else model.news_encoder
        drop = ne.drop_state(st.flat.data.device)
        snap = dict(data=st.flat.data.clone(), grad=st.flat.grad.clone(), m=inner.m.clone(), v=inner.v.clone(),
                    vmax=inner.vmax.clone(), step=inner.step_dev.clone(), seed=drop.seed.clone())
        cur = torch.cuda.current_stream()
        side = torch.cuda.Stream()
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            for _ in range(max(1, warmup)):                # first-call work: workspaces, smem attributes, NCCL buffers
                self._eager(self.static)
        cur.wait_stream(side)
        torch.cuda.synchronize()
        launches0 = ops.stats.launches
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph, capture_error_mode="thread_local"):
            self.out = self._eager(self.static)
        self.launches_per_step = ops.stats.launches - launches0
        # roll the warm-up steps back: parameters, optimizer state, step counter, dropout seed
        st.flat.data.copy_(snap["data"])
        st.flat.grad.copy_(snap["grad"])
        st.flat.refresh_shadow()
        inner.m.copy_(snap["m"])
        inner.v.copy_(snap["v"])
        inner.vmax.copy_(snap["vmax"])
        inner.step_dev.copy_(snap["step"])
        drop.seed.copy_(snap["seed"])
        torch.cuda.synchronize()

    def close(self):
        """Release the captured graph.  Call this before ``torch.distributed.destroy_process_group()``: the graph
        holds the NCCL all-reduce kernels of the step, and tearing the communicator down under a live graph hangs."""
        self.graph = None
        self.out = None
        torch.cuda.synchronize()

    @staticmethod
    def _clone(batch):
        h, m, c, l, th, tc = batch
        dev = lambda t: t.detach().to("cuda", copy=True) if not t.is_cuda else t.detach().clone()  # noqa: E731
        return (dev(h), dev(m), dev(c), dev(l), [dev(t) for t in th], [dev(t) for t in tc])

    def _eager(self, batch):
        self.opt.zero_grad()
        out = self.model(*batch)
        out[0].backward()
        self.opt.step()
        return tuple(o.detach() for o in out)

    def __call__(self, history, history_mask, candidate, label, teacher_history_embs, teacher_candidate_embs):
        s = self.static
        s[0].copy_(history, non_blocking=True)
        s[1].copy_(history_mask, non_blocking=True)
        s[2].copy_(candidate, non_blocking=True)
        s[3].copy_(label, non_blocking=True)
        for d, t in zip(s[4], teacher_history_embs):
            d.copy_(t, non_blocking=True)
        for d, t in zip(s[5], teacher_candidate_embs):
            d.copy_(t, non_blocking=True)
        self.graph.replay()
        ops.bump_param_generation()          # the replay stepped the optimizer: cached parameter copies are stale
        return self.out


def train(args, news_combined, teacher_embs, batches, model=None, category_dict=None, subcategory_dict=None,
          word_dict=None):
    """run.py:20-216.  ``batches`` yields (hist_idx, hist_mask, cand_idx, label) device tensors
    (see IndexBatches / batches_from_lines).  Returns the trained model."""
    world, rank, local = par.init_distributed(getattr(args, "enable_hvd", True))
    device = torch.device("cuda", local)
    if model is None:
        model = Model(args)
    model = model.to(device)
    apply_freeze_policy(model, list(args.bert_trainable_layer))
    optimizer = Adam(model, lr=args.lr, amsgrad=True)                       # run.py:134
    if world > 1:
        broadcast_parameters(model, root_rank=0)                            # run.py:142-143
        optimizer = DistributedOptimizer(optimizer)                         # run.py:144-149
    tables = dl.DeviceTables(news_combined, teacher_embs, device)
    K = args.npratio + 1
    batcher = dl.TrainBatcher(tables, args.batch_size, args.user_log_length, K)
    logging.info("Training...")
    for ep in range(getattr(args, "start_epoch", 0), args.epochs):
        loss_sum = torch.zeros((), device=device)
        acc_sum = torch.zeros((), device=device)
        cnt = -1
        for cnt, (hist_idx, hist_mask, cand_idx, label) in enumerate(batches):
            if cnt > args.max_steps_per_epoch:                              # run.py:176 (runs max+1 steps)
                break
            history, candidate, th, tc = batcher.assemble(hist_idx, cand_idx)
            total, distill, emb, target, y = model(history, hist_mask, candidate, label, th, tc)
            loss_sum += total.detach()
            acc_sum += accuracy(label, y.detach())
            optimizer.zero_grad()
            total.backward()
            optimizer.step()
            if cnt % args.log_steps == 0 and cnt > 0:
                logging.info("[{}] Ed: {}, train_loss: {:.5f}, acc: {:.5f}".format(
                    rank, cnt * args.batch_size, float(loss_sum) / cnt, float(acc_sum) / cnt))
        if rank == 0 and getattr(args, "model_dir", None):
            os.makedirs(args.model_dir, exist_ok=True)
            ckpt_path = os.path.join(args.model_dir, f"epoch-{ep + 1}.pt")
            ckpt.save_checkpoint(ckpt_path, model, category_dict, word_dict, subcategory_dict)       # run.py:205-214
            logging.info(f"Model saved to {ckpt_path}")
    return model


@torch.no_grad()
def build_news_table(news_encoder, news_combined, batch_size=2048, device=None, sharded=True):
    """run.py:261-290 / :432-447: encode every row of ``news_combined`` (row 0 = pad news included).
    Rows are sharded contiguously over ranks and all-gathered at the end -> fp32 [N+1, D] on device."""
    device = device or next(news_encoder.parameters()).device
    nc = torch.as_tensor(np.ascontiguousarray(news_combined))
    n = nc.shape[0]
    world, rank = (par.world_size(), par.get_rank()) if sharded else (1, 0)
    lo, hi = par.shard_rows(n, rank, world)
    D = news_encoder.dense.weight.shape[0]
    out = torch.empty(hi - lo, D, device=device, dtype=torch.float32)
    tab = nc[lo:hi].to(device)
    for s in range(0, hi - lo, batch_size):
        x = tab[s:s + batch_size].to(torch.int64)           # torch.LongTensor(arr), run.py:276-277
        out[s:s + x.shape[0]] = news_encoder(x)
    return par.allgather_rows(out, n) if world > 1 else out


@torch.no_grad()
def evaluate(user_encoder, news_scoring, hist_idx, hist_mask, cand_ptr, cand_idx, labels, batch_size=4096,
             return_per_impression=False):
    """run.py:301-379 on the device.  Impressions are CSR-packed: ``cand_ptr`` int64 [n+1],
    ``cand_idx`` int32 [nnz], ``labels`` int8 [nnz]; rank r scores impressions r-th contiguous slice.
    Returns (mean [auc, mrr, ndcg5, ndcg10] over ALL impressions as the reference's final line does,
    total impression count)."""
    device = news_scoring.device
    n = hist_idx.shape[0]
    world, rank = par.world_size(), par.get_rank()
    lo, hi = par.shard_rows(n, rank, world)
    sums = torch.zeros(5, device=device, dtype=torch.float64)
    per_all = []
    ptr_h = np.asarray(cand_ptr)
    # Index batches go host -> pinned staging -> device on a copy stream, one batch ahead of the scoring kernels
    # (the reference's loader thread did the same job with blocking copies, dataloader.py:303-314); nothing is read
    # back until the final reduction.
    copy_stream = torch.cuda.Stream(device=device)
    main = torch.cuda.current_stream(device)

    def stage(s):
        e = min(hi, s + batch_size)
        p0, p1 = int(ptr_h[s]), int(ptr_h[e])
        host = (np.ascontiguousarray(hist_idx[s:e]), np.ascontiguousarray(hist_mask[s:e]),
                np.ascontiguousarray(ptr_h[s:e + 1] - p0), np.ascontiguousarray(cand_idx[p0:p1]),
                np.ascontiguousarray(labels[p0:p1]))
        max_c = int(np.diff(host[2]).max()) if e > s else 0
        with torch.cuda.stream(copy_stream):
            pinned = [torch.from_numpy(a).pin_memory() for a in host]
            dev = [t.to(device, non_blocking=True) for t in pinned]
            ready = torch.cuda.Event()
            ready.record(copy_stream)
        return dict(e=e, dev=dev, pinned=pinned, max_c=max_c, ready=ready)

    nxt = stage(lo) if hi > lo else None
    s = lo
    while nxt is not None:
        cur = nxt
        e = cur["e"]
        nxt = stage(e) if e < hi else None                       # next batch's copies overlap this batch's kernels
        main.wait_event(cur["ready"])
        hi_t, hm_t, ptr_t, cand_t, lab_t = cur["dev"]
        for t in cur["dev"]:
            t.record_stream(main)
        if hasattr(user_encoder, "forward_gather"):      # news_scoring[log_ids] (dataloader.py:295) fused into the kernel
            user = user_encoder.forward_gather(news_scoring, hi_t, hm_t)
        else:
            user = user_encoder(dl.gather_history_vecs(news_scoring, hi_t), hm_t)
        per = torch.zeros(e - s, 5, device=device, dtype=torch.float64)
        ops.eval_metrics(news_scoring, user, ptr_t, cand_t, lab_t, cur["max_c"], per, sums)
        if return_per_impression:
            per_all.append(per)
        s = e
    mean, total = par.reduce_eval_sums(hi - lo, sums[:4])
    if return_per_impression:
        return mean, total, torch.cat(per_all) if per_all else torch.zeros(0, 5, dtype=torch.float64)
    return mean, total


def doc_sim(news_scoring, n_pairs=1000000, rng=random):
    """run.py:292-299: mean cosine similarity of ``n_pairs`` random row pairs drawn from rows 1..N with
    ``random.randrange`` exactly as the reference draws them (pairs with i == j add nothing but still count in the
    mean).  The pair list is drawn on the host, the 2 x n_pairs row reads and cosines run in one kernel."""
    n = news_scoring.shape[0]
    pairs = np.empty((n_pairs, 2), dtype=np.int32)
    for k in range(n_pairs):
        pairs[k, 0] = rng.randrange(1, n)
        pairs[k, 1] = rng.randrange(1, n)
    total = torch.zeros(1, device=news_scoring.device, dtype=torch.float64)
    ops.doc_sim(news_scoring, torch.from_numpy(pairs).to(news_scoring.device), total)
    return float(total.item()) / n_pairs


def test(args, model, news_combined, hist_idx, hist_mask, cand_ptr, cand_idx, labels):
    """run.py:219-379: student news table, then impression scoring with user_log_mask as in args."""
    world, rank, local = par.init_distributed(getattr(args, "enable_hvd", True))
    model.eval()
    news_scoring = build_news_table(model.student.news_encoder, news_combined, batch_size=args.batch_size * 16)
    logging.info("news scoring num: {}".format(news_scoring.shape[0]))
    if rank == 0 and getattr(args, "doc_sim", True):
        print(f"=== doc-sim: {doc_sim(news_scoring, getattr(args, 'doc_sim_pairs', 1000000))} ===")      # run.py:292-299
    mean, total = evaluate(model.student.user_encoder, news_scoring, hist_idx, hist_mask, cand_ptr, cand_idx, labels)
    if rank == 0:
        logging.info("[{}] Ed: {}: {}".format(rank, total, "\t".join("{:0.2f}".format(float(x) * 100) for x in mean)))
    return mean, total


def get_teacher_emb(args, teacher_state_dicts, news_combined, out_paths=None):
    """run.py:382-460: one [N+1, D] float32 table per teacher checkpoint, pickled like the reference."""
    from .model_bert_2 import ModelBert
    world, rank, local = par.init_distributed(getattr(args, "enable_hvd", True))
    device = torch.device("cuda", local)
    tables = []
    for i, tsd in enumerate(teacher_state_dicts):
        model = ModelBert(args)
        if tsd is not None:
            model.load_state_dict(tsd)
        model = model.to(device).eval()
        table = build_news_table(model.news_encoder, news_combined, batch_size=args.batch_size * 64)
        arr = table.cpu().numpy()
        logging.info("news scoring num: {}".format(arr.shape[0]))
        if out_paths is not None and rank == 0:
            ckpt.save_teacher_table(out_paths[i], arr)                                          # run.py:458-459
            logging.info(f"teacher embedding saved at {out_paths[i]}")
        tables.append(arr)
        del model
    return tables
