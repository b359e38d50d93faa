"""CPU-only tests: the C-ABI library builds, loads and exports every symbol the header
declares; host-side logic (index parsing / padding, rel-pos buckets, synthetic init,
state_dict keys, sharding + collectives over gloo with world_size 2)."""
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="session")
def built_lib():
    import __graft_entry__ as ge
    import tinyrec._lib as L
    if not os.path.exists(L.LIB_PATH):
        ge.build()
    return L.LIB_PATH


def _header_symbols():
    txt = open(os.path.join(ROOT, "include", "tinyrec.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(tnr_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_header_symbol(built_lib):
    import tinyrec._lib as L
    out = subprocess.run(["nm", "-D", "--defined-only", built_lib], capture_output=True, text=True, check=True).stdout
    exported = set(re.findall(r" T (tnr_[a-z0-9_]+)", out))
    declared = _header_symbols()
    assert declared, "no declarations parsed from include/tinyrec.h"
    missing = [s for s in declared if s not in exported]
    assert not missing, f"declared in tinyrec.h but not exported: {missing}"
    assert sorted(L.exported_names()) == declared, "ctypes binding table and header disagree"
    lib = L.load()
    assert lib.tnr_abi_version() == 6


def test_ctypes_signatures_match_header_declarations():
    """Every C-ABI entry point: the ctypes binding (tinyrec/_lib.py) declares as many arguments, of pointer / integer /
    float kind, as include/tinyrec.h does -- a mismatch would corrupt the call frame silently."""
    import ctypes
    import tinyrec._lib as L
    txt = open(os.path.join(ROOT, "include", "tinyrec.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    decls = re.findall(r"\b(?:int|long long|const char\*)\s+(tnr_[a-z0-9_]+)\s*\(([^;]*?)\)\s*;", txt, flags=re.S)
    assert len(decls) >= 30
    for name, args in decls:
        if name == "tnr_last_error":
            continue
        args = args.strip()
        params = [] if args in ("", "void") else [a.strip() for a in args.split(",")]
        sig = L._SIGNATURES[name][0]
        assert len(sig) == len(params), (name, len(sig), len(params))
        for ct, decl in zip(sig, params):
            if "*" in decl:
                assert ct is ctypes.c_void_p or hasattr(ct, "contents") or ct is ctypes.c_char_p, (name, decl, ct)
            elif decl.startswith("float"):
                assert ct is ctypes.c_float, (name, decl, ct)
            elif decl.startswith("long long"):
                assert ct is ctypes.c_int64, (name, decl, ct)
            elif decl.startswith("int"):
                assert ct is ctypes.c_int, (name, decl, ct)


def test_ops_fail_loudly_without_cuda(built_lib):
    import tinyrec._lib as L
    import tinyrec.ops as ops
    a = torch.zeros(64, 64, dtype=torch.bfloat16)
    with pytest.raises(L.TinyRecError):
        ops.gemm(a, a, torch.zeros(64, 64, dtype=torch.bfloat16))
    import tinyrec.model_bert as mb
    import tinyrec.synth as synth
    ne = mb.NewsEncoder(synth.demo_args(num_student_layers=1))
    with pytest.raises(L.TinyRecError):
        ne(torch.zeros(2, 60, dtype=torch.int64))       # CPU tensor -> no fallback


def test_missing_library_is_an_error(monkeypatch, tmp_path):
    import tinyrec._lib as L
    monkeypatch.setattr(L, "_lib", None)
    monkeypatch.setattr(L, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(L.TinyRecError):
        L.load()


def test_relpos_bucket_table_matches_reference(golden):
    from tinyrec.engine import rel_pos_bucket_table
    g = golden("relpos")
    lut = dict(zip(g["rel"].tolist(), g["bucket"].tolist()))
    for L in (1, 9, 30, 32, 180, 512):
        t = rel_pos_bucket_table(L)
        pos = np.arange(L)
        rel = pos[None, :] - pos[:, None]
        want = np.vectorize(lut.get)(rel)
        assert np.array_equal(t.numpy(), want)


def test_state_dict_keys_match_reference_layout():
    import tinyrec.model_bert as mb
    import tinyrec.model_bert_2 as mb2
    import tinyrec.synth as synth
    m = mb.Model(synth.demo_args(num_student_layers=4, num_teachers=4))
    want = synth.kd_model_state(4, 4, 0)
    got = m.state_dict()
    assert list(got.keys()) == list(want.keys())           # names AND order (113 keys)
    assert len(got) == 113
    for k in want:
        assert tuple(got[k].shape) == tuple(want[k].shape), k
    # NRMS variant (model_bert.py:145-148): W_Q / W_K / W_V in front of the pooling, same key order as the reference
    n = mb.Model(synth.demo_args(num_student_layers=1, num_teachers=2, model="NRMS", num_attention_heads=16))
    assert list(n.state_dict().keys()) == list(synth.kd_model_state(1, 2, 0, model="NRMS", n_heads=16).keys())
    t = mb2.ModelBert(synth.demo_args(num_hidden_layers=2))
    assert list(t.state_dict().keys()) == list(synth.model_bert_state("", 2, 0).keys())
    # attribute paths used by run.py:101-112,285,343,442
    assert len(m.student.news_encoder.bert_model.bert.encoder.layer) == 4
    assert hasattr(m.student, "user_encoder") and hasattr(m, "teachers") and hasattr(m, "transform_matrix")
    n_train = 0
    for p in m.teachers.parameters():
        p.requires_grad = False
    for p in m.student.news_encoder.bert_model.parameters():
        p.requires_grad = False
    for i in (2, 3):
        for p in m.student.news_encoder.bert_model.bert.encoder.layer[i].parameters():
            p.requires_grad = True
    trainable = [p for p in m.parameters() if p.requires_grad]
    assert len(trainable) == 51 and sum(p.numel() for p in trainable) == 14841634


def test_post_train_wrappers_keep_the_notebook_names():
    """Post-train_KD.ipynb cells 12 / 14 and Domian-specific_Post-train.ipynb cell 11: attribute paths and state_dict
    keys (``student.news_encoder.*`` + ``transform_matrix.{i}.*``; ``news_encoder.*``) and the freeze policy of cell 17."""
    from types import SimpleNamespace
    import tinyrec.post_train as pt
    import tinyrec.synth as synth
    args = SimpleNamespace(news_query_vector_dim=200, news_dim=256, num_hidden_layers=2, num_teachers=3)
    m = pt.DistillModel(args)
    full = synth.kd_model_state(2, 3, 0)
    want = [k for k in full if k.startswith(("student.news_encoder.", "transform_matrix."))]
    assert list(m.state_dict().keys()) == want
    for p in m.student.news_encoder.bert_model.parameters():
        p.requires_grad = False
    for index, layer in enumerate(m.student.news_encoder.bert_model.bert.encoder.layer):
        if index in [1]:
            for p in layer.parameters():
                p.requires_grad = True
    d = pt.DomainTitleBodySimModel(args)
    assert list(d.state_dict().keys()) == [k for k in synth.model_bert_state("", 2, 0) if k.startswith("news_encoder.")]
    with pytest.raises(Exception):
        m(torch.zeros(1, 2, 8, dtype=torch.int64), torch.zeros(1, 8, dtype=torch.int64), torch.zeros(1, dtype=torch.int64),
          [torch.zeros(1, 2, 256)] * 3, [torch.zeros(1, 256)] * 3)            # CPU tensors: no fallback path


def test_loader_index_logic_bit_exact(golden):
    import random
    import tinyrec.dataloader as dl
    g = golden("batching")
    H, npratio = int(g["H"]), int(g["npratio"])
    n_news = g["news_combined"].shape[0] - 1
    news_index = {f"N{i}": i for i in range(1, n_news + 1)}
    random.seed(123)
    hist, masks, cands, labels = [], [], [], []
    for clicks, pos, neg in zip(g["clicks"], g["pos"], g["neg"]):
        line = "\t".join(["0", "U1", "t", str(clicks), str(pos), str(neg)]).encode()
        h, m, c, lab = dl.parse_train_line(line, news_index, H, npratio)
        hist.append(h); masks.append(m); cands.append(c); labels.append(lab)
    assert np.array_equal(g["news_combined"][np.array(hist)].astype(np.int64), g["user_feature"])
    assert np.array_equal(g["news_combined"][np.array(cands)].astype(np.int64), g["news_feature"])
    assert np.array_equal(np.array(masks, dtype=np.float32), g["log_mask"])
    assert np.array_equal(np.array(labels), g["label"])
    assert dl.pad_to_fix_len([1, 2, 3], 2) == ([2, 3], [1, 1])
    assert dl.pad_to_fix_len([], 3) == ([0, 0, 0], [0, 0, 0])
    h, m, c, labs = dl.parse_eval_line("1\tU\tt\tN1 N2\tN3-1 N4-0 NX-0", news_index, 4)
    assert h == [0, 0, 1, 2] and m == [0, 0, 1, 1] and c == [3, 4, 0] and labs == [1, 0, 0]


def test_synth_shapes_and_determinism():
    import tinyrec.synth as synth
    t = synth.news_table(100, L=30, seed=1)
    assert t.shape == (101, 60) and t.dtype == np.int32 and not t[0].any()
    assert (t[1:, 0] == 101).all() and ((t[:, :30] > 0) == (t[:, 30:] == 1)).all()
    h, m, c, y = synth.train_impressions(64, 100, 50, 5, seed=2)
    assert h.shape == (64, 50) and ((h > 0) == (m > 0)).all() and (m[:, -1] == 1).all()
    assert c.shape == (64, 5) and y.max() < 5
    hh, mm, ptr, cand, lab = synth.eval_impressions(200, 100, seed=3)
    for i in range(200):
        seg = lab[ptr[i]:ptr[i + 1]]
        assert 0 < seg.sum() < len(seg)
    a = synth.kd_model_state(1, 1, 0)["student.news_encoder.dense.weight"]
    b = synth.kd_model_state(1, 1, 0)["student.news_encoder.dense.weight"]
    assert torch.equal(a, b)


def test_shard_helpers():
    import tinyrec.parallel as par
    assert par.shard_files(list("abcdefg"), 1, 3) == ["b", "e"]
    cover = []
    for r in range(8):
        lo, hi = par.shard_rows(161014, r, 8)
        cover += list(range(lo, hi))[:1] + [hi]
    assert par.shard_rows(161014, 0, 8)[0] == 0 and par.shard_rows(161014, 7, 8)[1] == 161014
    assert sum(par.shard_rows(10, r, 4)[1] - par.shard_rows(10, r, 4)[0] for r in range(4)) == 10


_GLOO_WORKER = r"""
import os, sys
sys.path.insert(0, {root!r})
import torch, torch.distributed as dist
import tinyrec.parallel as par
from oracle import optim as oopt
world, rank, local = par.init_distributed(backend="gloo")
assert world == 2
g = torch.Generator().manual_seed(7)
grads = [torch.randn(1000, generator=g) for _ in range(2)]          # identical on both ranks
mine = grads[rank].clone()
par.allreduce_mean_(mine)
assert torch.allclose(mine, oopt.allreduce_average(grads), atol=1e-7)
# SUM all-reduce + grad_scale = 1/world (what DistributedOptimizer does) == average
s = grads[rank].clone(); par.allreduce_sum_(s)
assert torch.allclose(s * 0.5, oopt.allreduce_average(grads), atol=1e-7)
# table build: shard rows, all-gather, equals the unsharded table
table = torch.arange(11 * 4, dtype=torch.float32).reshape(11, 4)
lo, hi = par.shard_rows(11, rank, world)
full = par.allgather_rows(table[lo:hi].clone(), 11)
assert torch.equal(full, table)
# eval reduction (run.py:372-379): sums over ranks / total count
mean, total = par.reduce_eval_sums(3 + rank, torch.tensor([1.0, 2.0, 3.0, 4.0]) * (rank + 1))
assert total == 7 and torch.allclose(mean, torch.tensor([3.0, 6.0, 9.0, 12.0], dtype=torch.float64) / 7)
assert par.shard_files(list(range(5)), rank, world) == list(range(rank, 5, 2))
dist.destroy_process_group()
sys.stdout.write("ok%d\n" % rank); sys.stdout.flush()
"""


def test_two_rank_gloo_collectives(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(_GLOO_WORKER.format(root=ROOT))
    import socket
    with socket.socket() as sk:                       # a free port: a fixed one can still be in TIME_WAIT
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", str(port), str(script)],
                       capture_output=True, text=True, timeout=240, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "ok0" in r.stdout and "ok1" in r.stdout


# ------------------------------------------------------------------ on-disk formats
def test_unilm_state_dict_conversion_matches_reference(golden):
    """tinyrec.checkpoint.convert_unilm_state_dict vs the reference's converter run on the same
    UniLM-format dict (fixture made by tests/golden/make_golden.py:gen_unilm_convert)."""
    import tinyrec.checkpoint as ck
    g = golden("unilm_convert")
    sd = {k[3:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("in/")}
    want = {k[4:]: g[k] for k in g.files if k.startswith("out/")}
    got = ck.convert_unilm_state_dict(sd)
    assert list(got.keys()) == [str(k) for k in g["out_keys"]]            # same keys, same order
    for k, v in want.items():
        assert got[k].shape == v.shape and np.array_equal(got[k].numpy(), v), k
    assert float(got["bert.encoder.layer.0.attention.self.key.bias"].abs().sum()) == 0.0


def test_position_resize_and_nonstrict_bin_load(tmp_path):
    """tnlrv3/modeling.py:88-118 semantics + loading a deeper .bin into a shallower encoder."""
    import tinyrec.checkpoint as ck
    old = torch.arange(12 * 4, dtype=torch.float32).reshape(12, 4)
    k = "bert.embeddings.position_embeddings.weight"
    grown = ck.resize_position_embeddings({k: old.clone()}, 30, reuse_position_embedding=True)[k]
    assert grown.shape == (30, 4) and torch.equal(grown[:12], old) and torch.equal(grown[12:24], old) \
        and torch.equal(grown[24:], old[:6])
    once = ck.resize_position_embeddings({k: old.clone()}, 30)[k]
    assert torch.equal(once[:12], old) and not torch.equal(once[12:24], old) and float(once[12:].abs().max()) < 0.2
    cut = ck.resize_position_embeddings({k: old.clone()}, 5)[k]
    assert torch.equal(cut, old[:5])
    assert ck.strip_prefix({"unilm.a": 1, "b": 2}, "unilm.") == {"a": 1, "b": 2}

    class Tiny(torch.nn.Module):                        # stands in for TuringNLRv3ForSequenceClassification
        def __init__(self):
            super().__init__()
            self.bert = torch.nn.Module()
            self.bert.embeddings = torch.nn.Module()
            self.bert.embeddings.position_embeddings = torch.nn.Embedding(20, 4)
            self.bert.rel_pos_bias = torch.nn.Linear(32, 2, bias=False)
            self.classifier = torch.nn.Linear(4, 2)
    m = Tiny()
    unilm = {k: old.clone(), "bert.encoder.rel_pos_bias.weight": torch.ones(2, 32),
             "bert.encoder.layer.11.output.dense.weight": torch.zeros(4, 16)}
    path = tmp_path / "unilm.bin"
    torch.save(unilm, path)
    missing, unexpected = ck.load_unilm_bin(m, str(path))
    assert torch.equal(m.bert.rel_pos_bias.weight.data, torch.ones(2, 32))
    assert torch.equal(m.bert.embeddings.position_embeddings.weight.data[:12], old)
    assert "classifier.weight" in missing and "bert.encoder.layer.11.output.dense.weight" in unexpected


def test_checkpoint_and_teacher_table_roundtrip(tmp_path):
    import tinyrec.checkpoint as ck
    m = torch.nn.Linear(3, 2)
    p = tmp_path / "epoch-1.pt"
    ck.save_checkpoint(str(p), m, category_dict={"a": 1}, word_dict=None, subcategory_dict={"b": 2})
    raw = torch.load(str(p), map_location="cpu")
    assert sorted(raw) == ["category_dict", "model_state_dict", "subcategory_dict", "word_dict"]     # run.py:205-214
    m2 = torch.nn.Linear(3, 2)
    ck.load_checkpoint(str(p), m2)
    assert torch.equal(m2.weight, m.weight) and raw["category_dict"] == {"a": 1}
    tab = np.random.default_rng(0).standard_normal((7, 4)).astype(np.float32)
    tp = tmp_path / "teacher_emb-0.pkl"
    ck.save_teacher_table(str(tp), torch.from_numpy(tab))
    import pickle
    with open(tp, "rb") as f:
        assert np.array_equal(pickle.load(f), tab)                                                   # run.py:458-459
    assert np.array_equal(ck.load_teacher_table(str(tp)), tab)


def test_batch_sources_shard_evenly_and_reiterate():
    """run.IndexBatches / run.LineBatches (streaming.py:53-54 sharding): every rank gets the SAME number of batches
    whatever the remainder (a ragged tail would un-pair the per-step all-reduces), ranks are disjoint, a second
    pass (epoch) yields again, and parsed lines equal the oracle's batch assembly (dataloader.py:118-148)."""
    import random
    import tinyrec.run as trun
    import tinyrec.synth as synth
    from oracle import batching as ob
    H, K, bs = 6, 5, 4
    hist_idx, hmask, cand_idx, label = synth.train_impressions(4 * 11 + 3, 200, H, K, seed=3)      # 11 batches + 3 left over
    for world in (1, 2, 3, 4):
        seen, lens = [], []
        for rank in range(world):
            src = trun.IndexBatches(hist_idx, hmask, cand_idx, label, bs, rank=rank, world=world, device="cpu")
            first = [tuple(t.clone() for t in b) for b in src]
            second = [b for b in src]                                                             # re-iterable
            assert len(first) == len(second) == len(src) == 11 // world
            lens.append(len(first))
            for b, b2 in zip(first, second):
                assert all(torch.equal(x, y) for x, y in zip(b, b2))
                assert b[0].dtype == torch.int32 and b[1].dtype == torch.float32 and b[3].dtype == torch.int64
                seen.append(b[0].numpy().tobytes())
        assert len(set(lens)) == 1 and len(set(seen)) == len(seen)                                # equal counts, disjoint
    # LineBatches
    args = synth.demo_args(user_log_length=H, npratio=K - 1, batch_size=bs)
    news_index = {f"N{i}": i for i in range(1, 50)}
    rnd = random.Random(0)
    lines = []
    for i in range(2 * 2 * bs + 3):
        clicks = " ".join(f"N{rnd.randrange(1, 60)}" for _ in range(rnd.randrange(0, 10)))      # ids >= 50 are unknown -> 0
        neg = " ".join(f"N{rnd.randrange(1, 50)}" for _ in range(K - 1))
        lines.append(f"{i}\tU\tt\t{clicks}\tN{rnd.randrange(1, 50)}\t{neg}")
    for rank in range(2):
        src = trun.batches_from_lines(lines, news_index, args, rank=rank, world=2, device="cpu", seed=5)
        ep0 = [tuple(t.clone() for t in b) for b in src]
        ep1 = [tuple(t.clone() for t in b) for b in src]
        assert len(ep0) == len(ep1) == 2                       # 19 lines // (2 ranks x 4) = 2 batches per rank
        rng = random.Random(5)                                 # epoch 0 draws labels from Random(seed + 0)
        for b, (h, m, c, lab) in enumerate(ep0):
            for j in range(bs):
                cols = lines[(b * bs + j) * 2 + rank].split("\t")
                want_h, want_m = ob.pad_history(ob.to_index(cols[3].split(), news_index), H)
                y = rng.randint(0, K - 1)
                want_c = ob.insert_positive(ob.to_index(cols[4].split(), news_index), ob.to_index(cols[5].split(), news_index), y)
                assert h[j].tolist() == want_h and m[j].tolist() == [float(x) for x in want_m]
                assert c[j].tolist() == want_c and int(lab[j]) == y
        assert all(torch.equal(a[0], b_[0]) for a, b_ in zip(ep0, ep1))      # same lines ...
        assert any(not torch.equal(a[3], b_[3]) for a, b_ in zip(ep0, ep1))  # ... fresh positive slots per epoch


def test_news_encoder_model_name_must_load(tmp_path):
    """model_bert.py:114 ``from_pretrained(args.model_name)``: a set model_name is loaded (UniLM .bin conversion) or
    is an error -- never a silent random init."""
    import tinyrec.model_bert as mb
    import tinyrec.synth as synth
    from tinyrec._lib import TinyRecError
    with pytest.raises(TinyRecError):
        mb.NewsEncoder(synth.demo_args(num_student_layers=1, model_name=str(tmp_path / "missing.bin")))
    src = synth.bert_model_state("", 2, 8)
    fused = {}
    for k, v in src.items():                                   # UniLM layout: fused qkv, q_bias / v_bias, no key bias
        if k.endswith("attention.self.query.weight"):
            base = k[:-len("query.weight")]
            fused[base + "qkv_linear.weight"] = torch.cat([v, src[base + "key.weight"], src[base + "value.weight"]], 0)
            fused[base + "q_bias"] = src[base + "query.bias"]
            fused[base + "v_bias"] = src[base + "value.bias"]
        elif ".attention.self." in k or k.startswith(("bert.pooler", "classifier")):
            continue
        elif k == "bert.rel_pos_bias.weight":
            fused["bert.encoder.rel_pos_bias.weight"] = v
        else:
            fused[k] = v
    path = str(tmp_path / "unilm.bin")
    torch.save(fused, path)
    ne = mb.NewsEncoder(synth.demo_args(num_student_layers=1, model_name=path))      # 2-layer file into a 1-layer student
    got = ne.bert_model.state_dict()
    assert torch.equal(got["bert.encoder.layer.0.attention.self.value.weight"], src["bert.encoder.layer.0.attention.self.value.weight"])
    assert torch.equal(got["bert.embeddings.word_embeddings.weight"], src["bert.embeddings.word_embeddings.weight"])
    assert torch.equal(got["bert.rel_pos_bias.weight"], src["bert.rel_pos_bias.weight"])
    assert float(got["bert.encoder.layer.0.attention.self.key.bias"].abs().max()) == 0.0
    assert all(k.startswith(("bert.pooler", "classifier")) for k in ne.pretrained_missing_keys)


def test_sass_is_blackwell_native(built_lib):
    """Static check of the shipped library (no GPU): the GEMM / scoring kernels really are tcgen05 + TMEM + TMA
    (UTCHMMA, LDTM, UTMALDG / UTMASTG in their SASS; the CTA-pair variants issue UTCHMMA.2CTA), the MMA issue loops keep
    their operands in uniform registers (no R2UR.BROADCAST waterfall around the UTCHMMAs -- that was 75 cycles per MMA),
    the NVLS exchange kernel reduces in the switch (LDGMC...ADD = multimem.ld_reduce) and the attention backward stages
    its fragments with stmatrix (STSM)."""
    import shutil
    import subprocess
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    sass = subprocess.run([cuobjdump, "-sass", built_lib], capture_output=True, text=True, check=True).stdout
    funcs = {}
    for blk in re.split(r"\n\s*Function : ", sass)[1:]:
        name, _, body = blk.partition("\n")
        funcs[name.strip()] = body
    gemms = {n: b for n, b in funcs.items() if "gemm_kernel" in n and "sgemm" not in n}
    assert len(gemms) >= 40, len(gemms)                                  # every (BN, layout, epilogue, CTA-pair) variant
    for n, b in gemms.items():
        assert b.count("UTCHMMA") == 4 and "LDTM" in b and "UTMALDG" in b, n
        assert "R2UR.BROADCAST" not in b, n
        cta2 = n.split("gemm_kernelI")[1].split("EEEv")[0].endswith("Lb1")       # last template argument: CTA2
        assert ("UTCHMMA.2CTA" in b) == cta2, n
        epi_tma = "ELb1ELb" in n.split("gemm_kernelI")[1][-12:]
        if epi_tma:
            assert "UTMASTG" in b, n
    logits = [b for n, b in funcs.items() if "ue_logits_kernel" in n]
    assert len(logits) == 2 and all("UTCHMMA.2CTA" in b and "LDTM" in b and "R2UR.BROADCAST" not in b for b in logits)
    nvls = [b for n, b in funcs.items() if "allreduce_nvls_kernel" in n]
    assert len(nvls) == 1 and "LDGMC" in nvls[0]
    attn_bwd = [b for n, b in funcs.items() if "tnr15attn_bwd_kernel" in n]       # (not nrms_attn_bwd_kernel)
    assert len(attn_bwd) == 1 and "STSM" in attn_bwd[0] and "HMMA" in attn_bwd[0]
