"""GPU parity tests of the module-level API (drop-in boundary) against
(a) the golden vectors produced by the real reference code (tests/golden/*.npz) and
(b) the CPU oracle on seeded synthetic inputs.

Tolerances: bf16 compute path -> 1e-2 relative (BASELINE.json north_star) on logits /
embeddings / losses (SCORE_TOL); gradients per tensor relative in norm (GRAD_TOL: bf16 activations and
activation gradients; every measured value is appended to gpurun_out/test_report.jsonl and the bounds are
kept at about twice the worst one); bit-exact for gathers; 1e-3 absolute for ranking metrics (they are
exact in practice)."""
import math

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


# Per-tensor relative gradient error, bf16 activations / activation gradients with fp32 accumulation.  Bounds = about
# twice the worst value measured on B200 (round 2, profiles/r02_test_report_summary.txt): 1.4e-2 over the 51 tensors of
# the benchmark configuration, 2.1e-2 with the NRMS user encoder, 2.9e-2 at 100-token rows, 6.8e-2 behind 12 layers.
GRAD_TOL = 3e-2
GRAD_TOL_LONG = 6e-2      # 100-token rows: streamed-KV attention backward, three more bf16 round trips per layer
GRAD_TOL_12L = 1e-1       # 12 bf16 layers in front of the loss; the scores are ~60 with gaps of ~1
SCORE_TOL = 1e-2          # north_star: logits / embeddings within 1e-2 relative on the bf16 path


def _chk(label, value, bound):
    """assert value < bound, recording the measured value (tests/conftest.py: gpurun_out/test_report.jsonl) so the
    bounds in this file can be kept at ~2x what the kernels actually deliver."""
    from conftest import report
    report(label, float(value), float(bound))
    assert value < bound, (label, value, bound)


def _rel(got, ref):
    got = torch.as_tensor(got).float().cpu()
    ref = torch.as_tensor(ref).float().cpu()
    return float((got - ref).norm() / (ref.norm() + 1e-12))


def _grad_err(name, got, ref, named):
    """Relative gradient error.  d loss / d key.bias is identically zero (softmax is invariant to a
    per-query constant q.b_k), so both sides hold rounding noise: measure it against the query-bias
    gradient scale instead of against itself."""
    got = torch.as_tensor(got).float().cpu()
    ref = torch.as_tensor(ref).float().cpu()
    if name.endswith("attention.self.key.bias"):
        scale = named[name.replace("key.bias", "query.bias")].grad.float().norm().cpu()
        return float((got - ref).norm() / (scale + 1e-12))
    if name.endswith("multi_head_self_attn.W_K.bias"):
        # NRMS: exp(q.(k + b_K)) / sum is invariant to b_K as well (up to the 1e-8 in the denominator)
        scale = named[name.replace("W_K.bias", "W_Q.bias")].grad.float().norm().cpu()
        return float((got - ref).norm() / (scale + 1e-12))
    if ".user_encoder.attn." in name and name.replace("attn." + name.split("attn.")[1], "multi_head_self_attn.W_V.weight") in named:
        # NRMS: the self-attention averages the history, so the rows the pooling sees are nearly equal and the true
        # gradient of the pooling parameters is ~0 (1e-10 against 1e-3 for the W_V next to it): both sides hold the
        # bf16 noise of the news vectors.  Measure against the W_V gradient of the same encoder.
        wv = named[name.replace("attn." + name.split("attn.")[1], "multi_head_self_attn.W_V.weight")].grad
        return float((got - ref).norm() / (max(float(ref.norm()), float(wv.float().norm().cpu())) + 1e-12))
    if name.endswith("attn.att_fc2.bias"):
        # additive-attention weights are shift invariant too (up to the 1e-8 in the denominator)
        scale = named[name.replace("att_fc2.bias", "att_fc2.weight")].grad.float().norm().cpu()
        return float((got - ref).norm() / (scale + 1e-12))
    return float((got - ref).norm() / (ref.norm() + 1e-12))


def _apply_freeze(model, trainable_layers, student=True):
    """run.py:101-112"""
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
            for p in layer.parameters():
                p.requires_grad = True


def test_news_encoder_vs_reference_golden(golden):
    import tinyrec.model_bert as mb
    import tinyrec.synth as synth
    g = golden("encoder")
    layers = int(g["layers"])
    sd = synth.model_bert_state("", layers, int(g["seed"]), noisy=True)
    sd = {k[len("news_encoder."):]: v for k, v in sd.items() if k.startswith("news_encoder.")}
    ne = mb.NewsEncoder(synth.demo_args(num_student_layers=layers))
    ne.load_state_dict(sd, strict=True)
    ne.cuda().eval()
    with torch.no_grad():
        vec = ne(torch.from_numpy(g["x"]).cuda())
    assert _rel(vec, g["news_vec"]) < 1e-2
    assert torch.isfinite(vec).all()


@pytest.mark.parametrize("tag,ulm", [("pad", False), ("mask", True)])
def test_model_bert_vs_reference_golden(golden, tag, ulm):
    import tinyrec.model_bert as mb
    import tinyrec.synth as synth
    g = golden("modelbert")
    layers = int(g["layers"])
    H = g["history"].shape[1]
    m = mb.ModelBert(synth.demo_args(num_student_layers=layers, user_log_length=H, user_log_mask=ulm))
    m.load_state_dict(synth.model_bert_state("", layers, int(g["seed"]), noisy=True), strict=True)
    m.cuda().eval()
    with torch.no_grad():
        score, hv, cv, uv = m(torch.from_numpy(g["history"]).cuda(), torch.from_numpy(g["history_mask"]).cuda(),
                              torch.from_numpy(g["candidate"]).cuda())
    assert _rel(hv, g[f"hist_{tag}"]) < 1e-2
    assert _rel(cv, g[f"cand_{tag}"]) < 1e-2
    assert _rel(uv, g[f"user_{tag}"]) < 1e-2
    _chk("test_model_bert_vs_reference_golden.score", _rel(score, g[f"score_{tag}"]), SCORE_TOL)
    if ulm:
        assert float(uv[1].abs().max()) == 0.0          # fully masked history -> exact zeros


def test_plmnr_loss_vs_reference_golden(golden):
    import tinyrec.model_bert_2 as mb2
    import tinyrec.synth as synth
    g = golden("modelbert")
    layers = int(g["layers"])
    H = g["history"].shape[1]
    m = mb2.ModelBert(synth.demo_args(num_hidden_layers=layers, user_log_length=H, user_log_mask=False))
    m.load_state_dict(synth.model_bert_state("", layers, int(g["seed"]), noisy=True), strict=True)
    m.cuda().eval()
    _apply_freeze(m, [0], student=False)
    with torch.no_grad():
        loss, score = m(torch.from_numpy(g["history"]).cuda(), torch.from_numpy(g["history_mask"]).cuda(),
                        torch.from_numpy(g["candidate"]).cuda(), torch.from_numpy(g["plmnr_label"]).cuda())
    assert abs(float(loss) - float(g["plmnr_loss"])) < 1e-2 * abs(float(g["plmnr_loss"])) + 1e-3
    _chk("test_plmnr_loss_vs_reference_golden.score", _rel(score, g["plmnr_score"]), SCORE_TOL)


def _kd_model(g, ulm):
    import tinyrec.model_bert as mb
    import tinyrec.synth as synth
    layers, M = int(g["layers"]), int(g["M"])
    H = g["history"].shape[1]
    m = mb.Model(synth.demo_args(num_student_layers=layers, user_log_length=H, user_log_mask=ulm, num_teachers=M,
                                 temperature=float(g["temperature"]), coef=float(g["coef"])))
    m.load_state_dict(synth.kd_model_state(layers, M, int(g["seed"]), noisy=True), strict=True)
    m.cuda().eval()                      # golden vectors are eval()-mode outputs of the reference
    _apply_freeze(m, [int(i) for i in g["trainable"]])
    inputs = (torch.from_numpy(g["history"]).cuda(), torch.from_numpy(g["history_mask"]).cuda(),
              torch.from_numpy(g["candidate"]).cuda(), torch.from_numpy(g["label"]).cuda(),
              [torch.from_numpy(g[f"th{i}"]).cuda() for i in range(M)], [torch.from_numpy(g[f"tc{i}"]).cuda() for i in range(M)])
    return m, inputs


@pytest.mark.parametrize("tag,ulm", [("pad", False), ("mask", True)])
def test_kd_forward_vs_reference_golden(golden, tag, ulm):
    g = golden("kd")
    m, inputs = _kd_model(g, ulm)
    with torch.no_grad():
        res = m(*inputs)
    for v, nm in zip(res[:4], ("total", "distill", "emb", "target")):
        ref = float(g[f"{nm}_{tag}"])
        assert abs(float(v) - ref) < 1e-2 * abs(ref) + 1e-4, (nm, float(v), ref)
    _chk("test_kd_forward_vs_reference_golden.score", _rel(res[4], g[f"score_{tag}"]), SCORE_TOL)


def test_kd_gradients_vs_reference_golden(golden):
    g = golden("kd")
    m, inputs = _kd_model(g, False)
    total = m(*inputs)[0]
    total.backward()
    named = dict(m.named_parameters())
    worst = 0.0
    for k in [str(s) for s in g["trainable_names"]]:
        gr = named[k].grad
        assert gr is not None, k
        if f"gfull/{k}" in g.files:
            ref = torch.from_numpy(g[f"gfull/{k}"])
            r = _grad_err(k, gr.reshape(ref.shape), ref, named)
        else:
            ref = torch.from_numpy(g[f"gslice/{k}"])
            r = _grad_err(k, gr[:16, :16], ref, named)
        worst = max(worst, r)
        _chk(f"test_kd_gradients_vs_reference_golden.grad.{k}", r, GRAD_TOL)
        ref_abs = float(g[f"gabs/{k}"])
        if not k.endswith(("key.bias", "att_fc2.bias")):
            assert abs(float(gr.double().abs().sum()) - ref_abs) < 3e-2 * ref_abs + 1e-7, k
    # frozen parameters got no gradient
    for k, p in named.items():
        if not p.requires_grad:
            assert p.grad is None
    # a second backward accumulates (optimizer.zero_grad() semantics are the caller's)
    g1 = named["transform_matrix.0.weight"].grad.clone()
    m(*inputs)[0].backward()
    assert _rel(named["transform_matrix.0.weight"].grad, 2 * g1) < 1e-3


# ------------------------------------------------------------------ NRMS user encoder
def _nrms_model(g, ulm):
    import tinyrec.model_bert as mb
    import tinyrec.synth as synth
    layers, M = int(g["layers"]), int(g["M"])
    H = g["history"].shape[1]
    m = mb.Model(synth.demo_args(num_student_layers=layers, user_log_length=H, user_log_mask=ulm, num_teachers=M,
                                 temperature=float(g["temperature"]), coef=float(g["coef"]), model="NRMS",
                                 num_attention_heads=16))
    m.load_state_dict(synth.kd_model_state(layers, M, int(g["seed"]), noisy=True, model="NRMS", n_heads=16), strict=True)
    m.cuda().eval()
    _apply_freeze(m, [0])
    inputs = (torch.from_numpy(g["history"]).cuda(), torch.from_numpy(g["history_mask"]).cuda(),
              torch.from_numpy(g["candidate"]).cuda(), torch.from_numpy(g["label"]).cuda(),
              [torch.from_numpy(g[f"th{i}"]).cuda() for i in range(M)], [torch.from_numpy(g[f"tc{i}"]).cuda() for i in range(M)])
    return m, inputs


@pytest.mark.parametrize("tag,ulm", [("pad", False), ("mask", True)])
def test_nrms_kd_vs_reference_golden(golden, tag, ulm):
    """KD ``Model`` with args.model = 'NRMS' against the reference's own outputs and gradients
    (tests/golden/nrms.npz, make_golden.py:gen_nrms): losses, scores, every trainable tensor's gradient."""
    g = golden("nrms")
    m, inputs = _nrms_model(g, ulm)
    res = m(*inputs)
    for v, nm in zip(res[:4], ("total", "distill", "emb", "target")):
        ref = float(g[f"{nm}_{tag}"])
        assert abs(float(v) - ref) < 1e-2 * abs(ref) + 1e-4, (nm, float(v), ref)
    _chk("test_nrms_kd_vs_reference_golden.score", _rel(res[4], g[f"score_{tag}"]), SCORE_TOL)
    res[0].backward()
    named = dict(m.named_parameters())
    for k in [str(s) for s in g[f"trainable_names_{tag}"]]:
        gr = named[k].grad
        assert gr is not None, k
        if f"{tag}/gfull/{k}" in g.files:
            ref = torch.from_numpy(g[f"{tag}/gfull/{k}"])
            r = _grad_err(k, gr.reshape(ref.shape), ref, named)
        else:
            ref = torch.from_numpy(g[f"{tag}/gslice/{k}"])
            r = _grad_err(k, gr[:16, :16], ref, named)
        _chk(f"test_nrms_kd_vs_reference_golden.grad.{k}", r, GRAD_TOL)
    # the stand-alone user encoder (run.py:343) gives the same user vector as the training path
    st = m.train_state()
    hist = st.last["news"][:inputs[0].shape[0] * inputs[0].shape[1]].view(inputs[0].shape[0], inputs[0].shape[1], -1)
    with torch.no_grad():
        user = m.student.user_encoder(hist, inputs[1])
    assert _rel(user, st.last["user"]) < 1e-5


def test_nrms_step_vs_oracle_at_demo_shape():
    """B=8 impressions at the demo history length (H=50, 16 heads) against the CPU oracle (itself pinned on the
    reference by tests/test_oracle_golden.py::test_nrms_kd_forward_and_gradients)."""
    import tinyrec.model_bert as mb
    import tinyrec.synth as synth
    from oracle import model as om
    B, H, K, L, M, layers, N = 8, 50, 5, 12, 2, 1, 500
    news = synth.news_table(N, L=L, seed=1)
    hist_idx, hmask, cand_idx, label = synth.train_impressions(B, N, H, K, seed=2)
    tables = synth.teacher_tables(N, M, 256, seed=3)
    history = torch.from_numpy(news[hist_idx].astype(np.int64))
    candidate = torch.from_numpy(news[cand_idx].astype(np.int64))
    th = [torch.from_numpy(t[hist_idx]) for t in tables]
    tc = [torch.from_numpy(t[cand_idx]) for t in tables]
    sd = synth.kd_model_state(layers, M, 5, noisy=True, model="NRMS", n_heads=16)
    for ulm in (False, True):
        args = synth.demo_args(num_student_layers=layers, num_teachers=M, user_log_length=H, model="NRMS", user_log_mask=ulm)
        m = mb.Model(args)
        m.load_state_dict(sd, strict=True)
        m.cuda().eval()
        _apply_freeze(m, [0])
        res = m(history.cuda(), torch.from_numpy(hmask).cuda(), candidate.cuda(), torch.from_numpy(label).cuda(),
                [t.cuda() for t in th], [t.cuda() for t in tc])
        res[0].backward()
        osd = {k: v.clone() for k, v in sd.items()}
        names = [k for k, p in m.named_parameters() if p.requires_grad]
        for k in names:
            osd[k].requires_grad_(True)
        ref = om.kd_model_forward(osd, history, torch.from_numpy(hmask), candidate, torch.from_numpy(label), th, tc,
                                  layers, ulm, args.temperature, args.coef)
        ref[0].backward()
        assert abs(float(res[0]) - float(ref[0])) < 1e-2 * abs(float(ref[0])) + 1e-4
        _chk("test_nrms_step_vs_oracle_at_demo_shape.score", _rel(res[4], ref[4].detach()), SCORE_TOL)
        named = dict(m.named_parameters())
        for k in names:
            if osd[k].grad is None:                 # pad_doc in the user_log_mask branch
                assert float(named[k].grad.abs().max()) == 0.0, k
                continue
            _chk(f"test_nrms_step_vs_oracle_at_demo_shape.grad.{k}", _grad_err(k, named[k].grad, osd[k].grad, named), GRAD_TOL)


def test_kd_step_vs_oracle_at_demo_shape():
    """B=4 impressions at the demo shape (H=50, K=5, L=30, M=4, 2 layers, layer 1 trainable):
    CUDA path vs the CPU oracle on the same seeded inputs; losses, scores and gradients."""
    import tinyrec.model_bert as mb
    import tinyrec.synth as synth
    from oracle import model as om
    B, H, K, L, M, layers, D = 4, 50, 5, 30, 4, 2, 256
    news = synth.news_table(500, L=L, seed=3)
    hist_idx, hmask, cand_idx, label = synth.train_impressions(B, 500, H, K, seed=4)
    tables = synth.teacher_tables(500, M, D, seed=5)
    history = torch.from_numpy(news[hist_idx].astype(np.int64))
    candidate = torch.from_numpy(news[cand_idx].astype(np.int64))
    th = [torch.from_numpy(t[hist_idx]) for t in tables]
    tc = [torch.from_numpy(t[cand_idx]) for t in tables]
    sd = synth.kd_model_state(layers, M, 11, noisy=True)
    args = synth.demo_args(num_student_layers=layers, num_teachers=M)
    m = mb.Model(args)
    m.load_state_dict(sd, strict=True)
    m.cuda().eval()
    _apply_freeze(m, [1])
    res = m(history.cuda(), torch.from_numpy(hmask).cuda(), candidate.cuda(), torch.from_numpy(label).cuda(),
            [t.cuda() for t in th], [t.cuda() for t in tc])
    res[0].backward()
    keys = om.trainable_keys(sd, [1])
    osd = {k: v.clone().requires_grad_(k in keys) for k, v in sd.items()}
    ref = om.kd_model_forward(osd, history, torch.from_numpy(hmask), candidate, torch.from_numpy(label), th, tc, layers,
                              False, args.temperature, args.coef)
    ref[0].backward()
    for v, r, nm in zip(res[:4], ref[:4], ("total", "distill", "emb", "target")):
        assert abs(float(v) - float(r)) < 1e-2 * abs(float(r)) + 1e-4, (nm, float(v), float(r))
    _chk("test_kd_step_vs_oracle_at_demo_shape.score", _rel(res[4], ref[4].detach()), SCORE_TOL)
    named = dict(m.named_parameters())
    for k in keys:
        r = _grad_err(k, named[k].grad, osd[k].grad, named)
        _chk(f"test_kd_step_vs_oracle_at_demo_shape.grad.{k}", r, GRAD_TOL)


def test_kd_training_mode_step_with_dropout_vs_oracle_with_injected_masks():
    """Training mode (the reference's train() never calls eval(): dropout 0.1 is active in the
    embeddings, attention probabilities and both dense outputs of every layer).  The CUDA path draws
    its masks from a counter-based generator; the oracle gets the SAME masks injected
    (oracle/dropout.py), so losses, scores and all gradients must agree like in eval mode."""
    import tinyrec.model_bert as mb
    import tinyrec.synth as synth
    from oracle import model as om
    from oracle.dropout import Plan
    B, H, K, L, M, layers, D = 3, 7, 3, 12, 2, 2, 256
    news = synth.news_table(300, L=L, seed=3)
    hist_idx, hmask, cand_idx, label = synth.train_impressions(B, 300, H, K, seed=4)
    tables = synth.teacher_tables(300, M, D, seed=5)
    history = torch.from_numpy(news[hist_idx].astype(np.int64))
    candidate = torch.from_numpy(news[cand_idx].astype(np.int64))
    th = [torch.from_numpy(t[hist_idx]) for t in tables]
    tc = [torch.from_numpy(t[cand_idx]) for t in tables]
    sd = synth.kd_model_state(layers, M, 11, noisy=True)
    args = synth.demo_args(num_student_layers=layers, num_teachers=M, user_log_length=H)
    m = mb.Model(args)
    m.load_state_dict(sd, strict=True)
    m.cuda()
    assert m.training                                  # default nn.Module mode, as in the reference's train()
    _apply_freeze(m, [0, 1])
    drop = m.student.news_encoder.drop_state(torch.device("cuda", 0))
    seed0 = int(drop.seed.item())
    gpu_in = (history.cuda(), torch.from_numpy(hmask).cuda(), candidate.cuda(), torch.from_numpy(label).cuda(),
              [t.cuda() for t in th], [t.cuda() for t in tc])
    res = m(*gpu_in)
    res[0].backward()
    assert int(drop.seed.item()) == seed0 + 1          # one seed per step
    keys = om.trainable_keys(sd, [0, 1])
    osd = {k: v.clone().requires_grad_(k in keys) for k, v in sd.items()}
    plan = Plan(seed0 + 1, drop.p_hidden, drop.p_attn)
    ref = om.kd_model_forward(osd, history, torch.from_numpy(hmask), candidate, torch.from_numpy(label), th, tc, layers,
                              False, args.temperature, args.coef, drop_plan=plan)
    ref[0].backward()
    for v, r, nm in zip(res[:4], ref[:4], ("total", "distill", "emb", "target")):
        assert abs(float(v) - float(r)) < 1e-2 * abs(float(r)) + 1e-4, (nm, float(v), float(r))
    _chk("test_kd_training_mode_step_with_dropout_vs_oracle_with_injected_masks.score", _rel(res[4], ref[4].detach()), SCORE_TOL)
    named = dict(m.named_parameters())
    for k in keys:
        r = _grad_err(k, named[k].grad, osd[k].grad, named)
        _chk(f"test_kd_training_mode_step_with_dropout_vs_oracle_with_injected_masks.grad.{k}", r, GRAD_TOL)
    # eval() switches dropout off and a new step draws new masks
    with torch.no_grad():
        a1 = float(m(*gpu_in)[0])
        a2 = float(m(*gpu_in)[0])
        m.eval()
        e1 = float(m(*gpu_in)[0])
        e2 = float(m(*gpu_in)[0])
    assert a1 != a2 and e1 == e2


def test_train_state_survives_optimizer_step_and_matches_oracle_adam(golden):
    """Fused Adam(amsgrad) kernel vs torch.optim.Adam golden trajectory."""
    import tinyrec.ops as ops
    g = golden("adam")
    p = torch.from_numpy(g["p0"].copy()).cuda()
    m, v, vmax = torch.zeros_like(p), torch.zeros_like(p), torch.zeros_like(p)
    shadow = torch.zeros(p.numel(), device="cuda", dtype=torch.bfloat16)
    for step in range(g["grads"].shape[0]):
        ops.adam_amsgrad(p, torch.from_numpy(g["grads"][step]).cuda(), m, v, vmax, shadow, 1e-4, 0.9, 0.999, 1e-8, step + 1)
        np.testing.assert_allclose(p.cpu().numpy(), g["params"][step], rtol=2e-6, atol=1e-7)
    assert torch.equal(shadow, p.to(torch.bfloat16))


def test_fused_adam_moves_model_and_refreshes_shadow(golden):
    import tinyrec.optim as topt
    g = golden("kd")
    m, inputs = _kd_model(g, False)
    opt = topt.Adam(m, lr=1e-3)
    l0 = float(m(*inputs)[0])
    for _ in range(5):
        opt.zero_grad()
        loss = m(*inputs)[0]
        loss.backward()
        opt.step()
    l1 = float(m(*inputs)[0])
    assert l1 < l0                                # same batch, loss must go down
    st = m.train_state()
    assert torch.equal(st.flat.shadow, st.flat.data.to(torch.bfloat16))


def test_gathers_bit_exact(golden):
    import tinyrec.ops as ops
    g = golden("batching")
    table = torch.from_numpy(g["news_combined"]).cuda()
    H = int(g["H"])
    uf = g["user_feature"]              # int64 [B, H, 2L] from the reference loader
    # recover the indices the reference used by matching rows is unnecessary: re-derive them like the oracle
    from oracle import batching as ob
    n_news = table.shape[0] - 1
    news_index = {f"N{i}": i for i in range(1, n_news + 1)}
    idx = [ob.pad_history(ob.to_index(str(c).split(), news_index), H)[0] for c in g["clicks"]]
    idx_t = torch.tensor(idx, dtype=torch.int32).cuda().reshape(-1)
    out = torch.empty(idx_t.numel(), table.shape[1], dtype=torch.int64, device="cuda")
    ops.gather_rows_i32_i64(table, idx_t, out)
    assert np.array_equal(out.cpu().numpy().reshape(uf.shape), uf)
    t0 = torch.from_numpy(g["t0"]).cuda()
    outf = torch.empty(idx_t.numel(), t0.shape[1], device="cuda")
    ops.gather_rows_f32(t0, idx_t, outf)
    assert np.array_equal(outf.cpu().numpy().reshape(g["th0"].shape), g["th0"])
    # out-of-range index -> row 0
    bad = torch.tensor([-1, 10 ** 6], dtype=torch.int32).cuda()
    o2 = torch.empty(2, t0.shape[1], device="cuda")
    ops.gather_rows_f32(t0, bad, o2)
    assert torch.equal(o2[0], t0[0]) and torch.equal(o2[1], t0[0])


def test_eval_metrics_vs_reference_golden(golden):
    """Scores are injected through a 1-D 'embedding' (D=32: table row = [score,0,...,0], user = e0)."""
    import tinyrec.ops as ops
    g = golden("metrics")
    ptr = torch.from_numpy(g["ptr"].astype(np.int64)).cuda()
    nnz = int(g["ptr"][-1])
    table = torch.zeros(nnz, 32, device="cuda")
    table[:, 0] = torch.from_numpy(g["score"]).cuda()
    n_imp = len(g["ptr"]) - 1
    user = torch.zeros(n_imp, 32, device="cuda")
    user[:, 0] = 1.0
    cand = torch.arange(nnz, dtype=torch.int32, device="cuda")
    label = torch.from_numpy(g["label"].astype(np.int8)).cuda()
    per = torch.zeros(n_imp, 5, dtype=torch.float64, device="cuda")
    sums = torch.zeros(5, dtype=torch.float64, device="cuda")
    ops.eval_metrics(table, user, ptr, cand, label, int(np.diff(g["ptr"]).max()), per, sums)
    per = per.cpu().numpy()
    vals = g["vals"]
    valid = ~np.isnan(vals[:, 0])
    assert np.array_equal(per[:, 4] == 1.0, valid)
    # AUC is tie-aware and exact; MRR / nDCG depend on the tie order of np.argsort (unspecified for
    # equal scores), so impressions whose scores contain ties among differently-labelled items get 1e-3.. skip
    np.testing.assert_allclose(per[valid, 0], vals[valid, 0], atol=1e-12)
    ptr_h = g["ptr"]
    for i in np.nonzero(valid)[0]:
        s = g["score"][ptr_h[i]:ptr_h[i + 1]]
        if len(np.unique(s)) == len(s):
            np.testing.assert_allclose(per[i, 1:4], vals[i, 1:4], atol=1e-12)
    np.testing.assert_allclose(sums.cpu().numpy()[:4], per[:, :4].sum(0), rtol=1e-12)
    assert int(sums[4]) == int(valid.sum())


def test_graphed_train_step_matches_eager_loop():
    """tinyrec.run.GraphedTrainStep (one CUDA graph per step) against the eager loop of run.py:178-197 on the same
    batches: same losses step by step, same parameters afterwards (split-K wgrads use fp32 atomics: not bitwise),
    the Adam step counter advances inside the graph, and constructing the object does not train the model."""
    import tinyrec.model_bert as mb
    import tinyrec.optim as topt
    import tinyrec.run as trun
    import tinyrec.synth as synth
    B, H, K, L, M, layers, N = 4, 10, 3, 12, 2, 2, 300
    news = synth.news_table(N, L=L, seed=1)
    tables = synth.teacher_tables(N, M, 256, seed=3)
    batches = []
    for s in range(3):
        hist_idx, hmask, cand_idx, label = synth.train_impressions(B, N, H, K, seed=10 + s)
        batches.append((torch.from_numpy(news[hist_idx].astype(np.int64)).cuda(), torch.from_numpy(hmask).cuda(),
                        torch.from_numpy(news[cand_idx].astype(np.int64)).cuda(), torch.from_numpy(label).cuda(),
                        [torch.from_numpy(t[hist_idx]).cuda() for t in tables], [torch.from_numpy(t[cand_idx]).cuda() for t in tables]))
    sd = synth.kd_model_state(layers, M, 5, noisy=True)
    runs = []
    for graphed in (False, True):
        m = mb.Model(synth.demo_args(num_student_layers=layers, num_teachers=M, user_log_length=H))
        m.load_state_dict(sd, strict=True)
        m.cuda().eval()                                   # dropout off: the two runs must see the same arithmetic
        _apply_freeze(m, [1])
        opt = topt.Adam(m, lr=1e-4)
        losses = []
        if graphed:
            before = m.student.news_encoder.dense.weight.detach().clone()
            step = trun.GraphedTrainStep(m, opt, batches[0])
            assert torch.equal(m.student.news_encoder.dense.weight.detach(), before)      # no side effect
            assert opt.steps_done() == 0
            for b in batches:
                losses.append(float(step(*b)[0]))
            assert opt.steps_done() == len(batches)
        else:
            for b in batches:
                opt.zero_grad()
                out = m(*b)
                out[0].backward()
                opt.step()
                losses.append(float(out[0]))
        runs.append((losses, {k: v.detach().clone() for k, v in m.named_parameters() if v.requires_grad}))
    (l0, p0), (l1, p1) = runs
    assert l0[0] != l0[1]
    for a, b in zip(l0, l1):                             # later steps inherit the +-lr noise steps of the zero-gradient biases
        assert abs(a - b) < 5e-3 * abs(a) + 1e-5, (l0, l1)
    for k in p0:
        if k.endswith(("key.bias", "att_fc2.bias")):     # gradient identically 0 up to rounding noise: Adam turns the
            continue                                     # noise into +-lr steps, different in any two runs
        assert _rel(p1[k], p0[k]) < 2e-3, k


def test_kd_step_long_sequence_vs_oracle():
    """KD train step on long news rows (100 tokens: body-length text): forward through the
    streamed-KV attention, backward through its dQ / dK-dV kernels; losses, scores and every trainable gradient
    against the CPU oracle."""
    import tinyrec.model_bert as mb
    import tinyrec.synth as synth
    from oracle import model as om
    B, H, K, L, M, layers, N = 2, 4, 3, 100, 2, 2, 120
    news = synth.news_table(N, L=L, seed=1)
    hist_idx, hmask, cand_idx, label = synth.train_impressions(B, N, H, K, seed=2)
    tables = synth.teacher_tables(N, M, 256, seed=3)
    history = torch.from_numpy(news[hist_idx].astype(np.int64))
    candidate = torch.from_numpy(news[cand_idx].astype(np.int64))
    th = [torch.from_numpy(t[hist_idx]) for t in tables]
    tc = [torch.from_numpy(t[cand_idx]) for t in tables]
    sd = synth.kd_model_state(layers, M, 5, noisy=True)
    args = synth.demo_args(num_student_layers=layers, num_teachers=M, user_log_length=H)
    m = mb.Model(args)
    m.load_state_dict(sd, strict=True)
    m.cuda().eval()
    _apply_freeze(m, [0, 1])
    res = m(history.cuda(), torch.from_numpy(hmask).cuda(), candidate.cuda(), torch.from_numpy(label).cuda(),
            [t.cuda() for t in th], [t.cuda() for t in tc])
    res[0].backward()
    osd = {k: v.clone() for k, v in sd.items()}
    names = [k for k, p in m.named_parameters() if p.requires_grad]
    for k in names:
        osd[k].requires_grad_(True)
    ref = om.kd_model_forward(osd, history, torch.from_numpy(hmask), candidate, torch.from_numpy(label), th, tc,
                              layers, False, args.temperature, args.coef)
    ref[0].backward()
    assert abs(float(res[0]) - float(ref[0])) < 1e-2 * abs(float(ref[0])) + 1e-4
    _chk("test_kd_step_long_sequence_vs_oracle.score", _rel(res[4], ref[4].detach()), SCORE_TOL)
    named = dict(m.named_parameters())
    for k in names:
        _chk(f"test_kd_step_long_sequence_vs_oracle.grad.{k}", _grad_err(k, named[k].grad, osd[k].grad, named), GRAD_TOL_LONG)


def test_doc_sim_matches_reference_loop():
    """run.py:292-299 (mean cosine of random row pairs) on the device against the oracle's verbatim loop, same
    ``random`` stream; i == j pairs are skipped but counted in the mean."""
    import random
    import tinyrec.run as trun
    from oracle import metrics as omet
    g = torch.Generator().manual_seed(3)
    table = torch.randn(40, 256, generator=g) * 0.3 + 0.05
    got = trun.doc_sim(table.cuda(), n_pairs=5000, rng=random.Random(11))
    want = omet.doc_sim(table.numpy(), 5000, random.Random(11))
    assert abs(got - want) < 1e-6, (got, want)


def test_post_train_distill_model_vs_oracle():
    """First-stage KD wrappers (Post-train_KD.ipynb cells 12, 14): 1+K titles of 12 tokens and a 72-token body per
    sample through the same encoder (two saved passes, the body through the streamed-KV attention), losses, scores and
    every trainable gradient against the oracle restatement; then the two learning-rate groups of cell 18."""
    from types import SimpleNamespace
    import tinyrec.optim as topt
    import tinyrec.post_train as pt
    import tinyrec.synth as synth
    from oracle import model as om
    B, K1, Lt, Lb, M, layers, D = 3, 4, 12, 72, 2, 2, 256
    rng = np.random.default_rng(5)
    title = torch.from_numpy(synth.news_table(B * K1 - 1, L=Lt, seed=1).astype(np.int64)).reshape(B, K1, 2 * Lt)
    body = torch.from_numpy(synth.news_table(B - 1, L=Lb, seed=2).astype(np.int64))
    labels = torch.tensor([0, 2, 1])
    tt = [torch.from_numpy(rng.standard_normal((B, K1, D)).astype(np.float32) * 0.3) for _ in range(M)]
    tb = [torch.from_numpy(rng.standard_normal((B, D)).astype(np.float32) * 0.3) for _ in range(M)]
    args = SimpleNamespace(news_query_vector_dim=200, news_dim=D, num_hidden_layers=layers, num_teachers=M)
    m = pt.DistillModel(args)
    full = synth.kd_model_state(layers, M, 7, noisy=True)
    sd = {k: v for k, v in full.items() if k.startswith(("student.news_encoder.", "transform_matrix."))}
    assert list(m.state_dict().keys()) == list(sd.keys())
    m.load_state_dict(sd, strict=True)
    m.cuda().eval()
    for p in m.student.news_encoder.bert_model.parameters():                # cell 17
        p.requires_grad = False
    for p in m.student.news_encoder.bert_model.bert.encoder.layer[1].parameters():
        p.requires_grad = True
    res = m(title.cuda(), body.cuda(), labels.cuda(), [t.cuda() for t in tt], [t.cuda() for t in tb])
    res[0].backward()
    osd = {k: v.clone() for k, v in sd.items()}
    names = [k for k, p in m.named_parameters() if p.requires_grad]
    for k in names:
        osd[k].requires_grad_(True)
    ref = om.post_train_distill_forward(osd, title, body, labels, tt, tb, layers)
    ref[0].backward()
    for got, want in zip(res[:4], ref[:4]):
        assert abs(float(got) - float(want)) < 1e-2 * abs(float(want)) + 1e-4, (float(got), float(want))
    _chk("test_post_train_distill_model_vs_oracle.score", _rel(res[4], ref[4].detach()), SCORE_TOL)
    named = dict(m.named_parameters())
    for k in names:
        _chk(f"test_post_train_distill_model_vs_oracle.grad.{k}", _grad_err(k, named[k].grad, osd[k].grad, named), GRAD_TOL)
    # inference form of cell 12
    with torch.no_grad():
        sc, te, be = m.student(title.cuda(), body.cuda())
    _chk("test_post_train_distill_model_vs_oracle.score_inference", _rel(sc, ref[4].detach()), SCORE_TOL)
    assert tuple(te.shape) == (B, K1, D) and tuple(be.shape) == (B, D)
    # cell 18: bert_model at 1e-6, the rest at 1e-5
    opt = topt.Adam(m, lr=1e-5)
    ranges = m.lr_ranges(1e-6, 1e-5)
    opt.set_lr_ranges(ranges)
    st = m.train_state()
    before = st.flat.data.clone()
    opt.step()
    step = (st.flat.data - before).abs()
    cut = ranges[0][1]
    assert 0 < cut < st.flat.numel
    assert float(step[:cut].max()) <= 1.05e-6 and float(step[cut:].max()) <= 1.05e-5      # + fp32 rounding of p - p0
    assert float(step[:cut].max()) > 5e-7 and float(step[cut:].max()) > 5e-6      # first Adam step = lr * sign(g)


def test_domain_post_train_model_vs_reference_golden(golden):
    """``tinyrec.post_train.DomainTitleBodySimModel`` against the notebook's own ``TitleBodySimModel`` (fixture generated
    by executing cells 10-11 of Domian-specific_Post-train.ipynb): 12 layers, 1+K titles and a 40-token body per sample
    through the same encoder -- the two-saved-passes path the first-stage KD wrapper shares."""
    from types import SimpleNamespace
    import tinyrec.post_train as pt
    import tinyrec.synth as synth
    g = golden("post_train")
    layers = int(g["layers"])
    m = pt.DomainTitleBodySimModel(SimpleNamespace(news_query_vector_dim=200, news_dim=256, num_hidden_layers=layers))
    full = synth.model_bert_state("", layers, int(g["seed"]), noisy=True)
    sd = {k: v for k, v in full.items() if k.startswith("news_encoder.")}
    assert list(m.state_dict().keys()) == list(sd.keys())
    m.load_state_dict(sd, strict=True)
    m.cuda().eval()
    for p in m.news_encoder.bert_model.parameters():
        p.requires_grad = False
    for i in [int(t) for t in g["trainable"]]:
        for p in m.news_encoder.bert_model.bert.encoder.layer[i].parameters():
            p.requires_grad = True
    scores, loss = m(torch.from_numpy(g["title"]).cuda(), torch.from_numpy(g["body"]).cuda(), torch.from_numpy(g["labels"]).cuda())
    ref = torch.from_numpy(g["scores"])
    assert _rel(scores, ref) < 1e-2
    serr = float((scores.detach().cpu() - ref).abs().max())
    assert abs(float(loss) - float(g["loss"])) < 2 * serr + 1e-3            # |d CE| <= 2 max |d score|
    loss.backward()
    named = dict(m.named_parameters())
    for k in [str(s) for s in g["trainable_names"]]:
        gr = named[k].grad
        assert gr is not None, k
        if f"gfull/{k}" in g.files:
            refg = torch.from_numpy(g[f"gfull/{k}"])
            r = _grad_err(k, gr.reshape(refg.shape), refg, named)
        else:
            refg = torch.from_numpy(g[f"gslice/{k}"])
            r = _grad_err(k, gr[:16, :16], refg, named)
        _chk(f"test_domain_post_train_model_vs_reference_golden.grad.{k}", r, GRAD_TOL_12L)             # 12 bf16 layers in front of the loss; the scores are ~60 with gaps of ~1


def test_scoring_path_sees_parameters_updated_by_the_fused_adam(golden):
    """The scoring kernels read a packed copy of att_fc1.weight that is cached per parameter; the fused Adam rewrites
    parameters through raw pointers (torch's version counters do not move), so the cache must key on the library's own
    parameter generation: user vectors computed after a training step must match the per-impression kernel that
    reads the live weights."""
    import tinyrec.ops as ops
    import tinyrec.optim as topt
    g = golden("kd")
    m, inputs = _kd_model(g, False)
    for p in m.student.user_encoder.parameters():
        p.requires_grad = True
    opt = topt.Adam(m, lr=1e-2)
    ue = m.student.user_encoder
    B, H, D = 96, inputs[0].shape[1], 256
    gen = torch.Generator(device="cuda").manual_seed(1)
    vecs = torch.randn(B, H, D, device="cuda", generator=gen) * 0.3
    mask = torch.ones(B, H, device="cuda")
    u_before = ue(vecs, mask).clone()                 # B >= 64: scoring path, packs W1
    opt.zero_grad()
    m(*inputs)[0].backward()
    opt.step()
    u_after = ue(vecs, mask)
    moved = _rel(u_after, u_before)
    assert moved > 1e-3                               # the weights moved
    at = ue.attn
    ref = torch.empty(B, D, device="cuda")
    a = torch.empty(B, H, device="cuda")
    ops.user_encoder_fwd(vecs.view(B * H, D), mask, ue.pad_doc.view(-1), at.att_fc1.weight, at.att_fc1.bias,
                         at.att_fc2.weight.view(-1), at.att_fc2.bias, False, ref, a, None, B, H)
    # the scoring GEMM truncates the history rows to TF32 (tcgen05 kind::tf32), the per-impression kernel rounds them:
    # ~1e-4 between the two, far below the movement a stale packed copy would leave
    err = _rel(u_after, ref)
    assert err < 5e-4 and moved > 5 * err, (err, moved)


def test_graphed_train_step_advances_dropout_seed_and_trains():
    """Training mode (dropout active) under graph replay: the device-resident seed advances inside the graph, so
    replays of the same batch draw different masks (different losses), the Adam step counter advances, and the
    parameters move by about lr per replay (first Adam steps are lr * sign(g)); equality with the eager loop is
    test_graphed_train_step_matches_eager_loop."""
    import tinyrec.model_bert as mb
    import tinyrec.optim as topt
    import tinyrec.run as trun
    import tinyrec.synth as synth
    B, H, K, L, M, layers, N = 4, 10, 3, 12, 2, 2, 300
    news = synth.news_table(N, L=L, seed=1)
    tables = synth.teacher_tables(N, M, 256, seed=3)
    hist_idx, hmask, cand_idx, label = synth.train_impressions(B, N, H, K, seed=10)
    batch = (torch.from_numpy(news[hist_idx].astype(np.int64)).cuda(), torch.from_numpy(hmask).cuda(),
             torch.from_numpy(news[cand_idx].astype(np.int64)).cuda(), torch.from_numpy(label).cuda(),
             [torch.from_numpy(t[hist_idx]).cuda() for t in tables], [torch.from_numpy(t[cand_idx]).cuda() for t in tables])
    m = mb.Model(synth.demo_args(num_student_layers=layers, num_teachers=M, user_log_length=H))
    m.load_state_dict(synth.kd_model_state(layers, M, 5, noisy=True), strict=True)
    m.cuda().train()
    _apply_freeze(m, [1])
    w0 = m.student.news_encoder.dense.weight.detach().clone()
    opt = topt.Adam(m, lr=1e-4)
    step = trun.GraphedTrainStep(m, opt, batch)
    l1 = float(step(*batch)[0])
    l2 = float(step(*batch)[0])
    assert l1 != l2                                           # new dropout masks (and new weights) every replay
    for _ in range(6):
        step(*batch)
    assert opt.steps_done() == 8
    step.close()
    moved = (m.student.news_encoder.dense.weight.detach() - w0).abs()
    assert 1e-4 < float(moved.max()) <= 8.5e-4 and float(moved.mean()) > 5e-5
    assert math.isfinite(l1) and math.isfinite(l2)
