"""Oracle (test infrastructure): fp32 restatement of the reference model path.

Functional style over a flat ``state_dict`` (the reference's own key names), so
the same tensors can be handed to the CUDA path and to the oracle.  Dropout
(active in the reference's train(), which never leaves training mode) is modelled
by INJECTED masks: ``drop=(plan, row0)`` with an ``oracle.dropout.Plan`` reproducing
the CUDA path's counter-based generator; ``drop=None`` is ``eval()`` mode.

Citations are relative to /root/reference.
"""
import math

import torch
import torch.nn.functional as F

LN_EPS = 1e-12  # tnlrv3/configuration_tnlrv3.py:61 (layer_norm_eps default)


# --------------------------------------------------------------------------
# relative-position bias  (Tiny-NewsRec/tnlrv3/modeling.py:345-373, 458-463)
# --------------------------------------------------------------------------
def rel_pos_bucket(rel, num_buckets=32, max_distance=128):
    """Bidirectional T5-style bucket of ``rel = pos_key - pos_query``.

    modeling.py:345-373: half the buckets per sign; |d| < 8 exact, log-spaced
    beyond up to ``max_distance``; keys to the right get +num_buckets/2.
    """
    half = num_buckets // 2
    out = (rel > 0).long() * half
    n = rel.abs()
    exact = half // 2
    big = exact + (torch.log(n.float() / exact) / math.log(max_distance / exact)
                   * (half - exact)).to(torch.long)
    big = torch.clamp(big, max=half - 1)
    return out + torch.where(n < exact, n, big)


def rel_pos_bias_table(rel_pos_weight, L, num_buckets=32, max_distance=128):
    """[A, L, L] additive bias.  The reference builds a one-hot [n, L, L, 32]
    per forward and applies Linear(32 -> A, no bias) (modeling.py:458-463);
    position_ids are always arange(L) (modeling.py:162-163) so the result is
    batch-invariant and equals a column lookup of the weight."""
    pos = torch.arange(L)
    rel = pos.unsqueeze(-2) - pos.unsqueeze(-1)          # [i, j] = j - i
    bucket = rel_pos_bucket(rel, num_buckets, max_distance)   # [L, L]
    return rel_pos_weight[:, bucket]                     # [A, L, L]


# --------------------------------------------------------------------------
# encoder
# --------------------------------------------------------------------------
def embeddings(sd, pfx, ids, drop=None):
    """dropout(LN(word[id] + pos[arange L] + type[0]))   (modeling.py:153-178)."""
    L = ids.shape[1]
    x = sd[pfx + "word_embeddings.weight"][ids]
    x = x + sd[pfx + "position_embeddings.weight"][:L].unsqueeze(0)
    x = x + sd[pfx + "token_type_embeddings.weight"][0]
    x = F.layer_norm(x, (x.shape[-1],), sd[pfx + "LayerNorm.weight"],
                     sd[pfx + "LayerNorm.bias"], LN_EPS)
    if drop is not None:
        x = x * drop[0].emb(drop[1], x.shape[0], L, x.shape[-1])
    return x


def self_attention(sd, pfx, x, ext_mask, relpos, heads, drop=None, layer=0):
    """modeling.py:205-231, 233-272: QKV linears, scores/sqrt(dh) + mask + relpos,
    softmax, PV, merge heads."""
    n, L, E = x.shape
    dh = E // heads

    def proj(name):
        y = F.linear(x, sd[pfx + name + ".weight"], sd[pfx + name + ".bias"])
        return y.view(n, L, heads, dh).permute(0, 2, 1, 3)

    q, k, v = proj("query"), proj("key"), proj("value")
    s = torch.matmul(q, k.transpose(-1, -2)) / math.sqrt(dh)
    s = s + ext_mask + relpos
    p = torch.softmax(s, dim=-1)
    if drop is not None:                                     # modeling.py:223
        p = p * drop[0].attn(layer, drop[1], n, heads, L)
    ctx = torch.matmul(p, v).permute(0, 2, 1, 3).reshape(n, L, E)
    return ctx


def encoder_layer(sd, pfx, x, ext_mask, relpos, heads, drop=None, layer=0):
    """modeling.py:275-308 plus transformers BertSelfOutput / BertIntermediate /
    BertOutput (third-party, imported at modeling.py:12-14): post-LN residual
    blocks with erf-GELU."""
    E = x.shape[-1]
    n, L = x.shape[0], x.shape[1]
    ctx = self_attention(sd, pfx + "attention.self.", x, ext_mask, relpos, heads, drop, layer)
    a = F.linear(ctx, sd[pfx + "attention.output.dense.weight"], sd[pfx + "attention.output.dense.bias"])
    if drop is not None:                                     # BertSelfOutput.dropout
        a = a * drop[0].dense(layer, 1, drop[1], n, L, E)
    a = F.layer_norm(a + x, (E,), sd[pfx + "attention.output.LayerNorm.weight"],
                     sd[pfx + "attention.output.LayerNorm.bias"], LN_EPS)
    h = F.gelu(F.linear(a, sd[pfx + "intermediate.dense.weight"], sd[pfx + "intermediate.dense.bias"]))
    o = F.linear(h, sd[pfx + "output.dense.weight"], sd[pfx + "output.dense.bias"])
    if drop is not None:                                     # BertOutput.dropout
        o = o * drop[0].dense(layer, 2, drop[1], n, L, E)
    return F.layer_norm(o + a, (E,), sd[pfx + "output.LayerNorm.weight"],
                        sd[pfx + "output.LayerNorm.bias"], LN_EPS)


def bert_last_hidden(sd, pfx, ids, mask, num_layers, heads=12, all_hidden=False, drop=None):
    """TuringNLRv3Model.forward (modeling.py:421-476) up to the last hidden
    state; the pooler/classifier outputs are discarded by the caller
    (model_bert.py:128-129) and are not computed here."""
    ext_mask = (1.0 - mask.to(torch.float32))[:, None, None, :] * -10000.0   # modeling.py:446-454
    x = embeddings(sd, pfx + "bert.embeddings.", ids, drop)
    relpos = rel_pos_bias_table(sd[pfx + "bert.rel_pos_bias.weight"], ids.shape[1]).unsqueeze(0)
    hs = [x]
    for l in range(num_layers):                                               # modeling.py:318-342
        x = encoder_layer(sd, f"{pfx}bert.encoder.layer.{l}.", x, ext_mask, relpos, heads, drop, l)
        hs.append(x)
    return hs if all_hidden else x


def attention_pooling(sd, pfx, x, mask=None):
    """model_bert.py:15-34: alpha = exp(fc2(tanh(fc1 x))) [* mask]; alpha /= sum+1e-8."""
    e = torch.tanh(F.linear(x, sd[pfx + "att_fc1.weight"], sd[pfx + "att_fc1.bias"]))
    alpha = torch.exp(F.linear(e, sd[pfx + "att_fc2.weight"], sd[pfx + "att_fc2.bias"]))
    if mask is not None:
        alpha = alpha * mask.unsqueeze(2)
    alpha = alpha / (alpha.sum(dim=1, keepdim=True) + 1e-8)
    return torch.bmm(x.permute(0, 2, 1), alpha).squeeze(-1)


def news_encoder(sd, pfx, x, num_layers, heads=12, drop=None):
    """model_bert.py:119-137 (pooling='att'): split ids|mask, encoder, UNMASKED
    additive pooling over words, dense E->D."""
    L = x.shape[1] // 2
    ids, mask = x[:, :L], x[:, L:]
    h = bert_last_hidden(sd, pfx + "bert_model.", ids, mask, num_layers, heads, drop=drop)
    pooled = attention_pooling(sd, pfx + "attn.", h)
    return F.linear(pooled, sd[pfx + "dense.weight"], sd[pfx + "dense.bias"])


def multi_head_self_attn(sd, pfx, x, mask=None, d_k=16, d_v=16):
    """NRMS user encoder, model_bert.py:37-100: Q/K/V projections, per-head
    ``exp(QK^T / sqrt(d_k))`` (NOT a softmax: no max subtraction) [* key mask],
    normalised by ``sum + 1e-8``, times V; heads concatenated."""
    B, H, _ = x.shape
    n_heads = sd[pfx + "W_Q.weight"].shape[0] // d_k
    q = F.linear(x, sd[pfx + "W_Q.weight"], sd[pfx + "W_Q.bias"]).view(B, H, n_heads, d_k).transpose(1, 2)
    k = F.linear(x, sd[pfx + "W_K.weight"], sd[pfx + "W_K.bias"]).view(B, H, n_heads, d_k).transpose(1, 2)
    v = F.linear(x, sd[pfx + "W_V.weight"], sd[pfx + "W_V.bias"]).view(B, H, n_heads, d_v).transpose(1, 2)
    scores = torch.exp(torch.matmul(q, k.transpose(-1, -2)) / math.sqrt(d_k))            # :53-54
    if mask is not None:
        scores = scores * mask[:, None, None, :]                                          # :56-57
    attn = scores / (scores.sum(dim=-1, keepdim=True) + 1e-8)                             # :59
    ctx = torch.matmul(attn, v)
    return ctx.transpose(1, 2).contiguous().view(B, H, n_heads * d_v)


def user_encoder(sd, pfx, vecs, log_mask, user_log_mask):
    """model_bert.py:155-176.  NAML branches: masked pool (:166) or pad_doc blend then unmasked pool
    (:168-175).  NRMS branches (a state dict with ``multi_head_self_attn.*``): the same with the
    multi-head self-attention in front of the pooling (:162-164, :171-173)."""
    nrms = (pfx + "multi_head_self_attn.W_Q.weight") in sd
    if user_log_mask:
        if nrms:
            vecs = multi_head_self_attn(sd, pfx + "multi_head_self_attn.", vecs, log_mask)
        return attention_pooling(sd, pfx + "attn.", vecs, log_mask)
    m = log_mask.unsqueeze(-1)
    blended = vecs * m + sd[pfx + "pad_doc"].unsqueeze(0) * (1 - m)
    if nrms:
        blended = multi_head_self_attn(sd, pfx + "multi_head_self_attn.", blended)
    return attention_pooling(sd, pfx + "attn.", blended)


def model_bert_forward(sd, pfx, history, history_mask, candidate, num_layers, user_log_mask,
                       heads=12, drop_plan=None):
    """ModelBert.forward, model_bert.py:187-205 ->
    (score[B,K], hist_vecs[B,H,D], cand_vecs[B,K,D], user_vec[B,D])."""
    B, H, W = history.shape
    # the CUDA path encodes [history rows | candidate rows] as one batch: dropout rows are offset accordingly
    dc = (drop_plan, B * H) if drop_plan is not None else None
    dh = (drop_plan, 0) if drop_plan is not None else None
    cand = news_encoder(sd, pfx + "news_encoder.", candidate.reshape(-1, W), num_layers, heads, dc)
    cand = cand.reshape(B, -1, cand.shape[-1])
    hist = news_encoder(sd, pfx + "news_encoder.", history.reshape(-1, W), num_layers, heads, dh)
    hist = hist.reshape(B, H, -1)
    user = user_encoder(sd, pfx + "user_encoder.", hist, history_mask, user_log_mask)
    score = torch.bmm(cand, user.unsqueeze(-1)).squeeze(-1)
    return score, hist, cand, user


def plmnr_forward(sd, history, history_mask, candidate, label, num_layers, user_log_mask):
    """Teacher / PLM-NR form: (loss, score)   (model_bert_2.py:192-213,
    PLM-NR/model_bert.py:187-207)."""
    score, _, _, _ = model_bert_forward(sd, "", history, history_mask, candidate, num_layers,
                                        user_log_mask)
    return F.cross_entropy(score, label), score


def kd_ce_loss(logits_s, logits_t, temperature=1.0):
    """model_bert.py:208-219 (no tau^2 factor)."""
    p_t = torch.softmax(logits_t / temperature, dim=-1)
    return -(p_t * torch.log_softmax(logits_s / temperature, dim=-1)).sum(-1).mean()


def kd_model_forward(sd, history, history_mask, candidate, label, teacher_history_embs,
                     teacher_candidate_embs, num_layers, user_log_mask, temperature=1.0,
                     coef=1.0, drop_plan=None):
    """Model.forward, model_bert.py:262-306 ->
    (total, distill, emb, target, student_score)."""
    s_score, s_hist, s_cand, s_user = model_bert_forward(
        sd, "student.", history, history_mask, candidate, num_layers, user_log_mask, drop_plan=drop_plan)
    s_news = torch.cat([s_hist, s_cand], dim=1)
    target = F.cross_entropy(s_score, label)                                   # :271
    t_scores, t_losses, ne, ue = [], [], [], []
    for i, (th, tc) in enumerate(zip(teacher_history_embs, teacher_candidate_embs)):
        W, b = sd[f"transform_matrix.{i}.weight"], sd[f"transform_matrix.{i}.bias"]
        t_news = F.linear(torch.cat([th, tc], dim=1), W, b)                    # :277-278
        ne.append(((s_news - t_news) ** 2).mean(-1).mean(-1))                  # :279-280 -> [B]
        t_user = user_encoder(sd, f"teachers.{i}.", th, history_mask, user_log_mask)   # :282
        ue.append(((s_user - F.linear(t_user, W, b)) ** 2).mean(-1))           # :283-284 -> [B]
        sc = torch.bmm(tc, t_user.unsqueeze(-1)).squeeze(-1)                   # :286-287
        t_scores.append(sc)
        t_losses.append(F.cross_entropy(sc, label, reduction="none"))          # :288
    w = torch.softmax(-torch.stack(t_losses, -1), dim=-1)                      # :292-293  [B,M]
    t_score = torch.bmm(torch.stack(t_scores, -1), w.unsqueeze(-1)).squeeze(-1)    # :295-297
    distill = kd_ce_loss(s_score, t_score, temperature)                        # :298
    emb = (torch.stack(ne, -1) * w).sum(-1).mean() + (torch.stack(ue, -1) * w).sum(-1).mean()  # :300-303
    total = distill + coef * target + emb                                      # :305
    return total, distill, emb, target, s_score


def post_train_distill_forward(sd, title, body, labels, teacher_titles, teacher_bodies, num_layers):
    """First-stage KD, Post-train_KD.ipynb cells 12 and 14 (TitleBodySimModel / DistillModel.forward) ->
    (loss, target_loss, distill_loss, emb_loss, student_score).  PARITY UNPINNED: cell 14 as published multiplies the
    Python list ``teacher_MSEs`` by a tensor (TypeError); this restates the evident intent,
    ``torch.stack(teacher_MSEs, dim=-1) * teacher_weights``, and cannot be checked against a run of the notebook."""
    bz, k1, w = title.shape
    body_emb = news_encoder(sd, "student.news_encoder.", body, num_layers)                         # cell 12
    title_emb = news_encoder(sd, "student.news_encoder.", title.reshape(-1, w), num_layers).reshape(bz, k1, -1)
    s_score = torch.bmm(title_emb, body_emb.unsqueeze(-1)).squeeze(-1)
    target = F.cross_entropy(s_score, labels)
    t_scores, t_losses, mses = [], [], []
    for i, (tt, tb) in enumerate(zip(teacher_titles, teacher_bodies)):
        W, b = sd[f"transform_matrix.{i}.weight"], sd[f"transform_matrix.{i}.bias"]
        sc = torch.bmm(tt, tb.unsqueeze(-1)).squeeze(-1)
        t_scores.append(sc)
        t_losses.append(F.cross_entropy(sc, labels, reduction="none"))
        tt_p, tb_p = F.linear(tt, W, b), F.linear(tb, W, b)
        mses.append(((title_emb - tt_p) ** 2).mean(-1).mean(-1) + ((body_emb - tb_p) ** 2).mean(-1))
    wts = torch.softmax(-torch.stack(t_losses, -1), dim=-1)
    t_score = torch.bmm(torch.stack(t_scores, -1), wts.unsqueeze(-1)).squeeze(-1)
    distill = kd_ce_loss(s_score, t_score)
    emb = (torch.stack(mses, -1) * wts).sum(-1).mean()
    return target + distill + emb, target, distill, emb, s_score


def domain_post_train_forward(sd, title, body, labels, num_layers):
    """Domian-specific_Post-train.ipynb cell 11 (TitleBodySimModel.forward) -> (scores, loss); pinned on the
    notebook's own code by tests/golden/post_train.npz."""
    bz, k1, w = title.shape
    body_emb = news_encoder(sd, "news_encoder.", body, num_layers)
    title_emb = news_encoder(sd, "news_encoder.", title.reshape(-1, w), num_layers).reshape(bz, k1, -1)
    scores = torch.bmm(title_emb, body_emb.unsqueeze(-1)).squeeze(-1)
    return scores, F.cross_entropy(scores, labels)


def accuracy(y_true, y_hat):
    """utils.py:79-83."""
    return (y_true == y_hat.argmax(-1)).sum().float() / y_true.shape[0]


# --------------------------------------------------------------------------
# parameter bookkeeping shared by tests / bench
# --------------------------------------------------------------------------
def trainable_keys(sd, trainable_layers, student_pfx="student."):
    """The 'requires_grad' policy of run.py:101-112: teachers frozen, whole
    bert_model frozen except encoder.layer[i] for i in bert_trainable_layer;
    everything else (pooling head, dense, user encoder, transform_matrix) trains."""
    keys = []
    for k in sd:
        if k.startswith("teachers."):
            continue
        bm = student_pfx + "news_encoder.bert_model."
        if k.startswith(bm):
            rest = k[len(bm):]
            if rest.startswith("bert.encoder.layer."):
                if int(rest.split(".")[3]) in trainable_layers:
                    keys.append(k)
            continue
        keys.append(k)
    return keys
