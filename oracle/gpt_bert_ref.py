"""Single-process restatement of the GPT and BERT layer math the reference runtime executes (TEST INFRASTRUCTURE ONLY).

Plain torch, no parallelism, no custom kernels -- what one rank of the reference computes when every group has size 1:
  GPT   gpt_hf/GPTModel_sequential.py:44-100 (wte + wpe, [b,s,h] -> [s,b,h]), GPTModel_tensor_parallel.py:14-82 (pre-LayerNorm
        blocks, biases on every projection, residual), tensor_parallel/transformer.py:150-160 (bias + tanh-GeLU, ``bias_gelu_impl``),
        :453-509 (causal softmax attention, scale 1/sqrt(hn)), GPTModel_sequential.py:120-186 (ln_f, untied lm_head, per-token CE)
  BERT  bert_hf/BertModel_sequential.py:41-108 (word + position + token-type embeddings, embedding LayerNorm),
        BertModel_tensor_parallel.py:11-72 (POST-LayerNorm blocks), padding mask over the keys (:12-18), :136-226 (MLM head:
        dense + GeLU + LayerNorm, decoder with bias, megatron's vocab-parallel cross entropy, in which a label outside the
        vocabulary -- the -100 of unmasked positions -- contributes logsumexp - max: cross_entropy.py:45-60)
The fused QKV weight is in the per-head layout [n_heads, 3 * hn] of transformer.py:733-756 (q_h | k_h | v_h per head).
``dtype`` = torch.bfloat16 rounds every op's output to bf16; float32/float64 is exact math.  Pinned against HF ``GPT2LMHeadModel`` /
``BertForMaskedLM`` in fp64 by tests/test_oracle_gpt_bert.py (the baselines of the reference's tests/models/test_model_correctness.py).
"""
import math

import torch
import torch.nn.functional as F


def _r(x, dtype):
    if dtype in (torch.float32, torch.float64):
        return x.to(dtype)
    return x.to(dtype).float()


def layer_norm(x, w, b, eps, dtype):
    xf = x.double() if dtype == torch.float64 else x.float()
    mean = xf.mean(-1, keepdim=True)
    var = (xf - mean).pow(2).mean(-1, keepdim=True)
    return _r((xf - mean) * torch.rsqrt(var + eps) * w + b, dtype)


def gelu(x, tanh_form):
    return F.gelu(x, approximate="tanh" if tanh_form else "none")


def attention(h, p, cfg, dtype, causal, key_mask):
    """h [s, b, hidden] (already normed where the family pre-norms) -> attention output + bias, [s, b, hidden]."""
    s, b, _ = h.shape
    n, hn = cfg["n_heads"], cfg["head_dim"]
    mixed = _r(h @ p["qkv"].t() + p["qkv_b"], dtype).view(s, b, n, 3 * hn)
    q, k, v = torch.split(mixed, [hn, hn, hn], dim=3)
    qb, kb, vb = [t.permute(1, 2, 0, 3) for t in (q, k, v)]               # [b, n, s, hn]
    scores = (qb @ kb.transpose(-1, -2)) / math.sqrt(hn)
    if causal:
        scores = scores.masked_fill(torch.triu(torch.ones(s, s, dtype=torch.bool, device=h.device), diagonal=1), float("-inf"))
    if key_mask is not None:
        scores = scores.masked_fill(~key_mask.bool()[:, None, None, :], float("-inf"))
    ctxt = _r(torch.softmax(scores, dim=-1) @ vb, dtype)
    ctxt = ctxt.permute(2, 0, 1, 3).reshape(s, b, n * hn)
    return _r(_r(ctxt @ p["dense"].t(), dtype) + p["dense_b"], dtype)


def mlp(h, p, cfg, dtype):
    inter = _r(h @ p["h_to_4h"].t(), dtype)                                   # the bias is added inside the fused GeLU
    act = _r(gelu(inter + p["h_to_4h_b"], cfg.get("gelu_tanh", True)), dtype)
    return _r(_r(act @ p["4h_to_h"].t(), dtype) + p["4h_to_h_b"], dtype)


def block(h, p, cfg, dtype, post_ln, causal, key_mask):
    if post_ln:   # BERT: sublayer, residual, LayerNorm
        h = layer_norm(_r(attention(h, p, cfg, dtype, causal, key_mask) + h, dtype), p["ln1"], p["ln1_b"], cfg["eps"], dtype)
        return layer_norm(_r(mlp(h, p, cfg, dtype) + h, dtype), p["ln2"], p["ln2_b"], cfg["eps"], dtype)
    x = layer_norm(h, p["ln1"], p["ln1_b"], cfg["eps"], dtype)
    h = _r(attention(x, p, cfg, dtype, causal, key_mask) + h, dtype)
    x = layer_norm(h, p["ln2"], p["ln2_b"], cfg["eps"], dtype)
    return _r(mlp(x, p, cfg, dtype) + h, dtype)


def _token_loss(logits, labels, dtype):
    """per-token CE [s, b] as megatron's vocab-parallel CE computes it (cross_entropy.py:22-100): the row maximum is subtracted
    first, the target's SHIFTED logit is looked up -- and set to 0 for a label outside [0, V) (BERT's -100 on unmasked positions,
    :45-60) -- so such a position contributes log(sum exp(l - max)) = logsumexp - max."""
    lf = logits if dtype == torch.float64 else logits.float()
    # (the maximum is a constant for the hand-written backward, cross_entropy.py:103-152: grad = softmax - onehot * inside)
    shifted = lf - lf.max(dim=-1, keepdim=True).values.detach()
    inside = (labels >= 0) & (labels < lf.shape[-1])
    pred = torch.where(inside, shifted.gather(-1, labels.clamp(0, lf.shape[-1] - 1).unsqueeze(-1)).squeeze(-1), torch.zeros_like(shifted[..., 0]))
    return torch.log(torch.exp(shifted).sum(-1)) - pred


def gpt_forward_loss(weights, tokens, labels, cfg, dtype=torch.float32):
    """tokens, labels [b, s] -> (per-token loss [b, s], scalar mean).  weights: wte [V,h], wpe [S,h], layers [{ln1, ln1_b, qkv, qkv_b,
    dense, dense_b, ln2, ln2_b, h_to_4h, h_to_4h_b, 4h_to_h, 4h_to_h_b}], norm, norm_b, lm_head [V,h]."""
    wd = lambda t: _r(t, dtype) if t.dtype != torch.float64 else t  # noqa: E731
    s = tokens.shape[1]
    h = _r(wd(weights["wte"])[tokens] + wd(weights["wpe"])[torch.arange(s)][None], dtype).transpose(0, 1)   # [s, b, h]
    for p in weights["layers"]:
        h = block(h, {k: wd(v) for k, v in p.items()}, cfg, dtype, post_ln=False, causal=True, key_mask=None)
    h = layer_norm(h, wd(weights["norm"]), wd(weights["norm_b"]), cfg["eps"], dtype)
    logits = _r(h @ wd(weights["lm_head"]).t(), dtype)
    loss = _token_loss(logits, labels.transpose(0, 1), dtype).transpose(0, 1)
    return loss, loss.mean()


def bert_forward_loss(weights, tokens, labels, cfg, dtype=torch.float32, attention_mask=None, token_type_ids=None):
    """weights: word [V,h], pos [S,h], type [2,h], emb_ln, emb_ln_b, layers [...], transform (dense [h,h]), transform_b, transform_ln,
    transform_ln_b, decoder [V,h], decoder_b [V]."""
    wd = lambda t: _r(t, dtype) if t.dtype != torch.float64 else t  # noqa: E731
    s = tokens.shape[1]
    tt = torch.zeros_like(tokens) if token_type_ids is None else token_type_ids
    e = wd(weights["word"])[tokens] + wd(weights["pos"])[torch.arange(s)][None] + wd(weights["type"])[tt]
    h = layer_norm(_r(e, dtype), wd(weights["emb_ln"]), wd(weights["emb_ln_b"]), cfg["eps"], dtype).transpose(0, 1)
    for p in weights["layers"]:
        h = block(h, {k: wd(v) for k, v in p.items()}, cfg, dtype, post_ln=True, causal=False, key_mask=attention_mask)
    t = _r(gelu(_r(h @ wd(weights["transform"]).t(), dtype) + wd(weights["transform_b"]), cfg.get("gelu_tanh", True)), dtype)
    t = layer_norm(t, wd(weights["transform_ln"]), wd(weights["transform_ln_b"]), cfg["eps"], dtype)
    logits = _r(_r(t @ wd(weights["decoder"]).t(), dtype) + wd(weights["decoder_b"]), dtype)
    loss = _token_loss(logits, labels.transpose(0, 1), dtype).transpose(0, 1)
    return loss, loss.mean()


def _layer(cfg, rnd, dtype):
    h, ffn, n, hn = cfg["hidden"], cfg["ffn"], cfg["n_heads"], cfg["head_dim"]
    one, zero = (lambda k: torch.ones(k, dtype=dtype)), (lambda k: torch.zeros(k, dtype=dtype))
    return {"ln1": one(h), "ln1_b": zero(h), "qkv": rnd(3 * n * hn, h), "qkv_b": zero(3 * n * hn), "dense": rnd(h, n * hn), "dense_b": zero(h),
            "ln2": one(h), "ln2_b": zero(h), "h_to_4h": rnd(ffn, h), "h_to_4h_b": zero(ffn), "4h_to_h": rnd(h, ffn), "4h_to_h_b": zero(h)}


def init_weights(cfg, family, seed=0, std=0.02, dtype=torch.float32):
    g = torch.Generator(device="cpu").manual_seed(seed)
    rnd = lambda *shape: (torch.randn(*shape, generator=g) * std).to(dtype)  # noqa: E731
    h, V, S = cfg["hidden"], cfg["vocab"], cfg["seq"]
    layers = [_layer(cfg, rnd, dtype) for _ in range(cfg["n_layers"])]
    if family == "gpt":
        return dict(wte=rnd(V, h), wpe=rnd(S, h), layers=layers, norm=torch.ones(h, dtype=dtype), norm_b=torch.zeros(h, dtype=dtype),
                    lm_head=rnd(V, h))
    return dict(word=rnd(V, h), pos=rnd(S, h), type=rnd(2, h), emb_ln=torch.ones(h, dtype=dtype), emb_ln_b=torch.zeros(h, dtype=dtype),
                layers=layers, transform=rnd(h, h), transform_b=torch.zeros(h, dtype=dtype), transform_ln=torch.ones(h, dtype=dtype),
                transform_ln_b=torch.zeros(h, dtype=dtype), decoder=rnd(V, h), decoder_b=torch.zeros(V, dtype=dtype))


def split_qkv(p, cfg):
    """fused per-head [n, 3*hn, h] weight and [n, 3*hn] bias -> (q, k, v) weights [n*hn, h] and biases [n*hn] (HF's order)."""
    n, hn = cfg["n_heads"], cfg["head_dim"]
    w = p["qkv"].view(n, 3, hn, -1)
    b = p["qkv_b"].view(n, 3, hn)
    return [w[:, i].reshape(n * hn, -1) for i in range(3)], [b[:, i].reshape(n * hn) for i in range(3)]


def to_hf_state_dict(w, cfg, family):
    """Megatron-layout oracle weights -> the HF ``GPT2LMHeadModel`` / ``BertForMaskedLM`` state dict they correspond to (the mapping
    tests/test_oracle_gpt_bert.py pins; inverse of what the families' checkpoint loaders do).  GPT-2's ``Conv1D`` weights are [in, out]."""
    sd = {}
    if family == "gpt":
        sd.update({"transformer.wte.weight": w["wte"], "transformer.wpe.weight": w["wpe"], "transformer.ln_f.weight": w["norm"],
                   "transformer.ln_f.bias": w["norm_b"], "lm_head.weight": w["lm_head"]})
        for i, p in enumerate(w["layers"]):
            (q, k, v), (qb, kb, vb) = split_qkv(p, cfg)
            pre = "transformer.h.%d." % i
            sd.update({pre + "ln_1.weight": p["ln1"], pre + "ln_1.bias": p["ln1_b"], pre + "ln_2.weight": p["ln2"], pre + "ln_2.bias": p["ln2_b"],
                       pre + "attn.c_attn.weight": torch.cat([q, k, v]).t(), pre + "attn.c_attn.bias": torch.cat([qb, kb, vb]),
                       pre + "attn.c_proj.weight": p["dense"].t(), pre + "attn.c_proj.bias": p["dense_b"],
                       pre + "mlp.c_fc.weight": p["h_to_4h"].t(), pre + "mlp.c_fc.bias": p["h_to_4h_b"],
                       pre + "mlp.c_proj.weight": p["4h_to_h"].t(), pre + "mlp.c_proj.bias": p["4h_to_h_b"]})
        return sd
    sd.update({"bert.embeddings.word_embeddings.weight": w["word"], "bert.embeddings.position_embeddings.weight": w["pos"],
               "bert.embeddings.token_type_embeddings.weight": w["type"], "bert.embeddings.LayerNorm.weight": w["emb_ln"],
               "bert.embeddings.LayerNorm.bias": w["emb_ln_b"], "cls.predictions.transform.dense.weight": w["transform"],
               "cls.predictions.transform.dense.bias": w["transform_b"], "cls.predictions.transform.LayerNorm.weight": w["transform_ln"],
               "cls.predictions.transform.LayerNorm.bias": w["transform_ln_b"], "cls.predictions.decoder.weight": w["decoder"],
               "cls.predictions.decoder.bias": w["decoder_b"], "cls.predictions.bias": w["decoder_b"]})
    for i, p in enumerate(w["layers"]):
        (q, k, v), (qb, kb, vb) = split_qkv(p, cfg)
        pre = "bert.encoder.layer.%d." % i
        sd.update({pre + "attention.self.query.weight": q, pre + "attention.self.query.bias": qb, pre + "attention.self.key.weight": k,
                   pre + "attention.self.key.bias": kb, pre + "attention.self.value.weight": v, pre + "attention.self.value.bias": vb,
                   pre + "attention.output.dense.weight": p["dense"], pre + "attention.output.dense.bias": p["dense_b"],
                   pre + "attention.output.LayerNorm.weight": p["ln1"], pre + "attention.output.LayerNorm.bias": p["ln1_b"],
                   pre + "intermediate.dense.weight": p["h_to_4h"], pre + "intermediate.dense.bias": p["h_to_4h_b"],
                   pre + "output.dense.weight": p["4h_to_h"], pre + "output.dense.bias": p["4h_to_h_b"],
                   pre + "output.LayerNorm.weight": p["ln2"], pre + "output.LayerNorm.bias": p["ln2_b"]})
    return sd
