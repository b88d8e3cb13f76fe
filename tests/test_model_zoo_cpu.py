"""Model families beyond GPT / Llama / MoE (the reference's v1 model zoo: hetu/v1/examples/{nlp/bert,cnn,ctr,rec}): parity with
HuggingFace / PyTorch on converted weights, and short training runs."""
import numpy as np
import pytest
import torch

import hetu_b200 as ht


def test_bert_matches_huggingface_on_converted_weights_and_trains():
    transformers = pytest.importorskip("transformers")
    from hetu_b200.models import BertConfig, BertForPreTraining, convert_bert_hf_to_ht
    torch.manual_seed(0)
    L, H, NH, F, V, S, B = 2, 32, 4, 64, 101, 10, 3
    hf_cfg = transformers.BertConfig(vocab_size=V, hidden_size=H, num_hidden_layers=L, num_attention_heads=NH, intermediate_size=F, max_position_embeddings=S,
                                     hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0)
    hf = transformers.BertForPreTraining(hf_cfg).eval()
    sd = convert_bert_hf_to_ht({k: v.detach().clone() for k, v in hf.state_dict().items()}, L)
    ids = torch.randint(0, V, (B, S))
    tt = torch.randint(0, 2, (B, S))
    mask = torch.ones(B, S, dtype=torch.long)
    mask[1, 7:] = 0
    mask[2, 4:] = 0
    with torch.no_grad():
        full = hf(ids, token_type_ids=tt)
        padded = hf(ids, token_type_ids=tt, attention_mask=mask)
    cfg = BertConfig(vocab_size=V, hidden_size=H, num_hidden_layers=L, num_attention_heads=NH, intermediate_size=F, max_position_embeddings=S,
                     hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0)
    with ht.graph("define_and_run", create_new=True) as g:
        m = BertForPreTraining(cfg)
        X = ht.placeholder("int64", [B, S], name="ids")
        T = ht.placeholder("int64", [B, S], name="tt")
        M = ht.placeholder("float32", [B, S], name="mask")
        mlm, nsp = m(X, T)
        mlm_m, nsp_m = m(X, T, attention_mask=M)
        missing = m.load_state_dict(sd, strict=False)
        out = g.run(mlm, [mlm, nsp, mlm_m, nsp_m], {X: ids, T: tt, M: mask.float()})
    assert torch.allclose(out[0].float(), full.prediction_logits.reshape(B * S, V), atol=3e-4, rtol=1e-4)
    assert torch.allclose(out[1].float(), full.seq_relationship_logits, atol=3e-4, rtol=1e-4)
    keep = mask.reshape(-1).bool()                                   # padded query rows are garbage in both implementations: compare real tokens
    assert torch.allclose(out[2].float()[keep], padded.prediction_logits.reshape(B * S, V)[keep], atol=3e-4, rtol=1e-4)
    assert torch.allclose(out[3].float(), padded.seq_relationship_logits, atol=3e-4, rtol=1e-4)

    # pre-training step: MLM (ignored positions = -100) + NSP loss falls on a fixed batch
    with ht.graph("define_and_run", create_new=True) as g:
        m = BertForPreTraining(BertConfig(vocab_size=V, hidden_size=H, num_hidden_layers=L, num_attention_heads=NH, intermediate_size=F,
                                          max_position_embeddings=S, hidden_dropout_prob=0.1, attention_probs_dropout_prob=0.1))
        X = ht.placeholder("int64", [B, S], name="ids")
        T = ht.placeholder("int64", [B, S], name="tt")
        Y = ht.placeholder("int64", [B, S], name="mlm_labels")
        N = ht.placeholder("int64", [B], name="nsp_labels")
        loss, _, _ = m(X, T, masked_lm_labels=Y, next_sentence_label=N)
        train = ht.AdamOptimizer(lr=2e-3).minimize(loss)
        labels = torch.full((B, S), -100, dtype=torch.long)
        labels[:, ::3] = ids[:, ::3]
        nsl = torch.tensor([0, 1, 0])
        losses = [float(g.run(loss, [loss, train], {X: ids, T: tt, Y: labels, N: nsl})[0]) for _ in range(30)]
    assert losses[-1] < 0.6 * losses[0], (losses[0], losses[-1])


def test_bert_sequence_classifier_learns_a_separable_task():
    from hetu_b200.models import BertConfig, BertForSequenceClassification
    rng = np.random.RandomState(0)
    V, S, B = 40, 8, 16
    cfg = BertConfig(vocab_size=V, hidden_size=32, num_hidden_layers=1, num_attention_heads=2, intermediate_size=64, max_position_embeddings=S,
                     hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0, num_labels=2)
    with ht.graph("define_and_run", create_new=True) as g:
        m = BertForSequenceClassification(cfg)
        X = ht.placeholder("int64", [B, S], name="ids")
        Y = ht.placeholder("int64", [B], name="y")
        loss, logits = m(X, labels=Y)
        train = ht.AdamOptimizer(lr=3e-3).minimize(loss)
        acc = []
        for step in range(60):
            ids = rng.randint(2, V, (B, S))
            y = rng.randint(0, 2, B)
            ids[:, 1] = y                                            # the label is the token at position 1
            out = g.run(loss, [loss, logits, train], {X: torch.as_tensor(ids), Y: torch.as_tensor(y)})
            acc.append(float((out[1].float().argmax(1).numpy() == y).mean()))
    assert np.mean(acc[-10:]) > 0.9, acc[-10:]
