"""Model families beyond GPT / Llama / MoE (the reference's v1 model zoo: hetu/v1/examples/{nlp/bert,cnn,ctr,rec}): parity with
HuggingFace / PyTorch on converted weights, and short training runs."""
import numpy as np
import pytest
import torch

import hetu_b200 as ht


@pytest.fixture(autouse=True)
def _fixed_seed():
    """parameter initialisers draw from the global seed: pin it so that the short training runs below do not depend on which tests
    ran before"""
    ht.set_seed(1234)
    torch.manual_seed(1234)
    yield


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


def test_lstm_matches_torch_lstm_with_shared_weights():
    from hetu_b200.models import LSTM
    torch.manual_seed(0)
    B, T, F, Hd = 4, 6, 5, 7
    ref = torch.nn.LSTM(F, Hd, batch_first=True)
    x = torch.randn(B, T, F)
    with torch.no_grad():
        seq, (hn, cn) = ref(x)
    with ht.graph("define_and_run", create_new=True) as g:
        m = LSTM(F, Hd, num_classes=3)
        X = ht.placeholder("float32", [B, T, F], name="x")
        hs, (h, c) = m(X, return_sequence=True)
        m.load_state_dict({"wx.weight": ref.weight_ih_l0.detach(), "wx.bias": ref.bias_ih_l0.detach(), "wh.weight": ref.weight_hh_l0.detach(),
                           "wh.bias": ref.bias_hh_l0.detach()}, strict=False)
        out = g.run(h, hs + [h, c], {X: x})
    for t in range(T):
        assert torch.allclose(out[t].float(), seq[:, t], atol=1e-5), t
    assert torch.allclose(out[T].float(), hn[0], atol=1e-5) and torch.allclose(out[T + 1].float(), cn[0], atol=1e-5)


@pytest.mark.parametrize("name", ["logreg", "mlp", "lenet", "cnn3", "alexnet", "vgg16", "resnet18", "rnn", "lstm"])
def test_image_and_sequence_classifiers_fit_a_small_batch(name):
    """every classifier of the zoo builds, back-propagates and drives the loss down on a memorisable batch"""
    from hetu_b200 import models as M
    rng = np.random.RandomState(0)
    B = 8
    if name in ("rnn", "lstm"):
        x, shape = rng.randn(B, 6, 10).astype(np.float32), [B, 6, 10]
        make = {"rnn": lambda: M.RNN(10, 32, 4), "lstm": lambda: M.LSTM(10, 32, 4)}[name]
    elif name in ("logreg", "lenet"):
        x, shape = rng.randn(B, 1, 28, 28).astype(np.float32), [B, 1, 28, 28]
        make = {"logreg": lambda: M.LogReg(784, 4), "lenet": lambda: M.LeNet(1, 4, 28)}[name]
    else:
        x, shape = rng.randn(B, 3, 32, 32).astype(np.float32), [B, 3, 32, 32]
        make = {"mlp": lambda: M.MLP(3072, (64,), 4), "cnn3": lambda: M.CNN3(3, 4, 32, width=8), "alexnet": lambda: M.AlexNet(3, 4, 32, dropout=0.0),
                "vgg16": lambda: M.VGG(16, 3, 4, 32, width_div=8), "resnet18": lambda: M.ResNet(18, 3, 4, width=8)}[name]
    y = rng.randint(0, 4, B)
    with ht.graph("define_and_run", create_new=True) as g:
        model = make()
        X = ht.placeholder("float32", shape, name="x")
        Y = ht.placeholder("int64", [B], name="y")
        loss, logits = model(X, Y)
        train = ht.AdamOptimizer(lr=3e-4 if name == "alexnet" else 3e-3).minimize(loss)      # no normalisation layers in AlexNet: smaller steps
        assert list(logits.shape) == [B, 4]
        losses = [float(g.run(loss, [loss, train], {X: torch.as_tensor(x), Y: torch.as_tensor(y)})[0]) for _ in range(60 if name == "alexnet" else 25)]
    assert np.isfinite(losses).all() and losses[-1] < 0.7 * losses[0], (name, losses[0], losses[-1])


def test_deep_crossing_and_ncf_train():
    from hetu_b200.models import NCF, DeepCrossing
    rng = np.random.RandomState(0)
    B, F, N = 64, 6, 500
    with ht.graph("define_and_run", create_new=True) as g:
        m = DeepCrossing(N, 8, num_fields=F, num_dense=4, num_units=2, unit_hidden=32)
        d, s, y = ht.placeholder("float32", [B, 4], name="d"), ht.placeholder("int64", [B, F], name="s"), ht.placeholder("float32", [B, 1], name="y")
        loss, _ = m(d, s, y)
        train = ht.AdamOptimizer(lr=1e-2).minimize(loss)
        ls = []
        for _ in range(40):
            dense, sparse = rng.randn(B, 4).astype(np.float32), rng.randint(0, N, (B, F))
            label = ((dense[:, :1] + (sparse[:, :1] % 2)) > 0.5).astype(np.float32)
            ls.append(float(g.run(loss, [loss, train], {d: torch.as_tensor(dense), s: torch.as_tensor(sparse), y: torch.as_tensor(label)})[0]))
    assert np.mean(ls[-5:]) < 0.85 * np.mean(ls[:5])
    with ht.graph("define_and_run", create_new=True) as g:
        m = NCF(50, 80, factors=4, mlp_layers=(16, 8, 4))
        u, i, y = ht.placeholder("int64", [B], name="u"), ht.placeholder("int64", [B], name="i"), ht.placeholder("float32", [B, 1], name="y")
        loss, logit = m(u, i, y)
        train = ht.AdamOptimizer(lr=2e-2).minimize(loss)
        like = (np.arange(50)[:, None] % 4) == (np.arange(80)[None, :] % 4)          # users like the items of their own group (low rank)
        ls = []
        for _ in range(150):
            uu, ii = rng.randint(0, 50, B), rng.randint(0, 80, B)
            ls.append(float(g.run(loss, [loss, train], {u: torch.as_tensor(uu), i: torch.as_tensor(ii), y: torch.as_tensor(like[uu, ii].astype(np.float32).reshape(B, 1))})[0]))
    assert np.mean(ls[-10:]) < 0.7 * np.mean(ls[:10])


def test_encoder_decoder_transformer_learns_to_reverse_sequences_and_decodes_greedily():
    """ref: hetu/v1/examples/nlp/hetu_transformer.py -- seq2seq Transformer with padding masks, causal decoder, cross attention and
    label smoothing learns the reversal task; greedy decoding reproduces the targets"""
    from hetu_b200.models import Transformer, TransformerConfig
    ht.set_seed(1234)                      # initialisers draw from the global seed: independent of the tests that ran before
    rng = np.random.RandomState(0)
    V, S, B = 12, 6, 32
    BOS, EOS, PAD = 1, 2, 0
    cfg = TransformerConfig(src_vocab_size=V, tgt_vocab_size=V, d_model=32, num_heads=4, d_ff=64, num_encoder_layers=2, num_decoder_layers=2,
                            max_len=16, dropout=0.0, label_smoothing=0.05)

    def batch():
        lens = rng.randint(3, S + 1, B)
        src = np.zeros((B, S), np.int64); tin = np.zeros((B, S + 1), np.int64); tout = np.zeros((B, S + 1), np.int64)
        for i, n in enumerate(lens):
            seq = rng.randint(3, V, n)
            src[i, :n] = seq
            tin[i, :n + 1] = np.concatenate([[BOS], seq[::-1]])
            tout[i, :n + 1] = np.concatenate([seq[::-1], [EOS]])
        return src, tin, tout
    with ht.graph("define_and_run", create_new=True) as g:
        m = Transformer(cfg)
        SRC = ht.placeholder("int64", [B, S], name="src")
        TIN = ht.placeholder("int64", [B, S + 1], name="tin")
        TOUT = ht.placeholder("int64", [B, S + 1], name="tout")
        SM = ht.placeholder("float32", [B, S], name="src_mask")
        TM = ht.placeholder("float32", [B, S + 1], name="tgt_mask")
        loss, logits = m(SRC, TIN, TOUT, src_mask=SM, tgt_mask=TM)
        train = ht.AdamOptimizer(lr=3e-3).minimize(loss)
        losses = []
        for step in range(260):
            src, tin, tout = batch()
            feed = {SRC: torch.as_tensor(src), TIN: torch.as_tensor(tin), TOUT: torch.as_tensor(tout), SM: torch.as_tensor((src != PAD).astype(np.float32)),
                    TM: torch.as_tensor((tin != PAD).astype(np.float32))}
            out = g.run(loss, [loss, logits, train], feed)
            losses.append(float(out[0]))
        pred = out[1].float().numpy().reshape(B, S + 1, V).argmax(-1)
        real = tout != PAD
        assert losses[-1] < 0.35 * losses[0], (losses[0], losses[-1])
        assert (pred[real] == tout[real]).mean() > 0.9
        # greedy decoding of full-length sources (no padding) reproduces the reversed sequence
        src = rng.randint(3, V, (4, S))
        dec = m.greedy_decode(g, src, max_len=S + 2, bos_id=BOS, eos_id=EOS)
        assert (dec[:, 1:S + 1] == src[:, ::-1]).mean() > 0.6, (dec, src)       # free-running decoding of the longest sequences: harder than teacher forcing


def test_generator_greedy_and_sampling_continue_a_learned_pattern():
    """a tiny GPT learns 'count upwards modulo 20'; greedy decoding continues the count from prompts of different lengths, nucleus /
    top-k sampling with a low temperature does too, eos stops a sequence, and the filters keep exactly the requested candidate sets"""
    from hetu_b200.models import Generator, GPTConfig, GPTLMHeadModel, generate_ds_parallel_config
    from hetu_b200.models.generation import _filter_logits
    V, S, B = 24, 16, 8
    cfg = GPTConfig(vocab_size=V, n_positions=S, n_embd=32, n_layer=2, n_head=2)
    dsc = [generate_ds_parallel_config(2, 1, 1, 1, 1, zero=False)]
    rng = np.random.RandomState(0)
    with ht.graph("define_and_run", create_new=True) as g:
        m = GPTLMHeadModel(cfg, dsc)
        X, P, Y = (ht.placeholder("int64", [B * S], name=n) for n in ("ids", "pos", "lab"))
        loss = m(X, P, Y, seq_len=S)
        train = ht.AdamOptimizer(lr=5e-3).minimize(loss)
        for _ in range(150):
            start = rng.randint(0, 20, B)
            seq = (start[:, None] + np.arange(S + 1)[None]) % 20 + 2
            g.run(loss, [loss, train], {X: torch.as_tensor(seq[:, :-1].reshape(-1)), P: torch.arange(S).repeat(B), Y: torch.as_tensor(seq[:, 1:].reshape(-1))})
        state = m.state_dict()
    gen = Generator(lambda: GPTLMHeadModel(cfg, dsc), batch=3, window=S)
    gen.load_state_dict(state)
    prompts = [[5, 6, 7], [12], [19, 20, 21, 2]]
    out = gen.generate(prompts, max_new_tokens=6)
    for p, o in zip(prompts, out):
        assert o[:len(p)] == p and len(o) == len(p) + 6
        want = [(p[-1] - 2 + k) % 20 + 2 for k in range(1, 7)]
        assert o[len(p):] == want, (p, o)
    sampled = gen.generate(prompts, max_new_tokens=4, temperature=0.2, top_k=3, top_p=0.9, seed=1)
    assert all(s[:len(p)] == p and len(s) == len(p) + 4 for p, s in zip(prompts, sampled))
    assert sum(s[len(p):] == o[len(p):len(p) + 4] for p, s, o in zip(prompts, sampled, out)) >= 2      # a sharp distribution: mostly the greedy path
    stopped = gen.generate([[5, 6, 7]], max_new_tokens=10, eos_id=10)
    assert stopped[0] == [5, 6, 7, 8, 9, 10]
    full = gen.generate([list(range(2, 2 + S - 2))], max_new_tokens=10)
    assert len(full[0]) == S                                                    # the window is the hard limit
    z = np.log(np.array([[0.5, 0.25, 0.15, 0.07, 0.03]]))
    assert np.isfinite(_filter_logits(z, 2, 1.0)).sum() == 2 and np.isfinite(_filter_logits(z, 0, 0.8)).sum() == 3
    assert np.isfinite(_filter_logits(z, 0, 1.0)).all()
