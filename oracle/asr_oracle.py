"""CPU oracle for the chinese-asr inference hot path.  TEST INFRASTRUCTURE ONLY.

This file is a CPU (torch fp32 / numpy) restatement of the reference algorithm for the path
named by BASELINE.json: log-mel features -> 4-layer residual BiLSTM encoder -> additive-attention
LSTM decoder with greedy / beam search -> optional second-pass n-gram LM rescoring.

  * It is the CHECKER: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
    `--impl reference` leg may import it.  The product (chinese_asr_b200/) never does; the
    product fails loudly when its CUDA library is missing.
  * Parity pin: tests/golden/*.npz hold outputs of the UNMODIFIED reference (imported from
    /root/reference behind tools/ref_shim.py) on seeded inputs; tests/test_oracle_golden.py checks
    this restatement against every one of them, plus the reference's single reproducible
    known-answer (encoder.py:636-652).  The n-gram LM is builder-defined because the reference's
    KenLM model/source is absent -> LM scoring itself is "parity unpinned" (the rescoring *rule*
    model.py:749-763 is pinned through the reference with this LM injected).

All arithmetic is float32 like the reference; indices are int64.  Citations are file:line into
the reference tree.
"""
from __future__ import annotations

import math
from collections import namedtuple

import numpy as np
import torch
import torch.nn.functional as F

# ----------------------------------------------------------------------------------------------
# constants of the path (gpd.py:4-133; captured at import time in the reference, so effectively
# architecture constants - SURVEY.md section 1)
SAMPLE_RATE = 16000
N_FFT = 512
HOP = 160            # gpd window_step .01 * 16000           (data.py:206)
WIN = 400            # gpd window_len .025 * 16000           (data.py:207)
N_BINS = 257
N_MELS = 80
F_MIN, F_MAX = 80.0, 7600.0                                   # data.py:378-379
PREEMPH = 0.97                                                # gpd.py:18
FEAT_DIM = 720       # 80 mel * 3 (delta stack) * 3 (frame stack)
ENC_H = 256          # per direction                          gpd.py:62
ENC_LAYERS = 4
ENC_OUT = 512
DEC_H = 512
EMB = 256
ATT = 128
VOCAB = 5004         # max_num_words 5000 + 4 specials        decoder.py:11-12
PAD, SOS, EOS, UNK = 0, 1, 2, 3
MAX_LEN = 40         # gpd.py:125
SPACE_ID = 781       # dict.pkl: the literal ' ' token

EvalOutput = namedtuple('EvalOutput', ('pred_text', 'score', 'text', 'wer', 'n', 'alignment',
                                       'audio_feat_len', 'text_len'))   # util.py:2403


# ----------------------------------------------------------------------------------------------
# features (data.py:21-57, 129-164, 167-280; main.py:36-37)

def mel_filterbank() -> torch.Tensor:
    """[257, 80] triangular filters, data.py:21-57.  Quirk kept: STFT bin j is given the
    frequency linspace(80, 7600, 257)[j] (data.py:43), HTK mel scale, un-normalised."""
    def hz2mel(f):
        return 2595.0 * torch.log10(torch.tensor(1.0) + (f / 700.0))

    def mel2hz(m):
        return 700.0 * (10 ** (m / 2595.0) - 1.0)

    bin_hz = torch.linspace(F_MIN, F_MAX, N_BINS)
    mel_pts = torch.linspace(hz2mel(F_MIN), hz2mel(F_MAX), N_MELS + 2)
    hz_pts = mel2hz(mel_pts)
    width = hz_pts[1:] - hz_pts[:-1]
    dist = hz_pts.unsqueeze(0) - bin_hz.unsqueeze(1)           # [257, 82]
    falling = (-1.0 * dist[:, :-2]) / width[:-1]
    rising = dist[:, 2:] / width[1:]
    return torch.max(torch.tensor(0.0), torch.min(falling, rising))


def hann_window() -> torch.Tensor:
    return torch.hann_window(WIN)                               # periodic, data.py:381-382


def delta_taps() -> np.ndarray:
    """[3, 9] taps: identity / delta / delta-delta, each L2-normalised (data.py:138-149)."""
    d = np.array([2, 1, 0, -1, -2], dtype=np.float64)
    dd = np.convolve(d, d, mode="full")
    taps = np.zeros((3, 9), dtype=np.float32)
    taps[0, 4] = 1.0
    taps[1, 2:7] = d
    taps[2] = dd
    taps /= np.sqrt((taps.astype(np.float32) ** 2).sum(axis=1, keepdims=True))
    return taps.astype(np.float32)


_FB = None
_WINDOW = None


def _consts():
    global _FB, _WINDOW
    if _FB is None:
        _FB, _WINDOW = mel_filterbank(), hann_window()
    return _FB, _WINDOW


def num_frames(n_samples: int) -> int:
    """STFT frame count T for a waveform of n_samples (after the 1-sample pre-emphasis loss)."""
    n = n_samples - 1
    return 0 if n < N_FFT else 1 + (n - N_FFT) // HOP


def log_mel_frames(pcm: np.ndarray) -> torch.Tensor:
    """pcm float32 [N] -> log-mel [T, 80]  (data.py:201-224)."""
    fb, window = _consts()
    x = np.asarray(pcm, dtype=np.float32)
    x = x[1:] - np.float32(PREEMPH) * x[:-1]                    # data.py:202 (float64 scalar * f32 -> f32)
    spec = torch.stft(torch.from_numpy(x).view(1, -1), n_fft=N_FFT, hop_length=HOP, win_length=WIN,
                      window=window, center=False, normalized=False, onesided=True,
                      return_complex=True)                      # [1, 257, T]
    spec = torch.view_as_real(spec).transpose(1, 2)             # [1, T, 257, 2]
    power = spec.pow(2).sum(-1)
    mel = torch.matmul(power, fb)                               # data.py:80
    mel.masked_fill_(mel == 0.0, torch.finfo(torch.float32).eps)
    return torch.log(mel[0])


def delta_stack(logmel: torch.Tensor) -> torch.Tensor:
    """[T, 80] -> [L = T // 3, 720]; column = c*240 + j*80 + m  (data.py:157-162, 244-249)."""
    taps = torch.from_numpy(delta_taps()).view(3, 1, 9, 1)
    x = F.pad(logmel[None, None], pad=(0, 0, 4, 4), mode="constant", value=0.0)
    y = F.conv2d(x, taps)[0]                                    # [3, T, 80]
    T = y.size(1)
    L = T // 3
    y = y[:, :3 * L].reshape(3, L, 240).transpose(0, 1).contiguous().view(L, FEAT_DIM)
    return y


def cmvn(feat: torch.Tensor, eps: float = 1e-6) -> torch.Tensor:
    """main.py:37 - per-utterance, per-column mean / unbiased std."""
    return (feat - feat.mean(dim=0)) / (feat.std(dim=0) + eps)


def features(pcm: np.ndarray, normalise: bool = True) -> torch.Tensor:
    """get_log_mel (data.py:167-280) + CMVN (main.py:37): pcm [N] -> [L, 720]."""
    f = delta_stack(log_mel_frames(pcm))
    return cmvn(f) if normalise else f


# ----------------------------------------------------------------------------------------------
# synthetic inputs and weights.  Defined HERE (not taken from the reference
# RNG stream) so the GPU box can regenerate them without /root/reference; the golden script loads
# the same dict into the reference through Model.load's checkpoint layout (model.py:357-369).

def synth_pcm(seed: int, n_samples: int) -> np.ndarray:
    return (0.1 * np.random.default_rng(seed).standard_normal(n_samples)).astype(np.float32)


def _xavier_normal(g, shape):
    fan_out, fan_in = shape[0], shape[1]
    std = math.sqrt(2.0 / (fan_in + fan_out))
    return (torch.randn(shape, generator=g) * std)


def _orthogonal(g, rows, cols):
    a = torch.randn(rows, cols, generator=g, dtype=torch.float64)
    q, r = torch.linalg.qr(a)
    q = q * torch.sign(torch.diagonal(r)).unsqueeze(0)
    return q.to(torch.float32).contiguous()


def _lstm_bias(n):
    b = torch.zeros(n)
    b[n // 4: n // 2] = 0.5                                     # util.py:101-107 forget-gate bias
    return b


def make_weights(seed: int = 1234, variant: str = "plain", proj_scale: float = 30.0,
                 emb_scale: float = 10.0, eos_bias: float = 8.0, v_scale: float = 10.0) -> dict:
    """Random-init weights with the reference's initialiser *distributions* (util.py:90-114,
    decoder.py:75-92, attention.py:53-65), in the Model.save/load layout (model.py:347-369):
    {'encoder_state_dict': 32 tensors, 'decoder_state_dict': 12 tensors}.

    variant 'sharp' (SURVEY.md section 8d): proj_linear.weight x proj_scale, embedding x emb_scale,
    attention v x v_scale and an EOS bias, so that the logit spread is a few nats (decision margins
    well above fp32 noise) and beams finish at varied steps."""
    g = torch.Generator().manual_seed(seed)
    enc = {}
    for layer in range(ENC_LAYERS):
        k_in = FEAT_DIM if layer == 0 else ENC_OUT
        for sfx in ("", "_reverse"):
            p = f"rnn.rnn.{layer}."
            enc[p + "weight_ih_l0" + sfx] = _xavier_normal(g, (4 * ENC_H, k_in))
            enc[p + "weight_hh_l0" + sfx] = _orthogonal(g, 4 * ENC_H, ENC_H)
            enc[p + "bias_ih_l0" + sfx] = _lstm_bias(4 * ENC_H)
            enc[p + "bias_hh_l0" + sfx] = _lstm_bias(4 * ENC_H)
    dec = {}
    emb = torch.randn(VOCAB, EMB, generator=g) * 0.1
    dec["embedding.weight"] = emb                                # decoder.py:77 (no pad zeroing after init)
    dec["cell.cell.0.weight_ih"] = _xavier_normal(g, (4 * DEC_H, EMB + ENC_OUT))
    dec["cell.cell.0.weight_hh"] = _orthogonal(g, 4 * DEC_H, DEC_H)
    dec["cell.cell.0.bias_ih"] = _lstm_bias(4 * DEC_H)
    dec["cell.cell.0.bias_hh"] = _lstm_bias(4 * DEC_H)
    dec["proj_linear.weight"] = _xavier_normal(g, (VOCAB, DEC_H + ENC_OUT))
    bound = 1.0 / math.sqrt(DEC_H + ENC_OUT)                    # nn.Linear default bias init
    dec["proj_linear.bias"] = (torch.rand(VOCAB, generator=g) * 2 - 1) * bound
    # attention params are stored [in, out] and used as x @ W (attention.py:29-32, 77, 92)
    dec["attn_mechanism.W_enc"] = _xavier_normal(g, (ENC_OUT, ATT))
    dec["attn_mechanism.b_attn"] = torch.zeros(ATT)
    dec["attn_mechanism.W_hidden"] = _xavier_normal(g, (DEC_H, ATT))
    dec["attn_mechanism.v"] = torch.randn(ATT, generator=g) * 0.1
    if variant == "sharp":
        dec["proj_linear.weight"] = dec["proj_linear.weight"] * proj_scale
        dec["embedding.weight"] = dec["embedding.weight"] * emb_scale
        dec["proj_linear.bias"] = dec["proj_linear.bias"].clone()
        dec["proj_linear.bias"][EOS] += eos_bias
        dec["attn_mechanism.v"] = dec["attn_mechanism.v"] * v_scale
    elif variant != "plain":
        raise ValueError(variant)
    return {"encoder_state_dict": enc, "decoder_state_dict": dec,
            "optimizer_state_dict": None, "args": None}


# ----------------------------------------------------------------------------------------------
# encoder (encoder.py:36-81, util.py:1223-1324)

def _lstm_layer_library(x_packed, w, layer):
    """One bidirectional nn.LSTM layer on a PackedSequence through the same ATen entry point the
    reference reaches (util.py:1259 -> nn.LSTM.forward)."""
    p = f"rnn.rnn.{layer}."
    flat = [w[p + n] for n in ("weight_ih_l0", "weight_hh_l0", "bias_ih_l0", "bias_hh_l0",
                               "weight_ih_l0_reverse", "weight_hh_l0_reverse",
                               "bias_ih_l0_reverse", "bias_hh_l0_reverse")]
    bsz = int(x_packed.batch_sizes[0])
    z = x_packed.data.new_zeros(2, bsz, ENC_H)
    out, h, c = torch._VF.lstm(x_packed.data, x_packed.batch_sizes, (z, z), flat, True, 1, 0.0,
                               False, True)
    return out, h, c


def lstm_cell_math(gates, c_prev):
    """gates [.., 4H] in i,f,g,o order -> (h, c)."""
    i, f, g, o = gates.chunk(4, dim=-1)
    c = torch.sigmoid(f) * c_prev + torch.sigmoid(i) * torch.tanh(g)
    h = torch.sigmoid(o) * torch.tanh(c)
    return h, c


def _lstm_layer_explicit(x_list, w, layer):
    """Explicit-loop restatement of one packed bidirectional layer (semantics of SURVEY 3.3):
    per sequence, forward runs 0..L-1, backward runs L-1..0, both from zero state."""
    p = f"rnn.rnn.{layer}."
    outs, hs, cs = [], [], []
    for x in x_list:
        L = x.size(0)
        y = x.new_zeros(L, ENC_OUT)
        hn, cn = [], []
        for d, sfx in enumerate(("", "_reverse")):
            wih, whh = w[p + "weight_ih_l0" + sfx], w[p + "weight_hh_l0" + sfx]
            b = w[p + "bias_ih_l0" + sfx] + w[p + "bias_hh_l0" + sfx]
            xg = x @ wih.t() + b
            h = x.new_zeros(ENC_H)
            c = x.new_zeros(ENC_H)
            order = range(L) if d == 0 else range(L - 1, -1, -1)
            for t in order:
                h, c = lstm_cell_math(xg[t] + whh @ h, c)
                y[t, d * ENC_H:(d + 1) * ENC_H] = h
            hn.append(h)
            cn.append(c)
        outs.append(y)
        hs.append(torch.cat(hn))
        cs.append(torch.cat(cn))
    return outs, torch.stack(hs), torch.stack(cs)


def encoder_forward(weights: dict, feats: list, lens: torch.Tensor, explicit: bool = False,
                    return_layers: bool = False):
    """list of B [L_i, 720] + lens[B] -> (out [Lmax, B, 512] zero padded, lens, (h, c) [B, 512]).
    encoder.py:36-81: sort by length, pack, 4 residual BiLSTM layers, unpack, un-sort; state is
    the LAST layer's final (h, c), laid out fwd|bwd."""
    w = weights["encoder_state_dict"]
    bsz = len(feats)
    if explicit:
        xs = [f for f in feats]
        layer_outs = []
        for layer in range(ENC_LAYERS):
            ys, h, c = _lstm_layer_explicit(xs, w, layer)
            xs = ys if layer == 0 else [a + b for a, b in zip(xs, ys)]   # util.py:1284-1291
            layer_outs.append(xs)
        out = torch.nn.utils.rnn.pad_sequence(xs)                # [Lmax, B, 512], zeros
        res = (out, lens, (h, c))
        return res + (layer_outs,) if return_layers else res
    order = lens.argsort(descending=True)
    inv = torch.empty(bsz, dtype=torch.long)
    inv[order] = torch.arange(bsz)
    packed = torch.nn.utils.rnn.pack_sequence([feats[i] for i in order])
    x = packed.data
    layer_outs = []
    for layer in range(ENC_LAYERS):
        y, h, c = _lstm_layer_library(
            torch.nn.utils.rnn.PackedSequence(x, packed.batch_sizes), w, layer)
        x = y if layer == 0 else x + y
        if return_layers:
            po, _ = torch.nn.utils.rnn.pad_packed_sequence(
                torch.nn.utils.rnn.PackedSequence(x, packed.batch_sizes))
            layer_outs.append(po[:, inv])
    out, _ = torch.nn.utils.rnn.pad_packed_sequence(
        torch.nn.utils.rnn.PackedSequence(x, packed.batch_sizes), padding_value=0.0)
    out = out[:, inv]
    h = h[:, inv].transpose(0, 1).contiguous().view(bsz, -1)     # encoder.py:67-72
    c = c[:, inv].transpose(0, 1).contiguous().view(bsz, -1)
    res = (out, lens, (h, c))
    return res + (layer_outs,) if return_layers else res


def softmax_mask(lens: torch.Tensor) -> torch.Tensor:
    """util.py:131-142 - additive mask [Lmax, B]: 0 on valid frames, -inf on padding."""
    lmax = int(lens.max())
    pad = torch.arange(lmax).expand(lens.size(0), -1) >= lens.view(-1, 1)
    m = pad.t().float()
    m.masked_fill_(m == 1.0, -np.inf)
    return m


def tile(t: torch.Tensor, k: int, batch_first: bool = False) -> torch.Tensor:
    """util.py:41-56 - repeat every utterance k times along the batch axis."""
    if batch_first:
        return t.unsqueeze(1).expand(-1, k, *t.shape[1:]).contiguous().view(-1, *t.shape[1:])
    return t.unsqueeze(2).expand(-1, -1, k, *t.shape[2:]).contiguous().view(
        t.size(0), t.size(1) * k, *t.shape[2:])


# ----------------------------------------------------------------------------------------------
# attention + decoder step (attention.py:67-95, decoder.py:94-137, util.py:1650-1661)

def attention_keys(weights, enc_out):
    d = weights["decoder_state_dict"]
    return torch.matmul(enc_out, d["attn_mechanism.W_enc"]) + d["attn_mechanism.b_attn"]


def attention(weights, values, mask, h, keys):
    d = weights["decoder_state_dict"]
    e = (torch.tanh(keys + torch.mm(h, d["attn_mechanism.W_hidden"])) * d["attn_mechanism.v"]).sum(dim=2)
    a = F.softmax(mask + e, dim=0)                               # [L, R]
    ctx = (a[..., None] * values).sum(dim=0)                     # [R, 512]
    return ctx, a


def decoder_step(weights, enc_out, mask, keys, token, h, c, ctx_prev):
    """One step (decoder.py:94-137, Bahdanau + input feeding): returns logit, ctx, align, h, c."""
    d = weights["decoder_state_dict"]
    x = F.embedding(token, d["embedding.weight"])
    if ctx_prev is None:
        ctx_prev = x.new_zeros(x.size(0), ENC_OUT)               # decoder.py:105-110
    x = torch.cat((x, ctx_prev), dim=1)
    gates = (x @ d["cell.cell.0.weight_ih"].t() + d["cell.cell.0.bias_ih"]
             + h @ d["cell.cell.0.weight_hh"].t() + d["cell.cell.0.bias_hh"])
    h, c = lstm_cell_math(gates, c)
    ctx, align = attention(weights, enc_out, mask, h, keys)
    logit = F.linear(torch.cat([h, ctx], -1), d["proj_linear.weight"], d["proj_linear.bias"])
    return logit, ctx, align, h, c


# ----------------------------------------------------------------------------------------------
# greedy (model.py:503-602)

@torch.no_grad()
def greedy_decode(weights, feats, lens, int2word=None, max_len=MAX_LEN, trace=None):
    bsz = len(feats)
    enc_out, enc_len, (h, c) = encoder_forward(weights, feats, lens)
    mask = softmax_mask(enc_len)
    keys = attention_keys(weights, enc_out)
    tokens = torch.full((bsz,), SOS, dtype=torch.long)
    ctx = None
    outputs, aligns = [], []
    finished = torch.zeros(bsz, dtype=torch.bool)
    final_lens = torch.zeros(bsz, dtype=torch.int32)
    accum = torch.zeros(bsz)
    for step in range(max_len):
        logit, ctx, align, h, c = decoder_step(weights, enc_out, mask, keys, tokens, h, c, ctx)
        if trace is not None:
            trace.setdefault("logit", []).append(logit.clone())
        aligns.append(align)
        logp = logit - torch.logsumexp(logit, dim=1).view(-1, 1)
        logp, tokens = logp.max(dim=1)
        outputs.append(tokens)
        now = tokens == EOS
        accum = accum + ((~finished) & now).float() * logp        # EOS log-prob counted once
        finished = finished | now
        final_lens = final_lens + (~finished).to(torch.int32)
        accum = accum + (~finished).float() * logp
        if bool(finished.all()):
            break
    out = torch.stack(outputs, dim=1)
    seqs = [s[:n] for s, n in zip(out.tolist(), final_lens.tolist())]
    scores = []
    for i, s in enumerate(seqs):
        if len(s) == 0:
            scores.append(0.0)
        else:
            scores.append(accum[i].item() / (final_lens[i].item() + finished[i].item()))
    text = None
    if int2word is not None:
        text = ["".join(int2word[t] for t in s) for s in seqs]
    return {"tokens": seqs, "score": scores, "pred_text": text, "text_len": final_lens,
            "alignment": aligns, "audio_feat_len": enc_len, "steps": len(outputs)}


# ----------------------------------------------------------------------------------------------
# builder-defined back-off n-gram LM with KenLM scoring semantics

class NGramLM:
    """Seeded random back-off trigram over token ids.  `.score(sentence, bos=True, eos=True)`
    mirrors kenlm.LanguageModel.score as called at model.py:755: the sentence is split on
    whitespace, each word is looked up (OOV -> <unk>), context starts at <s> when bos, </s> is
    appended when eos, and the return value is the total log10 probability.

    Rules stated explicitly (the reference's KenLM binary is absent, so these are ours):
      * words are the dict.pkl strings; id 781 (' ') vanishes under the whitespace split;
      * '<pad>', '<s>', '<unk>' appearing as hypothesis tokens are ordinary words 0/1/3;
      * accumulation is float32, left to right (like KenLM's `float total`).
    Tables: unigram logp[V] & backoff[V]; bigrams and trigrams in open-addressing hash tables
    keyed by packed ids; lookup is longest-match with back-off weights added, ARPA style."""

    def __init__(self, seed=7, vocab=VOCAB, n_bigrams=60000, n_trigrams=60000, word2int=None):
        rng = np.random.default_rng(seed)
        self.vocab = vocab
        self.word2int = word2int
        uni = rng.standard_normal(vocab).astype(np.float32) * 0.7 - 3.5
        self.uni_logp = uni
        self.uni_bo = (-rng.random(vocab).astype(np.float32) * 0.8)
        a = rng.integers(0, vocab, n_bigrams)
        b = rng.integers(0, vocab, n_bigrams)
        self.bi = {}
        for x, y in zip(a.tolist(), b.tolist()):
            self.bi[(x, y)] = (np.float32(-rng.random() * 3.0), np.float32(-rng.random() * 0.6))
        # dense block among the most frequent ids so that real hypotheses hit bigrams/trigrams
        hot = 48
        for x in range(hot):
            for y in range(hot):
                if rng.random() < 0.5:
                    self.bi[(x, y)] = (np.float32(-rng.random() * 3.0), np.float32(-rng.random() * 0.6))
        self.tri = {}
        keys = list(self.bi.keys())
        pick = rng.integers(0, len(keys), n_trigrams)
        c = rng.integers(0, vocab, n_trigrams)
        for i, z in zip(pick.tolist(), c.tolist()):
            x, y = keys[i]
            self.tri[(x, y, z)] = np.float32(-rng.random() * 2.5)

    # -- scoring on ids -------------------------------------------------------------------
    def _word_logp(self, ctx, w):
        """log10 p(w | ctx) with back-off; ctx is a tuple of at most 2 ids (oldest first)."""
        if len(ctx) == 2:
            t = self.tri.get((ctx[0], ctx[1], w))
            if t is not None:
                return t
            bo = self.bi.get((ctx[0], ctx[1]))
            bo = bo[1] if bo is not None else np.float32(0.0)
            return np.float32(bo + self._word_logp(ctx[1:], w))
        if len(ctx) == 1:
            b = self.bi.get((ctx[0], w))
            if b is not None:
                return b[0]
            return np.float32(self.uni_bo[ctx[0]] + self.uni_logp[w])
        return self.uni_logp[w]

    def score_ids(self, ids, bos=True, eos=True):
        ctx = (SOS,) if bos else ()
        total = np.float32(0.0)
        seq = [i for i in ids if i != SPACE_ID]
        if eos:
            seq = seq + [EOS]
        for w in seq:
            total = np.float32(total + self._word_logp(ctx, w))
            ctx = (ctx + (w,))[-2:]
        return float(total)

    def score(self, sentence, bos=True, eos=True):
        assert self.word2int is not None, "NGramLM needs word2int to score strings"
        ids = [self.word2int.get(wd, UNK) for wd in sentence.split()]
        return self.score_ids(ids, bos=bos, eos=eos)

    # -- flat tables for the device (uploaded by the product through the C ABI) -------------
    def tables(self):
        def build(d, key_fn, nvals):
            cap = 1
            while cap < 2 * len(d) + 8:
                cap *= 2
            keys = np.full(cap, -1, dtype=np.int64)
            vals = np.zeros((cap, nvals), dtype=np.float32)
            for k, v in d.items():
                kk = key_fn(k)
                slot = hash_slot(kk, cap)
                while keys[slot] != -1:
                    slot = (slot + 1) & (cap - 1)
                keys[slot] = kk
                vals[slot] = v
            return keys, vals
        bk, bv = build(self.bi, lambda k: k[0] * self.vocab + k[1], 2)
        tk, tv = build(self.tri, lambda k: (k[0] * self.vocab + k[1]) * self.vocab + k[2], 1)
        return {"uni_logp": self.uni_logp, "uni_bo": self.uni_bo, "bi_keys": bk, "bi_vals": bv,
                "tri_keys": tk, "tri_vals": tv.reshape(-1)}


def hash_slot(key: int, cap: int) -> int:
    """64-bit mix (splitmix64 finaliser) -> slot; identical on device (csrc/asr_lm.cuh)."""
    m = (1 << 64) - 1
    z = (key + 0x9E3779B97F4A7C15) & m
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & m
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & m
    z = z ^ (z >> 31)
    return int(z & (cap - 1))


# ----------------------------------------------------------------------------------------------
# beam search (model.py:604-987)

def _topk_sorted(scores, k):
    """torch.topk with a DEFINED tie order: value descending, then index ascending (the
    reference's torch.topk tie order is unspecified; seeded inputs have no exact ties among the
    selected candidates - the harness asserts that through `min_margin`)."""
    v, i = torch.sort(scores, dim=1, descending=True, stable=True)
    return v[:, :k].contiguous(), i[:, :k].contiguous()


@torch.no_grad()
def beam_decode(weights, k, feats, lens, int2word=None, second_pass=False, lm_model=None,
                lm_weight=0.0, length_weight=0.0, temperature=1.0, max_len=MAX_LEN, trace=None,
                batch_stop_step=None):
    """Restatement of Model.eval_one_batch_with_beam.  `trace` (dict) receives per-step
    internals: cand_scores/cand_beams/cand_tokens [B,2k], backptr/active_tokens [B,k],
    finished records, stop step, and per utterance the decision margins (`margin_utt`) and the
    first step whose rank-0 candidate was </s> (`top_done_step`, -1 = never).

    batch_stop_step: utterances interact only through the early stop (model.py:897-901: the loop
    ends at the first step where EVERY utterance of the batch has had </s> at rank 0).  To decode
    a few utterances of a large batch alone and still get what the reference returns for them
    inside that batch, the stop step of the whole batch is passed in: an int S ends the loop at
    step S, -1 never ends it early (some other utterance of the batch never finishes); None is
    the reference's own rule for the utterances given."""
    bsz = len(feats)
    R = bsz * k
    c2 = 2 * k
    enc_out, enc_len, (h, c) = encoder_forward(weights, feats, lens)
    mask = softmax_mask(enc_len)
    keys = attention_keys(weights, enc_out)
    # tile everything k times, including the whole encoder memory (model.py:660-669)
    mask, keys, enc_out = tile(mask, k), tile(keys, k), tile(enc_out, k)
    h, c = tile(h, k, True), tile(c, k, True)
    ctx = None
    hist = torch.full((max_len + 1, R), PAD, dtype=torch.long)
    hist[0] = SOS
    beam_scores = torch.zeros(R)
    base = k * torch.arange(bsz)
    rank_ids = torch.arange(c2).view(1, -1).expand(bsz, -1)
    top_done = torch.zeros(bsz, dtype=torch.bool)
    finished = []                  # (tokens [step, n], utt [n], score [n]) per step, model.py:889
    step = -1
    stopped_at = None
    for step in range(max_len):
        tokens = hist[step]
        logit, ctx, _align, h, c = decoder_step(weights, enc_out, mask, keys, tokens, h, c, ctx)
        logit = logit / temperature
        if trace is not None:
            trace.setdefault("logit", []).append(logit.clone())
        logp = logit - torch.logsumexp(logit, dim=1).view(-1, 1)
        logp = logp + beam_scores.view(-1, 1)
        flat = logp.view(bsz, -1)
        if step == 0:
            cand_s, cand_i = _topk_sorted(flat[:, :VOCAB], c2)   # all beams identical at step 0
            margin_src = flat[:, :VOCAB]
        else:
            cand_s, cand_i = _topk_sorted(flat, c2)
            margin_src = flat
        cand_beam = torch.div(cand_i, VOCAB, rounding_mode="floor")
        cand_tok = torch.fmod(cand_i, VOCAB)
        if trace is not None:
            trace.setdefault("cand_scores", []).append(cand_s.clone())
            trace.setdefault("cand_beams", []).append(cand_beam.clone())
            trace.setdefault("cand_tokens", []).append(cand_tok.clone())
            # decision margin: smallest gap between consecutive sorted candidates up to one past
            # the deepest candidate that is actually consumed (k-th non-EOS one)
            srt = torch.sort(margin_src, dim=1, descending=True)[0][:, :c2 + 1]
            non_eos_rank = torch.cumsum((cand_tok != EOS).long(), dim=1)
            deepest = (non_eos_rank < k).sum(dim=1).clamp(max=c2 - 1)       # index of k-th non-EOS
            gaps = srt[:, :-1] - srt[:, 1:]
            used = torch.arange(c2).view(1, -1) <= deepest.view(-1, 1)
            trace.setdefault("min_margin", []).append(float(gaps[used].min()))
            trace.setdefault("next_score", []).append(srt[:, c2].clone())     # best candidate left out of the 2k
            trace.setdefault("margin_utt", []).append(
                torch.where(used, gaps, torch.full_like(gaps, float("inf"))).min(dim=1)[0])
        # finished set: EOS among the top-k candidates (model.py:876-889)
        top_beam = cand_beam[:, :k] + base.view(-1, 1)
        is_eos_k = cand_tok[:, :k] == EOS
        fin_rows = top_beam.masked_select(is_eos_k)
        fin_tokens = hist[1:step + 1, fin_rows]
        fin_utt = torch.div(fin_rows, k, rounding_mode="floor")
        fin_scores = cand_s[:, :k].masked_select(is_eos_k)
        finished.append((fin_tokens, fin_utt, fin_scores))
        if trace is not None:
            first = trace.setdefault("top_done_step", torch.full((bsz,), -1, dtype=torch.long))
            first[(cand_tok[:, 0] == EOS) & ~top_done] = step
        top_done = top_done | (cand_tok[:, 0] == EOS)
        if bool(top_done.all()) if batch_stop_step is None else step == batch_stop_step:
            stopped_at = step
            break
        # active set: first k non-EOS candidates in rank order (model.py:904-909)
        order_key = rank_ids + (cand_tok == EOS).to(rank_ids.dtype) * c2
        _, active = torch.topk(order_key, k, largest=False)
        back = torch.gather(cand_beam, 1, active)
        rows = (back + base.view(-1, 1)).view(-1)
        new_tok = torch.gather(cand_tok, 1, active)
        if trace is not None:
            trace.setdefault("backptr", []).append(back.clone())
            trace.setdefault("active_tokens", []).append(new_tok.clone())
        # reorder by back-pointer (model.py:913-926) - the whole tiled memory moves
        enc_out, mask, keys = enc_out[:, rows], mask[:, rows], keys[:, rows]
        ctx, h, c = ctx[rows], h[rows], c[rows]
        hist = hist[:, rows]
        hist[step + 1] = new_tok.view(-1)
        beam_scores = torch.gather(cand_s, 1, active).view(-1)
    if trace is not None:
        trace["stopped_at"] = stopped_at
        trace["steps"] = step + 1
        trace["finished"] = finished

    # finalisation (model.py:708-765, 945-987)
    per_utt = {}
    for toks, utts, scs in finished:
        if utts.numel() == 0:
            continue
        toks, utts, scs = toks.t().tolist(), utts.tolist(), scs.tolist()
        for i, u in enumerate(utts):
            per_utt.setdefault(u, []).append((toks[i], scs[i]))
    nbest = {u: list(v) for u, v in per_utt.items()}
    chosen = {}
    for u, hyps in per_utt.items():
        if second_pass:
            if len(hyps) == 1:
                chosen[u] = hyps[0]
                continue
            lm = [lm_model.score(" ".join(int2word[t] for t in hy[0]), bos=True) for hy in hyps]
            total = [hy[1] + lm_weight * s + length_weight * len(hy[0]) for hy, s in zip(hyps, lm)]
            chosen[u] = hyps[int(np.argmax(total))]            # returned score stays un-rescored
        else:
            best = hyps[0]
            for hy in hyps[1:]:
                if hy[1] > best[1]:
                    best = hy                                    # first max wins
            chosen[u] = best
    missing = sorted(set(range(bsz)) - set(chosen))
    if missing:                                                  # model.py:961-972
        mt = torch.tensor(missing)
        act = beam_scores + length_weight * (step + 1)
        sc, idx = torch.topk(act.view(bsz, -1)[mt], k=1, dim=1)
        rows = idx.view(-1) + base[mt]
        toks = hist[1:step + 2, rows].t().tolist()
        for u, t, s in zip(missing, toks, sc.view(-1).tolist()):
            chosen[u] = (t, s)
    tokens = [chosen[u][0] for u in range(bsz)]
    scores = [chosen[u][1] for u in range(bsz)]
    text = None
    if int2word is not None:
        text = ["".join(int2word[t] for t in s) for s in tokens]
    return {"tokens": tokens, "score": scores, "pred_text": text, "nbest": nbest,
            "steps": step + 1, "stopped_at": stopped_at, "fallback": missing}


# ----------------------------------------------------------------------------------------------
# batched front end and error rate: the callers either side of
# the hot path.  Pinned by tests/golden/ref_frontend.npz (tools/make_golden_frontend.py).

def pcm_from_int16(x: np.ndarray) -> np.ndarray:
    """fast_read (data.py:109-121): soundfile.read(path, dtype='float32') on a 16-bit PCM file.
    soundfile is a third-party dependency absent from the reference tree and from this image
    (unpinned, no requirements file); libsndfile's documented int16 -> float conversion is
    x / 32768 (exact in float32)."""
    return np.asarray(x, dtype=np.int16).astype(np.float32) / np.float32(32768.0)


def synth_pcm_int16(seed: int, n_samples: int) -> np.ndarray:
    """16-bit synthetic waveform: the float waveform of synth_pcm quantised like a WAV writer does."""
    x = synth_pcm(seed, n_samples)
    return np.clip(np.round(x * 32768.0), -32768, 32767).astype(np.int16)


def convert_audio(x, rate: int, norm_db: float = -1.0) -> np.ndarray:
    """The device resampler's definition (chinese_asr_b200/csrc/resample.cu, include/asr_b200.h:
    asr_convert_audio) restated in float64 - NOT a restatement of reference code: main.py:19-24 runs ffmpeg
    (libswresample) and sox, neither of which is in the reference tree (parity unpinned).
    x: [n] or [n, channels], int16 or float in [-1, 1).  Down-mix = channel mean; y[m] = sum_i x[i] h(m rate /
    16000 - i) with h(d) = fc sinc(fc d) hann(d / W), fc = 0.97 min(1, 16000 / rate), W = 16 / fc; rate == 16000
    passes through; out = round(y 10^(dB / 20) / max |y| 32768) clipped to int16."""
    a = np.asarray(x)
    if a.dtype == np.int16:
        a = a.astype(np.float64) / 32768.0
    a = a.astype(np.float64)
    if a.ndim == 2:
        a = a.astype(np.float32).sum(axis=1, dtype=np.float32).astype(np.float64) / a.shape[1] if a.shape[1] > 1 else a[:, 0]
    n_in = a.shape[0]
    n_out = n_in * 16000 // rate
    if rate == 16000:
        y = a[:n_out].copy()
    else:
        fc = 0.97 * min(1.0, 16000.0 / rate)
        W = 16.0 / fc
        iw = int(W) + 1
        y = np.zeros(n_out)
        m = np.arange(n_out, dtype=np.int64)
        num = m * rate
        ip = num // 16000
        for off in range(-iw, iw + 2):
            i = ip + off
            ok = (i >= 0) & (i <= n_in - 1)
            d = (num - i * 16000) / 16000.0
            u = d / W
            ok &= np.abs(u) < 1.0
            h = fc * np.sinc(fc * d) * (0.5 + 0.5 * np.cos(np.pi * u))
            y += np.where(ok, a[np.clip(i, 0, n_in - 1)] * h, 0.0)
    peak = np.max(np.abs(y)) if n_out else 0.0
    scale = (10.0 ** (norm_db / 20.0)) / peak if peak > 0 else 1.0
    return np.clip(np.rint(y * scale * 32768.0), -32768, 32767).astype(np.int16)


def batch_audio(batch: list, eps: float = 1e-7):
    """AudioLoader.batch_audio (data.py:509-518) for the RNN encoders: instance normalisation of every
    [L, 720] feature matrix with eps 1e-7 (main.py:37 uses 1e-6), lens as IntTensor."""
    lens = torch.IntTensor([t.size(0) for t in batch])
    return [cmvn(t, eps) for t in batch], lens


def collate_fn(batch: list):
    """AudioLoader.collate_fn (data.py:496-507): [(feature, text_int)] or [(feature,)] -> (t, lens, text)."""
    t, lens = batch_audio([ele[0] for ele in batch])
    text = [ele[1] for ele in batch] if len(batch[0]) == 2 else None
    return t, lens, text


def edit_distance(pred, ref) -> int:
    """Unit-cost Levenshtein distance between two sequences: what get_wer (util.py:237-262) obtains
    from python-Levenshtein's distance() (third party, absent, unpinned) and what the reference's own
    get_wer_python (util.py:186-234) computes with one DP row."""
    n, m = len(pred), len(ref)
    if m == 0:
        return n
    if n == 0:
        return m
    dist = list(range(n + 1))
    for i in range(1, m + 1):
        pre = i
        for j in range(1, n + 1):
            if pred[j - 1] == ref[i - 1]:
                cur = dist[j - 1]
            else:
                cur = min(pre, dist[j], dist[j - 1]) + 1
            dist[j - 1] = pre
            pre = cur
        dist[n] = pre
    return dist[n]


def get_wer(pred: str, ref: str, normalize: bool = True):
    """util.py:237-246 (return_tuple=False): distance / len(ref) over the characters of the strings."""
    r = edit_distance(pred, ref)
    return r / (len(ref) * 1.) if normalize else r


def batch_wer(pred_tokens, ref_tokens, int2word):
    """The decode drivers' WER (model.py:595-598, 982-985): strings are ''.join(int2word[t]); mean of
    get_wer over the batch.  Returns (mean, per-utterance list)."""
    per = []
    for p, r in zip(pred_tokens, ref_tokens):
        per.append(get_wer("".join(int2word[t] for t in p), "".join(int2word[t] for t in r)))
    return float(np.mean(per)), per


def synth_reference_text(seed: int, hyp_tokens, n_vocab: int = VOCAB):
    """Seeded 'reference transcript' for WER tests: the hypothesis with random substitutions,
    deletions and insertions (token ids >= 3 so that '<unk>' - five characters - occurs)."""
    rng = np.random.default_rng(seed)
    out = []
    for t in hyp_tokens:
        r = rng.random()
        if r < 0.15:
            continue
        out.append(int(rng.integers(3, n_vocab)) if r < 0.35 else int(t))
        if rng.random() < 0.15:
            out.append(int(rng.integers(3, n_vocab)))
    if not out:
        out.append(int(rng.integers(3, n_vocab)))
    return out
