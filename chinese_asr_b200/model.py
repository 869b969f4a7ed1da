"""Model: the reference's decode drivers (model.py:18-82, 357-369, 503-602, 604-987) on B200.

Same entry points, argument meaning and result carriers as the reference:

    m = Model(); m.load(ckpt_path)                  # Model.save checkpoint layout, 44 tensors
    m.eval_one_batch_with_greedy(device, data, lens, int2word, text)
    m.eval_one_batch_with_beam(device, bmsz, data, lens, text, int2word,
                               second_pass, lm_model, lm_weight, length_weight)

Python only marshals buffers: features, encoder, the whole autoregressive loop, beam bookkeeping,
finalisation and LM rescoring run in libasr_b200.so (CUDA, sm_100a) without per-step host syncs.
torch tensors are buffer carriers (data_ptr()).  There is no CPU path: use_cuda=False raises."""
import ctypes as C

import numpy as np
import torch

from . import _cabi
from ._cabi import lib, check
from .gpd import gpd, check_frozen
from .util import EvalOutput
from . import data as _data

VOCAB = 5004
ENC_LAYERS = 4


def _f32(t):
    return np.ascontiguousarray(t.detach().cpu().numpy() if isinstance(t, torch.Tensor) else t, dtype=np.float32)


class Model(object):
    def __init__(self):
        check_frozen()
        if not gpd.get('use_cuda', True):
            raise RuntimeError("chinese_asr_b200 has no CPU path (gpd['use_cuda'] must be True)")
        if not torch.cuda.is_available():
            raise RuntimeError("chinese_asr_b200 needs a CUDA (sm_100a) device")
        self.device = torch.device('cuda', torch.cuda.current_device())
        self._h = None
        self._keep = []            # host arrays that must outlive asr_create
        self._reserved = None
        self._lm = None
        self._vocab_src = None
        self.optimizer = None
        self.logger = None

    # ---- weights -----------------------------------------------------------------------------
    def load(self, path):
        """model.py:357-369 - accepts exactly the Model.save dict."""
        if gpd['verbose']:
            print(f'[INFO] Loading weights from {path}...', end='')
        checkpoint = torch.load(path, map_location='cpu')
        self.load_state(checkpoint)
        if gpd['verbose']:
            print(' Loading done.')
        if 'args' in checkpoint:
            return checkpoint['args']

    def load_state(self, checkpoint):
        enc, dec = checkpoint['encoder_state_dict'], checkpoint['decoder_state_dict']
        w = _cabi.AsrWeights()
        keep = []

        def put(t, shape):
            a = _f32(t)
            if tuple(a.shape) != tuple(shape):
                raise ValueError(f"checkpoint tensor has shape {a.shape}, expected {shape}")
            keep.append(a)
            return _cabi.fptr(a)

        for layer in range(ENC_LAYERS):
            k_in = 720 if layer == 0 else 512
            for d, sfx in enumerate(('', '_reverse')):
                p = f'rnn.rnn.{layer}.'
                i = layer * 2 + d
                w.enc_w_ih[i] = put(enc[p + 'weight_ih_l0' + sfx], (1024, k_in))
                w.enc_w_hh[i] = put(enc[p + 'weight_hh_l0' + sfx], (1024, 256))
                w.enc_b_ih[i] = put(enc[p + 'bias_ih_l0' + sfx], (1024,))
                w.enc_b_hh[i] = put(enc[p + 'bias_hh_l0' + sfx], (1024,))
        w.embedding = put(dec['embedding.weight'], (VOCAB, 256))
        w.dec_w_ih = put(dec['cell.cell.0.weight_ih'], (2048, 768))
        w.dec_w_hh = put(dec['cell.cell.0.weight_hh'], (2048, 512))
        w.dec_b_ih = put(dec['cell.cell.0.bias_ih'], (2048,))
        w.dec_b_hh = put(dec['cell.cell.0.bias_hh'], (2048,))
        w.proj_w = put(dec['proj_linear.weight'], (VOCAB, 1024))
        w.proj_b = put(dec['proj_linear.bias'], (VOCAB,))
        w.att_w_enc = put(dec['attn_mechanism.W_enc'], (512, 128))
        w.att_b = put(dec['attn_mechanism.b_attn'], (128,))
        w.att_w_hidden = put(dec['attn_mechanism.W_hidden'], (512, 128))
        w.att_v = put(dec['attn_mechanism.v'], (128,))

        fc_np = _data.feature_consts()
        fc = _cabi.AsrFeatureConsts(_cabi.fptr(fc_np["mel_fb"]), _cabi.fptr(fc_np["window"]),
                                    _cabi.fptr(fc_np["taps"]), fc_np["preemphasis"])
        keep.append(fc_np)
        if self._h is not None:
            self.close()
        h = C.c_void_p()
        with torch.cuda.device(self.device):
            check(lib.asr_create(C.byref(h), C.byref(w), C.byref(fc)), "asr_create")
        self._h = h
        self._reserved = None
        self._lm = None                # the new handle has no LM tables / vocabulary yet
        self._vocab_src = None
        _data.set_default_engine(self)

    def close(self):
        if self._h is not None:
            lib.asr_destroy(self._h)
            self._h = None
        self._reserved = None
        self._lm = None
        self._vocab_src = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def model(self):
        return self          # main.py:90 calls model.model.eval()

    def eval(self):
        return self

    # ---- plumbing ----------------------------------------------------------------------------
    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _need(self):
        if self._h is None:
            raise RuntimeError("Model has no weights: call load() / load_state() first")

    def reserve(self, max_utts, max_rows, max_beam=16, max_samples=0, max_len=None):
        """Provision device workspaces (asr_reserve).  Called lazily with the sizes of the first
        batch; call explicitly to avoid re-allocation when batch sizes vary."""
        self._need()
        max_len = max_len or gpd['max_len']
        want = (int(max_utts), int(max_rows), int(max_beam), int(max_samples), int(max_len))
        cur = self._reserved
        if cur is not None and all(c >= w for c, w in zip(cur, want)):
            return
        if cur is not None:
            want = tuple(max(c, w) for c, w in zip(cur, want))
        check(lib.asr_reserve(self._h, *want), "asr_reserve")
        self._reserved = want

    # ---- features (data.py:167-280 + main.py:37) -------------------------------------------------
    def features(self, pcms, normalise=True, eps=1e-6):
        """list of waveforms (float32, or int16 as stored in 16-bit WAV files) -> list of CUDA tensors
        [L_i, 720].  eps: CMVN epsilon, 1e-6 = main.py:37, 1e-7 = AudioLoader.batch_audio (data.py:517)."""
        self._need()
        pcms = [p.detach().cpu().numpy() if isinstance(p, torch.Tensor) else np.asarray(p) for p in pcms]
        for p in pcms:
            if p.dtype != np.int16 and not np.issubdtype(p.dtype, np.floating):
                raise TypeError(f"waveform dtype {p.dtype}: only float (in [-1, 1)) or int16 PCM is accepted")
        s16 = len(pcms) > 0 and all(p.dtype == np.int16 for p in pcms)
        if s16:
            pcms = [np.ascontiguousarray(p) for p in pcms]
        else:       # a mixed batch: 16-bit items become x / 32768 like fast_read (data.py:111), exact in float32
            pcms = [p.astype(np.float32) / np.float32(32768.0) if p.dtype == np.int16
                    else np.ascontiguousarray(p, dtype=np.float32) for p in pcms]
        B = len(pcms)
        off = np.zeros(B + 1, dtype=np.int64)
        off[1:] = np.cumsum([len(p) for p in pcms])
        Ls = [int(lib.asr_num_frames(len(p))) for p in pcms]
        rows = int(sum(Ls))
        self.reserve(B, max(rows, 1), self._reserved[2] if self._reserved else 16, int(off[-1]))
        d_pcm = torch.from_numpy(np.concatenate(pcms)).to(self.device)
        d_feats = torch.empty(max(rows, 1), 720, dtype=torch.float32, device=self.device)
        h_L = np.zeros(B, dtype=np.int32)
        check(lib.asr_features_pcm(self._h, C.c_void_p(d_pcm.data_ptr()), _cabi.PCM_S16 if s16 else _cabi.PCM_F32,
                                   off.ctypes.data_as(_cabi.c_int64_p), B, C.c_void_p(d_feats.data_ptr()),
                                   h_L.ctypes.data_as(_cabi.c_int32_p), 1 if normalise else 0, float(eps),
                                   self._stream()), "asr_features_pcm")
        out, r = [], 0
        for n in h_L.tolist():
            out.append(d_feats[r:r + n])
            r += n
        return out

    def cmvn(self, feats, eps=1e-7):
        """AudioLoader.batch_audio's instance normalisation (data.py:513-518) on the device:
        list of [L_i, 720] tensors -> list of normalised CUDA tensors (views of one buffer)."""
        self._need()
        B = len(feats)
        lens = np.ascontiguousarray([int(t.size(0)) for t in feats], dtype=np.int32)
        for t in feats:
            if t.dim() != 2 or t.size(1) != 720:
                raise ValueError(f"feature tensor {tuple(t.shape)} is not [L, 720]")
        self.reserve(B, int(lens.sum()), self._reserved[2] if self._reserved else 1,
                     self._reserved[3] if self._reserved else 0)
        x = torch.cat([t.to(self.device, torch.float32) for t in feats], dim=0).contiguous()
        out = torch.empty_like(x)
        check(lib.asr_cmvn(self._h, C.c_void_p(x.data_ptr()), lens.ctypes.data_as(_cabi.c_int32_p), B, float(eps),
                           C.c_void_p(out.data_ptr()), self._stream()), "asr_cmvn")
        res, r = [], 0
        for n in lens.tolist():
            res.append(out[r:r + n])
            r += n
        return res

    # ---- character error rate (util.py:237-262) ----------------------------------------------------
    def set_vocab(self, int2word):
        """Upload int2word as code-point strings (asr_set_vocab) for the device edit distance."""
        self._need()
        if getattr(self, '_vocab_src', None) is int2word:
            return
        V = max(int2word) + 1
        off = np.zeros(V + 1, dtype=np.int32)
        cps = []
        for t in range(V):
            w = int2word.get(t, '')
            cps.extend(ord(ch) for ch in w)
            off[t + 1] = len(cps)
        cp = np.ascontiguousarray(cps if cps else [0], dtype=np.int32)
        check(lib.asr_set_vocab(self._h, cp.ctypes.data_as(_cabi.c_int32_p), off.ctypes.data_as(_cabi.c_int32_p), V),
              "asr_set_vocab")
        self._vocab_src = int2word

    def wer(self, text, int2word, hyp=None):
        """Per-utterance get_wer(pred, ref) (util.py:237-262, normalised by the reference length in
        characters) on the device.  text: reference transcripts as token-id lists (what collate_fn
        yields, data.py:503-504) or as strings; hyp: list of token-id lists, or None = the hypotheses
        of the last decode call, where they lie on the device."""
        self.set_vocab(int2word)
        B = len(text)
        refs = [t if isinstance(t, str) else ''.join(int2word[e] for e in t) for t in text]
        roff = np.zeros(B + 1, dtype=np.int64)
        roff[1:] = np.cumsum([len(t) for t in refs])
        ref = np.ascontiguousarray([ord(ch) for t in refs for ch in t] or [0], dtype=np.int32)
        dist = np.zeros(B, dtype=np.int32)
        hch = np.zeros(B, dtype=np.int32)
        i32 = _cabi.c_int32_p
        if hyp is None:
            hp, hl, ld = None, None, 0
        else:
            ld = max(1, max(len(x) for x in hyp))
            ha = np.zeros((B, ld), dtype=np.int32)
            hn = np.zeros(B, dtype=np.int32)
            for i, x in enumerate(hyp):
                ha[i, :len(x)] = x
                hn[i] = len(x)
            hp, hl = ha.ctypes.data_as(i32), hn.ctypes.data_as(i32)
        check(lib.asr_wer(self._h, hp, hl, ld, ref.ctypes.data_as(i32), roff.ctypes.data_as(_cabi.c_int64_p), B,
                          dist.ctypes.data_as(i32), hch.ctypes.data_as(i32), self._stream()), "asr_wer")
        # an empty reference divides by zero in the reference; here: 1.0 if anything was predicted
        return [d / len(r) if len(r) > 0 else float(n > 0) for d, r, n in zip(dist.tolist(), refs, hch.tolist())]

    # ---- encoder -----------------------------------------------------------------------------
    def _encode(self, data, lens, max_beam):
        self._need()
        if not isinstance(data, (list, tuple)):
            raise TypeError("data must be a list of [L_i, 720] tensors (RNN encoder path)")
        B = len(data)
        lens_np = np.ascontiguousarray(torch.as_tensor(lens).cpu().numpy(), dtype=np.int32)
        if lens_np.shape[0] != B:
            raise ValueError("lens and data disagree")
        for t, n in zip(data, lens_np.tolist()):
            if t.dim() != 2 or t.size(1) != 720 or t.size(0) != n:
                raise ValueError(f"feature tensor {tuple(t.shape)} does not match lens {n} x 720")
        rows = int(lens_np.sum())
        self.reserve(B, rows, max(max_beam, self._reserved[2] if self._reserved else 1),
                     self._reserved[3] if self._reserved else 0)
        feats = torch.cat([t.to(self.device, torch.float32) for t in data], dim=0).contiguous()
        check(lib.asr_encode(self._h, C.c_void_p(feats.data_ptr()), lens_np.ctypes.data_as(_cabi.c_int32_p), B,
                             self._stream()), "asr_encode")
        return B, lens_np

    def encode_export(self, data, lens):
        """(enc_out [Lmax,B,512], keys [Lmax,B,128], h [B,512], c [B,512]) in reference layouts."""
        B, lens_np = self._encode(data, lens, 1)
        lmax = int(lens_np.max())
        enc = torch.empty(lmax, B, 512, device=self.device)
        keys = torch.empty(lmax, B, 128, device=self.device)
        h = torch.empty(B, 512, device=self.device)
        c = torch.empty(B, 512, device=self.device)
        check(lib.asr_export_encoder(self._h, C.c_void_p(enc.data_ptr()), C.c_void_p(keys.data_ptr()),
                                     C.c_void_p(h.data_ptr()), C.c_void_p(c.data_ptr()), self._stream()),
              "asr_export_encoder")
        return enc, keys, h, c

    def encode_layer(self, data, lens, layer):
        """Residual-stream output of encoder layer `layer` as [Lmax, B, 512] (tests)."""
        self._need()
        B = len(data)
        lens_np = np.ascontiguousarray(torch.as_tensor(lens).cpu().numpy(), dtype=np.int32)
        self.reserve(B, int(lens_np.sum()), self._reserved[2] if self._reserved else 1,
                     self._reserved[3] if self._reserved else 0)
        feats = torch.cat([t.to(self.device, torch.float32) for t in data], dim=0).contiguous()
        out = torch.empty(int(lens_np.max()), B, 512, device=self.device)
        check(lib.asr_encode_layers(self._h, C.c_void_p(feats.data_ptr()), lens_np.ctypes.data_as(_cabi.c_int32_p),
                                    B, layer, C.c_void_p(out.data_ptr()), self._stream()), "asr_encode_layers")
        return out

    # ---- greedy (model.py:503-602) -------------------------------------------------------------
    @torch.no_grad()
    def eval_one_batch_with_greedy(self, device, data, lens, int2word=None, text=None, return_logits=False):
        B, lens_np = self._encode(data, lens, 1)
        max_len = gpd['max_len']
        lmax = int(lens_np.max())
        tokens = np.zeros((B, max_len), dtype=np.int32)
        tlen = np.zeros(B, dtype=np.int32)
        accum = np.zeros(B, dtype=np.float32)
        fin = np.zeros(B, dtype=np.int32)
        steps = C.c_int32(0)
        align = torch.zeros(max_len, lmax, B, dtype=torch.float32, device=self.device)
        logits = torch.empty(max_len, B, VOCAB, device=self.device) if return_logits else None
        check(lib.asr_decode_greedy(self._h, max_len, tokens.ctypes.data_as(_cabi.c_int32_p),
                                    tlen.ctypes.data_as(_cabi.c_int32_p), _cabi.fptr(accum),
                                    fin.ctypes.data_as(_cabi.c_int32_p), C.byref(steps),
                                    C.c_void_p(align.data_ptr()),
                                    C.c_void_p(logits.data_ptr()) if return_logits else None,
                                    self._stream()), "asr_decode_greedy")
        nsteps = steps.value
        outputs = [tokens[i, :tlen[i]].tolist() for i in range(B)]
        pred_text, score = [], []
        for i, ele in enumerate(outputs):
            if len(ele) == 0:
                pred_text.append('')
                score.append(.0)
            else:
                pred_text.append(''.join([int2word[e] for e in ele]) if int2word is not None else ele)
                score.append(float(accum[i]) / (int(tlen[i]) + int(fin[i])))      # model.py:593
        wer = None
        if text is not None:
            wer = np.mean(self.wer(text, int2word))        # hypotheses are scored where they lie (asr_wer)
            text = [t if isinstance(t, str) else ''.join([int2word[e] for e in t]) for t in text]
        out = EvalOutput(pred_text=pred_text, score=score, text=text, wer=wer, n=B,
                         alignment=[align[s] for s in range(nsteps)],
                         audio_feat_len=torch.as_tensor(lens), text_len=torch.from_numpy(tlen.copy()))
        if return_logits:
            return out, logits[:nsteps], tokens[:, :nsteps]
        return out

    # ---- LM ------------------------------------------------------------------------------------
    def set_lm(self, lm_model):
        """Upload the tables of an NGramLM (chinese_asr_b200.lm) - the device replacement of the
        kenlm object the reference passes as lm_model (model.py:755)."""
        self._need()
        if not hasattr(lm_model, 'tables'):
            raise TypeError("set_lm needs .tables() (chinese_asr_b200.lm.NGramLM); an lm_model that only offers "
                            ".score() is rescored on the host from asr_beam_nbest (see _host_rescore)")
        if self._lm is lm_model:
            return
        t = lm_model.tables()
        # typed, contiguous host copies that stay alive until asr_set_lm has copied them
        uni_logp = np.ascontiguousarray(t["uni_logp"], dtype=np.float32)
        uni_bo = np.ascontiguousarray(t["uni_bo"], dtype=np.float32)
        bi_keys = np.ascontiguousarray(t["bi_keys"], dtype=np.int64)
        bi_vals = np.ascontiguousarray(t["bi_vals"], dtype=np.float32).reshape(-1)
        tri_keys = np.ascontiguousarray(t["tri_keys"], dtype=np.int64)
        tri_vals = np.ascontiguousarray(t["tri_vals"], dtype=np.float32).reshape(-1)
        assert bi_vals.shape[0] == 2 * bi_keys.shape[0] and tri_vals.shape[0] == tri_keys.shape[0]
        id_map = np.ascontiguousarray(t["id_map"], dtype=np.int32) if t.get("id_map") is not None else None
        assert id_map is None or id_map.shape[0] == uni_logp.shape[0]
        tb = _cabi.AsrLmTables(
            _cabi.fptr(uni_logp), _cabi.fptr(uni_bo),
            bi_keys.ctypes.data_as(_cabi.c_int64_p), _cabi.fptr(bi_vals), int(bi_keys.shape[0]),
            tri_keys.ctypes.data_as(_cabi.c_int64_p), _cabi.fptr(tri_vals), int(tri_keys.shape[0]),
            int(uni_logp.shape[0]), int(getattr(lm_model, 'skip_id', 781)),
            id_map.ctypes.data_as(_cabi.c_int32_p) if id_map is not None else None)
        check(lib.asr_set_lm(self._h, C.byref(tb)), "asr_set_lm")
        self._lm = lm_model
        if hasattr(lm_model, 'bind'):
            lm_model.bind(self)

    def lm_score(self, id_lists):
        self._need()
        n = len(id_lists)
        max_n = max(1, max(len(x) for x in id_lists))
        ids = np.zeros((n, max_n), dtype=np.int32)
        ln = np.zeros(n, dtype=np.int32)
        for i, x in enumerate(id_lists):
            ids[i, :len(x)] = x
            ln[i] = len(x)
        out = np.zeros(n, dtype=np.float32)
        check(lib.asr_lm_score(self._h, ids.ctypes.data_as(_cabi.c_int32_p), ln.ctypes.data_as(_cabi.c_int32_p),
                               n, max_n, _cabi.fptr(out), self._stream()), "asr_lm_score")
        return out

    # ---- beam (model.py:604-987) -----------------------------------------------------------------
    @torch.no_grad()
    def eval_one_batch_with_beam(self, device, bmsz, data, lens, text, int2word,
                                 second_pass=gpd['second_pass'], lm_model=None,
                                 lm_weight=gpd['lm_weight'], length_weight=gpd['length_weight']):
        host_lm = self._prepare_lm(second_pass, lm_model)
        B, lens_np = self._encode(data, lens, bmsz)
        tokens, tlen, score, info = self._beam(B, bmsz, second_pass and not host_lm, lm_weight, length_weight)
        if host_lm:
            self._host_rescore(B, tokens, tlen, score, lm_model, lm_weight, length_weight, int2word)
        outputs = [tokens[i, :tlen[i]].tolist() for i in range(B)]
        pred_text = [''.join([int2word[idx] for idx in ele]) for ele in outputs]
        wer = None
        if text is not None:
            wer = np.mean(self.wer(text, int2word))        # hypotheses are scored where they lie (asr_wer)
            text = [t if isinstance(t, str) else ''.join([int2word[e] for e in t]) for t in text]
        self.last_beam_info = dict(steps=int(info[0]), stopped_at=int(info[1]), fallback=int(info[2]),
                                   finished=int(info[3]), tokens=outputs)
        return EvalOutput(pred_text=pred_text, score=[float(s) for s in score], text=text, wer=wer, n=B,
                          alignment=None, audio_feat_len=None, text_len=None)

    def _beam(self, B, bmsz, second_pass, lm_weight, length_weight):
        max_len = gpd['max_len']
        tokens = np.zeros((B, max_len), dtype=np.int32)
        tlen = np.zeros(B, dtype=np.int32)
        score = np.zeros(B, dtype=np.float32)
        info = np.zeros(4, dtype=np.int32)
        check(lib.asr_decode_beam(self._h, int(bmsz), max_len, float(gpd['temperature']), 1 if second_pass else 0,
                                  float(lm_weight), float(length_weight),
                                  tokens.ctypes.data_as(_cabi.c_int32_p), tlen.ctypes.data_as(_cabi.c_int32_p),
                                  _cabi.fptr(score), info.ctypes.data_as(_cabi.c_int32_p), self._stream()),
              "asr_decode_beam")
        return tokens, tlen, score, info

    def _prepare_lm(self, second_pass, lm_model):
        """Device tables when the LM offers them (NGramLM.tables()); any other object with the reference's
        duck-typed `.score(sentence, bos=True)` (model.py:755; main.py:82 passes a kenlm.LanguageModel) is
        rescored on the host from the device's n-best list.  Returns True for the host route."""
        if not second_pass:
            return False
        if lm_model is None:
            raise ValueError("second_pass=True needs lm_model")
        if hasattr(lm_model, 'tables'):
            self.set_lm(lm_model)
            return False
        if not hasattr(lm_model, 'score'):
            raise TypeError("lm_model needs .score(sentence, bos=True) (model.py:755)")
        return True

    def beam_nbest(self, B, cap=None):
        """parse_finished_tensors' per-utterance lists (model.py:708-747) of the last beam decode:
        [[(tokens, score), ...] in (step, rank) order] per utterance (asr_beam_nbest)."""
        self._need()
        i32 = _cabi.c_int32_p
        max_len = gpd['max_len']
        count = np.zeros(B, dtype=np.int32)
        if cap is None:
            check(lib.asr_beam_nbest(self._h, 0, max_len, count.ctypes.data_as(i32), None, None, None), "asr_beam_nbest")
            cap = max(1, int(count.max()))
        toks = np.zeros((B, cap, max_len), dtype=np.int32)
        lens = np.zeros((B, cap), dtype=np.int32)
        sc = np.zeros((B, cap), dtype=np.float32)
        check(lib.asr_beam_nbest(self._h, cap, max_len, count.ctypes.data_as(i32), toks.ctypes.data_as(i32),
                                 lens.ctypes.data_as(i32), _cabi.fptr(sc)), "asr_beam_nbest")
        return [[(toks[u, i, :lens[u, i]].tolist(), float(sc[u, i])) for i in range(min(int(count[u]), cap))]
                for u in range(B)]

    def _host_rescore(self, B, tokens, tlen, score, lm_model, lm_weight, length_weight, int2word):
        """model.py:749-763 with the caller's own lm_model: for every utterance with more than one finished
        hypothesis, argmax (first wins) of logp + lm_weight * lm_model.score(' '.join(chars), bos=True) +
        length_weight * len; the returned score stays the un-rescored logp (:762).  Utterances with one
        finished hypothesis or none (fallback, :961-972) keep what the device finaliser chose."""
        if int2word is None:
            raise ValueError("host LM rescoring needs int2word (the LM scores strings, model.py:755)")
        for u, hyps in enumerate(self.beam_nbest(B)):
            if len(hyps) < 2:
                continue
            total = [s + lm_weight * lm_model.score(' '.join(int2word[t] for t in toks), bos=True)
                     + length_weight * len(toks) for toks, s in hyps]
            best_toks, best_s = hyps[int(np.argmax(total))]
            tokens[u, :] = 0
            tokens[u, :len(best_toks)] = best_toks
            tlen[u] = len(best_toks)
            score[u] = best_s

    def check_guards(self):
        """asr_check_guards: raises if a kernel wrote past the end of any device buffer of this engine."""
        self._need()
        bad = int(lib.asr_check_guards(self._h))
        if bad != 0:
            msg = lib.asr_last_error()
            raise _cabi.AsrError(f"{bad} device buffer guard(s) overwritten: {msg.decode() if msg else ''}")

    def decode_info(self):
        """{steps, stopped_at (-1 = ran to max_len), fallback, finished} of the last decode (asr_decode_info)."""
        info = np.zeros(4, dtype=np.int32)
        check(lib.asr_decode_info(self._h, info.ctypes.data_as(_cabi.c_int32_p)), "asr_decode_info")
        return dict(steps=int(info[0]), stopped_at=int(info[1]), fallback=int(info[2]), finished=int(info[3]))

    def beam_trace(self, B, bmsz):
        """Per-step internals of the last beam decode (parity tests): dict of numpy arrays."""
        steps = self.decode_info()['steps']
        K = 2 * bmsz
        cs = np.zeros((steps, B, K), dtype=np.float32)
        cb = np.zeros((steps, B, K), dtype=np.int32)
        ct = np.zeros((steps, B, K), dtype=np.int32)
        bp = np.zeros((steps, B, bmsz), dtype=np.int32)
        at = np.zeros((steps, B, bmsz), dtype=np.int32)
        fs = np.zeros((steps, B, bmsz), dtype=np.float32)
        i32 = _cabi.c_int32_p
        check(lib.asr_beam_trace(self._h, _cabi.fptr(cs), cb.ctypes.data_as(i32), ct.ctypes.data_as(i32),
                                 bp.ctypes.data_as(i32), at.ctypes.data_as(i32), _cabi.fptr(fs)), "asr_beam_trace")
        return dict(cand_scores=cs, cand_beams=cb, cand_tokens=ct, backptr=bp, active_tokens=at, fin_scores=fs)

    # ---- dataset loop (model.py:1370-1439 test_model) ------------------------------------------------
    def test_model(self, loader, int2word, bw=None, second_pass=False, lm_model=None, lm_weight=0.0,
                   length_weight=0.0):
        """The reference's evaluation loop over an AudioLoader: per-batch WER weighted by batch size
        (model.py:1411-1430), plus the sentence error rate it prints at the end (:1432-1437)."""
        total_pred_text, total_text, wer_list = [], [], []
        eval_nb, eval_wer = 0, 0.
        for data, lens, text in getattr(loader, 'loader', loader):
            if bw is None:
                output = self.eval_one_batch_with_greedy(self.device, data, lens, int2word, text)
            else:
                output = self.eval_one_batch_with_beam(self.device, bw, data, lens, text, int2word,
                                                       second_pass=second_pass, lm_model=lm_model,
                                                       lm_weight=lm_weight, length_weight=length_weight)
            eval_nb += lens.size(0)
            eval_wer += output.wer * lens.size(0)
            wer_list.append(output.wer)
            total_pred_text.extend(output.pred_text)
            total_text.extend(output.text)
        eval_wer /= max(eval_nb, 1)
        wrong = sum(1 for a, b in zip(total_pred_text, total_text) if a != b)
        return dict(wer=float(eval_wer), n=eval_nb, wer_list=[float(w) for w in wer_list],
                    error_rate=wrong / max(eval_nb, 1), pred_text=total_pred_text, text=total_text)

    # ---- fused path: PCM in, hypotheses out ---------------------------------------------------------
    def transcribe(self, pcm, offsets, bw=None, second_pass=False, lm_model=None, lm_weight=0.0,
                   length_weight=0.0, int2word=None, resident=False, cmvn_eps=1e-6):
        """Whole path for a batch (parse() of main.py:27-65, batched): `pcm` is one float32 buffer
        (pinned host tensor / numpy array, or a CUDA tensor when resident=True) holding the
        concatenated waveforms, `offsets` [B+1] sample offsets.  An int16 buffer (16-bit WAV samples
        as stored) is converted on the device.  Returns (tokens, lens, scores) numpy arrays, plus texts
        when int2word is given."""
        self._need()
        off = np.ascontiguousarray(offsets, dtype=np.int64)
        B = off.shape[0] - 1
        host_lm = self._prepare_lm(second_pass, lm_model)
        max_len = gpd['max_len']
        rows = int(sum(int(lib.asr_num_frames(int(off[i + 1] - off[i]))) for i in range(B)))
        self.reserve(B, max(rows, 1), max(bw or 1, 1), int(off[-1] - off[0]))
        tokens = np.zeros((B, max_len), dtype=np.int32)
        tlen = np.zeros(B, dtype=np.int32)
        score = np.zeros(B, dtype=np.float32)
        fmt = self._pcm_format(pcm)
        if resident:
            fn, ptr = lib.asr_transcribe_device_pcm, pcm.data_ptr()
            assert pcm.is_cuda
        else:
            fn = lib.asr_transcribe_pcm
            ptr = pcm.data_ptr() if isinstance(pcm, torch.Tensor) else pcm.ctypes.data
        check(fn(self._h, C.c_void_p(ptr), fmt, float(cmvn_eps), off.ctypes.data_as(_cabi.c_int64_p), B,
                 int(bw or 0), max_len,
                 float(gpd['temperature']), 1 if (second_pass and not host_lm) else 0, float(lm_weight),
                 float(length_weight),
                 tokens.ctypes.data_as(_cabi.c_int32_p), tlen.ctypes.data_as(_cabi.c_int32_p),
                 _cabi.fptr(score), self._stream()), "asr_transcribe")
        if host_lm:
            self._host_rescore(B, tokens, tlen, score, lm_model, lm_weight, length_weight, int2word)
        if int2word is not None:
            texts = [''.join(int2word[t] for t in tokens[i, :tlen[i]].tolist()) for i in range(B)]
            return tokens, tlen, score, texts
        return tokens, tlen, score

    # ---- GEMM engine ---------------------------------------------------------------------------------
    def prefetch(self, pcm, offsets, bw=None):
        """Start the host->device copy of a batch (pinned host tensor / numpy array) so that it
        overlaps the decode of the previous batch; the next transcribe() call given the same buffer
        and offsets uses the staged copy.  At most two batches may be in flight.  `bw` sizes the
        workspace like the transcribe() call that follows (a later re-allocation would drop the copy)."""
        self._need()
        off = np.ascontiguousarray(offsets, dtype=np.int64)
        B = off.shape[0] - 1
        rows = int(sum(int(lib.asr_num_frames(int(off[i + 1] - off[i]))) for i in range(B)))
        self.reserve(B, max(rows, 1), max(bw or 1, 1), int(off[-1] - off[0]))
        ptr = pcm.data_ptr() if isinstance(pcm, torch.Tensor) else pcm.ctypes.data
        self._keep_prefetch = (pcm, off)           # the copy is asynchronous: keep the host buffer alive
        check(lib.asr_prefetch_pcm_fmt(self._h, C.c_void_p(ptr), self._pcm_format(pcm),
                                       off.ctypes.data_as(_cabi.c_int64_p), B), "asr_prefetch_pcm_fmt")

    @staticmethod
    def _pcm_format(pcm):
        dt = str(pcm.dtype).replace('torch.', '')
        if dt == 'float32':
            return _cabi.PCM_F32
        if dt == 'int16':
            return _cabi.PCM_S16
        raise TypeError(f"PCM buffer must be float32 or int16, not {dt}")

    def set_recurrence_chunks(self, chunks_per_direction):
        """asr_set_recurrence_chunks: 7 (default) = lowest latency for one batch; fewer = wider chunks on fewer SMs."""
        self._need()
        check(lib.asr_set_recurrence_chunks(self._h, int(chunks_per_direction)), "asr_set_recurrence_chunks")

    def test_gemm(self, A, W, bias):
        """C = A @ W.T + bias on the device through the split-precision tensor-core engine (tests)."""
        self._need()
        M, K = A.shape
        N = W.shape[0]
        Cout = torch.empty(M, N, dtype=torch.float32, device=self.device)
        check(lib.asr_test_gemm(self._h, C.c_void_p(A.data_ptr()), C.c_void_p(W.data_ptr()),
                                C.c_void_p(bias.data_ptr()), C.c_void_p(Cout.data_ptr()), M, N, K,
                                self._stream()), "asr_test_gemm")
        return Cout

    # ---- instrumentation ---------------------------------------------------------------------------
    def launch_count(self, reset=False):
        return int(lib.asr_launch_count(self._h, 1 if reset else 0))

    def stage_timing(self, enable=True, recurrence_timeline=False):
        check(lib.asr_stage_timing(self._h, (1 if enable else 0) | (2 if recurrence_timeline else 0)), "asr_stage_timing")

    def stage_times(self):
        """ms per pipeline stage (CUDA events on the launch stream) + kernel-level timers:
        `gemm_kernel_ms` = all GEMM-engine launches (nested in the stages), `operand_split_ms`,
        `gemm_gflop` = their algorithmic 2*M*N*K, `attention_kernel_ms` = the attention kernel alone."""
        ms = np.zeros(12, dtype=np.float32)
        check(lib.asr_stage_times(self._h, _cabi.fptr(ms), 12), "asr_stage_times")
        names = ("features", "enc_input_gemm", "enc_recurrence", "attn_keys", "dec_cell", "attention",
                 "vocab_proj", "topk_bookkeep", "gemm_kernel_ms", "operand_split_ms", "gemm_gflop",
                 "attention_kernel_ms")
        return dict(zip(names, ms.tolist()))
