#!/usr/bin/env python
"""Entry points of the reference's main.py (27-65 parse, 68-102 ASR) on the B200 path.

Knobs kept: greedy (bw=None) vs beam (bw=4/8/16), lm_path -> second pass with lm_weight=1.5 /
length_weight=1.5 (main.py:45-51), gpd['temperature' | 'max_len' | 'verbose'], dict.pkl vocab.
convert_audio (main.py:19-24 shells out to ffmpeg + sox) runs on the device (data.convert_audio: down-mix,
windowed-sinc resampling to 16 kHz, peak normalisation to -1 dBFS; builder-defined, the two programs are not
part of the reference tree): `path` is an integer-PCM WAV of any rate / channel count - decoding compressed
containers, which ffmpeg also does for the reference, is not part of this package.  lm_path: an ARPA file (order <= 3) is loaded into
device tables and the second pass runs on the GPU; anything else is handed to kenlm.LanguageModel like
main.py:82 does (when the kenlm module is installed) and the finished hypotheses are rescored on the host
with its .score() (Model._host_rescore, from asr_beam_nbest)."""
from time import time

import torch

import numpy as np

from .data import AudioBase, _waveform, convert_audio
from .gpd import gpd
from .lm import NGramLM
from .model import Model


def parse(path, model, audio_base, lm_model, bw):
    """main.py:27-65.  `path`: WAV file (any rate / channels: converted like main.py:30) or a 16 kHz waveform array."""
    if isinstance(path, str):
        path = convert_audio(path)          # main.py:30 -> 'tmp.wav'
    pcm = _waveform(path)        # fast_read (data.py:109): 16-bit samples are converted on the device
    # get_log_mel + per-utterance CMVN (main.py:36-37), fused in the feature kernels
    data = model.features([pcm], normalise=True)[0]
    lens = torch.tensor([data.shape[0]])
    data = [data]
    if bw is not None:
        if gpd['verbose']:
            print(f"[INFO] Beam Decode [bw={bw}]...")
        res = model.eval_one_batch_with_beam(model.device, bw, data, lens, None, audio_base.int2word,
                                             second_pass=True if lm_model is not None else False,
                                             lm_model=lm_model, lm_weight=1.5, length_weight=1.5)
    else:
        res = model.eval_one_batch_with_greedy(model.device, data, lens, audio_base.int2word, None)
    return res.pred_text[0]


class ASR:
    def __init__(self, lm_path=None, bw=None, ckpt_path='./pretrain-0.06328.ckpt', dict_path=None):
        audio_base = AudioBase(dict_path)
        if lm_path is not None and bw is not None and bw > 1:
            print('loading language model...')
            ts = time()
            if str(lm_path).endswith('.arpa'):
                lm_model = NGramLM.from_arpa(lm_path, audio_base.word2int)
            else:
                import kenlm                      # main.py:82 (not part of this image; the caller's environment)
                lm_model = kenlm.LanguageModel(lm_path)
            print('loading cost %.3fs' % (time() - ts))
        else:
            lm_model = None
        model = Model()
        model.load(ckpt_path)
        model.model.eval()
        self.audio_base = audio_base
        self.lm_model = lm_model
        self.model = model
        self.bw = bw

    def __call__(self, path):
        return parse(path, self.model, self.audio_base, self.lm_model, self.bw)


if __name__ == '__main__':
    import sys
    gpd['verbose'] = False
    gpd['use_cuda'] = True
    gpd['temperature'] = 1
    lm_path = None
    bw = None
    asr = ASR(lm_path, bw)
    path = sys.argv[1]
    text = asr(path)
    print(f'ENV: lm_path={lm_path}, bw={bw}\nINPUT PATH : {path}\nOUTPUT TEXT: {text}')
