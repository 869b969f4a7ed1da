"""Result carriers of the decode drivers (reference util.py:2401-2410)."""
from collections import namedtuple

EvalOutput = namedtuple('EvalOutput', ('pred_text', 'score', 'text', 'wer', 'n', 'alignment',
                                       'audio_feat_len', 'text_len'))
EncoderOutput = namedtuple('EncoderOutput', ('out', 'out_lens', 'state'))
