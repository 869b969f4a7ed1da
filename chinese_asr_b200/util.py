"""Result carriers and small host helpers (reference util.py:2401-2410, 186-234)."""
from collections import namedtuple

EvalOutput = namedtuple('EvalOutput', ('pred_text', 'score', 'text', 'wer', 'n', 'alignment',
                                       'audio_feat_len', 'text_len'))
EncoderOutput = namedtuple('EncoderOutput', ('out', 'out_lens', 'state'))


def get_wer(pred, ref):
    """Character error rate by edit distance (reference util.py:237-262 uses python-Levenshtein;
    this is the dynamic-programming form of util.py:186-234).  Only runs when `text` is given."""
    n, m = len(pred), len(ref)
    if m == 0:
        return float(n > 0)
    prev = list(range(m + 1))
    for i in range(1, n + 1):
        cur = [i] + [0] * m
        for j in range(1, m + 1):
            cur[j] = min(prev[j] + 1, cur[j - 1] + 1, prev[j - 1] + (pred[i - 1] != ref[j - 1]))
        prev = cur
    return prev[m] / m
